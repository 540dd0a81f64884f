// Performance build of the push kernels: FMA contraction allowed (-fmad=true).
#define EPB_NS epb_fast
#define EPB_FAST_MATH 1
#include "push.cuh"
void epb_launch_push_fast(const PushParams &P, int nd, bool tiled, cudaStream_t s, long long *launches) {
  epb_fast::launch_push(P, nd, tiled, s, launches);
}
void epb_launch_push_m_fast(const PushParams &P, cudaStream_t s, long long *launches) {
  epb_fast::launch_push_m(P, s, launches);
}
