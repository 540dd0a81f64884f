// Parity build of the push kernels: compiled with -fmad=false so every multiply and
// add rounds separately, as in the reference's gfortran -O3 build (epoch2d/Makefile:72).
#define EPB_NS epb_strict
#include "push.cuh"
void epb_launch_push_strict(const PushParams &P, int nd, bool tiled, cudaStream_t s, long long *launches) {
  epb_strict::launch_push(P, nd, tiled, s, launches);
}
void epb_launch_push_m_strict(const PushParams &P, cudaStream_t s, long long *launches) {
  epb_strict::launch_push_m(P, s, launches);
}
