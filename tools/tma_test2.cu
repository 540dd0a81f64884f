// canonical libcu++ TMA example (CUDA programming guide) as an environment check
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int BW = 16, BH = 4;
__global__ void k(const __grid_constant__ CUtensorMap m, double *out, int c0, int c1) {
  __shared__ alignas(128) double s[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&s, &m, c0, c1, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(s));
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = (&s[0][0])[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int W = 74, H = 74;
  double *d, *out, *h = (double *)malloc(W * H * 8);
  for (int i = 0; i < W * H; i++) h[i] = i;
  cudaMalloc(&d, W * H * 8); cudaMalloc(&out, 4096 * 8);
  cudaMemcpy(d, h, W * H * 8, cudaMemcpyHostToDevice);
  void *fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap m;
  cuuint64_t dims[2] = {W, H}, strides[1] = {W * 8};
  cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d qres=%d\n", (int)r, (int)q);
  k<<<1, 128>>>(m, out, 3, 27);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) { double o[2]; cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost); printf("out = %g %g (expect %d %d)\n", o[0], o[1], 27 * W + 3, 27 * W + 4); }
  return 0;
}
