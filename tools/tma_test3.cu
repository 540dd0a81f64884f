// 1-D bulk copy probe (cp.async.bulk, no tensor map) + tensor map via directly linked libcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void kb(const double *src, double *out, int n) {
  __shared__ __align__(128) double s[512];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(n * 8) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(s)), "l"(src), "r"(n * 8), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = s[i];
}
__global__ void kt(const __grid_constant__ CUtensorMap m, double *out, int c0, int c1) {
  __shared__ __align__(128) double s[64];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(64 * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(s)), "l"(&m), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < 64; i += blockDim.x) out[i] = s[i];
}
int main() {
  const int W = 74, H = 74;
  double *d, *out, *h = (double *)malloc(W * H * 8);
  for (int i = 0; i < W * H; i++) h[i] = i;
  cudaMalloc(&d, W * H * 8); cudaMalloc(&out, 4096 * 8);
  cudaMemcpy(d, h, W * H * 8, cudaMemcpyHostToDevice);
  kb<<<1, 128>>>(d + 74 * 3 + 4, out, 66);
  cudaError_t e = cudaDeviceSynchronize();
  printf("bulk copy kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) { double o[2]; cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost); printf("out = %g %g (expect %d %d)\n", o[0], o[1], 74 * 3 + 4, 74 * 3 + 5); }
  else return 1;
  alignas(64) CUtensorMap m;
  cuuint64_t dims[2] = {W, H}, strides[1] = {W * 8};
  cuuint32_t box[2] = {16, 4}, es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode (linked libcuda) rc=%d\n", (int)r);
  const unsigned long long *w = (const unsigned long long *)&m;
  for (int i = 0; i < 16; i++) printf("%016llx%c", w[i], i % 4 == 3 ? '\n' : ' ');
  kt<<<1, 128>>>(m, out, 4, 27);
  e = cudaDeviceSynchronize();
  printf("tensor kernel: %s\n", cudaGetErrorString(e));
  return 0;
}
