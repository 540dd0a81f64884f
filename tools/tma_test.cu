// standalone TMA probe: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_test tma_test.cu && ./tma_test BW BH
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap m, const CUtensorMap *gm, double *out, int bw, int bh, int c0, int c1) {
  extern __shared__ __align__(128) double s[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bw * bh * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(s)), "l"(gm ? gm : &m), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = s[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
  int bw = argc > 1 ? atoi(argv[1]) : 66, bh = argc > 2 ? atoi(argv[2]) : 9;
  int dt = argc > 3 ? atoi(argv[3]) : 0, loc = argc > 4 ? atoi(argv[4]) : 0;
  const int W = 74, H = 74;
  double *d, *out, *h = (double *)malloc(W * H * 8);
  for (int i = 0; i < W * H; i++) h[i] = i;
  cudaMalloc(&d, W * H * 8); cudaMalloc(&out, 256 * 256 * 8);
  cudaMemcpy(d, h, W * H * 8, cudaMemcpyHostToDevice);
  void *fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap m;
  const int mul = dt == 2 ? 2 : 1;
  cuuint64_t dims[2] = {(cuuint64_t)W * mul, H}, strides[1] = {W * 8};
  cuuint32_t box[2] = {(cuuint32_t)bw * mul, (cuuint32_t)bh}, es[2] = {1, 1};
  CUtensorMapDataType dty = dt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : dt == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = ((EncodeFn)fn)(&m, dty, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d box %dx%d dtype %d loc %d\n", (int)r, bw, bh, dt, loc);
  CUtensorMap *gm = nullptr;
  if (loc) { cudaMalloc(&gm, sizeof m); cudaMemcpy(gm, &m, sizeof m, cudaMemcpyHostToDevice); }
  k<<<1, 128, bw * bh * 8>>>(m, gm, out, bw, bh, 3 * mul, 27);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    double o[4]; cudaMemcpy(o, out, 32, cudaMemcpyDeviceToHost);
    printf("out[0..1] = %g %g (expect %d %d)\n", o[0], o[1], 27 * W + 3, 27 * W + 4);
  }
  return 0;
}
