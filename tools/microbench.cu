// microbench.cu — secondary ceilings of the push/deposit kernel on B200, which BASELINE.md asks
// to be measured because they bind before HBM for FP64 PIC: FP64 FMA rate, shared-memory
// load wavefront cost under the kernel's access patterns (same-address FP64 broadcast,
// two-address, 128-bit), warp shuffle rate, shared FP64 atomicAdd (CAS loop) rate and
// global FP64 reduction rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITER = 4096;

__global__ void k_dfma(double *out, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < ITER; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// mode 0: all lanes one 8-byte address; 1: two addresses (lanes 0-15 / 16-31); 2: lane-consecutive 8 B;
// 3: all lanes one 16-byte address (LDS.128); 4: 27 lanes, row pitch 33 doubles (the reduction's read);
// 5: four addresses (8 lanes each)
__global__ void k_lds(double *out, int mode, int stride) {
  extern __shared__ double s[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int base;
  if (mode == 0 || mode == 3) base = 0;
  else if (mode == 1) base = (lane >> 4) * 70;
  else if (mode == 2) base = lane;
  else if (mode == 4) base = (lane < 27 ? lane : 0) * 33;
  else base = (lane >> 3) * 70;
  double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  int o = base;
  for (int i = 0; i < ITER; i++) {
    if (mode == 3) {
      const double2 v0 = *reinterpret_cast<const double2 *>(&s[(o) & 4094]);
      const double2 v1 = *reinterpret_cast<const double2 *>(&s[(o + 2 * stride) & 4094]);
      acc0 += v0.x; acc1 += v0.y; acc2 += v1.x; acc3 += v1.y;
    } else {
      acc0 += s[(o) & 4095]; acc1 += s[(o + stride) & 4095];
      acc2 += s[(o + 2 * stride) & 4095]; acc3 += s[(o + 3 * stride) & 4095];
    }
    o += 4 * stride;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}

__global__ void k_shfl(double *out) {
  double x = threadIdx.x, y = x + 1;
  for (int i = 0; i < ITER; i++) {
    x += __shfl_xor_sync(0xffffffffu, y, 1);
    y += __shfl_xor_sync(0xffffffffu, x, 2);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
}

// shared FP64 atomicAdd: mode 0 each lane its own address (27 lanes like the flush), 1 all warps same addresses
__global__ void k_atoms(double *out, int mode) {
  __shared__ double s[32 * 32];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *p = s + (mode == 0 ? warp * 32 + lane : lane);
  for (int i = 0; i < ITER; i++) atomicAdd(p, 1.0);
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s[threadIdx.x];
}

// global FP64 reduction (RED.E.ADD.F64), addresses spread over `span` doubles
__global__ void k_red(double *buf, size_t span) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long h = i * 0x9E3779B97F4A7C15ull;
  for (int k = 0; k < 64; k++) {
    h = h * 6364136223846793005ull + 1442695040888963407ull;
    atomicAdd(buf + (h >> 20) % span, 1.0);
  }
}

template <typename F>
static float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, 0));
  const int sms = pr.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_ghz_max\": %.3f,\n", pr.name, sms, ghz);
  double *out;
  CK(cudaMalloc(&out, (size_t)sms * 8 * 1024 * sizeof(double)));
  {
    const int blocks = sms * 4, thr = 512;
    float ms = time_ms([&] { k_dfma<<<blocks, thr>>>(out, 1.0000001, 1e-9); });
    double fma = (double)blocks * thr * ITER * 8;
    printf(" \"fp64_fma_tflops\": %.2f, \"fp64_fma_per_clk_per_sm\": %.1f,\n", 2 * fma / ms * 1e-9,
           fma / (ms * 1e-3) / sms / (ghz * 1e9));
  }
  const char *names[] = {"lds64_same_addr", "lds64_two_addr", "lds64_lane_consecutive", "lds128_same_addr",
                         "lds64_27lanes_pitch33", "lds64_four_addr"};
  for (int mode = 0; mode < 6; mode++) {
    const int blocks = sms * 2, thr = 512;
    CK(cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    float ms = time_ms([&] { k_lds<<<blocks, thr, 65536>>>(out, mode, mode == 4 ? 1 : 3); });
    double warp_loads = (double)blocks * (thr / 32) * ITER * (mode == 3 ? 2 : 4);
    printf(" \"%s_cycles_per_warp_load_per_sm\": %.2f,\n", names[mode], (ms * 1e-3) * ghz * 1e9 * sms / warp_loads);
  }
  {
    const int blocks = sms * 2, thr = 512;
    float ms = time_ms([&] { k_shfl<<<blocks, thr>>>(out); });
    double n = (double)blocks * (thr / 32) * ITER * 2;  // 64-bit shuffles (2 SHFL.32 each)
    printf(" \"shfl64_cycles_per_warp_op_per_sm\": %.2f,\n", (ms * 1e-3) * ghz * 1e9 * sms / n);
  }
  for (int mode = 0; mode < 2; mode++) {
    const int blocks = sms * 2, thr = 256;
    float ms = time_ms([&] { k_atoms<<<blocks, thr>>>(out, mode); });
    double n = (double)blocks * (thr / 32) * ITER;
    printf(" \"smem_f64_atomicadd_%s_cycles_per_warp_op_per_sm\": %.2f,\n", mode ? "all_warps_same_addr" : "private_addr",
           (ms * 1e-3) * ghz * 1e9 * sms / n);
  }
  {
    double *buf;
    const size_t span = (size_t)3 * 4106 * 4106;  // the three J arrays of C2
    CK(cudaMalloc(&buf, span * sizeof(double)));
    CK(cudaMemset(buf, 0, span * sizeof(double)));
    const int blocks = sms * 64, thr = 256;
    float ms = time_ms([&] { k_red<<<blocks, thr>>>(buf, span); });
    printf(" \"global_f64_red_random_gops\": %.2f\n", (double)blocks * thr * 64 / ms * 1e-6);
    cudaFree(buf);
  }
  printf("}\n");
  cudaFree(out);
  return 0;
}
