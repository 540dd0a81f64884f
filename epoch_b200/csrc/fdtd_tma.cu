// fdtd_tma.cu — Yee FDTD half-step updates with TMA-staged stencil tiles (2D and 3D).
//
// update_e_field / update_b_field (fields.f90:206-225 / :422-439; 3D epoch3d/src/fields.f90:312-337 /
// :632-654), order 2, same expressions and operation order as the plain kernels in epb_api.cu
// (compiled -fmad=false), so results are bit-identical.  Only the arrays that are read through a
// stencil are staged: the E update reads B at (i, i-1), the B update reads E at (i, i+1).  One
// elected thread arms an mbarrier with the expected byte count and issues one
// cp.async.bulk.tensor load per array (the box includes the one-cell stencil halo; out-of-range
// coordinates are zero-filled by the TMA unit, and never used); everybody waits on the barrier
// and computes from shared memory.  The pointwise operands (the updated field itself and J) are
// read straight from global memory, coalesced.
//
// Tensor maps need 16-byte-multiple global strides, i.e. an even padded extent nx + 2 ng; other
// shapes and 1D use the plain grid-stride kernels.
#include <cuda.h>

#include <cstdlib>

#include "epb_internal.h"

namespace {

constexpr int NG = EPB_NG;
// tile of updated points and the staged box (one extra cell per stencil direction; the x extent
// is rounded up to an even number of doubles: TMA wants box rows that are multiples of 16 bytes)
constexpr int BX2 = 64, BY2 = 8;              // 2D: 512 points per CTA
constexpr int BX3 = 32, BY3 = 4, BZ3 = 4;     // 3D: 512 points per CTA
constexpr int THREADS = 256;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2,
                                            unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

struct TmaFieldParams {
  double *f[9];
  int n[3];
  int sz[3];
  double cx, cy, cz, fac;
};

// ---- 2D -------------------------------------------------------------------------------------
constexpr int SW2 = BX2 + 2, SH2 = BY2 + 1;  // staged box (doubles): 66 x 9
template <bool IS_E>
__global__ void __launch_bounds__(THREADS) k_fdtd_tma_2d(const __grid_constant__ TmaFieldParams F,
                                                        const __grid_constant__ CUtensorMap m0,
                                                        const __grid_constant__ CUtensorMap m1,
                                                        const __grid_constant__ CUtensorMap m2) {
  __shared__ __align__(128) double s[3][(SH2 * SW2 + 15) / 16 * 16];  // each array 128-byte aligned
  __shared__ __align__(8) unsigned long long bar;
  // first updated point of this tile (Fortran indices start at 0: the first ghost layer is updated too)
  const int ix0 = blockIdx.x * BX2, iy0 = blockIdx.y * BY2;
  // array coordinates (0-based, ghosts included) of the box origin: E reads (i-1..i), B reads (i..i+1).
  // The innermost coordinate must be even: the TMA unit raises an illegal-instruction fault when a
  // box row of 8-byte elements does not start on a 16-byte boundary.  XO = offset of point lx = 0.
  constexpr int XO = IS_E ? 2 : 0;
  const int bx0 = ix0 + NG - 1 - XO, by0 = iy0 + NG - 1 - (IS_E ? 1 : 0);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 3u * SH2 * SW2 * sizeof(double));
    tma_load_2d(s[0], &m0, bx0, by0, &bar);
    tma_load_2d(s[1], &m1, bx0, by0, &bar);
    tma_load_2d(s[2], &m2, bx0, by0, &bar);
  }
  mbar_wait(&bar, 0);
  const size_t sy = F.sz[0];
#pragma unroll
  for (int r = 0; r < (BX2 * BY2) / THREADS; r++) {
    const int t = threadIdx.x + r * THREADS;
    const int lx = t % BX2, ly = t / BX2;
    const int ix = ix0 + lx, iy = iy0 + ly;
    if (ix > F.n[0] || iy > F.n[1]) continue;
    const size_t o = (size_t)(ix + NG - 1) + sy * (size_t)(iy + NG - 1);
    if (IS_E) {
      // s[] = bx, by, bz with box origin at (ix0-2, iy0-1): point (lx, ly) sits at (lx+2, ly+1)
      const int q = (ly + 1) * SW2 + (lx + XO);
      const double *bx = s[0], *by = s[1], *bz = s[2];
      double *ex = F.f[0], *ey = F.f[1], *ez = F.f[2];
      const double *jx = F.f[6], *jy = F.f[7], *jz = F.f[8];
      ex[o] = ex[o] + F.cy * (bz[q] - bz[q - SW2]) - F.fac * jx[o];
      ey[o] = ey[o] - F.cx * (bz[q] - bz[q - 1]) - F.fac * jy[o];
      ez[o] = ez[o] + F.cx * (by[q] - by[q - 1]) - F.cy * (bx[q] - bx[q - SW2]) - F.fac * jz[o];
    } else {
      // s[] = ex, ey, ez with box origin at (ix0, iy0)
      const int q = ly * SW2 + lx;
      const double *ex = s[0], *ey = s[1], *ez = s[2];
      double *bx = F.f[3], *by = F.f[4], *bz = F.f[5];
      bx[o] = bx[o] - F.cy * (ez[q + SW2] - ez[q]);
      by[o] = by[o] + F.cx * (ez[q + 1] - ez[q]);
      bz[o] = bz[o] - F.cx * (ey[q + 1] - ey[q]) + F.cy * (ex[q + SW2] - ex[q]);
    }
  }
}

// ---- 3D -------------------------------------------------------------------------------------
constexpr int SW3 = BX3 + 2, SH3 = BY3 + 1, SD3 = BZ3 + 1;  // 34 x 5 x 5
template <bool IS_E>
__global__ void __launch_bounds__(THREADS) k_fdtd_tma_3d(const __grid_constant__ TmaFieldParams F,
                                                        const __grid_constant__ CUtensorMap m0,
                                                        const __grid_constant__ CUtensorMap m1,
                                                        const __grid_constant__ CUtensorMap m2) {
  __shared__ __align__(128) double s[3][(SD3 * SH3 * SW3 + 15) / 16 * 16];
  __shared__ __align__(8) unsigned long long bar;
  const int ix0 = blockIdx.x * BX3, iy0 = blockIdx.y * BY3, iz0 = blockIdx.z * BZ3;
  const int sh = IS_E ? 1 : 0;
  constexpr int XO = IS_E ? 2 : 0;  // even innermost box coordinate, see the 2D kernel
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 3u * SD3 * SH3 * SW3 * sizeof(double));
    tma_load_3d(s[0], &m0, ix0 + NG - 1 - XO, iy0 + NG - 1 - sh, iz0 + NG - 1 - sh, &bar);
    tma_load_3d(s[1], &m1, ix0 + NG - 1 - XO, iy0 + NG - 1 - sh, iz0 + NG - 1 - sh, &bar);
    tma_load_3d(s[2], &m2, ix0 + NG - 1 - XO, iy0 + NG - 1 - sh, iz0 + NG - 1 - sh, &bar);
  }
  mbar_wait(&bar, 0);
  const size_t sy = F.sz[0], szz = (size_t)F.sz[0] * F.sz[1];
  constexpr int PY = SW3, PZ = SW3 * SH3;
#pragma unroll
  for (int r = 0; r < (BX3 * BY3 * BZ3) / THREADS; r++) {
    const int t = threadIdx.x + r * THREADS;
    const int lx = t % BX3, ly = (t / BX3) % BY3, lz = t / (BX3 * BY3);
    const int ix = ix0 + lx, iy = iy0 + ly, iz = iz0 + lz;
    if (ix > F.n[0] || iy > F.n[1] || iz > F.n[2]) continue;
    const size_t o = (size_t)(ix + NG - 1) + sy * (size_t)(iy + NG - 1) + szz * (size_t)(iz + NG - 1);
    if (IS_E) {
      const int q = (lz + 1) * PZ + (ly + 1) * PY + (lx + XO);
      const double *bx = s[0], *by = s[1], *bz = s[2];
      double *ex = F.f[0], *ey = F.f[1], *ez = F.f[2];
      const double *jx = F.f[6], *jy = F.f[7], *jz = F.f[8];
      ex[o] = ex[o] + F.cy * (bz[q] - bz[q - PY]) - F.cz * (by[q] - by[q - PZ]) - F.fac * jx[o];
      ey[o] = ey[o] + F.cz * (bx[q] - bx[q - PZ]) - F.cx * (bz[q] - bz[q - 1]) - F.fac * jy[o];
      ez[o] = ez[o] + F.cx * (by[q] - by[q - 1]) - F.cy * (bx[q] - bx[q - PY]) - F.fac * jz[o];
    } else {
      const int q = lz * PZ + ly * PY + lx;
      const double *ex = s[0], *ey = s[1], *ez = s[2];
      double *bx = F.f[3], *by = F.f[4], *bz = F.f[5];
      bx[o] = bx[o] - F.cy * (ez[q + PY] - ez[q]) + F.cz * (ey[q + PZ] - ey[q]);
      by[o] = by[o] - F.cz * (ex[q + PZ] - ex[q]) + F.cx * (ez[q + 1] - ez[q]);
      bz[o] = bz[o] - F.cx * (ey[q + 1] - ey[q]) + F.cy * (ex[q + PY] - ex[q]);
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// Builds one tensor map per E/B array.  Returns false (plain kernels are used) if the shape does
// not satisfy TMA's alignment rules or the driver entry point is unavailable.
bool epb_fdtd_tma_setup(epb_handle *h) {
  h->tma_ok = false;
  const int nd = h->cfg.ndims;
  if (nd < 2) return false;
  if (epb_env("EPB_NO_TMA")) return false;
  if ((h->sz[0] * sizeof(double)) % 16 != 0) return false;
  if (nd == 3 && ((size_t)h->sz[0] * h->sz[1] * sizeof(double)) % 16 != 0) return false;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  EncodeFn encode = (EncodeFn)fn;
  static_assert(sizeof(CUtensorMap) == sizeof(TmapStorage), "tensor map storage");
  for (int q = 0; q < 6; q++) {
    cuuint64_t dims[3] = {(cuuint64_t)h->sz[0], (cuuint64_t)h->sz[1], (cuuint64_t)h->sz[2]};
    cuuint64_t strides[2] = {(cuuint64_t)h->sz[0] * sizeof(double), (cuuint64_t)h->sz[0] * h->sz[1] * sizeof(double)};
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    if (nd == 2) { box[0] = SW2; box[1] = SH2; box[2] = 1; }
    else { box[0] = SW3; box[1] = SH3; box[2] = SD3; }
    CUresult r = encode(reinterpret_cast<CUtensorMap *>(&h->tmap[q]), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)nd,
                        h->f(q), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
  }
  h->tma_ok = true;
  return true;
}

// is_e: update_e_field (reads B through the stencil) else update_b_field (reads E)
void epb_fdtd_tma_launch(epb_handle *h, bool is_e, double cx, double cy, double cz, double fac) {
  TmaFieldParams F;
  for (int q = 0; q < 9; q++) F.f[q] = h->f(q);
  for (int d = 0; d < 3; d++) { F.n[d] = h->cfg.n[d]; F.sz[d] = h->sz[d]; }
  F.cx = cx; F.cy = cy; F.cz = cz; F.fac = fac;
  const CUtensorMap *m = reinterpret_cast<const CUtensorMap *>(h->tmap) + (is_e ? 3 : 0);
  const int nd = h->cfg.ndims;
  if (nd == 2) {
    dim3 grid((h->cfg.n[0] + 1 + BX2 - 1) / BX2, (h->cfg.n[1] + 1 + BY2 - 1) / BY2);
    if (is_e) k_fdtd_tma_2d<true><<<grid, THREADS, 0, h->stream>>>(F, m[0], m[1], m[2]);
    else k_fdtd_tma_2d<false><<<grid, THREADS, 0, h->stream>>>(F, m[0], m[1], m[2]);
  } else {
    dim3 grid((h->cfg.n[0] + 1 + BX3 - 1) / BX3, (h->cfg.n[1] + 1 + BY3 - 1) / BY3, (h->cfg.n[2] + 1 + BZ3 - 1) / BZ3);
    if (is_e) k_fdtd_tma_3d<true><<<grid, THREADS, 0, h->stream>>>(F, m[0], m[1], m[2]);
    else k_fdtd_tma_3d<false><<<grid, THREADS, 0, h->stream>>>(F, m[0], m[1], m[2]);
  }
  h->launches++;
}
