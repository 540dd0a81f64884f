// push.cuh — Boris push + triangle-shape gather + Esirkepov deposit + particle BC
// classification, for sm_100a.  Included twice: push_strict.cu (compiled with
// -fmad=false: operation-for-operation the arithmetic of particles.F90, which the
// reference builds without FMA contraction, epoch2d/Makefile:72) and push_fast.cu
// (FMA contraction allowed; differs from the former at the 1e-16 level).
//
// Reference: epoch{1,2,3}d/src/particles.F90:28-650 and
// src/include/triangle/{gx,hx_dcell,e_part,b_part}.inc; boundary classification
// follows boundary.F90:1064-1433 (no thermal / CPML).
//
// Kernels:
//  * push_cell_2d<CTY,MINB> (default in 2D): one CTA per 16xCTY-cell tile, one warp per 16x2-cell
//    group, one LANE per CELL.  The particles of a group are stored interleaved by their rank in
//    the cell (sort.cu), so a round of a warp is one coalesced load; a lane keeps the 21
//    non-cancelling deposit sums of its cell's 3x3 stencil in registers for the whole tile and
//    touches shared memory for the deposit only when it flushes them.  It also emits the records
//    the next sort is built from and applies the previous sort's permutation (fused gather).
//  * push_tiled_2d<V21> (EPB_PUSH_VARIANT=0/1): one CTA per 16x16-cell tile of the cell-major
//    layout, one lane per particle; the 27 (21) deposit values of the lanes that share a cell are
//    summed through a per-warp transposition scratch, one shared update per (cell, value).
//  * push_tiled_3d: one CTA per 8x8x4-cell tile, J in shared memory, E/B through the read-only
//    path; lanes grouped by cell, plane-by-plane transposed reduction of the 54 values.
//  * push_generic<ND>: any particle range, fields through the read-only path, deposit with
//    global FP64 reductions (RED.E.ADD.F64).  1D, the unsorted tail (arrivals since the last
//    sort) and the per-particle fallback of the tiled kernels (push_one).
// In all tiled kernels a particle whose nearest cell moved by one cell along one axis keeps its
// core in the fast path (shifted weights) and only the values outside the core are queued
// (drain_edge / drain_edge_3d); shared FP64 atomicAdd is an ATOMS.CAST.SPIN loop on sm_100a, so
// shared updates per particle are what every kernel here avoids.
#pragma once
#include <cstdlib>

#include "epb_internal.h"

namespace EPB_NS {

constexpr int NG = EPB_NG;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void tri(double f, double &gm, double &g0, double &gp) {
  // include/triangle/gx.inc:1-4
  double cf2 = f * f;
  gm = 0.25 + cf2 + f;
  g0 = 1.5 - 2.0 * cf2;
  gp = 0.25 + cf2 - f;
}

// The same weights for the performance build in four FP64 instructions (a = f^2 + 1/4 fused, g0 = 2 - 2a);
// the parity build keeps the reference's expression tree.
__device__ __forceinline__ void tri_s(double f, double &gm, double &g0, double &gp) {
#ifdef EPB_FAST_MATH
  const double a = fma(f, f, 0.25);
  gm = a + f;
  gp = a - f;
  g0 = fma(-2.0, a, 2.0);
#else
  tri(f, gm, g0, gp);
#endif
}

// boundary.F90:1064-1433 for one particle.  Returns -1 if the particle stays on
// this rank, else the outbox slot: direction index (iz+1)*9+(iy+1)*3+(ix+1), or 13
// for a particle that left the system (open boundary, beyond x_min_outer).
template <int ND>
__device__ __forceinline__ int particle_bc(const PushParams &P, double *pos, double *mom) {
  bool cand = false;
#pragma unroll
  for (int d = 0; d < ND; d++) cand = cand || (pos[d] < P.bnd_min[d]) || (pos[d] > P.bnd_max[d]);
  if (!cand) return -1;
  int bd[3] = {0, 0, 0};
  bool oob = false;
#pragma unroll
  for (int d = 0; d < ND; d++) {
    const double part_pos = pos[d];
    if (part_pos < P.min_local[d]) {
      bd[d] = -1;
      const int bc = P.bc_min[d];
      if (bc == EPB_BC_REFLECT) {
        if (P.is_bnd_min[d]) {
          bd[d] = 0;
          pos[d] = 2.0 * P.gmin[d] - part_pos;
          mom[d] = -mom[d];
        }
      } else if (bc == EPB_BC_PERIODIC) {
        if (P.is_bnd_min[d]) pos[d] = part_pos + P.shift[d];
      } else if (bc == EPB_BC_THERMAL) {
        // boundary.F90:1104-1148: the particle stays on this rank; beyond x_min_outer it is re-emitted from the wall's
        // thermal distribution by k_thermal right after the push (epb_api.cu), which needs random numbers
        if (P.is_bnd_min[d]) bd[d] = 0;
      } else {
        if (part_pos < P.min_outer[d]) { bd[d] = 0; oob = true; }
        else if (P.is_bnd_min[d]) bd[d] = 0;
      }
    }
    if (part_pos >= P.max_local[d]) {
      bd[d] = 1;
      const int bc = P.bc_max[d];
      if (bc == EPB_BC_REFLECT) {
        if (P.is_bnd_max[d]) {
          bd[d] = 0;
          pos[d] = 2.0 * P.gmax[d] - part_pos;
          mom[d] = -mom[d];
        }
      } else if (bc == EPB_BC_PERIODIC) {
        if (P.is_bnd_max[d]) pos[d] = part_pos - P.shift[d];
      } else if (bc == EPB_BC_THERMAL) {
        if (P.is_bnd_max[d]) bd[d] = 0;
      } else {
        if (part_pos >= P.max_outer[d]) { bd[d] = 0; oob = true; }
        else if (P.is_bnd_max[d]) bd[d] = 0;
      }
    }
  }
  if (oob) return 13;
  const int dir = (bd[2] + 1) * 9 + (bd[1] + 1) * 3 + (bd[0] + 1);
  if (dir == 13) return -1;
  if (P.nbr_is_self[dir]) return -1;  // periodic wrap onto this rank: already shifted
  return dir;
}

__device__ __forceinline__ void outbox_put(const PushParams &P, long long i, int dir) {
  int slot = atomicAdd(&P.out_count[dir], 1);
  if (slot < P.out_cap) P.out_idx[(size_t)dir * P.out_cap + slot] = (int)i;
  P.gone[i] = 1;
}

template <int ND>
__device__ __forceinline__ size_t gofs(const PushParams &P, int cx, int cy, int cz) {
  size_t o = (size_t)(cx + NG - 1);
  if (ND >= 2) o += (size_t)P.sz[0] * (size_t)(cy + NG - 1);
  if (ND >= 3) o += (size_t)P.sz[0] * (size_t)P.sz[1] * (size_t)(cz + NG - 1);
  return o;
}

// include/triangle/e_part.inc: rows parenthesised, sums left to right
template <int ND>
__device__ __forceinline__ double gather_g(const PushParams &P, const double *__restrict__ F,
                                           const double *wx, int cx, const double *wy, int cy,
                                           const double *wz, int cz) {
  if (ND == 1) {
    size_t o = gofs<1>(P, cx, 1, 1);
    return wx[0] * __ldg(F + o - 1) + wx[1] * __ldg(F + o) + wx[2] * __ldg(F + o + 1);
  } else if (ND == 2) {
    double r = 0.0;
#pragma unroll
    for (int iy = 0; iy < 3; iy++) {
      size_t o = gofs<2>(P, cx, cy + iy - 1, 1);
      double row = wx[0] * __ldg(F + o - 1) + wx[1] * __ldg(F + o) + wx[2] * __ldg(F + o + 1);
      double t = wy[iy] * row;
      r = (iy == 0) ? t : r + t;
    }
    return r;
  } else {
    double r = 0.0;
#pragma unroll
    for (int iz = 0; iz < 3; iz++) {
      double pl = 0.0;
#pragma unroll
      for (int iy = 0; iy < 3; iy++) {
        size_t o = gofs<3>(P, cx, cy + iy - 1, cz + iz - 1);
        double row = wx[0] * __ldg(F + o - 1) + wx[1] * __ldg(F + o) + wx[2] * __ldg(F + o + 1);
        double t = wy[iy] * row;
        pl = (iy == 0) ? t : pl + t;
      }
      double t = wz[iz] * pl;
      r = (iz == 0) ? t : r + t;
    }
    return r;
  }
}

// One particle, everything through global memory.  Mirrors the oracle routine
// push_particles<ND> line by line.
template <int ND, bool HC = false>
__device__ __noinline__ void push_one(const PushParams &P, long long i) {
  const double c = EPB_C;
  const double part_weight = P.w[i];
  const double fcx = P.kfc[0] * part_weight;
  const double fcy = P.kfc[1] * part_weight;
  const double fcz = P.kfc[2] * part_weight;
  double part_pos[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < ND; d++) part_pos[d] = P.x[d][i] - P.grid_min_local[d];
  double part_ux = P.p[0][i] * P.ipart_mc;
  double part_uy = P.p[1][i] * P.ipart_mc;
  double part_uz = P.p[2][i] * P.ipart_mc;
  double gamma_rel = sqrt(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0);
  double root = P.dtco2 / gamma_rel;
  {
    const double u[3] = {part_ux, part_uy, part_uz};
#pragma unroll
    for (int d = 0; d < ND; d++) part_pos[d] = part_pos[d] + u[d] * root;
  }
  double G[3][5], H[3][5];
  int cell1[3] = {1, 1, 1}, cell2[3] = {1, 1, 1};
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int q = 0; q < 5; q++) { G[d][q] = 0.0; H[d][q] = 0.0; }
#pragma unroll
  for (int d = 0; d < ND; d++) {
    double cell_r = part_pos[d] * P.idx[d];
    int c1 = __double2int_rd(cell_r + 0.5);
    double cell_frac = (double)c1 - cell_r;
    cell1[d] = c1 + 1;
    tri(cell_frac, G[d][1], G[d][2], G[d][3]);
    int c2 = __double2int_rd(cell_r);
    cell_frac = (double)c2 - cell_r + 0.5;
    cell2[d] = c2 + 1;
    tri(cell_frac, H[d][1], H[d][2], H[d][3]);
  }
  const double *gx = &G[0][1], *gy = &G[1][1], *gz = &G[2][1];
  const double *hx = &H[0][1], *hy = &H[1][1], *hz = &H[2][1];
  const double ex_part = gather_g<ND>(P, P.e[0], hx, cell2[0], gy, cell1[1], gz, cell1[2]);
  const double ey_part = gather_g<ND>(P, P.e[1], gx, cell1[0], hy, cell2[1], gz, cell1[2]);
  const double ez_part = gather_g<ND>(P, P.e[2], gx, cell1[0], gy, cell1[1], hz, cell2[2]);
  const double bx_part = gather_g<ND>(P, P.b[0], gx, cell1[0], hy, cell2[1], hz, cell2[2]);
  const double by_part = gather_g<ND>(P, P.b[1], hx, cell2[0], gy, cell1[1], hz, cell2[2]);
  const double bz_part = gather_g<ND>(P, P.b[2], hx, cell2[0], hy, cell2[1], gz, cell1[2]);
  // particles.F90:382-428
  const double cmratio = P.cmratio;
  double uxm = part_ux + cmratio * ex_part;
  double uym = part_uy + cmratio * ey_part;
  double uzm = part_uz + cmratio * ez_part;
  if (HC) {  // particles.F90:386-398 (-DHC_PUSH), Higuera & Cary, Phys. Plasmas 24, 052104
    gamma_rel = uxm * uxm + uym * uym + uzm * uzm + 1.0;
    const double beta_x = P.hc_alpha * bx_part;
    const double beta_y = P.hc_alpha * by_part;
    const double beta_z = P.hc_alpha * bz_part;
    const double beta2 = beta_x * beta_x + beta_y * beta_y + beta_z * beta_z;
    const double sigma = gamma_rel - beta2;
    const double beta_dot_u = beta_x * uxm + beta_y * uym + beta_z * uzm;
    gamma_rel = sigma + sqrt(sigma * sigma + 4.0 * (beta2 + beta_dot_u * beta_dot_u));
    gamma_rel = sqrt(0.5 * gamma_rel);
  } else {
    gamma_rel = sqrt(uxm * uxm + uym * uym + uzm * uzm + 1.0);
  }
  root = P.ccmratio / gamma_rel;
  double taux = bx_part * root, tauy = by_part * root, tauz = bz_part * root;
  double taux2 = taux * taux, tauy2 = tauy * tauy, tauz2 = tauz * tauz;
  double tau = 1.0 / (1.0 + taux2 + tauy2 + tauz2);
  double uxp = ((1.0 + taux2 - tauy2 - tauz2) * uxm +
                2.0 * ((taux * tauy + tauz) * uym + (taux * tauz - tauy) * uzm)) * tau;
  double uyp = ((1.0 - taux2 + tauy2 - tauz2) * uym +
                2.0 * ((tauy * tauz + taux) * uzm + (tauy * taux - tauz) * uxm)) * tau;
  double uzp = ((1.0 - taux2 - tauy2 + tauz2) * uzm +
                2.0 * ((tauz * taux + tauy) * uxm + (tauz * tauy - taux) * uym)) * tau;
  part_ux = uxp + cmratio * ex_part;
  part_uy = uyp + cmratio * ey_part;
  part_uz = uzp + cmratio * ez_part;
  double part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz;
  gamma_rel = sqrt(part_u2 + 1.0);
  double delta[3] = {0, 0, 0}, part_vy = 0.0, part_vz = 0.0;
  if (ND == 1) {  // epoch1d particles.F90:392-396
    root = c / gamma_rel;
    delta[0] = part_ux * root * P.dto2;
    part_vy = part_uy * root;
    part_vz = part_uz * root;
  } else if (ND == 2) {  // epoch2d particles.F90:433-438
    double igamma = 1.0 / gamma_rel;
    root = P.dtco2 * igamma;
    delta[0] = part_ux * root;
    delta[1] = part_uy * root;
    part_vz = part_uz * c * igamma;
  } else {  // epoch3d particles.F90:470-474
    root = P.dtco2 / gamma_rel;
    delta[0] = part_ux * root;
    delta[1] = part_uy * root;
    delta[2] = part_uz * root;
  }
#pragma unroll
  for (int d = 0; d < ND; d++) part_pos[d] = part_pos[d] + delta[d];
  {
    double pos[3] = {0, 0, 0}, mom[3];
#pragma unroll
    for (int d = 0; d < ND; d++) pos[d] = part_pos[d] + P.grid_min_local[d];
    mom[0] = P.part_mc * part_ux;
    mom[1] = P.part_mc * part_uy;
    mom[2] = P.part_mc * part_uz;
    int dir = particle_bc<ND>(P, pos, mom);
#pragma unroll
    for (int d = 0; d < ND; d++) P.x[d][i] = pos[d];
#pragma unroll
    for (int d = 0; d < 3; d++) P.p[d][i] = mom[d];
    if (dir >= 0) outbox_put(P, i, dir);
  }
  if (!P.deposit) return;
  int dcell[3] = {0, 0, 0}, mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < ND; d++) {
    part_pos[d] = part_pos[d] + delta[d];
    double cell_r = part_pos[d] * P.idx[d];
    int c3 = __double2int_rd(cell_r + 0.5);
    double cell_frac = (double)c3 - cell_r;
    c3 = c3 + 1;
    dcell[d] = c3 - cell1[d];
    double wm, w0, wp;
    tri(cell_frac, wm, w0, wp);
#pragma unroll
    for (int q = 0; q < 5; q++) {
      // hx = 0; hx(dcell-1:dcell+1) = weights; hx = hx - gx   (particles.F90:521-538)
      int r = q - 2 - dcell[d];
      double hv = (r == -1) ? wm : (r == 0) ? w0 : (r == 1) ? wp : 0.0;
      H[d][q] = hv - G[d][q];
    }
    mn[d] = -1 + (dcell[d] - 1) / 2;
    mx[d] = 1 + (dcell[d] + 1) / 2;
  }
  gx = &G[0][2]; gy = &G[1][2]; gz = &G[2][2];
  hx = &H[0][2]; hy = &H[1][2]; hz = &H[2][2];
  const double third = P.third;
  if (ND == 1) {  // epoch1d particles.F90:489-507
    const double fjx = fcx * P.part_q;
    const double fjy = fcy * P.part_q * part_vy;
    const double fjz = fcy * P.part_q * part_vz;
    double jxh = 0.0;
    for (int ix = mn[0]; ix <= mx[0]; ix++) {
      size_t o = gofs<1>(P, cell1[0] + ix, 1, 1);
      double wx = hx[ix];
      double wy = gx[ix] + 0.5 * hx[ix];
      jxh = jxh - fjx * wx;
      atomicAdd(P.j[0] + o, jxh);
      atomicAdd(P.j[1] + o, fjy * wy);
      atomicAdd(P.j[2] + o, fjz * wy);
    }
  } else if (ND == 2) {  // epoch2d particles.F90:549-579
    const double fjx = fcx * P.part_q;
    const double fjy = fcy * P.part_q;
    const double fjz = fcz * P.part_q * part_vz;
    double jyh[5] = {0, 0, 0, 0, 0};
    for (int iy = mn[1]; iy <= mx[1]; iy++) {
      double yfac1 = gy[iy] + 0.5 * hy[iy];
      double yfac2 = third * hy[iy] + 0.5 * gy[iy];
      double hy_iy = hy[iy];
      double jxh = 0.0;
      for (int ix = mn[0]; ix <= mx[0]; ix++) {
        size_t o = gofs<2>(P, cell1[0] + ix, cell1[1] + iy, 1);
        double xfac1 = gx[ix] + 0.5 * hx[ix];
        double wx = hx[ix] * yfac1;
        double wy = hy_iy * xfac1;
        double wz = gx[ix] * yfac1 + hx[ix] * yfac2;
        jxh = jxh - fjx * wx;
        jyh[ix + 2] = jyh[ix + 2] - fjy * wy;
        double jzh = fjz * wz;
        atomicAdd(P.j[0] + o, jxh);
        atomicAdd(P.j[1] + o, jyh[ix + 2]);
        atomicAdd(P.j[2] + o, jzh);
      }
    }
  } else {  // epoch3d particles.F90:603-648
    const double fjx = fcx * P.part_q;
    const double fjy = fcy * P.part_q;
    const double fjz = fcz * P.part_q;
    double jzh[5][5];
    for (int a = 0; a < 5; a++)
      for (int b = 0; b < 5; b++) jzh[a][b] = 0.0;
    for (int iz = mn[2]; iz <= mx[2]; iz++) {
      double zfac1 = gz[iz] + 0.5 * hz[iz];
      double zfac2 = third * hz[iz] + 0.5 * gz[iz];
      double gz_iz = gz[iz], hz_iz = hz[iz];
      double jyh[5] = {0, 0, 0, 0, 0};
      for (int iy = mn[1]; iy <= mx[1]; iy++) {
        double yfac1 = gy[iy] + 0.5 * hy[iy];
        double yfac2 = third * hy[iy] + 0.5 * gy[iy];
        double hygz = hy[iy] * gz_iz;
        double hyhz = hy[iy] * hz_iz;
        double yzfac = gy[iy] * zfac1 + hy[iy] * zfac2;
        double hzyfac1 = hz_iz * yfac1;
        double hzyfac2 = hz_iz * yfac2;
        double jxh = 0.0;
        for (int ix = mn[0]; ix <= mx[0]; ix++) {
          size_t o = gofs<3>(P, cell1[0] + ix, cell1[1] + iy, cell1[2] + iz);
          double xfac1 = gx[ix] + 0.5 * hx[ix];
          double xfac2 = third * hx[ix] + 0.5 * gx[ix];
          double wx = hx[ix] * yzfac;
          double wy = xfac1 * hygz + xfac2 * hyhz;
          double wz = gx[ix] * hzyfac1 + hx[ix] * hzyfac2;
          jxh = jxh - fjx * wx;
          jyh[ix + 2] = jyh[ix + 2] - fjy * wy;
          jzh[iy + 2][ix + 2] = jzh[iy + 2][ix + 2] - fjz * wz;
          atomicAdd(P.j[0] + o, jxh);
          atomicAdd(P.j[1] + o, jyh[ix + 2]);
          atomicAdd(P.j[2] + o, jzh[iy + 2][ix + 2]);
        }
      }
    }
  }
}

template <int ND, bool HC = false>
__global__ void __launch_bounds__(256) push_generic(const __grid_constant__ PushParams P) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = P.first + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.last; i += stride)
    push_one<ND, HC>(P, i);
}

// ---------------------------------------------------------------------------
// Tiled 2D kernel
// ---------------------------------------------------------------------------
// One CTA per 16x16-cell tile of the cell-sorted layout, two CTAs per SM.  Shared memory:
//   sF  [6][TH][TW]   E/B tile + 3 halo cells
//   sJ  [3][TH][TW]   current accumulated by this CTA, flushed once with global reductions
//   sS  [warp][27][33] per-warp transposition scratch for the deposit reduction
//   sQ* [warp][...]    per-warp queue of particles that changed their nearest cell this step
//   sSlow             CTA list of particles outside the tile's halo (stale sort, wrapped)
// The 32 lanes of a warp hold 32 consecutive particles of the sorted range, i.e. mostly one
// or two cells.  Each lane writes its 27 deposit values (3 components x 3x3 cells around its
// nearest cell) to a column of sS; lane q < 27 then sums row q over the lanes that share a
// cell key and issues ONE shared-memory update per key (shared FP64 atomicAdd is a CAS loop
// on sm_100a, so updates per particle are what must be avoided).  The cost is independent of
// how many distinct cells the warp spans, which keeps the kernel efficient between sorts.
// Particles whose nearest cell changed during the step (a few %) have a wider stencil: they
// are queued and deposited densely, 32 at a time, with the reference's general loop.
constexpr int T2X = 16, T2Y = 16, HALO = 3;
constexpr int TW = T2X + 2 * HALO, TH = T2Y + 2 * HALO;
constexpr int TILE_ELEMS = TW * TH;
constexpr int PUSH2D_THREADS = 256, PUSH2D_WARPS = PUSH2D_THREADS / 32;
constexpr int SROWS = 27, SPITCH = 33;
constexpr int QCAP = 32, QDBL = 7;
constexpr int SLOWCAP = 510;
constexpr size_t PUSH2D_SMEM =
    sizeof(double) * ((size_t)9 * TILE_ELEMS + (size_t)PUSH2D_WARPS * SROWS * SPITCH + (size_t)PUSH2D_WARPS * QDBL * QCAP) +
    sizeof(int) * ((size_t)PUSH2D_WARPS * QCAP + SLOWCAP + 2);

__device__ __forceinline__ void smem_add(double *addr, double v) {
  // shared FP64 add: an ATOMS.CAST.SPIN loop; conflicts between warps are rare
  atomicAdd(addr, v);
}


// Performance build only: reciprocal square root and reciprocal of arguments that are known to
// be >= 1 (gamma^2 = u^2 + 1, 1 + tau^2), so the special-case paths of rsqrt() / the IEEE divide
// (denormals, infinities, the out-of-line slow path) are dead weight.  MUFU seed (about 2^-22
// relative error) + two Newton steps: within 1-2 ulp, like rsqrt().
#ifdef EPB_FAST_MATH
__device__ __forceinline__ double rsqrt_ge1(double s) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  double h = 0.5 * s;
  y = y * (1.5 - h * y * y);
  y = y * (1.5 - h * y * y);
  return y;
}
__device__ __forceinline__ double rcp_ge1(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}
#endif

// 1/sqrt(s) and friends.  The parity build keeps the reference's sqrt + divide sequence.
__device__ __forceinline__ void gamma_root(double s, double num, double &root) {
#ifdef EPB_FAST_MATH
  root = num * rsqrt_ge1(s);
#else
  root = num / sqrt(s);
#endif
}

// Edge deposit of a queued "semi-regular" particle of push_cell_2d: its nearest cell moved by one
// cell along exactly one axis (dc = +-1), its owner lane has already accumulated the 3x3 core of
// the stencil in registers, and only the values outside the core remain: the extra column/row at
// +-2, and -- because the lane keeps no sum for the third jx column / jy row -- the prefix value
// at +1 when dc = +1.  8 shared-memory updates instead of the general loop's 36.
__device__ __forceinline__ void drain_edge(const PushParams &P, double *sJ, double fxo, double fxn, double fyo,
                                           double fyn, double fjx, double fjy, double fjz, int key, int dcx,
                                           int dcy, int jstride, int pitch, bool core_has_third) {
  double gx[3], gy[3], nx[3], ny[3], hx[3], hy[3];
  tri(fxo, gx[0], gx[1], gx[2]);
  tri(fyo, gy[0], gy[1], gy[2]);
  tri(fxn, nx[0], nx[1], nx[2]);
  tri(fyn, ny[0], ny[1], ny[2]);
  const double third = P.third;
  if (dcx != 0) {
    const double hxe = dcx > 0 ? nx[2] : nx[0];
    if (dcx > 0) { hx[0] = 0.0 - gx[0]; hx[1] = nx[0] - gx[1]; hx[2] = nx[1] - gx[2]; }
    else { hx[0] = nx[1] - gx[0]; hx[1] = nx[2] - gx[1]; hx[2] = 0.0 - gx[2]; }
#pragma unroll
    for (int q = 0; q < 3; q++) hy[q] = ny[q] - gy[q];
    const double xfac1e = 0.5 * hxe;
    const double hxs = hx[0] + hx[1] + hx[2];
    const int oe = key + 2 * dcx;
    double jyh = 0.0;
#pragma unroll
    for (int iy = 0; iy < 3; iy++) {
      const double yfac1 = gy[iy] + 0.5 * hy[iy];
      const double yfac2 = third * hy[iy] + 0.5 * gy[iy];
      const int row = (iy - 1) * pitch;
      if (dcx < 0) smem_add(&sJ[oe + row], -(fjx * (hxe * yfac1)));
      else if (!core_has_third) smem_add(&sJ[key + 1 + row], -(fjx * (hxs * yfac1)));
      if (iy < 2) {
        jyh = jyh - fjy * (hy[iy] * xfac1e);
        smem_add(&sJ[jstride + oe + row], jyh);
      }
      smem_add(&sJ[2 * jstride + oe + row], fjz * (hxe * yfac2));
    }
  } else {
    const double hye = dcy > 0 ? ny[2] : ny[0];
    if (dcy > 0) { hy[0] = 0.0 - gy[0]; hy[1] = ny[0] - gy[1]; hy[2] = ny[1] - gy[2]; }
    else { hy[0] = ny[1] - gy[0]; hy[1] = ny[2] - gy[1]; hy[2] = 0.0 - gy[2]; }
#pragma unroll
    for (int q = 0; q < 3; q++) hx[q] = nx[q] - gx[q];
    const double yfac1e = 0.5 * hye, yfac2e = third * hye;
    const double hys = hy[0] + hy[1] + hy[2];
    const int oe = key + 2 * dcy * pitch;
    double jxh = 0.0;
#pragma unroll
    for (int ix = 0; ix < 3; ix++) {
      const double xfac1 = gx[ix] + 0.5 * hx[ix];
      if (ix < 2) {
        jxh = jxh - fjx * (hx[ix] * yfac1e);
        smem_add(&sJ[oe + ix - 1], jxh);
      }
      if (dcy < 0) smem_add(&sJ[jstride + oe + ix - 1], -(fjy * (hye * xfac1)));
      else if (!core_has_third) smem_add(&sJ[jstride + key + pitch + ix - 1], -(fjy * (hys * xfac1)));
      smem_add(&sJ[2 * jstride + oe + ix - 1], fjz * (gx[ix] * yfac1e + hx[ix] * yfac2e));
    }
  }
}

// Deposit of queued particles (nearest cell changed: dcell != 0 in x and/or y): the general
// loop of particles.F90:549-579 over xmin..xmax, ymin..ymax with shared-memory updates.
__device__ __forceinline__ void drain_extras(const PushParams &P, double *sJ, const double *Qd, const int *Qk,
                                             int n, int lane, int jstride = TILE_ELEMS, int pitch = TW) {
  if (lane >= n) return;
  const int pk = Qk[lane];
  const int key = pk & 1023, dcx = ((pk >> 10) & 3) - 1, dcy = ((pk >> 12) & 3) - 1;
  const double fjx = Qd[4 * QCAP + lane], fjy = Qd[5 * QCAP + lane], fjz = Qd[6 * QCAP + lane];
  if (pk & (1 << 14)) {  // push_cell_2d: core already accumulated by the owner lane
    drain_edge(P, sJ, Qd[0 * QCAP + lane], Qd[1 * QCAP + lane], Qd[2 * QCAP + lane], Qd[3 * QCAP + lane], fjx, fjy,
               fjz, key, dcx, dcy, jstride, pitch, (pk & (1 << 16)) != 0);
    return;
  }
  if (pk & (1 << 15)) {
    // push_cell_2d, particle outside its lane's cell (stale order) whose nearest cell did not
    // change: the 3x3 stencil without its six structurally cancelling values (last jx column,
    // last jy row), 21 updates instead of the general loop's 27
    double g3x[3], g3y[3], h3x[3], h3y[3];
    tri(Qd[0 * QCAP + lane], g3x[0], g3x[1], g3x[2]);
    tri(Qd[2 * QCAP + lane], g3y[0], g3y[1], g3y[2]);
    tri(Qd[1 * QCAP + lane], h3x[0], h3x[1], h3x[2]);
    tri(Qd[3 * QCAP + lane], h3y[0], h3y[1], h3y[2]);
#pragma unroll
    for (int q = 0; q < 3; q++) { h3x[q] = h3x[q] - g3x[q]; h3y[q] = h3y[q] - g3y[q]; }
    const double third = P.third;
    double jyh[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int iy = 0; iy < 3; iy++) {
      const double yfac1 = g3y[iy] + 0.5 * h3y[iy];
      const double yfac2 = third * h3y[iy] + 0.5 * g3y[iy];
      double jxh = 0.0;
#pragma unroll
      for (int ix = 0; ix < 3; ix++) {
        const int o = key + (iy - 1) * pitch + (ix - 1);
        if (ix < 2) {
          jxh = jxh - fjx * (h3x[ix] * yfac1);
          smem_add(&sJ[o], jxh);
        }
        if (iy < 2) {
          jyh[ix] = jyh[ix] - fjy * (h3y[iy] * (g3x[ix] + 0.5 * h3x[ix]));
          smem_add(&sJ[jstride + o], jyh[ix]);
        }
        smem_add(&sJ[2 * jstride + o], fjz * (g3x[ix] * yfac1 + h3x[ix] * yfac2));
      }
    }
    return;
  }
  double gx[5], gy[5], hx[5], hy[5];
  gx[0] = gx[4] = gy[0] = gy[4] = 0.0;
  tri(Qd[0 * QCAP + lane], gx[1], gx[2], gx[3]);
  tri(Qd[2 * QCAP + lane], gy[1], gy[2], gy[3]);
  double wm, w0, wp;
  tri(Qd[1 * QCAP + lane], wm, w0, wp);
#pragma unroll
  for (int q = 0; q < 5; q++) {
    const int r = q - 2 - dcx;
    hx[q] = ((r == -1) ? wm : (r == 0) ? w0 : (r == 1) ? wp : 0.0) - gx[q];
  }
  tri(Qd[3 * QCAP + lane], wm, w0, wp);
#pragma unroll
  for (int q = 0; q < 5; q++) {
    const int r = q - 2 - dcy;
    hy[q] = ((r == -1) ? wm : (r == 0) ? w0 : (r == 1) ? wp : 0.0) - gy[q];
  }
  const int xmin = -1 + (dcx - 1) / 2, xmax = 1 + (dcx + 1) / 2;
  const int ymin = -1 + (dcy - 1) / 2, ymax = 1 + (dcy + 1) / 2;
  const double third = P.third;
  double jyh[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int iy = -2; iy <= 2; iy++) {
    if (iy < ymin || iy > ymax) continue;
    const double yfac1 = gy[iy + 2] + 0.5 * hy[iy + 2];
    const double yfac2 = third * hy[iy + 2] + 0.5 * gy[iy + 2];
    double jxh = 0.0;
#pragma unroll
    for (int ix = -2; ix <= 2; ix++) {
      if (ix < xmin || ix > xmax) continue;
      const double xfac1 = gx[ix + 2] + 0.5 * hx[ix + 2];
      const double wx = hx[ix + 2] * yfac1;
      const double wy = hy[iy + 2] * xfac1;
      const double wz = gx[ix + 2] * yfac1 + hx[ix + 2] * yfac2;
      jxh = jxh - fjx * wx;
      jyh[ix + 2] = jyh[ix + 2] - fjy * wy;
      const double jzh = fjz * wz;
      const int o = key + iy * pitch + ix;
      smem_add(&sJ[o], jxh);
      smem_add(&sJ[jstride + o], jyh[ix + 2]);
      smem_add(&sJ[2 * jstride + o], jzh);
    }
  }
}

// V21: drop the six structurally cancelling stencil values (last jx column, last jy row)
template <bool V21>
__global__ void __launch_bounds__(PUSH2D_THREADS, 2) push_tiled_2d(const __grid_constant__ PushParams P) {
  extern __shared__ double sm[];
  double *sF = sm;                                   // [6][TH][TW]
  double *sJ = sF + 6 * TILE_ELEMS;                  // [3][TH][TW]
  double *sS_all = sJ + 3 * TILE_ELEMS;              // [warps][27][33]
  double *sQd_all = sS_all + PUSH2D_WARPS * SROWS * SPITCH;
  int *sQk_all = reinterpret_cast<int *>(sQd_all + PUSH2D_WARPS * QDBL * QCAP);
  int *sSlow = sQk_all + PUSH2D_WARPS * QCAP;
  int *sSlowCount = sSlow + SLOWCAP;
  const int tile = blockIdx.x;
  const int ttx = tile % P.tg.nt[0], tty = tile / P.tg.nt[0];
  const int ox = ttx * T2X + 1 - HALO;  // cell index of shared column 0
  const int oy = tty * T2Y + 1 - HALO;
  const long long start = P.tile_start[tile];
  long long end = P.tile_start[tile + 1];
  if (end > P.n_sorted_clip) end = P.n_sorted_clip;
  if (start >= end) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) *sSlowCount = 0;
  for (int q = tid; q < TILE_ELEMS; q += PUSH2D_THREADS) {
    const int lx = q % TW, ly = q / TW;
    const int cx = ox + lx, cy = oy + ly;
    const bool ok = (cx >= 1 - NG) && (cx <= P.n[0] + NG) && (cy >= 1 - NG) && (cy <= P.n[1] + NG);
    const size_t o = ok ? gofs<2>(P, cx, cy, 1) : 0;
#pragma unroll
    for (int f = 0; f < 3; f++) {
      sF[f * TILE_ELEMS + q] = ok ? __ldg(P.e[f] + o) : 0.0;
      sF[(3 + f) * TILE_ELEMS + q] = ok ? __ldg(P.b[f] + o) : 0.0;
      sJ[f * TILE_ELEMS + q] = 0.0;
    }
  }
  __syncthreads();

  const double c = EPB_C;
  const double third = P.third;
  const double *sEx = sF, *sEy = sF + TILE_ELEMS, *sEz = sF + 2 * TILE_ELEMS;
  const double *sBx = sF + 3 * TILE_ELEMS, *sBy = sF + 4 * TILE_ELEMS, *sBz = sF + 5 * TILE_ELEMS;
  double *S = sS_all + warp * SROWS * SPITCH;
  double *Qd = sQd_all + warp * QDBL * QCAP;
  int *Qk = sQk_all + warp * QCAP;
  int qcount = 0;  // warp-uniform
  // lane q < NR owns one deposit value: q = comp*9 + iy*3 + ix of the 3x3 stencil, or (V21)
  // rows 0..5 = jx(iy, ix<2), 6..11 = jy(iy<2, ix), 12..20 = jz(iy, ix)
  constexpr int NR = V21 ? 21 : SROWS;
  int offq;
  {
    int comp, diy, dix;
    if (V21) {
      if (lane < 6) { comp = 0; diy = lane / 2; dix = lane % 2; }
      else if (lane < 12) { comp = 1; diy = (lane - 6) / 3; dix = (lane - 6) % 3; }
      else { comp = 2; diy = (lane - 12) / 3; dix = (lane - 12) % 3; }
    } else { comp = lane / 9; diy = (lane % 9) / 3; dix = lane % 3; }
    offq = comp * TILE_ELEMS + (diy - 1) * TW + (dix - 1);
  }
  const unsigned lt_mask = (1u << lane) - 1u;

  // software pipeline: the next batch's particle loads are in flight while this one computes
  long long i = start + warp * 32 + lane;
  double n_x = 0, n_y = 0, n_px = 0, n_py = 0, n_pz = 0, n_w = 0;
  if (i < end) {
    n_w = P.w[i]; n_x = P.x[0][i]; n_y = P.x[1][i];
    n_px = P.p[0][i]; n_py = P.p[1][i]; n_pz = P.p[2][i];
  }
  for (; i - lane < end; i += PUSH2D_THREADS) {
    const bool active = i < end;
    const double part_weight = n_w;
    double px_ = n_x - P.grid_min_local[0];
    double py_ = n_y - P.grid_min_local[1];
    double part_ux = n_px * P.ipart_mc;
    double part_uy = n_py * P.ipart_mc;
    double part_uz = n_pz * P.ipart_mc;
    {
      const long long in = i + PUSH2D_THREADS;
      if (in < end) {
        n_w = P.w[in]; n_x = P.x[0][in]; n_y = P.x[1][in];
        n_px = P.p[0][in]; n_py = P.p[1][in]; n_pz = P.p[2][in];
      }
    }
    int key = -1;       // cell key of a lane that takes part in the transposed reduction
    bool extras = false;  // queued: general loop (moved diagonally)
    bool edge = false;    // queued: only the part of the stencil outside the 3x3 core
    int dcx = 0, dcy = 0;
    double q_fxo = 0, q_fxn = 0, q_fyo = 0, q_fyn = 0, fjx = 0, fjy = 0, fjz = 0;
    if (active) {
      double root;
      gamma_root(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0, P.dtco2, root);
      px_ = px_ + part_ux * root;
      py_ = py_ + part_uy * root;
      const double cell_x_r = px_ * P.idx[0];
      const double cell_y_r = py_ * P.idx[1];
      const int cx1 = __double2int_rd(cell_x_r + 0.5) + 1;
      const int cy1 = __double2int_rd(cell_y_r + 0.5) + 1;
      // the gather reads cell1-2..cell1+1, the deposit writes cell1-2..cell1+2
      const bool fast = (cx1 - 2 >= ox) && (cx1 + 2 <= ox + TW - 1) && (cy1 - 2 >= oy) && (cy1 + 2 <= oy + TH - 1);
      if (!fast) {
        const int slot = atomicAdd(sSlowCount, 1);
        if (slot < SLOWCAP) sSlow[slot] = (int)i;
        else push_one<2>(P, i);
      } else {
        double gx[3], gy[3], hx[3], hy[3];
        const double fxo = (double)(cx1 - 1) - cell_x_r, fyo = (double)(cy1 - 1) - cell_y_r;
        tri(fxo, gx[0], gx[1], gx[2]);
        tri(fyo, gy[0], gy[1], gy[2]);
        int cx2 = __double2int_rd(cell_x_r);
        tri((double)cx2 - cell_x_r + 0.5, hx[0], hx[1], hx[2]);
        cx2 += 1;
        int cy2 = __double2int_rd(cell_y_r);
        tri((double)cy2 - cell_y_r + 0.5, hy[0], hy[1], hy[2]);
        cy2 += 1;
        // shared-tile offsets of (cell-1, cell-1)
        const int o11 = (cy1 - 1 - oy) * TW + (cx1 - 1 - ox);
        const int o21 = (cy1 - 1 - oy) * TW + (cx2 - 1 - ox);
        const int o12 = (cy2 - 1 - oy) * TW + (cx1 - 1 - ox);
        const int o22 = (cy2 - 1 - oy) * TW + (cx2 - 1 - ox);
        auto gat = [&](const double *F, int o, const double *wx, const double *wy) {
          double r0 = wx[0] * F[o] + wx[1] * F[o + 1] + wx[2] * F[o + 2];
          double r1 = wx[0] * F[o + TW] + wx[1] * F[o + TW + 1] + wx[2] * F[o + TW + 2];
          double r2 = wx[0] * F[o + 2 * TW] + wx[1] * F[o + 2 * TW + 1] + wx[2] * F[o + 2 * TW + 2];
          return wy[0] * r0 + wy[1] * r1 + wy[2] * r2;
        };
        const double ex_part = gat(sEx, o21, hx, gy);
        const double ey_part = gat(sEy, o12, gx, hy);
        const double ez_part = gat(sEz, o11, gx, gy);
        const double bx_part = gat(sBx, o12, gx, hy);
        const double by_part = gat(sBy, o21, hx, gy);
        const double bz_part = gat(sBz, o22, hx, hy);
        const double cmratio = P.cmratio;
        const double uxm = part_ux + cmratio * ex_part;
        const double uym = part_uy + cmratio * ey_part;
        const double uzm = part_uz + cmratio * ez_part;
        gamma_root(uxm * uxm + uym * uym + uzm * uzm + 1.0, P.ccmratio, root);
        const double taux = bx_part * root, tauy = by_part * root, tauz = bz_part * root;
        const double taux2 = taux * taux, tauy2 = tauy * tauy, tauz2 = tauz * tauz;
#ifdef EPB_FAST_MATH
        const double tau = rcp_ge1(1.0 + taux2 + tauy2 + tauz2);
#else
        const double tau = 1.0 / (1.0 + taux2 + tauy2 + tauz2);
#endif
        const double uxp = ((1.0 + taux2 - tauy2 - tauz2) * uxm +
                            2.0 * ((taux * tauy + tauz) * uym + (taux * tauz - tauy) * uzm)) * tau;
        const double uyp = ((1.0 - taux2 + tauy2 - tauz2) * uym +
                            2.0 * ((tauy * tauz + taux) * uzm + (tauy * taux - tauz) * uxm)) * tau;
        const double uzp = ((1.0 - taux2 - tauy2 + tauz2) * uzm +
                            2.0 * ((tauz * taux + tauy) * uxm + (tauz * tauy - taux) * uym)) * tau;
        part_ux = uxp + cmratio * ex_part;
        part_uy = uyp + cmratio * ey_part;
        part_uz = uzp + cmratio * ez_part;
        const double part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz;
#ifdef EPB_FAST_MATH
        const double igamma = rsqrt_ge1(part_u2 + 1.0);
#else
        const double igamma = 1.0 / sqrt(part_u2 + 1.0);
#endif
        root = P.dtco2 * igamma;
        const double delta_x = part_ux * root;
        const double delta_y = part_uy * root;
        const double part_vz = part_uz * c * igamma;
        px_ = px_ + delta_x;
        py_ = py_ + delta_y;
        {
          double pos[3] = {px_ + P.grid_min_local[0], py_ + P.grid_min_local[1], 0.0};
          double mom[3] = {P.part_mc * part_ux, P.part_mc * part_uy, P.part_mc * part_uz};
          const int dir = particle_bc<2>(P, pos, mom);
          P.x[0][i] = pos[0];
          P.x[1][i] = pos[1];
          P.p[0][i] = mom[0];
          P.p[1][i] = mom[1];
          P.p[2][i] = mom[2];
          if (dir >= 0) outbox_put(P, i, dir);
        }
        if (P.deposit) {
          px_ = px_ + delta_x;
          py_ = py_ + delta_y;
          const double cxr = px_ * P.idx[0], cyr = py_ * P.idx[1];
          const int cx3 = __double2int_rd(cxr + 0.5), cy3 = __double2int_rd(cyr + 0.5);
          const double fxn = (double)cx3 - cxr, fyn = (double)cy3 - cyr;
          dcx = cx3 + 1 - cx1;
          dcy = cy3 + 1 - cy1;
          const double fcx = P.kfc[0] * part_weight;
          const double fcy = P.kfc[1] * part_weight;
          const double fcz = P.kfc[2] * part_weight;
          fjx = fcx * P.part_q;
          fjy = fcy * P.part_q;
          fjz = fcz * P.part_q * part_vz;
          const int k = (cy1 - oy) * TW + (cx1 - ox);
          key = k;
          q_fxo = fxo; q_fxn = fxn; q_fyo = fyo; q_fyn = fyn;
          if (dcx != 0 && dcy != 0) {
            extras = true;  // moved diagonally (rare): general loop; excluded from the reduction below
          } else {
            // Nearest cell unchanged, or moved by one cell along one axis.  New weights on the 3x3
            // core around the old cell (particles.F90:521-538 with the shift by dcell); the running
            // jx / jy prefixes of a particle that moved towards -x / -y enter the core with the
            // value of the outer column / row (hxa, hya).  What lies outside the core (5 or 8
            // values) is queued and deposited by drain_edge.
            double wm, w0, wp;
            tri(fxn, wm, w0, wp);
            hx[0] = (dcx == 0 ? wm : dcx > 0 ? 0.0 : w0) - gx[0];
            hx[1] = (dcx == 0 ? w0 : dcx > 0 ? wm : wp) - gx[1];
            hx[2] = (dcx == 0 ? wp : dcx > 0 ? w0 : 0.0) - gx[2];
            const double hxa = hx[0] + (dcx < 0 ? wm : 0.0);
            tri(fyn, wm, w0, wp);
            hy[0] = (dcy == 0 ? wm : dcy > 0 ? 0.0 : w0) - gy[0];
            hy[1] = (dcy == 0 ? w0 : dcy > 0 ? wm : wp) - gy[1];
            hy[2] = (dcy == 0 ? wp : dcy > 0 ? w0 : 0.0) - gy[2];
            const double hya = hy[0] + (dcy < 0 ? wm : 0.0);
            double xfac1[3], jyh[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int ix = 0; ix < 3; ix++) xfac1[ix] = gx[ix] + 0.5 * hx[ix];
            double *col = S + lane;
#pragma unroll
            for (int iy = 0; iy < 3; iy++) {
              const double yfac1 = gy[iy] + 0.5 * hy[iy];
              const double yfac2 = third * hy[iy] + 0.5 * gy[iy];
              const double hyw = iy == 0 ? hya : hy[iy];
              double jxh = 0.0;
#pragma unroll
              for (int ix = 0; ix < 3; ix++) {
                const double wx = (ix == 0 ? hxa : hx[ix]) * yfac1;
                const double wy = hyw * xfac1[ix];
                const double wz = gx[ix] * yfac1 + hx[ix] * yfac2;
                jxh = jxh - fjx * wx;
                jyh[ix] = jyh[ix] - fjy * wy;
                if (!V21 || ix < 2) col[(V21 ? iy * 2 + ix : iy * 3 + ix) * SPITCH] = jxh;
                if (!V21 || iy < 2) col[(V21 ? 6 + iy * 3 + ix : 9 + iy * 3 + ix) * SPITCH] = jyh[ix];
                col[((V21 ? 12 : 18) + iy * 3 + ix) * SPITCH] = fjz * wz;
              }
            }
            if ((dcx | dcy) != 0) { edge = true; }
          }
        }
      }
    }
    if (!P.deposit) continue;
    // ---- queue the particles with a wider stencil -------------------------------------
    const unsigned em = __ballot_sync(FULL, extras || edge);
    if (em) {
      const int ne = __popc(em);
      if (qcount + ne > QCAP) {
        __syncwarp();
        drain_extras(P, sJ, Qd, Qk, qcount, lane);
        __syncwarp();
        qcount = 0;
      }
      if (extras || edge) {
        const int slot = qcount + __popc(em & lt_mask);
        // bit 14: edge entry; bit 16: the core already holds the third jx column / jy row
        Qk[slot] = key | ((dcx + 1) << 10) | ((dcy + 1) << 12) | (edge ? (1 << 14) | (V21 ? 0 : 1 << 16) : 0);
        Qd[0 * QCAP + slot] = q_fxo; Qd[1 * QCAP + slot] = q_fxn;
        Qd[2 * QCAP + slot] = q_fyo; Qd[3 * QCAP + slot] = q_fyn;
        Qd[4 * QCAP + slot] = fjx; Qd[5 * QCAP + slot] = fjy; Qd[6 * QCAP + slot] = fjz;
        if (extras) key = -1;
      }
      qcount += ne;
    }
    // ---- transposed reduction: one shared update per (cell key, stencil value) ------------
    __syncwarp();
    unsigned rest = __ballot_sync(FULL, key >= 0);
    if (rest) {
      const int k0 = __shfl_sync(FULL, key, __ffs(rest) - 1);
      const unsigned m0 = __ballot_sync(FULL, key == k0);
      rest &= ~m0;
      int k1 = 0;
      unsigned m1 = 0;
      if (rest) {
        k1 = __shfl_sync(FULL, key, __ffs(rest) - 1);
        m1 = __ballot_sync(FULL, key == k1);
        rest &= ~m1;
      }
      if (lane < NR) {
        const double *row = S + lane * SPITCH;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const double v = row[j];
          if ((m0 >> j) & 1u) a0 += v;
          else if ((m1 >> j) & 1u) a1 += v;
        }
        smem_add(&sJ[offq + k0], a0);
        if (m1) smem_add(&sJ[offq + k1], a1);
      }
      while (rest) {  // lanes in further cells (stale sort): one update per lane and value
        const int j = __ffs(rest) - 1;
        rest &= rest - 1;
        const int kj = __shfl_sync(FULL, key, j);
        if (lane < NR) smem_add(&sJ[offq + kj], S[lane * SPITCH + j]);
      }
    }
    __syncwarp();
  }
  if (qcount) {
    __syncwarp();
    drain_extras(P, sJ, Qd, Qk, qcount, lane);
  }
  __syncthreads();
  {
    int ns = *sSlowCount;
    if (ns > SLOWCAP) ns = SLOWCAP;
    for (int q = tid; q < ns; q += PUSH2D_THREADS) push_one<2>(P, sSlow[q]);
  }
  for (int q = tid; q < TILE_ELEMS; q += PUSH2D_THREADS) {
    const int lx = q % TW, ly = q / TW;
    const int cx = ox + lx, cy = oy + ly;
    const bool ok = (cx >= 1 - NG) && (cx <= P.n[0] + NG) && (cy >= 1 - NG) && (cy <= P.n[1] + NG);
    if (!ok) continue;
    const size_t o = gofs<2>(P, cx, cy, 1);
#pragma unroll
    for (int f = 0; f < 3; f++) {
      const double val = sJ[f * TILE_ELEMS + q];
      if (val != 0.0) atomicAdd(P.j[f] + o, val);
    }
  }
}


// ---------------------------------------------------------------------------
// Cell-owner 2D kernel (layout 1, EPB_PUSH_VARIANT=2)
// ---------------------------------------------------------------------------
// One CTA per 16xCTY-cell tile, one warp per 16x2-cell group, ONE LANE PER CELL (a half-warp
// is one row of the tile; the shared row pitch of 32 doubles then keeps the 64-bit gather
// loads of a half-warp conflict-free up to the +-1 cell spread of the staggered stencils).  The sort
// (k_scatter_il) stores the particles of a group interleaved by their rank within the cell,
// so round r of a warp -- the r-th particle of each of its 32 cells -- is one coalesced load.
// A lane keeps the 21 non-cancelling deposit sums of its cell's 3x3 stencil in registers for
// the whole tile (raw sums; the running prefixes of particles.F90:563-571 are linear, so they
// are applied once per cell at the end) and touches shared memory for the deposit only in the
// final flush: no per-particle or per-batch shared FP64 updates, no transposition scratch.
// The sort key is the cell the particle will be gathered in (sort.cu, predict), so right
// after a sort every particle sits with the lane that owns its stencil.  Particles that are
// not "regular" for their lane -- nearest cell changed during the step (wider stencil), or
// stale order between sorts -- are queued per warp and deposited densely with the
// reference's general loop and shared-memory updates (drain_extras); because a round spans 32
// different cells those updates rarely collide.  Correctness never depends on the order.
// CTY = tile height in cells (16 or 8; width is always 16), MINB = CTAs per SM the register
// budget is sized for: <16,2> 128 registers, <8,3> 168, <16,1> 255.
constexpr int CPITCH = 32;  // shared row pitch (doubles) of the cell-owner kernel; 22 columns used
template <int CTY>
constexpr size_t pushcell_smem() {
  return sizeof(double) * ((size_t)9 * CPITCH * (CTY + 2 * HALO) + (size_t)(CTY / 2) * QDBL * QCAP) +
         sizeof(int) * ((size_t)(CTY / 2) * QCAP + SLOWCAP + 2);
}

template <int CTY, int MINB>
__global__ void __launch_bounds__(CTY * 16, MINB) push_cell_2d(const __grid_constant__ PushParams P) {
  constexpr int T2Y = CTY, TH = CTY + 2 * HALO, TW = CPITCH, TWU = T2X + 2 * HALO, TILE_ELEMS = TW * TH;
  constexpr int PUSH2D_THREADS = CTY * 16, PUSH2D_WARPS = CTY / 2;
  extern __shared__ double sm[];
  double *sF = sm;                                   // [6][TH][TW]
  double *sJ = sF + 6 * TILE_ELEMS;                  // [3][TH][TW]
  double *sQd_all = sJ + 3 * TILE_ELEMS;
  int *sQk_all = reinterpret_cast<int *>(sQd_all + PUSH2D_WARPS * QDBL * QCAP);
  int *sSlow = sQk_all + PUSH2D_WARPS * QCAP;
  int *sSlowCount = sSlow + SLOWCAP;
  const int tile = blockIdx.x;
  const int ttx = tile % P.tg.nt[0], tty = tile / P.tg.nt[0];
  const int ox = ttx * T2X + 1 - HALO;  // cell index of shared column 0
  const int oy = tty * T2Y + 1 - HALO;
  const int *cs = P.cell_start + (size_t)tile * (T2X * T2Y);
  const long long clip = P.n_sorted_clip;
  {
    const long long start = cs[0], end = cs[T2X * T2Y];
    if (start >= end || start >= clip) return;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) *sSlowCount = 0;
  for (int q = tid; q < TILE_ELEMS; q += PUSH2D_THREADS) {
    const int lx = q % TW, ly = q / TW;
    const int cx = ox + lx, cy = oy + ly;
    const bool ok = (lx < TWU) && (cx >= 1 - NG) && (cx <= P.n[0] + NG) && (cy >= 1 - NG) && (cy <= P.n[1] + NG);
    const size_t o = ok ? gofs<2>(P, cx, cy, 1) : 0;
#pragma unroll
    for (int f = 0; f < 3; f++) {
      sF[f * TILE_ELEMS + q] = ok ? __ldg(P.e[f] + o) : 0.0;
      sF[(3 + f) * TILE_ELEMS + q] = ok ? __ldg(P.b[f] + o) : 0.0;
      sJ[f * TILE_ELEMS + q] = 0.0;
    }
  }
  __syncthreads();

  const double c = EPB_C;
  const double third = P.third;
  const double *sEx = sF, *sEy = sF + TILE_ELEMS, *sEz = sF + 2 * TILE_ELEMS;
  const double *sBx = sF + 3 * TILE_ELEMS, *sBy = sF + 4 * TILE_ELEMS, *sBz = sF + 5 * TILE_ELEMS;
  double *Qd = sQd_all + warp * QDBL * QCAP;
  int *Qk = sQk_all + warp * QCAP;
  int qcount = 0;  // warp-uniform
  const unsigned lt_mask = (1u << lane) - 1u;

  // this lane's cell (1-based cell indices as in the reference) and its particle count
  const int hcx = ttx * T2X + (lane & 15) + 1;
  const int hcy = tty * T2Y + warp * 2 + (lane >> 4) + 1;
  const int my_key = tile * (T2X * T2Y) + warp * 32 + lane;
  int stay = 0;  // particles of this lane's cell that stay in it (their rank in the next order)
  const int my_start = cs[warp * 32 + lane];
  const int my_cnt = cs[warp * 32 + lane + 1] - my_start;
  const int maxcnt = __reduce_max_sync(FULL, my_cnt);
  long long rbase = __shfl_sync(FULL, my_start, 0);

  // raw deposit sums of this lane's cell: AX[iy][ix<2], AY[iy<2][ix], AZ[iy][ix]
  double AX[3][2], AY[2][3], AZ[3][3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int b = 0; b < 3; b++) {
      AZ[a][b] = 0.0;
      if (b < 2) AX[a][b] = 0.0;
      if (a < 2) AY[a][b] = 0.0;
    }
  }

  // software pipeline: the next round's particle loads are in flight while this one computes, and
  // (fused gather after a sort) the source index of the round after that is already being fetched
  const int *perm = P.perm;
  long long i, i2;      // slot of the current/next round, slot of the round after
  // source index of slot i2 (perm[i2], or i2 itself without a pending permutation).  Kept as the raw
  // 32-bit value: widening it here would make the warp wait for the perm load in the round that issues it
  int si2 = 0;
  bool act, act2;
  {
    const bool na = my_cnt > 0;
    const unsigned bal = __ballot_sync(FULL, na);
    i = rbase + __popc(bal & lt_mask);
    rbase += __popc(bal);
    act = na && i < clip;
  }
  double n_x = 0, n_y = 0, n_px = 0, n_py = 0, n_pz = 0, n_w = 0;
  if (act) {
    const long long si = perm ? (long long)perm[i] : i;
    n_w = P.ws[si]; n_x = P.xs[0][si]; n_y = P.xs[1][si];
    n_px = P.ps[0][si]; n_py = P.ps[1][si]; n_pz = P.ps[2][si];
  }
  {
    const bool na = my_cnt > 1;
    const unsigned bal = __ballot_sync(FULL, na);
    i2 = rbase + __popc(bal & lt_mask);
    rbase += __popc(bal);
    act2 = na && i2 < clip;
    if (act2) si2 = perm ? perm[i2] : (int)i2;
  }
  for (int r = 0; r < maxcnt; r++) {
    const bool active = act;
    const long long ci = i;
    const double part_weight = n_w;
    const double raw_x = n_x, raw_y = n_y, raw_px = n_px, raw_py = n_py, raw_pz = n_pz;
    double px_ = n_x - P.grid_min_local[0];
    double py_ = n_y - P.grid_min_local[1];
    double part_ux = n_px * P.ipart_mc;
    double part_uy = n_py * P.ipart_mc;
    double part_uz = n_pz * P.ipart_mc;
    {
      // loads of round r+1 (its source index arrived during the previous round)
      i = i2;
      act = act2;
      if (act) {
        const long long sl = si2;
        n_w = P.ws[sl]; n_x = P.xs[0][sl]; n_y = P.xs[1][sl];
        n_px = P.ps[0][sl]; n_py = P.ps[1][sl]; n_pz = P.ps[2][sl];
      }
      // slot and source index of round r+2
      const bool na = my_cnt > r + 2;
      const unsigned bal = __ballot_sync(FULL, na);
      i2 = rbase + __popc(bal & lt_mask);
      rbase += __popc(bal);
      act2 = na && i2 < clip;
      if (act2) si2 = perm ? perm[i2] : (int)i2;
    }
    bool extras = false;
    int key = 0, dcx = 0, dcy = 0;
    double q_fxo = 0, q_fxn = 0, q_fyo = 0, q_fyn = 0, fjx = 0, fjy = 0, fjz = 0;
    if (active) {
      double root;
      gamma_root(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0, P.dtco2, root);
      px_ = px_ + part_ux * root;
      py_ = py_ + part_uy * root;
      const double cell_x_r = px_ * P.idx[0];
      const double cell_y_r = py_ * P.idx[1];
      const int cx1 = __double2int_rd(cell_x_r + 0.5) + 1;
      const int cy1 = __double2int_rd(cell_y_r + 0.5) + 1;
      // the gather reads cell1-2..cell1+1, the deposit writes cell1-2..cell1+2
      const bool fast = (cx1 - 2 >= ox) && (cx1 + 2 <= ox + TWU - 1) && (cy1 - 2 >= oy) && (cy1 + 2 <= oy + TH - 1);
      if (!fast) {
        if (perm) {  // push_one works in place on the destination slot
          P.x[0][ci] = raw_x; P.x[1][ci] = raw_y;
          P.p[0][ci] = raw_px; P.p[1][ci] = raw_py; P.p[2][ci] = raw_pz;
          P.w[ci] = part_weight;
        }
        const int slot = atomicAdd(sSlowCount, 1);
        if (slot < SLOWCAP) sSlow[slot] = (int)ci;
        else push_one<2>(P, ci);
      } else {
        double gx[3], gy[3], hx[3], hy[3];
        const double fxo = (double)(cx1 - 1) - cell_x_r, fyo = (double)(cy1 - 1) - cell_y_r;
        tri(fxo, gx[0], gx[1], gx[2]);
        tri(fyo, gy[0], gy[1], gy[2]);
        int cx2 = __double2int_rd(cell_x_r);
        tri((double)cx2 - cell_x_r + 0.5, hx[0], hx[1], hx[2]);
        cx2 += 1;
        int cy2 = __double2int_rd(cell_y_r);
        tri((double)cy2 - cell_y_r + 0.5, hy[0], hy[1], hy[2]);
        cy2 += 1;
        // shared-tile offsets of (cell-1, cell-1)
        const int o11 = (cy1 - 1 - oy) * TW + (cx1 - 1 - ox);
        const int o21 = (cy1 - 1 - oy) * TW + (cx2 - 1 - ox);
        const int o12 = (cy2 - 1 - oy) * TW + (cx1 - 1 - ox);
        const int o22 = (cy2 - 1 - oy) * TW + (cx2 - 1 - ox);
        auto gat = [&](const double *F, int o, const double *wx, const double *wy) {
          double r0 = wx[0] * F[o] + wx[1] * F[o + 1] + wx[2] * F[o + 2];
          double r1 = wx[0] * F[o + TW] + wx[1] * F[o + TW + 1] + wx[2] * F[o + TW + 2];
          double r2 = wx[0] * F[o + 2 * TW] + wx[1] * F[o + 2 * TW + 1] + wx[2] * F[o + 2 * TW + 2];
          return wy[0] * r0 + wy[1] * r1 + wy[2] * r2;
        };
        const double ex_part = gat(sEx, o21, hx, gy);
        const double ey_part = gat(sEy, o12, gx, hy);
        const double ez_part = gat(sEz, o11, gx, gy);
        const double bx_part = gat(sBx, o12, gx, hy);
        const double by_part = gat(sBy, o21, hx, gy);
        const double bz_part = gat(sBz, o22, hx, hy);
        const double cmratio = P.cmratio;
        const double uxm = part_ux + cmratio * ex_part;
        const double uym = part_uy + cmratio * ey_part;
        const double uzm = part_uz + cmratio * ez_part;
        gamma_root(uxm * uxm + uym * uym + uzm * uzm + 1.0, P.ccmratio, root);
        const double taux = bx_part * root, tauy = by_part * root, tauz = bz_part * root;
        const double taux2 = taux * taux, tauy2 = tauy * tauy, tauz2 = tauz * tauz;
#ifdef EPB_FAST_MATH
        const double tau = rcp_ge1(1.0 + taux2 + tauy2 + tauz2);
#else
        const double tau = 1.0 / (1.0 + taux2 + tauy2 + tauz2);
#endif
        const double uxp = ((1.0 + taux2 - tauy2 - tauz2) * uxm +
                            2.0 * ((taux * tauy + tauz) * uym + (taux * tauz - tauy) * uzm)) * tau;
        const double uyp = ((1.0 - taux2 + tauy2 - tauz2) * uym +
                            2.0 * ((tauy * tauz + taux) * uzm + (tauy * taux - tauz) * uxm)) * tau;
        const double uzp = ((1.0 - taux2 - tauy2 + tauz2) * uzm +
                            2.0 * ((tauz * taux + tauy) * uxm + (tauz * tauy - taux) * uym)) * tau;
        part_ux = uxp + cmratio * ex_part;
        part_uy = uyp + cmratio * ey_part;
        part_uz = uzp + cmratio * ez_part;
        const double part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz;
#ifdef EPB_FAST_MATH
        const double igamma = rsqrt_ge1(part_u2 + 1.0);
#else
        const double igamma = 1.0 / sqrt(part_u2 + 1.0);
#endif
        root = P.dtco2 * igamma;
        const double delta_x = part_ux * root;
        const double delta_y = part_uy * root;
        const double part_vz = part_uz * c * igamma;
        px_ = px_ + delta_x;
        py_ = py_ + delta_y;
        int dir;
        {
          double pos[3] = {px_ + P.grid_min_local[0], py_ + P.grid_min_local[1], 0.0};
          double mom[3] = {P.part_mc * part_ux, P.part_mc * part_uy, P.part_mc * part_uz};
          dir = particle_bc<2>(P, pos, mom);
          P.x[0][ci] = pos[0];
          P.x[1][ci] = pos[1];
          P.p[0][ci] = mom[0];
          P.p[1][ci] = mom[1];
          P.p[2][ci] = mom[2];
          if (perm) P.w[ci] = part_weight;
          if (dir >= 0) outbox_put(P, ci, dir);
        }
        if (P.deposit) {
          bool emit_do = false, emit_arr = false;
          int emit_rank = 0;
          px_ = px_ + delta_x;
          py_ = py_ + delta_y;
          const double cxr = px_ * P.idx[0], cyr = py_ * P.idx[1];
          const int cx3 = __double2int_rd(cxr + 0.5), cy3 = __double2int_rd(cyr + 0.5);
          const double fxn = (double)cx3 - cxr, fyn = (double)cy3 - cyr;
          dcx = cx3 + 1 - cx1;
          dcy = cy3 + 1 - cy1;
          const double fcx = P.kfc[0] * part_weight;
          const double fcy = P.kfc[1] * part_weight;
          const double fcz = P.kfc[2] * part_weight;
          fjx = fcx * P.part_q;
          fjy = fcy * P.part_q;
          fjz = fcz * P.part_q * part_vz;
          if (P.emit && dir < 0 && cx3 >= 0 && cx3 < P.n[0] && cy3 >= 0 && cy3 < P.n[1]) {
            // record for the next sort: (cx3, cy3) is the cell the next push gathers in
            const int nkey = ((cy3 / T2Y) * P.tg.nt[0] + (cx3 >> 4)) * (T2X * T2Y) + (cy3 % T2Y) * T2X + (cx3 & 15);
            P.key_out[ci] = nkey;
            // the arrival counter's reply is only consumed after the deposit arithmetic below, so the
            // round trip of the global atomic is hidden instead of stalling the warp here
            emit_do = true;
            emit_arr = nkey != my_key;
            emit_rank = emit_arr ? atomicAdd(&P.arr_cnt[nkey], 1) : stay++;
          }
          key = (cy1 - oy) * TW + (cx1 - ox);
          q_fxo = fxo; q_fxn = fxn; q_fyo = fyo; q_fyn = fyn;
          if (cx1 != hcx || cy1 != hcy || (dcx != 0 && dcy != 0)) {
            extras = true;  // not this lane's cell (stale order), or moved diagonally
            if ((dcx | dcy) == 0) key |= 1 << 15;  // 3x3 stencil: 21-update drain
          } else {
            // This lane's own cell, nearest cell unchanged or moved by one cell along one axis.
            // New weights on the 3x3 core (particles.F90:521-538 with the shift by dcell); the
            // part of a moved particle's stencil outside the core is queued (drain_edge).
            double wm, w0, wp;
            tri(fxn, wm, w0, wp);
            hx[0] = (dcx == 0 ? wm : dcx > 0 ? 0.0 : w0) - gx[0];
            hx[1] = (dcx == 0 ? w0 : dcx > 0 ? wm : wp) - gx[1];
            hx[2] = (dcx == 0 ? wp : dcx > 0 ? w0 : 0.0) - gx[2];
            const double hxa = hx[0] + (dcx < 0 ? wm : 0.0);  // running jx prefix enters the core with column -2
            tri(fyn, wm, w0, wp);
            hy[0] = (dcy == 0 ? wm : dcy > 0 ? 0.0 : w0) - gy[0];
            hy[1] = (dcy == 0 ? w0 : dcy > 0 ? wm : wp) - gy[1];
            hy[2] = (dcy == 0 ? wp : dcy > 0 ? w0 : 0.0) - gy[2];
            const double hya = hy[0] + (dcy < 0 ? wm : 0.0);
            double xfac1[3], yfac1[3], yfac2[3];
#pragma unroll
            for (int q = 0; q < 3; q++) {
              xfac1[q] = gx[q] + 0.5 * hx[q];
              yfac1[q] = gy[q] + 0.5 * hy[q];
              yfac2[q] = third * hy[q] + 0.5 * gy[q];
            }
            const double fhx0 = fjx * hxa, fhx1 = fjx * hx[1];
            const double fhy0 = fjy * hya, fhy1 = fjy * hy[1];
#pragma unroll
            for (int iy = 0; iy < 3; iy++) {
              AX[iy][0] += fhx0 * yfac1[iy];
              AX[iy][1] += fhx1 * yfac1[iy];
            }
#pragma unroll
            for (int ix = 0; ix < 3; ix++) {
              AY[0][ix] += fhy0 * xfac1[ix];
              AY[1][ix] += fhy1 * xfac1[ix];
            }
#pragma unroll
            for (int ix = 0; ix < 3; ix++) {
              const double zg = fjz * gx[ix], zh = fjz * hx[ix];
#pragma unroll
              for (int iy = 0; iy < 3; iy++) AZ[iy][ix] += zg * yfac1[iy] + zh * yfac2[iy];
            }
            if ((dcx | dcy) != 0) { extras = true; key |= 1 << 14; }
          }
          if (emit_do) P.rank_out[ci] = emit_arr ? (emit_rank | EPB_RANK_ARRIVAL) : emit_rank;
        }
      }
    }
    if (!P.deposit) continue;
    // ---- queue the particles that are not regular for their lane ---------------------------
    const unsigned em = __ballot_sync(FULL, extras);
    if (em) {
      const int ne = __popc(em);
      if (qcount + ne > QCAP) {
        __syncwarp();
        drain_extras(P, sJ, Qd, Qk, qcount, lane, TILE_ELEMS, TW);
        __syncwarp();
        qcount = 0;
      }
      if (extras) {
        const int slot = qcount + __popc(em & lt_mask);
        Qk[slot] = key | ((dcx + 1) << 10) | ((dcy + 1) << 12);
        Qd[0 * QCAP + slot] = q_fxo; Qd[1 * QCAP + slot] = q_fxn;
        Qd[2 * QCAP + slot] = q_fyo; Qd[3 * QCAP + slot] = q_fyn;
        Qd[4 * QCAP + slot] = fjx; Qd[5 * QCAP + slot] = fjy; Qd[6 * QCAP + slot] = fjz;
      }
      qcount += ne;
    }
  }
  if (qcount) {
    __syncwarp();
    drain_extras(P, sJ, Qd, Qk, qcount, lane, TILE_ELEMS, TW);
  }
  if (P.emit) P.stay_cnt[my_key] = stay;
  // ---- flush this lane's cell sums: prefixes of particles.F90:563-571, one update per point ----
  if (P.deposit && my_cnt > 0) {
    const int hb = (hcy - oy) * TW + (hcx - ox);
#pragma unroll
    for (int iy = 0; iy < 3; iy++) {
      const double v0 = -AX[iy][0];
      const double v1 = v0 - AX[iy][1];
      smem_add(&sJ[hb + (iy - 1) * TW - 1], v0);
      smem_add(&sJ[hb + (iy - 1) * TW], v1);
    }
#pragma unroll
    for (int ix = 0; ix < 3; ix++) {
      const double v0 = -AY[0][ix];
      const double v1 = v0 - AY[1][ix];
      smem_add(&sJ[TILE_ELEMS + hb - TW + (ix - 1)], v0);
      smem_add(&sJ[TILE_ELEMS + hb + (ix - 1)], v1);
    }
#pragma unroll
    for (int iy = 0; iy < 3; iy++)
#pragma unroll
      for (int ix = 0; ix < 3; ix++) smem_add(&sJ[2 * TILE_ELEMS + hb + (iy - 1) * TW + (ix - 1)], AZ[iy][ix]);
  }
  __syncthreads();
  {
    int ns = *sSlowCount;
    if (ns > SLOWCAP) ns = SLOWCAP;
    for (int q = tid; q < ns; q += PUSH2D_THREADS) push_one<2>(P, sSlow[q]);
  }
  for (int q = tid; q < TILE_ELEMS; q += PUSH2D_THREADS) {
    const int lx = q % TW, ly = q / TW;
    const int cx = ox + lx, cy = oy + ly;
    const bool ok = (lx < TWU) && (cx >= 1 - NG) && (cx <= P.n[0] + NG) && (cy >= 1 - NG) && (cy <= P.n[1] + NG);
    if (!ok) continue;
    const size_t o = gofs<2>(P, cx, cy, 1);
#pragma unroll
    for (int f = 0; f < 3; f++) {
      const double val = sJ[f * TILE_ELEMS + q];
      if (val != 0.0) atomicAdd(P.j[f] + o, val);
    }
  }
}


// ---------------------------------------------------------------------------
// Slot-column 2D kernel (layout 2, the default): push_cell_2d's lane-per-cell register deposit on a
// particle store that needs no sort
// ---------------------------------------------------------------------------
// Every cell owns a column of R slots (epb_internal.h, slots.cu); row r of a warp's 32 columns is 32
// consecutive, 256-byte aligned doubles, so round r is one aligned coalesced access with no index
// arithmetic, no permutation and no prefix sums.  The lane walks its column; a particle whose next gather
// cell is still the lane's cell is written back into the column at the lane's write cursor (in-place
// compaction: the cursor never passes the read row), all others -- ~4 % per step in the C2 plasma -- leave
// through the mover buffer M:
//   * movers: pushed, next gather cell differs (or a boundary condition touched them): M entry, flag 0
//   * leavers: pushed, classified for another rank by particle_bc: M entry, flag 1, M index in the outbox
//   * particles whose stencil is not inside the tile's halo (only arrivals the prediction got wrong, or
//     columns left stale by a full M): unpushed M entry, flag 2, pushed by push_generic_m
//   * deleted particles (open boundary, beyond x_min_outer): dropped
// k_deliver (slots.cu) then inserts the flag-0 entries into their new columns.  A mover that finds M full
// simply stays where it is and is deposited through the general path next step: placement is an
// optimisation, never a correctness condition.  Deposit exactly as push_cell_2d (21 register sums per
// cell, core/edge split for one-axis movers, drain_extras for the rest).
constexpr int SLOT_LK = 16;   // arrivals one column takes from its group's inbox per step (more: through M)
template <int CTY>
constexpr size_t pushslots_smem() {
  return sizeof(double) * ((size_t)9 * CPITCH * (CTY + 2 * HALO) + (size_t)(CTY / 2) * QDBL * QCAP) +
         sizeof(int) * ((size_t)(CTY / 2) * QCAP + 2 + (size_t)(CTY / 2) * 32) + (size_t)(CTY / 2) * 32 * SLOT_LK +
         sizeof(double) * (size_t)(CTY / 2) * 6 * 32;
}

template <int CTY, int MINB, bool RB, bool HC = false>
__global__ void __launch_bounds__(CTY * 16, MINB) push_slots_2d(const __grid_constant__ PushParams P) {
  constexpr int T2Y = CTY, TH = CTY + 2 * HALO, TW = CPITCH, TWU = T2X + 2 * HALO, TILE_ELEMS = TW * TH;
  constexpr int PUSH2D_THREADS = CTY * 16, PUSH2D_WARPS = CTY / 2;
  extern __shared__ double sm[];
  double *sF = sm;                                   // [6][TH][TW]
  double *sJ = sF + 6 * TILE_ELEMS;                  // [3][TH][TW]
  double *sQd_all = sJ + 3 * TILE_ELEMS;
  int *sQk_all = reinterpret_cast<int *>(sQd_all + PUSH2D_WARPS * QDBL * QCAP);
  int *sAcnt_all = sQk_all + PUSH2D_WARPS * QCAP + 2;                               // [warp][32] arrivals per lane
  unsigned char *sAlist_all = reinterpret_cast<unsigned char *>(sAcnt_all + PUSH2D_WARPS * 32);  // [warp][32][SLOT_LK]
  double *sPend_all = reinterpret_cast<double *>(sAlist_all + PUSH2D_WARPS * 32 * SLOT_LK);       // [warp][6][32]
  const int tile = blockIdx.x;
  const int ttx = tile % P.tg.nt[0], tty = tile / P.tg.nt[0];
  const int ox = ttx * T2X + 1 - HALO;  // cell index of shared column 0
  const int oy = tty * T2Y + 1 - HALO;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int my_key = tile * (T2X * T2Y) + warp * 32 + lane;
  int my_cnt = P.cnt[my_key];
  if (my_cnt > P.R) my_cnt = P.R;
  // entries waiting in this warp's group inbox (filled by the previous push)
  const int grp = my_key >> 5;
  int inA = P.ic_in ? P.ic_in[grp] : 0;
  if (inA > P.IC) inA = P.IC;
  if (!__syncthreads_or(my_cnt > 0 || inA > 0)) return;
  for (int q = tid; q < TILE_ELEMS; q += PUSH2D_THREADS) {
    const int lx = q % TW, ly = q / TW;
    const int cx = ox + lx, cy = oy + ly;
    const bool ok = (lx < TWU) && (cx >= 1 - NG) && (cx <= P.n[0] + NG) && (cy >= 1 - NG) && (cy <= P.n[1] + NG);
    const size_t o = ok ? gofs<2>(P, cx, cy, 1) : 0;
#pragma unroll
    for (int f = 0; f < 3; f++) {
      sF[f * TILE_ELEMS + q] = ok ? __ldg(P.e[f] + o) : 0.0;
      sF[(3 + f) * TILE_ELEMS + q] = ok ? __ldg(P.b[f] + o) : 0.0;
      sJ[f * TILE_ELEMS + q] = 0.0;
    }
  }
  __syncthreads();

  const double c = EPB_C;
  const double third = P.third;
  const double *sEx = sF, *sEy = sF + TILE_ELEMS, *sEz = sF + 2 * TILE_ELEMS;
  const double *sBx = sF + 3 * TILE_ELEMS, *sBy = sF + 4 * TILE_ELEMS, *sBz = sF + 5 * TILE_ELEMS;
  double *Qd = sQd_all + warp * QDBL * QCAP;
  int *Qk = sQk_all + warp * QCAP;
  int qcount = 0;  // warp-uniform
  const unsigned lt_mask = (1u << lane) - 1u;

  // ---- this warp's inbox: which entries are for which lane (entry = 8 doubles, [6] = destination lane) ----
  int *sAcnt = sAcnt_all + warp * 32;
  unsigned char *sAlist = sAlist_all + warp * 32 * SLOT_LK;
  const double *ibg = P.ib_in ? P.ib_in + (size_t)grp * (size_t)P.IC * 8 : nullptr;
  sAcnt[lane] = 0;
  __syncwarp();
  for (int j0 = 0; j0 < inA; j0 += 32) {
    const int j = j0 + lane;
    if (j < inA) {
      const int ml = (int)__double_as_longlong(ibg[(size_t)j * 8 + 6]) & 31;
      const int pos = atomicAdd(&sAcnt[ml], 1);
      if (pos < SLOT_LK) {
        sAlist[ml * SLOT_LK + pos] = (unsigned char)j;
      } else {
        // more arrivals than one column takes per step: the entry goes on through the mover buffer (flag 0)
        const int m = atomicAdd(P.mcount, 1);
        if (m < P.mcap) {
          const double *e = ibg + (size_t)j * 8;
          P.mx[0][m] = e[0]; P.mx[1][m] = e[1];
          P.mp[0][m] = e[2]; P.mp[1][m] = e[3]; P.mp[2][m] = e[4];
          P.mw[m] = e[5];
          P.mflag[m] = 0;
        } else {
          atomicOr(P.err, 1);
        }
      }
    }
  }
  __syncwarp();
  int my_a = sAcnt[lane];
  if (my_a > SLOT_LK) my_a = SLOT_LK;
  const int my_tot = my_cnt + my_a;

  // this lane's cell (1-based cell indices as in the reference)
  const int hcx = ttx * T2X + (lane & 15) + 1;
  const int hcy = tty * T2Y + warp * 2 + (lane >> 4) + 1;
  const int maxcnt = __reduce_max_sync(FULL, my_tot);
  // row 0 of this lane's column: rows are blocks of 6 x 32 doubles (x y px py pz w), so every access below is
  // this pointer + row * ROWD + an immediate component offset
  // (RB = false: one plane per component, rows of 32 doubles; the component offsets are run-time strides then)
  constexpr int ROWD = RB ? 6 * 32 : 32;
  const size_t CS = RB ? (size_t)32 : (size_t)(P.x[1] - P.x[0]);
  const size_t OX = 0, OY = CS, OPX = 2 * CS, OPY = 3 * CS, OPZ = 4 * CS, OW = 5 * CS;
  double *const col = P.x[0] + ((size_t)(my_key >> 5) * (size_t)P.R) * ROWD + lane;
  int wcur = 0;  // write cursor: rows 0 .. wcur-1 hold the particles that stay in this column

  // raw deposit sums of this lane's cell: AX[iy][ix<2], AY[iy<2][ix], AZ[iy][ix]
  double AX[3][2], AY[2][3], AZ[3][3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int b = 0; b < 3; b++) {
      AZ[a][b] = 0.0;
      if (b < 2) AX[a][b] = 0.0;
      if (a < 2) AY[a][b] = 0.0;
    }
  }

  // A mover's inbox entry is written one round late: the reservation (an atomic round trip to L2) is issued in
  // the mover's own round, its payload waits in shared memory, and the reply is only looked at when the next
  // round starts -- a whole round of arithmetic hides the latency that 73 % of the rounds (those with at least
  // one mover among 32 lanes) used to wait for.
  double *sPend = sPend_all + warp * 6 * 32 + lane;
  int pend_key = -1, pend_slot = 0;
  auto flush_pending = [&]() {
    if (pend_key >= 0) {
      const double p_x = sPend[0], p_y = sPend[32], p_px = sPend[64], p_py = sPend[96], p_pz = sPend[128], p_w = sPend[160];
      if (pend_slot < P.IC) {
        double2 *e = reinterpret_cast<double2 *>(P.ib_out + ((size_t)(pend_key >> 5) * (size_t)P.IC + pend_slot) * 8);
        e[0] = make_double2(p_x, p_y);
        e[1] = make_double2(p_px, p_py);
        e[2] = make_double2(p_pz, p_w);
        e[3] = make_double2(__longlong_as_double((long long)(pend_key & 31)), 0.0);
      } else {
        // inbox full: on through the mover buffer (k_deliver finds the column); if that is full too the particle
        // goes back into this column (rows below the read row are free) and is deposited the general way next step
        const int m = atomicAdd(P.mcount, 1);
        if (m < P.mcap) {
          P.mx[0][m] = p_x; P.mx[1][m] = p_y;
          P.mp[0][m] = p_px; P.mp[1][m] = p_py; P.mp[2][m] = p_pz;
          P.mw[m] = p_w;
          P.mflag[m] = 0;
        } else if (wcur < P.R) {
          double *row = col + (size_t)wcur * ROWD;
          row[OX] = p_x; row[OY] = p_y; row[OPX] = p_px; row[OPY] = p_py; row[OPZ] = p_pz; row[OW] = p_w;
          wcur++;
        } else {
          atomicOr(P.err, 1);
        }
      }
      pend_key = -1;
    }
  };

  // software pipeline: the next row's loads are in flight while this one computes
  // (rounds 0 .. my_cnt-1 walk the column, rounds my_cnt .. my_tot-1 take the lane's arrivals from the inbox)
  double n_x = 0, n_y = 0, n_px = 0, n_py = 0, n_pz = 0, n_w = 0;
  auto fetch = [&](int rn) {
    if (rn < my_cnt) {
      const double *row = col + (size_t)rn * ROWD;
      n_w = row[OW]; n_x = row[OX]; n_y = row[OY];
      n_px = row[OPX]; n_py = row[OPY]; n_pz = row[OPZ];
    } else if (rn < my_tot) {
      const double2 *e = reinterpret_cast<const double2 *>(ibg + (size_t)sAlist[lane * SLOT_LK + (rn - my_cnt)] * 8);
      const double2 v0 = e[0], v1 = e[1], v2 = e[2];
      n_x = v0.x; n_y = v0.y; n_px = v1.x; n_py = v1.y; n_pz = v2.x; n_w = v2.y;
    }
  };
  fetch(0);
  for (int r = 0; r < maxcnt; r++) {
    const bool active = r < my_tot;
    const double part_weight = n_w;
    const double raw_x = n_x, raw_y = n_y, raw_px = n_px, raw_py = n_py, raw_pz = n_pz;
    double px_ = n_x - P.grid_min_local[0];
    double py_ = n_y - P.grid_min_local[1];
    double part_ux = n_px * P.ipart_mc;
    double part_uy = n_py * P.ipart_mc;
    double part_uz = n_pz * P.ipart_mc;
    fetch(r + 1);
    flush_pending();
    // what happens to the particle: 0 stays in this column, 1 pushed and leaves through M (flag 0, or 1 with
    // dir >= 0), 2 unpushed through M (flag 2), 3 deleted, 4 left through its next column's inbox (pending)
    int disp = 0, dir = -1, nkx, nky;
    bool touched = false;
    double o_x = raw_x, o_y = raw_y, o_px = raw_px, o_py = raw_py, o_pz = raw_pz;
    bool extras = false;
    // only read where `extras` (resp. a mover's disp) says they were set: deliberately not initialised, zeroing
    // them every round costs ~20 instructions
    int key, dcx, dcy;
    double q_fxo, q_fxn, q_fyo, q_fyn, fjx, fjy, fjz;
    if (active) {
      double root;
      gamma_root(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0, P.dtco2, root);
      px_ = px_ + part_ux * root;
      py_ = py_ + part_uy * root;
      const double cell_x_r = px_ * P.idx[0];
      const double cell_y_r = py_ * P.idx[1];
      const int cx1 = __double2int_rd(cell_x_r + 0.5) + 1;
      const int cy1 = __double2int_rd(cell_y_r + 0.5) + 1;
      // the gather reads cell1-2..cell1+1, the deposit writes cell1-2..cell1+2
      const bool fast = (cx1 - 2 >= ox) && (cx1 + 2 <= ox + TWU - 1) && (cy1 - 2 >= oy) && (cy1 + 2 <= oy + TH - 1);
      if (!fast) {
        disp = 2;
      } else {
        double gx[3], gy[3], hx[3], hy[3];
        const double fxo = (double)(cx1 - 1) - cell_x_r, fyo = (double)(cy1 - 1) - cell_y_r;
        tri_s(fxo, gx[0], gx[1], gx[2]);
        tri_s(fyo, gy[0], gy[1], gy[2]);
        int cx2 = __double2int_rd(cell_x_r);
        tri_s((double)cx2 - cell_x_r + 0.5, hx[0], hx[1], hx[2]);
        cx2 += 1;
        int cy2 = __double2int_rd(cell_y_r);
        tri_s((double)cy2 - cell_y_r + 0.5, hy[0], hy[1], hy[2]);
        cy2 += 1;
        // shared-tile offsets of (cell-1, cell-1)
        const int o11 = (cy1 - 1 - oy) * TW + (cx1 - 1 - ox);
        const int o21 = (cy1 - 1 - oy) * TW + (cx2 - 1 - ox);
        const int o12 = (cy2 - 1 - oy) * TW + (cx1 - 1 - ox);
        const int o22 = (cy2 - 1 - oy) * TW + (cx2 - 1 - ox);
        auto gat = [&](const double *F, int o, const double *wx, const double *wy) {
          double r0 = wx[0] * F[o] + wx[1] * F[o + 1] + wx[2] * F[o + 2];
          double r1 = wx[0] * F[o + TW] + wx[1] * F[o + TW + 1] + wx[2] * F[o + TW + 2];
          double r2 = wx[0] * F[o + 2 * TW] + wx[1] * F[o + 2 * TW + 1] + wx[2] * F[o + 2 * TW + 2];
          return wy[0] * r0 + wy[1] * r1 + wy[2] * r2;
        };
        const double ex_part = gat(sEx, o21, hx, gy);
        const double ey_part = gat(sEy, o12, gx, hy);
        const double ez_part = gat(sEz, o11, gx, gy);
        const double bx_part = gat(sBx, o12, gx, hy);
        const double by_part = gat(sBy, o21, hx, gy);
        const double bz_part = gat(sBz, o22, hx, hy);
        const double cmratio = P.cmratio;
        const double uxm = part_ux + cmratio * ex_part;
        const double uym = part_uy + cmratio * ey_part;
        const double uzm = part_uz + cmratio * ez_part;
        double gm2 = uxm * uxm + uym * uym + uzm * uzm + 1.0;
        if (HC) {  // particles.F90:386-398 (-DHC_PUSH), Higuera & Cary, Phys. Plasmas 24, 052104; the result is >= 1 too
          const double beta_x = P.hc_alpha * bx_part, beta_y = P.hc_alpha * by_part, beta_z = P.hc_alpha * bz_part;
          const double beta2 = beta_x * beta_x + beta_y * beta_y + beta_z * beta_z;
          const double sigma = gm2 - beta2;
          const double beta_dot_u = beta_x * uxm + beta_y * uym + beta_z * uzm;
          gm2 = 0.5 * (sigma + sqrt(sigma * sigma + 4.0 * (beta2 + beta_dot_u * beta_dot_u)));
        }
        gamma_root(gm2, P.ccmratio, root);
        const double taux = bx_part * root, tauy = by_part * root, tauz = bz_part * root;
        const double taux2 = taux * taux, tauy2 = tauy * tauy, tauz2 = tauz * tauz;
#ifdef EPB_FAST_MATH
        const double tau = rcp_ge1(1.0 + taux2 + tauy2 + tauz2);
#else
        const double tau = 1.0 / (1.0 + taux2 + tauy2 + tauz2);
#endif
        const double uxp = ((1.0 + taux2 - tauy2 - tauz2) * uxm +
                            2.0 * ((taux * tauy + tauz) * uym + (taux * tauz - tauy) * uzm)) * tau;
        const double uyp = ((1.0 - taux2 + tauy2 - tauz2) * uym +
                            2.0 * ((tauy * tauz + taux) * uzm + (tauy * taux - tauz) * uxm)) * tau;
        const double uzp = ((1.0 - taux2 - tauy2 + tauz2) * uzm +
                            2.0 * ((tauz * taux + tauy) * uxm + (tauz * tauy - taux) * uym)) * tau;
        part_ux = uxp + cmratio * ex_part;
        part_uy = uyp + cmratio * ey_part;
        part_uz = uzp + cmratio * ez_part;
        const double part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz;
#ifdef EPB_FAST_MATH
        const double igamma = rsqrt_ge1(part_u2 + 1.0);
#else
        const double igamma = 1.0 / sqrt(part_u2 + 1.0);
#endif
        root = P.dtco2 * igamma;
        const double delta_x = part_ux * root;
        const double delta_y = part_uy * root;
        const double part_vz = part_uz * c * igamma;
        px_ = px_ + delta_x;
        py_ = py_ + delta_y;
        {  // touched: a particle boundary condition looked at this particle (it is outside the local domain)
          double pos[3] = {px_ + P.grid_min_local[0], py_ + P.grid_min_local[1], 0.0};
          double mom[3] = {P.part_mc * part_ux, P.part_mc * part_uy, P.part_mc * part_uz};
          touched = (pos[0] < P.bnd_min[0]) || (pos[0] > P.bnd_max[0]) || (pos[1] < P.bnd_min[1]) || (pos[1] > P.bnd_max[1]);
          dir = particle_bc<2>(P, pos, mom);
          o_x = pos[0]; o_y = pos[1];
          o_px = mom[0]; o_py = mom[1]; o_pz = mom[2];
        }
        // the cell the next push gathers this particle in (its half step, particles.F90:289-322)
        px_ = px_ + delta_x;
        py_ = py_ + delta_y;
        const double cxr = px_ * P.idx[0], cyr = py_ * P.idx[1];
        const int cx3 = __double2int_rd(cxr + 0.5), cy3 = __double2int_rd(cyr + 0.5);
        {
          // (a particle predicted to gather in the first ghost cell belongs to the edge column: the clamp is
          // only paid by the particles that are about to move)
          const bool stays = !touched && (cx3 + 1 == hcx) && (cy3 + 1 == hcy);
          disp = (dir == 13) ? 3 : (stays ? 0 : 1);
          if (disp == 1 && dir < 0 && !touched) {
            nkx = cx3 < 0 ? 0 : (cx3 > P.n[0] - 1 ? P.n[0] - 1 : cx3);
            nky = cy3 < 0 ? 0 : (cy3 > P.n[1] - 1 ? P.n[1] - 1 : cy3);
            if (nkx + 1 == hcx && nky + 1 == hcy) {
              disp = 0;   // edge column after all
            } else if (P.ic_out) {
              // a mover whose next column is known goes straight into that column's group inbox: one 64-byte
              // entry (two full sectors, no read-modify-write), consumed by the next push.  Reserve the entry
              // now, park the payload, write it when the next round starts (flush_pending)
              const int nkey = ((nky / T2Y) * P.tg.nt[0] + (nkx >> 4)) * (T2X * T2Y) + (nky % T2Y) * T2X + (nkx & 15);
              pend_key = nkey;
              pend_slot = atomicAdd(&P.ic_out[nkey >> 5], 1);
              sPend[0] = o_x; sPend[32] = o_y; sPend[64] = o_px; sPend[96] = o_py; sPend[128] = o_pz; sPend[160] = part_weight;
              disp = 4;
            }
          }
        }
        if (P.deposit) {
          const double fxn = (double)cx3 - cxr, fyn = (double)cy3 - cyr;
          dcx = cx3 + 1 - cx1;
          dcy = cy3 + 1 - cy1;
          const double fcx = P.kfc[0] * part_weight;
          const double fcy = P.kfc[1] * part_weight;
          const double fcz = P.kfc[2] * part_weight;
          fjx = fcx * P.part_q;
          fjy = fcy * P.part_q;
          fjz = fcz * P.part_q * part_vz;
          key = (cy1 - oy) * TW + (cx1 - ox);
          q_fxo = fxo; q_fxn = fxn; q_fyo = fyo; q_fyn = fyn;
          if (cx1 != hcx || cy1 != hcy || (dcx != 0 && dcy != 0)) {
            extras = true;  // not this lane's cell, or moved diagonally
            if ((dcx | dcy) == 0) key |= 1 << 15;  // 3x3 stencil: 21-update drain
          } else {
            // This lane's own cell, nearest cell unchanged or moved by one cell along one axis.
            // New weights on the 3x3 core (particles.F90:521-538 with the shift by dcell); the
            // part of a moved particle's stencil outside the core is queued (drain_edge).
            double wm, w0, wp;
            tri_s(fxn, wm, w0, wp);
            hx[0] = (dcx == 0 ? wm : dcx > 0 ? 0.0 : w0) - gx[0];
            hx[1] = (dcx == 0 ? w0 : dcx > 0 ? wm : wp) - gx[1];
            hx[2] = (dcx == 0 ? wp : dcx > 0 ? w0 : 0.0) - gx[2];
            const double hxa = hx[0] + (dcx < 0 ? wm : 0.0);  // running jx prefix enters the core with column -2
            tri_s(fyn, wm, w0, wp);
            hy[0] = (dcy == 0 ? wm : dcy > 0 ? 0.0 : w0) - gy[0];
            hy[1] = (dcy == 0 ? w0 : dcy > 0 ? wm : wp) - gy[1];
            hy[2] = (dcy == 0 ? wp : dcy > 0 ? w0 : 0.0) - gy[2];
            const double hya = hy[0] + (dcy < 0 ? wm : 0.0);
            double xfac1[3], yfac1[3], yfac2[3];
#pragma unroll
            for (int q = 0; q < 3; q++) {
              xfac1[q] = gx[q] + 0.5 * hx[q];
              yfac1[q] = gy[q] + 0.5 * hy[q];
              yfac2[q] = third * hy[q] + 0.5 * gy[q];
            }
            const double fhx0 = fjx * hxa, fhx1 = fjx * hx[1];
            const double fhy0 = fjy * hya, fhy1 = fjy * hy[1];
#pragma unroll
            for (int iy = 0; iy < 3; iy++) {
              AX[iy][0] += fhx0 * yfac1[iy];
              AX[iy][1] += fhx1 * yfac1[iy];
            }
#pragma unroll
            for (int ix = 0; ix < 3; ix++) {
              AY[0][ix] += fhy0 * xfac1[ix];
              AY[1][ix] += fhy1 * xfac1[ix];
            }
#pragma unroll
            for (int ix = 0; ix < 3; ix++) {
              const double zg = fjz * gx[ix], zh = fjz * hx[ix];
#pragma unroll
              for (int iy = 0; iy < 3; iy++) {
#ifdef EPB_FAST_MATH
                AZ[iy][ix] = fma(zh, yfac2[iy], fma(zg, yfac1[iy], AZ[iy][ix]));   // two FMAs instead of mul + FMA + add
#else
                AZ[iy][ix] += zg * yfac1[iy] + zh * yfac2[iy];
#endif
              }
            }
            if ((dcx | dcy) != 0) { extras = true; key |= 1 << 14; }
          }
        }
      }
    }
    // ---- particles that leave this column through the mover buffer -------------------------------------
    // (boundary-touched or leaving particles, unpushed ones, movers when the inboxes are switched off)
    bool toM = (disp == 2) || (disp == 1);
    if (active && disp == 0 && wcur >= P.R) toM = true;   // no row left in this column: on through M (flag 0)
    {
      const unsigned bal = __ballot_sync(FULL, toM);
      if (bal) {
        int base = 0;
        if (lane == __ffs(bal) - 1) base = atomicAdd(P.mcount, __popc(bal));
        base = __shfl_sync(FULL, base, __ffs(bal) - 1);
        if (toM) {
          const int m = base + __popc(bal & lt_mask);
          if (m < P.mcap) {
            P.mx[0][m] = o_x; P.mx[1][m] = o_y;
            P.mp[0][m] = o_px; P.mp[1][m] = o_py; P.mp[2][m] = o_pz;
            P.mw[m] = part_weight;
            P.mflag[m] = (disp == 2) ? 2 : (dir >= 0 ? 1 : 0);
            if (disp == 1 && dir >= 0) {
              const int slot = atomicAdd(&P.out_count[dir], 1);
              if (slot < P.out_cap) P.out_idx[(size_t)dir * P.out_cap + slot] = m;
            }
            if (disp == 0) disp = 1;   // left through M after all
          } else if (disp == 1 && dir < 0 && wcur < P.R) {
            disp = 0;  // no room in M: the particle stays in this column and is deposited the general way next step
          } else {
            atomicOr(P.err, 1);  // a leaver, an unpushed particle or a full column cannot keep it: capacity error
            disp = 3;
          }
        }
      }
    }
    if (active && disp == 0) {
      double *row = col + (size_t)wcur * ROWD;
      row[OX] = o_x;
      row[OY] = o_y;
      row[OPX] = o_px;
      row[OPY] = o_py;
      row[OPZ] = o_pz;
      if (wcur != r || r >= my_cnt) row[OW] = part_weight;
      wcur++;
    }
    if (!P.deposit) continue;
    // ---- queue the particles that are not regular for their lane ---------------------------
    const unsigned em = __ballot_sync(FULL, extras);
    if (em) {
      const int ne = __popc(em);
      if (qcount + ne > QCAP) {
        __syncwarp();
        drain_extras(P, sJ, Qd, Qk, qcount, lane, TILE_ELEMS, TW);
        __syncwarp();
        qcount = 0;
      }
      if (extras) {
        const int slot = qcount + __popc(em & lt_mask);
        Qk[slot] = key | ((dcx + 1) << 10) | ((dcy + 1) << 12);
        Qd[0 * QCAP + slot] = q_fxo; Qd[1 * QCAP + slot] = q_fxn;
        Qd[2 * QCAP + slot] = q_fyo; Qd[3 * QCAP + slot] = q_fyn;
        Qd[4 * QCAP + slot] = fjx; Qd[5 * QCAP + slot] = fjy; Qd[6 * QCAP + slot] = fjz;
      }
      qcount += ne;
    }
  }
  if (qcount) {
    __syncwarp();
    drain_extras(P, sJ, Qd, Qk, qcount, lane, TILE_ELEMS, TW);
  }
  flush_pending();
  if (my_tot > 0) P.cnt[my_key] = wcur;
  // ---- flush this lane's cell sums: prefixes of particles.F90:563-571, one update per point ----
  if (P.deposit && my_tot > 0) {
    const int hb = (hcy - oy) * TW + (hcx - ox);
#pragma unroll
    for (int iy = 0; iy < 3; iy++) {
      const double v0 = -AX[iy][0];
      const double v1 = v0 - AX[iy][1];
      smem_add(&sJ[hb + (iy - 1) * TW - 1], v0);
      smem_add(&sJ[hb + (iy - 1) * TW], v1);
    }
#pragma unroll
    for (int ix = 0; ix < 3; ix++) {
      const double v0 = -AY[0][ix];
      const double v1 = v0 - AY[1][ix];
      smem_add(&sJ[TILE_ELEMS + hb - TW + (ix - 1)], v0);
      smem_add(&sJ[TILE_ELEMS + hb + (ix - 1)], v1);
    }
#pragma unroll
    for (int iy = 0; iy < 3; iy++)
#pragma unroll
      for (int ix = 0; ix < 3; ix++) smem_add(&sJ[2 * TILE_ELEMS + hb + (iy - 1) * TW + (ix - 1)], AZ[iy][ix]);
  }
  __syncthreads();
  if (!P.deposit) return;
  for (int q = tid; q < TILE_ELEMS; q += PUSH2D_THREADS) {
    const int lx = q % TW, ly = q / TW;
    const int cx = ox + lx, cy = oy + ly;
    const bool ok = (lx < TWU) && (cx >= 1 - NG) && (cx <= P.n[0] + NG) && (cy >= 1 - NG) && (cy <= P.n[1] + NG);
    if (!ok) continue;
    const size_t o = gofs<2>(P, cx, cy, 1);
#pragma unroll
    for (int f = 0; f < 3; f++) {
      const double val = sJ[f * TILE_ELEMS + q];
      if (val != 0.0) atomicAdd(P.j[f] + o, val);
    }
  }
}

// The mover buffer's unpushed entries (flag 2: no room in their column, or stencil outside their tile):
// push_one on the buffer itself.  A particle that leaves the rank is flagged 1 by outbox_put (P.gone is the
// flag array here), everything else becomes an ordinary flag-0 entry for k_deliver.
template <int ND, bool HC = false>
__global__ void __launch_bounds__(256) push_generic_m(const __grid_constant__ PushParams P) {
  int n = *P.mcount;
  if (n > P.mcap) n = P.mcap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (P.gone[i] != 2) continue;
    P.gone[i] = 0;
    push_one<ND, HC>(P, i);
  }
}


// ---------------------------------------------------------------------------
// Tiled 3D kernel
// ---------------------------------------------------------------------------
// One CTA (8 warps) per 8x8x4-cell tile of the cell-sorted layout, two CTAs per SM.  J is
// accumulated in a shared tile (+3 halo cells, 47 KB) and flushed once with global reductions;
// E/B are gathered through the read-only path (a 14^3 x 6 tile does not fit beside the J tile
// and the reduction scratch; the tile's fields stay L1/L2 resident).  The deposit uses the
// transposition of push_tiled_2d, generalised to any number of cells per warp (8 particles per
// cell in the C4 workload means ~4 cells per batch): the lanes of a batch are first grouped by
// cell key so that lanes of one cell own adjacent scratch columns; then, plane by plane of the
// 3x3x3 stencil (21 + 21 + 12 = 54 values that do not cancel structurally), every lane stores
// its values in its column and lane q sums row q run by run, one shared update per (cell, value).
// Particles whose nearest cell changed (wider stencil) are queued per warp and deposited with
// the reference's general loop (epoch3d particles.F90:603-648) on the shared tile.
// tile = 8 x 8 x 4 cells (T3 x T3 x T3Z), 8 warps, two CTAs per SM so that one CTA's J-tile prologue /
// epilogue overlaps the other's particle loop
constexpr int T3 = 8, T3Z = 4, HALO3 = 3, JW3 = T3 + 2 * HALO3, JD3 = T3Z + 2 * HALO3, JT3 = JW3 * JW3 * JD3;
constexpr int P3_THREADS = 256, P3_WARPS = P3_THREADS / 32;
constexpr int Q3CAP = 16, Q3DBL = 9, SLOW3CAP = 254, S3ROWS = 21;
constexpr size_t PUSH3D_SMEM =
    sizeof(double) * ((size_t)3 * JT3 + (size_t)P3_WARPS * S3ROWS * SPITCH + (size_t)P3_WARPS * Q3DBL * Q3CAP) +
    sizeof(int) * ((size_t)P3_WARPS * Q3CAP + (size_t)P3_WARPS * 32 + SLOW3CAP + 2);

// Edge deposit of a queued particle whose nearest cell moved by one cell along exactly one axis a
// (dc = +-1): its 3x3x3 core joined the transposed reduction with shifted new weights; what is left is
// J_a on the plane the core has no row for (a-index -2 when dc < 0, else +1: 9 values) and J_b, J_c on
// the outer plane at a-index +-2 (6 + 6 values).  Every weight of the reference's loop (epoch3d
// particles.F90:603-648) is h_a * S(g_b, h_b, g_c, h_c) with the bilinear form
// S = g_b g_c + (g_b h_c + h_b g_c) / 2 + h_b h_c / 3, cyclically in (a, b, c): 21 updates, not ~108.
__device__ __noinline__ void drain_edge_3d(const PushParams &P, double *sJ, const double *Qd, int slot, int key, int a,
                                           int dc, double fja, double fjb, double fjc) {
  const int b = (a + 1) % 3, c = (a + 2) % 3;
  const double third = P.third;
  double g[3][3], n[3][3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    tri(Qd[(2 * d) * Q3CAP + slot], g[d][0], g[d][1], g[d][2]);
    tri(Qd[(2 * d + 1) * Q3CAP + slot], n[d][0], n[d][1], n[d][2]);
  }
  // axis a: new weights shifted by dc onto the core; the outer cell only carries a new weight
  double hCs, hE;
  if (dc > 0) { hCs = (0.0 - g[a][0]) + (n[a][0] - g[a][1]) + (n[a][1] - g[a][2]); hE = n[a][2]; }
  else { hCs = (n[a][1] - g[a][0]) + (n[a][2] - g[a][1]) + (0.0 - g[a][2]); hE = n[a][0]; }
  double gb[3], hb[3], gc[3], hc[3];
#pragma unroll
  for (int q = 0; q < 3; q++) { gb[q] = g[b][q]; hb[q] = n[b][q] - g[b][q]; gc[q] = g[c][q]; hc[q] = n[c][q] - g[c][q]; }
  const int str[3] = {1, JW3, JW3 * JW3};
  const int sa = str[a], sb = str[b], sc = str[c];
  const int oe = key + 2 * dc * sa;                    // outer plane
  const int oa = dc < 0 ? oe : key + sa;               // plane of the J_a values the core has no row for
  const double ha = dc < 0 ? hE : hCs;
  // J_a: - fj_a * ha * S(b, c)
#pragma unroll
  for (int ic = 0; ic < 3; ic++)
#pragma unroll
    for (int ib = 0; ib < 3; ib++) {
      const double S = gb[ib] * gc[ic] + 0.5 * (gb[ib] * hc[ic] + hb[ib] * gc[ic]) + third * (hb[ib] * hc[ic]);
      smem_add(&sJ[a * JT3 + oa + (ib - 1) * sb + (ic - 1) * sc], -(fja * (ha * S)));
    }
  // J_b on the outer plane: running sum along b of - fj_b * h_b * S(a = outer cell, c); last row cancels
#pragma unroll
  for (int ic = 0; ic < 3; ic++) {
    const double Sac = hE * (0.5 * gc[ic] + third * hc[ic]);
    double run = 0.0;
#pragma unroll
    for (int ib = 0; ib < 2; ib++) {
      run = run - fjb * (hb[ib] * Sac);
      smem_add(&sJ[b * JT3 + oe + (ib - 1) * sb + (ic - 1) * sc], run);
    }
  }
  // J_c on the outer plane: running sum along c
#pragma unroll
  for (int ib = 0; ib < 3; ib++) {
    const double Sab = hE * (0.5 * gb[ib] + third * hb[ib]);
    double run = 0.0;
#pragma unroll
    for (int ic = 0; ic < 2; ic++) {
      run = run - fjc * (hc[ic] * Sab);
      smem_add(&sJ[c * JT3 + oe + (ib - 1) * sb + (ic - 1) * sc], run);
    }
  }
}

// General deposit of queued particles on the shared tile: epoch3d particles.F90:603-648.
__device__ __noinline__ void drain_extras_3d(const PushParams &P, double *sJ, const double *Qd, const int *Qk, int n,
                                             int lane) {
  if (lane >= n) return;
  const int pk = Qk[lane];
  const int key = pk & 4095;
  int dcell[3] = {((pk >> 12) & 3) - 1, ((pk >> 14) & 3) - 1, ((pk >> 16) & 3) - 1};
  const double fjx = Qd[6 * Q3CAP + lane], fjy = Qd[7 * Q3CAP + lane], fjz = Qd[8 * Q3CAP + lane];
  if (pk & (1 << 18)) {  // core already reduced with the batch: only the outer values are left
    const int a = dcell[0] != 0 ? 0 : dcell[1] != 0 ? 1 : 2;
    const double fj[3] = {fjx, fjy, fjz};
    drain_edge_3d(P, sJ, Qd, lane, key, a, dcell[a], fj[a], fj[(a + 1) % 3], fj[(a + 2) % 3]);
    return;
  }
  double G[3][5], H[3][5];
  int mn[3], mx[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    G[d][0] = G[d][4] = 0.0;
    tri(Qd[(2 * d) * Q3CAP + lane], G[d][1], G[d][2], G[d][3]);
    double wm, w0, wp;
    tri(Qd[(2 * d + 1) * Q3CAP + lane], wm, w0, wp);
#pragma unroll
    for (int q = 0; q < 5; q++) {
      const int r = q - 2 - dcell[d];
      H[d][q] = ((r == -1) ? wm : (r == 0) ? w0 : (r == 1) ? wp : 0.0) - G[d][q];
    }
    mn[d] = -1 + (dcell[d] - 1) / 2;
    mx[d] = 1 + (dcell[d] + 1) / 2;
  }
  const double *gx = &G[0][2], *gy = &G[1][2], *gz = &G[2][2];
  const double *hx = &H[0][2], *hy = &H[1][2], *hz = &H[2][2];
  const double third = P.third;
  double jzh[5][5];
  for (int a = 0; a < 5; a++)
    for (int b = 0; b < 5; b++) jzh[a][b] = 0.0;
  for (int iz = mn[2]; iz <= mx[2]; iz++) {
    const double zfac1 = gz[iz] + 0.5 * hz[iz];
    const double zfac2 = third * hz[iz] + 0.5 * gz[iz];
    const double gz_iz = gz[iz], hz_iz = hz[iz];
    double jyh[5] = {0, 0, 0, 0, 0};
    for (int iy = mn[1]; iy <= mx[1]; iy++) {
      const double yfac1 = gy[iy] + 0.5 * hy[iy];
      const double yfac2 = third * hy[iy] + 0.5 * gy[iy];
      const double hygz = hy[iy] * gz_iz;
      const double hyhz = hy[iy] * hz_iz;
      const double yzfac = gy[iy] * zfac1 + hy[iy] * zfac2;
      const double hzyfac1 = hz_iz * yfac1;
      const double hzyfac2 = hz_iz * yfac2;
      double jxh = 0.0;
      for (int ix = mn[0]; ix <= mx[0]; ix++) {
        const double xfac1 = gx[ix] + 0.5 * hx[ix];
        const double xfac2 = third * hx[ix] + 0.5 * gx[ix];
        const double wx = hx[ix] * yzfac;
        const double wy = xfac1 * hygz + xfac2 * hyhz;
        const double wz = gx[ix] * hzyfac1 + hx[ix] * hzyfac2;
        jxh = jxh - fjx * wx;
        jyh[ix + 2] = jyh[ix + 2] - fjy * wy;
        jzh[iy + 2][ix + 2] = jzh[iy + 2][ix + 2] - fjz * wz;
        const int o = key + (iz * JW3 + iy) * JW3 + ix;
        // the last column / row / plane of each running sum cancels structurally (sum of hx = 0)
        if (ix < mx[0]) smem_add(&sJ[o], jxh);
        if (iy < mx[1]) smem_add(&sJ[JT3 + o], jyh[ix + 2]);
        if (iz < mx[2]) smem_add(&sJ[2 * JT3 + o], jzh[iy + 2][ix + 2]);
      }
    }
  }
}

__global__ void __launch_bounds__(P3_THREADS, 2) push_tiled_3d(const __grid_constant__ PushParams P) {
  extern __shared__ double sm[];
  double *sJ = sm;                                       // [3][JW3][JW3][JW3]
  double *sS_all = sJ + 3 * JT3;                         // [warps][27][33]
  double *sQd_all = sS_all + P3_WARPS * S3ROWS * SPITCH;  // [warps][9][Q3CAP]
  int *sQk_all = reinterpret_cast<int *>(sQd_all + P3_WARPS * Q3DBL * Q3CAP);
  int *sRK_all = sQk_all + P3_WARPS * Q3CAP;
  int *sSlow = sRK_all + P3_WARPS * 32;
  int *sSlowCount = sSlow + SLOW3CAP;
  const int tile = blockIdx.x;
  const int ttx = tile % P.tg.nt[0], tty = (tile / P.tg.nt[0]) % P.tg.nt[1], ttz = tile / (P.tg.nt[0] * P.tg.nt[1]);
  const int ox = ttx * T3 + 1 - HALO3;  // cell index of shared column 0
  const int oy = tty * T3 + 1 - HALO3;
  const int oz = ttz * T3Z + 1 - HALO3;
  const long long start = P.tile_start[tile];
  long long end = P.tile_start[tile + 1];
  if (end > P.n_sorted_clip) end = P.n_sorted_clip;
  if (start >= end) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { *sSlowCount = 0; sSlowCount[1] = 0; }  // [1]: next batch of the tile (dynamic distribution)
  for (int q = tid; q < 3 * JT3; q += P3_THREADS) sJ[q] = 0.0;
  __syncthreads();

  const double third = P.third;
  double *S = sS_all + warp * S3ROWS * SPITCH;
  double *Qd = sQd_all + warp * Q3DBL * Q3CAP;
  int *Qk = sQk_all + warp * Q3CAP;
  int *RK = sRK_all + warp * 32;
  int qcount = 0;  // warp-uniform
  const unsigned lt_mask = (1u << lane) - 1u;
  // lane q owns row q of every plane pass: rows 0..5 = jx(iy, ix<2), 6..11 = jy(iy<2, ix), 12..20 = jz(iy, ix)
  int off0;
  {
    int comp, diy, dix;
    if (lane < 6) { comp = 0; diy = lane / 2; dix = lane % 2; }
    else if (lane < 12) { comp = 1; diy = (lane - 6) / 3; dix = (lane - 6) % 3; }
    else { comp = 2; diy = ((lane - 12) / 3) % 3; dix = (lane - 12) % 3; }
    off0 = comp * JT3 + ((0 - 1) * JW3 + (diy - 1)) * JW3 + (dix - 1);
  }

  // no software prefetch here: 16 warps per SM hide the load latency and the kernel is short of registers
  // batches are handed out dynamically: a warp that ran the wide-stencil drain takes fewer of them, so the
  // warps of a CTA reach the final barrier together (the static split left 19 % of the samples there)
  for (;;) {
    int bidx = 0;
    if (lane == 0) bidx = atomicAdd(&sSlowCount[1], 1);
    bidx = __shfl_sync(FULL, bidx, 0);
    const long long i = start + (long long)bidx * 32 + lane;
    if (i - lane >= end) break;
    const bool active = i < end;
    double part_weight = 0.0, pp[3] = {0.0, 0.0, 0.0}, part_ux = 0.0, part_uy = 0.0, part_uz = 0.0;
    if (active) {
      part_weight = P.w[i];
      pp[0] = P.x[0][i] - P.grid_min_local[0];
      pp[1] = P.x[1][i] - P.grid_min_local[1];
      pp[2] = P.x[2][i] - P.grid_min_local[2];
      part_ux = P.p[0][i] * P.ipart_mc;
      part_uy = P.p[1][i] * P.ipart_mc;
      part_uz = P.p[2][i] * P.ipart_mc;
    }
    // ---- phase A: half-step move and nearest cell (epoch3d particles.F90:320-370) ---------
    int key = -1;
    int cell1[3] = {0, 0, 0};
    double cell_r[3] = {0, 0, 0};
    if (active) {
      double root;
      gamma_root(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0, P.dtco2, root);
      pp[0] = pp[0] + part_ux * root;
      pp[1] = pp[1] + part_uy * root;
      pp[2] = pp[2] + part_uz * root;
      bool fast = true;
      const int org[3] = {ox, oy, oz};
#pragma unroll
      for (int d = 0; d < 3; d++) {
        cell_r[d] = pp[d] * P.idx[d];
        cell1[d] = __double2int_rd(cell_r[d] + 0.5) + 1;
        fast = fast && (cell1[d] - 2 >= org[d]) && (cell1[d] + 2 <= org[d] + (d == 2 ? JD3 : JW3) - 1);
      }
      if (!fast) {
        const int slot = atomicAdd(sSlowCount, 1);
        if (slot < SLOW3CAP) sSlow[slot] = (int)i;
        else push_one<3>(P, i);
      } else {
        key = ((cell1[2] - oz) * JW3 + (cell1[1] - oy)) * JW3 + (cell1[0] - ox);
      }
    }
    // ---- group the lanes by cell key: equal keys get adjacent scratch columns ----------------
    int pos = 0;
    unsigned bnd = 0;
    if (P.deposit) {
      unsigned rest = __ballot_sync(FULL, key >= 0);
      int base = 0;
      while (rest) {
        const int l = __ffs(rest) - 1;
        const int kl = __shfl_sync(FULL, key, l);
        const unsigned m = __ballot_sync(FULL, key == kl);
        if (key == kl) pos = base + __popc(m & lt_mask);
        base += __popc(m);
        bnd |= 1u << (base - 1);
        if (lane == l) RK[base - 1] = kl;
        rest &= ~m;
      }
    }
    // ---- phase B: gather, Boris rotation, move, store ------------------------------------
    bool regular = false, extras = false, edge = false;
    int dc[3] = {0, 0, 0};
    double hfold[3] = {0.0, 0.0, 0.0};
    double G[3][3], H[3][3];  // g weights at t+dt/2; H: staggered weights, later (new weights - g)
    double fo[3] = {0, 0, 0}, fn[3] = {0, 0, 0};
    double fjx = 0, fjy = 0, fjz = 0;
    if (key >= 0) {
      int cell2[3];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        fo[d] = (double)(cell1[d] - 1) - cell_r[d];
        tri(fo[d], G[d][0], G[d][1], G[d][2]);
        int c2 = __double2int_rd(cell_r[d]);
        tri((double)c2 - cell_r[d] + 0.5, H[d][0], H[d][1], H[d][2]);
        cell2[d] = c2 + 1;
      }
      const double ex_part = gather_g<3>(P, P.e[0], H[0], cell2[0], G[1], cell1[1], G[2], cell1[2]);
      const double ey_part = gather_g<3>(P, P.e[1], G[0], cell1[0], H[1], cell2[1], G[2], cell1[2]);
      const double ez_part = gather_g<3>(P, P.e[2], G[0], cell1[0], G[1], cell1[1], H[2], cell2[2]);
      const double bx_part = gather_g<3>(P, P.b[0], G[0], cell1[0], H[1], cell2[1], H[2], cell2[2]);
      const double by_part = gather_g<3>(P, P.b[1], H[0], cell2[0], G[1], cell1[1], H[2], cell2[2]);
      const double bz_part = gather_g<3>(P, P.b[2], H[0], cell2[0], H[1], cell2[1], G[2], cell1[2]);
      const double cmratio = P.cmratio;
      const double uxm = part_ux + cmratio * ex_part;
      const double uym = part_uy + cmratio * ey_part;
      const double uzm = part_uz + cmratio * ez_part;
      double root;
      gamma_root(uxm * uxm + uym * uym + uzm * uzm + 1.0, P.ccmratio, root);
      const double taux = bx_part * root, tauy = by_part * root, tauz = bz_part * root;
      const double taux2 = taux * taux, tauy2 = tauy * tauy, tauz2 = tauz * tauz;
#ifdef EPB_FAST_MATH
      const double tau = rcp_ge1(1.0 + taux2 + tauy2 + tauz2);
#else
      const double tau = 1.0 / (1.0 + taux2 + tauy2 + tauz2);
#endif
      const double uxp = ((1.0 + taux2 - tauy2 - tauz2) * uxm +
                          2.0 * ((taux * tauy + tauz) * uym + (taux * tauz - tauy) * uzm)) * tau;
      const double uyp = ((1.0 - taux2 + tauy2 - tauz2) * uym +
                          2.0 * ((tauy * tauz + taux) * uzm + (tauy * taux - tauz) * uxm)) * tau;
      const double uzp = ((1.0 - taux2 - tauy2 + tauz2) * uzm +
                          2.0 * ((tauz * taux + tauy) * uxm + (tauz * tauy - taux) * uym)) * tau;
      part_ux = uxp + cmratio * ex_part;
      part_uy = uyp + cmratio * ey_part;
      part_uz = uzp + cmratio * ez_part;
      // epoch3d particles.F90:470-474
      gamma_root(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0, P.dtco2, root);
      const double delta[3] = {part_ux * root, part_uy * root, part_uz * root};
#pragma unroll
      for (int d = 0; d < 3; d++) pp[d] = pp[d] + delta[d];
      {
        double pos3[3] = {pp[0] + P.grid_min_local[0], pp[1] + P.grid_min_local[1], pp[2] + P.grid_min_local[2]};
        double mom[3] = {P.part_mc * part_ux, P.part_mc * part_uy, P.part_mc * part_uz};
        const int dir = particle_bc<3>(P, pos3, mom);
#pragma unroll
        for (int d = 0; d < 3; d++) { P.x[d][i] = pos3[d]; P.p[d][i] = mom[d]; }
        if (dir >= 0) outbox_put(P, i, dir);
      }
      if (P.deposit) {
        double W[3][3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          pp[d] = pp[d] + delta[d];
          const double cr = pp[d] * P.idx[d];
          const int c3 = __double2int_rd(cr + 0.5);
          fn[d] = (double)c3 - cr;
          dc[d] = c3 + 1 - cell1[d];
          tri(fn[d], W[d][0], W[d][1], W[d][2]);
        }
        const double fcx = P.kfc[0] * part_weight;
        const double fcy = P.kfc[1] * part_weight;
        const double fcz = P.kfc[2] * part_weight;
        fjx = fcx * P.part_q;
        fjy = fcy * P.part_q;
        fjz = fcz * P.part_q;
        const int nmove = (dc[0] != 0) + (dc[1] != 0) + (dc[2] != 0);
        if (nmove >= 2) {
          extras = true;  // moved along two or three axes: general loop
        } else {
          // nearest cell unchanged, or moved by one cell along one axis: new weights shifted onto the
          // 3x3x3 core; the running prefix of that axis' own component enters the core with the value
          // of the outer cell at -2 (hfold); the outer values are queued (drain_edge_3d)
          regular = true;
#pragma unroll
          for (int d = 0; d < 3; d++) {
            const double n0 = dc[d] == 0 ? W[d][0] : dc[d] > 0 ? 0.0 : W[d][1];
            const double n1 = dc[d] == 0 ? W[d][1] : dc[d] > 0 ? W[d][0] : W[d][2];
            const double n2 = dc[d] == 0 ? W[d][2] : dc[d] > 0 ? W[d][1] : 0.0;
            H[d][0] = n0 - G[d][0];
            H[d][1] = n1 - G[d][1];
            H[d][2] = n2 - G[d][2];
            hfold[d] = H[d][0] + (dc[d] < 0 ? W[d][0] : 0.0);
          }
          if (nmove == 1) { extras = true; edge = true; }
        }
      }
    }
    if (!P.deposit) continue;
    // ---- queue the particles with a wider stencil ---------------------------------------
    const unsigned em = __ballot_sync(FULL, extras);
    if (em) {
      const int ne = __popc(em);
      if (qcount + ne > Q3CAP) {
        __syncwarp();
        drain_extras_3d(P, sJ, Qd, Qk, qcount, lane);
        __syncwarp();
        qcount = 0;
      }
      // a batch may hold more wide particles than the queue: the surplus lanes deposit directly
      const int myslot = qcount + __popc(em & lt_mask);
      if (extras) {
        if (myslot < Q3CAP) {
          Qk[myslot] = key | ((dc[0] + 1) << 12) | ((dc[1] + 1) << 14) | ((dc[2] + 1) << 16) | (edge ? 1 << 18 : 0);
#pragma unroll
          for (int d = 0; d < 3; d++) { Qd[(2 * d) * Q3CAP + myslot] = fo[d]; Qd[(2 * d + 1) * Q3CAP + myslot] = fn[d]; }
          Qd[6 * Q3CAP + myslot] = fjx; Qd[7 * Q3CAP + myslot] = fjy; Qd[8 * Q3CAP + myslot] = fjz;
        }
      }
      const int over = qcount + ne - Q3CAP;
      qcount = over > 0 ? Q3CAP : qcount + ne;
      if (over > 0) {  // warp-uniform: drain, then let the surplus lanes use the emptied queue
        __syncwarp();
        drain_extras_3d(P, sJ, Qd, Qk, qcount, lane);
        __syncwarp();
        qcount = 0;
        if (extras && myslot >= Q3CAP) {
          const int s2 = myslot - Q3CAP;
          Qk[s2] = key | ((dc[0] + 1) << 12) | ((dc[1] + 1) << 14) | ((dc[2] + 1) << 16) | (edge ? 1 << 18 : 0);
#pragma unroll
          for (int d = 0; d < 3; d++) { Qd[(2 * d) * Q3CAP + s2] = fo[d]; Qd[(2 * d + 1) * Q3CAP + s2] = fn[d]; }
          Qd[6 * Q3CAP + s2] = fjx; Qd[7 * Q3CAP + s2] = fjy; Qd[8 * Q3CAP + s2] = fjz;
        }
        qcount = over;
      }
    }
    // ---- transposed reduction, one z plane of the stencil per pass -------------------------
    {
      const double *gx = G[0], *gy = G[1], *gz = G[2], *hx = H[0], *hy = H[1], *hz = H[2];
      double jzh[3][3];
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) jzh[a][b] = 0.0;
      double *col = S + pos;
#pragma unroll
      for (int iz = 0; iz < 3; iz++) {
        const int nrows = iz < 2 ? 21 : 12;
        if (regular) {
          const double zfac1 = gz[iz] + 0.5 * hz[iz];
          const double zfac2 = third * hz[iz] + 0.5 * gz[iz];
          const double gz_iz = gz[iz], hz_iz = hz[iz];
          const double hzw = iz == 0 ? hfold[2] : hz_iz;   // jz prefix along z
          double jyh[3] = {0.0, 0.0, 0.0};
#pragma unroll
          for (int iy = 0; iy < 3; iy++) {
            const double yfac1 = gy[iy] + 0.5 * hy[iy];
            const double yfac2 = third * hy[iy] + 0.5 * gy[iy];
            const double hyw = iy == 0 ? hfold[1] : hy[iy];  // jy prefix along y
            const double hygz = hyw * gz_iz;
            const double hyhz = hyw * hz_iz;
            const double yzfac = gy[iy] * zfac1 + hy[iy] * zfac2;
            const double hzyfac1 = hzw * yfac1;
            const double hzyfac2 = hzw * yfac2;
            double jxh = 0.0;
#pragma unroll
            for (int ix = 0; ix < 3; ix++) {
              const double xfac1 = gx[ix] + 0.5 * hx[ix];
              const double xfac2 = third * hx[ix] + 0.5 * gx[ix];
              const double wx = (ix == 0 ? hfold[0] : hx[ix]) * yzfac;
              const double wy = xfac1 * hygz + xfac2 * hyhz;
              const double wz = gx[ix] * hzyfac1 + hx[ix] * hzyfac2;
              jxh = jxh - fjx * wx;
              jyh[ix] = jyh[ix] - fjy * wy;
              jzh[iy][ix] = jzh[iy][ix] - fjz * wz;
              if (ix < 2) col[(iy * 2 + ix) * SPITCH] = jxh;
              if (iy < 2) col[(6 + iy * 3 + ix) * SPITCH] = jyh[ix];
              if (iz < 2) col[(12 + iy * 3 + ix) * SPITCH] = jzh[iy][ix];
            }
          }
        } else if (key >= 0) {
          for (int q = 0; q < nrows; q++) col[q * SPITCH] = 0.0;
        }
        __syncwarp();
        if (lane < nrows) {
          const double *row = S + lane * SPITCH;
          const int off = off0 + iz * JW3 * JW3;
          unsigned b = bnd;
          int lo = 0;
          while (b) {  // one running sum per cell key (warp-uniform bounds)
            const int hi = __ffs(b);
            b &= b - 1;
            double a0 = 0.0, a1 = 0.0;
            int j = lo;
            for (; j + 2 <= hi; j += 2) { a0 += row[j]; a1 += row[j + 1]; }
            if (j < hi) a0 += row[j];
            smem_add(&sJ[off + RK[hi - 1]], a0 + a1);
            lo = hi;
          }
        }
        __syncwarp();
      }
    }
  }
  if (qcount) {
    __syncwarp();
    drain_extras_3d(P, sJ, Qd, Qk, qcount, lane);
  }
  __syncthreads();
  {
    int ns = *sSlowCount;
    if (ns > SLOW3CAP) ns = SLOW3CAP;
    for (int q = tid; q < ns; q += P3_THREADS) push_one<3>(P, sSlow[q]);
  }
  for (int q = tid; q < JT3; q += P3_THREADS) {
    const int lx = q % JW3, ly = (q / JW3) % JW3, lz = q / (JW3 * JW3);
    const int cx = ox + lx, cy = oy + ly, cz = oz + lz;
    const bool ok = (cx >= 1 - NG) && (cx <= P.n[0] + NG) && (cy >= 1 - NG) && (cy <= P.n[1] + NG) &&
                    (cz >= 1 - NG) && (cz <= P.n[2] + NG);
    if (!ok) continue;
    const size_t o = gofs<3>(P, cx, cy, cz);
#pragma unroll
    for (int f = 0; f < 3; f++) {
      const double val = sJ[f * JT3 + q];
      if (val != 0.0) atomicAdd(P.j[f] + o, val);
    }
  }
}


// ---------------------------------------------------------------------------
// Slot-column 3D kernel (layout 3, EPB_PUSH3D_VARIANT=1)
// ---------------------------------------------------------------------------
// The slot-column store of push_slots_2d in 3D: a column per cell, rows of 7 x 32 doubles (x y z px py pz w),
// in-place compaction by the owning lane, movers through the group inboxes, boundary-touched / leaving /
// out-of-tile particles through the mover buffer.  One CTA (6 warps) per 16 x 4 x 3-cell tile; a warp's 32 columns
// are 16 x 2 cells of one z plane, so a half-warp always works on 16 CONSECUTIVE cells of one x row: every shared-
// memory access of the gather and of the deposit is then conflict-free whatever the row pitch (the first version,
// 8 x 8 x 4 tiles with 8 x 4-cell groups, lost half of its shared-memory bandwidth to bank conflicts between the
// rows of a half-warp: profiles/r02_ncu_push_bag_3d_v1_cells.txt).  Because a particle is always gathered inside
// its tile (the prediction that places it is the push's own half step), the tile's E/B live in shared memory with
// a halo of only -2 / +1 cells and the gather never touches global memory -- the step push_tiled_3d could not
// take, since its sorted order goes stale between sorts.
// Deposit.  Shared memory has no FP64 add: atomicAdd on it is an ATOMS.CAST.SPIN loop, and 54 of those per
// particle were 36 % of the stall samples and a quarter of the instructions of the second version
// (profiles/r02_ncu_push_bag_3d_v2_16x4x2.txt).  So every warp owns a PRIVATE 18 x 4 x 3-point J tile (its 16 x 2
// cells and the 3x3x3 stencil's halo of one).  The lanes of a warp sit in 32 different cells, so for a particle
// whose nearest cell is unchanged (88 % of a thermal plasma's) each of the 54 non-cancelling values of the
// reference's loop lands on a different address in every lane: plain load / add / store, no atomics, no loops, and
// the 54 updates are independent so their latencies overlap.  A particle whose nearest cell moved (wider stencil)
// is queued per warp and drained densely with the reference's general loop into the CTA's J tile (halo 2, shared
// atomics); the private tiles are added to it once at the end, and the CTA tile goes to global memory as before.
constexpr int B3X = 16, B3Y = 4, B3Z = 3, B3N = B3X * B3Y * B3Z;          // tile = 192 cells = 6 warps x 32 columns
constexpr int B3_THREADS = 192, B3_WARPS = 6;
constexpr int EB3X = B3X + 3, EB3Y = B3Y + 3, EB3Z = B3Z + 3, EB3N = EB3X * EB3Y * EB3Z;   // E/B: cells origin-2 .. origin+T
constexpr int EB3P = (EB3N + 1) & ~1;                                      // padded component stride
constexpr int JB3H = 2, JB3X = B3X + 2 * JB3H, JB3Y = B3Y + 2 * JB3H, JB3Z = B3Z + 2 * JB3H, JB3N = JB3X * JB3Y * JB3Z;
constexpr int PV3X = B3X + 2, PV3Y = 4, PV3Z = 3, PV3N = PV3X * PV3Y * PV3Z;   // a warp's private J tile (one component)
constexpr size_t PUSHBAG3D_SMEM =
    sizeof(double) * ((size_t)6 * EB3P + (size_t)3 * JB3N + (size_t)B3_WARPS * 3 * PV3N + (size_t)B3_WARPS * Q3DBL * Q3CAP) +
    sizeof(int) * ((size_t)B3_WARPS * Q3CAP + (size_t)B3_WARPS * 32 + 2) + (size_t)B3_WARPS * 32 * SLOT_LK;

__device__ __forceinline__ void drain_edge_3d_slots(double *sJ, const double *Qd, int pk, int lane, double third);
// drain_extras_3d's general loop (epoch3d particles.F90:603-648) on a J tile of any geometry
__device__ __noinline__ void drain_general_3d(const PushParams &P, double *sJ, const double *Qd, const int *Qk, int n,
                                              int lane, int sy, int sz, int jt) {
  if (lane >= n) return;
  const int pk = Qk[lane];
  if (pk & (1 << 20)) {   // a one-axis mover whose core went to the private tile: only the 21 values beyond it
    drain_edge_3d_slots(sJ, Qd, pk, lane, P.third);
    return;
  }
  const int key = pk & 4095;
  const int dcell[3] = {((pk >> 12) & 3) - 1, ((pk >> 14) & 3) - 1, ((pk >> 16) & 3) - 1};
  const double fjx = Qd[6 * Q3CAP + lane], fjy = Qd[7 * Q3CAP + lane], fjz = Qd[8 * Q3CAP + lane];
  double G[3][5], H[3][5];
  int mn[3], mx[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    G[d][0] = G[d][4] = 0.0;
    tri(Qd[(2 * d) * Q3CAP + lane], G[d][1], G[d][2], G[d][3]);
    double wm, w0, wp;
    tri(Qd[(2 * d + 1) * Q3CAP + lane], wm, w0, wp);
#pragma unroll
    for (int q = 0; q < 5; q++) {
      const int r = q - 2 - dcell[d];
      H[d][q] = ((r == -1) ? wm : (r == 0) ? w0 : (r == 1) ? wp : 0.0) - G[d][q];
    }
    mn[d] = -1 + (dcell[d] - 1) / 2;
    mx[d] = 1 + (dcell[d] + 1) / 2;
  }
  const double *gx = &G[0][2], *gy = &G[1][2], *gz = &G[2][2];
  const double *hx = &H[0][2], *hy = &H[1][2], *hz = &H[2][2];
  const double third = P.third;
  double jzh[5][5];
  for (int a = 0; a < 5; a++)
    for (int b = 0; b < 5; b++) jzh[a][b] = 0.0;
  for (int iz = mn[2]; iz <= mx[2]; iz++) {
    const double zfac1 = gz[iz] + 0.5 * hz[iz];
    const double zfac2 = third * hz[iz] + 0.5 * gz[iz];
    const double gz_iz = gz[iz], hz_iz = hz[iz];
    double jyh[5] = {0, 0, 0, 0, 0};
    for (int iy = mn[1]; iy <= mx[1]; iy++) {
      const double yfac1 = gy[iy] + 0.5 * hy[iy];
      const double yfac2 = third * hy[iy] + 0.5 * gy[iy];
      const double hygz = hy[iy] * gz_iz;
      const double hyhz = hy[iy] * hz_iz;
      const double yzfac = gy[iy] * zfac1 + hy[iy] * zfac2;
      const double hzyfac1 = hz_iz * yfac1;
      const double hzyfac2 = hz_iz * yfac2;
      double jxh = 0.0;
      for (int ix = mn[0]; ix <= mx[0]; ix++) {
        const double xfac1 = gx[ix] + 0.5 * hx[ix];
        const double xfac2 = third * hx[ix] + 0.5 * gx[ix];
        const double wx = hx[ix] * yzfac;
        const double wy = xfac1 * hygz + xfac2 * hyhz;
        const double wz = gx[ix] * hzyfac1 + hx[ix] * hzyfac2;
        jxh = jxh - fjx * wx;
        jyh[ix + 2] = jyh[ix + 2] - fjy * wy;
        jzh[iy + 2][ix + 2] = jzh[iy + 2][ix + 2] - fjz * wz;
        const int o = key + iz * sz + iy * sy + ix;
        // the last column / row / plane of each running sum cancels structurally (sum of hx = 0)
        if (ix < mx[0]) smem_add(&sJ[o], jxh);
        if (iy < mx[1]) smem_add(&sJ[jt + o], jyh[ix + 2]);
        if (iz < mx[2]) smem_add(&sJ[2 * jt + o], jzh[iy + 2][ix + 2]);
      }
    }
  }
}

// One-axis movers (the nearest cell changed along exactly one axis a, by s = +-1): the 3x3x3 core around the old
// cell is deposited inline with every other particle (deposit_core_3d); what is left of the 4x3x3 stencil -- 21
// values -- comes here, one particle per lane, in closed form.  With g the old and h = new - old weights,
// Esirkepov's weight of component a is w_a = h_a S(b,c), S(b,c) = g_b g_c + (h_b g_c + g_b h_c)/2 + h_b h_c/3
// (epoch3d particles.F90:620-626 for the three components), and the current is the running sum of -fj w along
// its own axis.  Beyond the core there is one more plane, at 2s along a, where g_a = 0 and h_a = E (the new weight
// there):
//   component a: its running sum has one more non-zero point than a particle that stays -- the core's third point
//     for s = +1 (value +fj_a E S(b,c), since the weights h sum to zero), the plane at -2 for s = -1 (-fj_a E S(b,c));
//   component b (and c alike) on the plane: w_b = h_b E (g_c/2 + h_c/3), summed along b; the last point cancels.
__device__ __forceinline__ void drain_edge_3d_slots(double *sJ, const double *Qd, int pk, int lane, double third) {
  const int key = pk & 4095;
  const int d0 = ((pk >> 12) & 3) - 1, d1 = ((pk >> 14) & 3) - 1, d2 = ((pk >> 16) & 3) - 1;
  const int a = d0 != 0 ? 0 : (d1 != 0 ? 1 : 2);
  const int sgn = d0 + d1 + d2;                     // exactly one of them is non-zero
  const int b = a == 2 ? 0 : a + 1, c = a == 0 ? 2 : a - 1;
  double wm, w0, wp;
  tri(Qd[(2 * a + 1) * Q3CAP + lane], wm, w0, wp);
  const double E = sgn > 0 ? wp : wm;
  double Gb[3], Hb[3], Gc[3], Hc[3];
  tri(Qd[(2 * b) * Q3CAP + lane], Gb[0], Gb[1], Gb[2]);
  tri(Qd[(2 * b + 1) * Q3CAP + lane], wm, w0, wp);
  Hb[0] = wm - Gb[0]; Hb[1] = w0 - Gb[1]; Hb[2] = wp - Gb[2];
  tri(Qd[(2 * c) * Q3CAP + lane], Gc[0], Gc[1], Gc[2]);
  tri(Qd[(2 * c + 1) * Q3CAP + lane], wm, w0, wp);
  Hc[0] = wm - Gc[0]; Hc[1] = w0 - Gc[1]; Hc[2] = wp - Gc[2];
  const double va = Qd[(6 + a) * Q3CAP + lane] * E * (double)sgn;
  const double vb = -(Qd[(6 + b) * Q3CAP + lane] * E);
  const double vc = -(Qd[(6 + c) * Q3CAP + lane] * E);
  const int st0 = 1, st1 = JB3X, st2 = JB3X * JB3Y;
  const int sa = a == 0 ? st0 : (a == 1 ? st1 : st2);
  const int sb = b == 0 ? st0 : (b == 1 ? st1 : st2);
  const int sc = c == 0 ? st0 : (c == 1 ? st1 : st2);
  double *ja = sJ + a * JB3N + key + (sgn > 0 ? 1 : -2) * sa;   // component a: third core point / the plane at -2
  double *jb = sJ + b * JB3N + key + 2 * sgn * sa;               // components b, c: the plane at 2s
  double *jc = sJ + c * JB3N + key + 2 * sgn * sa;
  double f2b[3], f2c[3];
#pragma unroll
  for (int q = 0; q < 3; q++) {
    f2b[q] = third * Hb[q] + 0.5 * Gb[q];
    f2c[q] = third * Hc[q] + 0.5 * Gc[q];
  }
#pragma unroll
  for (int ib = 0; ib < 3; ib++)
#pragma unroll
    for (int ic = 0; ic < 3; ic++) {
      const double S = Gb[ib] * (Gc[ic] + 0.5 * Hc[ic]) + Hb[ib] * f2c[ic];
      smem_add(ja + (ib - 1) * sb + (ic - 1) * sc, va * S);
    }
#pragma unroll
  for (int ic = 0; ic < 3; ic++) {
    const double t = vb * f2c[ic];
    smem_add(jb - sb + (ic - 1) * sc, t * Hb[0]);
    smem_add(jb + (ic - 1) * sc, t * (Hb[0] + Hb[1]));
  }
#pragma unroll
  for (int ib = 0; ib < 3; ib++) {
    const double t = vc * f2b[ib];
    smem_add(jc + (ib - 1) * sb - sc, t * Hc[0]);
    smem_add(jc + (ib - 1) * sb, t * (Hc[0] + Hc[1]));
  }
}

// The 3x3x3 core of the deposit on the warp's private tile (epoch3d particles.F90:603-648 on the core): G the old
// weights, H = (new weights on the core points) - G, cin the new weight one point below the core (a particle that
// moved down along that axis; it enters the running sums before the core), fj the current factors.  The last
// column / row / plane of each running sum is left out: it cancels for a particle that stays and belongs to
// drain_edge_3d_slots for one that moved up.  Convergent code: the whole warp comes here; `on` says which lanes
// deposit.  Lanes are different cells, so one (ix, iy) step never hits one address twice, but the lanes of
// neighbouring cells hit the same address in DIFFERENT steps: the load / add / store of a step must be complete
// before the next step's loads, hence the __syncwarp() between steps (an ordering point for the compiler; the
// hardware runs a warp's shared-memory instructions in order).  Inside a step the up to eight updates (three
// planes x three components) are independent.
__device__ __forceinline__ void deposit_core_3d(bool on, double *pvb, const double (&G)[3][3], const double (&H)[3][3],
                                                const double (&cin)[3], const double (&fj)[3], double third) {
  const double *gx = G[0], *gy = G[1], *gz = G[2];
  const double *hx = H[0], *hy = H[1], *hz = H[2];
  double zfac1[3], zfac2[3], xz[3][3], jyh[3][3];
#pragma unroll
  for (int q = 0; q < 3; q++) {
    zfac1[q] = gz[q] + 0.5 * hz[q];
    zfac2[q] = third * hz[q] + 0.5 * gz[q];
  }
  const double c0 = -(fj[0] * cin[0]), c1 = -(fj[1] * cin[1]), c2 = -(fj[2] * cin[2]);
#pragma unroll
  for (int ix = 0; ix < 3; ix++) {
    const double xfac1 = gx[ix] + 0.5 * hx[ix];
    const double xfac2 = third * hx[ix] + 0.5 * gx[ix];
#pragma unroll
    for (int iz = 0; iz < 3; iz++) {
      xz[iz][ix] = xfac1 * gz[iz] + xfac2 * hz[iz];
      jyh[iz][ix] = c1 * xz[iz][ix];
    }
  }
#pragma unroll
  for (int iy = 0; iy < 3; iy++) {
    const double yfac1 = gy[iy] + 0.5 * hy[iy];
    const double yfac2 = third * hy[iy] + 0.5 * gy[iy];
    double yz[3], jxh[3];
#pragma unroll
    for (int iz = 0; iz < 3; iz++) {
      yz[iz] = gy[iy] * zfac1[iz] + hy[iy] * zfac2[iz];
      jxh[iz] = c0 * yz[iz];
    }
    const double fhy = fj[1] * hy[iy];
#pragma unroll
    for (int ix = 0; ix < 3; ix++) {
      const double xy = gx[ix] * yfac1 + hx[ix] * yfac2;
      const double fhx = fj[0] * hx[ix];
      double jzh = c2 * xy;
      double *jr = pvb + iy * PV3X + ix;
#pragma unroll
      for (int iz = 0; iz < 3; iz++) {
        jxh[iz] = jxh[iz] - fhx * yz[iz];
        jyh[iz][ix] = jyh[iz][ix] - fhy * xz[iz][ix];
        jzh = jzh - (fj[2] * hz[iz]) * xy;
        if (on) {
          if (ix < 2) jr[iz * PV3Y * PV3X] += jxh[iz];
          if (iy < 2) jr[PV3N + iz * PV3Y * PV3X] += jyh[iz][ix];
          if (iz < 2) jr[2 * PV3N + iz * PV3Y * PV3X] += jzh;
        }
      }
      __syncwarp();
    }
  }
}

template <bool HC>
__global__ void __launch_bounds__(B3_THREADS, 2) push_bag_3d(const __grid_constant__ PushParams P) {
  extern __shared__ double sm[];
  double *sF = sm;                                   // [6][EB3P]
  double *sJ = sF + 6 * EB3P;                        // [3][JB3N]
  double *sPV_all = sJ + 3 * JB3N;                   // [warp][3][PV3N] private deposit tiles
  double *sQd_all = sPV_all + B3_WARPS * 3 * PV3N;
  int *sQk_all = reinterpret_cast<int *>(sQd_all + B3_WARPS * Q3DBL * Q3CAP);
  int *sAcnt_all = sQk_all + B3_WARPS * Q3CAP;       // [warp][32] arrivals per lane
  unsigned char *sAlist_all = reinterpret_cast<unsigned char *>(sAcnt_all + B3_WARPS * 32 + 2);   // [warp][32][SLOT_LK]
  const int tile = blockIdx.x;
  const int ttx = tile % P.tg.nt[0], tty = (tile / P.tg.nt[0]) % P.tg.nt[1], ttz = tile / (P.tg.nt[0] * P.tg.nt[1]);
  const int cx0 = ttx * B3X + 1, cy0 = tty * B3Y + 1, cz0 = ttz * B3Z + 1;  // first cell of the tile (1-based)
  const int ex0 = cx0 - 2, ey0 = cy0 - 2, ez0 = cz0 - 2;                   // E/B tile origin
  const int ox = cx0 - JB3H, oy = cy0 - JB3H, oz = cz0 - JB3H;             // J tile origin
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int my_key = tile * B3N + warp * 32 + lane;
  int my_cnt = P.cnt[my_key];
  if (my_cnt > P.R) my_cnt = P.R;
  const int grp = my_key >> 5;
  int inA = P.ic_in ? P.ic_in[grp] : 0;
  if (inA > P.IC) inA = P.IC;
  if (!__syncthreads_or(my_cnt > 0 || inA > 0)) return;
  for (int q = tid; q < EB3N; q += B3_THREADS) {
    const int lx = q % EB3X, ly = (q / EB3X) % EB3Y, lz = q / (EB3X * EB3Y);
    const int cx = ex0 + lx, cy = ey0 + ly, cz = ez0 + lz;
    const bool ok = (cx >= 1 - NG) && (cx <= P.n[0] + NG) && (cy >= 1 - NG) && (cy <= P.n[1] + NG) &&
                    (cz >= 1 - NG) && (cz <= P.n[2] + NG);
    const size_t o = ok ? gofs<3>(P, cx, cy, cz) : 0;
#pragma unroll
    for (int f = 0; f < 3; f++) {
      sF[f * EB3P + q] = ok ? __ldg(P.e[f] + o) : 0.0;
      sF[(3 + f) * EB3P + q] = ok ? __ldg(P.b[f] + o) : 0.0;
    }
  }
  for (int q = tid; q < 3 * JB3N + B3_WARPS * 3 * PV3N; q += B3_THREADS) sJ[q] = 0.0;   // CTA tile and the private tiles
  // ---- this warp's inbox: which entries are for which lane (entry = 8 doubles, [7] = destination lane) ----
  int *sAcnt = sAcnt_all + warp * 32;
  unsigned char *sAlist = sAlist_all + warp * 32 * SLOT_LK;
  const double *ibg = P.ib_in ? P.ib_in + (size_t)grp * (size_t)P.IC * 8 : nullptr;
  sAcnt[lane] = 0;
  __syncwarp();
  for (int j0 = 0; j0 < inA; j0 += 32) {
    const int j = j0 + lane;
    if (j < inA) {
      const int ml = (int)__double_as_longlong(ibg[(size_t)j * 8 + 7]) & 31;
      const int pos = atomicAdd(&sAcnt[ml], 1);
      if (pos < SLOT_LK) {
        sAlist[ml * SLOT_LK + pos] = (unsigned char)j;
      } else {
        const int m = atomicAdd(P.mcount, 1);
        if (m < P.mcap) {
          const double *e = ibg + (size_t)j * 8;
          P.mx[0][m] = e[0]; P.mx[1][m] = e[1]; P.mx[2][m] = e[2];
          P.mp[0][m] = e[3]; P.mp[1][m] = e[4]; P.mp[2][m] = e[5];
          P.mw[m] = e[6];
          P.mflag[m] = 0;
        } else {
          atomicOr(P.err, 1);
        }
      }
    }
  }
  __syncthreads();
  int my_a = sAcnt[lane];
  if (my_a > SLOT_LK) my_a = SLOT_LK;
  const int my_tot = my_cnt + my_a;
  const int maxcnt = __reduce_max_sync(FULL, my_tot);

  const double third = P.third;
  double *Qd = sQd_all + warp * Q3DBL * Q3CAP;
  int *Qk = sQk_all + warp * Q3CAP;
  int qcount = 0;  // warp-uniform
  const unsigned lt_mask = (1u << lane) - 1u;
  constexpr int ROWD = 7 * 32;
  double *const col = P.x[0] + ((size_t)grp * (size_t)P.R) * ROWD + lane;
  int wcur = 0;
  // this lane's cell (1-based) and the (cell - 1) corner of its 3x3x3 stencil in the warp's private tile, whose
  // origin is (cx0 - 1, first row of the warp - 1, plane of the warp - 1)
  const int hcx = cx0 + (lane & 15), hcy = cy0 + 2 * (warp & 1) + (lane >> 4), hcz = cz0 + (warp >> 1);
  double *const pvb = sPV_all + warp * 3 * PV3N + (lane >> 4) * PV3X + (lane & 15);

  double n_v[7] = {0, 0, 0, 0, 0, 0, 0};   // x y z px py pz w of the next round
  auto fetch = [&](int rn) {
    if (rn < my_cnt) {
      const double *row = col + (size_t)rn * ROWD;
#pragma unroll
      for (int q = 0; q < 7; q++) n_v[q] = row[q * 32];
    } else if (rn < my_tot) {
      const double2 *e = reinterpret_cast<const double2 *>(ibg + (size_t)sAlist[lane * SLOT_LK + (rn - my_cnt)] * 8);
      const double2 v0 = e[0], v1 = e[1], v2 = e[2], v3 = e[3];
      n_v[0] = v0.x; n_v[1] = v0.y; n_v[2] = v1.x; n_v[3] = v1.y; n_v[4] = v2.x; n_v[5] = v2.y; n_v[6] = v3.x;
    }
  };
  fetch(0);
  for (int r = 0; r < maxcnt; r++) {
    const bool active = r < my_tot;
    double o_v[6] = {n_v[0], n_v[1], n_v[2], n_v[3], n_v[4], n_v[5]};   // what is written back / sent on
    const double part_weight = n_v[6];
    double pp[3] = {n_v[0] - P.grid_min_local[0], n_v[1] - P.grid_min_local[1], n_v[2] - P.grid_min_local[2]};
    double uu[3] = {n_v[3] * P.ipart_mc, n_v[4] * P.ipart_mc, n_v[5] * P.ipart_mc};
    fetch(r + 1);
    int disp = 0, dir = -1, ib_slot = -1, ntile = 0, nlane = 0;
    bool touched = false, extras = false, inl = false;
    int qkey, dcell[3];
    double q_f[6], fj[3] = {0.0, 0.0, 0.0};
    // inputs of the core deposit (deposit_core_3d runs in convergent code after this block)
    double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Hd[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, cin[3] = {0, 0, 0};
    if (active) {
      double root;
      gamma_root(uu[0] * uu[0] + uu[1] * uu[1] + uu[2] * uu[2] + 1.0, P.dtco2, root);
      int c1[3], c2[3];
      double H[3][3], cr[3];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        pp[d] = pp[d] + uu[d] * root;
        cr[d] = pp[d] * P.idx[d];
        c1[d] = __double2int_rd(cr[d] + 0.5) + 1;
      }
      const bool fast = (c1[0] >= cx0) && (c1[0] < cx0 + B3X) && (c1[1] >= cy0) && (c1[1] < cy0 + B3Y) &&
                        (c1[2] >= cz0) && (c1[2] < cz0 + B3Z);
      if (!fast) {
        disp = 2;
      } else {
        double fo[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          fo[d] = (double)(c1[d] - 1) - cr[d];
          tri_s(fo[d], G[d][0], G[d][1], G[d][2]);
          const int c2d = __double2int_rd(cr[d]);
          tri_s((double)c2d - cr[d] + 0.5, H[d][0], H[d][1], H[d][2]);
          c2[d] = c2d + 1;
        }
        // shared E/B tile offsets of (cell - 1) along every axis, for the nearest (1) and the staggered (2) cell
        const int a1[3] = {c1[0] - 1 - ex0, (c1[1] - 1 - ey0) * EB3X, (c1[2] - 1 - ez0) * EB3X * EB3Y};
        const int a2[3] = {c2[0] - 1 - ex0, (c2[1] - 1 - ey0) * EB3X, (c2[2] - 1 - ez0) * EB3X * EB3Y};
        auto gat = [&](const double *F, int o, const double *wx, const double *wy, const double *wz) {
          double acc = 0.0;
#pragma unroll
          for (int iz = 0; iz < 3; iz++) {
            double pl = 0.0;
#pragma unroll
            for (int iy = 0; iy < 3; iy++) {
              const double *row = F + o + iz * EB3X * EB3Y + iy * EB3X;
              const double rs = wx[0] * row[0] + wx[1] * row[1] + wx[2] * row[2];
              pl = (iy == 0) ? wy[iy] * rs : pl + wy[iy] * rs;
            }
            acc = (iz == 0) ? wz[iz] * pl : acc + wz[iz] * pl;
          }
          return acc;
        };
        // include/triangle/e_part.inc, b_part.inc: ex(hx,gy,gz) ey(gx,hy,gz) ez(gx,gy,hz) bx(gx,hy,hz) by(hx,gy,hz) bz(hx,hy,gz)
        const double ex_part = gat(sF + 0 * EB3P, a2[0] + a1[1] + a1[2], H[0], G[1], G[2]);
        const double ey_part = gat(sF + 1 * EB3P, a1[0] + a2[1] + a1[2], G[0], H[1], G[2]);
        const double ez_part = gat(sF + 2 * EB3P, a1[0] + a1[1] + a2[2], G[0], G[1], H[2]);
        const double bx_part = gat(sF + 3 * EB3P, a1[0] + a2[1] + a2[2], G[0], H[1], H[2]);
        const double by_part = gat(sF + 4 * EB3P, a2[0] + a1[1] + a2[2], H[0], G[1], H[2]);
        const double bz_part = gat(sF + 5 * EB3P, a2[0] + a2[1] + a1[2], H[0], H[1], G[2]);
        const double cmratio = P.cmratio;
        const double uxm = uu[0] + cmratio * ex_part;
        const double uym = uu[1] + cmratio * ey_part;
        const double uzm = uu[2] + cmratio * ez_part;
        double gm2 = uxm * uxm + uym * uym + uzm * uzm + 1.0;
        if (HC) {  // particles.F90:386-398 (-DHC_PUSH), Higuera & Cary, Phys. Plasmas 24, 052104; the result is >= 1 too
          const double beta_x = P.hc_alpha * bx_part, beta_y = P.hc_alpha * by_part, beta_z = P.hc_alpha * bz_part;
          const double beta2 = beta_x * beta_x + beta_y * beta_y + beta_z * beta_z;
          const double sigma = gm2 - beta2;
          const double beta_dot_u = beta_x * uxm + beta_y * uym + beta_z * uzm;
          gm2 = 0.5 * (sigma + sqrt(sigma * sigma + 4.0 * (beta2 + beta_dot_u * beta_dot_u)));
        }
        gamma_root(gm2, P.ccmratio, root);
        const double taux = bx_part * root, tauy = by_part * root, tauz = bz_part * root;
        const double taux2 = taux * taux, tauy2 = tauy * tauy, tauz2 = tauz * tauz;
#ifdef EPB_FAST_MATH
        const double tau = rcp_ge1(1.0 + taux2 + tauy2 + tauz2);
#else
        const double tau = 1.0 / (1.0 + taux2 + tauy2 + tauz2);
#endif
        const double uxp = ((1.0 + taux2 - tauy2 - tauz2) * uxm +
                            2.0 * ((taux * tauy + tauz) * uym + (taux * tauz - tauy) * uzm)) * tau;
        const double uyp = ((1.0 - taux2 + tauy2 - tauz2) * uym +
                            2.0 * ((tauy * tauz + taux) * uzm + (tauy * taux - tauz) * uxm)) * tau;
        const double uzp = ((1.0 - taux2 - tauy2 + tauz2) * uzm +
                            2.0 * ((tauz * taux + tauy) * uxm + (tauz * tauy - taux) * uym)) * tau;
        uu[0] = uxp + cmratio * ex_part;
        uu[1] = uyp + cmratio * ey_part;
        uu[2] = uzp + cmratio * ez_part;
        // epoch3d particles.F90:470-474
        gamma_root(uu[0] * uu[0] + uu[1] * uu[1] + uu[2] * uu[2] + 1.0, P.dtco2, root);
        double delta[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          delta[d] = uu[d] * root;
          pp[d] = pp[d] + delta[d];
        }
        {
          double pos[3] = {pp[0] + P.grid_min_local[0], pp[1] + P.grid_min_local[1], pp[2] + P.grid_min_local[2]};
          double mom[3] = {P.part_mc * uu[0], P.part_mc * uu[1], P.part_mc * uu[2]};
          touched = (pos[0] < P.bnd_min[0]) || (pos[0] > P.bnd_max[0]) || (pos[1] < P.bnd_min[1]) || (pos[1] > P.bnd_max[1]) ||
                    (pos[2] < P.bnd_min[2]) || (pos[2] > P.bnd_max[2]);
          dir = particle_bc<3>(P, pos, mom);
#pragma unroll
          for (int d = 0; d < 3; d++) { o_v[d] = pos[d]; o_v[3 + d] = mom[d]; }
        }
        // the cell the next push gathers this particle in, its new shape weights for the deposit
        int c3[3];
        double fn[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          pp[d] = pp[d] + delta[d];
          const double crn = pp[d] * P.idx[d];
          c3[d] = __double2int_rd(crn + 0.5);
          fn[d] = (double)c3[d] - crn;
          dcell[d] = c3[d] + 1 - c1[d];
        }
        {
          int k3[3];
#pragma unroll
          for (int d = 0; d < 3; d++) k3[d] = c3[d] < 0 ? 0 : (c3[d] > P.n[d] - 1 ? P.n[d] - 1 : c3[d]);
          ntile = ((k3[2] / B3Z) * P.tg.nt[1] + (k3[1] / B3Y)) * P.tg.nt[0] + (k3[0] / B3X);
          // in-tile index of the next gather cell: 32 consecutive ones are 16 x 2 cells of one z plane
          const int nin = ((k3[2] % B3Z) * B3Y + (k3[1] % B3Y)) * B3X + (k3[0] % B3X);
          const bool stays = !touched && ntile == tile && nin == warp * 32 + lane;
          disp = (dir == 13) ? 3 : (stays ? 0 : 1);
          if (disp == 1 && dir < 0 && !touched && P.ic_out) {
            // reserve the entry in the destination cell's group inbox now, look at the reply after the deposit
            ntile = ntile * B3_WARPS + (nin >> 5);
            nlane = nin & 31;
            ib_slot = atomicAdd(&P.ic_out[ntile], 1);
          }
        }
        if (P.deposit) {
          const double fcx = P.kfc[0] * part_weight, fcy = P.kfc[1] * part_weight, fcz = P.kfc[2] * part_weight;
          fj[0] = fcx * P.part_q;
          fj[1] = fcy * P.part_q;
          fj[2] = fcz * P.part_q;
          const int key = ((c1[2] - oz) * JB3Y + (c1[1] - oy)) * JB3X + (c1[0] - ox);
          const int nmov = (dcell[0] != 0) + (dcell[1] != 0) + (dcell[2] != 0);
          const bool own = c1[0] == hcx && c1[1] == hcy && c1[2] == hcz;
          qkey = key | ((dcell[0] + 1) << 12) | ((dcell[1] + 1) << 14) | ((dcell[2] + 1) << 16);
          if (nmov != 0 || !own) {
            // nearest cell moved: the stencil is wider than the core.  Along one axis the part beyond the core is
            // 21 values (drain_edge_3d_slots, bit 20 of the key); anything else (0.5 % of a thermal plasma's
            // particles, or a particle that is not in this lane's cell) takes the reference's general loop.
            extras = true;
#pragma unroll
            for (int d = 0; d < 3; d++) { q_f[2 * d] = fo[d]; q_f[2 * d + 1] = fn[d]; }
          }
          if (nmov <= 1 && own) {
            inl = true;
            if (nmov == 1) qkey |= 1 << 20;
            // new weights on the core points (shifted by the move), the carry from the point below the core
#pragma unroll
            for (int d = 0; d < 3; d++) {
              double wm, w0, wp;
              tri_s(fn[d], wm, w0, wp);
              Hd[d][0] = (dcell[d] == 0 ? wm : dcell[d] > 0 ? 0.0 : w0) - G[d][0];
              Hd[d][1] = (dcell[d] == 0 ? w0 : dcell[d] > 0 ? wm : wp) - G[d][1];
              Hd[d][2] = (dcell[d] == 0 ? wp : dcell[d] > 0 ? w0 : 0.0) - G[d][2];
              cin[d] = dcell[d] < 0 ? wm : 0.0;
            }
          }
        }
      }
    }
    if (P.deposit && __any_sync(FULL, inl)) deposit_core_3d(inl, pvb, G, Hd, cin, fj, third);
    // ---- particles that leave this column (see push_slots_2d) ----
    bool toM = (disp == 2) || (disp == 1 && (dir >= 0 || touched));
    if (disp == 1 && !toM) {
      if (ib_slot >= 0 && ib_slot < P.IC) {
        double2 *e = reinterpret_cast<double2 *>(P.ib_out + ((size_t)ntile * (size_t)P.IC + ib_slot) * 8);
        e[0] = make_double2(o_v[0], o_v[1]);
        e[1] = make_double2(o_v[2], o_v[3]);
        e[2] = make_double2(o_v[4], o_v[5]);
        e[3] = make_double2(part_weight, __longlong_as_double((long long)nlane));
      } else {
        toM = true;
      }
    }
    if (active && disp == 0 && wcur >= P.R) toM = true;
    {
      const unsigned bal = __ballot_sync(FULL, toM);
      if (bal) {
        int base = 0;
        if (lane == __ffs(bal) - 1) base = atomicAdd(P.mcount, __popc(bal));
        base = __shfl_sync(FULL, base, __ffs(bal) - 1);
        if (toM) {
          const int m = base + __popc(bal & lt_mask);
          if (m < P.mcap) {
            P.mx[0][m] = o_v[0]; P.mx[1][m] = o_v[1]; P.mx[2][m] = o_v[2];
            P.mp[0][m] = o_v[3]; P.mp[1][m] = o_v[4]; P.mp[2][m] = o_v[5];
            P.mw[m] = part_weight;
            P.mflag[m] = (disp == 2) ? 2 : (dir >= 0 ? 1 : 0);
            if (disp == 1 && dir >= 0) {
              const int slot = atomicAdd(&P.out_count[dir], 1);
              if (slot < P.out_cap) P.out_idx[(size_t)dir * P.out_cap + slot] = m;
            }
            if (disp == 0) disp = 1;
          } else if (disp == 1 && dir < 0 && wcur < P.R) {
            disp = 0;
          } else {
            atomicOr(P.err, 1);
            disp = 3;
          }
        }
      }
    }
    if (active && disp == 0) {
      double *row = col + (size_t)wcur * ROWD;
#pragma unroll
      for (int q = 0; q < 6; q++) row[q * 32] = o_v[q];
      if (wcur != r || r >= my_cnt) row[6 * 32] = part_weight;
      wcur++;
    }
    if (!P.deposit) continue;
    // ---- queue the particles whose nearest cell moved ----
    const unsigned em = __ballot_sync(FULL, extras);
    if (em) {
      // the queue holds Q3CAP entries: a round's extras go in in two halves if need be
      for (int half = 0; half < 2; half++) {
        const unsigned hm = half == 0 ? (em & 0xffffu) : (em & 0xffff0000u);
        if (!hm) continue;
        const int ne = __popc(hm);
        if (qcount + ne > Q3CAP) {
          __syncwarp();
          drain_general_3d(P, sJ, Qd, Qk, qcount, lane, JB3X, JB3X * JB3Y, JB3N);
          __syncwarp();
          qcount = 0;
        }
        if (extras && ((hm >> lane) & 1u)) {
          const int slot = qcount + __popc(hm & lt_mask);
          Qk[slot] = qkey;
#pragma unroll
          for (int q = 0; q < 6; q++) Qd[q * Q3CAP + slot] = q_f[q];
          Qd[6 * Q3CAP + slot] = fj[0]; Qd[7 * Q3CAP + slot] = fj[1]; Qd[8 * Q3CAP + slot] = fj[2];
        }
        qcount += ne;
      }
    }
  }
  if (qcount) {
    __syncwarp();
    drain_general_3d(P, sJ, Qd, Qk, qcount, lane, JB3X, JB3X * JB3Y, JB3N);
  }
  if (my_tot > 0) P.cnt[my_key] = wcur;
  if (P.deposit && maxcnt > 0) {
    // the warp's private tile joins the CTA tile: point (px, py, pz) of it is CTA-tile point
    // (px + JB3H - 1, py + first row of the warp + JB3H - 1, pz + plane of the warp + JB3H - 1)
    __syncwarp();
    const double *pv = sPV_all + warp * 3 * PV3N;
    const int jb = (((warp >> 1) + JB3H - 1) * JB3Y + 2 * (warp & 1) + JB3H - 1) * JB3X + JB3H - 1;
    for (int q = lane; q < PV3N; q += 32) {
      const int px = q % PV3X, py = (q / PV3X) % PV3Y, pz = q / (PV3X * PV3Y);
      const int o = jb + (pz * JB3Y + py) * JB3X + px;
#pragma unroll
      for (int f = 0; f < 3; f++) {
        const double v = pv[f * PV3N + q];
        if (v != 0.0) smem_add(&sJ[f * JB3N + o], v);
      }
    }
  }
  __syncthreads();
  if (!P.deposit) return;
  for (int q = tid; q < JB3N; q += B3_THREADS) {
    const int lx = q % JB3X, ly = (q / JB3X) % JB3Y, lz = q / (JB3X * JB3Y);
    const int cx = ox + lx, cy = oy + ly, cz = oz + lz;
    const bool ok = (cx >= 1 - NG) && (cx <= P.n[0] + NG) && (cy >= 1 - NG) && (cy <= P.n[1] + NG) &&
                    (cz >= 1 - NG) && (cz <= P.n[2] + NG);
    if (!ok) continue;
    const size_t o = gofs<3>(P, cx, cy, cz);
#pragma unroll
    for (int f = 0; f < 3; f++) {
      const double val = sJ[f * JB3N + q];
      if (val != 0.0) atomicAdd(P.j[f] + o, val);
    }
  }
}


inline void launch_push(const PushParams &P, int nd, bool tiled, cudaStream_t s, long long *launches) {
  static bool attr_set = false;
  static int variant = 0;
  if (tiled && nd == 2) {
    if (!attr_set) {
      cudaFuncSetAttribute(push_tiled_2d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PUSH2D_SMEM);
      cudaFuncSetAttribute(push_tiled_2d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PUSH2D_SMEM);
      cudaFuncSetAttribute(push_cell_2d<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pushcell_smem<16>());
      cudaFuncSetAttribute(push_cell_2d<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pushcell_smem<16>());
      cudaFuncSetAttribute(push_cell_2d<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pushcell_smem<8>());
      variant = epb_push_variant();
      attr_set = true;
    }
    if (P.tg.ntiles > 0) {
      if (P.tg.layout == 2) {
        static bool attr_s = false;
        if (!attr_s) {
          cudaFuncSetAttribute(push_slots_2d<8, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pushslots_smem<8>());
          cudaFuncSetAttribute(push_slots_2d<8, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pushslots_smem<8>());
          cudaFuncSetAttribute(push_slots_2d<8, 3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pushslots_smem<8>());
          cudaFuncSetAttribute(push_slots_2d<8, 3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pushslots_smem<8>());
          attr_s = true;
        }
        if (P.hc_push) {   // Higuera-Cary rotation: the same kernels with the other gamma (particles.F90:386-398)
          if (P.rowd == 32) push_slots_2d<8, 3, false, true><<<P.tg.ntiles, 128, pushslots_smem<8>(), s>>>(P);
          else push_slots_2d<8, 3, true, true><<<P.tg.ntiles, 128, pushslots_smem<8>(), s>>>(P);
        } else if (P.rowd == 32) push_slots_2d<8, 3, false><<<P.tg.ntiles, 128, pushslots_smem<8>(), s>>>(P);
        else push_slots_2d<8, 3, true><<<P.tg.ntiles, 128, pushslots_smem<8>(), s>>>(P);
      } else if (P.tg.layout == 1) {
        if (P.tg.T[1] == 8) push_cell_2d<8, 3><<<P.tg.ntiles, 128, pushcell_smem<8>(), s>>>(P);
        else if (variant == 4) push_cell_2d<16, 1><<<P.tg.ntiles, 256, pushcell_smem<16>(), s>>>(P);
        else push_cell_2d<16, 2><<<P.tg.ntiles, 256, pushcell_smem<16>(), s>>>(P);
      }
      else if (variant == 1) push_tiled_2d<true><<<P.tg.ntiles, PUSH2D_THREADS, PUSH2D_SMEM, s>>>(P);
      else push_tiled_2d<false><<<P.tg.ntiles, PUSH2D_THREADS, PUSH2D_SMEM, s>>>(P);
      (*launches)++;
    }
    return;
  }
  if (tiled && nd == 3) {
    static bool attr3 = false;
    if (!attr3) {
      cudaFuncSetAttribute(push_tiled_3d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PUSH3D_SMEM);
      attr3 = true;
    }
    if (P.tg.ntiles > 0) {
      if (P.tg.layout == 3) {
        static bool attrb = false;
        if (!attrb) {
          cudaFuncSetAttribute(push_bag_3d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PUSHBAG3D_SMEM);
          cudaFuncSetAttribute(push_bag_3d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PUSHBAG3D_SMEM);
          attrb = true;
        }
        if (P.hc_push) push_bag_3d<true><<<P.tg.ntiles, B3_THREADS, PUSHBAG3D_SMEM, s>>>(P);
        else push_bag_3d<false><<<P.tg.ntiles, B3_THREADS, PUSHBAG3D_SMEM, s>>>(P);
      } else {
        push_tiled_3d<<<P.tg.ntiles, P3_THREADS, PUSH3D_SMEM, s>>>(P);
      }
      (*launches)++;
    }
    return;
  }
  const long long cnt = P.last - P.first;
  if (cnt <= 0) return;
  long long blocks = (cnt + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  if (P.hc_push) {
    if (nd == 1) push_generic<1, true><<<(int)blocks, 256, 0, s>>>(P);
    else if (nd == 2) push_generic<2, true><<<(int)blocks, 256, 0, s>>>(P);
    else push_generic<3, true><<<(int)blocks, 256, 0, s>>>(P);
  } else if (nd == 1) push_generic<1><<<(int)blocks, 256, 0, s>>>(P);
  else if (nd == 2) push_generic<2><<<(int)blocks, 256, 0, s>>>(P);
  else push_generic<3><<<(int)blocks, 256, 0, s>>>(P);
  (*launches)++;
}

// layout 2: the unpushed entries of the mover buffer (P.x/p/w/gone point into it)
inline void launch_push_m(const PushParams &P, cudaStream_t s, long long *launches) {
  int blocks = (P.mcap + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (P.hc_push) {
    if (P.nd == 3) push_generic_m<3, true><<<blocks, 256, 0, s>>>(P);
    else push_generic_m<2, true><<<blocks, 256, 0, s>>>(P);
  } else if (P.nd == 3) push_generic_m<3><<<blocks, 256, 0, s>>>(P);
  else push_generic_m<2><<<blocks, 256, 0, s>>>(P);
  (*launches)++;
}

}  // namespace EPB_NS
