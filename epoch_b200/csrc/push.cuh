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
// Two kernels:
//  * push_generic<ND>: any particle range; gathers E/B through the read-only path
//    and deposits with native global FP64 reductions (RED.E.ADD.F64).  Used for 1D
//    and 3D, for the unsorted tail (arrivals since the last sort) and as the
//    per-particle fallback of the tiled kernel.
//  * push_tiled_2d: one CTA per 16x16-cell tile of the cell-sorted layout.  The
//    E/B tile (+3 halo cells) is staged in shared memory, J is accumulated in a
//    shared tile and flushed once with global reductions.  Because the particle
//    range is cell-ordered, the 32 lanes of a warp mostly sit in one or two cells:
//    their 3x3x3 deposit values are summed across the warp with a transposing
//    butterfly (31 shuffles) and only 27 lanes issue one shared-memory update
//    each, instead of 27 CAS loops per particle (shared FP64 atomicAdd is an
//    ATOMS.CAST.SPIN loop on sm_100a).
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
template <int ND>
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
  gamma_rel = sqrt(uxm * uxm + uym * uym + uzm * uzm + 1.0);
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

template <int ND>
__global__ void __launch_bounds__(256) push_generic(const __grid_constant__ PushParams P) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = P.first + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.last; i += stride)
    push_one<ND>(P, i);
}

// ---------------------------------------------------------------------------
// Tiled 2D kernel
// ---------------------------------------------------------------------------
// One CTA per 16x16-cell tile of the cell-sorted layout, two CTAs per SM.  The kernel is
// bound by the SM's shared-memory data path (wavefronts) and the FP64 pipe, not by HBM
// (tools/microbench.cu, profiles/).  Shared memory:
//   sF  [6][TH][TW]    E/B tile + 3 halo cells
//   sJ  [3] padded     current accumulated by this CTA, flushed once with global reductions
//   sS  [warp][27][33] per-warp transposition scratch for the deposit reduction
//   sQ  [warp][4][32]  double2: per-warp queue of particles whose nearest cell changed
//   sSlow              CTA list of particles outside the tile's halo (stale sort, wrapped)
// The 32 lanes of a warp hold 32 consecutive particles of the sorted range, i.e. mostly one
// or two cells.  Each lane writes its 27 deposit values (3 components x 3x3 cells around its
// nearest cell) into one column of sS; lane q < 27 then sums row q over the columns that
// share a cell key and issues ONE shared-memory update per (key, value) (shared FP64
// atomicAdd is a per-lane serialised CAS loop on sm_100a, so updates per particle are what
// must be avoided).  The partial sum of the last key is carried in registers into the next
// batch.  The cost is independent of how many distinct cells the warp spans, which keeps the
// kernel efficient between sorts.
// Particles whose nearest cell changed during the step (a few %) have a wider stencil: they
// are queued and deposited densely, 32 at a time, with the reference's general loop.
constexpr int T2X = 16, T2Y = 16, HALO = 3;
constexpr int TW = T2X + 2 * HALO, TH = T2Y + 2 * HALO;
constexpr int TILE_ELEMS = TW * TH;
constexpr int PUSH2D_THREADS = 256, PUSH2D_WARPS = PUSH2D_THREADS / 32;
constexpr int SROWS = 27, SPITCH = 34;  // even pitch: rows are read two columns at a time
// sJ is padded (row pitch 29, component stride 649 doubles) so that the 27 addresses of one
// flush fall into distinct 8-byte banks at most twice
constexpr int JP = 29, JC = 633;
constexpr int QCAP = 32;
constexpr int SLOWCAP = 254;
constexpr size_t PUSH2D_SMEM =
    sizeof(double2) * ((size_t)PUSH2D_WARPS * 4 * QCAP + 2 * TILE_ELEMS) +
    sizeof(double) * ((size_t)2 * TILE_ELEMS + 3 * JC + 1 + (size_t)PUSH2D_WARPS * SROWS * SPITCH) + sizeof(int) * (SLOWCAP + 2);

__device__ __forceinline__ void smem_add(double *addr, double v) {
  // shared FP64 add: an ATOMS.CAST.SPIN loop; conflicts between warps are rare
  atomicAdd(addr, v);
}
#define SMEM_ADD(addr, v) do { if (!(P.experiment & 1)) smem_add((addr), (v)); else if ((v) == 1.2345e300) *(addr) = 0.0; } while (0)

// num / sqrt(s).  The parity build keeps the reference's sqrt + divide sequence.
__device__ __forceinline__ double over_sqrt(double num, double s) {
#ifdef EPB_FAST_MATH
  return num * rsqrt(s);
#else
  return num / sqrt(s);
#endif
}

// Deposit of queued particles (nearest cell changed: dcell != 0 in x and/or y): the general
// loop of particles.F90:549-579 over xmin..xmax, ymin..ymax with shared-memory updates.
__device__ __forceinline__ void drain_extras(const PushParams &P, double *sJ, const double2 *Q, int n, int lane) {
  if (lane >= n) return;
  const double2 qx = Q[0 * QCAP + lane], qy = Q[1 * QCAP + lane], qj = Q[2 * QCAP + lane], qz = Q[3 * QCAP + lane];
  const int pk = __double2loint(qz.y);
  const int key = pk & 1023, dcx = ((pk >> 10) & 3) - 1, dcy = ((pk >> 12) & 3) - 1;
  const double fjx = qj.x, fjy = qj.y, fjz = qz.x;
  double gx[5], gy[5], hx[5], hy[5];
  gx[0] = gx[4] = gy[0] = gy[4] = 0.0;
  tri(qx.x, gx[1], gx[2], gx[3]);
  tri(qy.x, gy[1], gy[2], gy[3]);
  double wm, w0, wp;
  tri(qx.y, wm, w0, wp);
#pragma unroll
  for (int q = 0; q < 5; q++) {
    const int r = q - 2 - dcx;
    hx[q] = ((r == -1) ? wm : (r == 0) ? w0 : (r == 1) ? wp : 0.0) - gx[q];
  }
  tri(qy.y, wm, w0, wp);
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
      const int o = key + iy * JP + ix;
      smem_add(&sJ[o], jxh);
      smem_add(&sJ[JC + o], jyh[ix + 2]);
      smem_add(&sJ[2 * JC + o], jzh);
    }
  }
}

// sum of row `row` over columns [lo, hi); the row is 16-byte aligned and read with LDS.128
__device__ __forceinline__ double row_sum(const double *row, int lo, int hi) {
  double a = 0.0, b = 0.0;
  int j = lo;
  if ((j & 1) && j < hi) { a = row[j]; j++; }
  for (; j + 2 <= hi; j += 2) {
    const double2 u = *reinterpret_cast<const double2 *>(row + j);
    a += u.x;
    b += u.y;
  }
  if (j < hi) a += row[j];
  return a + b;
}

template <bool CHUNK, bool CARRY>
__global__ void __launch_bounds__(PUSH2D_THREADS, 2) push_tiled_2d(const __grid_constant__ PushParams P) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double2 *sQ_all = reinterpret_cast<double2 *>(smraw);         // [warps][4][32]
  double2 *sEB1 = sQ_all + PUSH2D_WARPS * 4 * QCAP;             // (ex, by) [TH][TW]
  double2 *sEB2 = sEB1 + TILE_ELEMS;                            // (ey, bx)
  double *sS_all = reinterpret_cast<double *>(sEB2 + TILE_ELEMS);   // [warps][27][34], 16-byte aligned rows
  double *sEz = sS_all + PUSH2D_WARPS * SROWS * SPITCH;
  double *sBz = sEz + TILE_ELEMS;
  double *sJ = sBz + TILE_ELEMS;                                // [3] padded tiles
  int *sSlow = reinterpret_cast<int *>(sJ + 3 * JC + 1);
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
    sEB1[q] = ok ? make_double2(__ldg(P.e[0] + o), __ldg(P.b[1] + o)) : make_double2(0.0, 0.0);
    sEB2[q] = ok ? make_double2(__ldg(P.e[1] + o), __ldg(P.b[0] + o)) : make_double2(0.0, 0.0);
    sEz[q] = ok ? __ldg(P.e[2] + o) : 0.0;
    sBz[q] = ok ? __ldg(P.b[2] + o) : 0.0;
  }
  for (int q = tid; q < 3 * JC; q += PUSH2D_THREADS) sJ[q] = 0.0;
  __syncthreads();

  const double c = EPB_C;
  const double third = P.third;
  double *S = sS_all + warp * SROWS * SPITCH;
  double2 *Q = sQ_all + warp * 4 * QCAP;
  int qcount = 0;  // warp-uniform
  // lane q < 27 owns deposit value q = comp*9 + iy*3 + ix of the 3x3 stencil
  const int offq = (lane / 9) * JC + ((lane % 9) / 3 - 1) * JP + (lane % 3 - 1);
  const bool owner = lane < SROWS;
  const double *row = S + (owner ? lane : 0) * SPITCH;
  const unsigned lt_mask = (1u << lane) - 1u;

  // each warp streams its own contiguous eighth of the tile's (cell-ordered) range, so that
  // concurrent warps work on different cells and consecutive batches of a warp share cells
  long long wend;
  long long i;
  if (CHUNK) {
    const long long total = end - start;
    const long long chunk = ((total + PUSH2D_WARPS - 1) / PUSH2D_WARPS + 31) / 32 * 32;
    const long long wstart = start + warp * chunk;
    wend = wstart + chunk < end ? wstart + chunk : end;
    i = wstart + lane;
  } else {  // batches interleaved between the warps
    wend = end;
    i = start + warp * 32 + lane;
  }
  constexpr int STEP = CHUNK ? 32 : PUSH2D_THREADS;
  // carried partial sum of the reduction: the lane's value for cell key ck (warp-uniform)
  int ck = -1;
  double ca = 0.0;
  // software pipeline: the next batch's particle loads are in flight while this one computes
  double n_x = 0, n_y = 0, n_px = 0, n_py = 0, n_pz = 0, n_w = 0;
  if (i < wend) {
    n_w = P.w[i]; n_x = P.x[0][i]; n_y = P.x[1][i];
    n_px = P.p[0][i]; n_py = P.p[1][i]; n_pz = P.p[2][i];
  }
  for (; i - lane < wend; i += STEP) {
    const bool active = i < wend;
    const double part_weight = n_w;
    double px_ = n_x - P.grid_min_local[0];
    double py_ = n_y - P.grid_min_local[1];
    double part_ux = n_px * P.ipart_mc;
    double part_uy = n_py * P.ipart_mc;
    double part_uz = n_pz * P.ipart_mc;
    {
      const long long in = i + STEP;
      if (in < wend) {
        n_w = P.w[in]; n_x = P.x[0][in]; n_y = P.x[1][in];
        n_px = P.p[0][in]; n_py = P.p[1][in]; n_pz = P.p[2][in];
      }
    }
    int key = -1;       // cell key (offset of the nearest cell in the shared tile)
    bool dep = false;   // lane takes part in the transposed reduction
    bool extras = false;
    int dcx = 0, dcy = 0;
    double fxo = 0, fxn = 0, fyo = 0, fyn = 0, fjx = 0, fjy = 0, fjz = 0;
    if (active) {
      double root = over_sqrt(P.dtco2, part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0);
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
        fxo = (double)(cx1 - 1) - cell_x_r;
        fyo = (double)(cy1 - 1) - cell_y_r;
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
        // include/triangle/e_part.inc, b_part.inc: rows parenthesised, sums left to right
        auto gat = [&](const double *F, int o, const double *wx, const double *wy) {
          const double r0 = wx[0] * F[o] + wx[1] * F[o + 1] + wx[2] * F[o + 2];
          const double r1 = wx[0] * F[o + TW] + wx[1] * F[o + TW + 1] + wx[2] * F[o + TW + 2];
          const double r2 = wx[0] * F[o + 2 * TW] + wx[1] * F[o + 2 * TW + 1] + wx[2] * F[o + 2 * TW + 2];
          return wy[0] * r0 + wy[1] * r1 + wy[2] * r2;
        };
        // (ex, by) share the weights (hx, gy) and the offset (cell_x2, cell_y1), (ey, bx) share
        // (gx, hy) at (cell_x1, cell_y2): one 16-byte shared load fetches both
        auto gat2 = [&](const double2 *F, int o, const double *wx, const double *wy, double &ra, double &rb) {
          double a[3], b[3];
#pragma unroll
          for (int r = 0; r < 3; r++) {
            const double2 f0 = F[o + r * TW], f1 = F[o + r * TW + 1], f2 = F[o + r * TW + 2];
            a[r] = wx[0] * f0.x + wx[1] * f1.x + wx[2] * f2.x;
            b[r] = wx[0] * f0.y + wx[1] * f1.y + wx[2] * f2.y;
          }
          ra = wy[0] * a[0] + wy[1] * a[1] + wy[2] * a[2];
          rb = wy[0] * b[0] + wy[1] * b[1] + wy[2] * b[2];
        };
        double ex_part, ey_part, bx_part, by_part;
        gat2(sEB1, (P.experiment & 4) ? 50 : o21, hx, gy, ex_part, by_part);
        gat2(sEB2, (P.experiment & 4) ? 50 : o12, gx, hy, ey_part, bx_part);
        const double ez_part = gat(sEz, o11, gx, gy);
        const double bz_part = gat(sBz, o22, hx, hy);
        const double cmratio = P.cmratio;
        const double uxm = part_ux + cmratio * ex_part;
        const double uym = part_uy + cmratio * ey_part;
        const double uzm = part_uz + cmratio * ez_part;
        root = over_sqrt(P.ccmratio, uxm * uxm + uym * uym + uzm * uzm + 1.0);
        const double taux = bx_part * root, tauy = by_part * root, tauz = bz_part * root;
        const double taux2 = taux * taux, tauy2 = tauy * tauy, tauz2 = tauz * tauz;
        const double tau = 1.0 / (1.0 + taux2 + tauy2 + tauz2);
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
        const double igamma = over_sqrt(1.0, part_u2 + 1.0);
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
          fxn = (double)cx3 - cxr;
          fyn = (double)cy3 - cyr;
          dcx = cx3 + 1 - cx1;
          dcy = cy3 + 1 - cy1;
          const double fcx = P.kfc[0] * part_weight;
          const double fcy = P.kfc[1] * part_weight;
          const double fcz = P.kfc[2] * part_weight;
          fjx = fcx * P.part_q;
          fjy = fcy * P.part_q;
          fjz = fcz * P.part_q * part_vz;
          key = (cy1 - oy) * JP + (cx1 - ox);
          extras = (dcx | dcy) != 0;
          dep = !extras;
        }
      }
    }
    if (!P.deposit) continue;
    // ---- queue the particles with a wider stencil -------------------------------------
    const unsigned em = __ballot_sync(FULL, extras && !(P.experiment & 8));
    if (em) {
      const int ne = __popc(em);
      if (qcount + ne > QCAP) {
        __syncwarp();
        drain_extras(P, sJ, Q, qcount, lane);
        __syncwarp();
        qcount = 0;
      }
      if (extras) {
        const int slot = qcount + __popc(em & lt_mask);
        Q[0 * QCAP + slot] = make_double2(fxo, fxn);
        Q[1 * QCAP + slot] = make_double2(fyo, fyn);
        Q[2 * QCAP + slot] = make_double2(fjx, fjy);
        Q[3 * QCAP + slot] = make_double2(fjz, __hiloint2double(0, key | ((dcx + 1) << 10) | ((dcy + 1) << 12)));
      }
      qcount += ne;
    }
    // ---- transposed reduction: one shared update per (cell key, stencil value) ------------
    // Lanes are grouped by cell key: group 0 = key of the first lane, group 1 = the next key,
    // the rest (further cells; stale sort) are handled one by one.  Columns of sS are assigned
    // group by group so that the reducing lanes sum plain index ranges.
    unsigned rest = __ballot_sync(FULL, dep && !(P.experiment & 2));
    if (rest) {
      const int k0 = __shfl_sync(FULL, key, __ffs(rest) - 1);
      const unsigned m0 = __ballot_sync(FULL, dep && key == k0);
      rest &= ~m0;
      int k1 = -1;
      unsigned m1 = 0;
      if (rest) {
        k1 = __shfl_sync(FULL, key, __ffs(rest) - 1);
        m1 = __ballot_sync(FULL, dep && key == k1);
        rest &= ~m1;
      }
      const int n0 = __popc(m0), n1 = __popc(m1);
      if (dep) {
        const unsigned me = 1u << lane;
        const int colidx = (m0 & me) ? __popc(m0 & lt_mask)
                         : (m1 & me) ? n0 + __popc(m1 & lt_mask)
                                     : n0 + n1 + __popc(rest & lt_mask);
        // dcell = 0: hx = new weights - gx on the same three cells (particles.F90:521-538)
        double gx[3], gy[3], hx[3], hy[3];
        tri(fxo, gx[0], gx[1], gx[2]);
        tri(fyo, gy[0], gy[1], gy[2]);
        tri(fxn, hx[0], hx[1], hx[2]);
        tri(fyn, hy[0], hy[1], hy[2]);
#pragma unroll
        for (int q = 0; q < 3; q++) { hx[q] = hx[q] - gx[q]; hy[q] = hy[q] - gy[q]; }
        double xfac1[3], jyh[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int ix = 0; ix < 3; ix++) xfac1[ix] = gx[ix] + 0.5 * hx[ix];
        double *col = S + colidx;
#pragma unroll
        for (int iy = 0; iy < 3; iy++) {
          const double yfac1 = gy[iy] + 0.5 * hy[iy];
          const double yfac2 = third * hy[iy] + 0.5 * gy[iy];
          double jxh = 0.0;
#pragma unroll
          for (int ix = 0; ix < 3; ix++) {
            const double wx = hx[ix] * yfac1;
            const double wy = hy[iy] * xfac1[ix];
            const double wz = gx[ix] * yfac1 + hx[ix] * yfac2;
            jxh = jxh - fjx * wx;
            jyh[ix] = jyh[ix] - fjy * wy;
            col[(iy * 3 + ix) * SPITCH] = jxh;
            col[(9 + iy * 3 + ix) * SPITCH] = jyh[ix];
            col[(18 + iy * 3 + ix) * SPITCH] = fjz * wz;
          }
        }
      }
      __syncwarp();
      {
        const double a0 = row_sum(row, 0, n0);
        if (CARRY && k0 == ck) ca += a0;
        else {
          if (ck >= 0 && owner) SMEM_ADD(&sJ[offq + ck], ca);
          ck = k0;
          ca = a0;
        }
      }
      if (n1) {
        const double a1 = row_sum(row, n0, n0 + n1);
        if (owner) SMEM_ADD(&sJ[offq + ck], ca);
        ck = k1;
        ca = a1;
      }
      if (!CARRY) {
        if (owner) SMEM_ADD(&sJ[offq + ck], ca);
        ck = -1;
      }
      int j = n0 + n1;
      while (rest) {  // lanes in further cells (stale sort): one update per lane and value
        const int kj = __shfl_sync(FULL, key, __ffs(rest) - 1);
        rest &= rest - 1;
        if (owner) SMEM_ADD(&sJ[offq + kj], row[j]);
        j++;
      }
    }
    __syncwarp();
  }
  if (ck >= 0 && owner) SMEM_ADD(&sJ[offq + ck], ca);
  if (qcount) {
    __syncwarp();
    drain_extras(P, sJ, Q, qcount, lane);
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
      const double val = sJ[f * JC + ly * JP + lx];
      if (val != 0.0) atomicAdd(P.j[f] + o, val);
    }
  }
}

inline void launch_push(const PushParams &P, int nd, bool tiled, cudaStream_t s, long long *launches) {
  static bool attr_set = false;
  static int variant = 0;
  if (tiled && nd == 2) {
    if (!attr_set) {
      cudaFuncSetAttribute(push_tiled_2d<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PUSH2D_SMEM);
      cudaFuncSetAttribute(push_tiled_2d<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PUSH2D_SMEM);
      cudaFuncSetAttribute(push_tiled_2d<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PUSH2D_SMEM);
      cudaFuncSetAttribute(push_tiled_2d<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PUSH2D_SMEM);
      if (const char *e = getenv("EPB_PUSH_VARIANT")) variant = atoi(e);
      attr_set = true;
    }
    if (P.tg.ntiles > 0) {
      switch (variant) {
        case 1: push_tiled_2d<true, false><<<P.tg.ntiles, PUSH2D_THREADS, PUSH2D_SMEM, s>>>(P); break;
        case 2: push_tiled_2d<false, true><<<P.tg.ntiles, PUSH2D_THREADS, PUSH2D_SMEM, s>>>(P); break;
        case 3: push_tiled_2d<false, false><<<P.tg.ntiles, PUSH2D_THREADS, PUSH2D_SMEM, s>>>(P); break;
        default: push_tiled_2d<true, true><<<P.tg.ntiles, PUSH2D_THREADS, PUSH2D_SMEM, s>>>(P); break;
      }
      (*launches)++;
    }
    return;
  }
  const long long cnt = P.last - P.first;
  if (cnt <= 0) return;
  long long blocks = (cnt + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  if (nd == 1) push_generic<1><<<(int)blocks, 256, 0, s>>>(P);
  else if (nd == 2) push_generic<2><<<(int)blocks, 256, 0, s>>>(P);
  else push_generic<3><<<(int)blocks, 256, 0, s>>>(P);
  (*launches)++;
}

}  // namespace EPB_NS
