// epb_api.cu — C ABI, device state, field solver and boundary kernels.
// Compiled with -fmad=false: these kernels are bandwidth-bound, so separate
// rounding costs nothing and keeps the field arithmetic operation-for-operation
// that of fields.f90 / boundary.F90 / laser.f90.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include <algorithm>

#include "epb_internal.h"

int epb_fail(epb_handle *h, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  else fprintf(stderr, "epoch_b200: %s\n", buf);
  return code;
}

namespace {

constexpr int NG = EPB_NG;

struct FieldParams {
  double *f[9];
  int nd;
  int n[3];
  int sz[3];
  double cx, cy, cz, fac;  // cnx.. (E update) or hdtx.. (B update)
  // general solver (k_update_e_gen / k_update_b_gen): nt = field_order / 2 difference terms per
  // derivative with coefficients c1..c3 * cx (fields.f90:128-204); ext: extended 2D B stencil
  int nt, ext;
  double kx[3], ky[3], kz[3];
  double alpha[3], beta[6], gamma[3], delta[3];  // beta[2a + k]: other axes of a, lower first
  // CPML (fields.f90:112-204, :306-420): cx1 = c1 * (cnx / cpml_kappa_ex(ix)) ... per cell; kap[a] == nullptr: no CPML
  const double *kap[3];
  double ck[3];
};

__device__ __forceinline__ size_t fofs(const int *sz, int nd, int i, int j, int k) {
  size_t o = (size_t)(i + NG - 1);
  if (nd >= 2) o += (size_t)sz[0] * (size_t)(j + NG - 1);
  if (nd >= 3) o += (size_t)sz[0] * (size_t)sz[1] * (size_t)(k + NG - 1);
  return o;
}

// fields.f90:206-225 (2D), epoch3d fields.f90:312-337, epoch1d fields.f90:150-166
template <int ND>
__global__ void __launch_bounds__(256) k_update_e(const __grid_constant__ FieldParams F) {
  const int ex_ = F.n[0] + 1, ey_ = ND >= 2 ? F.n[1] + 1 : 1, ez_ = ND >= 3 ? F.n[2] + 1 : 1;
  const size_t total = (size_t)ex_ * ey_ * ez_;
  const size_t sx = 1, sy = F.sz[0], szz = (size_t)F.sz[0] * F.sz[1];
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(t % ex_);
    const int iy = ND >= 2 ? (int)((t / ex_) % ey_) : 1;
    const int iz = ND >= 3 ? (int)(t / ((size_t)ex_ * ey_)) : 1;
    const size_t o = fofs(F.sz, ND, ix, iy, iz);
    double *ex = F.f[0], *ey = F.f[1], *ez = F.f[2];
    const double *bx = F.f[3], *by = F.f[4], *bz = F.f[5];
    const double *jx = F.f[6], *jy = F.f[7], *jz = F.f[8];
    if (ND == 1) {
      ex[o] = ex[o] - F.fac * jx[o];
      ey[o] = ey[o] - F.cx * (bz[o] - bz[o - sx]) - F.fac * jy[o];
      ez[o] = ez[o] + F.cx * (by[o] - by[o - sx]) - F.fac * jz[o];
    } else if (ND == 2) {
      ex[o] = ex[o] + F.cy * (bz[o] - bz[o - sy]) - F.fac * jx[o];
      ey[o] = ey[o] - F.cx * (bz[o] - bz[o - sx]) - F.fac * jy[o];
      ez[o] = ez[o] + F.cx * (by[o] - by[o - sx]) - F.cy * (bx[o] - bx[o - sy]) - F.fac * jz[o];
    } else {
      ex[o] = ex[o] + F.cy * (bz[o] - bz[o - sy]) - F.cz * (by[o] - by[o - szz]) - F.fac * jx[o];
      ey[o] = ey[o] + F.cz * (bx[o] - bx[o - szz]) - F.cx * (bz[o] - bz[o - sx]) - F.fac * jy[o];
      ez[o] = ez[o] + F.cx * (by[o] - by[o - sx]) - F.cy * (bx[o] - bx[o - sy]) - F.fac * jz[o];
    }
  }
}

// fields.f90:422-439 (2D), epoch3d fields.f90:632-654, epoch1d fields.f90:296-303
template <int ND>
__global__ void __launch_bounds__(256) k_update_b(const __grid_constant__ FieldParams F) {
  const int ex_ = F.n[0] + 1, ey_ = ND >= 2 ? F.n[1] + 1 : 1, ez_ = ND >= 3 ? F.n[2] + 1 : 1;
  const size_t total = (size_t)ex_ * ey_ * ez_;
  const size_t sx = 1, sy = F.sz[0], szz = (size_t)F.sz[0] * F.sz[1];
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(t % ex_);
    const int iy = ND >= 2 ? (int)((t / ex_) % ey_) : 1;
    const int iz = ND >= 3 ? (int)(t / ((size_t)ex_ * ey_)) : 1;
    const size_t o = fofs(F.sz, ND, ix, iy, iz);
    const double *ex = F.f[0], *ey = F.f[1], *ez = F.f[2];
    double *bx = F.f[3], *by = F.f[4], *bz = F.f[5];
    if (ND == 1) {
      by[o] = by[o] + F.cx * (ez[o + sx] - ez[o]);
      bz[o] = bz[o] - F.cx * (ey[o + sx] - ey[o]);
    } else if (ND == 2) {
      bx[o] = bx[o] - F.cy * (ez[o + sy] - ez[o]);
      by[o] = by[o] + F.cx * (ez[o + sx] - ez[o]);
      bz[o] = bz[o] - F.cx * (ey[o + sx] - ey[o]) + F.cy * (ex[o + sy] - ex[o]);
    } else {
      bx[o] = bx[o] - F.cy * (ez[o + sy] - ez[o]) + F.cz * (ey[o + szz] - ey[o]);
      by[o] = by[o] - F.cz * (ex[o + szz] - ex[o]) + F.cx * (ez[o + sx] - ez[o]);
      bz[o] = bz[o] - F.cx * (ey[o + sx] - ey[o]) + F.cy * (ex[o + sy] - ex[o]);
    }
  }
}

// Box copy / add inside one array set: dst(dlo + o) (=|+=) src(slo + o), o in [0,ext).
struct BoxOp {
  double *f[3];
  int nf;
  int nd;
  int sz[3];
  int slo[3], dlo[3], ext[3];
  int add;
};
__global__ void __launch_bounds__(256) k_box(const __grid_constant__ BoxOp B) {
  const size_t total = (size_t)B.ext[0] * B.ext[1] * B.ext[2];
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int a = (int)(t % B.ext[0]);
    const int b = (int)((t / B.ext[0]) % B.ext[1]);
    const int c_ = (int)(t / ((size_t)B.ext[0] * B.ext[1]));
    const size_t so = fofs(B.sz, B.nd, B.slo[0] + a, B.slo[1] + b, B.slo[2] + c_);
    const size_t dq = fofs(B.sz, B.nd, B.dlo[0] + a, B.dlo[1] + b, B.dlo[2] + c_);
    for (int q = 0; q < B.nf; q++) {
      if (B.add) B.f[q][dq] = B.f[q][dq] + B.f[q][so];
      else B.f[q][dq] = B.f[q][so];
    }
  }
}

// field_clamp_zero / field_zero_gradient (boundary.F90:416-530) for three components.
struct MirrorOp {
  double *f[3];
  int stag[3];   // stagger(dir, field)
  int nd, sz[3], n[3];
  int d;         // axis
  int is_max;
  double sign[3]; // per component: -1 clamp, +1 zero gradient
};
__global__ void __launch_bounds__(256) k_mirror(const __grid_constant__ MirrorOp M) {
  // threads cover the full transverse extent (incl. ghosts), as the Fortran ':' slices do
  int e[3] = {M.sz[0], M.nd >= 2 ? M.sz[1] : 1, M.nd >= 3 ? M.sz[2] : 1};
  e[M.d] = 1;
  const size_t total = (size_t)e[0] * e[1] * e[2];
  const size_t str[3] = {1, (size_t)M.sz[0], (size_t)M.sz[0] * M.sz[1]};
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int c3[3] = {(int)(t % e[0]), (int)((t / e[0]) % e[1]), (int)(t / ((size_t)e[0] * e[1]))};
    size_t base = 0;
    for (int q = 0; q < M.nd; q++)
      if (q != M.d) base += str[q] * c3[q];
    const size_t s = str[M.d];
    const int nn = M.n[M.d];
    for (int q = 0; q < 3; q++) {
      double *a = M.f[q] + base;
      const double sgn = M.sign[q];
      // Fortran index i along axis d lives at offset (i + NG - 1) * s
#define AT(i) a[(size_t)((i) + NG - 1) * s]
      if (!M.is_max) {
        if (M.stag[q]) {
          for (int i = 1; i <= NG - 1; i++) AT(i - NG) = sgn * AT(NG - i);
          if (sgn < 0) AT(0) = 0.0;
        } else {
          for (int i = 1; i <= NG; i++) AT(i - NG) = sgn * AT(NG + 1 - i);
        }
      } else {
        if (M.stag[q]) {
          if (sgn < 0) AT(nn) = 0.0;
          for (int i = 1; i <= NG - 1; i++) AT(nn + i) = sgn * AT(nn - i);
        } else {
          for (int i = 1; i <= NG; i++) AT(nn + i) = sgn * AT(nn + 1 - i);
        }
      }
#undef AT
    }
  }
}

// particle_reflection_bcs (boundary.F90:534-630) for one J component on one boundary
struct FoldOp {
  double *a;
  int nd, sz[3], n[3];
  int d, is_max, flip;
};
__global__ void __launch_bounds__(256) k_jfold(const __grid_constant__ FoldOp M) {
  int e[3] = {M.sz[0], M.nd >= 2 ? M.sz[1] : 1, M.nd >= 3 ? M.sz[2] : 1};
  e[M.d] = 1;
  const size_t total = (size_t)e[0] * e[1] * e[2];
  const size_t str[3] = {1, (size_t)M.sz[0], (size_t)M.sz[0] * M.sz[1]};
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int c3[3] = {(int)(t % e[0]), (int)((t / e[0]) % e[1]), (int)(t / ((size_t)e[0] * e[1]))};
    size_t base = 0;
    for (int q = 0; q < M.nd; q++)
      if (q != M.d) base += str[q] * c3[q];
    const size_t s = str[M.d];
    const int nn = M.n[M.d];
    double *a = M.a + base;
#define AT(i) a[(size_t)((i) + NG - 1) * s]
    if (!M.is_max) {
      if (M.flip) for (int i = 1; i <= NG - 1; i++) { AT(i) = AT(i) - AT(-i); AT(-i) = 0.0; }
      else for (int i = 1; i <= NG - 1; i++) { AT(i) = AT(i) + AT(1 - i); AT(1 - i) = 0.0; }
    } else {
      if (M.flip) for (int i = 1; i <= NG; i++) { AT(nn - i) = AT(nn - i) - AT(nn + i); AT(nn + i) = 0.0; }
      else for (int i = 1; i <= NG; i++) { AT(nn + 1 - i) = AT(nn + 1 - i) + AT(nn + i); AT(nn + i) = 0.0; }
    }
#undef AT
  }
}

// setup_field_boundaries (setup.F90:391-447), x boundaries
struct SnapOp {
  const double *f[6];
  double *snap;  // [2][6][plane]
  int nd, sz[3], n[3];
  int i0[2];     // plane of the snapshot per side: 1 / nx, or one outside the laser plane of a cpml_laser face (setup.F90:409-412)
  size_t plane;
};
__global__ void __launch_bounds__(256) k_snapshot(const __grid_constant__ SnapOp S) {
  const int ey_ = S.nd >= 2 ? S.sz[1] : 1, ez_ = S.nd >= 3 ? S.sz[2] : 1;
  const size_t total = (size_t)ey_ * ez_;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % ey_) + 1 - NG, k = (int)(t / ey_) + 1 - NG;
    const int jj = S.nd >= 2 ? j : 1, kk = S.nd >= 3 ? k : 1;
    for (int side = 0; side < 2; side++) {
      const int i0 = S.i0[side];
      const size_t o = fofs(S.sz, S.nd, i0, jj, kk);
      for (int q = 0; q < 6; q++) {
        const bool avg = (q == EPB_EX || q == EPB_BY || q == EPB_BZ);
        const double val = avg ? 0.5 * (S.f[q][o] + S.f[q][o - 1]) : S.f[q][o];
        S.snap[((size_t)side * 6 + q) * S.plane + t] = val;
      }
    }
  }
}

// outflow_bcs_x_min / x_max (laser.f90:310-458; 3D laser.f90:350-506; 1D laser.f90:260-392)
struct OutflowOp {
  double *f[9];
  const double *snap;  // [6][plane] of this side
  const double *s1, *s2;  // (0:ny, 0:nz)
  int nd, sz[3], n[3];
  size_t plane;
  int is_max;
  int lp;        // laserpos: 1 / nx, or cpml_x_min_laser_idx / cpml_x_max_laser_idx (laser.f90:320-323, :396-399)
  double lx, ly, lz, sum, diff, dt_eps;
};
template <int ND>
__global__ void __launch_bounds__(256) k_outflow(const __grid_constant__ OutflowOp O) {
  const double c = EPB_C;
  const int ey_ = ND >= 2 ? O.n[1] + 1 : 1, ez_ = ND >= 3 ? O.n[2] + 1 : 1;
  const size_t total = (size_t)ey_ * ez_;
  const size_t sy = O.sz[0], szz = (size_t)O.sz[0] * O.sz[1];
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int j = ND >= 2 ? (int)(t % ey_) : 1, k = ND >= 3 ? (int)(t / ey_) : 1;
    // snapshot planes are indexed over the full ghosted transverse extent
    const size_t sp = (size_t)(ND >= 2 ? (j + NG - 1) : 0) + (size_t)(ND >= 2 ? O.sz[1] : 1) * (size_t)(ND >= 3 ? (k + NG - 1) : 0);
    const double *snap = O.snap;
#define SN(q) snap[(size_t)(q) * O.plane + sp]
    double *bx = O.f[3], *by = O.f[4], *bz = O.f[5];
    const double *ey = O.f[1], *ez = O.f[2], *jy = O.f[7], *jz = O.f[8];
    const double src1 = O.s1[t], src2 = O.s2[t];
    if (!O.is_max) {
      const size_t o = fofs(O.sz, ND, O.lp, j, k);
      bx[o - 1] = SN(EPB_BX);
      double tz = 4.0 * src1 + 2.0 * (SN(EPB_EY) + c * SN(EPB_BZ)) - 2.0 * ey[o];
      if (ND == 3) tz = tz - O.lz * (bx[o] - bx[o - szz]);
      tz = tz + O.dt_eps * jy[o] + O.diff * bz[o];
      double ty = -4.0 * src2 - 2.0 * (SN(EPB_EZ) - c * SN(EPB_BY)) + 2.0 * ez[o];
      if (ND >= 2) ty = ty - O.ly * (bx[o] - bx[o - sy]);
      ty = ty - O.dt_eps * jz[o] + O.diff * by[o];
      bz[o - 1] = O.sum * tz;
      by[o - 1] = O.sum * ty;
    } else {
      const size_t o = fofs(O.sz, ND, O.lp, j, k);
      bx[o + 1] = SN(EPB_BX);
      double tz = -4.0 * src1 - 2.0 * (SN(EPB_EY) - c * SN(EPB_BZ)) + 2.0 * ey[o];
      if (ND == 3) tz = tz + O.lz * (bx[o] - bx[o - szz]);
      tz = tz - O.dt_eps * jy[o] + O.diff * bz[o - 1];
      double ty = 4.0 * src2 + 2.0 * (SN(EPB_EZ) + c * SN(EPB_BY)) - 2.0 * ez[o];
      if (ND >= 2) ty = ty + O.ly * (bx[o] - bx[o - sy]);
      ty = ty + O.dt_eps * jz[o] + O.diff * by[o - 1];
      bz[o] = O.sum * tz;
      by[o] = O.sum * ty;
    }
#undef SN
  }
}

// ---- laser / outflow boundaries on the y and z faces ---------------------------------------------
// setup_field_boundaries (setup.F90:430-447; epoch3d :431-500) and outflow_bcs_{y,z}_{min,max}
// (laser.f90:462-610; epoch3d laser.f90:510-830): the x-face expressions under the cyclic permutation
// of axes and components (a = face normal, b = a+1, cc = a+2).  Plane index t runs over the two other
// axes in axis order, lower axis fastest, ghosted extents for the snapshots and (0:n) for the update.
struct FaceOp {
  double *f[9];
  double *snap;            // [2][6][plane] of this axis
  const double *s1, *s2;   // sources of this side
  int nd, sz[3], n[3], a, is_max;
  int lp, snap_i[2];       // laser plane of the update; snapshot planes per side (cpml_laser faces move them inwards)
  size_t plane;
  double l[3], sum, diff, dt_eps;
};
__device__ __forceinline__ size_t face_plane_index(const FaceOp &O, const int *p) {
  // position of (p) in the ghosted plane of axis a
  size_t t = 0, mul = 1;
  for (int d = 0; d < 3; d++) {
    if (d == O.a || d >= O.nd) continue;
    t += mul * (size_t)(p[d] + NG - 1);
    mul *= (size_t)O.sz[d];
  }
  return t;
}
__global__ void __launch_bounds__(256) k_snapshot_face(const __grid_constant__ FaceOp O) {
  int ext[3], lo[3];
  size_t total = 1;
  for (int d = 0; d < 3; d++) {
    if (d == O.a || d >= O.nd) { ext[d] = 1; lo[d] = 1; }
    else { ext[d] = O.sz[d]; lo[d] = 1 - NG; }
    total *= (size_t)ext[d];
  }
  size_t str[3] = {1, (size_t)O.sz[0], (size_t)O.sz[0] * O.sz[1]};
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int p[3];
    size_t r = t;
    for (int d = 0; d < 3; d++) { p[d] = lo[d] + (int)(r % ext[d]); r /= ext[d]; }
    const size_t tp = face_plane_index(O, p);
    for (int side = 0; side < 2; side++) {
      p[O.a] = O.snap_i[side];
      const size_t o = fofs(O.sz, O.nd, p[0], p[1], p[2]);
      for (int q = 0; q < 6; q++) {
        const bool avg = q < 3 ? (q == O.a) : (q - 3 != O.a);
        const double val = avg ? 0.5 * (O.f[q][o] + O.f[q][o - str[O.a]]) : O.f[q][o];
        O.snap[((size_t)side * 6 + q) * O.plane + tp] = val;
      }
    }
  }
}
template <int ND>
__global__ void __launch_bounds__(256) k_outflow_face(const __grid_constant__ FaceOp O) {
  const double c = EPB_C;
  const int a = O.a, b = (a + 1) % 3, cc = (a + 2) % 3;
  int ext[3];
  size_t total = 1;
  for (int d = 0; d < 3; d++) { ext[d] = (d == a || d >= ND) ? 1 : O.n[d] + 1; total *= (size_t)ext[d]; }
  const ptrdiff_t str[3] = {1, (ptrdiff_t)O.sz[0], (ptrdiff_t)O.sz[0] * O.sz[1]};
  double *Ba = O.f[3 + a], *Bb = O.f[3 + b], *Bc = O.f[3 + cc];
  const double *Eb = O.f[b], *Ec = O.f[cc], *Jb = O.f[6 + b], *Jc = O.f[6 + cc];
  const double *snap = O.snap + (size_t)O.is_max * 6 * O.plane;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int p[3];
    size_t r = t;
    for (int d = 0; d < 3; d++) { p[d] = (d == a || d >= ND) ? 1 : (int)(r % ext[d]); if (!(d == a || d >= ND)) r /= ext[d]; }
    const size_t tp = face_plane_index(O, p);
#define SN(q) snap[(size_t)(q) * O.plane + tp]
    const double src1 = O.s1[t], src2 = O.s2[t];
    p[a] = O.lp;  // laserpos
    const ptrdiff_t o = (ptrdiff_t)fofs(O.sz, ND, p[0], p[1], p[2]);
    const ptrdiff_t sa = str[a];
    if (!O.is_max) {
      Ba[o - sa] = SN(3 + a);
      double tc = 4.0 * src1 + 2.0 * (SN(b) + c * SN(3 + cc)) - 2.0 * Eb[o];
      if (cc < ND) tc = tc - O.l[cc] * (Ba[o] - Ba[o - str[cc]]);
      tc = tc + O.dt_eps * Jb[o] + O.diff * Bc[o];
      double tb = -4.0 * src2 - 2.0 * (SN(cc) - c * SN(3 + b)) + 2.0 * Ec[o];
      if (b < ND) tb = tb - O.l[b] * (Ba[o] - Ba[o - str[b]]);
      tb = tb - O.dt_eps * Jc[o] + O.diff * Bb[o];
      Bc[o - sa] = O.sum * tc;
      Bb[o - sa] = O.sum * tb;
    } else {
      Ba[o + sa] = SN(3 + a);
      double tc = -4.0 * src1 - 2.0 * (SN(b) - c * SN(3 + cc)) + 2.0 * Eb[o];
      if (cc < ND) tc = tc + O.l[cc] * (Ba[o] - Ba[o - str[cc]]);
      tc = tc - O.dt_eps * Jb[o] + O.diff * Bc[o - sa];
      double tb = 4.0 * src2 + 2.0 * (SN(cc) + c * SN(3 + b)) - 2.0 * Ec[o];
      if (b < ND) tb = tb + O.l[b] * (Ba[o] - Ba[o - str[b]]);
      tb = tb + O.dt_eps * Jc[o] + O.diff * Bb[o - sa];
      Bc[o] = O.sum * tc;
      Bb[o] = O.sum * tb;
    }
#undef SN
  }
}

// calc_ppc (io/calc_df.F90:761-808): cell = FLOOR((pos - x_grid_min_local)/dx + 0.5) + 1
struct CountOp {
  const double *x[3];
  PRange r;
  int nd, nloc[3];
  double gmin[3], dx[3];
  int *out;
};
__global__ void __launch_bounds__(256) k_cell_counts(const __grid_constant__ CountOp C) {
  const long long nn = prange_n(C.r);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (long long)gridDim.x * blockDim.x) {
    if (!prange_valid(C.r, i)) continue;
    const long long e = prange_at(C.r, i);
    int cell[3] = {1, 1, 1};
    bool ok = true;
    for (int d = 0; d < C.nd; d++) {
      cell[d] = __double2int_rd((C.x[d][e] - C.gmin[d]) / C.dx[d] + 0.5) + 1;
      if (cell[d] < 1 || cell[d] > C.nloc[d]) ok = false;
    }
    if (ok) atomicAdd(&C.out[(size_t)(cell[0] - 1) + (size_t)C.nloc[0] * ((size_t)(cell[1] - 1) + (size_t)C.nloc[1] * (cell[2] - 1))], 1);
  }
}

// Diagnostics moments: io/calc_df.F90 calc_number_density :689-757, calc_charge_density :608-685,
// calc_mass_density :35-110 (1D/3D trees alike).  particle_to_grid.inc (division by dx, not *idx) with the
// normalised triangle weights of include/triangle/gxfac.inc; data(cell+ix, ...) += gx*gy*gz*wdata.
struct MomentOp {
  const double *x[3], *w;
  PRange r;
  int nd, sz[3];
  double gmin[3], dx[3];
  double scale;    // part_q (charge density, current), part_m (mass density)
  int use_scale;   // 0: number density, wdata = weight; 1: scale * weight; 2: calc_per_species_current (:1132-1235),
                   // wdata = part_q * weight * p(dir) / sqrt((m c)^2 + p^2)
  int dir;
  double part_mc;
  const double *p[3];
  double *out;
};
__global__ void __launch_bounds__(256) k_moment(const __grid_constant__ MomentOp M) {
  const long long nn = prange_n(M.r);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (long long)gridDim.x * blockDim.x) {
    if (!prange_valid(M.r, i)) continue;
    const long long e = prange_at(M.r, i);
    int cell[3] = {1, 1, 1};
    double g[3][3] = {{0.0, 1.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 1.0, 0.0}};
    bool ok = true;
    for (int d = 0; d < M.nd; d++) {
      const double cell_r = (M.x[d][e] - M.gmin[d]) / M.dx[d];
      const int cx = __double2int_rd(cell_r + 0.5);
      const double cf = (double)cx - cell_r;
      cell[d] = cx + 1;
      const double c2 = cf * cf;
      g[d][0] = 0.5 * (0.25 + c2 + cf);
      g[d][1] = 0.75 - c2;
      g[d][2] = 0.5 * (0.25 + c2 - cf);
      if (cell[d] - 1 < 1 - NG || cell[d] + 1 > M.sz[d] - NG) ok = false;  // outside the allocated extent
    }
    if (!ok) continue;
    double wdata = M.use_scale ? M.scale * M.w[e] : M.w[e];
    if (M.use_scale == 2) {
      const double part_px = M.p[0][e], part_py = M.p[1][e], part_pz = M.p[2][e];
      const double root = 1.0 / sqrt(M.part_mc * M.part_mc + part_px * part_px + part_py * part_py + part_pz * part_pz);
      wdata = wdata * (M.dir == 0 ? part_px : M.dir == 1 ? part_py : part_pz) * root;
    }
    if (M.nd == 1) {
      for (int ix = -1; ix <= 1; ix++) atomicAdd(M.out + fofs(M.sz, 1, cell[0] + ix, 1, 1), g[0][ix + 1] * wdata);
    } else if (M.nd == 2) {
      for (int iy = -1; iy <= 1; iy++)
        for (int ix = -1; ix <= 1; ix++)
          atomicAdd(M.out + fofs(M.sz, 2, cell[0] + ix, cell[1] + iy, 1), g[0][ix + 1] * g[1][iy + 1] * wdata);
    } else {
      for (int iz = -1; iz <= 1; iz++)
        for (int iy = -1; iy <= 1; iy++)
          for (int ix = -1; ix <= 1; ix++)
            atomicAdd(M.out + fofs(M.sz, 3, cell[0] + ix, cell[1] + iy, cell[2] + iz),
                      g[0][ix + 1] * g[1][iy + 1] * g[2][iz + 1] * wdata);
    }
  }
}
// calc_ekbar (io/calc_df.F90:116-221) and the two passes of calc_temperature (:877-1128)
struct Moment2Op {
  const double *x[3], *p[3], *w;
  PRange r;
  int nd, sz[3];
  double gmin[3], dx[3];
  int mode;        // 3: ekbar (a0 += g wdata, a1 += g w); 4: temperature pass 1 (mean[q] += g w p/sqrt(m), cnt += g w);
                   // 5: pass 2 (sig += g sum_q (p/sqrt(m) - mean[q])^2, cnt += g)
  int dir;         // temperature: -1 all components, else one
  int sub;         // mode 3: 0 ekbar; 1..6 ekflux -x,+x,-y,+y,-z,+z (:415-557); 7..9 average momentum px,py,pz (:1239-1317)
  double flux_fac; // ekflux: xfac / yfac / zfac of the direction
  double part_mc, sqrt_part_m;
  double *a0, *a1;           // ekbar: data, wt; temperature: sigma, count
  double *mean[3];
};
// mode 3: what one particle adds to the data array (ekbar, one ekflux direction, one momentum component)
__device__ __forceinline__ double moment2_wdata(const Moment2Op &M, long long e, double part_w) {
  const double c = EPB_C;
  if (M.sub >= 7) return part_w * M.p[M.sub - 7][e];
  const double fac = M.part_mc * part_w * c;
  const double part_ux = M.p[0][e] / M.part_mc, part_uy = M.p[1][e] / M.part_mc, part_uz = M.p[2][e] / M.part_mc;
  const double part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz;
  const double gamma_rel = sqrt(part_u2 + 1.0);
  const double gamma_rel_m1 = part_u2 / (gamma_rel + 1.0);
  double wdata = gamma_rel_m1 * fac;
  if (M.sub >= 1) {
    const int a = (M.sub - 1) / 2;
    const double part_flux = M.flux_fac * (a == 0 ? part_ux : a == 1 ? part_uy : part_uz) / gamma_rel;
    if ((M.sub - 1) % 2 == 0) wdata = -wdata * fmin(part_flux, 0.0);
    else wdata = wdata * fmax(part_flux, 0.0);
  }
  return wdata;
}
__global__ void __launch_bounds__(256) k_moment2(const __grid_constant__ Moment2Op M) {
  const long long nn = prange_n(M.r);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (long long)gridDim.x * blockDim.x) {
    if (!prange_valid(M.r, i)) continue;
    const long long e = prange_at(M.r, i);
    int cell[3] = {1, 1, 1};
    double g[3][3] = {{0.0, 1.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 1.0, 0.0}};
    bool ok = true;
    for (int d = 0; d < M.nd; d++) {
      const double cell_r = (M.x[d][e] - M.gmin[d]) / M.dx[d];
      const int cx = __double2int_rd(cell_r + 0.5);
      const double cf = (double)cx - cell_r;
      cell[d] = cx + 1;
      const double c2 = cf * cf;
      g[d][0] = 0.5 * (0.25 + c2 + cf);
      g[d][1] = 0.75 - c2;
      g[d][2] = 0.5 * (0.25 + c2 - cf);
      if (cell[d] - 1 < 1 - NG || cell[d] + 1 > M.sz[d] - NG) ok = false;
    }
    if (!ok) continue;
    const double part_w = M.w[e];
    double wdata = 0.0, pm[3] = {0.0, 0.0, 0.0};
    if (M.mode == 6) {  // calc_average_weight (:811-873): nearest cell only
      const size_t o = fofs(M.sz, M.nd, cell[0], cell[1], cell[2]);
      atomicAdd(M.a0 + o, part_w);
      atomicAdd(M.a1 + o, 1.0);
      continue;
    }
    if (M.mode == 3) {
      wdata = moment2_wdata(M, e, part_w);
    } else {
      for (int q = 0; q < 3; q++) pm[q] = M.p[q][e] / M.sqrt_part_m;
    }
    const int z0 = M.nd >= 3 ? -1 : 0, z1 = M.nd >= 3 ? 1 : 0;
    const int y0 = M.nd >= 2 ? -1 : 0, y1 = M.nd >= 2 ? 1 : 0;
    for (int iz = z0; iz <= z1; iz++)
      for (int iy = y0; iy <= y1; iy++)
        for (int ix = -1; ix <= 1; ix++) {
          // gx(ix) * gy(iy) * gz(iz), left to right
          double gg = g[0][ix + 1];
          if (M.nd >= 2) gg = gg * g[1][iy + 1];
          if (M.nd >= 3) gg = gg * g[2][iz + 1];
          const size_t o = fofs(M.sz, M.nd, cell[0] + ix, cell[1] + iy, cell[2] + iz);
          if (M.mode == 3) {
            atomicAdd(M.a0 + o, gg * wdata);
            atomicAdd(M.a1 + o, gg * part_w);
          } else if (M.mode == 4) {
            const double gf = gg * part_w;
            for (int q = 0; q < 3; q++)
              if (M.dir < 0 || M.dir == q) atomicAdd(M.mean[q] + o, gf * pm[q]);
            atomicAdd(M.a1 + o, gf);
          } else {
            double wd;
            if (M.dir < 0) {
              const double d0 = pm[0] - M.mean[0][o], d1 = pm[1] - M.mean[1][o], d2 = pm[2] - M.mean[2][o];
              wd = d0 * d0 + d1 * d1 + d2 * d2;
            } else {
              const double d0 = pm[M.dir] - M.mean[M.dir][o];
              wd = d0 * d0;
            }
            atomicAdd(M.a0 + o, gg * wd);
            atomicAdd(M.a1 + o, gg);
          }
        }
  }
}
// The same sums on the 2D slot columns (layout 2): one warp per group of 32 columns = 16 x 2 cells, one lane per
// column.  A column holds the particles that GATHER in its cell at the next push, so nearly all of them also have
// it as their nearest cell: those accumulate their nine stencil values per output array in registers, the warp
// folds the lanes' registers into its 18 x 4-cell shared tile (nine conflict-free steps) and adds the tile to the
// grid once -- 72 atomics per array and group instead of nine per particle.  The few particles whose nearest cell
// is a neighbour of the column's cell take the per-particle atomics of k_moment2.  The kernel is bound by the
// latency of one particle's dependent FP64 chain, so the divisions by loop constants go through div_rcp (no
// slow-path branch: they overlap) and the next row is loaded while the current one is worked on.
constexpr int MS_WARPS = 4;
template <int MODE>   // 3: NA = 2 (data, wt); 4: NA = 4 (mean x 3, count); 5: NA = 2 (sigma, count)
__global__ void __launch_bounds__(MS_WARPS * 32, 4) k_moment2_slots(const __grid_constant__ Moment2Op M, const TileGeom tg) {
  constexpr int NA = MODE == 4 ? 4 : 2;
  constexpr int TW = 18, TH = 4;
  __shared__ double tile_all[MS_WARPS][NA][TW * TH];
  __shared__ double mean_all[MODE == 5 ? MS_WARPS : 1][3][TW * TH];   // pass 2: the means around the warp's cells
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int ngroups = tg.nkeys / 32, gpt = tg.cpt / 32;
  double (*tile)[TW * TH] = tile_all[wib];
  double (*mean_s)[TW * TH] = mean_all[MODE == 5 ? wib : 0];
  double *out[NA];
  if (MODE == 4) { out[0] = M.mean[0]; out[1] = M.mean[1]; out[NA - 2] = M.mean[2]; out[NA - 1] = M.a1; }
  else { out[0] = M.a0; out[1] = M.a1; }
  const double rdx[2] = {1.0 / M.dx[0], 1.0 / M.dx[1]};
  const double pdiv = MODE == 3 ? M.part_mc : M.sqrt_part_m, rpdiv = 1.0 / pdiv;
  const int lx = lane & 15, ly = lane >> 4;
  const int home_o = (ly + 1) * TW + lx + 1;   // the column's cell inside the warp's tile
  for (int g = blockIdx.x * MS_WARPS + wib; g < ngroups; g += gridDim.x * MS_WARPS) {
    const int t = g / gpt, gi = g - t * gpt;
    const int ty = t / tg.nt[0], tx = t - ty * tg.nt[0];
    const int x0 = tx * tg.T[0], y0 = ty * tg.T[1] + gi * 2;   // Fortran index of tile cell (0, 0), the halo corner; the lane's column is tile cell (lx + 1, ly + 1)
    const int n = M.r.cnt[g * 32 + lane];
    if (MODE == 5) {
      for (int i = lane; i < TW * TH; i += 32) {
        const int cx = x0 + i % TW, cy = y0 + i / TW;
        const bool in = cx <= M.sz[0] - NG && cy <= M.sz[1] - NG;
        const size_t o = in ? fofs(M.sz, 2, cx, cy, 1) : 0;
#pragma unroll
        for (int q = 0; q < 3; q++) mean_s[q][i] = (in && (M.dir < 0 || M.dir == q)) ? M.mean[q][o] : 0.0;
      }
      __syncwarp();
    }
    double acc[NA][9];
#pragma unroll
    for (int a = 0; a < NA; a++)
#pragma unroll
      for (int k = 0; k < 9; k++) acc[a][k] = 0.0;
    // row r of the lane's column: element e0 + r * rstride of every component array (planes, or row blocks of K x 32)
    const long long e0 = ((((long long)g * M.r.R) * M.r.K) << 5) + lane, rstride = (long long)M.r.K << 5;
    const double *comp[6] = {M.x[0], M.x[1], M.p[0], M.p[1], M.p[2], M.w};
    double nxt[6];
    if (n > 0) {
#pragma unroll
      for (int q = 0; q < 6; q++) nxt[q] = comp[q][e0];
    }
    for (int r = 0; r < n; r++) {
      double cur[6];
#pragma unroll
      for (int q = 0; q < 6; q++) cur[q] = nxt[q];
      if (r + 1 < n) {
#pragma unroll
        for (int q = 0; q < 6; q++) nxt[q] = comp[q][e0 + (r + 1) * rstride];
      }
      int cell[2];
      double gw[2][3];
      bool ok = true;
#pragma unroll
      for (int d = 0; d < 2; d++) {
        const double cell_r = div_rcp(cur[d] - M.gmin[d], M.dx[d], rdx[d]);
        const int cx = __double2int_rd(cell_r + 0.5);
        const double cf = (double)cx - cell_r;
        cell[d] = cx + 1;
        const double c2 = cf * cf;
        gw[d][0] = 0.5 * (0.25 + c2 + cf);
        gw[d][1] = 0.75 - c2;
        gw[d][2] = 0.5 * (0.25 + c2 - cf);
        if (cell[d] - 1 < 1 - NG || cell[d] + 1 > M.sz[d] - NG) ok = false;
      }
      if (!ok) continue;
      const double part_w = cur[5];
      double wdata = 0.0, pm[3];
#pragma unroll
      for (int q = 0; q < 3; q++) pm[q] = div_rcp(cur[2 + q], pdiv, rpdiv);   // mode 3: u = p / (m c); else p / sqrt(m)
      if (MODE == 3) {
        if (M.sub >= 7) {
          wdata = part_w * (M.sub == 7 ? cur[2] : M.sub == 8 ? cur[3] : cur[4]);
        } else {
          const double fac = M.part_mc * part_w * EPB_C;
          const double part_u2 = pm[0] * pm[0] + pm[1] * pm[1] + pm[2] * pm[2];
          const double gamma_rel = sqrt(part_u2 + 1.0);
          const double gamma_rel_m1 = part_u2 / (gamma_rel + 1.0);
          wdata = gamma_rel_m1 * fac;
          if (M.sub >= 1) {
            const int a = (M.sub - 1) / 2;
            const double part_flux = M.flux_fac * (a == 0 ? pm[0] : a == 1 ? pm[1] : pm[2]) / gamma_rel;
            if ((M.sub - 1) % 2 == 0) wdata = -wdata * fmin(part_flux, 0.0);
            else wdata = wdata * fmax(part_flux, 0.0);
          }
        }
      }
      const bool home = cell[0] == x0 + lx + 1 && cell[1] == y0 + ly + 1;
      // what the particle adds at stencil point k = (iy + 1) * 3 + ix + 1; mean_k: the means there (pass 2)
      auto value = [&](int k, const double *mean_k, double *v) {
        const double gg = gw[0][k % 3] * gw[1][k / 3];
        if (MODE == 3) {
          v[0] = gg * wdata;
          v[1] = gg * part_w;
        } else if (MODE == 4) {
          const double gf = gg * part_w;
          v[0] = gf * pm[0]; v[1] = gf * pm[1]; v[NA - 2] = gf * pm[2];
          v[NA - 1] = gf;
        } else {
          double wd = 0.0;
#pragma unroll
          for (int q = 0; q < 3; q++)
            if (M.dir < 0 || M.dir == q) { const double dq = pm[q] - mean_k[q]; wd += dq * dq; }
          v[0] = gg * wd;
          v[1] = gg;
        }
      };
      if (home) {
#pragma unroll
        for (int k = 0; k < 9; k++) {
          double mk[3] = {0.0, 0.0, 0.0}, v[NA];
          if (MODE == 5) {
            const int o = home_o + (k / 3 - 1) * TW + (k % 3 - 1);
#pragma unroll
            for (int q = 0; q < 3; q++) mk[q] = mean_s[q][o];
          }
          value(k, mk, v);
#pragma unroll
          for (int a = 0; a < NA; a++) acc[a][k] += v[a];
        }
      } else {
        for (int k = 0; k < 9; k++) {
          const size_t o = fofs(M.sz, 2, cell[0] + (k % 3 - 1), cell[1] + (k / 3 - 1), 1);
          double mk[3] = {0.0, 0.0, 0.0}, v[NA];
          if (MODE == 5) {
#pragma unroll
            for (int q = 0; q < 3; q++)
              if (M.dir < 0 || M.dir == q) mk[q] = M.mean[q][o];
          }
          value(k, mk, v);
#pragma unroll
          for (int a = 0; a < NA; a++)
            if (MODE != 4 || a == NA - 1 || M.dir < 0 || M.dir == a) atomicAdd(out[a] + o, v[a]);
        }
      }
    }
    // fold the lanes' registers into the warp's tile: in step k every lane adds to a different cell
    for (int i = lane; i < NA * TW * TH; i += 32) (&tile[0][0])[i] = 0.0;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const int o = home_o + (k / 3 - 1) * TW + (k % 3 - 1);
#pragma unroll
      for (int a = 0; a < NA; a++) tile[a][o] += acc[a][k];
      __syncwarp();
    }
    for (int i = lane; i < TW * TH; i += 32) {
      const int cx = x0 + i % TW, cy = y0 + i / TW;
      if (cx > M.sz[0] - NG || cy > M.sz[1] - NG) continue;   // padding columns past the rank's extent hold no particle
      const size_t o = fofs(M.sz, 2, cx, cy, 1);
#pragma unroll
      for (int a = 0; a < NA; a++) {
        if (MODE == 4 && a < NA - 1 && M.dir >= 0 && M.dir != a) continue;   // one-component temperature: only that mean
        const double v = tile[a][i];
        if (v != 0.0) atomicAdd(out[a] + o, v);
      }
    }
    __syncwarp();
  }
}
// calc_poynt_flux (io/calc_df.F90:561-604, epoch1d :441-474, epoch3d :585-650): E x B / mu0 at the cell centres
struct PoyntOp {
  const double *f[6];
  int nd, n[3], sz[3], dir;
  double *out;
};
__device__ __forceinline__ double cell_centred(const PoyntOp &P, int field, int i, int j, int k) {
  // the active axes the component is staggered along (setup.F90:124-134); lower axis first
  int ax[2], na = 0;
  for (int d = 0; d < P.nd; d++) {
    const bool st = field < 3 ? (d == field) : (d != field - 3);
    if (st) ax[na++] = d;
  }
  const double *a = P.f[field];
  int q[3] = {i, j, k};
  if (na == 0) return a[fofs(P.sz, P.nd, q[0], q[1], q[2])];
  if (na == 1) {
    const double hi = a[fofs(P.sz, P.nd, q[0], q[1], q[2])];
    q[ax[0]] -= 1;
    const double lo = a[fofs(P.sz, P.nd, q[0], q[1], q[2])];
    return 0.5 * (lo + hi);
  }
  double v[4];
  for (int s = 0; s < 4; s++) {  // (lo,lo), (hi,lo), (lo,hi), (hi,hi)
    int r[3] = {i, j, k};
    if (!(s & 1)) r[ax[0]] -= 1;
    if (!(s & 2)) r[ax[1]] -= 1;
    v[s] = a[fofs(P.sz, P.nd, r[0], r[1], r[2])];
  }
  return 0.25 * (v[0] + v[1] + v[2] + v[3]);
}
__global__ void __launch_bounds__(256) k_poynt_flux(const __grid_constant__ PoyntOp P) {
  const double mu0 = 4.e-7 * 3.141592653589793238462643383279503;
  const size_t total = (size_t)P.n[0] * P.n[1] * P.n[2];
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(t % P.n[0]) + 1;
    const int iy = (int)((t / P.n[0]) % P.n[1]) + 1;
    const int iz = (int)(t / ((size_t)P.n[0] * P.n[1])) + 1;
    const int e1 = (P.dir + 1) % 3, e2 = (P.dir + 2) % 3;
    const double e1c = cell_centred(P, e1, ix, iy, iz), e2c = cell_centred(P, e2, ix, iy, iz);
    const double b1c = cell_centred(P, 3 + e1, ix, iy, iz), b2c = cell_centred(P, 3 + e2, ix, iy, iz);
    P.out[fofs(P.sz, P.nd, ix, iy, iz)] = (e1c * b2c - e2c * b1c) / mu0;
  }
}
// element-wise tails of calc_ekbar / calc_temperature
struct MomentPostOp { int op; size_t n; double *a, *b, *m[3]; double k1, k2; };
__global__ void __launch_bounds__(256) k_moment_post(const __grid_constant__ MomentPostOp P) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < P.n; t += (size_t)gridDim.x * blockDim.x) {
    if (P.op == 0) {         // data_array = data_array / MAX(wt, c_tiny)
      P.a[t] = P.a[t] / fmax(P.b[t], P.k1);
    } else if (P.op == 1) {  // part_count = MAX(part_count, 1.e-6); mean = mean / part_count
      const double pc = fmax(P.b[t], 1.e-6);
      P.b[t] = pc;
      for (int q = 0; q < 3; q++) P.m[q][t] = P.m[q][t] / pc;
    } else {                 // sigma = sigma / MAX(part_count, 1.e-6) / kb / dof
      P.a[t] = P.a[t] / fmax(P.b[t], 1.e-6) / P.k1 / P.k2;
    }
  }
}
__global__ void __launch_bounds__(256) k_scale(double *a, size_t n, double s) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) a[t] = a[t] * s;
}

// Device-side uniform thermal loader (stands in for auto_load on benchmark-size runs).
struct LoadOp {
  double *x[3], *p[3], *w;
  int nd, nloc[3];
  int ppc;
  double gmin_local[3], dx[3];
  double weight;
  double stdev[3], drift[3];
  unsigned long long seed;
  int mixed;  // 1: cell drawn at random per particle (Poisson counts, the long-time state of a thermal plasma)
  long long i0, i1;  // particles [i0, i1) of the species are generated, written at index i - i0
};
__device__ __forceinline__ unsigned long long splitmix(unsigned long long &s) {
  unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(unsigned long long &s) {
  return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0);
}
__global__ void __launch_bounds__(256) k_load_uniform(const __grid_constant__ LoadOp L) {
  const long long ncell = (long long)L.nloc[0] * L.nloc[1] * L.nloc[2];
  const long long total = ncell * L.ppc;
  const long long last = L.i1 < total ? L.i1 : total;
  for (long long i = L.i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < last; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i - L.i0;
    unsigned long long s = L.seed ^ (0xD1B54A32D192ED03ull * (unsigned long long)(i + 1));
    long long cellid = i / L.ppc;
    if (L.mixed) cellid = (long long)(u01(s) * (double)ncell) % ncell;
    int cell[3] = {(int)(cellid % L.nloc[0]), (int)((cellid / L.nloc[0]) % L.nloc[1]),
                   (int)(cellid / ((long long)L.nloc[0] * L.nloc[1]))};
    for (int d = 0; d < L.nd; d++)
      L.x[d][o] = (L.gmin_local[d] + (double)cell[d] * L.dx[d]) + (u01(s) - 0.5) * L.dx[d];
    double g[4];
    for (int q = 0; q < 2; q++) {
      double r1, r2, ww;
      do {
        r1 = 2.0 * u01(s) - 1.0;
        r2 = 2.0 * u01(s) - 1.0;
        ww = r1 * r1 + r2 * r2;
      } while (!(ww > 0.0 && ww < 1.0));
      ww = sqrt((-2.0 * log(ww)) / ww);
      g[2 * q] = r1 * ww;
      g[2 * q + 1] = r2 * ww;
    }
    for (int d = 0; d < 3; d++) L.p[d][o] = g[d] * L.stdev[d] + L.drift[d];
    L.w[o] = L.weight;
  }
}

// Thermal particle boundaries (boundary.F90:1104-1148 and its x_max / y / z copies; epoch3d :1496-1550, epoch1d
// :728-750), applied right after the push to the particles particle_bc left beyond x_min_outer / x_max_outer of a
// thermal wall: the wall temperature is interpolated with the triangle weights at the particle's transverse
// position, the momentum normal to the wall is drawn from the inward flux distribution (flux_momentum_from_
// temperature with zero drift: SQRT(g1^2 + g2^2) of two normal deviates, particle_temperature.F90:409-460), the other
// two from Maxwellians (momentum_from_temperature :388-398), and the position is mirrored about the outer edge.
// The reference draws from the rank's serial KISS stream; here every particle has its own counter-based stream.
struct ThermalOp {
  double *x[3], *p[3];
  PRange r;
  int nd, n[3];
  double gmin_local[3], dx[3], min_outer[3], max_outer[3];
  int th_min[3], th_max[3];      // thermal wall on this rank's boundary face
  const double *ext_temp[6];     // (plane, 3): the transverse axes in axis order with ghost cells, lower axis fastest
  double mass;
  unsigned long long seed;
};
__device__ __forceinline__ void normal_pair(unsigned long long &s, double &a, double &b) {
  double u1 = u01(s), u2 = u01(s);
  if (u1 < 1e-300) u1 = 1e-300;
  const double r = sqrt(-2.0 * log(u1));
  a = r * cos(6.283185307179586476925286766559 * u2);
  b = r * sin(6.283185307179586476925286766559 * u2);
}
template <int ND>
__global__ void __launch_bounds__(256) k_thermal(const __grid_constant__ ThermalOp T) {
  const long long nn = prange_n(T.r);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (long long)gridDim.x * blockDim.x) {
    if (!prange_valid(T.r, i)) continue;
    const long long e = prange_at(T.r, i);
    double pos[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < ND; d++) pos[d] = T.x[d][e];
    bool any = false;
#pragma unroll
    for (int d = 0; d < ND; d++)
      any = any || (T.th_min[d] && pos[d] < T.min_outer[d]) || (T.th_max[d] && pos[d] >= T.max_outer[d]);
    if (!any) continue;
    unsigned long long s = T.seed ^ (0xD1B54A32D192ED03ull * (unsigned long long)(i + 1));
    (void)splitmix(s);
    double mom[3] = {T.p[0][e], T.p[1][e], T.p[2][e]};
#pragma unroll
    for (int d = 0; d < ND; d++) {
      const double part_pos = pos[d];
      for (int side = 0; side < 2; side++) {
        const bool hit = side == 0 ? (T.th_min[d] && part_pos < T.min_outer[d]) : (T.th_max[d] && part_pos >= T.max_outer[d]);
        if (!hit) continue;
        // wall temperature at the particle's transverse position (always the triangle weighting)
        int tr[2] = {0, 0}, ntr = 0;
        for (int q = 0; q < ND; q++) if (q != d) tr[ntr++] = q;
        size_t plane = 1;
        for (int q = 0; q < ntr; q++) plane *= (size_t)(T.n[tr[q]] + 2 * NG);
        int cell[2] = {0, 0};
        double g[2][3] = {{0.0, 1.0, 0.0}, {0.0, 1.0, 0.0}};
        for (int q = 0; q < ntr; q++) {
          const double cell_r = (pos[tr[q]] - T.gmin_local[tr[q]]) / T.dx[tr[q]];
          const int c = __double2int_rd(cell_r + 0.5);
          const double cf = (double)c - cell_r;
          cell[q] = c + 1;
          const double cf2 = cf * cf;
          g[q][0] = 0.5 * (0.25 + cf2 + cf);
          g[q][1] = 0.75 - cf2;
          g[q][2] = 0.5 * (0.25 + cf2 - cf);
        }
        const double *ET = T.ext_temp[2 * d + side];
        double temp[3];
        for (int c3 = 0; c3 < 3; c3++) {
          double t = 0.0;
          if (ntr == 0) {
            t = ET[c3];
          } else if (ntr == 1) {
            for (int a = -1; a <= 1; a++) t = t + g[0][a + 1] * ET[(size_t)(cell[0] + a + NG - 1) + plane * c3];
          } else {
            const size_t e0 = (size_t)(T.n[tr[0]] + 2 * NG);
            for (int b = -1; b <= 1; b++)
              for (int a = -1; a <= 1; a++)
                t = t + g[0][a + 1] * g[1][b + 1] * ET[(size_t)(cell[0] + a + NG - 1) + e0 * (size_t)(cell[1] + b + NG - 1) + plane * c3];
          }
          temp[c3] = t;
        }
        const double direction = side == 0 ? 1.0 : -1.0;   // -REAL(sgn): into the domain
        double g1, g2, g3, g4;
        normal_pair(s, g1, g2);
        normal_pair(s, g3, g4);
        const double gq[2] = {g3, g4};
        int k = 0;
        for (int c3 = 0; c3 < 3; c3++) {
          const double stdev = sqrt(temp[c3] * EPB_KB * T.mass);
          if (c3 == d) mom[c3] = direction * sqrt((g1 * stdev) * (g1 * stdev) + (g2 * stdev) * (g2 * stdev));
          else mom[c3] = gq[k++] * stdev;
        }
        pos[d] = 2.0 * (side == 0 ? T.min_outer[d] : T.max_outer[d]) - part_pos;
      }
    }
#pragma unroll
    for (int d = 0; d < ND; d++) T.x[d][e] = pos[d];
    T.p[0][e] = mom[0]; T.p[1][e] = mom[1]; T.p[2][e] = mom[2];
  }
}

// calc_total_energy_sum (io/calc_df.F90:1321-1417)
struct EnergyOp {
  const double *f[6];
  int nd, n[3], sz[3];
  double *out;  // [2]
};
__global__ void __launch_bounds__(256) k_field_energy(const __grid_constant__ EnergyOp E) {
  const size_t total = (size_t)E.n[0] * E.n[1] * E.n[2];
  double se = 0.0, sb = 0.0;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(t % E.n[0]) + 1;
    const int iy = (int)((t / E.n[0]) % E.n[1]) + 1;
    const int iz = (int)(t / ((size_t)E.n[0] * E.n[1])) + 1;
    const size_t o = fofs(E.sz, E.nd, ix, iy, iz);
    se += E.f[0][o] * E.f[0][o] + E.f[1][o] * E.f[1][o] + E.f[2][o] * E.f[2][o];
    sb += E.f[3][o] * E.f[3][o] + E.f[4][o] * E.f[4][o] + E.f[5][o] * E.f[5][o];
  }
  for (int s = 16; s >= 1; s >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, s);
    sb += __shfl_xor_sync(0xffffffffu, sb, s);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&E.out[0], se); atomicAdd(&E.out[1], sb); }
}
__global__ void __launch_bounds__(256) k_kinetic_energy(const double *px, const double *py, const double *pz,
                                                        const double *w, const __grid_constant__ PRange R, double mc, double mc2, double *out) {
  double s = 0.0;
  const long long n = prange_n(R);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (!prange_valid(R, i)) continue;
    const long long e = prange_at(R, i);
    const double ux = px[e] / mc, uy = py[e] / mc, uz = pz[e] / mc;
    const double u2 = ux * ux + uy * uy + uz * uz;
    const double gamma = sqrt(u2 + 1.0);
    s += w[e] * (u2 / (gamma + 1.0)) * mc2;  // (gamma-1) m c^2 without cancellation
  }
  for (int q = 16; q >= 1; q >>= 1) s += __shfl_xor_sync(0xffffffffu, s, q);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

inline int nblocks(size_t total, int cap = 148 * 16) {
  size_t b = (total + 255) / 256;
  if (b < 1) b = 1;
  if (b > (size_t)cap) b = cap;
  return (int)b;
}

void fill_field_params(epb_handle *h, FieldParams &F) {
  for (int q = 0; q < 9; q++) F.f[q] = h->f(q);
  F.nd = h->cfg.ndims;
  for (int d = 0; d < 3; d++) { F.n[d] = h->cfg.n[d]; F.sz[d] = h->sz[d]; }
}

// ---- field orders 4 / 6 and the extended 2D stencils ---------------------------------------------
// Every c*(difference) term of the order-2 expressions becomes the group c1*c*(d1) + c2*c*(d2)
// [+ c3*c*(d3)], added one after the other in the reference's textual order (fields.f90:128-204,
// :468-529; epoch3d fields.f90:337-430, :734-830; epoch1d fields.f90:166-215, :315-360).
template <int ND>
__global__ void __launch_bounds__(256) k_update_e_gen(const __grid_constant__ FieldParams F) {
  const int ex_ = F.n[0] + 1, ey_ = ND >= 2 ? F.n[1] + 1 : 1, ez_ = ND >= 3 ? F.n[2] + 1 : 1;
  const size_t total = (size_t)ex_ * ey_ * ez_;
  const ptrdiff_t sx = 1, sy = F.sz[0], szz = (ptrdiff_t)F.sz[0] * F.sz[1];
  const int nt = F.nt;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(t % ex_);
    const int iy = ND >= 2 ? (int)((t / ex_) % ey_) : 1;
    const int iz = ND >= 3 ? (int)(t / ((size_t)ex_ * ey_)) : 1;
    const ptrdiff_t o = (ptrdiff_t)fofs(F.sz, ND, ix, iy, iz);
    double *ex = F.f[0], *ey = F.f[1], *ez = F.f[2];
    const double *bx = F.f[3], *by = F.f[4], *bz = F.f[5];
    const double *jx = F.f[6], *jy = F.f[7], *jz = F.f[8];
    // backward difference k along stride s: f(i+k) - f(i-k-1)
    auto db = [&](const double *f, ptrdiff_t s, int k) { return f[o + k * s] - f[o - (k + 1) * s]; };
    double kx[3] = {F.kx[0], F.kx[1], F.kx[2]}, ky[3] = {F.ky[0], F.ky[1], F.ky[2]}, kz[3] = {F.kz[0], F.kz[1], F.kz[2]};
    if (F.kap[0]) {   // CPML: the stretching divides every cell's coefficients
      const double bx_ = F.cx / F.kap[0][ix + NG - 1];
      const double by_ = ND >= 2 ? F.cy / F.kap[1][iy + NG - 1] : 0.0;
      const double bz_ = ND >= 3 ? F.cz / F.kap[2][iz + NG - 1] : 0.0;
      for (int k = 0; k < 3; k++) { kx[k] = F.ck[k] * bx_; ky[k] = F.ck[k] * by_; kz[k] = F.ck[k] * bz_; }
    }
    double v = ex[o];
    if (ND >= 2) for (int k = 0; k < nt; k++) v = v + ky[k] * db(bz, sy, k);
    if (ND >= 3) for (int k = 0; k < nt; k++) v = v - kz[k] * db(by, szz, k);
    ex[o] = v - F.fac * jx[o];
    v = ey[o];
    if (ND >= 3) for (int k = 0; k < nt; k++) v = v + kz[k] * db(bx, szz, k);
    for (int k = 0; k < nt; k++) v = v - kx[k] * db(bz, sx, k);
    ey[o] = v - F.fac * jy[o];
    v = ez[o];
    for (int k = 0; k < nt; k++) v = v + kx[k] * db(by, sx, k);
    if (ND >= 2) for (int k = 0; k < nt; k++) v = v - ky[k] * db(bx, sy, k);
    ez[o] = v - F.fac * jz[o];
  }
}

template <int ND>
__global__ void __launch_bounds__(256) k_update_b_gen(const __grid_constant__ FieldParams F) {
  const int ex_ = F.n[0] + 1, ey_ = ND >= 2 ? F.n[1] + 1 : 1, ez_ = ND >= 3 ? F.n[2] + 1 : 1;
  const size_t total = (size_t)ex_ * ey_ * ez_;
  const ptrdiff_t sx = 1, sy = F.sz[0], szz = (ptrdiff_t)F.sz[0] * F.sz[1];
  const int nt = F.nt;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(t % ex_);
    const int iy = ND >= 2 ? (int)((t / ex_) % ey_) : 1;
    const int iz = ND >= 3 ? (int)(t / ((size_t)ex_ * ey_)) : 1;
    const ptrdiff_t o = (ptrdiff_t)fofs(F.sz, ND, ix, iy, iz);
    const double *ex = F.f[0], *ey = F.f[1], *ez = F.f[2];
    double *bx = F.f[3], *by = F.f[4], *bz = F.f[5];
    // forward difference k along stride s: f(i+k+1) - f(i-k)
    auto df = [&](const double *f, ptrdiff_t s, int k) { return f[o + (k + 1) * s] - f[o - k * s]; };
    double kx[3] = {F.kx[0], F.kx[1], F.kx[2]}, ky[3] = {F.ky[0], F.ky[1], F.ky[2]}, kz[3] = {F.kz[0], F.kz[1], F.kz[2]};
    double hx_ = F.cx, hy_ = F.cy, hz_ = F.cz;
    if (F.kap[0]) {   // CPML: cx1 = hdtx / cpml_kappa_bx(ix) (times c1.. for the higher orders), fields.f90:306-420
      hx_ = F.cx / F.kap[0][ix + NG - 1];
      hy_ = ND >= 2 ? F.cy / F.kap[1][iy + NG - 1] : 0.0;
      hz_ = ND >= 3 ? F.cz / F.kap[2][iz + NG - 1] : 0.0;
      for (int q = 0; q < 3; q++) { kx[q] = F.ck[q] * hx_; ky[q] = F.ck[q] * hy_; kz[q] = F.ck[q] * hz_; }
    }
    if (F.ext) {
      // fields.f90:441-465, epoch3d fields.f90:655-730, epoch1d fields.f90:304-312: derivative along
      // axis a = alpha + the two betas (lower other axis first; +1 then -1) + gamma (3D; first other
      // axis + then -, second other axis - then +) + delta, added in that order
      auto dd = [&](const double *f, int a) {
        const ptrdiff_t sa = a == 0 ? sx : a == 1 ? sy : szz;
        double v = F.alpha[a] * (f[o + sa] - f[o]);
        const int b0 = a == 0 ? 1 : 0, b1 = a == 2 ? 1 : 2;
        const ptrdiff_t s0 = b0 == 0 ? sx : sy, s1 = b1 == 1 ? sy : szz;
        if (b0 < ND) v = v + F.beta[2 * a] * (f[o + sa + s0] - f[o + s0] + f[o + sa - s0] - f[o - s0]);
        if (b1 < ND) v = v + F.beta[2 * a + 1] * (f[o + sa + s1] - f[o + s1] + f[o + sa - s1] - f[o - s1]);
        if (ND == 3)
          v = v + F.gamma[a] * (f[o + sa + s0 - s1] - f[o + s0 - s1] + f[o + sa - s0 - s1] - f[o - s0 - s1] +
                                f[o + sa + s0 + s1] - f[o + s0 + s1] + f[o + sa - s0 + s1] - f[o - s0 + s1]);
        v = v + F.delta[a] * (f[o + 2 * sa] - f[o - sa]);
        return v;
      };
      if (ND == 1) {
        by[o] = by[o] + hx_ * dd(ez, 0);
        bz[o] = bz[o] - hx_ * dd(ey, 0);
      } else if (ND == 2) {
        bx[o] = bx[o] - hy_ * dd(ez, 1);
        by[o] = by[o] + hx_ * dd(ez, 0);
        bz[o] = bz[o] - hx_ * dd(ey, 0) + hy_ * dd(ex, 1);
      } else {
        bx[o] = bx[o] - hy_ * dd(ez, 1) + hz_ * dd(ey, 2);
        by[o] = by[o] - hz_ * dd(ex, 2) + hx_ * dd(ez, 0);
        bz[o] = bz[o] - hx_ * dd(ey, 0) + hy_ * dd(ex, 1);
      }
      continue;
    }
    double v;
    if (ND >= 2) {
      v = bx[o];
      for (int k = 0; k < nt; k++) v = v - ky[k] * df(ez, sy, k);
      if (ND >= 3) for (int k = 0; k < nt; k++) v = v + kz[k] * df(ey, szz, k);
      bx[o] = v;
    }
    v = by[o];
    if (ND >= 3) for (int k = 0; k < nt; k++) v = v - kz[k] * df(ex, szz, k);
    for (int k = 0; k < nt; k++) v = v + kx[k] * df(ez, sx, k);
    by[o] = v;
    v = bz[o];
    for (int k = 0; k < nt; k++) v = v - kx[k] * df(ey, sx, k);
    if (ND >= 2) for (int k = 0; k < nt; k++) v = v + ky[k] * df(ex, sy, k);
    bz[o] = v;
  }
}

// smooth_array (current_smooth.F90:111-137; 1D :104-126; 3D :114-144): one strided binomial pass
// of the three current components, work arrays -> J on the interior cells
struct SmoothOp {
  const double *wk[3];
  double *a[3];
  int nd, n[3], sz[3], cs;
  double alpha, beta;
};
__global__ void __launch_bounds__(256) k_smooth(const __grid_constant__ SmoothOp S) {
  const size_t total = (size_t)S.n[0] * S.n[1] * S.n[2];
  const ptrdiff_t sx = S.cs, sy = (ptrdiff_t)S.sz[0] * S.cs, szz = (ptrdiff_t)S.sz[0] * S.sz[1] * S.cs;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(t % S.n[0]) + 1;
    const int iy = (int)((t / S.n[0]) % S.n[1]) + 1;
    const int iz = (int)(t / ((size_t)S.n[0] * S.n[1])) + 1;
    const ptrdiff_t o = (ptrdiff_t)fofs(S.sz, S.nd, ix, iy, iz);
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const double *w = S.wk[q];
      double nb = w[o - sx] + w[o + sx];
      if (S.nd >= 2) nb = nb + w[o - sy] + w[o + sy];
      if (S.nd >= 3) nb = nb + w[o - szz] + w[o + szz];
      S.a[q][o] = S.alpha * w[o] + nb * S.beta;
    }
  }
}
__global__ void __launch_bounds__(256) k_smooth_copyback(const __grid_constant__ SmoothOp S) {
  const size_t total = (size_t)S.n[0] * S.n[1] * S.n[2];
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(t % S.n[0]) + 1;
    const int iy = (int)((t / S.n[0]) % S.n[1]) + 1;
    const int iz = (int)(t / ((size_t)S.n[0] * S.n[1])) + 1;
    const size_t o = fofs(S.sz, S.nd, ix, iy, iz);
#pragma unroll
    for (int q = 0; q < 3; q++) const_cast<double *>(S.wk[q])[o] = S.a[q][o];
  }
}

// cpml_advance_e_currents / cpml_advance_b_currents (boundary.F90:1813-2023; epoch3d :2365-2790, epoch1d :929-1040)
// for one layer of axis a, (b, c) its cyclic successors:
//   E: psi_Eb = bco psi_Eb + cco (B_c(i) - B_c(i-1)), E_b -= fac psi_Eb;  psi_Ec from B_b, E_c += fac psi_Ec
//   B: psi_Bb = bco psi_Bb + cco (E_c(i+1) - E_c(i)), B_b += tstep psi_Bb; psi_Bc from E_b, B_c -= tstep psi_Bc
// over the interior of the other axes.  bco / cco are the reference's bcoeff / ccoeff_d per layer position, evaluated
// on the host (the same libm exp as the CPU oracle) for the half step.
struct CpmlOp {
  double *fb, *fc;             // updated components b, c
  const double *gb, *gc;       // the other field's components b, c
  double *psb, *psc;
  const double *bco, *cco;     // indexed i + NG - 1 along the axis
  int nd, sz[3], n[3], a, i0, i1, efield;
  double fac;                  // tstep c^2 (E) or tstep (B)
};
__global__ void __launch_bounds__(256) k_cpml(const __grid_constant__ CpmlOp O) {
  int ext[3], lo[3];
  size_t total = 1;
  for (int d = 0; d < 3; d++) {
    if (d == O.a) { ext[d] = O.i1 - O.i0 + 1; lo[d] = O.i0; }
    else if (d < O.nd) { ext[d] = O.n[d]; lo[d] = 1; }
    else { ext[d] = 1; lo[d] = 1; }
    total *= (size_t)ext[d];
  }
  const ptrdiff_t str[3] = {1, (ptrdiff_t)O.sz[0], (ptrdiff_t)O.sz[0] * O.sz[1]};
  const ptrdiff_t sa = str[O.a];
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int p[3];
    size_t r = t;
    for (int d = 0; d < 3; d++) { p[d] = lo[d] + (int)(r % ext[d]); r /= ext[d]; }
    const ptrdiff_t o = (ptrdiff_t)fofs(O.sz, O.nd, p[0], p[1], p[2]);
    const double bcoeff = O.bco[p[O.a] + NG - 1], ccoeff_d = O.cco[p[O.a] + NG - 1];
    if (O.efield) {
      O.psb[o] = bcoeff * O.psb[o] + ccoeff_d * (O.gc[o] - O.gc[o - sa]);
      O.fb[o] = O.fb[o] - O.fac * O.psb[o];
      O.psc[o] = bcoeff * O.psc[o] + ccoeff_d * (O.gb[o] - O.gb[o - sa]);
      O.fc[o] = O.fc[o] + O.fac * O.psc[o];
    } else {
      O.psb[o] = bcoeff * O.psb[o] + ccoeff_d * (O.gc[o + sa] - O.gc[o]);
      O.fb[o] = O.fb[o] + O.fac * O.psb[o];
      O.psc[o] = bcoeff * O.psc[o] + ccoeff_d * (O.gb[o + sa] - O.gb[o]);
      O.fc[o] = O.fc[o] - O.fac * O.psc[o];
    }
  }
}

static int cpml_advance(epb_handle *h, double tstep, bool efield) {
  const epb_config &c = h->cfg;
  const double cc = EPB_C;
  for (int a = 0; a < c.ndims; a++) {
    const int b = (a + 1) % 3, c3 = (a + 2) % 3;
    for (int sd = 0; sd < 2; sd++) {
      const int bc = c.bc_field[2 * a + sd];
      if (bc != EPB_BC_CPML_LASER && bc != EPB_BC_CPML_OUTFLOW) continue;
      int i0 = h->cp_start[a][sd], i1 = h->cp_end[a][sd];
      if (!efield && sd == 1) { i0 -= 1; i1 -= 1; }
      if (i1 < i0) continue;
      CpmlOp O;
      O.fb = h->f((efield ? EPB_EX : EPB_BX) + b);
      O.fc = h->f((efield ? EPB_EX : EPB_BX) + c3);
      O.gb = h->f((efield ? EPB_BX : EPB_EX) + b);
      O.gc = h->f((efield ? EPB_BX : EPB_EX) + c3);
      O.psb = h->cp_psi[a] + (size_t)(efield ? 0 : 2) * h->fsize;
      O.psc = h->cp_psi[a] + (size_t)(efield ? 1 : 3) * h->fsize;
      O.bco = h->cp_bco[a][efield ? 0 : 1];
      O.cco = h->cp_cco[a][efield ? 0 : 1];
      O.nd = c.ndims;
      for (int d = 0; d < 3; d++) { O.sz[d] = h->sz[d]; O.n[d] = c.n[d]; }
      O.a = a; O.i0 = i0; O.i1 = i1; O.efield = efield ? 1 : 0;
      O.fac = efield ? tstep * (cc * cc) : tstep;
      size_t total = (size_t)(i1 - i0 + 1);
      for (int d = 0; d < c.ndims; d++) if (d != a) total *= (size_t)c.n[d];
      k_cpml<<<nblocks(total, 148 * 8), 256, 0, h->stream>>>(O);
      h->launches++;
    }
  }
  return EPB_OK;
}

static void fd_coeffs(int order, double base, double *cc) {
  if (order == 4) { cc[0] = (9.0 / 8.0) * base; cc[1] = (-1.0 / 24.0) * base; cc[2] = 0.0; }
  else if (order == 6) { cc[0] = (75.0 / 64.0) * base; cc[1] = (-25.0 / 384.0) * base; cc[2] = (3.0 / 640.0) * base; }
  else { cc[0] = base; cc[1] = 0.0; cc[2] = 0.0; }
}
// true: the deck asks for something the TMA order-2 Yee kernels do not cover
static bool general_solver(const epb_handle *h, FieldParams &F) {
  const epb_config &c = h->cfg;
  const int order = c.field_order ? c.field_order : 2;
  F.nt = order / 2;
  F.ext = c.maxwell_solver != 0 ? 1 : 0;
  fd_coeffs(order, F.cx, F.kx);
  fd_coeffs(order, F.cy, F.ky);
  fd_coeffs(order, F.cz, F.kz);
  for (int q = 0; q < 3; q++) { F.alpha[q] = c.stencil[q]; F.gamma[q] = c.stencil[9 + q]; F.delta[q] = c.stencil[12 + q]; }
  for (int q = 0; q < 6; q++) F.beta[q] = c.stencil[3 + q];
  fd_coeffs(order, 1.0, F.ck);
  for (int q = 0; q < 3; q++) F.kap[q] = nullptr;
  return order != 2 || F.ext || h->cpml;
}

int update_e(epb_handle *h, double hdt) {
  FieldParams F;
  fill_field_params(h, F);
  const double c = EPB_C;
  F.cx = hdt / h->cfg.dx[0] * (c * c);
  F.cy = F.nd >= 2 ? hdt / h->cfg.dx[1] * (c * c) : 0.0;
  F.cz = F.nd >= 3 ? hdt / h->cfg.dx[2] * (c * c) : 0.0;
  F.fac = hdt / EPB_EPS0;
  if (general_solver(h, F)) {
    if (h->cpml) for (int q = 0; q < F.nd; q++) F.kap[q] = h->cp_kap[q][0];
    size_t tot = (size_t)(F.n[0] + 1) * (F.nd >= 2 ? F.n[1] + 1 : 1) * (F.nd >= 3 ? F.n[2] + 1 : 1);
    int nbg = nblocks(tot, 148 * 32);
    if (F.nd == 1) k_update_e_gen<1><<<nbg, 256, 0, h->stream>>>(F);
    else if (F.nd == 2) k_update_e_gen<2><<<nbg, 256, 0, h->stream>>>(F);
    else k_update_e_gen<3><<<nbg, 256, 0, h->stream>>>(F);
    h->launches++;
    if (h->cpml) cpml_advance(h, hdt, true);   // fields.f90:204
    return EPB_OK;
  }
  if (h->tma_ok) { epb_fdtd_tma_launch(h, true, F.cx, F.cy, F.cz, F.fac); return EPB_OK; }
  size_t total = (size_t)(F.n[0] + 1) * (F.nd >= 2 ? F.n[1] + 1 : 1) * (F.nd >= 3 ? F.n[2] + 1 : 1);
  int nb = nblocks(total, 148 * 32);
  if (F.nd == 1) k_update_e<1><<<nb, 256, 0, h->stream>>>(F);
  else if (F.nd == 2) k_update_e<2><<<nb, 256, 0, h->stream>>>(F);
  else k_update_e<3><<<nb, 256, 0, h->stream>>>(F);
  h->launches++;
  return EPB_OK;
}

int update_b(epb_handle *h, double hdt) {
  FieldParams F;
  fill_field_params(h, F);
  F.cx = hdt / h->cfg.dx[0];
  F.cy = F.nd >= 2 ? hdt / h->cfg.dx[1] : 0.0;
  F.cz = F.nd >= 3 ? hdt / h->cfg.dx[2] : 0.0;
  F.fac = 0.0;
  if (general_solver(h, F)) {
    if (h->cpml) for (int q = 0; q < F.nd; q++) F.kap[q] = h->cp_kap[q][1];
    size_t tot = (size_t)(F.n[0] + 1) * (F.nd >= 2 ? F.n[1] + 1 : 1) * (F.nd >= 3 ? F.n[2] + 1 : 1);
    int nbg = nblocks(tot, 148 * 32);
    if (F.nd == 1) k_update_b_gen<1><<<nbg, 256, 0, h->stream>>>(F);
    else if (F.nd == 2) k_update_b_gen<2><<<nbg, 256, 0, h->stream>>>(F);
    else k_update_b_gen<3><<<nbg, 256, 0, h->stream>>>(F);
    h->launches++;
    if (h->cpml) cpml_advance(h, hdt, false);   // after the B update (fields.f90, end of update_b_field)
    return EPB_OK;
  }
  if (h->tma_ok) { epb_fdtd_tma_launch(h, false, F.cx, F.cy, F.cz, F.fac); return EPB_OK; }
  size_t total = (size_t)(F.n[0] + 1) * (F.nd >= 2 ? F.n[1] + 1 : 1) * (F.nd >= 3 ? F.n[2] + 1 : 1);
  int nb = nblocks(total, 148 * 32);
  if (F.nd == 1) k_update_b<1><<<nb, 256, 0, h->stream>>>(F);
  else if (F.nd == 2) k_update_b<2><<<nb, 256, 0, h->stream>>>(F);
  else k_update_b<3><<<nb, 256, 0, h->stream>>>(F);
  h->launches++;
  return EPB_OK;
}

inline bool stagger(int dir, int field) {  // setup.F90:124-134
  switch (field) {
    case EPB_EX: return dir == 0;
    case EPB_EY: return dir == 1;
    case EPB_EZ: return dir == 2;
    case EPB_BX: return dir != 0;
    case EPB_BY: return dir != 1;
    case EPB_BZ: return dir != 2;
  }
  return false;
}

int mirror3(epb_handle *h, int f0, int boundary, double sign, int conduct = 0) {
  const epb_config &c = h->cfg;
  if (c.bc_field[boundary] == EPB_BC_PERIODIC) return EPB_OK;
  if (!c.is_boundary[boundary]) return EPB_OK;
  MirrorOp M;
  M.nd = c.ndims;
  M.d = boundary / 2;
  M.is_max = boundary & 1;
  // conduct: c_bc_conduct (boundary.F90:817-832, :870-885) clamps the E component normal to the wall and the
  // B components along it, and gives the others a zero gradient
  for (int q = 0; q < 3; q++) {
    M.sign[q] = sign;
    if (conduct) {
      const bool normal = (q == M.d);
      M.sign[q] = (f0 == EPB_EX) ? (normal ? -1.0 : +1.0) : (normal ? +1.0 : -1.0);
    }
  }
  size_t total = 1;
  for (int d = 0; d < 3; d++) {
    M.sz[d] = h->sz[d];
    M.n[d] = c.n[d];
    if (d < c.ndims && d != M.d) total *= h->sz[d];
  }
  for (int q = 0; q < 3; q++) { M.f[q] = h->f(f0 + q); M.stag[q] = stagger(M.d, f0 + q); }
  k_mirror<<<nblocks(total), 256, 0, h->stream>>>(M);
  h->launches++;
  return EPB_OK;
}

// efield_bcs / bfield_bcs (boundary.F90:808-907)
int field_bcs3(epb_handle *h, int f0, bool mpi_only) {
  int rc = epb_halo_exchange(h, f0, 3, false);
  if (rc) return rc;
  if (mpi_only) return EPB_OK;
  for (int i = 0; i < 2 * h->cfg.ndims; i++)
    if (h->cfg.bc_field[i] == EPB_BC_CONDUCT) mirror3(h, f0, i, 0.0, 1);
  for (int i = 0; i < 2 * h->cfg.ndims; i++) {
    int b = h->cfg.bc_field[i];
    if (b == EPB_BC_CLAMP || b == EPB_BC_SIMPLE_LASER || b == EPB_BC_SIMPLE_OUTFLOW) mirror3(h, f0, i, -1.0);
    if (b == EPB_BC_ZERO_GRADIENT || b == EPB_BC_CPML_LASER || b == EPB_BC_CPML_OUTFLOW) mirror3(h, f0, i, +1.0);   // boundary.F90:845-851, :898-904
  }
  return EPB_OK;
}

int outflow_x(epb_handle *h, int side, double dt) {
  const epb_config &c = h->cfg;
  OutflowOp O;
  for (int q = 0; q < 9; q++) O.f[q] = h->f(q);
  O.snap = h->snap + (size_t)side * 6 * h->plane;
  O.s1 = h->src + ((size_t)side * 2 + 0) * h->plane;
  O.s2 = h->src + ((size_t)side * 2 + 1) * h->plane;
  O.nd = c.ndims;
  for (int d = 0; d < 3; d++) { O.sz[d] = h->sz[d]; O.n[d] = c.n[d]; }
  O.plane = h->plane;
  O.is_max = side;
  O.lp = c.bc_field[side] == EPB_BC_CPML_LASER ? h->cp_laser_idx[0][side] : (side ? c.n[0] : 1);
  const double cc = EPB_C;
  const double dtc2 = dt * (cc * cc);
  O.lx = dtc2 / c.dx[0];
  O.ly = c.ndims >= 2 ? dtc2 / c.dx[1] : 0.0;
  O.lz = c.ndims >= 3 ? dtc2 / c.dx[2] : 0.0;
  O.sum = 1.0 / (O.lx + cc);
  O.diff = O.lx - cc;
  O.dt_eps = dt / EPB_EPS0;
  size_t total = (size_t)(c.ndims >= 2 ? c.n[1] + 1 : 1) * (c.ndims >= 3 ? c.n[2] + 1 : 1);
  int nb = nblocks(total);
  if (c.ndims == 1) k_outflow<1><<<nb, 256, 0, h->stream>>>(O);
  else if (c.ndims == 2) k_outflow<2><<<nb, 256, 0, h->stream>>>(O);
  else k_outflow<3><<<nb, 256, 0, h->stream>>>(O);
  h->launches++;
  return EPB_OK;
}

static void fill_face_op(epb_handle *h, int a, FaceOp &O) {
  const epb_config &c = h->cfg;
  for (int q = 0; q < 9; q++) O.f[q] = h->f(q);
  O.snap = h->snapA[a];
  O.nd = c.ndims;
  for (int d = 0; d < 3; d++) { O.sz[d] = h->sz[d]; O.n[d] = c.n[d]; }
  O.a = a;
  O.plane = h->planeA[a];
  for (int sd = 0; sd < 2; sd++) {
    O.snap_i[sd] = sd ? c.n[a] : 1;
    if (c.bc_field[2 * a + sd] == EPB_BC_CPML_LASER) O.snap_i[sd] = sd ? h->cp_laser_idx[a][1] + 1 : h->cp_laser_idx[a][0] - 1;
  }
  O.lp = 1;
}
int outflow_face(epb_handle *h, int boundary, double dt) {
  const epb_config &c = h->cfg;
  const int a = boundary / 2, side = boundary & 1;
  FaceOp O;
  fill_face_op(h, a, O);
  O.is_max = side;
  O.lp = c.bc_field[boundary] == EPB_BC_CPML_LASER ? h->cp_laser_idx[a][side] : (side ? c.n[a] : 1);
  O.s1 = h->srcA[a] + ((size_t)side * 2 + 0) * h->planeA[a];
  O.s2 = h->srcA[a] + ((size_t)side * 2 + 1) * h->planeA[a];
  const double cc = EPB_C;
  const double dtc2 = dt * (cc * cc);
  for (int d = 0; d < 3; d++) O.l[d] = d < c.ndims ? dtc2 / c.dx[d] : 0.0;
  O.sum = 1.0 / (O.l[a] + cc);
  O.diff = O.l[a] - cc;
  O.dt_eps = dt / EPB_EPS0;
  size_t total = 1;
  for (int d = 0; d < c.ndims; d++) if (d != a) total *= (size_t)(c.n[d] + 1);
  const int nb = nblocks(total);
  if (c.ndims == 2) k_outflow_face<2><<<nb, 256, 0, h->stream>>>(O);
  else k_outflow_face<3><<<nb, 256, 0, h->stream>>>(O);
  h->launches++;
  return EPB_OK;
}

// bfield_final_bcs (boundary.F90:911-944)
int bfield_final_bcs(epb_handle *h, double dt) {
  int rc = field_bcs3(h, EPB_BX, false);
  if (rc) return rc;
  // add_laser(i) .OR. simple_outflow (boundary.F90:918-940): a cpml_laser face has add_laser only on the rank that
  // holds its laser plane (boundary.F90:1572-1577); a cpml_outflow face only absorbs
  for (int side = 0; side < 2; side++) {
    int b = h->cfg.bc_field[side];
    if (h->cfg.is_boundary[side] && (b == EPB_BC_SIMPLE_LASER || b == EPB_BC_SIMPLE_OUTFLOW ||
                                     (b == EPB_BC_CPML_LASER && h->cp_add_laser[0][side]))) outflow_x(h, side, dt);
  }
  for (int bd = 2; bd < 2 * h->cfg.ndims; bd++) {
    int b = h->cfg.bc_field[bd];
    if (h->cfg.is_boundary[bd] && (b == EPB_BC_SIMPLE_LASER || b == EPB_BC_SIMPLE_OUTFLOW ||
                                   (b == EPB_BC_CPML_LASER && h->cp_add_laser[bd / 2][bd & 1]))) outflow_face(h, bd, dt);
  }
  return field_bcs3(h, EPB_BX, true);
}

int bc_allspecies(const epb_handle *h, int i) {  // deck_species_block.F90:182-199
  if (h->sp.empty()) return h->cfg.bc_field[i] == EPB_BC_PERIODIC ? EPB_BC_PERIODIC : EPB_BC_OPEN;
  int b = h->sp[h->bc_species >= 0 ? h->bc_species : 0].cfg.bc_particle[i];
  if (b != EPB_BC_REFLECT && b != EPB_BC_PERIODIC) b = EPB_BC_OPEN;
  return b;
}

void fill_push_params(epb_handle *h, int is, PushParams &P) {
  const epb_config &c = h->cfg;
  SpeciesDev &S = h->sp[is];
  const int nd = c.ndims;
  const double cc = EPB_C;
  memset(&P, 0, sizeof P);
  P.nd = nd;
  for (int d = 0; d < 3; d++) {
    P.n[d] = c.n[d];
    P.sz[d] = h->sz[d];
    P.e[d] = h->f(EPB_EX + d);
    P.b[d] = h->f(EPB_BX + d);
    P.j[d] = h->f(EPB_JX + d);
    P.idx[d] = d < nd ? 1.0 / c.dx[d] : 0.0;
    P.grid_min_local[d] = c.grid_min_local[d];
    P.x[d] = S.buf[S.cur][d];
    P.p[d] = S.buf[S.cur][3 + d];
  }
  P.w = S.buf[S.cur][6];
  // particles.F90:128-136, 155-167 (1D: epoch1d particles.F90:151-158; 3D: epoch3d :162-174)
  double fac = 1.0;
  for (int d = 0; d < nd; d++) fac *= 0.5;
  const double dt = c.dt;
  const double idt = 1.0 / dt;
  P.dto2 = dt / 2.0;
  P.dtco2 = cc * P.dto2;
  const double dtfac = 0.5 * dt * fac;
  P.third = 1.0 / 3.0;
  if (nd == 1) { P.kfc[0] = idt * fac; P.kfc[1] = P.idx[0] * fac; P.kfc[2] = 0.0; }
  else if (nd == 2) { P.kfc[0] = idt * P.idx[1] * fac; P.kfc[1] = idt * P.idx[0] * fac; P.kfc[2] = P.idx[0] * P.idx[1] * fac; }
  else { P.kfc[0] = idt * P.idx[1] * P.idx[2] * fac; P.kfc[1] = idt * P.idx[0] * P.idx[2] * fac; P.kfc[2] = idt * P.idx[0] * P.idx[1] * fac; }
  P.part_q = S.cfg.charge;
  P.part_mc = cc * S.cfg.mass;
  P.ipart_mc = 1.0 / P.part_mc;
  P.cmratio = P.part_q * dtfac * P.ipart_mc;
  P.ccmratio = cc * P.cmratio;
  P.deposit = !S.cfg.zero_current;
  P.hc_push = c.hc_push ? 1 : 0;
  P.hc_alpha = 0.5 * P.part_q * dt / S.cfg.mass;
  P.tile_start = S.tile_start;
  P.cell_start = S.cell_start;
  P.emit = 0;
  P.perm = nullptr;
  for (int d = 0; d < 3; d++) { P.xs[d] = P.x[d]; P.ps[d] = P.p[d]; }
  P.ws = P.w;
  P.key_out = S.key;
  P.rank_out = S.rank;
  P.stay_cnt = S.stay_cnt;
  P.arr_cnt = S.arr_cnt;
  P.tg = h->tg;
  for (int d = 0; d < 3; d++) {
    P.bnd_min[d] = c.min_local[d];
    P.bnd_max[d] = c.max_local[d];
    P.min_local[d] = c.min_local[d];
    P.max_local[d] = c.max_local[d];
    P.gmin[d] = c.gmin[d];
    P.gmax[d] = c.gmax[d];
    P.shift[d] = (c.gmax[d] - c.gmin[d]) + 2.0 * c.dx[d] * (double)c.cpml_thickness;  // length_x + 2 dx cpml_thickness (boundary.F90:1047-1048)
    P.min_outer[d] = c.min_outer[d];
    P.max_outer[d] = c.max_outer[d];
    P.bc_min[d] = S.cfg.bc_particle[2 * d];
    P.bc_max[d] = S.cfg.bc_particle[2 * d + 1];
    P.is_bnd_min[d] = c.is_boundary[2 * d];
    P.is_bnd_max[d] = c.is_boundary[2 * d + 1];
  }
  for (int q = 0; q < 27; q++) {
    P.nbr_is_self[q] = (c.neighbour[q] == c.rank);
    P.nbr_valid[q] = (c.neighbour[q] >= 0);
  }
  P.out_count = h->out_count;
  P.out_idx = h->out_idx;
  P.out_cap = h->out_cap;
  P.gone = S.gone;
  static int experiment = epb_env("EPB_PUSH_EXPERIMENT") ? atoi(epb_env("EPB_PUSH_EXPERIMENT")) : 0;
  P.experiment = experiment;
}

}  // namespace

// ---------------------------------------------------------------------------
// Ghost-cell exchange on a single rank (self-wrap); multi-rank path in exchange.cu
// ---------------------------------------------------------------------------
int epb_halo_local(epb_handle *h, int f0, int nf, bool add, int d, int pass) {
  const epb_config &c = h->cfg;
  BoxOp B;
  B.nf = nf;
  B.nd = c.ndims;
  B.add = add ? 1 : 0;
  for (int q = 0; q < nf; q++) B.f[q] = h->f(f0 + q);
  for (int q = 0; q < 3; q++) {
    B.sz[q] = h->sz[q];
    B.slo[q] = B.dlo[q] = (q < c.ndims) ? 1 - NG : 1;
    B.ext[q] = h->sz[q];
  }
  B.ext[d] = NG;
  const int n = c.n[d];
  if (!add) {
    // do_field_mpi_with_lengths: pass 0 low interior -> high ghosts, pass 1 high interior -> low ghosts
    if (pass == 0) { B.slo[d] = 1; B.dlo[d] = n + 1; }
    else { B.slo[d] = n + 1 - NG; B.dlo[d] = 1 - NG; }
  } else {
    // particle_periodic_bcs: pass 0 high ghosts -> += low interior, pass 1 low ghosts -> += high interior
    if (pass == 0) { B.slo[d] = n + 1; B.dlo[d] = 1; }
    else { B.slo[d] = 1 - NG; B.dlo[d] = n + 1 - NG; }
  }
  size_t total = (size_t)B.ext[0] * B.ext[1] * B.ext[2];
  k_box<<<nblocks(total), 256, 0, h->stream>>>(B);
  h->launches++;
  return EPB_OK;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

const char *epb_version(void) { return "epoch_b200 0.1 (sm_100a)"; }
int epb_abi_info(int32_t out[4]) {
  if (!out) return EPB_ERR_ARG;
  out[0] = (int32_t)sizeof(epb_config);
  out[1] = (int32_t)sizeof(epb_species);
  out[2] = NG;
  out[3] = EPB_NFIELD;
  return EPB_OK;
}
const char *epb_last_error(const epb_handle *h) { return h ? h->err.c_str() : "null handle"; }
int64_t epb_launch_count(const epb_handle *h) { return h ? h->launches : 0; }

static int create_device_state(epb_handle *h, const epb_config *cfg, const epb_species *species);

int epb_create(const epb_config *cfg, const epb_species *species, epb_handle **out) {
  if (!cfg || !out) return EPB_ERR_ARG;
  *out = nullptr;
  if (cfg->ndims < 1 || cfg->ndims > 3) return epb_fail(nullptr, EPB_ERR_ARG, "ndims must be 1..3");
  if (cfg->ng != NG) return epb_fail(nullptr, EPB_ERR_UNSUPPORTED, "ng must be %d (triangle shape)", NG);
  {
    const int fo = cfg->field_order;
    if (fo != 0 && fo != 2 && fo != 4 && fo != 6)
      return epb_fail(nullptr, EPB_ERR_UNSUPPORTED, "field_order %d (2, 4 or 6)", fo);
    for (int q = 0; q < 4; q++)
      if (((cfg->smooth_strides >> (4 * q)) & 15) > NG)
        return epb_fail(nullptr, EPB_ERR_UNSUPPORTED, "smoothing stride > %d ghost cells", NG);
    if (cfg->maxwell_solver != 0 && fo != 0 && fo != 2)
      return epb_fail(nullptr, EPB_ERR_UNSUPPORTED,
                      "extended Maxwell stencils (maxwell_solver %d) exist for field_order 2 only", cfg->maxwell_solver);
  }
  for (int d = 0; d < cfg->ndims; d++)
    if (cfg->n[d] < NG) return epb_fail(nullptr, EPB_ERR_UNSUPPORTED, "local extent %d < ng", cfg->n[d]);
  for (int i = 0; i < 2 * cfg->ndims; i++) {
    int b = cfg->bc_field[i];
    bool ok = b == EPB_BC_PERIODIC || b == EPB_BC_CLAMP || b == EPB_BC_ZERO_GRADIENT ||
              b == EPB_BC_SIMPLE_LASER || b == EPB_BC_SIMPLE_OUTFLOW || b == EPB_BC_CONDUCT ||
              b == EPB_BC_CPML_LASER || b == EPB_BC_CPML_OUTFLOW;
    if (!ok) return epb_fail(nullptr, EPB_ERR_UNSUPPORTED, "field boundary code %d on boundary %d not implemented on the device path", b, i);
  }
  for (int s = 0; s < cfg->n_species; s++) {
    for (int i = 0; i < 2 * cfg->ndims; i++) {
      int b = species[s].bc_particle[i];
      if (!(b == EPB_BC_PERIODIC || b == EPB_BC_REFLECT || b == EPB_BC_OPEN || b == EPB_BC_THERMAL))
        return epb_fail(nullptr, EPB_ERR_UNSUPPORTED, "particle boundary code %d not implemented on the device path", b);
    }
  }
  // everything that can fail after the handle exists runs in create_device_state, so that a failure there
  // (out of memory, a CUDA error) releases what was already allocated instead of leaking a half-built handle
  epb_handle *h = new epb_handle;
  h->cfg = *cfg;
  const int rc = create_device_state(h, cfg, species);
  if (rc != EPB_OK) {
    fprintf(stderr, "epoch_b200: epb_create failed: %s\n", h->err.c_str());
    epb_destroy(h);
    return rc;
  }
  *out = h;
  return EPB_OK;
}

static int create_device_state(epb_handle *h, const epb_config *cfg, const epb_species *species) {
  {  // c_bc_mixed: do the species disagree on a particle boundary?
    auto norm = [](int b) { return (b == EPB_BC_REFLECT || b == EPB_BC_PERIODIC) ? b : EPB_BC_OPEN; };
    for (int s = 1; s < cfg->n_species; s++)
      for (int i = 0; i < 2 * cfg->ndims; i++)
        if (norm(species[s].bc_particle[i]) != norm(species[0].bc_particle[i])) h->bc_mixed = true;
    if (cfg->n_species > 0 && epb_env("EPB_FORCE_MIXED") && atoi(epb_env("EPB_FORCE_MIXED"))) h->bc_mixed = true;
  }
  const int nd = cfg->ndims;
  h->fsize = 1;
  for (int d = 0; d < 3; d++) {
    h->sz[d] = d < nd ? cfg->n[d] + 2 * NG : 1;
    if (d >= nd) h->cfg.n[d] = 1;
    h->fsize *= h->sz[d];
  }
  h->plane = (size_t)h->sz[1] * h->sz[2];
  EPB_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  // ex..jz, plus five work arrays (field ids 9..13): smooth_current uses 9..11, epb_calc_moment 9..13
  const int nfields = 14;
  EPB_CUDA(h, cudaMalloc(&h->fields, nfields * h->fsize * sizeof(double)));
  EPB_CUDA(h, cudaMemsetAsync(h->fields, 0, nfields * h->fsize * sizeof(double), h->stream));
  EPB_CUDA(h, cudaMalloc(&h->snap, 12 * h->plane * sizeof(double)));
  EPB_CUDA(h, cudaMemsetAsync(h->snap, 0, 12 * h->plane * sizeof(double), h->stream));
  EPB_CUDA(h, cudaMalloc(&h->src, 4 * h->plane * sizeof(double)));
  EPB_CUDA(h, cudaMemsetAsync(h->src, 0, 4 * h->plane * sizeof(double), h->stream));
  for (int a = 1; a < nd; a++) {  // y / z faces
    size_t pl = 1;
    for (int d = 0; d < nd; d++) if (d != a) pl *= (size_t)h->sz[d];
    h->planeA[a] = pl;
    EPB_CUDA(h, cudaMalloc(&h->snapA[a], 12 * pl * sizeof(double)));
    EPB_CUDA(h, cudaMemsetAsync(h->snapA[a], 0, 12 * pl * sizeof(double), h->stream));
    EPB_CUDA(h, cudaMalloc(&h->srcA[a], 4 * pl * sizeof(double)));
    EPB_CUDA(h, cudaMemsetAsync(h->srcA[a], 0, 4 * pl * sizeof(double), h->stream));
  }
  {  // set_cpml_helpers + allocate_cpml_fields (boundary.F90:1479-1794; epoch3d :1891-2330, epoch1d :783-925), per axis
    const int t = cfg->cpml_thickness;
    bool any = false;
    for (int i = 0; i < 2 * nd; i++)
      if (cfg->bc_field[i] == EPB_BC_CPML_LASER || cfg->bc_field[i] == EPB_BC_CPML_OUTFLOW) any = true;
    for (int d = 0; d < 3; d++)
      for (int sd = 0; sd < 2; sd++) { h->cp_start[d][sd] = cfg->n[d] + 1; h->cp_end[d][sd] = 0; }
    if (any && t <= 0) return epb_fail(h, EPB_ERR_ARG, "a CPML field boundary needs cpml_thickness > 0");
    if (any) {
      h->cpml = true;
      const double cc = EPB_C;
      const int cpml_m = 3, cpml_ma = 1;
      // fng: field_order / 2 (fields.f90:37), but 2 for the Lehe solvers (deck_control_block.F90:117-120)
      int fng = (cfg->field_order ? cfg->field_order : 2) / 2;
      if (cfg->maxwell_solver >= 2 && cfg->maxwell_solver <= 4) fng = 2;   // c_maxwell_solver_lehe_x / _y / _z
      const double tstep = 0.5 * cfg->dt;                               // both half steps use hdt
      auto pw = [](double x, int e) { double r = 1.0; for (int q = 0; q < e; q++) r = r * x; return r; };
      // dx of the first axis for every axis, as boundary.F90:1517 has it
      const double sigma_maxval = cfg->cpml_sigma_max * cc * 0.8 * (cpml_m + 1.0) / cfg->dx[0];
      for (int d = 0; d < nd; d++) {
        const int n = cfg->n[d], len = n + 2 * NG;
        std::vector<double> kap[2], sig[2], aa[2];
        for (int q = 0; q < 2; q++) { kap[q].assign(len, 1.0); sig[q].assign(len, 0.0); aa[q].assign(len, 0.0); }
        auto at = [&](std::vector<double> &v, int i) -> double & { return v[i + NG - 1]; };
        const int gmin = cfg->n_global_min[d], gmax = gmin + n - 1, ng_ = cfg->n_global[d];
        auto profile = [&](int i, int ib, int ig) {   // E point i, B point ib, distance index ig into the layer
          double x_pos = 1.0 - (double)(ig - 1) / (double)t;
          at(kap[0], i) = 1.0 + (cfg->cpml_kappa_max - 1.0) * pw(x_pos, cpml_m);
          at(sig[0], i) = sigma_maxval * pw(x_pos, cpml_m);
          at(aa[0], i) = cfg->cpml_a_max * pw(1.0 - x_pos, cpml_ma);
          x_pos = 1.0 - ((double)ig - 0.5) / (double)t;
          at(kap[1], ib) = 1.0 + (cfg->cpml_kappa_max - 1.0) * pw(x_pos, cpml_m);
          at(sig[1], ib) = sigma_maxval * pw(x_pos, cpml_m);
          at(aa[1], ib) = cfg->cpml_a_max * pw(1.0 - x_pos, cpml_ma);
        };
        const int bmin = cfg->bc_field[2 * d], bmax = cfg->bc_field[2 * d + 1];
        if (bmin == EPB_BC_CPML_LASER || bmin == EPB_BC_CPML_OUTFLOW) {
          if (gmin <= t) {
            h->cp_start[d][0] = 1;
            h->cp_end[d][0] = gmax >= t ? t - gmin + 1 : n;
            for (int i = h->cp_start[d][0]; i <= h->cp_end[d][0]; i++) profile(i, i, i + gmin - 1);
          }
          if (gmin <= t + fng + 1 && gmax >= t + fng + 1) { h->cp_add_laser[d][0] = true; h->cp_laser_idx[d][0] = t + fng + 1 - gmin; }
        }
        if (bmax == EPB_BC_CPML_LASER || bmax == EPB_BC_CPML_OUTFLOW) {
          if (gmax >= ng_ - t + 1) {
            h->cp_end[d][1] = n;
            h->cp_start[d][1] = gmin <= ng_ - t + 1 ? ng_ - t + 1 - gmin + 1 : 1;
            for (int i = h->cp_start[d][1]; i <= h->cp_end[d][1]; i++) profile(i, i - 1, ng_ - (i + gmin - 1) + 1);
          }
          if (gmin <= ng_ - t - fng + 2 && gmax >= ng_ - t - fng + 2) {
            h->cp_add_laser[d][1] = true;
            h->cp_laser_idx[d][1] = ng_ - t - fng + 2 - gmin;
          }
        }
        for (int q = 0; q < 2; q++) {
          // bcoeff = EXP(-(sigma / kappa + acoeff) * tstep), ccoeff_d = (bcoeff - 1) * sigma / kappa / (sigma + kappa * acoeff) / dx
          // (boundary.F90:1832-1835); outside the layers (sigma = a = 0) the values are never read
          std::vector<double> bco(len, 0.0), cco(len, 0.0);
          for (int i = 0; i < len; i++) {
            const double kappa = kap[q][i], sigma = sig[q][i], acoeff = aa[q][i];
            if (sigma == 0.0 && acoeff == 0.0) continue;
            bco[i] = std::exp(-(sigma / kappa + acoeff) * tstep);
            cco[i] = (bco[i] - 1.0) * sigma / kappa / (sigma + kappa * acoeff) / cfg->dx[d];
          }
          EPB_CUDA(h, cudaMalloc(&h->cp_kap[d][q], len * sizeof(double)));
          EPB_CUDA(h, cudaMalloc(&h->cp_bco[d][q], len * sizeof(double)));
          EPB_CUDA(h, cudaMalloc(&h->cp_cco[d][q], len * sizeof(double)));
          EPB_CUDA(h, cudaMemcpy(h->cp_kap[d][q], kap[q].data(), len * sizeof(double), cudaMemcpyHostToDevice));
          EPB_CUDA(h, cudaMemcpy(h->cp_bco[d][q], bco.data(), len * sizeof(double), cudaMemcpyHostToDevice));
          EPB_CUDA(h, cudaMemcpy(h->cp_cco[d][q], cco.data(), len * sizeof(double), cudaMemcpyHostToDevice));
        }
        EPB_CUDA(h, cudaMalloc(&h->cp_psi[d], 4 * h->fsize * sizeof(double)));
        EPB_CUDA(h, cudaMemsetAsync(h->cp_psi[d], 0, 4 * h->fsize * sizeof(double), h->stream));
      }
    }
  }
  epb_fdtd_tma_setup(h);
  epb_make_tiles(h->cfg, h->tg);
  // 0 = library default: the cell-owner kernel wants a fresh order (its sort is cheap), the
  // transposition kernel tolerates a stale one
  if (h->cfg.sort_interval < 1) h->cfg.sort_interval = (h->tg.layout == 1) ? 2 : 8;   // layout 2 never sorts
  EPB_CUDA(h, cudaMalloc(&h->cell_count, ((size_t)h->tg.nkeys + 1) * sizeof(int)));
  EPB_CUDA(h, cudaMalloc(&h->cell_start, ((size_t)h->tg.nkeys + 1) * sizeof(int)));
  long long maxcap = 0;
  h->sp.resize(cfg->n_species);
  for (int s = 0; s < cfg->n_species; s++) {
    SpeciesDev &S = h->sp[s];
    S.cfg = species[s];
    S.cap = species[s].capacity > 0 ? species[s].capacity : 1024;
    if (S.cap >= (1LL << 31) - 1024) return epb_fail(h, EPB_ERR_CAPACITY, "species capacity must be < 2^31");
    maxcap = S.cap > maxcap ? S.cap : maxcap;
    if (h->tg.layout >= 2) {  // slot columns / tile bags: the arena is sized at the first upload / load (slots.cu)
      int rcs = epb_slots_alloc(h, s);
      if (!rcs) {  // a first arena from the mean occupancy the capacity implies; re-sized by upload / load if denser
        long long ncell = 1;
        for (int d = 0; d < nd; d++) ncell *= cfg->n[d];
        rcs = epb_slots_reset(h, s, 0, (int)std::min<long long>(S.cap / std::max<long long>(1, ncell), 4096));
      }
      if (rcs) { epb_destroy(h); return rcs; }
      continue;
    }
    for (int b = 0; b < 2; b++)
      for (int q = 0; q < 7; q++) {
        if (q < 3 && q >= nd) continue;
        EPB_CUDA(h, cudaMalloc(&S.buf[b][q], (size_t)S.cap * sizeof(double)));
      }
    EPB_CUDA(h, cudaMalloc(&S.key, (size_t)S.cap * sizeof(int)));
    EPB_CUDA(h, cudaMalloc(&S.tile_start, ((size_t)h->tg.ntiles + 1) * sizeof(int)));
    EPB_CUDA(h, cudaMemsetAsync(S.tile_start, 0, ((size_t)h->tg.ntiles + 1) * sizeof(int), h->stream));
    if (h->tg.layout == 1) {
      EPB_CUDA(h, cudaMalloc(&S.cell_start, ((size_t)h->tg.nkeys + 1) * sizeof(int)));
      EPB_CUDA(h, cudaMemsetAsync(S.cell_start, 0, ((size_t)h->tg.nkeys + 1) * sizeof(int), h->stream));
      EPB_CUDA(h, cudaMalloc(&S.rank, (size_t)S.cap * sizeof(int)));
      EPB_CUDA(h, cudaMalloc(&S.perm, (size_t)S.cap * sizeof(int)));
      EPB_CUDA(h, cudaMalloc(&S.stay_cnt, ((size_t)h->tg.nkeys + 1) * sizeof(int)));
      EPB_CUDA(h, cudaMalloc(&S.arr_cnt, ((size_t)h->tg.nkeys + 1) * sizeof(int)));
    }
    EPB_CUDA(h, cudaMalloc(&S.gone, (size_t)S.cap));
    EPB_CUDA(h, cudaMemsetAsync(S.gone, 0, (size_t)S.cap, h->stream));
  }
  // an outbox is only needed if a particle can leave this rank: a remote neighbour, or an
  // open physical boundary (deletion).  Fully periodic/reflecting single-rank runs skip the
  // per-step host read of the leaver counts altogether.
  bool needs_outbox = false;
  for (int q = 0; q < 27; q++)
    if (cfg->neighbour[q] >= 0 && cfg->neighbour[q] != cfg->rank) needs_outbox = true;
  for (int s = 0; s < cfg->n_species; s++)
    for (int i = 0; i < 2 * nd; i++)
      if (cfg->is_boundary[i] && species[s].bc_particle[i] == EPB_BC_OPEN) needs_outbox = true;
  h->out_cap = 0;
  if (needs_outbox) {
    h->out_cap = (int)(maxcap / 64 > 65536 ? maxcap / 64 : 65536);
    if (h->out_cap > maxcap && maxcap > 0) h->out_cap = (int)maxcap;
  }
  if (cfg->n_species > 0) {
    EPB_CUDA(h, cudaMalloc(&h->out_count, 64 * sizeof(int)));
    EPB_CUDA(h, cudaMemsetAsync(h->out_count, 0, 64 * sizeof(int), h->stream));
    EPB_CUDA(h, cudaMalloc(&h->out_idx, ((size_t)27 * h->out_cap + 1) * sizeof(int)));
    EPB_CUDA(h, cudaMalloc(&h->movers, ((size_t)27 * h->out_cap + 1) * sizeof(int)));
  }
  EPB_CUDA(h, cudaMalloc(&h->d_scratch, 1024 * sizeof(int)));
  EPB_CUDA(h, cudaMallocHost(&h->h_counts, 256 * sizeof(int)));
  EPB_CUDA(h, cudaEventCreate(&h->ev0));
  EPB_CUDA(h, cudaEventCreate(&h->ev1));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPB_OK;
}

int epb_destroy(epb_handle *h) {
  if (!h) return EPB_OK;
  cudaStreamSynchronize(h->stream);
  epb_comm_destroy(h);
  cudaFree(h->fields); cudaFree(h->snap); cudaFree(h->src);
  for (int a = 1; a < 3; a++) { cudaFree(h->snapA[a]); cudaFree(h->srcA[a]); }
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); cudaFree(h->dump_stage); cudaEventDestroy(h->dump_ready); cudaEventDestroy(h->dump_done); }
  cudaFree(h->cell_count); cudaFree(h->cell_start); cudaFree(h->cub_tmp); cudaFree(h->movers);
  cudaFree(h->out_count); cudaFree(h->out_idx); cudaFree(h->d_scratch); cudaFree(h->d_err); cudaFree(h->aos_stage); cudaFree(h->coll_work); cudaFree(h->prof_scratch); cudaFree(h->scal_dev);
  for (int d = 0; d < 3; d++) {
    cudaFree(h->cp_psi[d]);
    for (int q = 0; q < 2; q++) { cudaFree(h->cp_kap[d][q]); cudaFree(h->cp_bco[d][q]); cudaFree(h->cp_cco[d][q]); }
  }
  for (int q = 0; q < 4; q++) if (h->scal_ev[q]) cudaEventDestroy(h->scal_ev[q]);
  for (int q = 0; q < 8; q++) if (h->src_ev[q]) cudaEventDestroy(h->src_ev[q]);
  if (h->src_stage) cudaFreeHost(h->src_stage);
  cudaFree(h->sendbuf); cudaFree(h->recvbuf);
  for (auto &S : h->sp) {
    // slot columns first: their buf[0][*] point INTO the arena and are cleared by epb_slots_free
    if (S.slots) epb_slots_free(S);
    for (int q = 0; q < 6; q++) cudaFree(S.ext_temp[q]);
    for (int b = 0; b < 2; b++)
      for (int q = 0; q < 7; q++) cudaFree(S.buf[b][q]);
    cudaFree(S.key); cudaFree(S.tile_start); cudaFree(S.cell_start); cudaFree(S.rank); cudaFree(S.perm); cudaFree(S.stay_cnt); cudaFree(S.arr_cnt); cudaFree(S.gone);
  }
  if (h->h_counts) cudaFreeHost(h->h_counts);
  for (auto &e : h->ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return EPB_OK;
}

int epb_set_stream(epb_handle *h, void *s) {
  if (!h) return EPB_ERR_ARG;
  cudaStreamSynchronize(h->stream);
  if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  h->stream = (cudaStream_t)s;
  return EPB_OK;
}

int epb_synchronize(epb_handle *h) {
  if (!h) return EPB_ERR_ARG;
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPB_OK;
}

int epb_upload_field(epb_handle *h, int field, const double *host) {
  if (!h || field < 0 || field >= 9) return EPB_ERR_ARG;
  EPB_CUDA(h, cudaMemcpyAsync(h->f(field), host, h->fsize * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPB_OK;
}
int epb_download_field(epb_handle *h, int field, double *host) {
  if (!h || field < 0 || field >= 9) return EPB_ERR_ARG;
  EPB_CUDA(h, cudaMemcpyAsync(host, h->f(field), h->fsize * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPB_OK;
}
// Field dump that overlaps the following steps: the array is snapshotted on the device in stream
// order (so later kernels may overwrite it at once) and copied to the host on a second stream.
int epb_download_field_async(epb_handle *h, int field, double *host) {
  if (!h || field < 0 || field >= 9 || !host) return EPB_ERR_ARG;
  if (!h->copy_stream) {
    EPB_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    EPB_CUDA(h, cudaMalloc(&h->dump_stage, h->fsize * sizeof(double)));
    EPB_CUDA(h, cudaEventCreateWithFlags(&h->dump_ready, cudaEventDisableTiming));
    EPB_CUDA(h, cudaEventCreateWithFlags(&h->dump_done, cudaEventDisableTiming));
  }
  if (h->dump_pending) {  // one staging buffer: the previous dump must have left the device
    EPB_CUDA(h, cudaStreamWaitEvent(h->stream, h->dump_done, 0));
  }
  EPB_CUDA(h, cudaMemcpyAsync(h->dump_stage, h->f(field), h->fsize * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  EPB_CUDA(h, cudaEventRecord(h->dump_ready, h->stream));
  EPB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->dump_ready, 0));
  EPB_CUDA(h, cudaMemcpyAsync(host, h->dump_stage, h->fsize * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
  EPB_CUDA(h, cudaEventRecord(h->dump_done, h->copy_stream));
  h->dump_pending = true;
  return EPB_OK;
}
// blocks until the host buffers of all asynchronous dumps are complete
int epb_wait_downloads(epb_handle *h) {
  if (!h) return EPB_ERR_ARG;
  if (h->dump_pending) {
    EPB_CUDA(h, cudaEventSynchronize(h->dump_done));
    h->dump_pending = false;
  }
  return EPB_OK;
}
int epb_field_device_ptr(epb_handle *h, int field, void **dptr) {
  if (!h || field < 0 || field >= 9 || !dptr) return EPB_ERR_ARG;
  *dptr = h->f(field);
  return EPB_OK;
}

// pack_particle order (partlist.F90:414-486): pos(1..ndims), p(1..3), weight
int epb_upload_species(epb_handle *h, int is, int64_t n, const double *packed) {
  if (!h || is < 0 || is >= (int)h->sp.size() || n < 0) return EPB_ERR_ARG;
  SpeciesDev &S = h->sp[is];
  if (n > S.cap) return epb_fail(h, EPB_ERR_CAPACITY, "species %d: %lld particles > capacity %lld", is, (long long)n, S.cap);
  if (S.slots) return epb_slots_upload(h, is, n, packed);
  const int nd = h->cfg.ndims, nv = nd + 4;
  std::vector<double> tmp((size_t)n);
  for (int q = 0; q < nv; q++) {
    for (int64_t i = 0; i < n; i++) tmp[i] = packed[i * nv + q];
    int comp = q < nd ? q : 3 + (q - nd);
    EPB_CUDA(h, cudaMemcpyAsync(S.buf[S.cur][comp], tmp.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  EPB_CUDA(h, cudaMemsetAsync(S.gone, 0, (size_t)S.cap, h->stream));
  S.n = n;
  S.n_sorted = 0;
  S.info_valid = false;
  S.pending_perm = false;
  h->pushes_since_sort = 1 << 30;  // force a sort before the next push
  return EPB_OK;
}
// Particles the host creates in mid-run (run_injectors, injectors.F90:150-330; insert_particles of the moving window,
// window.F90:182-320 -- both evaluate deck expressions and draw from the host's random stream, so they stay EPOCH's):
// appended to the species in the wire layout, like append_partlist onto attached_list.  The arena of a slot-layout
// species must exist (epb_upload_species / epb_load_uniform came first).
int epb_append_species(epb_handle *h, int is, int64_t n, const double *packed) {
  if (!h || is < 0 || is >= (int)h->sp.size() || n < 0 || (n > 0 && !packed)) return EPB_ERR_ARG;
  if (n == 0) return EPB_OK;
  SpeciesDev &S = h->sp[is];
  if (S.slots && !S.arena_ready) return epb_upload_species(h, is, n, packed);
  const int nv = h->cfg.ndims + 4;
  const int64_t CH = 2 << 20;
  if (!h->aos_stage) EPB_CUDA(h, cudaMalloc(&h->aos_stage, (size_t)CH * 7 * sizeof(double)));
  for (int64_t i0 = 0; i0 < n; i0 += CH) {
    const int64_t mm = std::min<int64_t>(CH, n - i0);
    EPB_CUDA(h, cudaMemcpyAsync(h->aos_stage, packed + i0 * nv, (size_t)mm * nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int rc = epb_species_insert_aos(h, is, h->aos_stage, mm);
    if (rc) return rc;
    EPB_CUDA(h, cudaStreamSynchronize(h->stream));   // the staging buffer and the caller's array are free again
  }
  return S.slots ? epb_slots_check(h) : EPB_OK;
}

int epb_download_species(epb_handle *h, int is, int64_t n, double *packed) {
  if (!h || is < 0 || is >= (int)h->sp.size()) return EPB_ERR_ARG;
  SpeciesDev &S = h->sp[is];
  if (S.slots) return epb_slots_download(h, is, n, packed);
  if (n > S.n) n = S.n;
  const int nd = h->cfg.ndims, nv = nd + 4;
  std::vector<double> tmp((size_t)n);
  for (int q = 0; q < nv; q++) {
    int comp = q < nd ? q : 3 + (q - nd);
    EPB_CUDA(h, cudaMemcpyAsync(tmp.data(), S.buf[S.cur][comp], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < n; i++) packed[i * nv + q] = tmp[i];
  }
  return EPB_OK;
}
int epb_species_count(epb_handle *h, int is, int64_t *n) {
  if (!h || is < 0 || is >= (int)h->sp.size() || !n) return EPB_ERR_ARG;
  if (h->sp[is].slots) {
    long long v = 0;
    int rc = epb_slots_count(h, is, &v);
    *n = v;
    return rc;
  }
  *n = h->sp[is].n;
  return EPB_OK;
}

int epb_load_uniform(epb_handle *h, int is, int32_t ppc_arg, double density, const double temp_k[3],
                     const double drift[3], uint64_t seed) {
  if (!h || is < 0 || is >= (int)h->sp.size() || ppc_arg == 0) return EPB_ERR_ARG;
  // ppc < 0: |ppc| particles per cell on average, every particle's cell drawn at random (Poisson counts: the state a
  // thermal plasma relaxes to, against the loader's exactly |ppc| per cell)
  const int32_t ppc = ppc_arg < 0 ? -ppc_arg : ppc_arg;
  const epb_config &c = h->cfg;
  SpeciesDev &S = h->sp[is];
  long long total = (long long)c.n[0] * c.n[1] * c.n[2] * ppc;
  if (total > S.cap) return epb_fail(h, EPB_ERR_CAPACITY, "load_uniform: %lld particles > capacity %lld", total, S.cap);
  LoadOp L;
  memset(&L, 0, sizeof L);
  for (int d = 0; d < 3; d++) {
    if (!S.slots) {
      L.x[d] = S.buf[S.cur][d];
      L.p[d] = S.buf[S.cur][3 + d];
    }
    L.nloc[d] = c.n[d];
    L.gmin_local[d] = c.grid_min_local[d];
    L.dx[d] = c.dx[d];
    L.stdev[d] = sqrt(temp_k[d] * EPB_KB * S.cfg.mass);
    L.drift[d] = drift[d];
  }
  if (!S.slots) L.w = S.buf[S.cur][6];
  L.nd = c.ndims;
  L.ppc = ppc;
  L.i0 = 0;
  L.i1 = total;
  L.mixed = ppc_arg < 0 ? 1 : (epb_env("EPB_LOAD_MIXED") ? atoi(epb_env("EPB_LOAD_MIXED")) : 0);
  double vol = 1.0;
  for (int d = 0; d < c.ndims; d++) vol *= c.dx[d];
  L.weight = density * vol / ppc;
  L.seed = seed + 0x632BE59BD9B4E019ull * (unsigned long long)(c.rank + 1);
  if (S.slots) {
    // slot columns: generated chunk by chunk into the mover buffer and inserted by k_deliver (slots.cu); the
    // mixed state has Poisson counts, so its densest cell is ~6 sigma above the mean
    const int hint = L.mixed ? ppc + (int)ceil(6.0 * sqrt((double)ppc)) : ppc;
    int rc = epb_slots_reset(h, is, total, hint);
    if (rc) return rc;
    long long i0 = 0;
    while (i0 < total) {
      int waiting = 0;
      rc = epb_slots_waiting(h, is, &waiting);
      if (rc) return rc;
      const long long mm = std::min<long long>(total - i0, S.mcap - waiting);
      if (mm <= 0) return epb_fail(h, EPB_ERR_CAPACITY, "load_uniform: the mover buffer is full of particles that do not fit their columns");
      for (int d = 0; d < 3; d++) {
        L.x[d] = S.mbuf[S.mcur][d] ? S.mbuf[S.mcur][d] + waiting : nullptr;
        L.p[d] = S.mbuf[S.mcur][3 + d] + waiting;
      }
      L.w = S.mbuf[S.mcur][6] + waiting;
      L.i0 = i0;
      L.i1 = i0 + mm;
      k_load_uniform<<<nblocks((size_t)mm, 148 * 32), 256, 0, h->stream>>>(L);
      h->launches++;
      rc = epb_slots_commit(h, is, waiting, mm);
      if (rc) return rc;
      i0 += mm;
    }
    EPB_CUDA(h, cudaGetLastError());
    return epb_slots_check(h);
  }
  k_load_uniform<<<nblocks((size_t)total, 148 * 32), 256, 0, h->stream>>>(L);
  h->launches++;
  EPB_CUDA(h, cudaMemsetAsync(S.gone, 0, (size_t)S.cap, h->stream));
  S.n = total;
  S.n_sorted = 0;
  S.info_valid = false;
  S.pending_perm = false;
  h->pushes_since_sort = 1 << 30;
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

int epb_cell_counts(epb_handle *h, int is, int32_t *out) {
  if (!h || is < 0 || is >= (int)h->sp.size() || !out) return EPB_ERR_ARG;
  const epb_config &c = h->cfg;
  SpeciesDev &S = h->sp[is];
  size_t ncell = (size_t)c.n[0] * c.n[1] * c.n[2];
  int *d_out;
  EPB_CUDA(h, cudaMalloc(&d_out, ncell * sizeof(int)));
  EPB_CUDA(h, cudaMemsetAsync(d_out, 0, ncell * sizeof(int), h->stream));
  CountOp C;
  for (int d = 0; d < 3; d++) {
    C.nloc[d] = c.n[d];
    C.gmin[d] = c.grid_min_local[d];
    C.dx[d] = c.dx[d];
  }
  C.nd = c.ndims;
  C.out = d_out;
  SlotView V[2];
  const int nv = epb_species_views(h, is, V);
  for (int v = 0; v < nv; v++) {
    for (int d = 0; d < 3; d++) C.x[d] = V[v].a[d];
    C.r = V[v].r;
    k_cell_counts<<<nblocks((size_t)V[v].r.n), 256, 0, h->stream>>>(C);
    h->launches++;
  }
  (void)S;
  EPB_CUDA(h, cudaMemcpyAsync(out, d_out, ncell * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_out);
  return EPB_OK;
}

int epb_set_laser_source(epb_handle *h, int side, const double *s1, const double *s2) {
  if (!h || side < 0 || side >= 2 * h->cfg.ndims) return EPB_ERR_ARG;
  const epb_config &c = h->cfg;
  const int a = side / 2, sd = side & 1;
  size_t n = 1;
  for (int d = 0; d < c.ndims; d++) if (d != a) n *= (size_t)(c.n[d] + 1);
  double *base = a == 0 ? h->src : h->srcA[a];
  const size_t plane = a == 0 ? h->plane : h->planeA[a];
  if (!base) return EPB_ERR_ARG;
  // The caller may reuse its buffers right after return, and the call must not wait for the stream (the host
  // enqueues step n+1 while step n runs): the planes go through a ring of page-locked staging slots; a slot is
  // reused eight calls later, after its copy's event.
  size_t maxn = 1;
  for (int q = 0; q < c.ndims; q++) {
    size_t m = 1;
    for (int d = 0; d < c.ndims; d++) if (d != q) m *= (size_t)(c.n[d] + 1);
    if (m > maxn) maxn = m;
  }
  if (!h->src_stage) {
    h->src_stage_slot = 2 * maxn;
    EPB_CUDA(h, cudaMallocHost(&h->src_stage, 8 * h->src_stage_slot * sizeof(double)));
    for (int q = 0; q < 8; q++) EPB_CUDA(h, cudaEventCreateWithFlags(&h->src_ev[q], cudaEventDisableTiming));
  }
  const int slot = (int)(h->src_calls % 8);
  if (h->src_calls >= 8) EPB_CUDA(h, cudaEventSynchronize(h->src_ev[slot]));
  double *st = h->src_stage + (size_t)slot * h->src_stage_slot;
  memcpy(st, s1, n * sizeof(double));
  memcpy(st + n, s2, n * sizeof(double));
  EPB_CUDA(h, cudaMemcpyAsync(base + ((size_t)sd * 2 + 0) * plane, st, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EPB_CUDA(h, cudaMemcpyAsync(base + ((size_t)sd * 2 + 1) * plane, st + n, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EPB_CUDA(h, cudaEventRecord(h->src_ev[slot], h->stream));
  h->src_calls++;
  return EPB_OK;
}

int epb_init_boundaries(epb_handle *h) {
  if (!h) return EPB_ERR_ARG;
  const epb_config &c = h->cfg;
  // setup_field_boundaries (setup.F90:391-447)
  SnapOp S;
  for (int q = 0; q < 6; q++) S.f[q] = h->f(q);
  S.snap = h->snap;
  S.nd = c.ndims;
  for (int d = 0; d < 3; d++) { S.sz[d] = h->sz[d]; S.n[d] = c.n[d]; }
  S.plane = h->plane;
  S.i0[0] = c.bc_field[0] == EPB_BC_CPML_LASER ? h->cp_laser_idx[0][0] - 1 : 1;
  S.i0[1] = c.bc_field[1] == EPB_BC_CPML_LASER ? h->cp_laser_idx[0][1] + 1 : c.n[0];
  k_snapshot<<<nblocks(h->plane), 256, 0, h->stream>>>(S);
  h->launches++;
  for (int a = 1; a < c.ndims; a++) {
    if (!h->snapA[a]) continue;
    FaceOp O;
    fill_face_op(h, a, O);
    O.is_max = 0; O.s1 = O.s2 = nullptr; O.sum = O.diff = O.dt_eps = 0.0;
    k_snapshot_face<<<nblocks(h->planeA[a]), 256, 0, h->stream>>>(O);
    h->launches++;
  }
  // setup_bc_lists + particle_bcs (epoch2d.F90:144-145): classification of the loaded particles
  // happens inside epb_push's kernel; uploaded particles are expected inside the local domain,
  // which the loaders guarantee (helper.F90:658-659 already ran particle_bcs).
  int rc = field_bcs3(h, EPB_EX, false);  // efield_bcs
  if (rc) return rc;
  rc = bfield_final_bcs(h, c.dt / 2.0);   // dt halved (epoch2d.F90:158-162)
  if (rc) return rc;
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

int epb_fields_half(epb_handle *h) {
  if (!h) return EPB_ERR_ARG;
  const double hdt = 0.5 * h->cfg.dt;
  update_e(h, hdt);
  int rc = field_bcs3(h, EPB_EX, false);
  if (rc) return rc;
  update_b(h, hdt);
  rc = field_bcs3(h, EPB_BX, true);
  if (rc) return rc;
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

int epb_fields_final(epb_handle *h) {
  if (!h) return EPB_ERR_ARG;
  const double hdt = 0.5 * h->cfg.dt;
  update_b(h, hdt);
  int rc = bfield_final_bcs(h, h->cfg.dt);
  if (rc) return rc;
  update_e(h, hdt);
  rc = field_bcs3(h, EPB_EX, false);
  if (rc) return rc;
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

int epb_sort(epb_handle *h) {
  if (!h) return EPB_ERR_ARG;
  for (int is = 0; is < (int)h->sp.size(); is++) {
    if (h->sp[is].slots) continue;   // slot columns are always in order
    int rc = h->sp[is].info_valid ? epb_sort_species_emitted(h, is) : epb_sort_species(h, is);
    if (rc) return rc;
  }
  h->pushes_since_sort = 0;
  return EPB_OK;
}

static int current_bcs_species(epb_handle *h, int is);

// thermal walls of this rank for species `is`: re-emission of the particles the push left beyond the outer edge.
// `views`: the particle ranges to scan (slot columns: only the mover buffer holds boundary-touched particles)
static int thermal_apply(epb_handle *h, int is, const SlotView *V, int nv) {
  const epb_config &c = h->cfg;
  SpeciesDev &S = h->sp[is];
  ThermalOp T;
  memset(&T, 0, sizeof T);
  bool any = false;
  for (int d = 0; d < c.ndims; d++) {
    T.th_min[d] = c.is_boundary[2 * d] && S.cfg.bc_particle[2 * d] == EPB_BC_THERMAL;
    T.th_max[d] = c.is_boundary[2 * d + 1] && S.cfg.bc_particle[2 * d + 1] == EPB_BC_THERMAL;
    any = any || T.th_min[d] || T.th_max[d];
  }
  if (!any) return EPB_OK;
  for (int side = 0; side < 2 * c.ndims; side++) {
    const int d = side / 2;
    if (((side & 1) ? T.th_max[d] : T.th_min[d]) && !S.ext_temp[side])
      return epb_fail(h, EPB_ERR_ARG, "species %d: thermal boundary %d without epb_set_boundary_temperature", is, side);
    T.ext_temp[side] = S.ext_temp[side];
  }
  T.nd = c.ndims;
  for (int d = 0; d < 3; d++) {
    T.n[d] = c.n[d];
    T.gmin_local[d] = c.grid_min_local[d];
    T.dx[d] = c.dx[d];
    T.min_outer[d] = c.min_outer[d];
    T.max_outer[d] = c.max_outer[d];
  }
  T.mass = S.cfg.mass;
  h->thermal_calls++;
  T.seed = 0x2545F4914F6CDD1Dull * (unsigned long long)h->thermal_calls + 0x9E3779B97F4A7C15ull * (unsigned long long)(c.rank + 1) + (unsigned long long)is;
  for (int v = 0; v < nv; v++) {
    for (int d = 0; d < 3; d++) { T.x[d] = V[v].a[d]; T.p[d] = V[v].a[3 + d]; }
    T.r = V[v].r;
    const int nb = nblocks((size_t)V[v].r.n, 148 * 8);
    if (c.ndims == 1) k_thermal<1><<<nb, 256, 0, h->stream>>>(T);
    else if (c.ndims == 2) k_thermal<2><<<nb, 256, 0, h->stream>>>(T);
    else k_thermal<3><<<nb, 256, 0, h->stream>>>(T);
    h->launches++;
  }
  return EPB_OK;
}

// ext_temp_<side> of a species (shared_data.F90:255-256; set from the deck's temperature by the host): the wall
// temperature of a thermal particle boundary, (plane, 3) doubles -- the transverse axes in axis order with ghost
// cells (1-ng:n+ng), lower axis fastest, then the three momentum components
int epb_set_boundary_temperature(epb_handle *h, int is, int side, const double *temp) {
  if (!h || is < 0 || is >= (int)h->sp.size() || side < 0 || side >= 2 * h->cfg.ndims || !temp) return EPB_ERR_ARG;
  SpeciesDev &S = h->sp[is];
  size_t plane = 1;
  for (int d = 0; d < h->cfg.ndims; d++) if (d != side / 2) plane *= (size_t)h->sz[d];
  if (!S.ext_temp[side]) EPB_CUDA(h, cudaMalloc(&S.ext_temp[side], 3 * plane * sizeof(double)));
  EPB_CUDA(h, cudaMemcpyAsync(S.ext_temp[side], temp, 3 * plane * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPB_OK;
}

int epb_push(epb_handle *h) {
  if (!h) return EPB_ERR_ARG;
  const epb_config &c = h->cfg;
  if (h->pushes_since_sort >= c.sort_interval) {
    int rc = epb_sort(h);
    if (rc) return rc;
  }
  // jx = jy = jz = 0 (particles.F90:148-150)
  EPB_CUDA(h, cudaMemsetAsync(h->f(EPB_JX), 0, 3 * h->fsize * sizeof(double), h->stream));
  for (int is = 0; is < (int)h->sp.size(); is++) {
    SpeciesDev &S = h->sp[is];
    if (S.cfg.immobile) continue;
    if (S.n == 0 && !S.slots) {  // nothing to push, but neighbours may still send us particles (and, mixed: current)
      if (h->bc_mixed) {
        int rcm = current_bcs_species(h, is);
        if (rcm) return rcm;
      }
      int rc0 = epb_particle_exchange(h, is);
      if (rc0) return rc0;
      continue;
    }
    PushParams P;
    fill_push_params(h, is, P);
    auto launch = c.strict_fp ? epb_launch_push_strict : epb_launch_push_fast;
    if (S.slots) {
      // slot columns (layout 2): the lanes walk their columns and compact them in place; movers, leavers and
      // arrivals go through the mover buffer and are inserted by k_deliver after the exchange.  No sort.
      epb_slots_fill_push(h, is, P);
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (h->time_push) {
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, h->stream);
      }
      launch(P, c.ndims, true, h->stream, &h->launches);
      {  // entries of the mover buffer that still have to be pushed (column full / stencil outside the tile)
        PushParams Pm = P;
        for (int d = 0; d < 3; d++) { Pm.x[d] = P.mx[d]; Pm.p[d] = P.mp[d]; }
        Pm.w = P.mw;
        Pm.gone = P.mflag;
        (c.strict_fp ? epb_launch_push_m_strict : epb_launch_push_m_fast)(Pm, h->stream, &h->launches);
      }
      if (h->time_push) {
        cudaEventRecord(e1, h->stream);
        h->ev_pool.push_back({e0, e1});
      }
      EPB_CUDA(h, cudaGetLastError());
      if (h->bc_mixed) {
        int rcm = current_bcs_species(h, is);
        if (rcm) return rcm;
      }
      int rc = epb_slots_after_push(h, is);
      if (rc) return rc;
      {  // thermal walls: the boundary-touched particles sit in the mover buffer
        SlotView V[2];
        epb_slots_views(h, is, V);
        rc = thermal_apply(h, is, V + 1, 1);
        if (rc) return rc;
      }
      rc = epb_particle_exchange(h, is);
      if (rc) return rc;
      rc = epb_slots_deliver(h, is);
      if (rc) return rc;
      continue;
    }
    static const int no3d = epb_env("EPB_NO_TILED_3D") ? atoi(epb_env("EPB_NO_TILED_3D")) : 0;
    // HC_PUSH builds of the reference: the slot-column kernels carry the rotation as a template flag (handled
    // above); the sorted layouts' tiled kernels hold the Boris gamma only, there every particle takes push_generic<ND, true>
    const bool tiled = !c.hc_push && ((c.ndims == 2) || (c.ndims == 3 && !no3d));
    long long sorted = S.n_sorted < S.n ? S.n_sorted : S.n;
    // layout 1: the last push before a sort also records every particle's place in the next
    // order, so that sort needs neither a key pass nor rank atomics (sort.cu)
    static const int no_emit = epb_env("EPB_NO_EMIT") ? atoi(epb_env("EPB_NO_EMIT")) : 0;
    S.info_valid = false;
    if (tiled && sorted > 0 && h->tg.layout == 1 && P.deposit && !no_emit && h->pushes_since_sort + 1 >= c.sort_interval) {
      const size_t kb = ((size_t)h->tg.nkeys + 1) * sizeof(int);
      EPB_CUDA(h, cudaMemsetAsync(S.key, 0xff, (size_t)S.n * sizeof(int), h->stream));
      EPB_CUDA(h, cudaMemsetAsync(S.stay_cnt, 0, kb, h->stream));
      EPB_CUDA(h, cudaMemsetAsync(S.arr_cnt, 0, kb, h->stream));
      P.emit = 1;
      S.info_valid = true;
    }
    const bool fuse_gather = S.pending_perm && tiled && sorted > 0 && h->tg.layout == 1 && sorted == S.n;
    if (S.pending_perm && !fuse_gather) {
      int rcp = epb_apply_pending_perm(h, is);
      if (rcp) return rcp;
      fill_push_params(h, is, P);  // the buffers were swapped
      if (S.info_valid) P.emit = 1;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->time_push) {
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0, h->stream);
    }
    if (tiled && sorted > 0) {
      P.n_sorted_clip = sorted;
      if (fuse_gather) {
        // first push after an emitted sort: read the old order through perm, write the new one
        P.perm = S.perm;
        for (int d = 0; d < 3; d++) {
          P.xs[d] = S.buf[S.cur][d];
          P.ps[d] = S.buf[S.cur][3 + d];
          P.x[d] = S.buf[S.cur ^ 1][d];
          P.p[d] = S.buf[S.cur ^ 1][3 + d];
        }
        P.ws = S.buf[S.cur][6];
        P.w = S.buf[S.cur ^ 1][6];
        S.cur ^= 1;
        S.pending_perm = false;
      }
      launch(P, c.ndims, true, h->stream, &h->launches);
      P.first = sorted;
    } else {
      P.first = 0;
    }
    P.last = S.n;
    if (P.last > P.first) launch(P, c.ndims, false, h->stream, &h->launches);
    if (h->time_push) {
      cudaEventRecord(e1, h->stream);
      h->ev_pool.push_back({e0, e1});
    }
    EPB_CUDA(h, cudaGetLastError());
    if (h->bc_mixed) {  // current_bcs(species = ispecies), particles.F90:645
      int rcm = current_bcs_species(h, is);
      if (rcm) return rcm;
    }
    {  // thermal walls (part of particle_bcs): scan the species for particles beyond the outer edge
      SlotView V[2];
      const int nv = epb_species_views(h, is, V);
      int rct = thermal_apply(h, is, V, nv);
      if (rct) return rct;
    }
    // particle_bcs (particles.F90:648) for this species.  The outbox is shared by all
    // species, so it is drained before the next species is pushed; the reference runs
    // particle_bcs after the species loop, which is equivalent because a species' push
    // reads no other species' particles.
    int rc = epb_particle_exchange(h, is);
    if (rc) return rc;
  }
  h->pushes_since_sort++;
  return EPB_OK;
}

// current_bcs -> processor_summation_bcs (boundary.F90:783-804): reflection fold, then periodic / neighbour
// sum, per component, with the boundary codes bc_allspecies() currently answers with
static int current_sum_bcs(epb_handle *h) {
  const epb_config &c = h->cfg;
  for (int q = 0; q < 3; q++) {
    for (int d = 0; d < c.ndims; d++)
      for (int side = 0; side < 2; side++) {
        int bd = 2 * d + side;
        if (c.is_boundary[bd] && bc_allspecies(h, bd) == EPB_BC_REFLECT) {
          FoldOp M;
          M.a = h->f(EPB_JX + q);
          M.nd = c.ndims;
          size_t total = 1;
          for (int k = 0; k < 3; k++) {
            M.sz[k] = h->sz[k];
            M.n[k] = c.n[k];
            if (k < c.ndims && k != d) total *= h->sz[k];
          }
          M.d = d;
          M.is_max = side;
          M.flip = (q == d);
          k_jfold<<<nblocks(total), 256, 0, h->stream>>>(M);
          h->launches++;
        }
      }
  }
  return epb_halo_exchange(h, EPB_JX, 3, true);
}

// particle_clear_bcs (boundary.F90:755-779): zero everything outside 1..n of the three current arrays
struct ClearOp { double *a[3]; int nd, n[3], sz[3]; };
__global__ void __launch_bounds__(256) k_clear_ghosts(const __grid_constant__ ClearOp C) {
  const size_t total = (size_t)C.sz[0] * C.sz[1] * C.sz[2];
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t % C.sz[0]) + 1 - NG;
    const int j = C.nd >= 2 ? (int)((t / C.sz[0]) % C.sz[1]) + 1 - NG : 1;
    const int k = C.nd >= 3 ? (int)(t / ((size_t)C.sz[0] * C.sz[1])) + 1 - NG : 1;
    const bool in = i >= 1 && i <= C.n[0] && (C.nd < 2 || (j >= 1 && j <= C.n[1])) && (C.nd < 3 || (k >= 1 && k <= C.n[2]));
    if (!in) { C.a[0][t] = 0.0; C.a[1][t] = 0.0; C.a[2][t] = 0.0; }
  }
}
// current_bcs(species) of a c_bc_mixed run, called after that species was pushed
static int current_bcs_species(epb_handle *h, int is) {
  h->bc_species = is;
  int rc = current_sum_bcs(h);
  h->bc_species = -1;
  if (rc) return rc;
  ClearOp C;
  for (int q = 0; q < 3; q++) C.a[q] = h->f(EPB_JX + q);
  C.nd = h->cfg.ndims;
  for (int k = 0; k < 3; k++) { C.n[k] = h->cfg.n[k]; C.sz[k] = h->sz[k]; }
  k_clear_ghosts<<<nblocks(h->fsize, 148 * 16), 256, 0, h->stream>>>(C);
  h->launches++;
  return EPB_OK;
}

int epb_current_finish(epb_handle *h) {
  if (!h) return EPB_ERR_ARG;
  const epb_config &c = h->cfg;
  int rc = EPB_OK;
  if (!h->bc_mixed) {  // with c_bc_mixed the sums were done per species inside epb_push
    rc = current_sum_bcs(h);
    if (rc) return rc;
  }
  rc = epb_halo_exchange(h, EPB_JX, 3, false);  // field_bc(jx|jy|jz, jng)
  if (rc) return rc;
  if (c.smooth_its + c.smooth_comp_its > 0) {  // smooth_current (current_smooth.F90:50-141)
    const int WK0 = 9;
    EPB_CUDA(h, cudaMemcpyAsync(h->f(WK0), h->f(EPB_JX), 3 * h->fsize * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    int strides[4] = {1, 0, 0, 0}, ns = 0;
    for (int q = 0; q < 4; q++) {
      const int v = (c.smooth_strides >> (4 * q)) & 15;
      if (v) strides[ns++] = v;
    }
    if (ns == 0) ns = 1;
    SmoothOp S;
    for (int q = 0; q < 3; q++) { S.wk[q] = h->f(WK0 + q); S.a[q] = h->f(EPB_JX + q); }
    S.nd = c.ndims;
    for (int k = 0; k < 3; k++) { S.n[k] = c.n[k]; S.sz[k] = h->sz[k]; }
    double alpha = 0.5;
    S.beta = c.ndims == 1 ? (1.0 - alpha) * 0.5 : c.ndims == 2 ? (1.0 - alpha) * 0.25 : (1.0 - alpha) / 6.0;
    const size_t total = (size_t)c.n[0] * c.n[1] * c.n[2];
    for (int it = 1; it <= c.smooth_its + c.smooth_comp_its; it++) {
      for (int is = 0; is < ns; is++) {
        rc = epb_halo_exchange(h, WK0, 3, false);
        if (rc) return rc;
        S.cs = strides[is];
        S.alpha = alpha;
        k_smooth<<<nblocks(total, 148 * 32), 256, 0, h->stream>>>(S);
        k_smooth_copyback<<<nblocks(total, 148 * 32), 256, 0, h->stream>>>(S);
        h->launches += 2;
      }
      if (it > c.smooth_its) alpha = (double)c.smooth_its * 0.5 + 1.0;  // as in the reference: after the pass
    }
  }
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

int epb_field_energy(epb_handle *h, double out[2]) {
  if (!h || !out) return EPB_ERR_ARG;
  const epb_config &c = h->cfg;
  double *d = (double *)(h->d_scratch ? h->d_scratch + 256 : nullptr);
  bool tmp = false;
  if (!d) { EPB_CUDA(h, cudaMalloc(&d, 2 * sizeof(double))); tmp = true; }
  EPB_CUDA(h, cudaMemsetAsync(d, 0, 2 * sizeof(double), h->stream));
  EnergyOp E;
  for (int q = 0; q < 6; q++) E.f[q] = h->f(q);
  E.nd = c.ndims;
  for (int q = 0; q < 3; q++) { E.n[q] = c.n[q]; E.sz[q] = h->sz[q]; }
  E.out = d;
  size_t total = (size_t)c.n[0] * c.n[1] * c.n[2];
  k_field_energy<<<nblocks(total, 148 * 8), 256, 0, h->stream>>>(E);
  h->launches++;
  double v[2];
  EPB_CUDA(h, cudaMemcpyAsync(v, d, sizeof v, cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (tmp) cudaFree(d);
  double dv = 1.0;
  for (int q = 0; q < c.ndims; q++) dv *= c.dx[q];
  const double mu0 = 4.e-7 * 3.141592653589793238462643383279503;
  out[0] = 0.5 * EPB_EPS0 * v[0] * dv;
  out[1] = 0.5 / mu0 * v[1] * dv;
  return EPB_OK;
}

// The scalars EPOCH's host looks at after every step -- update_particle_count's global count of every species
// (partlist.F90:984-1003) and calc_total_energy_sum's field energies (io/calc_df.F90:1321-1417, summed over the
// ranks like its MPI_ALLREDUCE) -- evaluated in stream order and copied to page-locked host memory WITHOUT stopping
// the host, so that it can enqueue the next step while this one runs.  host[0] = 0.5 eps0 sum E^2 dV, host[1] =
// 0.5/mu0 sum B^2 dV, host[2 + is] = global count of species is (exact: an integer below 2^53), host[2 + n_species]
// = the device error word (non-zero: a capacity overflow lost particles).  Returns a ticket for epb_wait_scalars.
namespace {
struct ScalOp {
  double *out;
  double ce, cb;
  int nsp;
  const long long *cnt2;          // [2 * is]: columns, inbox
  const int *mcount[EPB_SCAL_MAXSP];
  int mcap[EPB_SCAL_MAXSP], slots[EPB_SCAL_MAXSP];
  long long fixed[EPB_SCAL_MAXSP];
  const int *err;
};
__global__ void k_scalars(const __grid_constant__ ScalOp O) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  O.out[0] *= O.ce;
  O.out[1] *= O.cb;
  for (int is = 0; is < O.nsp; is++) {
    long long v = O.fixed[is];
    if (O.slots[is]) {
      const int w = *O.mcount[is];
      v = O.cnt2[2 * is] + O.cnt2[2 * is + 1] + (w < O.mcap[is] ? w : O.mcap[is]);
    }
    O.out[2 + is] = (double)v;
  }
  O.out[2 + O.nsp] = O.err ? (double)*O.err : 0.0;
}
}  // namespace

int epb_step_scalars_async(epb_handle *h, double *host, int64_t *ticket) {
  if (!h || !host) return EPB_ERR_ARG;
  const epb_config &c = h->cfg;
  const int nsp = (int)h->sp.size();
  if (nsp > EPB_SCAL_MAXSP) return epb_fail(h, EPB_ERR_UNSUPPORTED, "epb_step_scalars_async: more than %d species", EPB_SCAL_MAXSP);
  if (!h->scal_dev) {
    EPB_CUDA(h, cudaMalloc(&h->scal_dev, (size_t)(2 * (3 + EPB_SCAL_MAXSP)) * sizeof(double)));
    for (int q = 0; q < 4; q++) EPB_CUDA(h, cudaEventCreateWithFlags(&h->scal_ev[q], cudaEventDisableTiming));
  }
  double *d = h->scal_dev;
  const int len = 3 + nsp;
  EPB_CUDA(h, cudaMemsetAsync(d, 0, 2 * (size_t)(3 + EPB_SCAL_MAXSP) * sizeof(double), h->stream));
  EnergyOp E;
  for (int q = 0; q < 6; q++) E.f[q] = h->f(q);
  E.nd = c.ndims;
  for (int q = 0; q < 3; q++) { E.n[q] = c.n[q]; E.sz[q] = h->sz[q]; }
  E.out = d;
  const size_t total = (size_t)c.n[0] * c.n[1] * c.n[2];
  k_field_energy<<<nblocks(total, 148 * 8), 256, 0, h->stream>>>(E);
  ScalOp O;
  memset(&O, 0, sizeof O);
  double dv = 1.0;
  for (int q = 0; q < c.ndims; q++) dv *= c.dx[q];
  const double mu0 = 4.e-7 * 3.141592653589793238462643383279503;
  O.out = d;
  O.ce = 0.5 * EPB_EPS0 * dv;
  O.cb = 0.5 / mu0 * dv;
  O.nsp = nsp;
  long long *cnt2 = (long long *)(h->d_scratch + 512);
  O.cnt2 = cnt2;
  O.err = h->d_err;
  for (int is = 0; is < nsp; is++) {
    SpeciesDev &S = h->sp[is];
    O.slots[is] = S.slots ? 1 : 0;
    O.fixed[is] = S.n;
    if (S.slots) {
      int rc = epb_slots_count_enqueue(h, is, cnt2 + 2 * is);
      if (rc) return rc;
      O.mcount[is] = S.mcount + S.mcur;
      O.mcap[is] = S.mcap;
    }
  }
  k_scalars<<<1, 32, 0, h->stream>>>(O);
  h->launches += 2;
  EPB_CUDA(h, cudaGetLastError());
  double *res = d;
  if (c.nranks > 1 && h->nccl) {
    int rc = epb_allreduce_sum_f64(h, d, d + (3 + EPB_SCAL_MAXSP), len);
    if (rc) return rc;
    res = d + (3 + EPB_SCAL_MAXSP);
  }
  EPB_CUDA(h, cudaMemcpyAsync(host, res, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  const long long t = h->scal_ticket++;
  EPB_CUDA(h, cudaEventRecord(h->scal_ev[t % 4], h->stream));
  if (ticket) *ticket = t;
  return EPB_OK;
}

int epb_wait_scalars(epb_handle *h, int64_t ticket) {
  if (!h || ticket < 0 || ticket >= h->scal_ticket) return EPB_ERR_ARG;
  if (h->scal_ticket - ticket > 4) return EPB_OK;   // its event was recorded again since: that request completed long ago
  EPB_CUDA(h, cudaEventSynchronize(h->scal_ev[ticket % 4]));
  return EPB_OK;
}

int epb_kinetic_energy(epb_handle *h, int is, double *out) {
  if (!h || is < 0 || is >= (int)h->sp.size() || !out) return EPB_ERR_ARG;
  SpeciesDev &S = h->sp[is];
  double *d = (double *)(h->d_scratch + 256);
  EPB_CUDA(h, cudaMemsetAsync(d, 0, sizeof(double), h->stream));
  const double mc = EPB_C * S.cfg.mass;
  SlotView V[2];
  const int nv = epb_species_views(h, is, V);
  for (int v = 0; v < nv; v++) {
    k_kinetic_energy<<<nblocks((size_t)V[v].r.n, 148 * 8), 256, 0, h->stream>>>(V[v].a[3], V[v].a[4], V[v].a[5], V[v].a[6], V[v].r,
                                                                          mc, mc * EPB_C, d);
    h->launches++;
  }
  EPB_CUDA(h, cudaMemcpyAsync(out, d, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return EPB_OK;
}

// calc_boundary (io/calc_df.F90:24-31) = processor_summation_bcs without flip_direction on the work array:
// reflecting fold (no sign change), then the periodic / neighbour sum (boundary.F90:783-804)
static int moment_sum_bcs(epb_handle *h, int f) {
  const epb_config &c = h->cfg;
  for (int d = 0; d < c.ndims; d++)
    for (int side = 0; side < 2; side++) {
      const int bd = 2 * d + side;
      if (c.is_boundary[bd] && bc_allspecies(h, bd) == EPB_BC_REFLECT) {
        FoldOp M;
        M.a = h->f(f);
        M.nd = c.ndims;
        size_t total = 1;
        for (int k = 0; k < 3; k++) {
          M.sz[k] = h->sz[k];
          M.n[k] = c.n[k];
          if (k < c.ndims && k != d) total *= h->sz[k];
        }
        M.d = d;
        M.is_max = side;
        M.flip = 0;
        k_jfold<<<nblocks(total), 256, 0, h->stream>>>(M);
        h->launches++;
      }
    }
  return epb_halo_exchange(h, f, 1, true);
}

// field_zero_gradient(array, c_stagger_centre, bd) on every boundary (boundary.F90:416-469)
static int moment_zero_gradient(epb_handle *h, int f) {
  const epb_config &c = h->cfg;
  for (int bd = 0; bd < 2 * c.ndims; bd++) {
    if (c.bc_field[bd] == EPB_BC_PERIODIC || !c.is_boundary[bd]) continue;
    MirrorOp M;
    M.nd = c.ndims;
    M.d = bd / 2;
    M.is_max = bd & 1;
    M.sign[0] = M.sign[1] = M.sign[2] = 1.0;
    size_t total = 1;
    for (int d = 0; d < 3; d++) {
      M.sz[d] = h->sz[d];
      M.n[d] = c.n[d];
      if (d < c.ndims && d != M.d) total *= h->sz[d];
    }
    for (int q = 0; q < 3; q++) { M.f[q] = h->f(f); M.stag[q] = 0; }  // one array: the copy is idempotent
    k_mirror<<<nblocks(total), 256, 0, h->stream>>>(M);
    h->launches++;
  }
  return EPB_OK;
}

// calc_boundary(array, ispecies): only under c_bc_mixed, followed by particle_clear_bcs (boundary.F90:749, :790-792)
static int moment_bcs_species(epb_handle *h, int f, int is) {
  if (!h->bc_mixed) return EPB_OK;
  h->bc_species = is;
  int rc = moment_sum_bcs(h, f);
  h->bc_species = -1;
  if (rc) return rc;
  ClearOp C;
  for (int q = 0; q < 3; q++) C.a[q] = h->f(f);
  C.nd = h->cfg.ndims;
  for (int k = 0; k < 3; k++) { C.n[k] = h->cfg.n[k]; C.sz[k] = h->sz[k]; }
  k_clear_ghosts<<<nblocks(h->fsize, 148 * 16), 256, 0, h->stream>>>(C);
  h->launches++;
  return EPB_OK;
}
// calc_boundary(array): only without c_bc_mixed (boundary.F90:794-796)
static int moment_bcs_all(epb_handle *h, int f) { return h->bc_mixed ? EPB_OK : moment_sum_bcs(h, f); }

static void moment2_fill(epb_handle *h, int is, const SlotView &V, Moment2Op &M) {
  const epb_config &c = h->cfg;
  SpeciesDev &S = h->sp[is];
  for (int d = 0; d < 3; d++) {
    M.x[d] = V.a[d];
    M.p[d] = V.a[3 + d];
    M.sz[d] = h->sz[d];
    M.gmin[d] = c.grid_min_local[d];
    M.dx[d] = c.dx[d];
  }
  M.w = V.a[6];
  M.r = V.r;
  M.nd = c.ndims;
  M.part_mc = EPB_C * S.cfg.mass;
  M.sqrt_part_m = sqrt(S.cfg.mass);
}

// One pass of calc_ekbar / calc_temperature over a range of particles: the slot-column kernel for the arena of the 2D
// default layout, the per-particle scatter for everything else (EPB_DEBUG=1 EPB_MOMENT_GENERIC=1: always the latter).
static void launch_moment2(epb_handle *h, const Moment2Op &M) {
  static const int generic = epb_env("EPB_MOMENT_GENERIC") ? atoi(epb_env("EPB_MOMENT_GENERIC")) : 0;
  if (!generic && h->tg.layout == 2 && h->tg.T[0] == 16 && h->tg.cpt % 32 == 0 && M.nd == 2 && M.r.cnt && M.r.K && M.mode >= 3 && M.mode <= 5) {
    const int ngroups = h->tg.nkeys / 32;
    const int blocks = std::min((ngroups + MS_WARPS - 1) / MS_WARPS, 148 * 16);
    if (M.mode == 3) k_moment2_slots<3><<<blocks, MS_WARPS * 32, 0, h->stream>>>(M, h->tg);
    else if (M.mode == 4) k_moment2_slots<4><<<blocks, MS_WARPS * 32, 0, h->stream>>>(M, h->tg);
    else k_moment2_slots<5><<<blocks, MS_WARPS * 32, 0, h->stream>>>(M, h->tg);
  } else {
    k_moment2<<<nblocks((size_t)M.r.n, 148 * 32), 256, 0, h->stream>>>(M);
  }
  h->launches++;
}

// calc_ekbar (io/calc_df.F90:116-221, sub 0), calc_ekflux (:415-557, sub 1..6), calc_average_momentum (:1239-1317,
// sub 7..9), calc_average_weight (:811-873, sub 10: nearest cell, no ghost-cell sums or fill): result in work array 9
static int calc_ratio_dev(epb_handle *h, int ispecies, int sub) {
  const epb_config &c = h->cfg;
  const int A = 9, WT = 10;
  const bool avg_weight = sub == 10;
  double flux_fac = 0.0;
  if (sub >= 1 && sub <= 6) {
    const int a = (sub - 1) / 2, nd = c.ndims;
    const double cc = EPB_C, dx = c.dx[0], dy = c.dx[1], dz = c.dx[2];
    if (nd == 1) flux_fac = a == 0 ? cc : cc * dx;
    else if (nd == 2) flux_fac = a == 0 ? cc * dy : a == 1 ? cc * dx : cc * dx * dy;
    else flux_fac = a == 0 ? cc * dy * dz : a == 1 ? cc * dx * dz : cc * dx * dy;
  }
  EPB_CUDA(h, cudaMemsetAsync(h->f(A), 0, 2 * h->fsize * sizeof(double), h->stream));
  const bool spec_sum = ispecies < 0;
  for (int is = spec_sum ? 0 : ispecies; is < (spec_sum ? (int)h->sp.size() : ispecies + 1); is++) {
    SpeciesDev &S = h->sp[is];
    if (spec_sum && S.cfg.zero_current) continue;
    SlotView V[2];
    const int nv = epb_species_views(h, is, V);
    for (int v = 0; v < nv; v++) {
      Moment2Op M;
      moment2_fill(h, is, V[v], M);
      M.mode = avg_weight ? 6 : 3;
      M.dir = -1;
      M.sub = sub;
      M.flux_fac = flux_fac;
      M.a0 = h->f(A);
      M.a1 = h->f(WT);
      for (int q = 0; q < 3; q++) M.mean[q] = nullptr;
      launch_moment2(h, M);
    }
    if (avg_weight) continue;
    int rc = moment_bcs_species(h, A, is);
    if (!rc) rc = moment_bcs_species(h, WT, is);
    if (rc) return rc;
  }
  int rc = EPB_OK;
  if (!avg_weight) {
    rc = moment_bcs_all(h, A);
    if (!rc) rc = moment_bcs_all(h, WT);
  }
  if (rc) return rc;
  MomentPostOp P;
  P.op = 0; P.n = h->fsize; P.a = h->f(A); P.b = h->f(WT);
  for (int q = 0; q < 3; q++) P.m[q] = nullptr;
  P.k1 = 2.2250738585072014e-308;  // c_tiny = TINY(1.0_num), constants.F90:29
  P.k2 = 0.0;
  k_moment_post<<<nblocks(h->fsize, 148 * 16), 256, 0, h->stream>>>(P);
  h->launches++;
  (void)c;
  return EPB_OK;
}

// calc_temperature (io/calc_df.F90:877-1128): sigma in work array 9, means 10..12, part_count 13
static int calc_temperature_dev(epb_handle *h, int ispecies, int dir) {
  const int SIG = 9, MEAN0 = 10, CNT = 13;
  EPB_CUDA(h, cudaMemsetAsync(h->f(SIG), 0, 5 * h->fsize * sizeof(double), h->stream));
  const bool spec_sum = ispecies < 0;
  const int s0 = spec_sum ? 0 : ispecies, s1 = spec_sum ? (int)h->sp.size() : ispecies + 1;
  for (int pass = 0; pass < 2; pass++) {
    for (int is = s0; is < s1; is++) {
      SpeciesDev &S = h->sp[is];
      if (spec_sum && S.cfg.zero_current) continue;
      SlotView V[2];
      const int nv = epb_species_views(h, is, V);
      for (int v = 0; v < nv; v++) {
        Moment2Op M;
        moment2_fill(h, is, V[v], M);
        M.mode = pass == 0 ? 4 : 5;
        M.dir = dir;
        M.sub = 0;
        M.flux_fac = 0.0;
        M.a0 = h->f(SIG);
        M.a1 = h->f(CNT);
        for (int q = 0; q < 3; q++) M.mean[q] = h->f(MEAN0 + q);
        launch_moment2(h, M);
      }
      int rc = EPB_OK;
      if (pass == 0) {
        for (int q = 0; q < 3 && !rc; q++)
          if (dir < 0 || dir == q) rc = moment_bcs_species(h, MEAN0 + q, is);
      } else {
        rc = moment_bcs_species(h, SIG, is);
      }
      if (!rc) rc = moment_bcs_species(h, CNT, is);
      if (rc) return rc;
    }
    int rc = EPB_OK;
    if (pass == 0) {
      for (int q = 0; q < 3 && !rc; q++)
        if (dir < 0 || dir == q) rc = moment_bcs_all(h, MEAN0 + q);
    } else {
      rc = moment_bcs_all(h, SIG);
    }
    if (!rc) rc = moment_bcs_all(h, CNT);
    if (rc) return rc;
    MomentPostOp P;
    P.n = h->fsize; P.a = h->f(SIG); P.b = h->f(CNT);
    for (int q = 0; q < 3; q++) P.m[q] = h->f(MEAN0 + q);
    if (pass == 0) {
      P.op = 1; P.k1 = 0.0; P.k2 = 0.0;
    } else {
      P.op = 2; P.k1 = EPB_KB; P.k2 = dir < 0 ? 3.0 : 1.0;
    }
    k_moment_post<<<nblocks(h->fsize, 148 * 16), 256, 0, h->stream>>>(P);
    h->launches++;
    if (pass == 0) {
      // restore the ghost cells of the means (field_bc), then part_count = 0
      for (int q = 0; q < 3; q++)
        if (dir < 0 || dir == q) {
          rc = epb_halo_exchange(h, MEAN0 + q, 1, false);
          if (rc) return rc;
        }
      EPB_CUDA(h, cudaMemsetAsync(h->f(CNT), 0, h->fsize * sizeof(double), h->stream));
    }
  }
  return EPB_OK;
}

}  // extern "C"

// calc_coll_ekbar (what = 0; collisions.F90:1487-1575, J) / calc_coll_temperature_ev (what = 1; :1367-1483, left in K)
// of one species into a device array: the same grid quantities as calc_ekbar / calc_temperature, whose device
// versions are reused (collide.cu turns the temperature into eV where it uses it)
int epb_coll_moment_dev(epb_handle *h, int what, int ispecies, double *dst) {
  int rc = what == 0 ? calc_ratio_dev(h, ispecies, 0) : calc_temperature_dev(h, ispecies, -1);
  if (rc) return rc;
  if (what == 0) {
    rc = moment_zero_gradient(h, 9);
    if (rc) return rc;
  }
  EPB_CUDA(h, cudaMemcpyAsync(dst, h->f(9), h->fsize * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return EPB_OK;
}

extern "C" {

int epb_calc_moment(epb_handle *h, int kind, int ispecies, double *host) {
  if (!h || !host || kind < 0 || kind > EPB_MOMENT_POYNT_FLUX_Z || ispecies >= (int)h->sp.size()) return EPB_ERR_ARG;
  if (kind >= EPB_MOMENT_POYNT_FLUX_X) {  // field-only: ghost cells of the result are zero
    EPB_CUDA(h, cudaMemsetAsync(h->f(9), 0, h->fsize * sizeof(double), h->stream));
    PoyntOp P;
    for (int q = 0; q < 6; q++) P.f[q] = h->f(q);
    P.nd = h->cfg.ndims;
    for (int q = 0; q < 3; q++) { P.n[q] = h->cfg.n[q]; P.sz[q] = h->sz[q]; }
    P.dir = kind - EPB_MOMENT_POYNT_FLUX_X;
    P.out = h->f(9);
    const size_t total = (size_t)P.n[0] * P.n[1] * P.n[2];
    k_poynt_flux<<<nblocks(total, 148 * 16), 256, 0, h->stream>>>(P);
    h->launches++;
    EPB_CUDA(h, cudaMemcpyAsync(host, h->f(9), h->fsize * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPB_CUDA(h, cudaStreamSynchronize(h->stream));
    EPB_CUDA(h, cudaGetLastError());
    return EPB_OK;
  }
  const bool density_like = kind <= EPB_MOMENT_MASS_DENSITY || (kind >= EPB_MOMENT_JX && kind <= EPB_MOMENT_JZ);
  if (!density_like) {
    int rc;
    bool fill = true;  // field_zero_gradient(data_array, c_stagger_centre, bd): not for temperature / average weight
    if (kind == EPB_MOMENT_EKBAR) rc = calc_ratio_dev(h, ispecies, 0);
    else if (kind <= EPB_MOMENT_TEMPERATURE_Z) { rc = calc_temperature_dev(h, ispecies, kind - 5); fill = false; }
    else if (kind <= EPB_MOMENT_AVERAGE_PZ) rc = calc_ratio_dev(h, ispecies, kind - 7);  // ekflux 1..6, momentum 7..9
    else { rc = calc_ratio_dev(h, ispecies, 10); fill = false; }
    if (rc) return rc;
    if (fill) {
      int rcz = moment_zero_gradient(h, 9);
      if (rcz) return rcz;
    }
    EPB_CUDA(h, cudaMemcpyAsync(host, h->f(9), h->fsize * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPB_CUDA(h, cudaStreamSynchronize(h->stream));
    EPB_CUDA(h, cudaGetLastError());
    return EPB_OK;
  }
  const epb_config &c = h->cfg;
  const int WK = 9;
  double *wk = h->f(WK);
  EPB_CUDA(h, cudaMemsetAsync(wk, 0, h->fsize * sizeof(double), h->stream));
  const bool current = kind >= EPB_MOMENT_JX;
  double idx;
  if (kind == EPB_MOMENT_NUMBER_DENSITY) {  // vol = dx * dy; idx = 1 / vol
    double vol = c.dx[0];
    for (int d = 1; d < c.ndims; d++) vol = vol * c.dx[d];
    idx = 1.0 / vol;
  } else {                                  // idx = 1 / dx / dy
    idx = 1.0 / c.dx[0];
    for (int d = 1; d < c.ndims; d++) idx = idx / c.dx[d];
    if (current) idx = EPB_C * idx;       // calc_per_species_current: idx = c * idx
  }
  const bool spec_sum = ispecies < 0;
  for (int is = spec_sum ? 0 : ispecies; is < (spec_sum ? (int)h->sp.size() : ispecies + 1); is++) {
    SpeciesDev &S = h->sp[is];
    if (spec_sum && S.cfg.zero_current) continue;  // tracers are left out of a species sum
    SlotView V[2];
    const int nv = epb_species_views(h, is, V);
    for (int v = 0; v < nv; v++) {
      MomentOp M;
      for (int d = 0; d < 3; d++) {
        M.x[d] = V[v].a[d];
        M.sz[d] = h->sz[d];
        M.gmin[d] = c.grid_min_local[d];
        M.dx[d] = c.dx[d];
      }
      M.w = V[v].a[6];
      M.r = V[v].r;
      M.nd = c.ndims;
      M.use_scale = current ? 2 : kind != EPB_MOMENT_NUMBER_DENSITY;
      M.scale = (kind == EPB_MOMENT_CHARGE_DENSITY || current) ? S.cfg.charge : S.cfg.mass;
      M.dir = current ? kind - EPB_MOMENT_JX : 0;
      M.part_mc = EPB_C * S.cfg.mass;
      for (int d = 0; d < 3; d++) M.p[d] = V[v].a[3 + d];
      M.out = wk;
      k_moment<<<nblocks((size_t)V[v].r.n, 148 * 32), 256, 0, h->stream>>>(M);
      h->launches++;
    }
    if (h->bc_mixed) {  // calc_boundary(data_array, ispecies) + particle_clear_bcs
      h->bc_species = is;
      int rc = moment_sum_bcs(h, WK);
      h->bc_species = -1;
      if (rc) return rc;
      ClearOp C;
      for (int q = 0; q < 3; q++) C.a[q] = wk;
      C.nd = c.ndims;
      for (int k = 0; k < 3; k++) { C.n[k] = c.n[k]; C.sz[k] = h->sz[k]; }
      k_clear_ghosts<<<nblocks(h->fsize, 148 * 16), 256, 0, h->stream>>>(C);
      h->launches++;
    }
  }
  if (!h->bc_mixed) {  // calc_boundary(data_array)
    int rc = moment_sum_bcs(h, WK);
    if (rc) return rc;
  }
  k_scale<<<nblocks(h->fsize, 148 * 16), 256, 0, h->stream>>>(wk, h->fsize, idx);
  h->launches++;
  {
    int rcz = moment_zero_gradient(h, WK);
    if (rcz) return rcz;
  }
  EPB_CUDA(h, cudaMemcpyAsync(host, wk, h->fsize * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

int epb_push_kernel_ms(epb_handle *h, double *avg_ms, int64_t *launches, int reset) {
  if (!h) return EPB_ERR_ARG;
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (auto &e : h->ev_pool) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) { h->push_ms_sum += ms; h->push_ms_n++; }
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  h->ev_pool.clear();
  if (avg_ms) *avg_ms = h->push_ms_n ? h->push_ms_sum / h->push_ms_n : 0.0;
  if (launches) *launches = h->push_ms_n;
  if (reset == 1) { h->push_ms_sum = 0; h->push_ms_n = 0; h->time_push = 1; }
  if (reset == 2) { h->push_ms_sum = 0; h->push_ms_n = 0; h->time_push = 0; }
  return EPB_OK;
}

}  // extern "C"
