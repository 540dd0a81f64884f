// collide.cu — binary Coulomb collisions on the cell-resident particle layout (BASELINE config 5).
//
// Replaces particle_collisions (epoch2d/src/physics_packages/collisions.F90:86-214) and what it needs around it:
// reorder_particles_to_grid / reattach_particles_to_mainlist (housekeeping/split_particle.F90:29-106) fall away,
// because the slot columns of layout 2 ARE the per-cell lists (secondary_list(ix,iy)): one warp takes a group of
// 32 cells, one lane a cell, and row r of the group is a coalesced access exactly as in the push.  Per cell the
// reference's sequence is followed: number density of the cell (calc_coll_number_density :1320-1363), the pair
// weights (np, factor), then the pairs in list order -- (1,2), (3,4), ... with the list closed into a ring, so an
// odd cell's last particle meets the head again (:262-276, :296-431) -- with either scattering operator:
//   * Nanbu / Perez cumulative scattering (intra_collisions_np :446-646, inter_collisions_np :900-1115), the
//     reference's default (use_nanbu = T, shared_data.F90:575);
//   * Sentoku-Kemp (intra_collisions_sk :218-442, inter_collisions_sk :650-893) incl. the weighted-particle
//     correction (:1146-1185).
// Compiled with -fmad=false: the pair arithmetic is the reference's, operation for operation (the oracle holds the
// same restatement on the CPU; tests/test_collisions.py compares the two pair by pair on identical random numbers).
//
// Random numbers: the reference draws from one global KISS stream in cell-and-list order, which no parallel code
// can reproduce; here every pair gets its own counter-based stream (splitmix64 of seed, call number, species pair,
// cell key, pair index), so a run is reproducible whatever the launch geometry.  Parity with the reference is
// therefore statistical for whole steps (conservation per cell, relaxation rates) and exact per pair.
//
// The reference shuffles the list of the outer species in every cell on every collision step (reorder_particles_to_
// grid clears is_shuffled; shuffle_particle_list_random :1224-1284, Durstenfeld).  Here the lane draws the same kind
// of permutation of its column's rows (in local memory, nothing is moved) and walks the pairs through it.
// Coulomb logarithm: fixed, or calc_coulomb_log (:1288-1316) from the device moments (ekbar of species 1,
// temperature of species 2: the same grid quantities calc_coll_ekbar / calc_coll_temperature_ev build).
#include <cmath>
#include <cstring>

#include "epb_internal.h"

int epb_coll_moment_dev(epb_handle *h, int what, int ispecies, double *dst);   // epb_api.cu

namespace {

constexpr int NGC = EPB_NG;
constexpr double C_ = EPB_C;
constexpr double CC = EPB_C * EPB_C;
constexpr double PI = 3.141592653589793238462643383279503;
constexpr double Q0 = 1.602176565e-19;
constexpr double M0 = 9.10938291e-31;
constexpr double MC0 = 2.73092429345209278e-22;
constexpr double H_BAR = 1.054571725336289397963133257349698e-34;
constexpr double EPS = 2.220446049250313e-16;            // EPSILON(1.0_num)
constexpr double C_TINY = 2.2250738585072014e-308;        // TINY(1.0_num)
constexpr double C_LARGEST = 1.7976931348623157e308;      // HUGE(1.0_num)

struct Rng {   // counter-based: one independent stream per pair
  unsigned long long s;
  __host__ __device__ double next() {
    unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
  }
};
// test hook: the random numbers of a pair come from an array instead (at most 4 per pair)
struct RanSrc {
  Rng g;
  const double *fixed;
  int k;
  __device__ double next() { return fixed ? fixed[k++] : g.next(); }
};

__device__ __forceinline__ double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// new_coords (collisions.F90:1189-1220)
__device__ void new_coords(const double *v, double *c1, double *c2, double *c3) {
  const double vmag = sqrt(dot3(v, v));
  const double vtrans = sqrt(v[1] * v[1] + v[2] * v[2]);
  if (vtrans > C_TINY) {
    for (int d = 0; d < 3; d++) c1[d] = v[d] / vmag;
    c2[0] = 0.0 / vtrans; c2[1] = v[2] / vtrans; c2[2] = -v[1] / vtrans;
    const double den = vmag * vtrans;
    c3[0] = (vtrans * vtrans) / den; c3[1] = -(v[0] * v[1]) / den; c3[2] = -(v[0] * v[2]) / den;
  } else {
    c1[0] = 1.0; c1[1] = 0.0; c1[2] = 0.0;
    c2[0] = 0.0; c2[1] = 1.0; c2[2] = 0.0;
    c3[0] = 0.0; c3[1] = 0.0; c3[2] = 1.0;
  }
}

// coll_freq (collisions.F90:1119-1142)
__device__ double coll_freq(double vrabs, double log_lambda, double m1, double m2, double q1, double q2, double jdens) {
  const double fac = 4.0 * PI * (EPB_EPS0 * EPB_EPS0);
  const double mu = (m1 * m2) / (m1 + m2);
  if (!(vrabs > 0.0)) return 0.0;
  const double q12 = q1 * q2;
  const double numerator = q12 * q12 * jdens * log_lambda;
  const double denominator = fac * (mu * mu) * (vrabs * vrabs * vrabs);
  int en, ed;
  frexp(numerator, &en);
  frexp(denominator, &ed);
  if (denominator <= 0.0 || en - ed >= 1024) return 0.0;
  return numerator / denominator;
}

// weighted_particles_correction (collisions.F90:1146-1185)
__device__ void weighted_correction(double wtr, const double *p, double *p_scat, double en, double en_scat, double mass,
                                    RanSrc &R) {
  double p_after[3];
  const double en_after = (1.0 - wtr) * en + wtr * en_scat;
  for (int d = 0; d < 3; d++) p_after[d] = (1.0 - wtr) * p[d] + wtr * p_scat[d];
  const double p_mag = sqrt(dot3(p_after, p_after));
  const double gamma_en = en_after / (mass * CC);
  const double pm = p_mag / mass / C_;
  const double gamma_p = sqrt(1.0 + pm * pm);
  if (gamma_p < gamma_en) {
    const double delta_p = mass * C_ * sqrt(gamma_en * gamma_en - gamma_p * gamma_p);
    double c1[3], c2[3], c3[3];
    new_coords(p_after, c1, c2, c3);
    const double phi = 2.0 * PI * R.next();
    const double cp = cos(phi), sp = sin(phi);
    for (int d = 0; d < 3; d++) p_scat[d] = p_after[d] + delta_p * (c2[d] * cp + c3[d] * sp);
  }
}

struct PairEnv {
  double m1, m2, q1, q2;
  double dens;         // SK: jdens of coll_freq (intra: dens; inter: MIN(idens, jdens))
  double log_lambda;
  double factor, np, dt_coll;   // SK: nu * factor * np * dt_coll, left to right as in the reference
  double s_fac;        // NP: cell_fac * log_lambda / (4 pi eps0^2 c^4)
  double s_fac_prime;  // NP: intra cell_fac * pi_fac / dens^(2/3); inter cell_fac * pi_fac
  double sp_den;       // NP: intra MAX(m1, m2); inter MAX(m1 idens^(2/3), m2 jdens^(2/3))
  int inter;           // NP: inter-species pairs draw a third number and update by the weight ratio
};

// One Sentoku-Kemp pair (collisions.F90:296-431, :750-880).  Returns false if the pair was skipped (CYCLE).
__device__ bool pair_sk(const PairEnv &E, double *p1, double *p2, double w1, double w2, RanSrc &R) {
  const double m1 = E.m1, m2 = E.m2;
  double p1n[3], p2n[3], vc[3];
  for (int d = 0; d < 3; d++) { p1n[d] = p1[d] / MC0; p2n[d] = p2[d] / MC0; }
  if (dot3(p1n, p1n) < EPS && dot3(p2n, p2n) < EPS) return false;
  for (int d = 0; d < 3; d++) vc[d] = p1n[d] - p2n[d];
  if (dot3(vc, vc) < EPS) return false;
  const double e1 = C_ * sqrt(dot3(p1, p1) + (m1 * C_) * (m1 * C_));
  const double e2 = C_ * sqrt(dot3(p2, p2) + (m2 * C_) * (m2 * C_));
  for (int d = 0; d < 3; d++) vc[d] = (p1[d] + p2[d]) * CC / (e1 + e2);
  const double vc_sq = dot3(vc, vc);
  const double vc_sq_cc = vc_sq / CC;
  const double gamma_rel2 = 1.0 / (1.0 - vc_sq_cc);
  const double gamma_rel = sqrt(gamma_rel2);
  const double gamma_rel_m1 = gamma_rel2 * vc_sq_cc / (gamma_rel + 1.0);
  const double p1_vc = dot3(p1, vc), p2_vc = dot3(p2, vc);
  double p3[3], p4[3], v3[3], v4[3], vr[3];
  double tvar = p1_vc * gamma_rel_m1 / (vc_sq + C_TINY);
  for (int d = 0; d < 3; d++) p3[d] = p1[d] + vc[d] * (tvar - gamma_rel * e1 / CC);
  tvar = p2_vc * gamma_rel_m1 / (vc_sq + C_TINY);
  for (int d = 0; d < 3; d++) p4[d] = p2[d] + vc[d] * (tvar - gamma_rel * e2 / CC);
  const double p3_mag = sqrt(dot3(p3, p3));
  const double e3 = gamma_rel * (e1 - p1_vc);
  const double e4 = gamma_rel * (e2 - p2_vc);
  for (int d = 0; d < 3; d++) { v3[d] = p3[d] * CC / e3; v4[d] = p4[d] * CC / e4; }
  tvar = 1.0 - (dot3(v3, v4) / CC);
  for (int d = 0; d < 3; d++) vr[d] = (v3[d] - v4[d]) / tvar;
  const double vrabs = sqrt(dot3(vr, vr));
  double nu = coll_freq(vrabs, E.log_lambda, m1, m2, E.q1, E.q2, E.dens);
  nu = fmin(nu * E.factor * E.np * E.dt_coll, 0.02);
  double c1[3], c2[3], c3[3];
  new_coords(vr, c1, c2, c3);
  const double ran1 = (1.0 - 1.0e-10) * R.next() + 0.5e-10;
  double ran2 = 2.0 * PI * R.next();
  const double delta = sqrt(-2.0 * nu * log(ran1)) * sin(ran2);
  ran2 = 2.0 * PI * R.next();
  double sin_theta = 2.0 * delta / (1.0 + delta * delta);
  double cos_theta = (1.0 - delta * delta) / (1.0 + delta * delta);
  const double *vcr = (m1 > m2) ? v3 : v4;
  const double vcr2 = dot3(vcr, vcr);
  const double gamma_rel_r = 1.0 / sqrt(1.0 - (vcr2 / CC));
  const double denominator = gamma_rel_r * (cos_theta - sqrt(vcr2) / fmax(vrabs, C_TINY));
  double tan_theta_cm, tan_theta_cm2;
  if (fabs(denominator) > sqrt(C_TINY)) {
    tan_theta_cm = sin_theta / denominator;
    tan_theta_cm2 = tan_theta_cm * tan_theta_cm;
  } else {
    tan_theta_cm = C_LARGEST;
    tan_theta_cm2 = C_LARGEST;
  }
  sin_theta = tan_theta_cm / sqrt(1.0 + tan_theta_cm2);
  cos_theta = 1.0 / sqrt(1.0 + tan_theta_cm2);
  const double cr = cos(ran2), sr = sin(ran2);
  for (int d = 0; d < 3; d++) {
    p3[d] = p3_mag * (c1[d] * cos_theta + c2[d] * sin_theta * cr + c3[d] * sin_theta * sr);
    p4[d] = -p3[d];
  }
  double p5[3], p6[3];
  tvar = dot3(p3, vc) * gamma_rel_m1 / vc_sq;
  for (int d = 0; d < 3; d++) p5[d] = p3[d] + vc[d] * (tvar + gamma_rel * e3 / CC);
  tvar = dot3(p4, vc) * gamma_rel_m1 / vc_sq;
  for (int d = 0; d < 3; d++) p6[d] = p4[d] + vc[d] * (tvar + gamma_rel * e4 / CC);
  const double wr = w1 / w2;
  const double e5 = C_ * sqrt(dot3(p5, p5) + (m1 * C_) * (m1 * C_));
  const double e6 = C_ * sqrt(dot3(p6, p6) + (m2 * C_) * (m2 * C_));
  if (wr > 1.0 + 2.0 * EPS) weighted_correction(w2 / w1, p1, p5, e1, e5, m1, R);
  else if (wr < 1.0 - 2.0 * EPS) weighted_correction(w1 / w2, p2, p6, e2, e6, m2, R);
  for (int d = 0; d < 3; d++) { p1[d] = p5[d]; p2[d] = p6[d]; }
  return true;
}

// One Nanbu / Perez pair (collisions.F90:516-633, :984-1101)
__device__ bool pair_np(const PairEnv &E, double *q1p, double *q2p, double w1, double w2, RanSrc &R) {
  const double m1 = E.m1, m2 = E.m2;
  double p1[3], p2[3], p1n[3], p2n[3], vc[3], v1[3], v2[3], p3[3], p4[3];
  // the quotients by constants and the three-component quotients by one divisor go through div_rcp (epb_internal.h):
  // the compiler's division carries a slow-path branch that keeps neighbouring divisions from overlapping
  const double rc = 1.0 / C_, rm0 = 1.0 / M0, rm1 = 1.0 / m1, rm2 = 1.0 / m2;
  for (int d = 0; d < 3; d++) { p1[d] = div_rcp(q1p[d], C_, rc); p2[d] = div_rcp(q2p[d], C_, rc); }
  for (int d = 0; d < 3; d++) { p1n[d] = div_rcp(p1[d], M0, rm0); p2n[d] = div_rcp(p2[d], M0, rm0); }
  if (dot3(p1n, p1n) < EPS && dot3(p2n, p2n) < EPS) return false;
  for (int d = 0; d < 3; d++) vc[d] = p1n[d] - p2n[d];
  if (dot3(vc, vc) < EPS) return false;
  for (int d = 0; d < 3; d++) p1n[d] = div_rcp(p1[d], m1, rm1);
  const double gm1 = sqrt(dot3(p1n, p1n) + 1.0) * m1;
  for (int d = 0; d < 3; d++) p2n[d] = div_rcp(p2[d], m2, rm2);
  const double gm2 = sqrt(dot3(p2n, p2n) + 1.0) * m2;
  const double gm = gm1 + gm2;
  const double rgm1 = 1.0 / gm1, rgm2 = 1.0 / gm2, rgm = 1.0 / gm;
  for (int d = 0; d < 3; d++) { v1[d] = div_rcp(p1[d], gm1, rgm1); v2[d] = div_rcp(p2[d], gm2, rgm2); }
  for (int d = 0; d < 3; d++) vc[d] = div_rcp(p1[d] + p2[d], gm, rgm);
  const double vc_sq = dot3(vc, vc);
  const double gamma_rel_inv = sqrt(1.0 - vc_sq);
  const double gc = 1.0 / gamma_rel_inv;
  const double gc_m1_vc = (gc - 1.0) / vc_sq;
  {
    const double t = (gc_m1_vc * dot3(vc, v1) - gc) * gm1;
    for (int d = 0; d < 3; d++) p3[d] = p1[d] + t * vc[d];
  }
  double v_sq = dot3(vc, v1);
  const double gm3 = (1.0 - v_sq) * gc * gm1;
  v_sq = dot3(vc, v2);
  const double gm4 = (1.0 - v_sq) * gc * gm2;
  const double p_mag2 = dot3(p3, p3);
  const double p_mag = sqrt(p_mag2);
  const double q12 = E.q1 * E.q2;
  const double fac = q12 * q12 * E.s_fac / (gm1 * gm2);
  const double t1 = gm3 * gm4 / p_mag2 + 1.0;
  double s12 = fac * gc * p_mag * C_ / gm * (t1 * t1);
  const double v_rel = gm * p_mag * C_ / (gm3 * gm4 * gc);
  const double s_prime = E.s_fac_prime * (m1 + m2) * v_rel / E.sp_den;
  s12 = fmin(s12, s_prime);
  double ran1 = R.next();
  const double ran2 = R.next() * 2.0 * PI;
  double cosp;
  if (s12 < 0.1) {
    cosp = 1.0 + s12 * log(fmax(ran1, 5e-9));
  } else if (s12 >= 0.1 && s12 < 3.0) {
    const double a_inv = 0.0056958 + (0.9560202 + (-0.508139 + (0.47913906 + (-0.12788975 + 0.02389567 * s12) * s12) * s12) * s12) * s12;
    const double a = 1.0 / a_inv;
    cosp = a_inv * log(exp(-a) + 2.0 * ran1 * sinh(a));
  } else if (s12 >= 3.0 && s12 < 6.0) {
    const double a = 3.0 * exp(-s12);
    cosp = log(exp(-a) + 2.0 * ran1 * sinh(a)) / a;
  } else {
    cosp = 2.0 * ran1 - 1.0;
  }
  cosp = fmax(fmin(cosp, 1.0), -1.0);
  const double sinp = sin(acos(cosp));
  const double p_perp2 = p3[0] * p3[0] + p3[1] * p3[1];
  const double p_perp = sqrt(p_perp2);
  const double p_tot = sqrt(p_perp2 + p3[2] * p3[2]);
  const double p_perp_inv = 1.0 / (p_perp + C_TINY);
  const double m11 = p3[0] * p3[2] * p_perp_inv, m12 = -p3[1] * p_tot * p_perp_inv, m13 = p3[0];
  const double m21 = p3[1] * p3[2] * p_perp_inv, m22 = p3[0] * p_tot * p_perp_inv, m23 = p3[1];
  const double m31 = -p_perp, m32 = 0.0, m33 = p3[2];
  const double sinp_cos = sinp * cos(ran2), sinp_sin = sinp * sin(ran2);
  p3[0] = m11 * sinp_cos + m12 * sinp_sin + m13 * cosp;
  p3[1] = m21 * sinp_cos + m22 * sinp_sin + m23 * cosp;
  p3[2] = m31 * sinp_cos + m32 * sinp_sin + m33 * cosp;
  for (int d = 0; d < 3; d++) p4[d] = -p3[d];
  const double t5 = gc_m1_vc * dot3(vc, p3) + gm3 * gc;
  const double t6 = gc_m1_vc * dot3(vc, p4) + gm4 * gc;
  if (E.inter) {
    ran1 = R.next();
    if (ran1 < w2 / w1)
      for (int d = 0; d < 3; d++) q1p[d] = (p3[d] + t5 * vc[d]) * C_;
    if (ran1 < w1 / w2)
      for (int d = 0; d < 3; d++) q2p[d] = (p4[d] + t6 * vc[d]) * C_;
  } else {
    for (int d = 0; d < 3; d++) { q1p[d] = (p3[d] + t5 * vc[d]) * C_; q2p[d] = (p4[d] + t6 * vc[d]) * C_; }
  }
  return true;
}

// calc_coulomb_log (collisions.F90:1288-1316) for one cell; temp2 in eV
__device__ double coulomb_log_cell(double ekbar1, double temp2, double dens1, double dens2, double q1, double q2, double m1) {
  const double local_ekbar1 = fmax(ekbar1, 100.0 * Q0);
  const double local_temp2 = fmax(temp2, 100.0);
  if (dens1 <= 1.0 || dens2 <= 1.0) return 1.0;
  const double bmax = sqrt(EPB_EPS0 * Q0 * local_temp2 / (fabs(q2) * Q0 * dens2));
  const double b0 = fabs(q1 * q2) / (8.0 * PI * EPB_EPS0 * local_ekbar1);
  const double gamm = (local_ekbar1 / (m1 * CC)) + 1.0;
  const double dB = 2.0 * PI * H_BAR / (sqrt(gamm * gamm - 1.0) * m1 * C_);
  const double bmin = fmax(b0, dB);
  return fmax(1.0, log(bmax / bmin));
}

struct CollOp {
  // species 1 / species 2 columns (the same pointers for an intra-species call): momenta and weight of row 0
  double *p1[3], *w1, *p2[3], *w2;
  const int *cnt1, *cnt2;
  int R1, R2, rowd;
  int ngroups;
  int intra, nanbu;
  double m1, m2, q1, q2;
  double user_factor, dt_coll, idxy, dx, dy;
  double log_lambda;            // > 0: fixed
  const double *ekbar1, *temp2; // coulomb_log_auto: grid arrays (ex-like extent), temp2 in K
  int sz0, n0, n1;
  TileGeom tg;
  unsigned long long seed;
};

__device__ __forceinline__ size_t field_ofs_of_key(const CollOp &O, int key) {
  const int tile = key / O.tg.cpt, in = key - tile * O.tg.cpt;
  const int tx = tile % O.tg.nt[0], ty = tile / O.tg.nt[0];
  const int cx = tx * O.tg.T[0] + in % O.tg.T[0] + 1, cy = ty * O.tg.T[1] + in / O.tg.T[0] + 1;   // 1-based cell
  return (size_t)(cx + NGC - 1) + (size_t)O.sz0 * (size_t)(cy + NGC - 1);
}

// One warp per group of 32 cells, one lane per cell.
__global__ void __launch_bounds__(128) k_collide(const __grid_constant__ CollOp O) {
  const int lane = threadIdx.x & 31;
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= O.ngroups) return;
  const int key = g * 32 + lane;
  int n1 = O.cnt1[key], n2 = O.cnt2[key];
  if (n1 > O.R1) n1 = O.R1;
  if (n2 > O.R2) n2 = O.R2;
  const size_t c1 = ((size_t)g * O.R1) * O.rowd + lane, c2 = ((size_t)g * O.R2) * O.rowd + lane;
  auto at1 = [&](int r) { return c1 + (size_t)r * O.rowd; };
  auto at2 = [&](int r) { return c2 + (size_t)r * O.rowd; };
  int pcount;
  bool live;
  if (O.intra) { live = n1 > 1; pcount = n1 / 2 + (n1 & 1); }
  else { live = n1 > 0 && n2 > 0; pcount = n1 > n2 ? n1 : n2; }
  if (!live) pcount = 0;
  // cell sums: number density (calc_coll_number_density: the sum of the cell's own particles), np, factor
  double sw1 = 0.0, sw2 = 0.0, np = 0.0, factor = 0.0;
  for (int r = 0; r < n1; r++) sw1 += O.w1[at1(r)];
  if (!O.intra) for (int r = 0; r < n2; r++) sw2 += O.w2[at2(r)];
  // Durstenfeld's shuffle of species 1's rows (collisions.F90:1224-1284); p_num <= 2: nothing to be done
  unsigned char perm[256];
  const bool shuffled = live && n1 > 2 && n1 <= 256;
  if (shuffled) {
    Rng rg;
    rg.s = O.seed ^ (0xC2B2AE3D27D4EB4Full * ((unsigned long long)key + 1));
    for (int i = 0; i < n1; i++) perm[i] = (unsigned char)i;
    for (int idx = n1; idx >= 2; idx--) {
      int sw = (int)floor((double)idx * rg.next());
      if (sw > idx - 1) sw = idx - 1;
      const unsigned char t = perm[idx - 1];
      perm[idx - 1] = perm[sw];
      perm[sw] = t;
    }
  }
  auto row1 = [&](int r) { return shuffled ? (int)perm[r] : r; };
  if (live) {
    if (O.intra) {
      for (int k = 0; k < pcount; k++) {   // ring: the partner of an odd cell's last particle is the head
        const double wa = O.w1[at1(row1(2 * k))], wb = O.w1[at1(row1((2 * k + 1) % n1))];
        np = np + wa + wb;
        factor = factor + fmin(wa, wb);
      }
      factor = O.user_factor / factor / 2.0;
    } else {
      np = n1 >= n2 ? sw1 : sw2;
      for (int k = 0; k < pcount; k++) factor = factor + fmin(O.w1[at1(row1(k % n1))], O.w2[at2(k % n2)]);
      factor = O.user_factor / factor;
    }
  }
  const double idens = sw1 * O.idxy, jdens = O.intra ? idens : sw2 * O.idxy;
  PairEnv E;
  E.m1 = O.m1; E.m2 = O.m2; E.q1 = O.q1; E.q2 = O.q2;
  E.inter = !O.intra;
  double log_lambda = O.log_lambda;
  if (live && !(log_lambda > 0.0)) {
    const size_t fo = field_ofs_of_key(O, key);
    log_lambda = coulomb_log_cell(O.ekbar1[fo], O.temp2[fo] * (EPB_KB / Q0), idens, jdens, O.q1, O.q2, O.m1);
  }
  E.log_lambda = log_lambda;
  if (O.nanbu) {
    const double pi4_eps2_c4 = 4.0 * PI * (EPB_EPS0 * EPB_EPS0) * (CC * CC);
    const double two_thirds = 2.0 / 3.0;
    const double pi_fac = pow(4.0 * PI / 3.0, 1.0 / 3.0);
    if (O.intra) {
      const double cell_fac = idens * idens * O.dt_coll * factor * O.dx * O.dy;
      E.s_fac = cell_fac * log_lambda / pi4_eps2_c4;
      E.s_fac_prime = cell_fac * pi_fac / pow(idens, two_thirds);
      E.sp_den = fmax(O.m1, O.m2);
    } else {
      const double cell_fac = idens * jdens * O.dt_coll * factor * O.dx * O.dy;
      E.s_fac = cell_fac * log_lambda / pi4_eps2_c4;
      E.s_fac_prime = cell_fac * pi_fac;
      E.sp_den = fmax(O.m1 * pow(idens, two_thirds), O.m2 * pow(jdens, two_thirds));
    }
    E.dens = 0.0; E.factor = E.np = E.dt_coll = 0.0;
  } else {
    E.dens = O.intra ? idens : fmin(idens, jdens);
    E.factor = factor; E.np = np; E.dt_coll = O.dt_coll;
    E.s_fac = E.s_fac_prime = E.sp_den = 0.0;
  }
  for (int k = 0; k < pcount; k++) {
    const int ra = O.intra ? row1(2 * k) : row1(k % n1);
    const int rb = O.intra ? row1((2 * k + 1) % n1) : k % n2;
    const size_t ia = at1(ra), ib = O.intra ? at1(rb) : at2(rb);
    double pa[3], pb[3];
    for (int d = 0; d < 3; d++) { pa[d] = O.p1[d][ia]; pb[d] = (O.intra ? O.p1[d] : O.p2[d])[ib]; }
    const double wa = O.w1[ia], wb = (O.intra ? O.w1 : O.w2)[ib];
    RanSrc R;
    R.fixed = nullptr; R.k = 0;
    R.g.s = O.seed ^ (0xD1B54A32D192ED03ull * ((unsigned long long)key + 1)) ^ (0x8CB92BA72F3D8DD7ull * ((unsigned long long)k + 1));
    const bool done = O.nanbu ? pair_np(E, pa, pb, wa, wb, R) : pair_sk(E, pa, pb, wa, wb, R);
    if (done) {
      for (int d = 0; d < 3; d++) { O.p1[d][ia] = pa[d]; (O.intra ? O.p1[d] : O.p2[d])[ib] = pb[d]; }
    }
  }
}

// test entry: pairs given explicitly, random numbers from an array (4 per pair)
struct PairTestOp {
  double *p1, *p2;          // [n][3]
  const double *w1, *w2;    // [n]
  const double *ran;        // [n][4]
  int *done;
  int n, nanbu;
  PairEnv E;
};
__global__ void k_pair_test(const __grid_constant__ PairTestOp T) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T.n; i += gridDim.x * blockDim.x) {
    double a[3] = {T.p1[3 * i], T.p1[3 * i + 1], T.p1[3 * i + 2]}, b[3] = {T.p2[3 * i], T.p2[3 * i + 1], T.p2[3 * i + 2]};
    RanSrc R;
    R.fixed = T.ran + 4 * (size_t)i;
    R.k = 0;
    R.g.s = 0;
    const bool d = T.nanbu ? pair_np(T.E, a, b, T.w1[i], T.w2[i], R) : pair_sk(T.E, a, b, T.w1[i], T.w2[i], R);
    T.done[i] = d ? 1 : 0;
    for (int q = 0; q < 3; q++) { T.p1[3 * i + q] = a[q]; T.p2[3 * i + q] = b[q]; }
  }
}

}  // namespace

extern "C" int epb_collide(epb_handle *h, const epb_collisions *c) {
  if (!h || !c || !c->coll_pairs) return EPB_ERR_ARG;
  const epb_config &cf = h->cfg;
  const int nsp = (int)h->sp.size();
  if (c->n_species != nsp) return epb_fail(h, EPB_ERR_ARG, "epb_collide: n_species mismatch");
  if (cf.ndims != 2 || h->tg.layout != 2)
    return epb_fail(h, EPB_ERR_UNSUPPORTED, "epb_collide: binary collisions run on the 2D slot-column layout (epoch2d, default kernel)");
  const int coll_n_step = c->coll_n_step > 0 ? c->coll_n_step : 1;
  const bool auto_log = !(c->coulomb_log > 0.0);
  h->coll_calls++;
  for (int is = 0; is < nsp; is++) {   // the columns must hold every particle (no mover in flight between two columns)
    int rc = epb_slots_settle(h, is);
    if (rc) return rc;
  }
  const int ngroups = h->tg.nkeys / 32;
  const int blocks = (ngroups * 32 + 127) / 128;
  if (auto_log && !h->coll_work) {
    EPB_CUDA(h, cudaMalloc(&h->coll_work, 2 * h->fsize * sizeof(double)));
  }
  for (int is = 0; is < nsp; is++) {
    SpeciesDev &S1 = h->sp[is];
    if (fabs(S1.cfg.charge) <= C_TINY) continue;
    bool any = false;
    for (int js = is; js < nsp; js++) any = any || c->coll_pairs[is * nsp + js] > 0.0;
    if (!any) continue;
    if (auto_log) {
      int rc = epb_coll_moment_dev(h, 0, is, h->coll_work);   // calc_coll_ekbar(iekbar, ispecies)
      if (rc) return rc;
    }
    for (int js = is; js < nsp; js++) {
      const double user_factor = c->coll_pairs[is * nsp + js];
      if (!(user_factor > 0.0)) continue;
      SpeciesDev &S2 = h->sp[js];
      if (fabs(S2.cfg.charge) <= C_TINY) continue;
      if (auto_log) {
        int rc = epb_coll_moment_dev(h, 1, js, h->coll_work + h->fsize);   // calc_coll_temperature_ev(jtemp, jspecies)
        if (rc) return rc;
      }
      CollOp O;
      memset(&O, 0, sizeof O);
      for (int d = 0; d < 3; d++) { O.p1[d] = S1.buf[0][3 + d]; O.p2[d] = S2.buf[0][3 + d]; }
      O.w1 = S1.buf[0][6]; O.w2 = S2.buf[0][6];
      O.cnt1 = S1.cnt; O.cnt2 = S2.cnt;
      O.R1 = S1.R; O.R2 = S2.R; O.rowd = S1.rowd;
      O.ngroups = ngroups;
      O.intra = (is == js);
      O.nanbu = c->use_nanbu ? 1 : 0;
      O.m1 = S1.cfg.mass; O.m2 = S2.cfg.mass; O.q1 = S1.cfg.charge; O.q2 = S2.cfg.charge;
      O.user_factor = user_factor;
      O.dt_coll = cf.dt * (double)coll_n_step;
      O.idxy = 1.0 / cf.dx[0] / cf.dx[1];
      O.dx = cf.dx[0]; O.dy = cf.dx[1];
      O.log_lambda = auto_log ? 0.0 : c->coulomb_log;
      O.ekbar1 = h->coll_work; O.temp2 = h->coll_work ? h->coll_work + h->fsize : nullptr;
      O.sz0 = h->sz[0]; O.n0 = cf.n[0]; O.n1 = cf.n[1];
      O.tg = h->tg;
      O.seed = (c->seed + 0x632BE59BD9B4E019ull * (unsigned long long)h->coll_calls) ^
               (0xA24BAED4963EE407ull * (unsigned long long)(is * nsp + js + 1)) ^ (0x9FB21C651E98DF25ull * (unsigned long long)(cf.rank + 1));
      k_collide<<<blocks, 128, 0, h->stream>>>(O);
      h->launches++;
    }
  }
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

// Test entry point (tests/test_collisions.py): the device's pair operator on explicit pairs with given random
// numbers, for a pair-by-pair comparison with the oracle's restatement.  env[14]: m1 m2 q1 q2 dens log_lambda factor
// np dt_coll s_fac s_fac_prime sp_den inter nanbu
extern "C" int epb_collide_pairs_test(int n, double *p1, double *p2, const double *w1, const double *w2, const double *ran,
                                      const double *env, int *done) {
  if (n <= 0) return EPB_OK;
  double *d_p1, *d_p2, *d_w1, *d_w2, *d_ran;
  int *d_done;
  if (cudaMalloc(&d_p1, 3 * n * sizeof(double)) || cudaMalloc(&d_p2, 3 * n * sizeof(double)) || cudaMalloc(&d_w1, n * sizeof(double)) ||
      cudaMalloc(&d_w2, n * sizeof(double)) || cudaMalloc(&d_ran, 4 * n * sizeof(double)) || cudaMalloc(&d_done, n * sizeof(int)))
    return EPB_ERR_CUDA;
  cudaMemcpy(d_p1, p1, 3 * n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(d_p2, p2, 3 * n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(d_w1, w1, n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(d_w2, w2, n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(d_ran, ran, 4 * n * sizeof(double), cudaMemcpyHostToDevice);
  PairTestOp T;
  T.p1 = d_p1; T.p2 = d_p2; T.w1 = d_w1; T.w2 = d_w2; T.ran = d_ran; T.done = d_done; T.n = n;
  T.E.m1 = env[0]; T.E.m2 = env[1]; T.E.q1 = env[2]; T.E.q2 = env[3]; T.E.dens = env[4]; T.E.log_lambda = env[5];
  T.E.factor = env[6]; T.E.np = env[7]; T.E.dt_coll = env[8];
  T.E.s_fac = env[9]; T.E.s_fac_prime = env[10]; T.E.sp_den = env[11]; T.E.inter = (int)env[12];
  T.nanbu = (int)env[13];
  k_pair_test<<<(n + 127) / 128, 128>>>(T);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(p1, d_p1, 3 * n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(p2, d_p2, 3 * n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(done, d_done, n * sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_p1); cudaFree(d_p2); cudaFree(d_w1); cudaFree(d_w2); cudaFree(d_ran); cudaFree(d_done);
  return e == cudaSuccess ? EPB_OK : EPB_ERR_CUDA;
}
