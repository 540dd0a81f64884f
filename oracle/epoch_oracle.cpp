// epoch_oracle.cpp — CPU restatement of EPOCH's per-timestep PIC hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load it.  The product library (epoch_b200/csrc) never links or calls it.
//
// Parity status: the FIELD half (FDTD + field BCs + laser/outflow boundary) is
// pinned against the reference's own golden scalars
// (epoch{1,2,3}d/tests/test_laser.py:70-80).  The PARTICLE half (push, deposit,
// particle exchange) is "parity unpinned": the reference ships no numeric
// assertion for it (tests/test_twostream.py, test_landau.py only plot), and the
// reference binary (Fortran 2003 + MPI) cannot be built in this image.  It is
// checked instead through exact discrete charge conservation, which is the
// property particles.F90:30-34 claims.
//
// Each routine cites the reference file:line it restates.  2D line numbers are
// epoch2d/src/..., with the 1D/3D trees cited where they differ textually.
// Arithmetic keeps the reference's operation order; build with
// -ffp-contract=off (the reference's gfortran -O3 build has no -march flag and
// therefore no FMA contraction, epoch2d/Makefile:72).
//
// Supported: ndims 1/2/3, triangle shape (default build), per-particle weight,
// Boris push, Yee order-2 FDTD, field BCs periodic/clamp/zero_gradient/
// simple_laser/simple_outflow (laser/outflow on x_min/x_max), particle BCs
// periodic/reflect/open, in-process multi-"rank" domain decomposition that
// replays the MPI_SENDRECV sequences of boundary.F90.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <vector>

namespace {

// constants.F90:192-199
constexpr double pi = 3.141592653589793238462643383279503;
constexpr double q0 = 1.602176565e-19;
constexpr double m0 = 9.10938291e-31;
constexpr double c = 2.99792458e8;
constexpr double kb = 1.3806488e-23;
constexpr double epsilon0 = 8.854187817620389850536563031710750e-12;

// constants.F90:75-90
enum {
  c_bc_periodic = 1,
  c_bc_other = 2,
  c_bc_simple_laser = 3,
  c_bc_simple_outflow = 4,
  c_bc_open = 5,
  c_bc_zero_gradient = 7,
  c_bc_clamp = 8,
  c_bc_reflect = 9,
  c_bc_conduct = 10,
  c_bc_thermal = 11,
  c_bc_cpml_laser = 12,
  c_bc_cpml_outflow = 13,
};

// constants.F90:549-559 (triangle): sf_min=-1, sf_max=1, png=3, ng=png+2
constexpr int sf_min = -1, sf_max = 1, png = 3, NG = png + 2;

enum { EX, EY, EZ, BX, BY, BZ, JX, JY, JZ, WK, WK1, WK2, WK3, WK4, NFIELD };  // WK: work array of smooth_array; WK..WK4: calc_df

// Fortran-style array  a(1-g:n1+g [, 1-g:n2+g [, 1-g:n3+g]])
struct Arr {
  int lo[3] = {1, 1, 1}, sz[3] = {1, 1, 1};
  std::vector<double> v;
  void init(const int n[3], int nd, int g = NG) {
    for (int d = 0; d < 3; d++) {
      if (d < nd) { lo[d] = 1 - g; sz[d] = n[d] + 2 * g; }
      else { lo[d] = 1; sz[d] = 1; }
    }
    v.assign((size_t)sz[0] * sz[1] * sz[2], 0.0);
  }
  inline size_t idx(int i, int j, int k) const {
    return (size_t)(i - lo[0]) + (size_t)sz[0] * ((size_t)(j - lo[1]) + (size_t)sz[1] * (size_t)(k - lo[2]));
  }
  inline double &operator()(int i, int j = 1, int k = 1) { return v[idx(i, j, k)]; }
  inline double operator()(int i, int j = 1, int k = 1) const { return v[idx(i, j, k)]; }
};

struct Particle {  // shared_data.F90:93-142 (default flags)
  double pos[3];
  double p[3];
  double w;
};

struct SpeciesCfg {  // C-ABI mirror, see orc_species in epoch_oracle.h
  double charge, mass;
  int bc_particle[6];
  double npart_per_cell;
  double density;      // uniform number density inside the box below
  double box_lo[3], box_hi[3];  // density = 0 outside [lo,hi) per active dim
  double temp[3];      // K
  double drift[3];     // kg m/s
  int zero_current;
  int immobile;
};

struct Config {
  int ndims;
  int n_global[3];
  int nproc[3];
  double xmin[3], xmax[3];
  int bc_field[6];
  double dt;
  int n_species;
  int seed;  // 7842432 by default (setup.F90:567)
  // fields.f90:32-100: finite-difference order (2, 4, 6) of the Yee solver; extended 2D stencils
  // (order 2 only): maxwell_solver != 0 with the coefficients set_maxwell_solver derives
  int field_order;
  int maxwell_solver;
  // alpha[a], beta[a][other axis, lower first], gamma[a], delta[a] (epoch3d fields.f90:53-162)
  double st_alpha[3], st_beta[6], st_gamma[3], st_delta[3];
  // current_smooth.F90:50-141: smooth_currents with smooth_its (+ smooth_comp_its) passes over strides
  int smooth_its, smooth_comp_its, smooth_nstrides, smooth_strides[4];
  int force_mixed;  // test hook: take the per-species current_bcs path even if all species agree
  int hc_push;      // -DHC_PUSH: Higuera-Cary gamma in the rotation (particles.F90:386-398)
  // CPML boundaries (boundary.F90:1479-2025); defaults of setup.F90:80-83: 6, 20, 0.15, 0.7.  Ignored (thickness 0)
  // unless some bc_field is cpml_laser / cpml_outflow (mpi_routines.F90:285)
  int cpml_thickness;
  double cpml_kappa_max, cpml_a_max, cpml_sigma_max;
};

// random_generator.f90:23-78 (KISS), :112-173 (polar Box-Muller)
struct Rng {
  int32_t x, y, z, w;
  bool cached = false;
  double cached_value = 0.0;
  static inline int32_t wrap(int64_t a) { return (int32_t)(uint32_t)(uint64_t)a; }
  double random() {
    x = wrap((int64_t)69069 * x + 1327217885);
    uint32_t a1 = (uint32_t)y;
    uint32_t a2 = a1 ^ (a1 << 13);
    uint32_t a3 = a2 ^ (a2 >> 17);
    y = (int32_t)(a3 ^ (a3 << 5));
    z = wrap((int64_t)18000 * (z & 65535) + (int64_t)((uint32_t)z >> 16));
    w = wrap((int64_t)30903 * (w & 65535) + (int64_t)((uint32_t)w >> 16));
    int32_t kiss = wrap((int64_t)x + (int64_t)y + (int64_t)(int32_t)((uint32_t)z << 16) + (int64_t)w);
    return ((double)kiss + 2147483648.0) / 4294967296.0;
  }
  void init(int seed) {
    x = wrap((int64_t)123456789 + seed);
    y = wrap((int64_t)362436069 + seed);
    z = wrap((int64_t)521288629 + seed);
    w = wrap((int64_t)916191069 + seed);
    cached = false;
    cached_value = 0.0;
    for (int i = 0; i < 1000; i++) (void)random();
  }
  double box_muller(double stdev, double mu) {
    const double c_tiny = 2.2250738585072014e-308;
    double r;
    if (cached) {
      cached = false;
      r = cached_value * stdev + mu;
    } else {
      cached = true;
      double rand1, rand2, ww;
      for (;;) {
        rand1 = random();
        rand2 = random();
        rand1 = 2.0 * rand1 - 1.0;
        rand2 = 2.0 * rand2 - 1.0;
        ww = rand1 * rand1 + rand2 * rand2;
        if (ww > c_tiny && ww < 1.0) break;
      }
      ww = std::sqrt((-2.0 * std::log(ww)) / ww);
      r = rand1 * ww * stdev + mu;
      cached_value = rand2 * ww;
    }
    return r;
  }
};

struct Rank {
  int rank;
  int coords[3] = {0, 0, 0};
  int n[3] = {1, 1, 1};           // nx, ny, nz (local)
  int gmin[3] = {1, 1, 1};        // nx_global_min ...
  bool is_bnd[6] = {false, false, false, false, false, false};
  int neighbour[3][3][3];         // [iz+1][iy+1][ix+1], -1 = MPI_PROC_NULL
  double grid_min_local[3] = {0, 0, 0};
  double min_local[3] = {0, 0, 0}, max_local[3] = {0, 0, 0};
  Arr f[NFIELD];
  // setup.F90:391-447 boundary snapshots on x_min / x_max (index by field 0..5)
  Arr snap_min[6], snap_max[6];
  // the same for the y and z faces: snapA[axis][side][field], srcA[axis][side][source1|source2]
  // (planes with a unit extent along the axis; axis 0 keeps the arrays above)
  Arr snapA[3][2][6], srcA[3][2][2];
  Arr src1[2], src2[2];  // laser sources on x_min / x_max, set by the host each step
  // thermal particle boundaries: ext_temp_x_min ... (shared_data.F90:255-256), [species][side] = (plane, 3): the two
  // transverse axes in axis order with ghost cells, lower axis fastest; empty = never set
  std::vector<std::vector<std::vector<double>>> ext_temp;
  // CPML (boundary.F90:1479-1770): per axis the stretching profiles on the E and the B points (index i - (1 - NG),
  // 1.0 / 0.0 outside the layers), the local index range of the two layers [axis][side] (start > end: none here),
  // the particle-domain offsets, the laser plane of a cpml_laser face, and the four auxiliary arrays of the axis
  // (psi of E_b, E_c, B_b, B_c for the cyclic (a, b, c))
  std::vector<double> kap_e[3], kap_b[3], sig_e[3], sig_b[3], a_e[3], a_b[3];
  int cp_start[3][2], cp_end[3][2], cp_off[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  int laser_idx[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  bool add_laser_cpml[3][2] = {{false, false}, {false, false}, {false, false}};
  Arr psi[3][4];
  std::vector<std::vector<Particle>> part;       // per species
  std::vector<std::vector<int64_t>> bnd_cand;    // boundary candidate indices per species
  std::vector<std::vector<Particle>> inserted;   // per species: what the window's insert_particles created since orc_window_clear_inserted
  Rng rng;
};

struct World {
  Config cfg;
  std::vector<SpeciesCfg> sp;
  int nd;
  int nranks;
  double d[3] = {1, 1, 1};  // dx,dy,dz
  double length[3] = {0, 0, 0};
  double grid_min[3] = {0, 0, 0};  // x_grid_min (cell centre of global cell 1)
  double xb_min[3] = {0, 0, 0};    // xb_global(1): the low edge of global cell 1 (setup.F90:176-177), moved by the window
  double min_outer[3], max_outer[3];
  std::vector<int> cell_min[3], cell_max[3];
  bool periods[3] = {false, false, false};
  int bc_field[6];
  int bc_allspecies[6];
  int cpml_t = 0;          // cpml_thickness, 0 without CPML boundaries
  int n_ext[3] = {1, 1, 1};  // nx_global after mpi_routines.F90:295-296 (deck value + 2 cpml_thickness)
  bool bc_mixed = false;   // c_bc_mixed on some boundary: J is folded per species (boundary.F90:547-556, 790-796)
  double dt;
  std::vector<Rank> r;
};

inline double x_global(const World &w, int d, int i) {
  // setup.F90:188  x_global(ix) = x_grid_min + (ix - 1) * dx
  return w.grid_min[d] + (double)(i - 1) * w.d[d];
}

// set_cpml_helpers + allocate_cpml_fields (boundary.F90:1479-1794; epoch3d :1891-2330, epoch1d :783-925): the same
// code per axis.  kappa / sigma / a on the E points (integer positions) and on the B points (half positions), the
// layer's local index range, the offset that keeps the layer out of the particle domain (utilities.f90:364-365) and
// the laser plane cpml_thickness + fng + 1 cells in.
inline bool is_cpml(int b) { return b == c_bc_cpml_laser || b == c_bc_cpml_outflow; }
void set_cpml_helpers(World &w, Rank &R) {
  const Config &cf = w.cfg;
  const int t = w.cpml_t;
  const int cpml_m = 3, cpml_ma = 1;
  // fng, "the number of ghost cells needed by the field solver": field_order / 2 (fields.f90:37), but 2 for the Lehe
  // solvers (deck_control_block.F90:117-120; epoch3d :118-122 adds lehe_z) -- it decides where the laser plane sits
  int fng = (cf.field_order ? cf.field_order : 2) / 2;
  if (cf.maxwell_solver >= 2 && cf.maxwell_solver <= 4) fng = 2;   // c_maxwell_solver_lehe_x / _y / _z
  for (int d = 0; d < w.nd; d++) {
    const int n = R.n[d], len = n + 2 * NG;
    R.kap_e[d].assign(len, 1.0); R.kap_b[d].assign(len, 1.0);
    R.sig_e[d].assign(len, 0.0); R.sig_b[d].assign(len, 0.0);
    R.a_e[d].assign(len, 0.0); R.a_b[d].assign(len, 0.0);
    for (int sd = 0; sd < 2; sd++) { R.cp_start[d][sd] = n + 1; R.cp_end[d][sd] = 0; R.cp_off[d][sd] = 0; }
    if (t == 0) continue;
    const int gmin = R.gmin[d], gmax = R.gmin[d] + n - 1, ng_ = w.n_ext[d];
    // note: dx of the FIRST axis for every axis's sigma (boundary.F90:1517 uses dx for x and y alike)
    const double sigma_maxval = cf.cpml_sigma_max * c * 0.8 * (cpml_m + 1.0) / w.d[0];
    auto pw = [](double x, int e) { double r = 1.0; for (int q = 0; q < e; q++) r = r * x; return r; };
    auto at = [&](std::vector<double> &v, int i) -> double & { return v[i - (1 - NG)]; };
    if (is_cpml(w.bc_field[2 * d])) {
      if (gmin <= t) {
        R.cp_start[d][0] = 1;
        if (gmax >= t) { R.cp_end[d][0] = t - gmin + 1; R.cp_off[d][0] = t - gmin + 1; }
        else { R.cp_end[d][0] = n; R.cp_off[d][0] = t; }
        for (int i = R.cp_start[d][0]; i <= R.cp_end[d][0]; i++) {
          const int ig = i + gmin - 1;
          double x_pos = 1.0 - (double)(ig - 1) / (double)t;
          at(R.kap_e[d], i) = 1.0 + (cf.cpml_kappa_max - 1.0) * pw(x_pos, cpml_m);
          at(R.sig_e[d], i) = sigma_maxval * pw(x_pos, cpml_m);
          at(R.a_e[d], i) = cf.cpml_a_max * pw(1.0 - x_pos, cpml_ma);
          x_pos = 1.0 - ((double)ig - 0.5) / (double)t;
          at(R.kap_b[d], i) = 1.0 + (cf.cpml_kappa_max - 1.0) * pw(x_pos, cpml_m);
          at(R.sig_b[d], i) = sigma_maxval * pw(x_pos, cpml_m);
          at(R.a_b[d], i) = cf.cpml_a_max * pw(1.0 - x_pos, cpml_ma);
        }
      }
      if (gmin <= t + fng + 1 && gmax >= t + fng + 1) {
        R.add_laser_cpml[d][0] = true;
        R.laser_idx[d][0] = t + fng + 1 - gmin;
      }
    }
    if (is_cpml(w.bc_field[2 * d + 1])) {
      if (gmax >= ng_ - t + 1) {
        R.cp_end[d][1] = n;
        if (gmin <= ng_ - t + 1) { R.cp_start[d][1] = ng_ - t + 1 - gmin + 1; R.cp_off[d][1] = t - ng_ + gmax; }
        else { R.cp_start[d][1] = 1; R.cp_off[d][1] = t; }
        for (int i = R.cp_start[d][1]; i <= R.cp_end[d][1]; i++) {
          const int ig = ng_ - (i + gmin - 1) + 1;
          double x_pos = 1.0 - (double)(ig - 1) / (double)t;
          at(R.kap_e[d], i) = 1.0 + (cf.cpml_kappa_max - 1.0) * pw(x_pos, cpml_m);
          at(R.sig_e[d], i) = sigma_maxval * pw(x_pos, cpml_m);
          at(R.a_e[d], i) = cf.cpml_a_max * pw(1.0 - x_pos, cpml_ma);
          x_pos = 1.0 - ((double)ig - 0.5) / (double)t;
          at(R.kap_b[d], i - 1) = 1.0 + (cf.cpml_kappa_max - 1.0) * pw(x_pos, cpml_m);
          at(R.sig_b[d], i - 1) = sigma_maxval * pw(x_pos, cpml_m);
          at(R.a_b[d], i - 1) = cf.cpml_a_max * pw(1.0 - x_pos, cpml_ma);
        }
      }
      if (gmin <= ng_ - t - fng + 2 && gmax >= ng_ - t - fng + 2) {
        R.add_laser_cpml[d][1] = true;
        R.laser_idx[d][1] = ng_ - t - fng + 2 - gmin;
      }
    }
    // utilities.f90:364-365
    R.min_local[d] = R.min_local[d] + R.cp_off[d][0] * w.d[d];
    R.max_local[d] = R.max_local[d] - R.cp_off[d][1] * w.d[d];
    for (int q = 0; q < 4; q++) R.psi[d][q].init(R.n, w.nd);
  }
}

// ---------------------------------------------------------------------------
// Setup: boundaries.F90:30-74, mpi_routines.F90:179-275,279-365,
// setup.F90:162-204, utilities.f90:343-421
// ---------------------------------------------------------------------------
void setup_world(World &w) {
  const Config &cf = w.cfg;
  const int nd = w.nd = cf.ndims;
  w.dt = cf.dt;
  for (int i = 0; i < 6; i++) {
    int b = cf.bc_field[i];
    // boundary.F90:44-57
    if (b == c_bc_other) b = c_bc_clamp;
    if (b == c_bc_reflect) b = c_bc_clamp;
    if (b == c_bc_open) b = c_bc_simple_outflow;
    w.bc_field[i] = b;
  }
  for (auto &s : w.sp)
    for (int i = 0; i < 2 * nd; i++) {
      int &b = s.bc_particle[i];
      // boundary.F90:108-122
      if (b == c_bc_other || b == c_bc_conduct) b = c_bc_reflect;
      if (b == c_bc_simple_laser || b == c_bc_simple_outflow || b == c_bc_cpml_laser || b == c_bc_cpml_outflow) b = c_bc_open;
    }
  // deck_species_block.F90:182-199
  for (int i = 0; i < 6; i++) w.bc_allspecies[i] = c_bc_open;
  for (size_t is = 0; is < w.sp.size(); is++)
    for (int i = 0; i < 2 * nd; i++) {
      int b = w.sp[is].bc_particle[i];
      if (b != c_bc_reflect && b != c_bc_periodic) b = c_bc_open;
      if (is == 0) w.bc_allspecies[i] = b;
      else if (w.bc_allspecies[i] != b) { w.bc_allspecies[i] = -99; w.bc_mixed = true; }  // c_bc_mixed
    }
  if (cf.force_mixed && !w.sp.empty()) w.bc_mixed = true;
  if (w.sp.empty())
    for (int i = 0; i < 2 * nd; i++)
      w.bc_allspecies[i] = (cf.bc_field[i] == c_bc_periodic) ? c_bc_periodic : c_bc_open;

  // boundary.F90:43-48, mpi_routines.F90:285, :295-296: every axis grows by 2 cpml_thickness cells as soon as one
  // boundary is a CPML; setup.F90:167-181: dx from the deck's cell count, the grid starts cpml_thickness cells out
  w.cpml_t = 0;
  for (int i = 0; i < 2 * nd; i++)
    if (w.bc_field[i] == c_bc_cpml_laser || w.bc_field[i] == c_bc_cpml_outflow) w.cpml_t = cf.cpml_thickness;
  for (int d = 0; d < nd; d++) {
    w.n_ext[d] = cf.n_global[d] + 2 * w.cpml_t;
    w.length[d] = cf.xmax[d] - cf.xmin[d];
    w.d[d] = w.length[d] / (double)(w.n_ext[d] - 2 * w.cpml_t);
    double g = cf.xmin[d] - w.d[d] * w.cpml_t;
    w.xb_min[d] = g;
    w.grid_min[d] = g + w.d[d] / 2.0;
    // utilities.f90:367-369
    double boundary_shift = (double)((1 + png + w.cpml_t) / 2);
    w.min_outer[d] = cf.xmin[d] - boundary_shift * w.d[d];
    w.max_outer[d] = cf.xmax[d] + boundary_shift * w.d[d];
  }
  // mpi_routines.F90:194-221
  for (int d = 0; d < nd; d++) {
    bool per = (w.bc_field[2 * d] == c_bc_periodic);
    for (auto &s : w.sp)
      if (s.bc_particle[2 * d] == c_bc_periodic) per = true;
    w.periods[d] = per;
  }
  // mpi_routines.F90:317-351
  int np[3] = {1, 1, 1};
  for (int d = 0; d < nd; d++) {
    np[d] = std::max(1, cf.nproc[d]);
    int ng_ = w.n_ext[d];
    int n0 = ng_ / np[d];
    int nxp = (n0 * np[d] != ng_) ? (n0 + 1) * np[d] - ng_ : np[d];
    w.cell_min[d].resize(np[d]);
    w.cell_max[d].resize(np[d]);
    for (int i = 1; i <= nxp; i++) {
      w.cell_min[d][i - 1] = (i - 1) * n0 + 1;
      w.cell_max[d][i - 1] = i * n0;
    }
    for (int i = nxp + 1; i <= np[d]; i++) {
      w.cell_min[d][i - 1] = nxp * n0 + (i - nxp - 1) * (n0 + 1) + 1;
      w.cell_max[d][i - 1] = nxp * n0 + (i - nxp) * (n0 + 1);
    }
  }
  w.nranks = np[0] * np[1] * np[2];
  w.r.resize(w.nranks);
  for (int rk = 0; rk < w.nranks; rk++) {
    Rank &R = w.r[rk];
    R.rank = rk;
    // MPI_CART_CREATE with dims = (nprocz, nprocy, nprocx), row major:
    // x_coords varies fastest (mpi_routines.F90:186-187, 239-245)
    R.coords[0] = rk % np[0];
    R.coords[1] = (rk / np[0]) % np[1];
    R.coords[2] = rk / (np[0] * np[1]);
    for (int d = 0; d < 3; d++) {
      if (d < nd) {
        R.gmin[d] = w.cell_min[d][R.coords[d]];
        R.n[d] = w.cell_max[d][R.coords[d]] - R.gmin[d] + 1;
        R.is_bnd[2 * d] = (R.coords[d] == 0);
        R.is_bnd[2 * d + 1] = (R.coords[d] == np[d] - 1);
        // utilities.f90:349-365
        R.grid_min_local[d] = x_global(w, d, w.cell_min[d][R.coords[d]]);
        double hdx = 0.5 * w.d[d];
        R.min_local[d] = R.grid_min_local[d] - hdx;
        R.max_local[d] = x_global(w, d, w.cell_max[d][R.coords[d]] + 1) - hdx;
      }
    }
    // mpi_routines.F90:256-273
    for (int iz = -1; iz <= 1; iz++)
      for (int iy = -1; iy <= 1; iy++)
        for (int ix = -1; ix <= 1; ix++) {
          int t[3] = {R.coords[0] + ix, R.coords[1] + iy, R.coords[2] + iz};
          bool op = true;
          for (int d = 0; d < 3; d++) {
            if (d >= nd) { if (t[d] != 0) op = false; continue; }
            if (t[d] < 0 || t[d] >= np[d]) {
              if (!w.periods[d]) op = false;
              else t[d] = (t[d] + np[d]) % np[d];
            }
          }
          R.neighbour[iz + 1][iy + 1][ix + 1] = op ? (t[2] * np[1] + t[1]) * np[0] + t[0] : -1;
        }
    for (int i = 0; i < NFIELD; i++) R.f[i].init(R.n, nd);
    set_cpml_helpers(w, R);
    int pn[3] = {1, R.n[1], R.n[2]};
    for (int i = 0; i < 6; i++) {
      // planes (1-ng:ny+ng, 1-ng:nz+ng) stored with a unit x extent
      Arr &a = R.snap_min[i], &b = R.snap_max[i];
      for (Arr *q : {&a, &b}) {
        q->lo[0] = 1; q->sz[0] = 1;
        for (int d = 1; d < 3; d++) {
          if (d < nd) { q->lo[d] = 1 - NG; q->sz[d] = pn[d] + 2 * NG; }
          else { q->lo[d] = 1; q->sz[d] = 1; }
        }
        q->v.assign((size_t)q->sz[1] * q->sz[2], 0.0);
      }
    }
    for (int s = 0; s < 2; s++)
      for (Arr *q : {&R.src1[s], &R.src2[s]}) {
        *q = R.snap_min[0];
        std::fill(q->v.begin(), q->v.end(), 0.0);
      }
    for (int a = 1; a < nd; a++)
      for (int sd = 0; sd < 2; sd++) {
        auto shape = [&](Arr &q) {
          size_t tot = 1;
          for (int d = 0; d < 3; d++) {
            if (d == a || d >= nd) { q.lo[d] = 1; q.sz[d] = 1; }
            else { q.lo[d] = 1 - NG; q.sz[d] = R.n[d] + 2 * NG; }
            tot *= (size_t)q.sz[d];
          }
          q.v.assign(tot, 0.0);
        };
        for (int f = 0; f < 6; f++) shape(R.snapA[a][sd][f]);
        for (int q = 0; q < 2; q++) shape(R.srcA[a][sd][q]);
      }
    R.part.resize(w.sp.size());
    R.bnd_cand.resize(w.sp.size());
    R.inserted.resize(w.sp.size());
    // setup.F90:566-571
    R.rng.init(cf.seed + rk);
  }
}

inline int nbr(const Rank &R, int d, int s) {  // neighbour along one axis
  int o[3] = {0, 0, 0};
  o[d] = s;
  return R.neighbour[o[2] + 1][o[1] + 1][o[0] + 1];
}

// ---------------------------------------------------------------------------
// Ghost-cell exchange: boundary.F90:222-315 (2D), epoch3d boundary.F90:317-470,
// epoch1d boundary.F90:143-190
// ---------------------------------------------------------------------------
struct Box { int lo[3], hi[3]; };

inline Box full_box(const Arr &a) {
  Box b;
  for (int d = 0; d < 3; d++) { b.lo[d] = a.lo[d]; b.hi[d] = a.lo[d] + a.sz[d] - 1; }
  return b;
}

void copy_out(const Arr &a, const Box &b, std::vector<double> &t) {
  t.clear();
  for (int k = b.lo[2]; k <= b.hi[2]; k++)
    for (int j = b.lo[1]; j <= b.hi[1]; j++)
      for (int i = b.lo[0]; i <= b.hi[0]; i++) t.push_back(a(i, j, k));
}

void field_bc(World &w, int which) {
  const int nd = w.nd;
  std::vector<std::vector<double>> temp(w.nranks);
  for (int d = 0; d < nd; d++) {
    // pass 1: send low interior strip to proc_min, receive from proc_max into high ghosts
    // pass 2: send high interior strip to proc_max, receive from proc_min into low ghosts
    for (int pass = 0; pass < 2; pass++) {
      for (int rk = 0; rk < w.nranks; rk++) {
        Rank &R = w.r[rk];
        temp[rk].clear();
        int src = nbr(R, d, pass == 0 ? +1 : -1);
        if (src < 0) continue;
        const Rank &S = w.r[src];
        const Arr &a = S.f[which];
        Box b = full_box(a);
        if (pass == 0) { b.lo[d] = 1; b.hi[d] = NG; }
        else { b.lo[d] = S.n[d] + 1 - NG; b.hi[d] = S.n[d]; }
        copy_out(a, b, temp[rk]);
      }
      for (int rk = 0; rk < w.nranks; rk++) {
        Rank &R = w.r[rk];
        if (temp[rk].empty()) continue;
        int bd = (pass == 0) ? 2 * d + 1 : 2 * d;
        if (R.is_bnd[bd] && w.bc_field[bd] != c_bc_periodic) continue;
        Arr &a = R.f[which];
        Box b = full_box(a);
        if (pass == 0) { b.lo[d] = R.n[d] + 1; b.hi[d] = R.n[d] + NG; }
        else { b.lo[d] = 1 - NG; b.hi[d] = 0; }
        size_t n = 0;
        for (int k = b.lo[2]; k <= b.hi[2]; k++)
          for (int j = b.lo[1]; j <= b.hi[1]; j++)
            for (int i = b.lo[0]; i <= b.hi[0]; i++) a(i, j, k) = temp[rk][n++];
      }
    }
  }
}

// stagger(dir, field): setup.F90:124-134
inline bool stagger(int dir, int field) {
  switch (field) {
    case EX: return dir == 0;
    case EY: return dir == 1;
    case EZ: return dir == 2;
    case BX: return dir != 0;
    case BY: return dir != 1;
    case BZ: return dir != 2;
  }
  return false;
}

// boundary.F90:416-469 (sign=+1) and :473-530 (sign=-1, zero on the staggered plane)
void field_mirror(World &w, int which, int boundary, double sign) {
  if (w.bc_field[boundary] == c_bc_periodic) return;
  const int d = boundary / 2;
  if (d >= w.nd) return;
  const bool is_max = boundary & 1;
  for (Rank &R : w.r) {
    if (!R.is_bnd[boundary]) continue;
    Arr &a = R.f[which];
    Box b = full_box(a);
    auto plane_copy = [&](int dst, int src, bool zero) {
      Box q = b;
      q.lo[d] = q.hi[d] = dst;
      for (int k = q.lo[2]; k <= q.hi[2]; k++)
        for (int j = q.lo[1]; j <= q.hi[1]; j++)
          for (int i = q.lo[0]; i <= q.hi[0]; i++) {
            int s[3] = {i, j, k};
            s[d] = src;
            a(i, j, k) = zero ? 0.0 : sign * a(s[0], s[1], s[2]);
          }
    };
    const int nn = R.n[d];
    const bool clamp = sign < 0;
    if (!is_max) {
      if (stagger(d, which)) {
        for (int i = 1; i <= NG - 1; i++) plane_copy(i - NG, NG - i, false);
        if (clamp) plane_copy(0, 0, true);
      } else {
        for (int i = 1; i <= NG; i++) plane_copy(i - NG, NG + 1 - i, false);
      }
    } else {
      if (stagger(d, which)) {
        if (clamp) plane_copy(nn, nn, true);
        for (int i = 1; i <= NG - 1; i++) plane_copy(nn + i, nn - i, false);
      } else {
        for (int i = 1; i <= NG; i++) plane_copy(nn + i, nn + 1 - i, false);
      }
    }
  }
}

// boundary.F90:808-854 / :858-907
void field_bcs3(World &w, int f0, bool mpi_only) {
  for (int i = 0; i < 3; i++) field_bc(w, f0 + i);
  if (mpi_only) return;
  // perfectly conducting boundaries (boundary.F90:817-832 efield, :870-885 bfield; epoch3d adds the z pair,
  // epoch1d has the x pair only): E normal to the wall and B along it are clamped, the others zero-gradient
  for (int i = 0; i < 2 * w.nd; i++) {
    if (w.bc_field[i] != c_bc_conduct) continue;
    for (int q = 0; q < 3; q++) {
      const bool normal = (q == i / 2);
      const double sgn = (f0 == EX) ? (normal ? -1.0 : +1.0) : (normal ? +1.0 : -1.0);
      field_mirror(w, f0 + q, i, sgn);
    }
  }
  for (int i = 0; i < 2 * w.nd; i++) {
    int b = w.bc_field[i];
    if (b == c_bc_clamp || b == c_bc_simple_laser || b == c_bc_simple_outflow)
      for (int q = 0; q < 3; q++) field_mirror(w, f0 + q, i, -1.0);
    if (b == c_bc_zero_gradient || b == c_bc_cpml_laser || b == c_bc_cpml_outflow)   // boundary.F90:845-851, :898-904
      for (int q = 0; q < 3; q++) field_mirror(w, f0 + q, i, +1.0);
  }
}
void efield_bcs(World &w) { field_bcs3(w, EX, false); }
void bfield_bcs(World &w, bool mpi_only) { field_bcs3(w, BX, mpi_only); }

// ---------------------------------------------------------------------------
// FDTD: fields.f90:206-225 (E), :422-439 (B) for 2D; epoch3d fields.f90:312-337,
// :632-654; epoch1d fields.f90:150-166, :296-303.  Yee, order 2, no CPML.
// ---------------------------------------------------------------------------
// Orders 4 and 6 replace every c*(difference) term of the order-2 expressions by the group
// c1*c*(d1) [+ c2*c*(d2) [+ c3*c*(d3)]] with (c1,c2,c3) = (9/8,-1/24) resp. (75/64,-25/384,3/640),
// added one after the other in the textual order of fields.f90:128-204 (2D), epoch3d
// fields.f90:337-430, epoch1d fields.f90:166-215.  nt = field_order / 2.
static inline void fd_coeffs(int order, double base, double *cc) {
  if (order == 4) { cc[0] = (9.0 / 8.0) * base; cc[1] = (-1.0 / 24.0) * base; cc[2] = 0.0; }
  else if (order == 6) { cc[0] = (75.0 / 64.0) * base; cc[1] = (-25.0 / 384.0) * base; cc[2] = (3.0 / 640.0) * base; }
  else { cc[0] = base; cc[1] = 0.0; cc[2] = 0.0; }
}

// cpml_advance_e_currents / cpml_advance_b_currents (boundary.F90:1813-2023; epoch3d :2365-2790, epoch1d :929-1040),
// one routine for every axis a with (b, c) the cyclic successors: in the layer
//   psi_Eb = bcoeff psi_Eb + ccoeff_d (B_c(i) - B_c(i-1)),  E_b -= fac psi_Eb;   psi_Ec likewise from B_b,  E_c += fac psi_Ec
//   psi_Bb = bcoeff psi_Bb + ccoeff_d (E_c(i+1) - E_c(i)),  B_b += tstep psi_Bb; psi_Bc likewise from E_b,  B_c -= tstep psi_Bc
// over the interior of the other axes; the B form of a max layer runs one point lower (start-1 .. end-1).
template <int ND>
void cpml_advance_currents(World &w, double tstep, bool efield) {
  const double fac = tstep * (c * c);
  for (Rank &R : w.r)
    for (int a = 0; a < ND; a++) {
      const int b = (a + 1) % 3, cc = (a + 2) % 3;
      for (int sd = 0; sd < 2; sd++) {
        if (!is_cpml(w.bc_field[2 * a + sd])) continue;
        int i0 = R.cp_start[a][sd], i1 = R.cp_end[a][sd];
        if (!efield && sd == 1) { i0 -= 1; i1 -= 1; }
        const std::vector<double> &kap = efield ? R.kap_e[a] : R.kap_b[a];
        const std::vector<double> &sig = efield ? R.sig_e[a] : R.sig_b[a];
        const std::vector<double> &aa = efield ? R.a_e[a] : R.a_b[a];
        Arr &Fb = R.f[(efield ? EX : BX) + b], &Fc = R.f[(efield ? EX : BX) + cc];
        const Arr &Gb = R.f[(efield ? BX : EX) + b], &Gc = R.f[(efield ? BX : EX) + cc];
        Arr &psb = R.psi[a][efield ? 0 : 2], &psc = R.psi[a][efield ? 1 : 3];
        int lo[3] = {1, 1, 1}, hi[3] = {1, 1, 1};
        for (int d = 0; d < ND; d++) if (d != a) hi[d] = R.n[d];
        int e[3] = {0, 0, 0};
        e[a] = 1;
        for (int ipos = i0; ipos <= i1; ipos++) {
          const double kappa = kap[ipos - (1 - NG)], sigma = sig[ipos - (1 - NG)], acoeff = aa[ipos - (1 - NG)];
          const double bcoeff = std::exp(-(sigma / kappa + acoeff) * tstep);
          const double ccoeff_d = (bcoeff - 1.0) * sigma / kappa / (sigma + kappa * acoeff) / w.d[a];
          for (int k = lo[2]; k <= hi[2]; k++)
            for (int j = lo[1]; j <= hi[1]; j++)
              for (int i = lo[0]; i <= hi[0]; i++) {
                int p[3] = {i, j, k};
                p[a] = ipos;
                const int q0 = p[0], q1 = p[1], q2 = p[2];
                if (efield) {
                  const int m0 = q0 - e[0], m1 = q1 - e[1], m2 = q2 - e[2];
                  psb(q0, q1, q2) = bcoeff * psb(q0, q1, q2) + ccoeff_d * (Gc(q0, q1, q2) - Gc(m0, m1, m2));
                  Fb(q0, q1, q2) = Fb(q0, q1, q2) - fac * psb(q0, q1, q2);
                  psc(q0, q1, q2) = bcoeff * psc(q0, q1, q2) + ccoeff_d * (Gb(q0, q1, q2) - Gb(m0, m1, m2));
                  Fc(q0, q1, q2) = Fc(q0, q1, q2) + fac * psc(q0, q1, q2);
                } else {
                  const int u0 = q0 + e[0], u1 = q1 + e[1], u2 = q2 + e[2];
                  psb(q0, q1, q2) = bcoeff * psb(q0, q1, q2) + ccoeff_d * (Gc(u0, u1, u2) - Gc(q0, q1, q2));
                  Fb(q0, q1, q2) = Fb(q0, q1, q2) + tstep * psb(q0, q1, q2);
                  psc(q0, q1, q2) = bcoeff * psc(q0, q1, q2) + ccoeff_d * (Gb(u0, u1, u2) - Gb(q0, q1, q2));
                  Fc(q0, q1, q2) = Fc(q0, q1, q2) - tstep * psc(q0, q1, q2);
                }
              }
        }
      }
    }
}

template <int ND>
void update_e_field(World &w, double hdt) {
  const double cnx = hdt / w.d[0] * (c * c);
  const double cny = ND >= 2 ? hdt / w.d[1] * (c * c) : 0.0;
  const double cnz = ND >= 3 ? hdt / w.d[2] * (c * c) : 0.0;
  const double fac = hdt / epsilon0;
  const int order = w.cfg.field_order ? w.cfg.field_order : 2;
  const int nt = order / 2;
  double cx[3], cy[3], cz[3];
  fd_coeffs(order, cnx, cx);
  fd_coeffs(order, cny, cy);
  fd_coeffs(order, cnz, cz);
  const bool cpml = w.cpml_t > 0;
  for (Rank &R : w.r) {
    Arr &ex = R.f[EX], &ey = R.f[EY], &ez = R.f[EZ];
    const Arr &bx = R.f[BX], &by = R.f[BY], &bz = R.f[BZ];
    const Arr &jx = R.f[JX], &jy = R.f[JY], &jz = R.f[JZ];
    const int k0 = ND >= 3 ? 0 : 1, k1 = ND >= 3 ? R.n[2] : 1;
    const int j0 = ND >= 2 ? 0 : 1, j1 = ND >= 2 ? R.n[1] : 1;
    for (int iz = k0; iz <= k1; iz++)
      for (int iy = j0; iy <= j1; iy++)
        for (int ix = 0; ix <= R.n[0]; ix++) {
          if (cpml) {   // fields.f90:112-204: cpml_x = cnx / cpml_kappa_ex(ix), cx1 = c1 * cpml_x, ...
            fd_coeffs(order, cnx / R.kap_e[0][ix - (1 - NG)], cx);
            if (ND >= 2) fd_coeffs(order, cny / R.kap_e[1][iy - (1 - NG)], cy);
            if (ND >= 3) fd_coeffs(order, cnz / R.kap_e[2][iz - (1 - NG)], cz);
          }
          // backward differences: term k is f(i+k) - f(i-k-1)
          auto dxb = [&](const Arr &f, int k) {
            return ND == 1 ? f(ix + k) - f(ix - k - 1)
                 : ND == 2 ? f(ix + k, iy) - f(ix - k - 1, iy) : f(ix + k, iy, iz) - f(ix - k - 1, iy, iz);
          };
          auto dyb = [&](const Arr &f, int k) {
            return ND == 2 ? f(ix, iy + k) - f(ix, iy - k - 1) : f(ix, iy + k, iz) - f(ix, iy - k - 1, iz);
          };
          auto dzb = [&](const Arr &f, int k) { return f(ix, iy, iz + k) - f(ix, iy, iz - k - 1); };
          if (ND == 1) {
            double v = ex(ix);
            ex(ix) = v - fac * jx(ix);
            v = ey(ix);
            for (int k = 0; k < nt; k++) v = v - cx[k] * dxb(bz, k);
            ey(ix) = v - fac * jy(ix);
            v = ez(ix);
            for (int k = 0; k < nt; k++) v = v + cx[k] * dxb(by, k);
            ez(ix) = v - fac * jz(ix);
          } else if (ND == 2) {
            double v = ex(ix, iy);
            for (int k = 0; k < nt; k++) v = v + cy[k] * dyb(bz, k);
            ex(ix, iy) = v - fac * jx(ix, iy);
            v = ey(ix, iy);
            for (int k = 0; k < nt; k++) v = v - cx[k] * dxb(bz, k);
            ey(ix, iy) = v - fac * jy(ix, iy);
            v = ez(ix, iy);
            for (int k = 0; k < nt; k++) v = v + cx[k] * dxb(by, k);
            for (int k = 0; k < nt; k++) v = v - cy[k] * dyb(bx, k);
            ez(ix, iy) = v - fac * jz(ix, iy);
          } else {
            double v = ex(ix, iy, iz);
            for (int k = 0; k < nt; k++) v = v + cy[k] * dyb(bz, k);
            for (int k = 0; k < nt; k++) v = v - cz[k] * dzb(by, k);
            ex(ix, iy, iz) = v - fac * jx(ix, iy, iz);
            v = ey(ix, iy, iz);
            for (int k = 0; k < nt; k++) v = v + cz[k] * dzb(bx, k);
            for (int k = 0; k < nt; k++) v = v - cx[k] * dxb(bz, k);
            ey(ix, iy, iz) = v - fac * jy(ix, iy, iz);
            v = ez(ix, iy, iz);
            for (int k = 0; k < nt; k++) v = v + cx[k] * dxb(by, k);
            for (int k = 0; k < nt; k++) v = v - cy[k] * dyb(bx, k);
            ez(ix, iy, iz) = v - fac * jz(ix, iy, iz);
          }
        }
  }
  if (cpml) cpml_advance_currents<ND>(w, hdt, true);   // fields.f90:204
}

template <int ND>
void update_b_field(World &w, double hdt) {
  const double hdtx = hdt / w.d[0];
  const double hdty = ND >= 2 ? hdt / w.d[1] : 0.0;
  const double hdtz = ND >= 3 ? hdt / w.d[2] : 0.0;
  const int order = w.cfg.field_order ? w.cfg.field_order : 2;
  const int nt = order / 2;
  double cx[3], cy[3], cz[3];
  fd_coeffs(order, hdtx, cx);
  fd_coeffs(order, hdty, cy);
  fd_coeffs(order, hdtz, cz);
  const bool ext = w.cfg.maxwell_solver != 0;  // fields.f90:441-465, epoch3d :655-730, epoch1d :304-312
  const Config &cf = w.cfg;
  const bool cpml = w.cpml_t > 0;
  for (Rank &R : w.r) {
    const Arr &ex = R.f[EX], &ey = R.f[EY], &ez = R.f[EZ];
    Arr &bx = R.f[BX], &by = R.f[BY], &bz = R.f[BZ];
    const int k0 = ND >= 3 ? 0 : 1, k1 = ND >= 3 ? R.n[2] : 1;
    const int j0 = ND >= 2 ? 0 : 1, j1 = ND >= 2 ? R.n[1] : 1;
    for (int iz = k0; iz <= k1; iz++)
      for (int iy = j0; iy <= j1; iy++)
        for (int ix = 0; ix <= R.n[0]; ix++) {
          double hx_ = hdtx, hy_ = hdty, hz_ = hdtz;
          if (cpml) {   // fields.f90:306-420: cx1 = hdtx / cpml_kappa_bx(ix) (times c1.. for the higher orders)
            hx_ = hdtx / R.kap_b[0][ix - (1 - NG)];
            fd_coeffs(order, hx_, cx);
            if (ND >= 2) { hy_ = hdty / R.kap_b[1][iy - (1 - NG)]; fd_coeffs(order, hy_, cy); }
            if (ND >= 3) { hz_ = hdtz / R.kap_b[2][iz - (1 - NG)]; fd_coeffs(order, hz_, cz); }
          }
          // forward differences: term k is f(i+k+1) - f(i-k)
          auto dxf = [&](const Arr &f, int k) {
            return ND == 1 ? f(ix + k + 1) - f(ix - k)
                 : ND == 2 ? f(ix + k + 1, iy) - f(ix - k, iy) : f(ix + k + 1, iy, iz) - f(ix - k, iy, iz);
          };
          auto dyf = [&](const Arr &f, int k) {
            return ND == 2 ? f(ix, iy + k + 1) - f(ix, iy - k) : f(ix, iy + k + 1, iz) - f(ix, iy - k, iz);
          };
          auto dzf = [&](const Arr &f, int k) { return f(ix, iy, iz + k + 1) - f(ix, iy, iz - k); };
          if (ND == 1 && !ext) {
            double v = by(ix);
            for (int k = 0; k < nt; k++) v = v + cx[k] * dxf(ez, k);
            by(ix) = v;
            v = bz(ix);
            for (int k = 0; k < nt; k++) v = v - cx[k] * dxf(ey, k);
            bz(ix) = v;
          } else if (ext) {
            // derivative of f along axis a at the staggered point: alpha, the two betas (lower other axis
            // first; +1 then -1), gamma (3D; first other axis + then -, second other axis - then +), delta
            auto F = [&](const Arr &f, int dx_, int dy_, int dz_) {
              return ND == 1 ? f(ix + dx_) : ND == 2 ? f(ix + dx_, iy + dy_) : f(ix + dx_, iy + dy_, iz + dz_);
            };
            auto dd = [&](const Arr &f, int a) {
              int e[3] = {0, 0, 0};
              e[a] = 1;
              auto df = [&](int ox, int oy, int oz) {  // f(+a) - f(0), both shifted by (ox, oy, oz)
                return F(f, e[0] + ox, e[1] + oy, e[2] + oz) - F(f, ox, oy, oz);
              };
              double v = cf.st_alpha[a] * df(0, 0, 0);
              int oth[2], no = 0;
              for (int d = 0; d < 3; d++) if (d != a) oth[no++] = d;
              for (int k = 0; k < 2; k++) {
                const int b_ = oth[k];
                if (b_ >= ND) continue;
                int p[3] = {0, 0, 0}, m[3] = {0, 0, 0};
                p[b_] = 1; m[b_] = -1;
                // (f(+a,+b) - f(0,+b) + f(+a,-b)) - f(0,-b), left to right
                const double t = F(f, e[0] + p[0], e[1] + p[1], e[2] + p[2]) - F(f, p[0], p[1], p[2]) +
                                 F(f, e[0] + m[0], e[1] + m[1], e[2] + m[2]) - F(f, m[0], m[1], m[2]);
                v = v + cf.st_beta[2 * a + k] * t;
              }
              if (ND == 3) {
                const int b_ = oth[0], c_ = oth[1];
                double t = 0.0;
                bool first = true;
                for (int sc = -1; sc <= 1; sc += 2)
                  for (int sb = 1; sb >= -1; sb -= 2) {
                    int o[3] = {0, 0, 0};
                    o[b_] = sb; o[c_] = sc;
                    const double hi = F(f, e[0] + o[0], e[1] + o[1], e[2] + o[2]), lo = F(f, o[0], o[1], o[2]);
                    if (first) { t = hi - lo; first = false; }
                    else t = t + hi - lo;
                  }
                v = v + cf.st_gamma[a] * t;
              }
              v = v + cf.st_delta[a] * (F(f, 2 * e[0], 2 * e[1], 2 * e[2]) - F(f, -e[0], -e[1], -e[2]));
              return v;
            };
            if (ND == 1) {
              by(ix) = by(ix) + hx_ * dd(ez, 0);
              bz(ix) = bz(ix) - hx_ * dd(ey, 0);
            } else if (ND == 2) {
              bx(ix, iy) = bx(ix, iy) - hy_ * dd(ez, 1);
              by(ix, iy) = by(ix, iy) + hx_ * dd(ez, 0);
              bz(ix, iy) = bz(ix, iy) - hx_ * dd(ey, 0) + hy_ * dd(ex, 1);
            } else {
              bx(ix, iy, iz) = bx(ix, iy, iz) - hy_ * dd(ez, 1) + hz_ * dd(ey, 2);
              by(ix, iy, iz) = by(ix, iy, iz) - hz_ * dd(ex, 2) + hx_ * dd(ez, 0);
              bz(ix, iy, iz) = bz(ix, iy, iz) - hx_ * dd(ey, 0) + hy_ * dd(ex, 1);
            }
          } else if (ND == 2) {
            double v = bx(ix, iy);
            for (int k = 0; k < nt; k++) v = v - cy[k] * dyf(ez, k);
            bx(ix, iy) = v;
            v = by(ix, iy);
            for (int k = 0; k < nt; k++) v = v + cx[k] * dxf(ez, k);
            by(ix, iy) = v;
            v = bz(ix, iy);
            for (int k = 0; k < nt; k++) v = v - cx[k] * dxf(ey, k);
            for (int k = 0; k < nt; k++) v = v + cy[k] * dyf(ex, k);
            bz(ix, iy) = v;
          } else {
            double v = bx(ix, iy, iz);
            for (int k = 0; k < nt; k++) v = v - cy[k] * dyf(ez, k);
            for (int k = 0; k < nt; k++) v = v + cz[k] * dzf(ey, k);
            bx(ix, iy, iz) = v;
            v = by(ix, iy, iz);
            for (int k = 0; k < nt; k++) v = v - cz[k] * dzf(ex, k);
            for (int k = 0; k < nt; k++) v = v + cx[k] * dxf(ez, k);
            by(ix, iy, iz) = v;
            v = bz(ix, iy, iz);
            for (int k = 0; k < nt; k++) v = v - cx[k] * dxf(ey, k);
            for (int k = 0; k < nt; k++) v = v + cy[k] * dyf(ex, k);
            bz(ix, iy, iz) = v;
          }
        }
  }
  if (cpml) cpml_advance_currents<ND>(w, hdt, false);   // fields.f90:420 region: after the B update
}

// ---------------------------------------------------------------------------
// Laser / outflow boundary on x_min, x_max: laser.f90:310-458 (2D),
// epoch3d laser.f90:350-506, epoch1d laser.f90:260-392.
// source1/source2 are evaluated by the host (deck expressions) and stored in
// R.src1/src2 before the call.
// ---------------------------------------------------------------------------
template <int ND>
void outflow_bcs_x(World &w, bool is_max, double dt) {
  const double dtc2 = dt * (c * c);
  const double lx = dtc2 / w.d[0];
  const double ly = ND >= 2 ? dtc2 / w.d[1] : 0.0;
  const double lz = ND >= 3 ? dtc2 / w.d[2] : 0.0;
  const double sum = 1.0 / (lx + c);
  const double diff = lx - c;
  const double dt_eps = dt / epsilon0;
  const bool cpml_face = w.bc_field[is_max ? 1 : 0] == c_bc_cpml_laser;
  for (Rank &R : w.r) {
    if (!R.is_bnd[is_max ? 1 : 0]) continue;
    // a cpml_laser face: only where the laser plane lies on this rank (add_laser, boundary.F90:1572-1577), and the
    // plane is cpml_x_min_laser_idx / cpml_x_max_laser_idx instead of 1 / nx (laser.f90:320-323, :396-399)
    if (cpml_face && !R.add_laser_cpml[0][is_max ? 1 : 0]) continue;
    Arr &bx = R.f[BX], &by = R.f[BY], &bz = R.f[BZ];
    const Arr &ey = R.f[EY], &ez = R.f[EZ], &jy = R.f[JY], &jz = R.f[JZ];
    const Arr *snap = is_max ? R.snap_max : R.snap_min;
    const Arr &s1 = R.src1[is_max ? 1 : 0], &s2 = R.src2[is_max ? 1 : 0];
    const int k0 = ND >= 3 ? 0 : 1, k1 = ND >= 3 ? R.n[2] : 1;
    const int j0 = ND >= 2 ? 0 : 1, j1 = ND >= 2 ? R.n[1] : 1;
    const int nx = cpml_face ? R.laser_idx[0][1] : R.n[0];
    const int lp_min = cpml_face ? R.laser_idx[0][0] : 1;
    // all right-hand sides use pre-update values (Fortran array assignment);
    // bz is written before by is evaluated but by's RHS never reads bz.
    for (int k = k0; k <= k1; k++)
      for (int j = j0; j <= j1; j++) {
        if (!is_max) bx(lp_min - 1, j, k) = snap[BX](1, j, k);
        else bx(nx + 1, j, k) = snap[BX](1, j, k);
      }
    for (int k = k0; k <= k1; k++)
      for (int j = j0; j <= j1; j++) {
        if (!is_max) {
          const int lp = lp_min;
          double t = 4.0 * s1(1, j, k) + 2.0 * (snap[EY](1, j, k) + c * snap[BZ](1, j, k)) -
                     2.0 * ey(lp, j, k);
          if (ND == 3) t = t - lz * (bx(lp, j, k) - bx(lp, j, k - 1));
          t = t + dt_eps * jy(lp, j, k) + diff * bz(lp, j, k);
          bz(lp - 1, j, k) = sum * t;
        } else {
          const int lp = nx;
          double t = -4.0 * s1(1, j, k) - 2.0 * (snap[EY](1, j, k) - c * snap[BZ](1, j, k)) +
                     2.0 * ey(lp, j, k);
          if (ND == 3) t = t + lz * (bx(lp, j, k) - bx(lp, j, k - 1));
          t = t - dt_eps * jy(lp, j, k) + diff * bz(lp - 1, j, k);
          bz(lp, j, k) = sum * t;
        }
      }
    for (int k = k0; k <= k1; k++)
      for (int j = j0; j <= j1; j++) {
        if (!is_max) {
          const int lp = lp_min;
          double t = -4.0 * s2(1, j, k) - 2.0 * (snap[EZ](1, j, k) - c * snap[BY](1, j, k)) +
                     2.0 * ez(lp, j, k);
          if (ND >= 2) t = t - ly * (bx(lp, j, k) - bx(lp, j - 1, k));
          t = t - dt_eps * jz(lp, j, k) + diff * by(lp, j, k);
          by(lp - 1, j, k) = sum * t;
        } else {
          const int lp = nx;
          double t = 4.0 * s2(1, j, k) + 2.0 * (snap[EZ](1, j, k) + c * snap[BY](1, j, k)) -
                     2.0 * ez(lp, j, k);
          if (ND >= 2) t = t + ly * (bx(lp, j, k) - bx(lp, j - 1, k));
          t = t + dt_eps * jz(lp, j, k) + diff * by(lp - 1, j, k);
          by(lp, j, k) = sum * t;
        }
      }
  }
}

// setup.F90:391-447 (x boundaries only)
void setup_field_boundaries(World &w) {
  for (Rank &R : w.r) {
    // setup.F90:409-412: a cpml_laser face takes its snapshot one plane outside the laser plane
    int nx0 = 1, nx1 = R.n[0];
    if (w.bc_field[0] == c_bc_cpml_laser) nx0 = R.laser_idx[0][0] - 1;
    if (w.bc_field[1] == c_bc_cpml_laser) nx1 = R.laser_idx[0][1] + 1;
    for (int f = 0; f < 6; f++) {
      const Arr &a = R.f[f];
      Arr &lo = R.snap_min[f], &hi = R.snap_max[f];
      const bool avg = (f == EX || f == BY || f == BZ);
      for (int k = lo.lo[2]; k < lo.lo[2] + lo.sz[2]; k++)
        for (int j = lo.lo[1]; j < lo.lo[1] + lo.sz[1]; j++) {
          lo(1, j, k) = avg ? 0.5 * (a(nx0, j, k) + a(nx0 - 1, j, k)) : a(nx0, j, k);
          hi(1, j, k) = avg ? 0.5 * (a(nx1, j, k) + a(nx1 - 1, j, k)) : a(nx1, j, k);
        }
    }
    // y and z faces (setup.F90:430-447; epoch3d setup.F90:431-500): components staggered along the
    // face normal (E normal, B transverse) are averaged across the boundary
    for (int ax = 1; ax < w.nd; ax++)
      for (int sd = 0; sd < 2; sd++)
        for (int f = 0; f < 6; f++) {
          const Arr &a = R.f[f];
          Arr &q = R.snapA[ax][sd][f];
          const bool avg = f < 3 ? (f == ax) : (f - 3 != ax);
          int n0 = sd == 0 ? 1 : R.n[ax];
          if (w.bc_field[2 * ax + sd] == c_bc_cpml_laser) n0 = sd == 0 ? R.laser_idx[ax][0] - 1 : R.laser_idx[ax][1] + 1;
          int e[3] = {0, 0, 0};
          e[ax] = 1;
          for (int k = q.lo[2]; k < q.lo[2] + q.sz[2]; k++)
            for (int j = q.lo[1]; j < q.lo[1] + q.sz[1]; j++)
              for (int i = q.lo[0]; i < q.lo[0] + q.sz[0]; i++) {
                int p[3] = {i, j, k};
                p[ax] = n0;
                const double v0 = a(p[0], p[1], p[2]);
                q(i, j, k) = avg ? 0.5 * (v0 + a(p[0] - e[0], p[1] - e[1], p[2] - e[2])) : v0;
              }
        }
  }
}

// outflow_bcs_{y,z}_{min,max} (laser.f90:462-610; epoch3d laser.f90:510-830): the x-face formulas under
// the cyclic permutation x -> y -> z of axes and components (a = normal, b = a+1, cc = a+2).
template <int ND>
void outflow_bcs_axis(World &w, int a, bool is_max, double dt) {
  const int b = (a + 1) % 3, cc = (a + 2) % 3;
  const double dtc2 = dt * (c * c);
  double l[3];
  for (int d = 0; d < 3; d++) l[d] = d < ND ? dtc2 / w.d[d] : 0.0;
  const double sum = 1.0 / (l[a] + c);
  const double diff = l[a] - c;
  const double dt_eps = dt / epsilon0;
  const bool cpml_face = w.bc_field[2 * a + (is_max ? 1 : 0)] == c_bc_cpml_laser;
  for (Rank &R : w.r) {
    if (!R.is_bnd[2 * a + (is_max ? 1 : 0)]) continue;
    if (cpml_face && !R.add_laser_cpml[a][is_max ? 1 : 0]) continue;
    Arr &Ba = R.f[BX + a], &Bb = R.f[BX + b], &Bc = R.f[BX + cc];
    const Arr &Eb = R.f[EX + b], &Ec = R.f[EX + cc], &Jb = R.f[JX + b], &Jc = R.f[JX + cc];
    const Arr *snap = R.snapA[a][is_max ? 1 : 0];
    const Arr &s1 = R.srcA[a][is_max ? 1 : 0][0], &s2 = R.srcA[a][is_max ? 1 : 0][1];
    int lo[3] = {1, 1, 1}, hi[3] = {1, 1, 1};
    for (int d = 0; d < ND; d++) if (d != a) { lo[d] = 0; hi[d] = R.n[d]; }
    const int lp = cpml_face ? R.laser_idx[a][is_max ? 1 : 0] : (is_max ? R.n[a] : 1);
    int ea[3] = {0, 0, 0}, eb[3] = {0, 0, 0}, ec[3] = {0, 0, 0};
    ea[a] = 1; eb[b] = 1; ec[cc] = 1;
    auto at = [&](const Arr &A, const int *p, const int *e, int m) {
      return A(p[0] + m * e[0], p[1] + m * e[1], p[2] + m * e[2]);
    };
    auto pl = [&](const Arr &A, const int *p) {  // plane array: unit extent along a
      int q[3] = {p[0], p[1], p[2]};
      q[a] = 1;
      return A(q[0], q[1], q[2]);
    };
    // the normal component first (array assignment in the reference), then the two transverse ones;
    // their right-hand sides only read the laserpos plane of B_a, never the one just written
    for (int pass = 0; pass < 3; pass++)
      for (int k = lo[2]; k <= hi[2]; k++)
        for (int j = lo[1]; j <= hi[1]; j++)
          for (int i = lo[0]; i <= hi[0]; i++) {
            int p[3] = {i, j, k};
            p[a] = lp;
            int pw[3] = {p[0], p[1], p[2]};       // the plane that is written
            pw[a] = is_max ? lp + 1 : lp - 1;
            if (pass == 0) {
              Ba(pw[0], pw[1], pw[2]) = pl(snap[BX + a], p);
            } else if (pass == 1) {
              if (!is_max) {
                double t = 4.0 * pl(s1, p) + 2.0 * (pl(snap[EX + b], p) + c * pl(snap[BX + cc], p)) - 2.0 * at(Eb, p, ea, 0);
                if (cc < ND) t = t - l[cc] * (at(Ba, p, ec, 0) - at(Ba, p, ec, -1));
                t = t + dt_eps * at(Jb, p, ea, 0) + diff * at(Bc, p, ea, 0);
                Bc(pw[0], pw[1], pw[2]) = sum * t;
              } else {
                double t = -4.0 * pl(s1, p) - 2.0 * (pl(snap[EX + b], p) - c * pl(snap[BX + cc], p)) + 2.0 * at(Eb, p, ea, 0);
                if (cc < ND) t = t + l[cc] * (at(Ba, p, ec, 0) - at(Ba, p, ec, -1));
                t = t - dt_eps * at(Jb, p, ea, 0) + diff * at(Bc, p, ea, -1);
                Bc(p[0], p[1], p[2]) = sum * t;
              }
            } else {
              if (!is_max) {
                double t = -4.0 * pl(s2, p) - 2.0 * (pl(snap[EX + cc], p) - c * pl(snap[BX + b], p)) + 2.0 * at(Ec, p, ea, 0);
                if (b < ND) t = t - l[b] * (at(Ba, p, eb, 0) - at(Ba, p, eb, -1));
                t = t - dt_eps * at(Jc, p, ea, 0) + diff * at(Bb, p, ea, 0);
                Bb(pw[0], pw[1], pw[2]) = sum * t;
              } else {
                double t = 4.0 * pl(s2, p) + 2.0 * (pl(snap[EX + cc], p) + c * pl(snap[BX + b], p)) - 2.0 * at(Ec, p, ea, 0);
                if (b < ND) t = t + l[b] * (at(Ba, p, eb, 0) - at(Ba, p, eb, -1));
                t = t + dt_eps * at(Jc, p, ea, 0) + diff * at(Bb, p, ea, -1);
                Bb(p[0], p[1], p[2]) = sum * t;
              }
            }
          }
  }
}

// boundary.F90:911-944
template <int ND>
void bfield_final_bcs(World &w, double dt) {
  bfield_bcs(w, false);
  // add_laser(i) .OR. simple_outflow (boundary.F90:918-940): a cpml_laser face has add_laser on the rank that holds
  // the laser plane (the outflow routines skip the others); a cpml_outflow face only absorbs
  for (int s = 0; s < 2; s++) {
    int b = w.bc_field[s];
    if (b == c_bc_simple_laser || b == c_bc_simple_outflow || b == c_bc_cpml_laser) outflow_bcs_x<ND>(w, s == 1, dt);
  }
  for (int s = 2; s < 2 * ND; s++) {
    int b = w.bc_field[s];
    if (b == c_bc_simple_laser || b == c_bc_simple_outflow || b == c_bc_cpml_laser)
      outflow_bcs_axis<ND>(w, s / 2, (s & 1) == 1, dt);
  }
  bfield_bcs(w, true);
}

// fields.f90:533-582
template <int ND>
void update_eb_fields_half(World &w) {
  const double hdt = 0.5 * w.dt;
  update_e_field<ND>(w, hdt);
  efield_bcs(w);
  update_b_field<ND>(w, hdt);
  bfield_bcs(w, true);
}
template <int ND>
void update_eb_fields_final(World &w) {
  const double hdt = 0.5 * w.dt;
  update_b_field<ND>(w, hdt);
  bfield_final_bcs<ND>(w, w.dt);
  update_e_field<ND>(w, hdt);
  efield_bcs(w);
}

// ---------------------------------------------------------------------------
// Particle push: particles.F90:28-650 (2D), epoch3d particles.F90, epoch1d
// particles.F90.  Triangle shape: include/triangle/{gx,hx_dcell,e_part,b_part}.inc
// ---------------------------------------------------------------------------
inline void tri_weights(double f, double *g) {  // g points at index 0
  // include/triangle/gx.inc:1-4
  double cf2 = f * f;
  g[-1] = 0.25 + cf2 + f;
  g[0] = 1.5 - 2.0 * cf2;
  g[1] = 0.25 + cf2 - f;
}

template <int ND>
inline double gather(const Arr &F, const double *wx, int cx, const double *wy, int cy,
                     const double *wz, int cz) {
  // include/triangle/e_part.inc: rows are parenthesised, sums left to right
  if (ND == 1) {
    return wx[-1] * F(cx - 1) + wx[0] * F(cx) + wx[1] * F(cx + 1);
  } else if (ND == 2) {
    double r = 0.0;
    for (int iy = -1; iy <= 1; iy++) {
      double row = wx[-1] * F(cx - 1, cy + iy) + wx[0] * F(cx, cy + iy) + wx[1] * F(cx + 1, cy + iy);
      double t = wy[iy] * row;
      r = (iy == -1) ? t : r + t;
    }
    return r;
  } else {
    double r = 0.0;
    for (int iz = -1; iz <= 1; iz++) {
      double pl = 0.0;
      for (int iy = -1; iy <= 1; iy++) {
        double row = wx[-1] * F(cx - 1, cy + iy, cz + iz) + wx[0] * F(cx, cy + iy, cz + iz) +
                     wx[1] * F(cx + 1, cy + iy, cz + iz);
        double t = wy[iy] * row;
        pl = (iy == -1) ? t : pl + t;
      }
      double t = wz[iz] * pl;
      r = (iz == -1) ? t : r + t;
    }
    return r;
  }
}

template <int ND>
void push_particles(World &w, int only = -1, bool zero_j = true) {
  const double dt = w.dt;
  for (Rank &R : w.r) {
    if (zero_j) for (int q = JX; q <= JZ; q++) std::fill(R.f[q].v.begin(), R.f[q].v.end(), 0.0);
    // particles.F90:128-136, 155-167
    double fac = 1.0;
    for (int d = 0; d < ND; d++) fac *= 0.5;
    const double idx = 1.0 / w.d[0];
    const double idy = ND >= 2 ? 1.0 / w.d[1] : 0.0;
    const double idz = ND >= 3 ? 1.0 / w.d[2] : 0.0;
    const double idt = 1.0 / dt;
    const double dto2 = dt / 2.0;
    const double dtco2 = c * dto2;
    const double dtfac = 0.5 * dt * fac;
    const double third = 1.0 / 3.0;
    // 2D: idty, idtx, idxy; 3D: idtyz, idtxz, idtxy; 1D: idtf, idxf
    double k_fcx, k_fcy, k_fcz;
    if (ND == 1) { k_fcx = idt * fac; k_fcy = idx * fac; k_fcz = 0.0; }
    else if (ND == 2) { k_fcx = idt * idy * fac; k_fcy = idt * idx * fac; k_fcz = idx * idy * fac; }
    else { k_fcx = idt * idy * idz * fac; k_fcy = idt * idx * idz * fac; k_fcz = idt * idx * idy * fac; }

    const Arr &ex = R.f[EX], &ey = R.f[EY], &ez = R.f[EZ];
    const Arr &bx = R.f[BX], &by = R.f[BY], &bz = R.f[BZ];
    Arr &jx = R.f[JX], &jy = R.f[JY], &jz = R.f[JZ];

    for (size_t is = 0; is < w.sp.size(); is++) {
      if (only >= 0 && (int)is != only) continue;
      const SpeciesCfg &S = w.sp[is];
      R.bnd_cand[is].clear();
      if (S.immobile) continue;
      // particles.F90:189-221 (no thermal / cpml): candidate bounds
      double bnd_min[3], bnd_max[3];
      for (int d = 0; d < ND; d++) { bnd_min[d] = R.min_local[d]; bnd_max[d] = R.max_local[d]; }
      // particles.F90:251-256
      const double part_q = S.charge;
      const double part_mc = c * S.mass;
      const double ipart_mc = 1.0 / part_mc;
      const double cmratio = part_q * dtfac * ipart_mc;
      const double ccmratio = c * cmratio;
      const bool deposit = !S.zero_current;

      // gx,gy,gz: entries sf_min-1 and sf_max+1 stay zero (particles.F90:152-153)
      double G[3][5] = {{0}}, H[3][5];
      std::vector<Particle> &pl = R.part[is];
      for (size_t ip = 0; ip < pl.size(); ip++) {
        Particle &P = pl[ip];
        const double part_weight = P.w;
        const double fcx = k_fcx * part_weight;
        const double fcy = k_fcy * part_weight;
        const double fcz = k_fcz * part_weight;
        // :289-302
        double part_pos[3];
        for (int d = 0; d < ND; d++) part_pos[d] = P.pos[d] - R.grid_min_local[d];
        double part_ux = P.p[0] * ipart_mc;
        double part_uy = P.p[1] * ipart_mc;
        double part_uz = P.p[2] * ipart_mc;
        double gamma_rel = std::sqrt(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0);
        double root = dtco2 / gamma_rel;
        const double u[3] = {part_ux, part_uy, part_uz};
        for (int d = 0; d < ND; d++) part_pos[d] = part_pos[d] + u[d] * root;
        // :319-354
        const double id[3] = {idx, idy, idz};
        int cell1[3] = {1, 1, 1}, cell2[3] = {1, 1, 1};
        for (int d = 0; d < ND; d++) {
          double cell_r = part_pos[d] * id[d];
          int c1 = (int)std::floor(cell_r + 0.5);
          double cell_frac = (double)c1 - cell_r;
          cell1[d] = c1 + 1;
          tri_weights(cell_frac, &G[d][2]);
          int c2 = (int)std::floor(cell_r);
          cell_frac = (double)c2 - cell_r + 0.5;
          cell2[d] = c2 + 1;
          for (int i = 0; i < 5; i++) H[d][i] = 0.0;
          tri_weights(cell_frac, &H[d][2]);
        }
        const double *gx = &G[0][2], *gy = &G[1][2], *gz = &G[2][2];
        const double *hx = &H[0][2], *hy = &H[1][2], *hz = &H[2][2];
        // e_part.inc / b_part.inc
        double ex_part = gather<ND>(ex, hx, cell2[0], gy, cell1[1], gz, cell1[2]);
        double ey_part = gather<ND>(ey, gx, cell1[0], hy, cell2[1], gz, cell1[2]);
        double ez_part = gather<ND>(ez, gx, cell1[0], gy, cell1[1], hz, cell2[2]);
        double bx_part = gather<ND>(bx, gx, cell1[0], hy, cell2[1], hz, cell2[2]);
        double by_part = gather<ND>(by, hx, cell2[0], gy, cell1[1], hz, cell2[2]);
        double bz_part = gather<ND>(bz, hx, cell2[0], hy, cell2[1], gz, cell1[2]);
        // :382-428 Boris
        double uxm = part_ux + cmratio * ex_part;
        double uym = part_uy + cmratio * ey_part;
        double uzm = part_uz + cmratio * ez_part;
        if (w.cfg.hc_push) {  // :386-398 (1D :345-357, 3D :423-435), Higuera & Cary, Phys. Plasmas 24, 052104
          gamma_rel = uxm * uxm + uym * uym + uzm * uzm + 1.0;
          const double alpha = 0.5 * part_q * dt / S.mass;
          const double beta_x = alpha * bx_part;
          const double beta_y = alpha * by_part;
          const double beta_z = alpha * bz_part;
          const double beta2 = beta_x * beta_x + beta_y * beta_y + beta_z * beta_z;
          const double sigma = gamma_rel - beta2;
          const double beta_dot_u = beta_x * uxm + beta_y * uym + beta_z * uzm;
          gamma_rel = sigma + std::sqrt(sigma * sigma + 4.0 * (beta2 + beta_dot_u * beta_dot_u));
          gamma_rel = std::sqrt(0.5 * gamma_rel);
        } else {
          gamma_rel = std::sqrt(uxm * uxm + uym * uym + uzm * uzm + 1.0);
        }
        root = ccmratio / gamma_rel;
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
        // :431-442; the three trees differ textually here
        double part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz;
        gamma_rel = std::sqrt(part_u2 + 1.0);
        double delta[3] = {0, 0, 0}, part_vy = 0.0, part_vz = 0.0;
        if (ND == 1) {
          // epoch1d particles.F90:392-396
          root = c / gamma_rel;
          delta[0] = part_ux * root * dto2;
          part_vy = part_uy * root;
          part_vz = part_uz * root;
        } else if (ND == 2) {
          // epoch2d particles.F90:433-438
          double igamma = 1.0 / gamma_rel;
          root = dtco2 * igamma;
          delta[0] = part_ux * root;
          delta[1] = part_uy * root;
          part_vz = part_uz * c * igamma;
        } else {
          // epoch3d particles.F90:470-474
          root = dtco2 / gamma_rel;
          delta[0] = part_ux * root;
          delta[1] = part_uy * root;
          delta[2] = part_uz * root;
        }
        for (int d = 0; d < ND; d++) part_pos[d] = part_pos[d] + delta[d];
        // :446-459
        bool cand = false;
        for (int d = 0; d < ND; d++) {
          P.pos[d] = part_pos[d] + R.grid_min_local[d];
          if (P.pos[d] < bnd_min[d] || P.pos[d] > bnd_max[d]) cand = true;
        }
        P.p[0] = part_mc * part_ux;
        P.p[1] = part_mc * part_uy;
        P.p[2] = part_mc * part_uz;
        if (cand) R.bnd_cand[is].push_back((int64_t)ip);

        if (!deposit) continue;
        // :494-547
        int dcell[3] = {0, 0, 0}, mn[3], mx[3];
        for (int d = 0; d < ND; d++) {
          part_pos[d] = part_pos[d] + delta[d];
          double cell_r = part_pos[d] * id[d];
          int c3 = (int)std::floor(cell_r + 0.5);
          double cell_frac = (double)c3 - cell_r;
          c3 = c3 + 1;
          for (int i = 0; i < 5; i++) H[d][i] = 0.0;
          dcell[d] = c3 - cell1[d];
          tri_weights(cell_frac, &H[d][2 + dcell[d]]);
          for (int i = 0; i < 5; i++) H[d][i] = H[d][i] - G[d][i];
          mn[d] = sf_min + (dcell[d] - 1) / 2;
          mx[d] = sf_max + (dcell[d] + 1) / 2;
        }
        if (ND == 1) {
          // epoch1d particles.F90:489-507
          const double fjx = fcx * part_q;
          const double fjy = fcy * part_q * part_vy;
          const double fjz = fcy * part_q * part_vz;
          double jxh = 0.0;
          for (int ix = mn[0]; ix <= mx[0]; ix++) {
            int cx = cell1[0] + ix;
            double wx = hx[ix];
            double wy = gx[ix] + 0.5 * hx[ix];
            jxh = jxh - fjx * wx;
            double jyh = fjy * wy;
            double jzh = fjz * wy;
            jx(cx) = jx(cx) + jxh;
            jy(cx) = jy(cx) + jyh;
            jz(cx) = jz(cx) + jzh;
          }
        } else if (ND == 2) {
          // epoch2d particles.F90:549-579
          const double fjx = fcx * part_q;
          const double fjy = fcy * part_q;
          const double fjz = fcz * part_q * part_vz;
          double jyh[5] = {0, 0, 0, 0, 0};
          for (int iy = mn[1]; iy <= mx[1]; iy++) {
            int cy = cell1[1] + iy;
            double yfac1 = gy[iy] + 0.5 * hy[iy];
            double yfac2 = third * hy[iy] + 0.5 * gy[iy];
            double hy_iy = hy[iy];
            double jxh = 0.0;
            for (int ix = mn[0]; ix <= mx[0]; ix++) {
              int cx = cell1[0] + ix;
              double xfac1 = gx[ix] + 0.5 * hx[ix];
              double wx = hx[ix] * yfac1;
              double wy = hy_iy * xfac1;
              double wz = gx[ix] * yfac1 + hx[ix] * yfac2;
              jxh = jxh - fjx * wx;
              jyh[ix + 2] = jyh[ix + 2] - fjy * wy;
              double jzh = fjz * wz;
              jx(cx, cy) = jx(cx, cy) + jxh;
              jy(cx, cy) = jy(cx, cy) + jyh[ix + 2];
              jz(cx, cy) = jz(cx, cy) + jzh;
            }
          }
        } else {
          // epoch3d particles.F90:603-648
          const double fjx = fcx * part_q;
          const double fjy = fcy * part_q;
          const double fjz = fcz * part_q;
          double jzh[5][5];
          for (auto &row : jzh) for (double &v : row) v = 0.0;
          for (int iz = mn[2]; iz <= mx[2]; iz++) {
            int cz = cell1[2] + iz;
            double zfac1 = gz[iz] + 0.5 * hz[iz];
            double zfac2 = third * hz[iz] + 0.5 * gz[iz];
            double gz_iz = gz[iz], hz_iz = hz[iz];
            double jyh[5] = {0, 0, 0, 0, 0};
            for (int iy = mn[1]; iy <= mx[1]; iy++) {
              int cy = cell1[1] + iy;
              double yfac1 = gy[iy] + 0.5 * hy[iy];
              double yfac2 = third * hy[iy] + 0.5 * gy[iy];
              double hygz = hy[iy] * gz_iz;
              double hyhz = hy[iy] * hz_iz;
              double yzfac = gy[iy] * zfac1 + hy[iy] * zfac2;
              double hzyfac1 = hz_iz * yfac1;
              double hzyfac2 = hz_iz * yfac2;
              double jxh = 0.0;
              for (int ix = mn[0]; ix <= mx[0]; ix++) {
                int cx = cell1[0] + ix;
                double xfac1 = gx[ix] + 0.5 * hx[ix];
                double xfac2 = third * hx[ix] + 0.5 * gx[ix];
                double wx = hx[ix] * yzfac;
                double wy = xfac1 * hygz + xfac2 * hyhz;
                double wz = gx[ix] * hzyfac1 + hx[ix] * hzyfac2;
                jxh = jxh - fjx * wx;
                jyh[ix + 2] = jyh[ix + 2] - fjy * wy;
                jzh[iy + 2][ix + 2] = jzh[iy + 2][ix + 2] - fjz * wz;
                jx(cx, cy, cz) = jx(cx, cy, cz) + jxh;
                jy(cx, cy, cz) = jy(cx, cy, cz) + jyh[ix + 2];
                jz(cx, cy, cz) = jz(cx, cy, cz) + jzh[iy + 2][ix + 2];
              }
            }
          }
        }
      }
    }
  }
}

// boundary.F90:948-1025
void setup_bc_lists(World &w) {
  for (Rank &R : w.r)
    for (size_t is = 0; is < w.sp.size(); is++) {
      R.bnd_cand[is].clear();
      auto &pl = R.part[is];
      for (size_t ip = 0; ip < pl.size(); ip++) {
        bool cand = false;
        for (int d = 0; d < w.nd; d++)
          if (pl[ip].pos[d] < R.min_local[d] || pl[ip].pos[d] > R.max_local[d]) cand = true;
        if (cand) R.bnd_cand[is].push_back((int64_t)ip);
      }
    }
}

// boundary.F90:1029-1462 (2D); 3D/1D identical per axis.  No thermal / CPML.
// The thermal branch of particle_bcs (boundary.F90:1104-1148 and its x_max / y / z copies; epoch3d :1496-1550,
// epoch1d :728-750): the boundary temperature is interpolated with the triangle weights at the particle's
// transverse position, the momentum normal to the wall is drawn from the flux distribution pointing inwards
// (flux_momentum_from_temperature with zero drift: a Rayleigh deviate, particle_temperature.F90:409-460), the
// other two from Maxwellians (momentum_from_temperature :388-398), all from the rank's KISS stream.
void thermal_reemit(World &w, Rank &R, int is, Particle &cur, int d, int side, double direction) {
  const int nd = w.nd;
  const double mass = w.sp[is].mass;
  double temp[3] = {0.0, 0.0, 0.0};
  const std::vector<double> &T = R.ext_temp[is][2 * d + side];
  int tr[2], ntr = 0;
  for (int q = 0; q < nd; q++) if (q != d) tr[ntr++] = q;
  size_t plane = 1;
  for (int q = 0; q < ntr; q++) plane *= (size_t)(R.n[tr[q]] + 2 * NG);
  if (T.empty()) throw std::runtime_error("thermal boundary without ext_temp");
  int cell[2] = {0, 0};
  double g[2][3] = {{0.0, 1.0, 0.0}, {0.0, 1.0, 0.0}};
  for (int q = 0; q < ntr; q++) {
    const double cell_r = (cur.pos[tr[q]] - R.grid_min_local[tr[q]]) / w.d[tr[q]];
    int c = (int)std::floor(cell_r + 0.5);
    const double cell_frac = (double)c - cell_r;
    cell[q] = c + 1;
    const double cf2 = cell_frac * cell_frac;
    g[q][0] = 0.5 * (0.25 + cf2 + cell_frac);
    g[q][1] = 0.75 - cf2;
    g[q][2] = 0.5 * (0.25 + cf2 - cell_frac);
  }
  for (int i = 0; i < 3; i++) {
    temp[i] = 0.0;
    if (ntr == 0) {
      temp[i] = T[i];
    } else if (ntr == 1) {
      for (int a = -1; a <= 1; a++) temp[i] = temp[i] + g[0][a + 1] * T[(size_t)(cell[0] + a + NG - 1) + plane * i];
    } else {
      const size_t e0 = (size_t)(R.n[tr[0]] + 2 * NG);
      for (int b = -1; b <= 1; b++)
        for (int a = -1; a <= 1; a++)
          temp[i] = temp[i] + g[0][a + 1] * g[1][b + 1] * T[(size_t)(cell[0] + a + NG - 1) + e0 * (size_t)(cell[1] + b + NG - 1) + plane * i];
    }
  }
  // the wall-normal component first (i = d), then the others in index order -- for an x wall this is the
  // reference's x, y, z order of draws
  auto mft = [&](double t) { return R.rng.box_muller(std::sqrt(t * kb * mass), 0.0); };
  for (int i = 0; i < 3; i++) {
    const int comp = i == 0 ? d : (i <= d ? i - 1 : i);
    if (comp == d) {
      const double mom1 = mft(temp[comp]), mom2 = mft(temp[comp]);
      cur.p[comp] = direction * std::sqrt(mom1 * mom1 + mom2 * mom2);
    } else {
      cur.p[comp] = mft(temp[comp]);
    }
  }
}

void particle_bcs(World &w) {
  const int nd = w.nd;
  double shift[3];
  for (int d = 0; d < nd; d++) shift[d] = w.length[d] + 2.0 * w.d[d] * (double)w.cpml_t;   // boundary.F90:1047-1048
  for (size_t is = 0; is < w.sp.size(); is++) {
    const SpeciesCfg &S = w.sp[is];
    // send lists per rank per direction
    std::vector<std::vector<Particle>> send(w.nranks * 27);
    for (Rank &R : w.r) {
      auto &pl = R.part[is];
      std::vector<char> gone(pl.size(), 0);
      for (int64_t ip : R.bnd_cand[is]) {
        Particle &cur = pl[ip];
        int bd[3] = {0, 0, 0};
        bool out_of_bounds = false;
        for (int d = 0; d < nd; d++) {
          const double part_pos = cur.pos[d];
          // min side (:1076-1159)
          int sgn = -1;
          if (is_cpml(w.bc_field[2 * d])) {   // boundary.F90:1077-1089: the layer belongs to the boundary rank
            if (R.is_bnd[2 * d]) {
              if (part_pos < w.min_outer[d]) { bd[d] = 0; out_of_bounds = true; }
            } else if (part_pos < R.min_local[d]) bd[d] = sgn;
          } else if (part_pos < R.min_local[d]) {
            bd[d] = sgn;
            int bc = S.bc_particle[2 * d];
            if (bc == c_bc_reflect) {
              if (R.is_bnd[2 * d]) {
                bd[d] = 0;
                cur.pos[d] = 2.0 * w.cfg.xmin[d] - part_pos;
                cur.p[d] = -cur.p[d];
              }
            } else if (bc == c_bc_periodic) {
              if (R.is_bnd[2 * d]) cur.pos[d] = part_pos - sgn * shift[d];
            } else if (bc == c_bc_thermal) {   // boundary.F90:1104-1148
              if (part_pos < w.min_outer[d]) {
                bd[d] = 0;
                thermal_reemit(w, R, (int)is, cur, d, 0, -(double)sgn);
                cur.pos[d] = 2.0 * w.min_outer[d] - part_pos;
              } else if (R.is_bnd[2 * d]) bd[d] = 0;
            } else {
              if (part_pos < w.min_outer[d]) { bd[d] = 0; out_of_bounds = true; }
              else if (R.is_bnd[2 * d]) bd[d] = 0;
            }
          }
          // max side (:1161-1244)
          sgn = 1;
          if (is_cpml(w.bc_field[2 * d + 1])) {   // boundary.F90:1162-1174
            if (R.is_bnd[2 * d + 1]) {
              if (part_pos >= w.max_outer[d]) { bd[d] = 0; out_of_bounds = true; }
            } else if (part_pos >= R.max_local[d]) bd[d] = sgn;
          } else if (part_pos >= R.max_local[d]) {
            bd[d] = sgn;
            int bc = S.bc_particle[2 * d + 1];
            if (bc == c_bc_reflect) {
              if (R.is_bnd[2 * d + 1]) {
                bd[d] = 0;
                cur.pos[d] = 2.0 * w.cfg.xmax[d] - part_pos;
                cur.p[d] = -cur.p[d];
              }
            } else if (bc == c_bc_periodic) {
              if (R.is_bnd[2 * d + 1]) cur.pos[d] = part_pos - sgn * shift[d];
            } else if (bc == c_bc_thermal) {   // boundary.F90:1189-1233
              if (part_pos >= w.max_outer[d]) {
                bd[d] = 0;
                thermal_reemit(w, R, (int)is, cur, d, 1, -(double)sgn);
                cur.pos[d] = 2.0 * w.max_outer[d] - part_pos;
              } else if (R.is_bnd[2 * d + 1]) bd[d] = 0;
            } else {
              if (part_pos >= w.max_outer[d]) { bd[d] = 0; out_of_bounds = true; }
              else if (R.is_bnd[2 * d + 1]) bd[d] = 0;
            }
          }
        }
        if (out_of_bounds) {
          gone[ip] = 1;
        } else if (std::abs(bd[0]) + std::abs(bd[1]) + std::abs(bd[2]) > 0) {
          gone[ip] = 1;
          send[R.rank * 27 + (bd[2] + 1) * 9 + (bd[1] + 1) * 3 + (bd[0] + 1)].push_back(cur);
        }
      }
      size_t o = 0;
      for (size_t ip = 0; ip < pl.size(); ip++)
        if (!gone[ip]) pl[o++] = pl[ip];
      pl.resize(o);
      R.bnd_cand[is].clear();
    }
    // :1436-1446: recv(-ix,-iy) from neighbour(-ix,-iy) gets that rank's send(ix,iy)
    for (Rank &R : w.r) {
      const int z0 = nd >= 3 ? -1 : 0, y0 = nd >= 2 ? -1 : 0;
      for (int iz = z0; iz <= -z0; iz++)
        for (int iy = y0; iy <= -y0; iy++)
          for (int ix = -1; ix <= 1; ix++) {
            if (std::abs(ix) + std::abs(iy) + std::abs(iz) == 0) continue;
            int src = R.neighbour[-iz + 1][-iy + 1][-ix + 1];
            if (src < 0) continue;
            auto &sv = send[src * 27 + (iz + 1) * 9 + (iy + 1) * 3 + (ix + 1)];
            // the sender addressed neighbour(ix,iy,iz); it must be us
            if (w.r[src].neighbour[iz + 1][iy + 1][ix + 1] != R.rank) continue;
            R.part[is].insert(R.part[is].end(), sv.begin(), sv.end());
          }
    }
  }
}

// boundary.F90:534-630: reflecting fold of J ghost cells
void particle_reflection_bcs(World &w, int which, int flip_dir, const int *bcs = nullptr) {
  if (!bcs) bcs = w.bc_allspecies;
  for (Rank &R : w.r) {
    Arr &a = R.f[which];
    for (int d = 0; d < w.nd; d++) {
      const int nn = R.n[d];
      Box b = full_box(a);
      auto fold = [&](int dst, int src, double sgn) {
        Box q = b;
        q.lo[d] = q.hi[d] = dst;
        for (int k = q.lo[2]; k <= q.hi[2]; k++)
          for (int j = q.lo[1]; j <= q.hi[1]; j++)
            for (int i = q.lo[0]; i <= q.hi[0]; i++) {
              int s[3] = {i, j, k};
              s[d] = src;
              a(i, j, k) = a(i, j, k) + sgn * a(s[0], s[1], s[2]);
              a(s[0], s[1], s[2]) = 0.0;
            }
      };
      if (R.is_bnd[2 * d] && bcs[2 * d] == c_bc_reflect) {
        if (flip_dir == d) for (int i = 1; i <= NG - 1; i++) fold(i, -i, -1.0);
        else for (int i = 1; i <= NG - 1; i++) fold(i, 1 - i, +1.0);
      }
      if (R.is_bnd[2 * d + 1] && bcs[2 * d + 1] == c_bc_reflect) {
        if (flip_dir == d) for (int i = 1; i <= NG; i++) fold(nn - i, nn + i, -1.0);
        else for (int i = 1; i <= NG; i++) fold(nn + 1 - i, nn + i, +1.0);
      }
    }
  }
}

// boundary.F90:634-751: send ghost layers, add into the neighbour's interior
void particle_periodic_bcs(World &w, int which, const int *bcs = nullptr) {
  if (!bcs) bcs = w.bc_allspecies;
  std::vector<std::vector<double>> temp(w.nranks);
  for (int d = 0; d < w.nd; d++) {
    for (int pass = 0; pass < 2; pass++) {
      // pass 0: send high ghosts (nn+1..nn+ng) to +1, receive from -1, add into 1..ng
      // pass 1: send low ghosts (1-ng..0) to -1, receive from +1, add into nn+1-ng..nn
      for (int rk = 0; rk < w.nranks; rk++) {
        Rank &R = w.r[rk];
        temp[rk].clear();
        int nl[2] = {nbr(R, d, -1), nbr(R, d, +1)};
        if (R.is_bnd[2 * d] && bcs[2 * d] != c_bc_periodic) nl[0] = -1;
        if (R.is_bnd[2 * d + 1] && bcs[2 * d + 1] != c_bc_periodic) nl[1] = -1;
        int src = pass == 0 ? nl[0] : nl[1];
        if (src < 0) continue;
        const Rank &S = w.r[src];
        // the sender must also be willing to send towards us
        int snl = pass == 0 ? nbr(S, d, +1) : nbr(S, d, -1);
        int sb = pass == 0 ? 2 * d + 1 : 2 * d;
        if (S.is_bnd[sb] && bcs[sb] != c_bc_periodic) snl = -1;
        if (snl != rk) continue;
        const Arr &a = S.f[which];
        Box b = full_box(a);
        if (pass == 0) { b.lo[d] = S.n[d] + 1; b.hi[d] = S.n[d] + NG; }
        else { b.lo[d] = 1 - NG; b.hi[d] = 0; }
        copy_out(a, b, temp[rk]);
      }
      for (int rk = 0; rk < w.nranks; rk++) {
        if (temp[rk].empty()) continue;
        Rank &R = w.r[rk];
        Arr &a = R.f[which];
        Box b = full_box(a);
        if (pass == 0) { b.lo[d] = 1; b.hi[d] = NG; }
        else { b.lo[d] = R.n[d] + 1 - NG; b.hi[d] = R.n[d]; }
        size_t n = 0;
        for (int k = b.lo[2]; k <= b.hi[2]; k++)
          for (int j = b.lo[1]; j <= b.hi[1]; j++)
            for (int i = b.lo[0]; i <= b.hi[0]; i++) a(i, j, k) = a(i, j, k) + temp[rk][n++];
      }
    }
  }
}

// current_smooth.F90:61-141 (strided binomial filter; 1D :104-126, 3D :114-144).  As in the
// reference, beta is fixed by the initial alpha and alpha only changes after the first
// compensation pass has already been done.
void smooth_array(World &w, int which) {
  const int nd = w.nd;
  const int its = w.cfg.smooth_its, comp_its = w.cfg.smooth_comp_its;
  int strides[4] = {1, 0, 0, 0}, ns = 1;
  if (w.cfg.smooth_nstrides > 0) {
    ns = w.cfg.smooth_nstrides;
    for (int i = 0; i < ns; i++) strides[i] = w.cfg.smooth_strides[i];
  }
  double alpha = 0.5;
  const double beta = nd == 1 ? (1.0 - alpha) * 0.5 : nd == 2 ? (1.0 - alpha) * 0.25 : (1.0 - alpha) / 6.0;
  for (Rank &R : w.r) R.f[WK].v = R.f[which].v;
  for (int it = 1; it <= its + comp_its; it++) {
    for (int is = 0; is < ns; is++) {
      field_bc(w, WK);
      const int cs = strides[is];
      for (Rank &R : w.r) {
        Arr &a = R.f[which];
        Arr &wk = R.f[WK];
        const int k1 = nd >= 3 ? R.n[2] : 1, j1 = nd >= 2 ? R.n[1] : 1;
        for (int iz = 1; iz <= k1; iz++)
          for (int iy = 1; iy <= j1; iy++)
            for (int ix = 1; ix <= R.n[0]; ix++) {
              if (nd == 1)
                a(ix) = alpha * wk(ix) + (wk(ix - cs) + wk(ix + cs)) * beta;
              else if (nd == 2)
                a(ix, iy) = alpha * wk(ix, iy) +
                            (wk(ix - cs, iy) + wk(ix + cs, iy) + wk(ix, iy - cs) + wk(ix, iy + cs)) * beta;
              else
                a(ix, iy, iz) = alpha * wk(ix, iy, iz) +
                                (wk(ix - cs, iy, iz) + wk(ix + cs, iy, iz) + wk(ix, iy - cs, iz) + wk(ix, iy + cs, iz) +
                                 wk(ix, iy, iz - cs) + wk(ix, iy, iz + cs)) * beta;
            }
        for (int iz = 1; iz <= k1; iz++)
          for (int iy = 1; iy <= j1; iy++)
            for (int ix = 1; ix <= R.n[0]; ix++) {
              if (nd == 1) wk(ix) = a(ix);
              else if (nd == 2) wk(ix, iy) = a(ix, iy);
              else wk(ix, iy, iz) = a(ix, iy, iz);
            }
      }
    }
    if (it > its) alpha = (double)its * 0.5 + 1.0;
  }
}

// current_smooth.F90:29-45 with boundary.F90:1466-1475, 783-804
// current_bcs(species) with c_bc_mixed (particles.F90:645; boundary.F90:783-804, :749 particle_clear_bcs):
// the ghost cells hold this species' current only (they are cleared after every species)
void current_bcs_species(World &w, int is) {
  int bcs[6];
  for (int i = 0; i < 6; i++) {
    int b = i < 2 * w.nd ? w.sp[is].bc_particle[i] : c_bc_open;
    if (b != c_bc_reflect && b != c_bc_periodic) b = c_bc_open;
    bcs[i] = b;
  }
  for (int q = 0; q < 3; q++) {
    particle_reflection_bcs(w, JX + q, q, bcs);
    particle_periodic_bcs(w, JX + q, bcs);
    for (Rank &R : w.r) {  // particle_clear_bcs: everything outside 1..n
      Arr &a = R.f[JX + q];
      for (int k = a.lo[2]; k < a.lo[2] + a.sz[2]; k++)
        for (int j = a.lo[1]; j < a.lo[1] + a.sz[1]; j++)
          for (int i = a.lo[0]; i < a.lo[0] + a.sz[0]; i++) {
            const bool in = (i >= 1 && i <= R.n[0]) && (w.nd < 2 || (j >= 1 && j <= R.n[1])) &&
                            (w.nd < 3 || (k >= 1 && k <= R.n[2]));
            if (!in) a(i, j, k) = 0.0;
          }
    }
  }
}

void current_finish(World &w) {
  for (int q = 0; q < 3 && !w.bc_mixed; q++) {
    particle_reflection_bcs(w, JX + q, q);
    particle_periodic_bcs(w, JX + q);
  }
  for (int q = 0; q < 3; q++) field_bc(w, JX + q);
  if (w.cfg.smooth_its + w.cfg.smooth_comp_its > 0)
    for (int q = 0; q < 3; q++) smooth_array(w, JX + q);
}

// ---------------------------------------------------------------------------
// Loader: helper.F90:371-661 (load_particles, npart_per_cell >= 0 branch),
// :667-776 (setup_particle_density), particle_temperature.F90:30-81,388-398,
// include/particle_to_grid.inc, include/triangle/gxfac.inc
// ---------------------------------------------------------------------------
inline void particle_to_grid(const World &w, const Rank &R, const Particle &P, int cell[3],
                             double g[3][3]) {
  for (int d = 0; d < w.nd; d++) {
    double cell_r = (P.pos[d] - R.grid_min_local[d]) / w.d[d];
    int cx = (int)std::floor(cell_r + 0.5);
    double cf = (double)cx - cell_r;
    cell[d] = cx + 1;
    double c2 = cf * cf;
    g[d][0] = 0.5 * (0.25 + c2 + cf);
    g[d][1] = 0.75 - c2;
    g[d][2] = 0.5 * (0.25 + c2 - cf);
  }
  for (int d = w.nd; d < 3; d++) { cell[d] = 1; g[d][0] = g[d][2] = 0.0; g[d][1] = 1.0; }
}

// io/calc_df.F90: calc_number_density :689-757 (kind 0), calc_charge_density :608-685 (kind 1),
// calc_mass_density :35-110 (kind 2) into the work array WK (1D/3D trees: same routines, gx*wdata resp.
// gx*gy*gz*wdata).  species < 0 sums all species (tracers skipped, :647-649); calc_boundary = processor_
// summation_bcs without a sign flip, per species under c_bc_mixed, once otherwise (boundary.F90:783-804);
// then the scaling (number: 1/(dx*dy), the others 1/dx/dy) and field_zero_gradient(c_stagger_centre) on
// every boundary.  Ghost cells of periodic / inter-rank edges keep their deposits, as in the reference.
void calc_moment(World &w, int kind, int species) {
  const int nd = w.nd;
  for (Rank &R : w.r) std::fill(R.f[WK].v.begin(), R.f[WK].v.end(), 0.0);
  double idx;
  if (kind == 0) {
    double vol = w.d[0];
    for (int d = 1; d < nd; d++) vol = vol * w.d[d];
    idx = 1.0 / vol;
  } else {
    idx = 1.0 / w.d[0];
    for (int d = 1; d < nd; d++) idx = idx / w.d[d];
  }
  const bool spec_sum = species < 0;
  const int no_flip = nd + 3;  // flip_direction absent: no axis matches
  for (int is = spec_sum ? 0 : species; is < (spec_sum ? (int)w.sp.size() : species + 1); is++) {
    const SpeciesCfg &S = w.sp[is];
    if (spec_sum && S.zero_current) continue;
    for (Rank &R : w.r) {
      Arr &a = R.f[WK];
      for (const Particle &P : R.part[is]) {
        const double wdata = kind == 0 ? P.w : (kind == 1 ? S.charge : S.mass) * P.w;
        int cell[3];
        double g[3][3];
        particle_to_grid(w, R, P, cell, g);
        if (nd == 1) {
          for (int ix = -1; ix <= 1; ix++) a(cell[0] + ix) = a(cell[0] + ix) + g[0][ix + 1] * wdata;
        } else if (nd == 2) {
          for (int iy = -1; iy <= 1; iy++)
            for (int ix = -1; ix <= 1; ix++)
              a(cell[0] + ix, cell[1] + iy) = a(cell[0] + ix, cell[1] + iy) + g[0][ix + 1] * g[1][iy + 1] * wdata;
        } else {
          for (int iz = -1; iz <= 1; iz++)
            for (int iy = -1; iy <= 1; iy++)
              for (int ix = -1; ix <= 1; ix++)
                a(cell[0] + ix, cell[1] + iy, cell[2] + iz) =
                    a(cell[0] + ix, cell[1] + iy, cell[2] + iz) + g[0][ix + 1] * g[1][iy + 1] * g[2][iz + 1] * wdata;
        }
      }
    }
    if (w.bc_mixed) {  // calc_boundary(data_array, ispecies)
      int bcs[6];
      for (int i = 0; i < 6; i++) {
        int b = i < 2 * nd ? S.bc_particle[i] : c_bc_open;
        if (b != c_bc_reflect && b != c_bc_periodic) b = c_bc_open;
        bcs[i] = b;
      }
      particle_reflection_bcs(w, WK, no_flip, bcs);
      particle_periodic_bcs(w, WK, bcs);
      for (Rank &R : w.r) {  // particle_clear_bcs
        Arr &a = R.f[WK];
        for (int k = a.lo[2]; k < a.lo[2] + a.sz[2]; k++)
          for (int j = a.lo[1]; j < a.lo[1] + a.sz[1]; j++)
            for (int i = a.lo[0]; i < a.lo[0] + a.sz[0]; i++) {
              const bool in = (i >= 1 && i <= R.n[0]) && (nd < 2 || (j >= 1 && j <= R.n[1])) &&
                              (nd < 3 || (k >= 1 && k <= R.n[2]));
              if (!in) a(i, j, k) = 0.0;
            }
      }
    }
  }
  if (!w.bc_mixed) {  // calc_boundary(data_array)
    particle_reflection_bcs(w, WK, no_flip);
    particle_periodic_bcs(w, WK);
  }
  for (Rank &R : w.r)
    for (double &v : R.f[WK].v) v = v * idx;
  for (int i = 0; i < 2 * nd; i++) field_mirror(w, WK, i, +1.0);
}

// calc_boundary(array, ispecies) / calc_boundary(array) of io/calc_df.F90:24-31 on a work array
static void calc_boundary_species(World &w, int which, const SpeciesCfg &S) {
  if (!w.bc_mixed) return;  // boundary.F90:790-792
  const int nd = w.nd;
  int bcs[6];
  for (int i = 0; i < 6; i++) {
    int b = i < 2 * nd ? S.bc_particle[i] : c_bc_open;
    if (b != c_bc_reflect && b != c_bc_periodic) b = c_bc_open;
    bcs[i] = b;
  }
  particle_reflection_bcs(w, which, nd + 3, bcs);
  particle_periodic_bcs(w, which, bcs);
  for (Rank &R : w.r) {  // particle_clear_bcs
    Arr &a = R.f[which];
    for (int k = a.lo[2]; k < a.lo[2] + a.sz[2]; k++)
      for (int j = a.lo[1]; j < a.lo[1] + a.sz[1]; j++)
        for (int i = a.lo[0]; i < a.lo[0] + a.sz[0]; i++) {
          const bool in = (i >= 1 && i <= R.n[0]) && (nd < 2 || (j >= 1 && j <= R.n[1])) &&
                          (nd < 3 || (k >= 1 && k <= R.n[2]));
          if (!in) a(i, j, k) = 0.0;
        }
  }
}
static void calc_boundary_all(World &w, int which) {
  if (w.bc_mixed) return;  // boundary.F90:794-796
  particle_reflection_bcs(w, which, w.nd + 3);
  particle_periodic_bcs(w, which);
}
// data(cell + offsets) += gx*gy*gz * wdata for up to 4 (array, wdata) pairs; the weight product is formed
// first, left to right, as in the reference's `gx(ix) * gy(iy) * wdata`
template <class F>
static inline void stencil3(int nd, const int cell[3], const double g[3][3], F &&f) {
  if (nd == 1) {
    for (int ix = -1; ix <= 1; ix++) f(cell[0] + ix, 1, 1, g[0][ix + 1]);
  } else if (nd == 2) {
    for (int iy = -1; iy <= 1; iy++)
      for (int ix = -1; ix <= 1; ix++) f(cell[0] + ix, cell[1] + iy, 1, g[0][ix + 1] * g[1][iy + 1]);
  } else {
    for (int iz = -1; iz <= 1; iz++)
      for (int iy = -1; iy <= 1; iy++)
        for (int ix = -1; ix <= 1; ix++)
          f(cell[0] + ix, cell[1] + iy, cell[2] + iz, g[0][ix + 1] * g[1][iy + 1] * g[2][iz + 1]);
  }
}

// The "ratio" family of io/calc_df.F90: sum(g wdata) / MAX(sum(g w), c_tiny), zero-gradient ghosts.
//   sub 0      calc_ekbar :116-221            wdata = (gamma - 1) m c^2 w
//   sub 1..6   calc_ekflux :415-557           direction -x, +x, -y, +y, -z, +z: wdata = -/+ wdata * MIN/MAX(flux, 0),
//                                             flux = xfac u_x / gamma (xfac = c dy, yfac = c dx, zfac = c dx dy in 2D;
//                                             1D: c, c dx, c dx; 3D: c dy dz, c dx dz, c dx dy)
//   sub 7..9   calc_average_momentum :1239-1317  wdata = w p(direction)
void calc_ratio(World &w, int species, int sub) {
  const int nd = w.nd;
  for (Rank &R : w.r) {
    std::fill(R.f[WK].v.begin(), R.f[WK].v.end(), 0.0);
    std::fill(R.f[WK1].v.begin(), R.f[WK1].v.end(), 0.0);
  }
  double flux_fac = 0.0;
  if (sub >= 1 && sub <= 6) {
    const int a = (sub - 1) / 2;
    const double dx = w.d[0], dy = nd >= 2 ? w.d[1] : 0.0, dz = nd >= 3 ? w.d[2] : 0.0;
    if (nd == 1) flux_fac = a == 0 ? c : c * dx;
    else if (nd == 2) flux_fac = a == 0 ? c * dy : a == 1 ? c * dx : c * dx * dy;
    else flux_fac = a == 0 ? c * dy * dz : a == 1 ? c * dx * dz : c * dx * dy;
  }
  const bool spec_sum = species < 0;
  for (int is = spec_sum ? 0 : species; is < (spec_sum ? (int)w.sp.size() : species + 1); is++) {
    const SpeciesCfg &S = w.sp[is];
    if (spec_sum && S.zero_current) continue;
    const double part_mc = c * S.mass;
    for (Rank &R : w.r) {
      Arr &a = R.f[WK], &wt = R.f[WK1];
      for (const Particle &P : R.part[is]) {
        const double part_w = P.w;
        double wdata;
        if (sub <= 6) {
          const double fac = part_mc * part_w * c;
          const double part_ux = P.p[0] / part_mc, part_uy = P.p[1] / part_mc, part_uz = P.p[2] / part_mc;
          const double part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz;
          const double gamma_rel = std::sqrt(part_u2 + 1.0);
          const double gamma_rel_m1 = part_u2 / (gamma_rel + 1.0);
          wdata = gamma_rel_m1 * fac;
          if (sub >= 1) {
            const double u[3] = {part_ux, part_uy, part_uz};
            const double part_flux = flux_fac * u[(sub - 1) / 2] / gamma_rel;
            if ((sub - 1) % 2 == 0) wdata = -wdata * std::min(part_flux, 0.0);
            else wdata = wdata * std::max(part_flux, 0.0);
          }
        } else {
          wdata = part_w * P.p[sub - 7];
        }
        int cell[3];
        double g[3][3];
        particle_to_grid(w, R, P, cell, g);
        stencil3(nd, cell, g, [&](int i, int j, int k, double gg) {
          a(i, j, k) = a(i, j, k) + gg * wdata;
          wt(i, j, k) = wt(i, j, k) + gg * part_w;
        });
      }
    }
    calc_boundary_species(w, WK, S);
    calc_boundary_species(w, WK1, S);
  }
  calc_boundary_all(w, WK);
  calc_boundary_all(w, WK1);
  const double c_tiny = std::numeric_limits<double>::min();
  for (Rank &R : w.r)
    for (size_t q = 0; q < R.f[WK].v.size(); q++) R.f[WK].v[q] = R.f[WK].v[q] / std::max(R.f[WK1].v[q], c_tiny);
  for (int i = 0; i < 2 * nd; i++) field_mirror(w, WK, i, +1.0);
}

// calc_per_species_current, io/calc_df.F90:1132-1235: q w c p_dir / sqrt((m c)^2 + p^2) on the triangle stencil,
// scaled by c / dx / dy
void calc_species_current(World &w, int species, int dir) {
  const int nd = w.nd;
  for (Rank &R : w.r) std::fill(R.f[WK].v.begin(), R.f[WK].v.end(), 0.0);
  double idx = 1.0 / w.d[0];
  for (int d = 1; d < nd; d++) idx = idx / w.d[d];
  const bool spec_sum = species < 0;
  for (int is = spec_sum ? 0 : species; is < (spec_sum ? (int)w.sp.size() : species + 1); is++) {
    const SpeciesCfg &S = w.sp[is];
    if (spec_sum && S.zero_current) continue;
    const double part_mc = c * S.mass;
    for (Rank &R : w.r) {
      Arr &a = R.f[WK];
      for (const Particle &P : R.part[is]) {
        double wdata = S.charge * P.w;
        const double root = 1.0 / std::sqrt(part_mc * part_mc + P.p[0] * P.p[0] + P.p[1] * P.p[1] + P.p[2] * P.p[2]);
        wdata = wdata * P.p[dir] * root;
        int cell[3];
        double g[3][3];
        particle_to_grid(w, R, P, cell, g);
        stencil3(nd, cell, g, [&](int i, int j, int k, double gg) { a(i, j, k) = a(i, j, k) + gg * wdata; });
      }
    }
    calc_boundary_species(w, WK, S);
  }
  calc_boundary_all(w, WK);
  idx = c * idx;
  for (Rank &R : w.r)
    for (double &v : R.f[WK].v) v = v * idx;
  for (int i = 0; i < 2 * nd; i++) field_mirror(w, WK, i, +1.0);
}

// calc_average_weight, io/calc_df.F90:811-873: nearest cell, no ghost-cell sums, no ghost fill
void calc_average_weight(World &w, int species) {
  const int nd = w.nd;
  for (Rank &R : w.r) {
    std::fill(R.f[WK].v.begin(), R.f[WK].v.end(), 0.0);
    std::fill(R.f[WK1].v.begin(), R.f[WK1].v.end(), 0.0);
  }
  const bool spec_sum = species < 0;
  for (int is = spec_sum ? 0 : species; is < (spec_sum ? (int)w.sp.size() : species + 1); is++) {
    if (spec_sum && w.sp[is].zero_current) continue;
    for (Rank &R : w.r)
      for (const Particle &P : R.part[is]) {
        int cell[3] = {1, 1, 1};
        for (int d = 0; d < nd; d++)
          cell[d] = (int)std::floor((P.pos[d] - R.grid_min_local[d]) / w.d[d] + 0.5) + 1;
        R.f[WK](cell[0], cell[1], cell[2]) = R.f[WK](cell[0], cell[1], cell[2]) + P.w;
        R.f[WK1](cell[0], cell[1], cell[2]) = R.f[WK1](cell[0], cell[1], cell[2]) + 1.0;
      }
  }
  const double c_tiny = std::numeric_limits<double>::min();
  for (Rank &R : w.r)
    for (size_t q = 0; q < R.f[WK].v.size(); q++) R.f[WK].v[q] = R.f[WK].v[q] / std::max(R.f[WK1].v[q], c_tiny);
}

// calc_temperature, io/calc_df.F90:877-1128: dir < 0: all three momentum components (dof 3), else one
// (dof 1).  Pass 1: weighted mean of p/sqrt(m) per cell (ghosts restored with field_bc); pass 2: un-weighted
// spread around the mean of the cell each stencil point lies in; sigma / MAX(count, 1e-6) / kb / dof.
void calc_temperature(World &w, int species, int dir) {
  const int nd = w.nd;
  const int MEAN[3] = {WK1, WK2, WK3}, CNT = WK4, SIG = WK;
  for (Rank &R : w.r)
    for (int q : {WK, WK1, WK2, WK3, WK4}) std::fill(R.f[q].v.begin(), R.f[q].v.end(), 0.0);
  const double dof = dir < 0 ? 3.0 : 1.0;
  const bool spec_sum = species < 0;
  const int s0 = spec_sum ? 0 : species, s1 = spec_sum ? (int)w.sp.size() : species + 1;
  auto use = [&](int q) { return dir < 0 || dir == q; };
  for (int is = s0; is < s1; is++) {
    const SpeciesCfg &S = w.sp[is];
    if (spec_sum && S.zero_current) continue;
    const double sqrt_part_m = std::sqrt(S.mass);
    for (Rank &R : w.r) {
      for (const Particle &P : R.part[is]) {
        const double part_w = P.w;
        const double pm[3] = {P.p[0] / sqrt_part_m, P.p[1] / sqrt_part_m, P.p[2] / sqrt_part_m};
        int cell[3];
        double g[3][3];
        particle_to_grid(w, R, P, cell, g);
        stencil3(nd, cell, g, [&](int i, int j, int k, double gg) {
          const double gf = gg * part_w;
          for (int q = 0; q < 3; q++)
            if (use(q)) R.f[MEAN[q]](i, j, k) = R.f[MEAN[q]](i, j, k) + gf * pm[q];
          R.f[CNT](i, j, k) = R.f[CNT](i, j, k) + gf;
        });
      }
    }
    for (int q = 0; q < 3; q++)
      if (use(q)) calc_boundary_species(w, MEAN[q], S);
    calc_boundary_species(w, CNT, S);
  }
  for (int q = 0; q < 3; q++)
    if (use(q)) calc_boundary_all(w, MEAN[q]);
  calc_boundary_all(w, CNT);
  for (Rank &R : w.r) {
    std::vector<double> &pc = R.f[CNT].v;
    for (size_t t = 0; t < pc.size(); t++) pc[t] = std::max(pc[t], 1.e-6);
    for (int q = 0; q < 3; q++)  // the reference divides all three, also the unused (zero) ones
      for (size_t t = 0; t < pc.size(); t++) R.f[MEAN[q]].v[t] = R.f[MEAN[q]].v[t] / pc[t];
  }
  for (int q = 0; q < 3; q++)
    if (use(q)) field_bc(w, MEAN[q]);
  for (Rank &R : w.r) std::fill(R.f[CNT].v.begin(), R.f[CNT].v.end(), 0.0);
  for (int is = s0; is < s1; is++) {
    const SpeciesCfg &S = w.sp[is];
    if (spec_sum && S.zero_current) continue;
    const double sqrt_part_m = std::sqrt(S.mass);
    for (Rank &R : w.r) {
      for (const Particle &P : R.part[is]) {
        const double pm[3] = {P.p[0] / sqrt_part_m, P.p[1] / sqrt_part_m, P.p[2] / sqrt_part_m};
        int cell[3];
        double g[3][3];
        particle_to_grid(w, R, P, cell, g);
        stencil3(nd, cell, g, [&](int i, int j, int k, double gf) {
          double wdata;
          if (dir < 0) {
            const double dx_ = pm[0] - R.f[MEAN[0]](i, j, k), dy_ = pm[1] - R.f[MEAN[1]](i, j, k),
                         dz_ = pm[2] - R.f[MEAN[2]](i, j, k);
            wdata = dx_ * dx_ + dy_ * dy_ + dz_ * dz_;
          } else {
            const double d_ = pm[dir] - R.f[MEAN[dir]](i, j, k);
            wdata = d_ * d_;
          }
          R.f[SIG](i, j, k) = R.f[SIG](i, j, k) + gf * wdata;
          R.f[CNT](i, j, k) = R.f[CNT](i, j, k) + gf;
        });
      }
    }
    calc_boundary_species(w, SIG, S);
    calc_boundary_species(w, CNT, S);
  }
  calc_boundary_all(w, SIG);
  calc_boundary_all(w, CNT);
  for (Rank &R : w.r)
    for (size_t t = 0; t < R.f[SIG].v.size(); t++)
      R.f[SIG].v[t] = R.f[SIG].v[t] / std::max(R.f[CNT].v[t], 1.e-6) / kb / dof;
}

// calc_poynt_flux, io/calc_df.F90:561-604 (epoch1d :441-474, epoch3d :585-650): E x B / mu0 at the cell centre of
// the interior cells.  Every component is averaged over the active axes it is staggered along (setup.F90:124-134):
// two-point means, four-point means in the order (lo,lo) + (hi,lo) + (lo,hi) + (hi,hi) with the lower axis first.
// data_array is INTENT(OUT) and only 1..n is written: the ghost cells are returned as zero here.
static double cell_centred(const Arr &a, int nd, int field, int i, int j, int k) {
  int ax[2], na = 0;
  for (int d = 0; d < nd; d++)
    if (stagger(d, field)) ax[na++] = d;
  auto at = [&](int s0, int s1) {
    int q[3] = {i, j, k};
    if (na >= 1) q[ax[0]] -= s0;
    if (na >= 2) q[ax[1]] -= s1;
    return a(q[0], q[1], q[2]);
  };
  if (na == 0) return at(0, 0);
  if (na == 1) return 0.5 * (at(1, 0) + at(0, 0));
  return 0.25 * (at(1, 1) + at(0, 1) + at(1, 0) + at(0, 0));
}
void calc_poynt_flux(World &w, int dir) {
  const int nd = w.nd;
  const double mu0 = 4.e-7 * pi;
  for (Rank &R : w.r) {
    Arr &out = R.f[WK];
    std::fill(out.v.begin(), out.v.end(), 0.0);
    for (int k = 1; k <= R.n[2]; k++)
      for (int j = 1; j <= R.n[1]; j++)
        for (int i = 1; i <= R.n[0]; i++) {
          const int e1 = EX + (dir + 1) % 3, e2 = EX + (dir + 2) % 3;
          const int b1 = BX + (dir + 1) % 3, b2 = BX + (dir + 2) % 3;
          const double e1c = cell_centred(R.f[e1], nd, e1, i, j, k), e2c = cell_centred(R.f[e2], nd, e2, i, j, k);
          const double b1c = cell_centred(R.f[b1], nd, b1, i, j, k), b2c = cell_centred(R.f[b2], nd, b2, i, j, k);
          out(i, j, k) = (e1c * b2c - e2c * b1c) / mu0;  // x: ey bz - ez by; y: ez bx - ex bz; z: ex by - ey bx
        }
  }
}

void auto_load(World &w) {
  const int nd = w.nd;
  for (size_t is = 0; is < w.sp.size(); is++) {
    const SpeciesCfg &S = w.sp[is];
    for (Rank &R : w.r) {
      // density(ix,iy) evaluated at cell centres incl. ghosts, then field_bc
      // (periodic wrap / neighbour copy reproduces the same analytic values)
      Arr dens, map;
      dens.init(R.n, nd);
      map.init(R.n, nd);
      Box b = full_box(dens);
      for (int k = b.lo[2]; k <= b.hi[2]; k++)
        for (int j = b.lo[1]; j <= b.hi[1]; j++)
          for (int i = b.lo[0]; i <= b.hi[0]; i++) {
            int ii[3] = {i, j, k};
            bool in = true;
            for (int d = 0; d < nd; d++) {
              int gi = ii[d] + R.gmin[d] - 1;
              // periodic images of ghost cells map back into the domain
              if (w.bc_field[2 * d] == c_bc_periodic) {
                int ng_ = w.n_ext[d];
                gi = ((gi - 1) % ng_ + ng_) % ng_ + 1;
              }
              double xc = x_global(w, d, gi);
              if (xc < S.box_lo[d] || xc >= S.box_hi[d]) in = false;
            }
            double v = in ? S.density : 0.0;
            const double density_min = 2.220446049250313e-16;  // helper.F90:198
            if (v >= density_min) map(i, j, k) = 1.0;
            else v = 0.0;
            dens(i, j, k) = v;
          }
      // load_particles, helper.F90:556-583
      std::vector<Particle> &pl = R.part[is];
      pl.clear();
      const int64_t npc = (int64_t)std::floor(S.npart_per_cell);
      const int k1 = nd >= 3 ? R.n[2] : 1, j1 = nd >= 2 ? R.n[1] : 1;
      for (int iz = 1; iz <= k1; iz++)
        for (int iy = 1; iy <= j1; iy++)
          for (int ix = 1; ix <= R.n[0]; ix++) {
            if (map(ix, iy, iz) == 0.0) continue;
            int ii[3] = {ix, iy, iz};
            for (int64_t ipart = 0; ipart < npc; ipart++) {
              Particle P;
              std::memset(&P, 0, sizeof P);
              for (int d = 0; d < nd; d++) {
                double xc = x_global(w, d, ii[d] + R.gmin[d] - 1);
                P.pos[d] = xc + (R.rng.random() - 0.5) * w.d[d];
              }
              pl.push_back(P);
            }
          }
    }
    // helper.F90:658-659
    setup_bc_lists(w);
    particle_bcs(w);
    for (Rank &R : w.r) {
      std::vector<Particle> &pl = R.part[is];
      // recompute density map for weights (same as above)
      Arr dens, map, cnt;
      dens.init(R.n, nd);
      map.init(R.n, nd);
      cnt.init(R.n, nd);
      Box b = full_box(dens);
      for (int k = b.lo[2]; k <= b.hi[2]; k++)
        for (int j = b.lo[1]; j <= b.hi[1]; j++)
          for (int i = b.lo[0]; i <= b.hi[0]; i++) {
            int ii[3] = {i, j, k};
            bool in = true;
            for (int d = 0; d < nd; d++) {
              int gi = ii[d] + R.gmin[d] - 1;
              if (w.bc_field[2 * d] == c_bc_periodic) {
                int ng_ = w.n_ext[d];
                gi = ((gi - 1) % ng_ + ng_) % ng_ + 1;
              }
              double xc = x_global(w, d, gi);
              if (xc < S.box_lo[d] || xc >= S.box_hi[d]) in = false;
            }
            double v = in ? S.density : 0.0;
            if (v >= 2.220446049250313e-16) map(i, j, k) = 1.0;
            else v = 0.0;
            dens(i, j, k) = v;
          }
      // helper.F90:711-757
      for (Particle &P : pl) {
        int cell[3];
        double g[3][3];
        particle_to_grid(w, R, P, cell, g);
        double wdata = 0.0;
        const int z0 = nd >= 3 ? -1 : 0, y0 = nd >= 2 ? -1 : 0;
        for (int sz_ = z0; sz_ <= -z0; sz_++) {
          int i = cell[0], j = cell[1], k = cell[2] + sz_;
          if (nd >= 3 && map(i, j, k) == 0.0) k = cell[2] + sz_ / 2;
          for (int sy = y0; sy <= -y0; sy++) {
            i = cell[0];
            j = cell[1] + sy;
            if (nd >= 2 && map(i, j, k) == 0.0) j = cell[1] + sy / 2;
            for (int sx = -1; sx <= 1; sx++) {
              i = cell[0] + sx;
              if (map(i, j, k) == 0.0) i = cell[0] + sx / 2;
              double wgt = g[0][sx + 1];
              if (nd >= 2) wgt = wgt * g[1][sy + 1];
              if (nd >= 3) wgt = wgt * g[2][sz_ + 1];
              wdata = wdata + wgt * dens(i, j, k);
            }
          }
        }
        P.w = wdata;
        cnt(cell[0], cell[1], cell[2]) += 1.0;
      }
      double vol = 1.0;
      if (nd == 1) vol = w.d[0];
      else if (nd == 2) vol = w.d[0] * w.d[1];
      else vol = w.d[0] * w.d[1] * w.d[2];
      for (Particle &P : pl) {
        int cell[3] = {1, 1, 1};
        for (int d = 0; d < nd; d++)
          cell[d] = (int)std::floor((P.pos[d] - R.grid_min_local[d]) / w.d[d] + 1.5);
        P.w = P.w * vol / cnt(cell[0], cell[1], cell[2]);
      }
    }
    // helper.F90:142-145: x pass over all particles, then y, then z
    for (Rank &R : w.r)
      for (int dir = 0; dir < 3; dir++)
        for (Particle &P : R.part[is]) {
          // uniform temperature / drift: the interpolation sum of a constant
          int cell[3];
          double g[3][3];
          particle_to_grid(w, R, P, cell, g);
          double temp_local = 0.0, drift_local = 0.0;
          const int z0 = nd >= 3 ? -1 : 0, y0 = nd >= 2 ? -1 : 0;
          if (nd == 1) {
            for (int sx = -1; sx <= 1; sx++) {
              temp_local = temp_local + g[0][sx + 1] * S.temp[dir];
              drift_local = drift_local + g[0][sx + 1] * S.drift[dir];
            }
          } else if (nd == 2) {
            for (int sy = y0; sy <= -y0; sy++)
              for (int sx = -1; sx <= 1; sx++) {
                temp_local = temp_local + g[0][sx + 1] * g[1][sy + 1] * S.temp[dir];
                drift_local = drift_local + g[0][sx + 1] * g[1][sy + 1] * S.drift[dir];
              }
          } else {
            for (int sz_ = z0; sz_ <= -z0; sz_++)
              for (int sy = y0; sy <= -y0; sy++)
                for (int sx = -1; sx <= 1; sx++) {
                  temp_local = temp_local + g[0][sx + 1] * g[1][sy + 1] * g[2][sz_ + 1] * S.temp[dir];
                  drift_local = drift_local + g[0][sx + 1] * g[1][sy + 1] * g[2][sz_ + 1] * S.drift[dir];
                }
          }
          // particle_temperature.F90:388-398
          double stdev = std::sqrt(temp_local * kb * S.mass);
          P.p[dir] = R.rng.box_muller(stdev, drift_local);
        }
  }
}

// ---------------------------------------------------------------------------
// Moving window: housekeeping/window.F90 (epoch1d / epoch2d / epoch3d; the x direction only, as there)
// ---------------------------------------------------------------------------
// The species description of this harness is a uniform density inside a box with uniform temperature and drift
// (SpeciesCfg).  The reference evaluates the deck's density / temperature / drift functions at pack_ix = nx and
// pack_iy = 0 .. ny+1 (ghost positions included, no periodic wrap); here a transverse position outside the domain
// takes the domain's edge value, which is what a constant deck expression gives.
static double window_density(const World &w, const Rank &R, const SpeciesCfg &S, int iy, int iz) {
  const int nd = w.nd;
  int ii[3] = {R.n[0], iy, iz};
  bool in = true;
  for (int d = 0; d < nd; d++) {
    int gi = ii[d] + R.gmin[d] - 1;
    if (d > 0) gi = std::min(std::max(gi, 1), w.n_ext[d]);
    const double xc = x_global(w, d, gi);
    if (xc < S.box_lo[d] || xc >= S.box_hi[d]) in = false;
  }
  double v = in ? S.density : 0.0;
  const double dmin = 2.220446049250313e-16;   // initial_conditions%density_min = EPSILON(1.0_num), deck_species_block.F90
  if (v < dmin) v = 0.0;
  return v;
}

// insert_particles (epoch2d window.F90:182-320, epoch1d :158-262, epoch3d :197-352): one cell of fresh plasma beyond
// the right-hand edge, on the x_max ranks, BEFORE the grid moves
template <int ND>
void window_insert_particles(World &w) {
  const double dmin = 2.220446049250313e-16;
  for (Rank &R : w.r) {
    if (!R.is_bnd[1]) continue;
    for (size_t is = 0; is < w.sp.size(); is++) {
      const SpeciesCfg &S = w.sp[is];
      const int64_t npart_per_cell = (int64_t)std::floor(S.npart_per_cell);
      const double npart_frac = S.npart_per_cell - (double)npart_per_cell;
      const double x_grid_max = x_global(w, 0, w.n_ext[0]);
      const double x0 = x_grid_max + 0.5 * w.d[0];
      std::vector<Particle> app;
      auto momentum = [&](Particle &P, const double *temp_local, const double *drift_local) {
        for (int i = 0; i < 3; i++) {
          // momentum_from_temperature, particle_temperature.F90:388-398
          const double stdev = std::sqrt(temp_local[i] * kb * S.mass);
          P.p[i] = R.rng.box_muller(stdev, drift_local[i]);
        }
      };
      if (ND == 1) {
        const double density = window_density(w, R, S, 1, 1);
        if (density < dmin) continue;
        int64_t n_frac = 0;
        if (npart_frac > 0.0 && R.rng.random() < npart_frac) n_frac = 1;
        const double wdata = w.d[0] / (double)(npart_per_cell + n_frac);
        for (int64_t ipart = 1; ipart <= npart_per_cell + n_frac; ipart++) {
          Particle P;
          std::memset(&P, 0, sizeof P);
          P.pos[0] = x0 + R.rng.random() * w.d[0];
          momentum(P, S.temp, S.drift);
          P.w = density * wdata;
          app.push_back(P);
        }
      } else {
        const int ny = R.n[1], nz = ND >= 3 ? R.n[2] : 1;
        const int kz0 = ND >= 3 ? 0 : 1, kz1 = ND >= 3 ? nz + 1 : 1;
        // density(0:ny+1[, 0:nz+1]); temperature and drift are uniform here
        std::vector<double> dens((size_t)(ny + 2) * (ND >= 3 ? nz + 2 : 1));
        auto D = [&](int iy, int iz) -> double & { return dens[(size_t)(ND >= 3 ? iz : 0) * (ny + 2) + iy]; };
        for (int iz = kz0; iz <= kz1; iz++)
          for (int iy = 0; iy <= ny + 1; iy++) D(iy, ND >= 3 ? iz : 0) = window_density(w, R, S, iy, iz);
        for (int iz = 1; iz <= nz; iz++)
          for (int iy = 1; iy <= ny; iy++) {
            const int kz = ND >= 3 ? iz : 0;
            if (D(iy, kz) < dmin) continue;
            int64_t n_frac = 0;
            if (npart_frac > 0.0 && R.rng.random() < npart_frac) n_frac = 1;
            double wdata = w.d[0] * w.d[1];
            if (ND >= 3) wdata = wdata * w.d[2];
            wdata = wdata / (double)(npart_per_cell + n_frac);
            for (int64_t ipart = 1; ipart <= npart_per_cell + n_frac; ipart++) {
              Particle P;
              std::memset(&P, 0, sizeof P);
              const double cell_frac_y = 0.5 - R.rng.random();
              double cell_frac_z = 0.0;
              if (ND >= 3) cell_frac_z = 0.5 - R.rng.random();
              P.pos[0] = x0 + R.rng.random() * w.d[0];
              P.pos[1] = x_global(w, 1, iy + R.gmin[1] - 1) - cell_frac_y * w.d[1];
              if (ND >= 3) P.pos[2] = x_global(w, 2, iz + R.gmin[2] - 1) - cell_frac_z * w.d[2];
              double gy[3], gz[3] = {0.0, 1.0, 0.0};
              const double cy2 = cell_frac_y * cell_frac_y;
              gy[0] = 0.5 * (0.25 + cy2 + cell_frac_y);
              gy[1] = 0.75 - cy2;
              gy[2] = 0.5 * (0.25 + cy2 - cell_frac_y);
              if (ND >= 3) {
                const double cz2 = cell_frac_z * cell_frac_z;
                gz[0] = 0.5 * (0.25 + cz2 + cell_frac_z);
                gz[1] = 0.75 - cz2;
                gz[2] = 0.5 * (0.25 + cz2 - cell_frac_z);
              }
              double temp_local[3] = {0.0, 0.0, 0.0}, drift_local[3] = {0.0, 0.0, 0.0};
              for (int i = 0; i < 3; i++) {
                if (ND == 2) {
                  for (int sy = -1; sy <= 1; sy++) {
                    temp_local[i] = temp_local[i] + gy[sy + 1] * S.temp[i];
                    drift_local[i] = drift_local[i] + gy[sy + 1] * S.drift[i];
                  }
                } else {
                  for (int sz_ = -1; sz_ <= 1; sz_++)
                    for (int sy = -1; sy <= 1; sy++) {
                      temp_local[i] = temp_local[i] + gy[sy + 1] * gz[sz_ + 1] * S.temp[i];
                      drift_local[i] = drift_local[i] + gy[sy + 1] * gz[sz_ + 1] * S.drift[i];
                    }
                }
              }
              momentum(P, temp_local, drift_local);
              double weight_local = 0.0;
              if (ND == 2) {
                for (int sy = -1; sy <= 1; sy++) weight_local = weight_local + gy[sy + 1] * D(iy + sy, 0);
              } else {
                for (int sz_ = -1; sz_ <= 1; sz_++)
                  for (int sy = -1; sy <= 1; sy++) weight_local = weight_local + gy[sy + 1] * gz[sz_ + 1] * D(iy + sy, iz + sz_);
              }
              P.w = weight_local * wdata;
              app.push_back(P);
            }
          }
      }
      R.part[is].insert(R.part[is].end(), app.begin(), app.end());
      R.inserted[is].insert(R.inserted[is].end(), app.begin(), app.end());
    }
  }
}

// shift_window (window.F90:62-94) for one cell: insert, move the grid (x_grid_min, xb_min, x_min, x_max accumulate
// one dx at a time exactly as there; dx and length_x are not recomputed), setup_grid_x (utilities.f90:343-381),
// remove_particles (:324-345), shift_fields (:98-178)
template <int ND>
void shift_window_once(World &w) {
  const double dx = w.d[0];
  window_insert_particles<ND>(w);
  w.grid_min[0] = x_global(w, 0, 1) + dx;
  w.xb_min[0] = w.xb_min[0] + dx;                       // xb_global(1) + dx
  w.cfg.xmin[0] = w.xb_min[0] + dx * (double)w.cpml_t;
  w.cfg.xmax[0] = (w.xb_min[0] + (double)(w.n_ext[0] + 1 - 1) * dx) - dx * (double)w.cpml_t;   // xb_global(nx_global+1) - ...
  for (Rank &R : w.r) {
    R.grid_min_local[0] = x_global(w, 0, w.cell_min[0][R.coords[0]]);
    const double hdx = 0.5 * dx;
    R.min_local[0] = R.grid_min_local[0] - hdx;
    R.max_local[0] = x_global(w, 0, w.cell_max[0][R.coords[0]] + 1) - hdx;
  }
  const double boundary_shift = (double)((1 + png + w.cpml_t) / 2);
  w.min_outer[0] = w.cfg.xmin[0] - boundary_shift * dx;
  w.max_outer[0] = w.cfg.xmax[0] + boundary_shift * dx;
  // remove_particles: only processors on the left
  for (Rank &R : w.r) {
    if (!R.is_bnd[0]) continue;
    for (auto &pl : R.part) {
      size_t k = 0;
      for (size_t i = 0; i < pl.size(); i++)
        if (!(pl[i].pos[0] < w.cfg.xmin[0])) pl[k++] = pl[i];
      pl.resize(k);
    }
  }
  // shift_fields: every array one cell to the left (ghost cells included), then field_bc
  for (int which = 0; which < NFIELD; which++) {
    for (Rank &R : w.r) {
      Arr &a = R.f[which];
      const Box b = full_box(a);
      for (int k = b.lo[2]; k <= b.hi[2]; k++)
        for (int j = b.lo[1]; j <= b.hi[1]; j++)
          for (int i = 1 - NG; i <= R.n[0] + NG - 1; i++) a(i, j, k) = a(i + 1, j, k);
    }
    field_bc(w, which);
  }
  // fix the incoming field cell on the x_max ranks (window.F90:126-143)
  for (Rank &R : w.r) {
    if (!R.is_bnd[1]) continue;
    const int nx = R.n[0];
    Arr &ex = R.f[EX], &ey = R.f[EY], &ez = R.f[EZ], &bx = R.f[BX], &by = R.f[BY], &bz = R.f[BZ];
    const Box b = full_box(ex);
    for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++) {
        ex(nx, j, k) = R.snap_max[EX](1, j, k);
        ex(nx + 1, j, k) = R.snap_max[EX](1, j, k);
        ey(nx + 1, j, k) = R.snap_max[EY](1, j, k);
        ez(nx + 1, j, k) = R.snap_max[EZ](1, j, k);
        ex(nx - 1, j, k) = 0.5 * (ex(nx - 2, j, k) + ex(nx, j, k));
        ey(nx, j, k) = 0.5 * (ey(nx - 1, j, k) + ey(nx + 1, j, k));
        ez(nx, j, k) = 0.5 * (ez(nx - 1, j, k) + ez(nx + 1, j, k));
        bx(nx + 1, j, k) = R.snap_max[BX](1, j, k);
        by(nx, j, k) = R.snap_max[BY](1, j, k);
        bz(nx, j, k) = R.snap_max[BZ](1, j, k);
        bx(nx, j, k) = 0.5 * (bx(nx - 1, j, k) + bx(nx + 1, j, k));
        by(nx - 1, j, k) = 0.5 * (by(nx - 2, j, k) + by(nx, j, k));
        bz(nx - 1, j, k) = 0.5 * (bz(nx - 2, j, k) + bz(nx, j, k));
      }
  }
}

// the part of moving_window (window.F90:350-397) that follows the decision to shift: shift_window(cells),
// setup_bc_lists, particle_bcs.  The decision (window_shift_fraction += dt * window_v_x / dx ...) is the host's.
void shift_window(World &w, int cells) {
  for (int i = 0; i < cells; i++) {
    if (w.nd == 1) shift_window_once<1>(w);
    else if (w.nd == 2) shift_window_once<2>(w);
    else shift_window_once<3>(w);
  }
  setup_bc_lists(w);
  particle_bcs(w);
}

template <int ND>
void init_sequence(World &w) {
  // epoch2d.F90:144-162
  setup_field_boundaries(w);  // after_deck_last, setup.F90:208 → :391
  setup_bc_lists(w);
  particle_bcs(w);
  efield_bcs(w);
  bfield_final_bcs<ND>(w, w.dt / 2.0);
}

}  // namespace

// ---------------------------------------------------------------------------
// C interface (ctypes)
// ---------------------------------------------------------------------------
extern "C" {

void *orc_create(const Config *cfg, const SpeciesCfg *sp) {
  World *w = new World;
  w->cfg = *cfg;
  for (int i = 0; i < cfg->n_species; i++) w->sp.push_back(sp[i]);
  setup_world(*w);
  return w;
}
void orc_destroy(void *h) { delete (World *)h; }
int orc_nranks(void *h) { return ((World *)h)->nranks; }
double orc_dx(void *h, int d) { return ((World *)h)->d[d]; }

// out: n[3], gmin[3], coords[3], is_bnd[6], neighbour[27]
void orc_rank_info(void *h, int rk, int *n, int *gmin, int *coords, int *is_bnd, int *neighbour,
                   double *grid_min_local, double *min_local, double *max_local) {
  Rank &R = ((World *)h)->r[rk];
  for (int d = 0; d < 3; d++) {
    n[d] = R.n[d]; gmin[d] = R.gmin[d]; coords[d] = R.coords[d];
    grid_min_local[d] = R.grid_min_local[d]; min_local[d] = R.min_local[d]; max_local[d] = R.max_local[d];
  }
  for (int i = 0; i < 6; i++) is_bnd[i] = R.is_bnd[i];
  for (int i = 0; i < 27; i++) neighbour[i] = (&R.neighbour[0][0][0])[i];
}
void orc_outer(void *h, double *min_outer, double *max_outer) {
  World &w = *(World *)h;
  for (int d = 0; d < 3; d++) { min_outer[d] = w.min_outer[d]; max_outer[d] = w.max_outer[d]; }
}
double *orc_field(void *h, int rk, int which) { return ((World *)h)->r[rk].f[which].v.data(); }
int64_t orc_field_size(void *h, int rk) { return (int64_t)((World *)h)->r[rk].f[0].v.size(); }
int64_t orc_species_count(void *h, int rk, int is) { return (int64_t)((World *)h)->r[rk].part[is].size(); }
// packed layout = pack_particle order (partlist.F90:414-486): pos(1..ndims), p(1..3), weight
void orc_get_particles(void *h, int rk, int is, double *out) {
  World &w = *(World *)h;
  const int nv = w.nd + 4;
  auto &pl = w.r[rk].part[is];
  for (size_t i = 0; i < pl.size(); i++) {
    double *o = out + i * nv;
    for (int d = 0; d < w.nd; d++) o[d] = pl[i].pos[d];
    for (int d = 0; d < 3; d++) o[w.nd + d] = pl[i].p[d];
    o[w.nd + 3] = pl[i].w;
  }
}
void orc_set_particles(void *h, int rk, int is, int64_t n, const double *in) {
  World &w = *(World *)h;
  const int nv = w.nd + 4;
  auto &pl = w.r[rk].part[is];
  pl.resize(n);
  for (int64_t i = 0; i < n; i++) {
    const double *o = in + i * nv;
    std::memset(&pl[i], 0, sizeof(Particle));
    for (int d = 0; d < w.nd; d++) pl[i].pos[d] = o[d];
    for (int d = 0; d < 3; d++) pl[i].p[d] = o[w.nd + d];
    pl[i].w = o[w.nd + 3];
  }
}
// src arrays are local planes over the two transverse axes (0:n), lower axis fastest; side = 2*axis + is_max
void orc_set_laser_source(void *h, int rk, int side, const double *s1, const double *s2) {
  World &w = *(World *)h;
  Rank &R = w.r[rk];
  const int a = side / 2, sd = side & 1;
  int lo[3] = {1, 1, 1}, hi[3] = {1, 1, 1};
  for (int d = 0; d < w.nd; d++) if (d != a) { lo[d] = 0; hi[d] = R.n[d]; }
  size_t n = 0;
  for (int k = lo[2]; k <= hi[2]; k++)
    for (int j = lo[1]; j <= hi[1]; j++)
      for (int i = lo[0]; i <= hi[0]; i++, n++) {
        if (a == 0) { R.src1[sd](1, j, k) = s1[n]; R.src2[sd](1, j, k) = s2[n]; }
        else {
          int q[3] = {i, j, k};
          q[a] = 1;
          R.srcA[a][sd][0](q[0], q[1], q[2]) = s1[n];
          R.srcA[a][sd][1](q[0], q[1], q[2]) = s2[n];
        }
      }
}
void orc_auto_load(void *h) { auto_load(*(World *)h); }
void orc_init(void *h) {
  World &w = *(World *)h;
  if (w.nd == 1) init_sequence<1>(w);
  else if (w.nd == 2) init_sequence<2>(w);
  else init_sequence<3>(w);
}
void orc_fields_half(void *h) {
  World &w = *(World *)h;
  if (w.nd == 1) update_eb_fields_half<1>(w);
  else if (w.nd == 2) update_eb_fields_half<2>(w);
  else update_eb_fields_half<3>(w);
}
void orc_fields_final(void *h) {
  World &w = *(World *)h;
  if (w.nd == 1) update_eb_fields_final<1>(w);
  else if (w.nd == 2) update_eb_fields_final<2>(w);
  else update_eb_fields_final<3>(w);
}
// push_particles = zero J + push + deposit (+ current_bcs(species) no-op) + particle_bcs
// push_particles with the per-species current_bcs of c_bc_mixed (particles.F90:169-646)
static void push_all(World &w) {
  auto one = [&](int only, bool zero) {
    if (w.nd == 1) push_particles<1>(w, only, zero);
    else if (w.nd == 2) push_particles<2>(w, only, zero);
    else push_particles<3>(w, only, zero);
  };
  if (!w.bc_mixed) { one(-1, true); return; }
  if (w.sp.empty()) { one(-1, true); return; }
  for (size_t is = 0; is < w.sp.size(); is++) {
    one((int)is, is == 0);
    current_bcs_species(w, (int)is);
  }
}
void orc_push(void *h) {
  World &w = *(World *)h;
  push_all(w);
  particle_bcs(w);
}
// push without the trailing particle_bcs (for kernel-level parity checks)
void orc_push_only(void *h) {
  World &w = *(World *)h;
  push_all(w);
}
void orc_particle_bcs(void *h) { particle_bcs(*(World *)h); }
// moving window: shift by `cells` cells (returns -1 for a configuration window.F90 is not restated for)
int orc_shift_window(void *h, int cells) {
  World &w = *(World *)h;
  if (w.cpml_t != 0 || w.periods[0]) return -1;
  shift_window(w, cells);
  return 0;
}
// the particles insert_particles created since the last clear: count, then the data in orc_get_particles' layout
int64_t orc_window_inserted_count(void *h, int rk, int is) { return (int64_t)((World *)h)->r[rk].inserted[is].size(); }
void orc_window_inserted(void *h, int rk, int is, double *out) {
  World &w = *(World *)h;
  const int nd = w.nd, nv = nd + 4;
  const auto &pl = w.r[rk].inserted[is];
  for (size_t i = 0; i < pl.size(); i++) {
    double *o = out + i * nv;
    for (int d = 0; d < nd; d++) o[d] = pl[i].pos[d];
    for (int d = 0; d < 3; d++) o[nd + d] = pl[i].p[d];
    o[nd + 3] = pl[i].w;
  }
}
void orc_window_clear_inserted(void *h) {
  for (Rank &R : ((World *)h)->r)
    for (auto &v : R.inserted) v.clear();
}
// x_grid_min, xb_min, x_min, x_max of the x axis as the window has left them
void orc_window_geometry(void *h, double out[4]) {
  World &w = *(World *)h;
  out[0] = w.grid_min[0]; out[1] = w.xb_min[0]; out[2] = w.cfg.xmin[0]; out[3] = w.cfg.xmax[0];
}
void orc_setup_bc_lists(void *h) { setup_bc_lists(*(World *)h); }
void orc_current_finish(void *h) { current_finish(*(World *)h); }
void orc_efield_bcs(void *h) { efield_bcs(*(World *)h); }
void orc_bfield_bcs(void *h, int mpi_only) { bfield_bcs(*(World *)h, mpi_only != 0); }

// calc_ppc, io/calc_df.F90:761-808 (triangle): cell = FLOOR((pos-x_grid_min_local)/dx + 0.5) + 1
// out has the local interior shape (nx,ny,nz), x fastest; out-of-range particles are skipped
// ext_temp_<side> of one species on one rank: (plane, 3) doubles, see Rank::ext_temp
void orc_set_boundary_temperature(void *h, int rk, int is, int side, const double *t, int64_t n) {
  World &w = *(World *)h;
  Rank &R = w.r[rk];
  if (R.ext_temp.empty()) R.ext_temp.assign(w.sp.size(), std::vector<std::vector<double>>(6));
  R.ext_temp[is][side].assign(t, t + n);
}

void orc_cell_counts(void *h, int rk, int is, int32_t *out) {
  World &w = *(World *)h;
  Rank &R = w.r[rk];
  std::fill(out, out + (size_t)R.n[0] * R.n[1] * R.n[2], 0);
  for (const Particle &P : R.part[is]) {
    int cell[3] = {1, 1, 1};
    bool ok = true;
    for (int d = 0; d < w.nd; d++) {
      cell[d] = (int)std::floor((P.pos[d] - R.grid_min_local[d]) / w.d[d] + 0.5) + 1;
      if (cell[d] < 1 || cell[d] > R.n[d]) ok = false;
    }
    if (ok) out[(size_t)(cell[0] - 1) + (size_t)R.n[0] * ((size_t)(cell[1] - 1) + (size_t)R.n[1] * (cell[2] - 1))]++;
  }
}

// kind 0..2: number / charge / mass density; 3: ekbar; 4: temperature; 5..7: temperature_x/y/z; 8..13: ekflux -x,+x,
// -y,+y,-z,+z; 14..16: average px,py,pz; 17..19: per-species current jx,jy,jz; 20: average weight; 21..23: Poynting flux x,y,z.  Result in WK.
void orc_calc_moment(void *h, int kind, int species) {
  World &w = *(World *)h;
  if (kind <= 2) calc_moment(w, kind, species);
  else if (kind == 3) calc_ratio(w, species, 0);
  else if (kind <= 7) calc_temperature(w, species, kind - 5);
  else if (kind <= 13) calc_ratio(w, species, kind - 7);      // ekflux -x, +x, -y, +y, -z, +z
  else if (kind <= 16) calc_ratio(w, species, kind - 7);      // average momentum px, py, pz (sub 7..9)
  else if (kind <= 19) calc_species_current(w, species, kind - 17);
  else if (kind == 20) calc_average_weight(w, species);
  else calc_poynt_flux(w, kind - 21);
}

// KISS stream check hook: fills out[n] with successive random() values for `seed`
void orc_kiss(int seed, int n, double *out) {
  Rng g;
  g.init(seed);
  for (int i = 0; i < n; i++) out[i] = g.random();
}

}  // extern "C"

// binary collisions (physics_packages/collisions.F90): SURVEY.md 8 f1
#include "collisions_oracle.inc"
