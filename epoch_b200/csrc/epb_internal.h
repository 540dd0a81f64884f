// Internal declarations shared by the translation units of libepoch_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/epoch_b200.h"

#define EPB_NG 5  // triangle shape: png = 3, ng = png + 2 (constants.F90:549-559)

// constants.F90:192-199
#define EPB_C 2.99792458e8
#define EPB_EPS0 8.854187817620389850536563031710750e-12
#define EPB_KB 1.3806488e-23

// Tile geometry of the cell-sorted particle layout.  A tile is TX*TY*TZ cells; the
// sort key of a particle is tile_id * cells_per_tile + cell-in-tile, so every tile
// (one CTA of the push kernel) owns a contiguous, cell-ordered particle range.
struct TileGeom {
  int T[3];    // cells per tile along x,y,z (1 on inactive dims)
  int nt[3];   // tiles along x,y,z
  int cpt;     // cells per tile
  int ntiles;
  int nkeys;   // ntiles * cpt
  int layout;  // 0: cell-major inside a tile; 1 (2D): 8x4-cell warp groups, particles of a group interleaved by rank (see push_cell_2d);
               // 2 (2D, default): slot columns, one fixed-capacity column per cell, no sort at all (see push_slots_2d);
               // 3 (3D, EPB_PUSH3D_VARIANT=1): slot columns in 3D, 16x4x3-cell tiles (push_bag_3d)
};

// Layout 2 ("slot columns", 2D): the particles of cell key k = group * 32 + lane live in rows 0 .. cnt[k]-1 of
// column k; row r of a group's 32 columns is one ROW BLOCK of NC x 32 doubles -- x[32] y[32] px[32] py[32] pz[32]
// w[32], 256-byte aligned each -- at block index group * R + r of ONE allocation, so a round of the owning warp is
// NC aligned, coalesced accesses off a single pointer with immediate offsets, and a particle that stays in its
// cell is never moved by anything but its own lane.  (S.buf[0][q] point at component q of block 0.)  Particles that change cell
// ("movers"), leave the rank, or whose stencil leaves the tile go through the mover buffer M (SoA + a flag
// byte per entry) and are inserted into their next column by k_deliver; what does not fit its column (or M)
// waits in the other M buffer and is pushed by the generic kernel.
// A range of particle slots a diagnostic kernel walks: the classic contiguous [0, n) (cnt == nullptr,
// n_dev == nullptr), the slot arena (cnt != nullptr: slot i is a particle iff its row < cnt[its column]), or a
// mover buffer (n_dev != nullptr: the count lives on the device, R holds the buffer's capacity, entries with
// flag == 1 have left).
struct PRange {
  long long n;             // contiguous: count; arena: nkeys * R slots
  const int *cnt;
  int R;
  const int *n_dev;
  const unsigned char *flag;
  int K;                   // arena: components per row block (the arrays are interleaved row by row, see below); else 0
};
// element offset of slot i inside a component array of the range (arena: row blocks of K x 32 doubles)
__device__ __forceinline__ long long prange_at(const PRange &V, long long i) {
  return V.K ? (((i >> 5) * V.K) << 5) + (i & 31) : i;
}
__device__ __forceinline__ long long prange_n(const PRange &V) {
  if (V.cnt || !V.n_dev) return V.n;
  const long long n = (long long)*V.n_dev;
  return n > V.R ? V.R : n;
}
__device__ __forceinline__ bool prange_valid(const PRange &V, long long i) {
  if (!V.cnt) return !V.flag || V.flag[i] != 1;
  const long long rowslot = i >> 5;            // group * R + r
  const int lane = (int)(i & 31);
  const long long g = rowslot / V.R;
  const int r = (int)(rowslot - g * V.R);
  return r < V.cnt[g * 32 + lane];
}
struct SlotView {
  double *a[7];            // x, y, (z), px, py, pz, w
  PRange r;
};

// Everything the push kernels need, passed by value (__grid_constant__).
struct PushParams {
  int nd;
  int n[3];
  int sz[3];               // array extents incl. ghosts
  const double *e[3];
  const double *b[3];
  double *j[3];
  double idx[3];           // 1/dx
  double grid_min_local[3];
  double dto2, dtco2, third;
  double kfc[3];           // idty/idtx/idxy (2D), idtyz/idtxz/idtxy (3D), idtf/idxf (1D)
  // species (particles.F90:251-256)
  double part_q, part_mc, ipart_mc, cmratio, ccmratio;
  double hc_alpha;         // HC_PUSH: alpha = 0.5 * part_q * dt / part_m (particles.F90:390)
  int deposit;
  int hc_push;             // Higuera-Cary rotation (template flag of the slot-column kernels; push_generic<ND, true> elsewhere)
  // particle SoA
  double *x[3];
  double *p[3];
  double *w;
  long long first, last;   // particle range of a generic launch
  const int *tile_start;   // tiled launch: ntiles+1 offsets into the sorted prefix
  const int *cell_start;   // layout 1: nkeys+1 offsets (exclusive scan of the per-key counts)
  // layout 1, push before a sort: the kernel records where every particle goes in the next
  // order (key_out = next gather cell's key, rank_out = rank inside that key; bit 30 set = rank
  // among the key's arrivals, to be offset by stay_cnt[key]) so the sort needs no atomics
  int emit;
  // layout 1, first push after a sort: the particle of slot i is read from xs/ps/ws[perm[i]] (the
  // old order, other buffer) and written to x/p/w[i]; perm == nullptr: in place
  const int *perm;
  const double *xs[3], *ps[3], *ws;
  int *key_out, *rank_out;
  int *stay_cnt, *arr_cnt;
  long long n_sorted_clip; // tile ranges are clipped to this count
  TileGeom tg;
  // particle boundary conditions (particles.F90:189-221, boundary.F90:1029-1462)
  double bnd_min[3], bnd_max[3];
  double min_local[3], max_local[3];
  double gmin[3], gmax[3], shift[3];
  double min_outer[3], max_outer[3];
  int bc_min[3], bc_max[3];
  int is_bnd_min[3], is_bnd_max[3];
  int nbr_is_self[27];     // neighbour(ix,iy,iz) == this rank
  int nbr_valid[27];
  // outbox: particles that leave this rank (index lists per direction; slot 13 = deleted)
  int *out_count;          // [27]
  int *out_idx;            // [27][out_cap]
  unsigned char *gone;     // per particle flag
  int out_cap;
  int experiment;          // profiling only (EPB_PUSH_EXPERIMENT): disables parts of the tiled kernel
  // layout 2 (slot columns): x/p/w above are the arena; cnt = particles per column, R = rows per column
  int *cnt;
  int R;
  int rowd;                // doubles per row block: 32 x components (interleaved rows) or 32 (one plane per component)
  // mover buffer the kernel appends to: entry m holds a particle that must be (re)inserted by k_deliver
  // (mflag 0), one that left this rank (1, listed in the outbox by its M index) or one that still has to be
  // pushed by the generic kernel (2: stencil outside the tile, or no room in its column)
  double *mx[3], *mp[3], *mw;
  unsigned char *mflag;
  int *mcount;             // device counter (may run past mcap: readers clamp)
  int mcap;
  int *err;                // device error word: bit 0 = mover buffer overflow lost a particle
  // group inboxes (layout 2): a mover whose next column is known goes straight into the inbox of that column's
  // group (64-byte entries: x, y, px, py, pz, w, destination lane, pad -- two full sectors, written by one lane),
  // and the NEXT push reads it from there in the rounds after the column's own rows; ib_in / ic_in are the
  // inbox this push consumes, ib_out / ic_out the one it fills.  nullptr: every mover takes the M path.
  const double *ib_in;
  const int *ic_in;
  double *ib_out;
  int *ic_out;
  int IC;                  // entries per group inbox (<= 256)
};

struct SpeciesDev {
  epb_species cfg;
  long long n = 0;         // attached_list%count
  long long n_sorted = 0;  // [0,n_sorted) follows tile_start
  long long cap = 0;
  double *buf[2][7] = {{0}};  // double buffer for the out-of-place sort: x,y,z,px,py,pz,w
  int cur = 0;
  int *key = nullptr;
  int *tile_start = nullptr;
  int *cell_start = nullptr;  // layout 1: this species' copy of the key offsets
  int *perm = nullptr;        // layout 1: scratch of the emitted sort (destination -> source index)
  int *rank = nullptr;        // layout 1: per-particle rank emitted by the push (see PushParams::emit)
  int *stay_cnt = nullptr, *arr_cnt = nullptr;  // layout 1: per-key counts emitted by the push
  bool info_valid = false;    // key/rank/stay_cnt/arr_cnt describe the current particle set
  bool pending_perm = false;  // the sort left the data in place: perm[new slot] = old index, applied by the next push
  unsigned char *gone = nullptr;
  double *ext_temp[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // thermal walls: ext_temp_x_min ... (plane, 3)
  // layout 2 (slot columns): buf[0] is the arena of nkeys * R slots; n / n_sorted / key / rank / perm are unused
  bool slots = false;
  int R = 0;                  // rows per column
  int rowd = 0;               // doubles per row block (see PushParams)
  int *cnt = nullptr;         // [nkeys] particles per column
  double *mbuf[2][7] = {{0}}; // mover buffers (SoA), mcur = the one the push appends to
  unsigned char *mflag[2] = {nullptr, nullptr};
  int *mcount = nullptr;      // device [2]
  int mcur = 0;
  long long mcap = 0;
  double *arena = nullptr;    // the one allocation behind buf[0][*]
  bool arena_ready = false;   // R chosen and the arena allocated (at the first upload / load, when the density is known)
  double *inbox[2] = {nullptr, nullptr};   // [ngroups][IC][8] group inboxes (see PushParams), ping-pong
  int *icnt[2] = {nullptr, nullptr};       // [ngroups] entries per inbox (may run past IC: readers clamp)
  int icur = 0;                            // the inbox the next push consumes
  int IC = 0;
  bool inbox_dirty = false;                // inbox[icur] may hold particles (between two pushes)
};

// opaque storage of a CUtensorMap (128 bytes, 64-byte aligned), see fdtd_tma.cu
struct alignas(64) TmapStorage { unsigned char b[128]; };

constexpr int EPB_SCAL_MAXSP = 16;   // species covered by epb_step_scalars_async
int epb_allreduce_sum_f64(epb_handle *h, const double *src, double *dst, int n);   // exchange.cu

// push_per_field (shared_data.F90:821): the weight of a particle against a cell in the balancer's load; deck.PUSH_PER_FIELD
constexpr int EPB_PUSH_PER_FIELD = 5;

struct epb_handle {
  epb_config cfg;
  std::vector<SpeciesDev> sp;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sz[3] = {1, 1, 1};
  size_t fsize = 0;             // elements per field array
  double *fields = nullptr;     // 9 * fsize
  size_t plane = 0;             // (ny+2ng)*(nz+2ng)
  double *snap = nullptr;       // [2 sides][6 fields][plane]
  double *src = nullptr;        // [2 sides][2][plane]
  // the same for the y and z faces (index 1, 2): planes over the other two ghosted extents
  size_t planeA[3] = {0, 0, 0};
  double *snapA[3] = {nullptr, nullptr, nullptr};  // [2 sides][6 fields][planeA]
  double *srcA[3] = {nullptr, nullptr, nullptr};   // [2 sides][2][planeA]
  unsigned long long *prof_scratch = nullptr;   // epb_load_profile: local + summed histogram
  size_t prof_cap = 0;
  TileGeom tg;
  int *cell_count = nullptr;    // nkeys + 1
  int *cell_start = nullptr;    // nkeys + 1
  void *cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  int *out_count = nullptr;     // device [27]
  int *out_idx = nullptr;       // device [27][out_cap]
  int out_cap = 0;
  int *h_counts = nullptr;      // pinned [64]
  int *d_scratch = nullptr;     // device ints
  double *coll_work = nullptr;  // collisions with coulomb_log_auto: ekbar of species 1, temperature of species 2 (2 x fsize)
  long long thermal_calls = 0;  // thermal re-emission launches so far (seed of the per-particle streams)
  long long coll_calls = 0;     // epb_collide calls so far (feeds the pair streams' seed)
  double *aos_stage = nullptr;  // 2 Mi particles in the pack_particle wire layout (upload / download staging)
  int *d_err = nullptr;         // device error word (layout 2), checked at the synchronising entry points
  int *movers = nullptr;        // exchange: tail survivors that fill holes (27*out_cap+1)
  // asynchronous field dump: device staging copy + second stream (epb_download_field_async)
  cudaStream_t copy_stream = nullptr;
  double *dump_stage = nullptr;
  cudaEvent_t dump_ready = nullptr, dump_done = nullptr;
  bool dump_pending = false;
  // CPML (boundary.F90:1479-2025), built by epb_create when cfg.cpml_thickness > 0: per axis the kappa profiles on the E and
  // B points and the recursion coefficients bcoeff / ccoeff_d of the half step (device, length n + 2 ng), the layers'
  // local index ranges [axis][side], the laser plane of a cpml_laser face, four auxiliary arrays per axis (field extent)
  bool cpml = false;
  double *cp_kap[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};   // [axis][0 E points, 1 B points]
  double *cp_bco[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  double *cp_cco[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  int cp_start[3][2], cp_end[3][2];
  int cp_laser_idx[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  bool cp_add_laser[3][2] = {{false, false}, {false, false}, {false, false}};
  double *cp_psi[3] = {nullptr, nullptr, nullptr};   // [axis]: 4 x fsize (psi of E_b, E_c, B_b, B_c; (a, b, c) cyclic)
  // epb_step_scalars_async: device block [EPB_SCAL_MAX doubles], a ring of completion events (ticket % 4)
  double *scal_dev = nullptr;
  cudaEvent_t scal_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  long long scal_ticket = 0;
  // epb_set_laser_source: page-locked staging ring, so the call neither waits for the stream nor pins the caller's buffer
  double *src_stage = nullptr;
  size_t src_stage_slot = 0;     // doubles per slot
  cudaEvent_t src_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  long long src_calls = 0;
  // c_bc_mixed (deck_species_block.F90:182-199): the species disagree on some particle boundary, so J is
  // folded / summed / cleared after every species with that species' boundary codes (particles.F90:645)
  bool bc_mixed = false;
  int bc_species = -1;          // species whose codes the current exchange uses (-1: bc_allspecies)
  double *sendbuf = nullptr, *recvbuf = nullptr;  // halo + particle staging
  int xcap[27] = {0};            // records per direction of the fixed-size particle messages (agreed at epb_set_comm)
  size_t sendbuf_elems = 0, recvbuf_elems = 0;
  void *nccl = nullptr;         // ncclComm_t
  std::string err;
  long long launches = 0;
  int pushes_since_sort = 0;
  // push-kernel timing
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  double push_ms_sum = 0.0;
  long long push_ms_n = 0;
  int time_push = 0;
  // TMA descriptors of ex,ey,ez,bx,by,bz for the staged FDTD kernels
  TmapStorage tmap[6];
  bool tma_ok = false;

  double *f(int which) const { return fields + (size_t)which * fsize; }
};

// 2D push kernel (EPB_PUSH_VARIANT): 5 (default) = push_slots_2d<8,3>: one lane per cell, register-resident
// deposit sums, slot-column layout (no sort), 16x8-cell tiles; 3 = push_cell_2d<8,3>: the same deposit on the
// rank-interleaved sorted layout (emitted sort every 2 steps), 168 registers; 2 / 4 = the same kernel on
// 16x16-cell tiles with 128 / 255 registers; 0 = push_tiled_2d (lane per particle, 27-value
// transposed reduction, cell-major layout); 1 = its 21-value form.
// Every kernel-selection / experiment switch of the library is read through epb_env: the variables are honoured
// only when EPB_DEBUG is set to a non-zero value, so a stray variable in a production host's environment cannot
// change kernels or particle loading.  (EPB_DEBUG=1 EPB_PUSH_VARIANT=3 ... is how the A/B lines under profiles/ were made.)
inline const char *epb_env(const char *name) {
  static const int debug = getenv("EPB_DEBUG") ? atoi(getenv("EPB_DEBUG")) : 0;
  return debug ? getenv(name) : nullptr;
}
// a / b for a divisor whose reciprocal rb = 1.0 / b is at hand: the product, then one residual correction through
// two FMAs.  Equal to the IEEE quotient except for rare last-place cases, and free of the slow-path branch of the
// compiler's division, so several of them overlap in the pipeline (moments and collisions: results held to 1e-12).
__device__ __forceinline__ double div_rcp(double a, double b, double rb) {
  const double q = a * rb;
  return fma(fma(-q, b, a), rb, q);
}
inline int epb_push_variant() {
  static int v = -1;
  if (v < 0) { const char *e = epb_env("EPB_PUSH_VARIANT"); v = e ? atoi(e) : 5; }
  return v;
}

// push_*.cu
void epb_launch_push_strict(const PushParams &P, int nd, bool tiled, cudaStream_t s, long long *launches);
void epb_launch_push_fast(const PushParams &P, int nd, bool tiled, cudaStream_t s, long long *launches);

// fdtd_tma.cu
bool epb_fdtd_tma_setup(epb_handle *h);
void epb_fdtd_tma_launch(epb_handle *h, bool is_e, double cx, double cy, double cz, double fac);

// slots.cu (layout 2)
int epb_slots_alloc(epb_handle *h, int is);                       // mover buffers + counts (at create)
void epb_slots_free(SpeciesDev &S);
int epb_slots_reset(epb_handle *h, int is, long long n_expected, int max_ppc_hint);  // empty the species; (re)size the arena
int epb_slots_ensure_rows(epb_handle *h, int is, int R);          // empty species: at least R rows per column
int epb_slots_deliver(epb_handle *h, int is);                     // insert the mover buffer's particles into their columns
int epb_slots_waiting(epb_handle *h, int is, int *waiting);       // entries of the current mover buffer that found their column full
int epb_slots_commit(epb_handle *h, int is, int waiting, long long m);  // m staged particles behind them: flags, count, deliver
int epb_slots_upload(epb_handle *h, int is, int64_t n, const double *packed);
int epb_slots_download(epb_handle *h, int is, int64_t n, double *packed);
int epb_slots_settle(epb_handle *h, int is);                      // group inboxes -> columns (before anything but a push walks the species)
int epb_slots_after_push(epb_handle *h, int is);                  // the consumed inbox is emptied, the filled one becomes current
int epb_slots_count(epb_handle *h, int is, long long *n);         // synchronises
int epb_slots_count_enqueue(epb_handle *h, int is, long long *d_out);   // d_out[0] columns, d_out[1] inbox; no host round trip
// chunked walk over a species' particles as contiguous SoA device arrays, any layout (slots.cu)
struct SpeciesIter {
  std::vector<int> hstart;   // slot columns: scanned column counts
  int k0 = 0, mc = 0;
  long long woff = 0, lin = 0;
};
int epb_species_iter_begin(epb_handle *h, int is, SpeciesIter &I);
int epb_species_iter_next(epb_handle *h, int is, SpeciesIter &I, long long CH, double *st[7], long long *m);
int epb_species_insert_aos(epb_handle *h, int is, const double *aos_dev, long long n);   // device block in the wire layout
int epb_slots_check(epb_handle *h);                               // device error word -> EPB_ERR_CAPACITY (synchronises)
void epb_slots_views(epb_handle *h, int is, SlotView V[2]);       // [0] arena, [1] waiting entries of the mover buffer
int epb_species_views(epb_handle *h, int is, SlotView V[2]);      // any layout: the ranges that hold the species' particles; returns how many
void epb_slots_fill_push(epb_handle *h, int is, PushParams &P);
void epb_launch_push_m_strict(const PushParams &P, cudaStream_t s, long long *launches);
void epb_launch_push_m_fast(const PushParams &P, cudaStream_t s, long long *launches);

// sort.cu
int epb_sort_species(epb_handle *h, int is);
int epb_sort_species_emitted(epb_handle *h, int is);
int epb_apply_pending_perm(epb_handle *h, int is);  // materialise a deferred reordering (gather)
#define EPB_RANK_ARRIVAL (1 << 30)
void epb_make_tiles(const epb_config &cfg, TileGeom &tg);

// exchange.cu
int epb_particle_exchange(epb_handle *h, int is);
int epb_halo_exchange(epb_handle *h, int f0, int nf, bool add);
int epb_comm_init(epb_handle *h, const void *id128);
void epb_comm_destroy(epb_handle *h);

int epb_fail(epb_handle *h, int code, const char *fmt, ...);
#define EPB_CUDA(h, call)                                                              \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return epb_fail((h), EPB_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,  \
                      cudaGetErrorString(e_));                                         \
  } while (0)
