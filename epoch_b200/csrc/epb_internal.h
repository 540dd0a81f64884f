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
  int layout;  // 0: cell-major inside a tile; 1 (2D): 8x4-cell warp groups, particles of a group interleaved by rank (see push_cell_2d)
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
  int hc_push;             // Higuera-Cary rotation: every particle goes through push_generic<ND, true>
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
};

// opaque storage of a CUtensorMap (128 bytes, 64-byte aligned), see fdtd_tma.cu
struct alignas(64) TmapStorage { unsigned char b[128]; };

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
  int *movers = nullptr;        // exchange: tail survivors that fill holes (27*out_cap+1)
  // asynchronous field dump: device staging copy + second stream (epb_download_field_async)
  cudaStream_t copy_stream = nullptr;
  double *dump_stage = nullptr;
  cudaEvent_t dump_ready = nullptr, dump_done = nullptr;
  bool dump_pending = false;
  // c_bc_mixed (deck_species_block.F90:182-199): the species disagree on some particle boundary, so J is
  // folded / summed / cleared after every species with that species' boundary codes (particles.F90:645)
  bool bc_mixed = false;
  int bc_species = -1;          // species whose codes the current exchange uses (-1: bc_allspecies)
  double *sendbuf = nullptr, *recvbuf = nullptr;  // halo + particle staging
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

// 2D push kernel (EPB_PUSH_VARIANT): 3 (default) = push_cell_2d<8,3>: one lane per cell, register-resident
// deposit sums, rank-interleaved layout, 16x8-cell tiles, 168 registers; 2 / 4 = the same kernel on
// 16x16-cell tiles with 128 / 255 registers; 0 = push_tiled_2d (lane per particle, 27-value
// transposed reduction, cell-major layout); 1 = its 21-value form.
inline int epb_push_variant() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("EPB_PUSH_VARIANT"); v = e ? atoi(e) : 3; }
  return v;
}

// push_*.cu
void epb_launch_push_strict(const PushParams &P, int nd, bool tiled, cudaStream_t s, long long *launches);
void epb_launch_push_fast(const PushParams &P, int nd, bool tiled, cudaStream_t s, long long *launches);

// fdtd_tma.cu
bool epb_fdtd_tma_setup(epb_handle *h);
void epb_fdtd_tma_launch(epb_handle *h, bool is_e, double cx, double cy, double cz, double fac);

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
