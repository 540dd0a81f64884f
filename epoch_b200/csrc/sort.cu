// sort.cu — on-GPU counting sort of a species into tile-major cell order.
//
// Replaces the reference's per-particle heap nodes (TYPE particle_list,
// shared_data.F90:159-171) and supersedes reorder_particles_to_grid
// (housekeeping/split_particle.F90:29-77): after the sort, every tile of the push
// kernel owns one contiguous particle range and, inside it, particles of the same
// cell are adjacent.  Out of place (double buffer), three passes:
//   keys + histogram  ->  exclusive scan  ->  scatter.
#include <cub/cub.cuh>

#include "epb_internal.h"

namespace {

struct KeyOp {
  const double *x[3];
  const double *p[3];      // layout 1: momenta, for the predicted gather cell
  double ipart_mc, dtco2, idx[3];
  int predict;
  long long n;
  int nd, nloc[3];
  double gmin[3], dx[3];
  TileGeom tg;
  int *key;
  int *count;
};

__device__ __forceinline__ int cell_key(const KeyOp &K, long long i) {
  int t[3] = {0, 0, 0}, in[3] = {0, 0, 0};
  double root = 0.0;
  if (K.predict) {
    // the cell the NEXT push gathers in: position advanced by half a step with the stored
    // momentum, exactly as particles.F90:289-322 does at the top of push_particles
    const double ux = K.p[0][i] * K.ipart_mc, uy = K.p[1][i] * K.ipart_mc, uz = K.p[2][i] * K.ipart_mc;
    root = K.dtco2 / sqrt(ux * ux + uy * uy + uz * uz + 1.0);
  }
  for (int d = 0; d < K.nd; d++) {
    int cell;
    if (K.predict) {
      double part_x = K.x[d][i] - K.gmin[d];
      part_x = part_x + (K.p[d][i] * K.ipart_mc) * root;
      cell = __double2int_rd(part_x * K.idx[d] + 0.5);
    } else {
      // nearest cell as calc_ppc defines it (io/calc_df.F90:795-796), clamped into 1..n
      cell = __double2int_rd((K.x[d][i] - K.gmin[d]) / K.dx[d] + 0.5);
    }
    cell = cell < 0 ? 0 : (cell > K.nloc[d] - 1 ? K.nloc[d] - 1 : cell);
    t[d] = cell / K.tg.T[d];
    in[d] = cell - t[d] * K.tg.T[d];
  }
  const int tile = (t[2] * K.tg.nt[1] + t[1]) * K.tg.nt[0] + t[0];
  // layout 1 (2D, 16-wide tiles): a warp group of push_cell_2d = two tile rows = 32 consecutive keys
  const int intile = (in[2] * K.tg.T[1] + in[1]) * K.tg.T[0] + in[0];
  return tile * K.tg.cpt + intile;
}

// warp-aggregated increment: lanes with equal keys share one atomic
__device__ __forceinline__ int agg_inc(int *ctr, int key, bool active) {
  const int lane = threadIdx.x & 31;
  const unsigned m = __match_any_sync(0xffffffffu, active ? key : -1 - lane);
  int base = 0;
  if (active) {
    const int leader = __ffs(m) - 1;
    if (lane == leader) base = atomicAdd(&ctr[key], __popc(m));
    base = __shfl_sync(m, base, leader);
    base += __popc(m & ((1u << lane) - 1u));
  }
  return base;
}

__global__ void __launch_bounds__(256) k_keys(const __grid_constant__ KeyOp K) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long nround = (K.n + 31) / 32 * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    const bool active = i < K.n;
    int key = 0;
    if (active) { key = cell_key(K, i); K.key[i] = key; }
    (void)agg_inc(K.count, key, active);
  }
}

struct ScatterOp {
  const double *src[7];
  double *dst[7];
  long long n;
  const int *key;
  const int *start;
  int *cursor;
};
__global__ void __launch_bounds__(256) k_scatter(const __grid_constant__ ScatterOp S) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long nround = (S.n + 31) / 32 * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    const bool active = i < S.n;
    const int key = active ? S.key[i] : 0;
    const int r = agg_inc(S.cursor, key, active);
    if (active) {
      const long long dst = (long long)S.start[key] + r;
#pragma unroll
      for (int q = 0; q < 7; q++)
        if (S.src[q]) S.dst[q][dst] = S.src[q][i];
    }
  }
}

// Layout 1: inside a group of 32 keys the particles are interleaved by their rank r within the
// key: all rank-0 particles of the group in key order, then all rank-1 particles, ... (keys
// that have run out are skipped).  A warp of push_cell_2d whose lane l owns key l then reads
// one round with a single coalesced load.  Position of (key l, rank r) inside the group =
// sum over l' of min(count[l'], r + (l' < l)).
__global__ void __launch_bounds__(256) k_scatter_il(const __grid_constant__ ScatterOp S) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long nround = (S.n + 31) / 32 * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    const bool active = i < S.n;
    const int key = active ? S.key[i] : 0;
    const int r = agg_inc(S.cursor, key, active);
    if (active) {
      const int l = key & 31;
      const int *st = S.start + (key - l);
      const int s0 = __ldg(st);
      int prev = s0, acc = 0;
#pragma unroll 8
      for (int q = 0; q < 32; q++) {
        const int nxt = __ldg(st + q + 1);
        const int c = nxt - prev;
        prev = nxt;
        const int lim = r + (q < l ? 1 : 0);
        acc += c < lim ? c : lim;
      }
      const long long dst = (long long)s0 + acc;
#pragma unroll
      for (int q = 0; q < 7; q++)
        if (S.src[q]) S.dst[q][dst] = S.src[q][i];
    }
  }
}

// ---- sort from the records the push emitted (PushParams::emit) ---------------------------------
// Particles without a record (pushed by the generic kernel, arrived from a neighbour, wrapped)
// get the predicted key here and a rank among their key's arrivals.
__global__ void __launch_bounds__(256) k_keys_fix(const __grid_constant__ KeyOp K, int *rank, int *arr_cnt) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long nround = (K.n + 31) / 32 * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    const bool missing = i < K.n && K.key[i] < 0;
    if (!__any_sync(0xffffffffu, missing)) continue;
    int key = 0;
    if (missing) { key = cell_key(K, i); K.key[i] = key; }
    const int r = agg_inc(arr_cnt, key, missing);
    if (missing) rank[i] = r | EPB_RANK_ARRIVAL;
  }
}
__global__ void __launch_bounds__(256) k_add_counts(const int *a, const int *b, int *out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = a[i] + b[i];
}
// The emitted ranks of a warp's 32 source particles differ by the number of particles that
// left each cell before them, so a direct scatter writes partial sectors (measured: 21 sectors
// per store, 2x DRAM traffic).  Instead the 4-byte source index is scattered (perm[dst] = src)
// and the seven arrays are then gathered with fully coalesced stores.
struct PermOp {
  long long n;
  const int *key, *rank, *stay, *start;
  int *perm;
};
__global__ void __launch_bounds__(256) k_perm_emitted(const __grid_constant__ PermOp S) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += stride) {
    const int key = S.key[i];
    int r = S.rank[i];
    if (r & EPB_RANK_ARRIVAL) r = (r & ~EPB_RANK_ARRIVAL) + __ldg(S.stay + key);
    const int l = key & 31;
    const int *st = S.start + (key - l);
    const int s0 = __ldg(st);
    int prev = s0, acc = 0;
#pragma unroll 8
    for (int q = 0; q < 32; q++) {
      const int nxt = __ldg(st + q + 1);
      const int c = nxt - prev;
      prev = nxt;
      const int lim = r + (q < l ? 1 : 0);
      acc += c < lim ? c : lim;
    }
    S.perm[(long long)s0 + acc] = (int)i;
  }
}
struct GatherOp {
  const double *src[7];
  double *dst[7];
  long long n;
  const int *perm;
};
__global__ void __launch_bounds__(256) k_gather_perm(const __grid_constant__ GatherOp S) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < S.n; j += stride) {
    const int i = S.perm[j];
#pragma unroll
    for (int q = 0; q < 7; q++)
      if (S.src[q]) S.dst[q][j] = S.src[q][i];
  }
}

__global__ void k_tile_start(const int *cell_start, int *tile_start, int ntiles, int cpt) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t <= ntiles) tile_start[t] = cell_start[(size_t)t * cpt];
}

}  // namespace

void epb_make_tiles(const epb_config &cfg, TileGeom &tg) {
  const int nd = cfg.ndims;
  int T[3] = {1, 1, 1};
  if (nd == 1) T[0] = 256;
  else if (nd == 2) { T[0] = 16; T[1] = 16; }  // must match T2X/T2Y in push.cuh
  else { T[0] = 8; T[1] = 8; T[2] = 4; }  // must match T3 / T3Z in push.cuh
  const int variant = epb_push_variant();
  if (nd == 2 && (variant == 3 || variant == 5)) T[1] = 8;      // push_cell_2d<8,3> / push_slots_2d<8,3>: 16x8-cell tiles
  tg.layout = (nd == 2 && variant >= 2 && variant <= 4) ? 1 : 0;
  // slot columns (the Higuera-Cary rotation is a template flag of the same kernels)
  if (nd == 2 && variant == 5) tg.layout = 2;
  // 3D: tile bags (layout 3, push_bag_3d) unless EPB_PUSH3D_VARIANT=0 asks for the sorted layout of push_tiled_3d
  static const int v3 = epb_env("EPB_PUSH3D_VARIANT") ? atoi(epb_env("EPB_PUSH3D_VARIANT")) : 1;
  if (nd == 3 && v3 != 0) {
    tg.layout = 3;
    T[0] = 16; T[1] = 4; T[2] = 3;   // must match B3X / B3Y / B3Z in push.cuh (half-warp = 16 consecutive cells of a row)
  }
  tg.cpt = 1;
  tg.ntiles = 1;
  for (int d = 0; d < 3; d++) {
    tg.T[d] = T[d];
    tg.nt[d] = d < nd ? (cfg.n[d] + T[d] - 1) / T[d] : 1;
    tg.cpt *= T[d];
    tg.ntiles *= tg.nt[d];
  }
  tg.nkeys = tg.ntiles * tg.cpt;
}

static void fill_keyop(epb_handle *h, SpeciesDev &S, KeyOp &K) {
  const epb_config &c = h->cfg;
  for (int d = 0; d < 3; d++) {
    K.x[d] = S.buf[S.cur][d];
    K.nloc[d] = c.n[d];
    K.gmin[d] = c.grid_min_local[d];
    K.dx[d] = c.dx[d];
    K.p[d] = S.buf[S.cur][3 + d];
    K.idx[d] = d < c.ndims ? 1.0 / c.dx[d] : 0.0;
  }
  K.n = S.n;
  K.nd = c.ndims;
  K.tg = h->tg;
  K.key = S.key;
  K.count = h->cell_count;
  K.predict = (h->tg.layout == 1);
  K.ipart_mc = 1.0 / (EPB_C * S.cfg.mass);
  K.dtco2 = EPB_C * (c.dt / 2.0);
}

// Sort from the records of the last push (layout 1): fix-up of the particles without a record,
// counts = stayers + arrivals, scan, atomic-free scatter.
int epb_apply_pending_perm(epb_handle *h, int is) {
  SpeciesDev &S = h->sp[is];
  if (!S.pending_perm) return EPB_OK;
  S.pending_perm = false;
  if (S.n <= 0) return EPB_OK;
  GatherOp Ga;
  for (int q = 0; q < 7; q++) { Ga.src[q] = S.buf[S.cur][q]; Ga.dst[q] = S.buf[S.cur ^ 1][q]; }
  Ga.n = S.n;
  Ga.perm = S.perm;
  long long nb = (S.n + 255) / 256;
  if (nb > 148LL * 32) nb = 148LL * 32;
  k_gather_perm<<<(int)nb, 256, 0, h->stream>>>(Ga);
  h->launches++;
  S.cur ^= 1;
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

int epb_sort_species_emitted(epb_handle *h, int is) {
  SpeciesDev &S = h->sp[is];
  const int nkeys = h->tg.nkeys;
  {
    int rc = epb_apply_pending_perm(h, is);
    if (rc) return rc;
  }
  if (S.n > 0) {
    KeyOp K;
    fill_keyop(h, S, K);
    long long nb = (S.n + 255) / 256;
    if (nb > 148LL * 32) nb = 148LL * 32;
    k_keys_fix<<<(int)nb, 256, 0, h->stream>>>(K, S.rank, S.arr_cnt);
    h->launches++;
  }
  k_add_counts<<<148 * 8, 256, 0, h->stream>>>(S.stay_cnt, S.arr_cnt, h->cell_count, nkeys + 1);
  h->launches++;
  size_t need = 0;
  EPB_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, need, h->cell_count, S.cell_start, nkeys + 1, h->stream));
  if (need > h->cub_tmp_bytes || !h->cub_tmp) {
    cudaFree(h->cub_tmp);
    EPB_CUDA(h, cudaMalloc(&h->cub_tmp, need));
    h->cub_tmp_bytes = need;
  }
  EPB_CUDA(h, cub::DeviceScan::ExclusiveSum(h->cub_tmp, need, h->cell_count, S.cell_start, nkeys + 1, h->stream));
  h->launches++;
  if (S.n > 0) {
    PermOp Pm;
    Pm.n = S.n;
    Pm.key = S.key;
    Pm.rank = S.rank;
    Pm.stay = S.stay_cnt;
    Pm.start = S.cell_start;
    Pm.perm = S.perm;
    long long nb = (S.n + 255) / 256;
    if (nb > 148LL * 32) nb = 148LL * 32;
    k_perm_emitted<<<(int)nb, 256, 0, h->stream>>>(Pm);
    h->launches++;
    // the gather itself is left to the next push of this species (PushParams::perm), which reads
    // the old order through perm and writes the new one: no separate 100 B/particle pass
    S.pending_perm = true;
    static const int no_fuse = epb_env("EPB_NO_FUSED_GATHER") ? atoi(epb_env("EPB_NO_FUSED_GATHER")) : 0;
    if (no_fuse) {
      int rc = epb_apply_pending_perm(h, is);
      if (rc) return rc;
    }
  }
  k_tile_start<<<(h->tg.ntiles + 1 + 255) / 256, 256, 0, h->stream>>>(S.cell_start, S.tile_start, h->tg.ntiles, h->tg.cpt);
  h->launches++;
  S.n_sorted = S.n;
  S.info_valid = false;
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

int epb_sort_species(epb_handle *h, int is) {
  SpeciesDev &S = h->sp[is];
  const epb_config &c = h->cfg;
  const int nkeys = h->tg.nkeys;
  S.info_valid = false;
  {
    int rc = epb_apply_pending_perm(h, is);
    if (rc) return rc;
  }
  EPB_CUDA(h, cudaMemsetAsync(h->cell_count, 0, ((size_t)nkeys + 1) * sizeof(int), h->stream));
  if (S.n > 0) {
    KeyOp K;
    for (int d = 0; d < 3; d++) {
      K.x[d] = S.buf[S.cur][d];
      K.nloc[d] = c.n[d];
      K.gmin[d] = c.grid_min_local[d];
      K.dx[d] = c.dx[d];
    }
    K.n = S.n;
    K.nd = c.ndims;
    K.tg = h->tg;
    K.key = S.key;
    K.count = h->cell_count;
    static const int predict_env = epb_env("EPB_SORT_PREDICT") ? atoi(epb_env("EPB_SORT_PREDICT")) : 0;
    K.predict = (h->tg.layout == 1) || predict_env;
    for (int d = 0; d < 3; d++) {
      K.p[d] = S.buf[S.cur][3 + d];
      K.idx[d] = d < c.ndims ? 1.0 / c.dx[d] : 0.0;
    }
    K.ipart_mc = 1.0 / (EPB_C * S.cfg.mass);
    K.dtco2 = EPB_C * (c.dt / 2.0);
    long long nb = (S.n + 255) / 256;
    if (nb > 148LL * 32) nb = 148LL * 32;
    k_keys<<<(int)nb, 256, 0, h->stream>>>(K);
    h->launches++;
  }
  int *cstart = (h->tg.layout == 1 && S.cell_start) ? S.cell_start : h->cell_start;
  size_t need = 0;
  // (checked: a stale error picked up by the size query would leave need = 0 and turn the real call below
  // into a second size query -- the scan would silently not run)
  EPB_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, need, h->cell_count, cstart, nkeys + 1, h->stream));
  if (need > h->cub_tmp_bytes || !h->cub_tmp) {
    cudaFree(h->cub_tmp);
    EPB_CUDA(h, cudaMalloc(&h->cub_tmp, need));
    h->cub_tmp_bytes = need;
  }
  EPB_CUDA(h, cub::DeviceScan::ExclusiveSum(h->cub_tmp, need, h->cell_count, cstart, nkeys + 1, h->stream));
  h->launches++;
  if (S.n > 0) {
    EPB_CUDA(h, cudaMemsetAsync(h->cell_count, 0, ((size_t)nkeys + 1) * sizeof(int), h->stream));
    ScatterOp Sc;
    for (int q = 0; q < 7; q++) { Sc.src[q] = S.buf[S.cur][q]; Sc.dst[q] = S.buf[S.cur ^ 1][q]; }
    Sc.n = S.n;
    Sc.key = S.key;
    Sc.start = cstart;
    Sc.cursor = h->cell_count;
    long long nb = (S.n + 255) / 256;
    if (nb > 148LL * 32) nb = 148LL * 32;
    if (h->tg.layout == 1) k_scatter_il<<<(int)nb, 256, 0, h->stream>>>(Sc);
    else k_scatter<<<(int)nb, 256, 0, h->stream>>>(Sc);
    h->launches++;
    S.cur ^= 1;
  }
  k_tile_start<<<(h->tg.ntiles + 1 + 255) / 256, 256, 0, h->stream>>>(cstart, S.tile_start, h->tg.ntiles, h->tg.cpt);
  h->launches++;
  S.n_sorted = S.n;
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}
