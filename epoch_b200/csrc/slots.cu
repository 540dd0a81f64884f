// slots.cu — layout 2 ("slot columns") of the 2D particle store: everything around push_slots_2d.
//
// Replaces, for 2D species, the cell-sorted contiguous arrays + periodic counting sort (sort.cu) that stood in
// for the reference's linked lists (TYPE particle_list, shared_data.F90:159-171; reorder_particles_to_grid,
// housekeeping/split_particle.F90:29-77): every cell owns a fixed-capacity column of slots, the push keeps a
// particle that stays in its cell where it is, and only the few per cent that change cell are moved -- through
// the mover buffer, by k_deliver below.  There is no sort and no second copy of the particle state.
//
//   arena  : 7 SoA arrays of nkeys * R slots; slot of (key k, row r) = ((k >> 5) * R + r) * 32 + (k & 31)
//   cnt    : particles per column (rows 0 .. cnt-1 are particles, packed)
//   M, M'  : mover buffers (SoA + flag byte + device counter).  After a push M holds the step's movers
//            (flag 0), the particles that left the rank (flag 1, indexed by the outbox) and arrivals from the
//            neighbours (flag 0); k_deliver inserts the flag-0 entries into their columns, what finds its
//            column full goes to M' with flag 2 and is pushed by the generic kernel next step.  M and M' swap.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "epb_internal.h"

namespace {

// components per particle = nd + 4 (2D: x y px py pz w; 3D: x y z px py pz w): a row block is NC x 32 doubles

inline int nblk(size_t n, int cap = 148 * 16) {
  size_t b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > (size_t)cap) b = cap;
  return (int)b;
}

struct DeliverOp {
  // source: mover buffer
  const double *sx[3], *sp[3], *sw;
  const unsigned char *sflag;
  const int *scount;
  int scap;
  // destination: arena + counts
  double *ax[3], *ap[3], *aw;
  int *cnt;
  int R;
  // overflow: the other mover buffer
  double *ox[3], *op[3], *ow;
  unsigned char *oflag;
  int *ocount;
  int ocap;
  int *err;
  // predicted gather cell (particles.F90:289-322, the top of the next push)
  int nd, nloc[3];
  double gmin[3], idx[3], ipart_mc, dtco2;
  TileGeom tg;
  int rowd;        // doubles per row block (32 x components when the arrays are interleaved row by row, else 32)
};

// The cell the NEXT push gathers this particle in: position advanced by half a step with the stored
// momentum, operation for operation what push_slots_2d does at the top of its round, so the particle meets
// the lane that owns its stencil.  Clamped into the interior (a particle within half a step of the rank's
// edge can gather in the first ghost cell: it lives in the edge column and takes the general deposit there).
__device__ __forceinline__ int predicted_key(const DeliverOp &D, const double *x, double px, double py, double pz) {
  const double ux = px * D.ipart_mc, uy = py * D.ipart_mc, uz = pz * D.ipart_mc;
  const double root = D.dtco2 / sqrt(ux * ux + uy * uy + uz * uz + 1.0);
  const double u[3] = {ux, uy, uz};
  int cell[3] = {0, 0, 0};
  for (int d = 0; d < D.nd; d++) {
    double part_x = x[d] - D.gmin[d];
    part_x = part_x + u[d] * root;
    int c = __double2int_rd(part_x * D.idx[d] + 0.5);
    cell[d] = c < 0 ? 0 : (c > D.nloc[d] - 1 ? D.nloc[d] - 1 : c);
  }
  const int tx = cell[0] / D.tg.T[0], ty = cell[1] / D.tg.T[1], tz = cell[2] / D.tg.T[2];
  const int tile = (tz * D.tg.nt[1] + ty) * D.tg.nt[0] + tx;
  // a column per cell: 32 consecutive keys are 16 x 2 cells of one z plane
  return tile * D.tg.cpt + ((cell[2] - tz * D.tg.T[2]) * D.tg.T[1] + (cell[1] - ty * D.tg.T[1])) * D.tg.T[0] +
         (cell[0] - tx * D.tg.T[0]);
}

__global__ void __launch_bounds__(256) k_deliver(const __grid_constant__ DeliverOp D) {
  int n = *D.scount;
  if (n > D.scap) n = D.scap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (D.sflag[i] == 1) continue;   // left this rank or the system
    double x[3] = {0.0, 0.0, 0.0};
    for (int d = 0; d < D.nd; d++) x[d] = D.sx[d][i];
    const double px = D.sp[0][i], py = D.sp[1][i], pz = D.sp[2][i], w = D.sw[i];
    const int key = predicted_key(D, x, px, py, pz);
    const int r = atomicAdd(&D.cnt[key], 1);
    if (r < D.R) {
      const size_t o = ((size_t)(key >> 5) * D.R + r) * D.rowd + (key & 31);
      for (int d = 0; d < D.nd; d++) D.ax[d][o] = x[d];
      D.ap[0][o] = px; D.ap[1][o] = py; D.ap[2][o] = pz;
      D.aw[o] = w;
    } else {
      // column full: every thread that saw r >= R takes its increment back, so the count settles at R
      atomicSub(&D.cnt[key], 1);
      const int m = atomicAdd(D.ocount, 1);
      if (m < D.ocap) {
        for (int d = 0; d < D.nd; d++) D.ox[d][m] = x[d];
        D.op[0][m] = px; D.op[1][m] = py; D.op[2][m] = pz;
        D.ow[m] = w;
        D.oflag[m] = 2;
      } else {
        atomicOr(D.err, 1);
      }
    }
  }
}

struct SettleOp {
  const double *ib;
  const int *ic;
  int IC, ngroups, nd, rowd;
  double *ax[3], *ap[3], *aw;
  int *cnt;
  int R;
  double *ox[3], *op[3], *ow;
  unsigned char *oflag;
  int *ocount;
  int ocap;
  int *err;
};
// inbox entry: nd position components, px, py, pz, w, then the destination lane (as an integer's bits) in double nd + 4
__global__ void __launch_bounds__(256) k_settle(const __grid_constant__ SettleOp O) {
  // one warp per group inbox
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < O.ngroups; g += warps) {
    int n = O.ic[g];
    if (n > O.IC) n = O.IC;
    for (int j = lane; j < n; j += 32) {
      const double *e = O.ib + ((size_t)g * O.IC + j) * 8;
      const int key = g * 32 + ((int)__double_as_longlong(e[O.nd + 4]) & 31);
      const int r = atomicAdd(&O.cnt[key], 1);
      if (r < O.R) {
        const size_t o = ((size_t)g * O.R + r) * O.rowd + (key & 31);
        for (int d = 0; d < O.nd; d++) O.ax[d][o] = e[d];
        O.ap[0][o] = e[O.nd]; O.ap[1][o] = e[O.nd + 1]; O.ap[2][o] = e[O.nd + 2];
        O.aw[o] = e[O.nd + 3];
      } else {
        atomicSub(&O.cnt[key], 1);
        const int m = atomicAdd(O.ocount, 1);
        if (m < O.ocap) {
          for (int d = 0; d < O.nd; d++) O.ox[d][m] = e[d];
          O.op[0][m] = e[O.nd]; O.op[1][m] = e[O.nd + 1]; O.op[2][m] = e[O.nd + 2];
          O.ow[m] = e[O.nd + 3];
          O.oflag[m] = 2;
        } else {
          atomicOr(O.err, 1);
        }
      }
    }
  }
}

// arena -> contiguous SoA staging (download): columns [k0, k1), offsets from the scanned counts
struct CompactOp {
  const double *a[7];
  double *dst[7];
  const int *cnt, *start;
  int R, k0, k1, rowd;
  long long base;   // start[k0]
};
__global__ void __launch_bounds__(256) k_compact(const __grid_constant__ CompactOp C) {
  // one warp per group of 32 columns, row by row (coalesced reads)
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int g = (C.k0 >> 5) + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); g < (C.k1 >> 5); g += warps) {
    const int key = g * 32 + lane;
    const int c = min(C.cnt[key], C.R);
    const long long st = (long long)C.start[key] - C.base;
    const int mx = __reduce_max_sync(0xffffffffu, c);
    for (int r = 0; r < mx; r++) {
      if (r < c) {
        const size_t o = ((size_t)g * C.R + r) * C.rowd + lane;
#pragma unroll
        for (int q = 0; q < 7; q++)
          if (C.a[q]) C.dst[q][st + r] = C.a[q][o];
      }
    }
  }
}

// pack_particle wire layout (partlist.F90:414-486: pos(1:ndims), p(1:3), weight) <-> the SoA arrays
struct AosOp {
  double *soa[7];
  double *aos;
  long long n;
  int nd, nv;
  int to_soa;
};
__global__ void __launch_bounds__(256) k_aos(const __grid_constant__ AosOp A) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += (long long)gridDim.x * blockDim.x) {
    double *o = A.aos + i * A.nv;
    int q = 0;
    if (A.to_soa) {
      for (int d = 0; d < A.nd; d++) A.soa[d][i] = o[q++];
      for (int d = 3; d < 7; d++) A.soa[d][i] = o[q++];
    } else {
      for (int d = 0; d < A.nd; d++) o[q++] = A.soa[d][i];
      for (int d = 3; d < 7; d++) o[q++] = A.soa[d][i];
    }
  }
}

__global__ void k_clamp_counts(const int *cnt, int *out, int n, int R) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = min(cnt[i], R);
}
__global__ void k_bump(int *count_dev, int n) { *count_dev += n; }

}  // namespace

// ---------------------------------------------------------------------------------------------
int epb_slots_alloc(epb_handle *h, int is) {
  SpeciesDev &S = h->sp[is];
  const int nd = h->cfg.ndims;
  S.slots = true;
  S.mcap = std::max<long long>(65536, S.cap / 8);
  if (S.mcap > S.cap && S.cap >= 4096) S.mcap = S.cap;
  for (int b = 0; b < 2; b++) {
    for (int q = 0; q < 7; q++) {
      if (q < 3 && q >= nd) continue;
      EPB_CUDA(h, cudaMalloc(&S.mbuf[b][q], (size_t)S.mcap * sizeof(double)));
    }
    EPB_CUDA(h, cudaMalloc(&S.mflag[b], (size_t)S.mcap));
    EPB_CUDA(h, cudaMemsetAsync(S.mflag[b], 0, (size_t)S.mcap, h->stream));
  }
  EPB_CUDA(h, cudaMalloc(&S.mcount, 2 * sizeof(int)));
  EPB_CUDA(h, cudaMemsetAsync(S.mcount, 0, 2 * sizeof(int), h->stream));
  EPB_CUDA(h, cudaMalloc(&S.cnt, ((size_t)h->tg.nkeys + 1) * sizeof(int)));
  EPB_CUDA(h, cudaMemsetAsync(S.cnt, 0, ((size_t)h->tg.nkeys + 1) * sizeof(int), h->stream));
  if (!h->d_err) {
    EPB_CUDA(h, cudaMalloc(&h->d_err, sizeof(int)));
    EPB_CUDA(h, cudaMemsetAsync(h->d_err, 0, sizeof(int), h->stream));
  }
  // group inboxes (EPB_SLOTS_INBOX=0 sends every mover through the mover buffer and k_deliver instead)
  static const int use_inbox = epb_env("EPB_SLOTS_INBOX") ? atoi(epb_env("EPB_SLOTS_INBOX")) : 1;
  if (use_inbox) {
    const size_t ngroups = (size_t)h->tg.nkeys / 32;
    // entries per group inbox: a step's arrivals of 32 columns.  192 covers 64 particles per cell at ~9 % movers;
    // runs with few particles per cell (3D: 8) get proportionally less (the mover buffer takes what does not fit)
    {
      long long ncell = 1;
      for (int d = 0; d < nd; d++) ncell *= h->cfg.n[d];
      const long long ppc = S.cap / std::max<long long>(1, ncell);
      S.IC = (int)std::min<long long>(192, std::max<long long>(48, 4 * ppc));
    }
    for (int b = 0; b < 2; b++) {
      EPB_CUDA(h, cudaMalloc(&S.inbox[b], ngroups * (size_t)S.IC * 8 * sizeof(double)));
      EPB_CUDA(h, cudaMalloc(&S.icnt[b], (ngroups + 1) * sizeof(int)));
      EPB_CUDA(h, cudaMemsetAsync(S.icnt[b], 0, (ngroups + 1) * sizeof(int), h->stream));
    }
  }
  return EPB_OK;
}

void epb_slots_free(SpeciesDev &S) {
  for (int b = 0; b < 2; b++) {
    for (int q = 0; q < 7; q++) cudaFree(S.mbuf[b][q]);
    cudaFree(S.mflag[b]);
  }
  cudaFree(S.mcount);
  cudaFree(S.cnt);
  cudaFree(S.arena);
  for (int q = 0; q < 7; q++) S.buf[0][q] = nullptr;   // they pointed into the arena
  for (int b = 0; b < 2; b++) { cudaFree(S.inbox[b]); cudaFree(S.icnt[b]); }
}

// Empty the species and make sure the arena fits what is about to be loaded: rows per column R from the
// densest cell expected (max_ppc_hint) and the mean of the occupied ones, capped by the memory the host
// reserved (2 x capacity slots: what the sorted layout's double buffer took).
static int slots_set_rows(epb_handle *h, int is, long long R) {
  SpeciesDev &S = h->sp[is];
  const int nd = h->cfg.ndims;
  const long long nkeys = h->tg.nkeys;
  const int NC = nd + 4;
  if ((size_t)R * (size_t)nkeys * NC >= ((size_t)1 << 38)) return epb_fail(h, EPB_ERR_CAPACITY, "species %d: slot arena too large", is);
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  // one allocation of nkeys / 32 * R row blocks of NC x 32 doubles; buf[0][q] = component q of block 0
  cudaFree(S.arena);
  S.arena = nullptr;
  for (int q = 0; q < 7; q++) S.buf[0][q] = nullptr;
  EPB_CUDA(h, cudaMalloc(&S.arena, (size_t)R * nkeys * NC * sizeof(double)));
  {
    // 2D default: the components as separate planes of R * nkeys doubles (measured 0.3 - 2.4 % faster than interleaved
    // row blocks on the same box, profiles/r02_call12_*); EPB_SLOTS_ROWBLOCK=1 selects the row blocks, which 3D always uses
    static const int rowblock = epb_env("EPB_SLOTS_ROWBLOCK") ? atoi(epb_env("EPB_SLOTS_ROWBLOCK")) : 0;
    const bool rb = rowblock != 0 || nd == 3;
    S.rowd = rb ? 32 * NC : 32;
    const size_t cstride = rb ? 32 : (size_t)R * nkeys;
    int cq = 0;
    for (int q = 0; q < 7; q++) {
      if (q < 3 && q >= nd) continue;
      S.buf[0][q] = S.arena + cstride * cq++;
    }
  }
  S.R = (int)R;
  S.arena_ready = true;
  return EPB_OK;
}
// memory a species' arena may take: what the sorted layout's double buffer took for the capacity the host
// reserved (2 x capacity slots), or 32 GiB, whichever is larger (non-uniform decks: few dense cells, many empty)
static long long slots_row_budget(epb_handle *h, int is) {
  SpeciesDev &S = h->sp[is];
  const long long nkeys = std::max<long long>(1, h->tg.nkeys);
  const long long by_cap = (2 * S.cap) / nkeys;
  const long long by_mem = (long long)((32ull << 30) / (8ull * (h->cfg.ndims + 4) * (unsigned long long)nkeys));
  return std::max<long long>(8, std::max(by_cap, by_mem));
}

// Empty the species and make sure the arena fits what is about to be loaded: rows per column R from the
// densest cell expected (max_ppc_hint), capped by slots_row_budget.
int epb_slots_reset(epb_handle *h, int is, long long n_expected, int max_ppc_hint) {
  SpeciesDev &S = h->sp[is];
  const long long nkeys = h->tg.nkeys;
  long long want = max_ppc_hint + (long long)std::ceil(5.0 * std::sqrt((double)std::max(1, max_ppc_hint))) + 4;
  if (want < 8) want = 8;
  long long R = std::min(want, slots_row_budget(h, is));
  (void)n_expected;
  // keep an arena that is large enough and not grossly oversized
  if (S.arena_ready && S.R >= R && S.R <= 2 * R + 16) R = S.R;
  if (!S.arena_ready || S.R != (int)R) {
    int rc = slots_set_rows(h, is, R);
    if (rc) return rc;
  }
  S.cur = 0;
  S.mcur = 0;
  EPB_CUDA(h, cudaMemsetAsync(S.cnt, 0, ((size_t)nkeys + 1) * sizeof(int), h->stream));
  EPB_CUDA(h, cudaMemsetAsync(S.mcount, 0, 2 * sizeof(int), h->stream));
  if (S.icnt[0]) {
    for (int b = 0; b < 2; b++) EPB_CUDA(h, cudaMemsetAsync(S.icnt[b], 0, ((size_t)nkeys / 32 + 1) * sizeof(int), h->stream));
    S.icur = 0;
    S.inbox_dirty = false;
  }
  S.n = 0;
  S.n_sorted = 0;
  return EPB_OK;
}

// an EMPTY species gets at least R rows per column (redistribution: as many as the state it replaces)
int epb_slots_ensure_rows(epb_handle *h, int is, int R) {
  SpeciesDev &S = h->sp[is];
  if (!S.slots || (S.arena_ready && S.R >= R)) return EPB_OK;
  return slots_set_rows(h, is, std::min<long long>(R, slots_row_budget(h, is)));
}

static void fill_deliver(epb_handle *h, int is, DeliverOp &D) {
  SpeciesDev &S = h->sp[is];
  const epb_config &c = h->cfg;
  memset(&D, 0, sizeof D);
  const int m = S.mcur, o = S.mcur ^ 1;
  for (int d = 0; d < 3; d++) {
    D.sx[d] = S.mbuf[m][d]; D.sp[d] = S.mbuf[m][3 + d];
    D.ax[d] = S.buf[0][d]; D.ap[d] = S.buf[0][3 + d];
    D.ox[d] = S.mbuf[o][d]; D.op[d] = S.mbuf[o][3 + d];
    D.nloc[d] = c.n[d];
    D.gmin[d] = c.grid_min_local[d];
    D.idx[d] = d < c.ndims ? 1.0 / c.dx[d] : 0.0;
  }
  D.sw = S.mbuf[m][6]; D.aw = S.buf[0][6]; D.ow = S.mbuf[o][6];
  D.sflag = S.mflag[m]; D.oflag = S.mflag[o];
  D.scount = S.mcount + m; D.ocount = S.mcount + o;
  D.scap = D.ocap = (int)S.mcap;
  D.cnt = S.cnt;
  D.R = S.R;
  D.err = h->d_err;
  D.nd = c.ndims;
  D.ipart_mc = 1.0 / (EPB_C * S.cfg.mass);
  D.dtco2 = EPB_C * (c.dt / 2.0);
  D.tg = h->tg;
  D.rowd = S.rowd;
}

// Insert the current mover buffer's particles into their columns; the buffers swap.
int epb_slots_deliver(epb_handle *h, int is) {
  SpeciesDev &S = h->sp[is];
  DeliverOp D;
  fill_deliver(h, is, D);
  k_deliver<<<nblk((size_t)S.mcap, 148 * 8), 256, 0, h->stream>>>(D);
  h->launches++;
  EPB_CUDA(h, cudaMemsetAsync(S.mcount + S.mcur, 0, sizeof(int), h->stream));
  S.mcur ^= 1;
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

void epb_slots_views(epb_handle *h, int is, SlotView V[2]) {
  SpeciesDev &S = h->sp[is];
  memset(V, 0, 2 * sizeof(SlotView));
  for (int q = 0; q < 7; q++) { V[0].a[q] = S.buf[0][q]; V[1].a[q] = S.mbuf[S.mcur][q]; }
  V[0].r.cnt = S.cnt;
  V[0].r.K = S.rowd / 32;
  V[0].r.R = S.R;
  V[0].r.n = S.arena_ready ? (long long)h->tg.nkeys * S.R : 0;
  V[1].r.n_dev = S.mcount + S.mcur;
  V[1].r.R = (int)S.mcap;
  V[1].r.n = S.mcap;       // upper bound, for sizing the grid
  V[1].r.flag = S.mflag[S.mcur];
}

int epb_species_views(epb_handle *h, int is, SlotView V[2]) {
  SpeciesDev &S = h->sp[is];
  if (S.slots) {
    if (!S.arena_ready) return 0;
    epb_slots_settle(h, is);
    epb_slots_views(h, is, V);
    return 2;
  }
  memset(V, 0, 2 * sizeof(SlotView));
  if (S.n <= 0) return 0;
  for (int q = 0; q < 7; q++) V[0].a[q] = S.buf[S.cur][q];
  V[0].r.n = S.n;
  return 1;
}

void epb_slots_fill_push(epb_handle *h, int is, PushParams &P) {
  SpeciesDev &S = h->sp[is];
  P.cnt = S.cnt;
  P.R = S.R;
  P.rowd = S.rowd;
  for (int d = 0; d < 3; d++) { P.mx[d] = S.mbuf[S.mcur][d]; P.mp[d] = S.mbuf[S.mcur][3 + d]; }
  P.mw = S.mbuf[S.mcur][6];
  P.mflag = S.mflag[S.mcur];
  P.mcount = S.mcount + S.mcur;
  P.mcap = (int)S.mcap;
  P.err = h->d_err;
  P.ib_in = S.inbox[S.icur];
  P.ic_in = S.icnt[S.icur];
  P.ib_out = S.inbox[S.icur ^ 1];
  P.ic_out = S.icnt[S.icur ^ 1];
  P.IC = S.IC;
}

int epb_slots_after_push(epb_handle *h, int is) {
  SpeciesDev &S = h->sp[is];
  if (!S.icnt[0]) return EPB_OK;
  EPB_CUDA(h, cudaMemsetAsync(S.icnt[S.icur], 0, ((size_t)h->tg.nkeys / 32 + 1) * sizeof(int), h->stream));
  S.icur ^= 1;
  S.inbox_dirty = true;
  return EPB_OK;
}

// Group inboxes -> columns: what the next push would do on the fly, done now because something else (a
// diagnostic, a download, the balancer) is about to walk the species.  An entry that finds its column full waits in
// the current mover buffer like any other (flag 2).
int epb_slots_settle(epb_handle *h, int is) {
  SpeciesDev &S = h->sp[is];
  if (!S.slots || !S.icnt[0] || !S.inbox_dirty) return EPB_OK;
  SettleOp O;
  memset(&O, 0, sizeof O);
  for (int d = 0; d < 3; d++) {
    O.ax[d] = S.buf[0][d]; O.ap[d] = S.buf[0][3 + d];
    O.ox[d] = S.mbuf[S.mcur][d]; O.op[d] = S.mbuf[S.mcur][3 + d];
  }
  O.aw = S.buf[0][6]; O.ow = S.mbuf[S.mcur][6];
  O.oflag = S.mflag[S.mcur];
  O.ocount = S.mcount + S.mcur;
  O.ocap = (int)S.mcap;
  O.cnt = S.cnt;
  O.R = S.R;
  O.err = h->d_err;
  O.ib = S.inbox[S.icur];
  O.ic = S.icnt[S.icur];
  O.IC = S.IC;
  O.nd = h->cfg.ndims;
  O.rowd = S.rowd;
  O.ngroups = h->tg.nkeys / 32;
  k_settle<<<nblk((size_t)O.ngroups * 32, 148 * 8), 256, 0, h->stream>>>(O);
  h->launches++;
  EPB_CUDA(h, cudaMemsetAsync(S.icnt[S.icur], 0, ((size_t)O.ngroups + 1) * sizeof(int), h->stream));
  S.inbox_dirty = false;
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

int epb_slots_check(epb_handle *h) {
  if (!h->d_err) return EPB_OK;
  int e = 0;
  EPB_CUDA(h, cudaMemcpyAsync(&e, h->d_err, sizeof e, cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (e) return epb_fail(h, EPB_ERR_CAPACITY, "slot layout: particles were lost (error word %d:%s%s%s%s); raise the species capacity", e,
                         (e & 1) ? " a mover buffer overflowed" : "", (e & 2) ? " arrivals from a neighbour did not fit the mover buffer" : "",
                         (e & 4) ? " more particles left through one face than an exchange message holds" : "",
                         (e & 8) ? " the outbox list of one direction overflowed" : "");
  return EPB_OK;
}

// The count of a slot-layout species without a host round trip: d_out[0] = particles in the columns, d_out[1] = on
// their way between two columns (inbox); the caller adds the waiting entries of the mover buffer (S.mcount[S.mcur]).
int epb_slots_count_enqueue(epb_handle *h, int is, long long *d_out) {
  SpeciesDev &S = h->sp[is];
  EPB_CUDA(h, cudaMemsetAsync(d_out, 0, 2 * sizeof(long long), h->stream));
  if (!S.arena_ready) return EPB_OK;
  const int nkeys = h->tg.nkeys;
  size_t need = 0;
  int *tmp = h->cell_count;   // [nkeys + 1] scratch: clamped counts
  k_clamp_counts<<<nblk((size_t)nkeys), 256, 0, h->stream>>>(S.cnt, tmp, nkeys, S.R);
  EPB_CUDA(h, cub::DeviceReduce::Sum(nullptr, need, tmp, d_out, nkeys, h->stream));
  if (need > h->cub_tmp_bytes || !h->cub_tmp) {
    EPB_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->cub_tmp);
    EPB_CUDA(h, cudaMalloc(&h->cub_tmp, need));
    h->cub_tmp_bytes = need;
  }
  cub::DeviceReduce::Sum(h->cub_tmp, need, tmp, d_out, nkeys, h->stream);
  h->launches += 2;
  if (S.icnt[0] && S.inbox_dirty) {   // particles on their way between two columns
    const int ng = nkeys / 32;
    k_clamp_counts<<<nblk((size_t)ng), 256, 0, h->stream>>>(S.icnt[S.icur], tmp, ng, S.IC);
    cub::DeviceReduce::Sum(h->cub_tmp, need, tmp, d_out + 1, ng, h->stream);
    h->launches += 2;
  }
  return EPB_OK;
}

int epb_slots_count(epb_handle *h, int is, long long *n) {
  SpeciesDev &S = h->sp[is];
  *n = 0;
  if (!S.arena_ready) return EPB_OK;
  long long *d_out = (long long *)(h->d_scratch + 512);
  int rc = epb_slots_count_enqueue(h, is, d_out);
  if (rc) return rc;
  long long v[2] = {0, 0};
  int mc[2] = {0, 0};
  EPB_CUDA(h, cudaMemcpyAsync(v, d_out, sizeof v, cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaMemcpyAsync(mc, S.mcount, sizeof mc, cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  long long waiting = std::min<long long>(mc[S.mcur], S.mcap);   // between steps: flag-2 entries only
  *n = v[0] + waiting + v[1];
  return epb_slots_check(h);
}

// Staging through the mover buffer (upload, device loader): the caller writes m particles behind the `waiting`
// entries of the current buffer (particles that found their column full), then commits: flags, count, k_deliver.
int epb_slots_waiting(epb_handle *h, int is, int *waiting) {
  SpeciesDev &S = h->sp[is];
  EPB_CUDA(h, cudaMemcpyAsync(waiting, S.mcount + S.mcur, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (*waiting > S.mcap) *waiting = (int)S.mcap;
  return EPB_OK;
}
int epb_slots_commit(epb_handle *h, int is, int waiting, long long m) {
  SpeciesDev &S = h->sp[is];
  EPB_CUDA(h, cudaMemsetAsync(S.mflag[S.mcur] + waiting, 0, (size_t)m, h->stream));
  k_bump<<<1, 1, 0, h->stream>>>(S.mcount + S.mcur, (int)m);
  h->launches++;
  // the waiting entries (flag 2) are offered to their columns again as well: harmless
  return epb_slots_deliver(h, is);
}

// Host block (pack_particle order, partlist.F90:414-486) -> mover buffer chunk -> k_deliver.
int epb_slots_upload(epb_handle *h, int is, int64_t n, const double *packed) {
  SpeciesDev &S = h->sp[is];
  const epb_config &c = h->cfg;
  const int nd = c.ndims, nv = nd + 4;
  // densest cell (nearest cell of the stored position, io/calc_df.F90:795-796): sizes the columns
  int max_ppc = 0;
  {
    std::vector<int> hist((size_t)c.n[0] * c.n[1] * c.n[2], 0);
    for (int64_t i = 0; i < n; i++) {
      size_t o = 0, str = 1;
      for (int d = 0; d < nd; d++) {
        int cd = (int)std::floor((packed[i * nv + d] - c.grid_min_local[d]) / c.dx[d] + 0.5);
        cd = cd < 0 ? 0 : (cd > c.n[d] - 1 ? c.n[d] - 1 : cd);
        o += str * (size_t)cd;
        str *= (size_t)c.n[d];
      }
      const int v = ++hist[o];
      if (v > max_ppc) max_ppc = v;
    }
  }
  int rc = epb_slots_reset(h, is, n, max_ppc);
  if (rc) return rc;
  // the packed block goes to the device as it is (chunks of <= 2 Mi particles through a staging buffer) and
  // is taken apart into the SoA mover buffer there
  const int64_t CH = std::min<int64_t>(S.mcap, 2 << 20);
  if (!h->aos_stage) {
    EPB_CUDA(h, cudaMalloc(&h->aos_stage, (size_t)(2 << 20) * 7 * sizeof(double)));
  }
  int64_t i0 = 0;
  while (i0 < n) {
    int waiting = 0;
    rc = epb_slots_waiting(h, is, &waiting);
    if (rc) return rc;
    const int64_t mm = std::min<int64_t>(std::min<int64_t>(n - i0, CH), S.mcap - waiting);
    if (mm <= 0) return epb_fail(h, EPB_ERR_CAPACITY, "species %d: %d particles fit neither their columns (R = %d) nor the mover buffer", is, waiting, S.R);
    EPB_CUDA(h, cudaMemcpyAsync(h->aos_stage, packed + i0 * nv, (size_t)mm * nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    AosOp A;
    for (int q = 0; q < 7; q++) A.soa[q] = S.mbuf[S.mcur][q] ? S.mbuf[S.mcur][q] + waiting : nullptr;
    A.aos = h->aos_stage; A.n = mm; A.nd = nd; A.nv = nv; A.to_soa = 1;
    k_aos<<<nblk((size_t)mm), 256, 0, h->stream>>>(A);
    h->launches++;
    rc = epb_slots_commit(h, is, waiting, mm);
    if (rc) return rc;
    i0 += mm;
  }
  return epb_slots_check(h);
}

// ---- walking a species chunk by chunk (download, redistribution): any layout ---------------------------------
// Every call to epb_species_iter_next hands out the next <= CH particles as contiguous SoA device arrays
// (slot columns: gathered into the idle mover buffer by k_compact; the waiting entries of the current mover
// buffer and the classic contiguous layout are handed out in place).  The species must not change meanwhile.
int epb_species_iter_begin(epb_handle *h, int is, SpeciesIter &I) {
  SpeciesDev &S = h->sp[is];
  I = SpeciesIter();
  if (!S.slots) return EPB_OK;
  if (!S.arena_ready) return EPB_OK;
  {
    int rcs = epb_slots_settle(h, is);
    if (rcs) return rcs;
  }
  const int nkeys = h->tg.nkeys;
  // scan of the clamped counts -> start[nkeys + 1]
  int *tmpc = h->cell_count, *start = h->cell_start;
  k_clamp_counts<<<nblk((size_t)nkeys), 256, 0, h->stream>>>(S.cnt, tmpc, nkeys, S.R);
  EPB_CUDA(h, cudaMemsetAsync(tmpc + nkeys, 0, sizeof(int), h->stream));
  size_t need = 0;
  EPB_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, need, tmpc, start, nkeys + 1, h->stream));
  if (need > h->cub_tmp_bytes || !h->cub_tmp) {
    EPB_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->cub_tmp);
    EPB_CUDA(h, cudaMalloc(&h->cub_tmp, need));
    h->cub_tmp_bytes = need;
  }
  cub::DeviceScan::ExclusiveSum(h->cub_tmp, need, tmpc, start, nkeys + 1, h->stream);
  h->launches += 2;
  I.hstart.resize((size_t)nkeys + 1);
  EPB_CUDA(h, cudaMemcpyAsync(I.hstart.data(), start, ((size_t)nkeys + 1) * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaMemcpyAsync(&I.mc, S.mcount + S.mcur, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (I.mc > S.mcap) I.mc = (int)S.mcap;
  return EPB_OK;
}

int epb_species_iter_next(epb_handle *h, int is, SpeciesIter &I, long long CH, double *st[7], long long *m) {
  SpeciesDev &S = h->sp[is];
  *m = 0;
  for (int q = 0; q < 7; q++) st[q] = nullptr;
  if (!S.slots) {
    if (I.lin >= S.n) return EPB_OK;
    *m = std::min<long long>(CH, S.n - I.lin);
    for (int q = 0; q < 7; q++) st[q] = S.buf[S.cur][q] ? S.buf[S.cur][q] + I.lin : nullptr;
    I.lin += *m;
    return EPB_OK;
  }
  if (!S.arena_ready) return EPB_OK;
  const int nkeys = h->tg.nkeys;
  if (CH > S.mcap) CH = S.mcap;
  const int stg = S.mcur ^ 1;
  while (I.k0 < nkeys) {
    // largest group-aligned key range whose particles fit one chunk
    int k1 = I.k0;
    while (k1 < nkeys) {
      const int kn = std::min(nkeys, k1 + 32);
      if ((long long)I.hstart[kn] - I.hstart[I.k0] > CH) break;
      k1 = kn;
    }
    if (k1 == I.k0) return epb_fail(h, EPB_ERR_CAPACITY, "one group of slot columns exceeds the staging buffer");
    const long long mm = (long long)I.hstart[k1] - I.hstart[I.k0];
    const int k0 = I.k0;
    I.k0 = k1;
    if (mm > 0) {
      CompactOp C;
      for (int q = 0; q < 7; q++) { C.a[q] = S.buf[0][q]; C.dst[q] = S.mbuf[stg][q]; }
      C.cnt = S.cnt; C.start = h->cell_start; C.R = S.R; C.k0 = k0; C.k1 = k1; C.base = I.hstart[k0];
      C.rowd = S.rowd;
      k_compact<<<nblk((size_t)(k1 - k0), 148 * 8), 256, 0, h->stream>>>(C);
      h->launches++;
      for (int q = 0; q < 7; q++) st[q] = S.mbuf[stg][q];
      *m = mm;
      return EPB_OK;
    }
  }
  // the entries waiting in the current mover buffer (no room in their column): all are live between steps
  if (I.woff < I.mc) {
    *m = std::min<long long>(I.mc - I.woff, CH);
    for (int q = 0; q < 7; q++) st[q] = S.mbuf[S.mcur][q] ? S.mbuf[S.mcur][q] + I.woff : nullptr;
    I.woff += *m;
  }
  return EPB_OK;
}

int epb_slots_download(epb_handle *h, int is, int64_t n, double *packed) {
  SpeciesDev &S = h->sp[is];
  const epb_config &c = h->cfg;
  const int nd = c.ndims, nv = nd + 4;
  if (!S.arena_ready || n <= 0) return EPB_OK;
  const long long CH = std::min<long long>(S.mcap, 2 << 20);
  if (!h->aos_stage) {
    EPB_CUDA(h, cudaMalloc(&h->aos_stage, (size_t)(2 << 20) * 7 * sizeof(double)));
  }
  SpeciesIter I;
  int rc = epb_species_iter_begin(h, is, I);
  if (rc) return rc;
  int64_t done = 0;
  while (done < n) {
    double *st[7];
    long long m = 0;
    rc = epb_species_iter_next(h, is, I, CH, st, &m);
    if (rc) return rc;
    if (m == 0) break;
    const long long take = std::min<long long>(m, n - done);
    AosOp A;
    for (int q = 0; q < 7; q++) A.soa[q] = st[q];
    A.aos = h->aos_stage; A.n = take; A.nd = nd; A.nv = nv; A.to_soa = 0;
    k_aos<<<nblk((size_t)take), 256, 0, h->stream>>>(A);
    h->launches++;
    EPB_CUDA(h, cudaMemcpyAsync(packed + done * nv, h->aos_stage, (size_t)take * nv * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EPB_CUDA(h, cudaStreamSynchronize(h->stream));
    done += take;
  }
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

// Particles that are already on the device in the wire layout (redistribution): appended to the species.
int epb_species_insert_aos(epb_handle *h, int is, const double *aos_dev, long long n) {
  SpeciesDev &S = h->sp[is];
  const int nd = h->cfg.ndims, nv = nd + 4;
  if (n <= 0) return EPB_OK;
  if (!S.slots) {
    if (S.n + n > S.cap) return epb_fail(h, EPB_ERR_CAPACITY, "species %d: %lld particles > capacity %lld", is, S.n + n, S.cap);
    AosOp A;
    for (int q = 0; q < 7; q++) A.soa[q] = S.buf[S.cur][q] ? S.buf[S.cur][q] + S.n : nullptr;
    A.aos = const_cast<double *>(aos_dev); A.n = n; A.nd = nd; A.nv = nv; A.to_soa = 1;
    k_aos<<<nblk((size_t)n), 256, 0, h->stream>>>(A);
    h->launches++;
    S.n += n;
    S.info_valid = false;
    h->pushes_since_sort = 1 << 30;   // sort before the next push
    return EPB_OK;
  }
  long long i0 = 0;
  while (i0 < n) {
    int waiting = 0;
    int rc = epb_slots_waiting(h, is, &waiting);
    if (rc) return rc;
    const long long mm = std::min<long long>(n - i0, S.mcap - waiting);
    if (mm <= 0) return epb_fail(h, EPB_ERR_CAPACITY, "species %d: %d particles fit neither their columns (R = %d) nor the mover buffer", is, waiting, S.R);
    AosOp A;
    for (int q = 0; q < 7; q++) A.soa[q] = S.mbuf[S.mcur][q] ? S.mbuf[S.mcur][q] + waiting : nullptr;
    A.aos = const_cast<double *>(aos_dev) + i0 * nv; A.n = mm; A.nd = nd; A.nv = nv; A.to_soa = 1;
    k_aos<<<nblk((size_t)mm), 256, 0, h->stream>>>(A);
    h->launches++;
    rc = epb_slots_commit(h, is, waiting, mm);
    if (rc) return rc;
    i0 += mm;
  }
  return EPB_OK;
}
