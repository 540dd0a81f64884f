// exchange.cu — ghost-cell halos, current-sum halos and particle migration.
//
// Replaces the MPI_SENDRECV sequences of boundary.F90:
//   do_field_mpi_with_lengths (:222-315)  -> epb_halo_exchange(add=false)
//   particle_periodic_bcs     (:634-751)  -> epb_halo_exchange(add=true)
//   particle_bcs exchange     (:1436-1446) + partlist_sendrecv (partlist.F90:830-884)
//                                         -> epb_particle_exchange
// One rank per GPU.  A dimension with a single rank wraps onto itself with device
// copies; otherwise strips are packed by a kernel, moved with grouped
// ncclSend/ncclRecv over NVLink, and unpacked (copy or add) by a kernel, all on the
// handle's stream.  The three components of E, B or J travel in one message.
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cstring>

#include "epb_internal.h"

int epb_halo_local(epb_handle *h, int f0, int nf, bool add, int d, int pass);

namespace {

constexpr int NG = EPB_NG;

struct PackOp {
  double *f[3];
  int nf, nd, sz[3];
  int lo[3], ext[3];
  double *buf;
  int mode;  // 0 pack (array -> buf), 1 unpack copy, 2 unpack add
};
__device__ __forceinline__ size_t pofs(const int *sz, int nd, int i, int j, int k) {
  size_t o = (size_t)(i + NG - 1);
  if (nd >= 2) o += (size_t)sz[0] * (size_t)(j + NG - 1);
  if (nd >= 3) o += (size_t)sz[0] * (size_t)sz[1] * (size_t)(k + NG - 1);
  return o;
}
__global__ void __launch_bounds__(256) k_pack(const __grid_constant__ PackOp B) {
  const size_t total = (size_t)B.ext[0] * B.ext[1] * B.ext[2];
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int a = (int)(t % B.ext[0]);
    const int b = (int)((t / B.ext[0]) % B.ext[1]);
    const int c_ = (int)(t / ((size_t)B.ext[0] * B.ext[1]));
    const size_t o = pofs(B.sz, B.nd, B.lo[0] + a, B.lo[1] + b, B.lo[2] + c_);
    for (int q = 0; q < B.nf; q++) {
      double *p = B.buf + (size_t)q * total + t;
      if (B.mode == 0) *p = B.f[q][o];
      else if (B.mode == 1) B.f[q][o] = *p;
      else B.f[q][o] = B.f[q][o] + *p;
    }
  }
}

inline int nblk(size_t total, int cap = 148 * 16) {
  size_t b = (total + 255) / 256;
  if (b < 1) b = 1;
  if (b > (size_t)cap) b = cap;
  return (int)b;
}
inline int nbr1(const epb_config &c, int d, int s) {
  int o[3] = {0, 0, 0};
  o[d] = s;
  return c.neighbour[(o[2] + 1) * 9 + (o[1] + 1) * 3 + (o[0] + 1)];
}
int bc_allspecies(const epb_handle *h, int i) {
  if (h->sp.empty()) return h->cfg.bc_field[i] == EPB_BC_PERIODIC ? EPB_BC_PERIODIC : EPB_BC_OPEN;
  int b = h->sp[h->bc_species >= 0 ? h->bc_species : 0].cfg.bc_particle[i];
  if (b != EPB_BC_REFLECT && b != EPB_BC_PERIODIC) b = EPB_BC_OPEN;
  return b;
}
int ensure_buf(epb_handle *h, double **buf, size_t *have, size_t need) {
  if (need <= *have) return EPB_OK;
  if (*buf) { cudaStreamSynchronize(h->stream); cudaFree(*buf); *buf = nullptr; }
  size_t n = need + need / 4 + 1024;
  EPB_CUDA(h, cudaMalloc(buf, n * sizeof(double)));
  *have = n;
  return EPB_OK;
}
#define EPB_NCCL(h, call)                                                                         \
  do {                                                                                            \
    ncclResult_t r_ = (call);                                                                     \
    if (r_ != ncclSuccess)                                                                        \
      return epb_fail((h), EPB_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_)); \
  } while (0)

void run_pack(epb_handle *h, int f0, int nf, const int lo[3], const int ext[3], double *buf, int mode) {
  PackOp B;
  B.nf = nf;
  B.nd = h->cfg.ndims;
  for (int q = 0; q < nf; q++) B.f[q] = h->f(f0 + q);
  for (int q = 0; q < 3; q++) { B.sz[q] = h->sz[q]; B.lo[q] = lo[q]; B.ext[q] = ext[q]; }
  B.buf = buf;
  B.mode = mode;
  size_t total = (size_t)ext[0] * ext[1] * ext[2];
  k_pack<<<nblk(total), 256, 0, h->stream>>>(B);
  h->launches++;
}

// ---- particle migration kernels -------------------------------------------------
struct PPackOp {
  const double *src[7];
  int nv, nd;
  const int *idx;   // outbox list of one direction
  int count;
  double *buf;      // AoS, pack_particle order
};
__global__ void __launch_bounds__(256) k_ppack(const __grid_constant__ PPackOp O) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < O.count; t += gridDim.x * blockDim.x) {
    const int i = O.idx[t];
    double *o = O.buf + (size_t)t * O.nv;
    int q = 0;
    for (int d = 0; d < O.nd; d++) o[q++] = O.src[d][i];
    for (int d = 3; d < 7; d++) o[q++] = O.src[d][i];
  }
}
struct PUnpackOp {
  double *dst[7];
  int *key;  // emitted sort records (layout 1) or null
  int nv, nd;
  long long first;
  int count;
  const double *buf;
  // layout 2: arrivals are appended to the mover buffer, whose count lives on the device
  const int *first_dev;
  unsigned char *flag;
  int cap;
  int *err;
};
__global__ void k_bump_count(int *count_dev, int n) { *count_dev += n; }
__global__ void __launch_bounds__(256) k_punpack(const __grid_constant__ PUnpackOp O) {
  const long long first = O.first_dev ? (long long)*O.first_dev : O.first;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < O.count; t += gridDim.x * blockDim.x) {
    const double *o = O.buf + (size_t)t * O.nv;
    const long long i = first + t;
    if (O.first_dev) {
      if (i >= O.cap) { atomicOr(O.err, 2); continue; }
      O.flag[i] = 0;
    }
    int q = 0;
    for (int d = 0; d < O.nd; d++) O.dst[d][i] = o[q++];
    for (int d = 3; d < 7; d++) O.dst[d][i] = o[q++];
    if (O.key) O.key[i] = -1;  // no emitted sort record for an arrival
  }
}
// ---- the same exchange without a host round trip (slot layout) ------------------------------------------------
// Every direction's message has a FIXED size agreed at epb_set_comm: a header record whose first double is the
// number of particles that follow, then room for xcap records.  The counts never visit the host, so the host can
// enqueue the whole step (and the next one) without waiting for the push kernel.  NVLink makes the padding cheap:
// at C2 the four face messages are 13 MB each, microseconds per step.
struct PPackDevOp {
  const double *src[7];
  int nv, nd;
  const int *idx;        // outbox list of this direction
  const int *count_dev;  // its device counter
  int xcap, out_cap;
  double *buf;           // header record + xcap records
  int *err;
};
__global__ void __launch_bounds__(256) k_ppack_dev(const __grid_constant__ PPackDevOp O) {
  int n = *O.count_dev;
  if (n > O.out_cap) {                    // the outbox list itself overflowed: those leavers were not recorded
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(O.err, 8);
    n = O.out_cap;
  }
  if (n > O.xcap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(O.err, 4);   // more leavers than the agreed message holds
    n = O.xcap;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) O.buf[0] = (double)n;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int i = O.idx[t];
    double *o = O.buf + (size_t)(t + 1) * O.nv;
    int q = 0;
    for (int d = 0; d < O.nd; d++) o[q++] = O.src[d][i];
    for (int d = 3; d < 7; d++) o[q++] = O.src[d][i];
  }
}
struct PUnpackDevOp {
  double *dst[7];
  int nv, nd;
  const double *buf;     // header record + records
  int xcap;
  int *mcount;           // device count of the mover buffer the arrivals are appended to
  unsigned char *flag;
  int cap;
  int *err;
};
__global__ void __launch_bounds__(256) k_punpack_dev(const __grid_constant__ PUnpackDevOp O) {
  int n = (int)O.buf[0];
  if (n < 0) n = 0;
  if (n > O.xcap) n = O.xcap;
  const long long first = (long long)*O.mcount;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const double *o = O.buf + (size_t)(t + 1) * O.nv;
    const long long i = first + t;
    if (i >= O.cap) { atomicOr(O.err, 2); continue; }
    O.flag[i] = 0;
    int q = 0;
    for (int d = 0; d < O.nd; d++) O.dst[d][i] = o[q++];
    for (int d = 3; d < 7; d++) O.dst[d][i] = o[q++];
  }
}
__global__ void k_bump_count_dev(int *count_dev, const double *header) {
  int n = (int)header[0];
  if (n > 0) *count_dev += n;
}
// survivors in the tail [n_new, n_old) that must move into holes below n_new
__global__ void __launch_bounds__(256) k_tail_movers(const unsigned char *gone, long long n_new, long long n_old,
                                                     int *movers, int *counter) {
  for (long long j = n_new + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_old; j += (long long)gridDim.x * blockDim.x)
    if (!gone[j]) movers[atomicAdd(counter, 1)] = (int)j;
}
struct FillOp {
  double *a[7];
  int *key, *rank;  // emitted sort records travel with the particle (layout 1), else null
  unsigned char *gone;
  const int *idx;   // outbox list of one direction
  int count;
  long long n_new;
  const int *movers;
  int *cursor;
};
__global__ void __launch_bounds__(256) k_fill_holes(const __grid_constant__ FillOp F) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < F.count; t += gridDim.x * blockDim.x) {
    const int hole = F.idx[t];
    F.gone[hole] = 0;
    if (hole < F.n_new) {
      const int src = F.movers[atomicAdd(F.cursor, 1)];
#pragma unroll
      for (int q = 0; q < 7; q++)
        if (F.a[q]) F.a[q][hole] = F.a[q][src];
      if (F.key) { F.key[hole] = F.key[src]; F.rank[hole] = F.rank[src]; }
    }
  }
}

}  // namespace

int epb_comm_init(epb_handle *h, const void *id128) {
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclComm_t comm;
  EPB_NCCL(h, ncclCommInitRank(&comm, h->cfg.nranks, id, h->cfg.rank));
  h->nccl = comm;
  return EPB_OK;
}
void epb_comm_destroy(epb_handle *h) {
  if (h->nccl) { ncclCommDestroy((ncclComm_t)h->nccl); h->nccl = nullptr; }
}

extern "C" int epb_nccl_unique_id(void *id128) {
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return EPB_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(id128, &id, sizeof id);
  return EPB_OK;
}
// Message capacities of the host-round-trip-free particle exchange (particle_exchange_async): per axis, the particles
// of one cell layer of the fullest species buffer (what can cross a face in one step at most in a uniform plasma,
// v dt < dx; a thermal plasma sends a few per cent of that), at least 64 Ki; edges and corners a 64th of the largest.
// Both ends of a message must use the same number: maximum over the ranks.  Collective; called when the communicator
// is attached and again when the balancer has re-cut the slabs (new extents, new capacities).
int epb_agree_exchange_caps(epb_handle *h) {
  const epb_config &c = h->cfg;
  for (int q = 0; q < 27; q++) h->xcap[q] = 0;
  if (c.nranks <= 1 || !h->nccl) return EPB_OK;
  long long prop[4] = {0, 0, 0, 0};
  long long maxcap = 0;
  for (auto &S : h->sp) maxcap = std::max<long long>(maxcap, S.cap);
  for (int d = 0; d < c.ndims; d++) {
    prop[d] = std::max<long long>(65536, maxcap / std::max(1, c.n[d]));
    prop[3] = std::max(prop[3], prop[d] / 64);
  }
  prop[3] = std::max<long long>(prop[3], 4096);
  long long *d = (long long *)(h->d_scratch + 896);
  EPB_CUDA(h, cudaMemcpyAsync(d, prop, sizeof prop, cudaMemcpyHostToDevice, h->stream));
  EPB_NCCL(h, ncclAllReduce(d, d, 4, ncclInt64, ncclMax, (ncclComm_t)h->nccl, h->stream));
  EPB_CUDA(h, cudaMemcpyAsync(prop, d, sizeof prop, cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int q = 0; q < 27; q++) {
    const int o[3] = {q % 3 - 1, (q / 3) % 3 - 1, q / 9 - 1};
    const int nz = (o[0] != 0) + (o[1] != 0) + (o[2] != 0);
    if (nz == 1) h->xcap[q] = (int)prop[o[0] != 0 ? 0 : (o[1] != 0 ? 1 : 2)];
    else if (nz > 1) h->xcap[q] = (int)prop[3];
  }
  return EPB_OK;
}

extern "C" int epb_set_comm(epb_handle *h, const void *id128) {
  if (!h || !id128) return EPB_ERR_ARG;
  if (h->cfg.nranks <= 1) return EPB_OK;
  int rc = epb_comm_init(h, id128);
  if (rc) return rc;
  return epb_agree_exchange_caps(h);
}

// get_load_x / get_load_y (balance.F90:1766-1844; epoch3d :2247-2362; epoch1d :980-1006): histogram of the
// particles of all species over the GLOBAL cells of one axis, cell = FLOOR((pos - x_grid_min) / dx + 1.5) + ng
namespace {
__global__ void __launch_bounds__(256) k_load_profile(const double *x, const __grid_constant__ PRange R, double grid_min, double dx, int len,
                                                      unsigned long long *load) {
  const long long n = prange_n(R);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (!prange_valid(R, i)) continue;
    const int cell = __double2int_rd((x[prange_at(R, i)] - grid_min) / dx + 1.5) + NG;   // index into load(1:len)
    if (cell >= 1 && cell <= len) atomicAdd(load + (cell - 1), 1ULL);
  }
}
}  // namespace

extern "C" int epb_load_profile(epb_handle *h, int axis, int64_t *load) {
  if (!h || !load || axis < 0 || axis >= h->cfg.ndims) return EPB_ERR_ARG;
  const epb_config &c = h->cfg;
  const int len = c.n_global[axis] + 2 * NG;
  // handle-owned scratch (two histograms: local and summed), grown on demand and released by epb_destroy
  if (h->prof_cap < 2 * (size_t)len) {
    EPB_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->prof_scratch);
    h->prof_scratch = nullptr;
    h->prof_cap = 0;
    EPB_CUDA(h, cudaMalloc(&h->prof_scratch, 2 * (size_t)len * sizeof(unsigned long long)));
    h->prof_cap = 2 * (size_t)len;
  }
  unsigned long long *d = h->prof_scratch;
  EPB_CUDA(h, cudaMemsetAsync(d, 0, 2 * (size_t)len * sizeof(unsigned long long), h->stream));
  const double grid_min = c.gmin[axis] + c.dx[axis] / 2.0;   // x_grid_min (setup.F90:169,180; no CPML)
  for (size_t is = 0; is < h->sp.size(); is++) {
    SlotView V[2];
    const int nv = epb_species_views(h, (int)is, V);
    for (int v = 0; v < nv; v++) {
      k_load_profile<<<nblk((size_t)V[v].r.n), 256, 0, h->stream>>>(V[v].a[axis], V[v].r, grid_min, c.dx[axis], len, d);
      h->launches++;
    }
  }
  EPB_CUDA(h, cudaGetLastError());
  unsigned long long *res = d;
  if (c.nranks > 1 && h->nccl) {   // MPI_ALLREDUCE(MPI_IN_PLACE, load, st, MPI_INTEGER8, MPI_SUM)
    ncclResult_t r = ncclAllReduce(d, d + len, len, ncclInt64, ncclSum, (ncclComm_t)h->nccl, h->stream);
    if (r != ncclSuccess) return epb_fail(h, EPB_ERR_NCCL, "ncclAllReduce: %s", ncclGetErrorString(r));
    res = d + len;
  }
  cudaError_t e = cudaMemcpyAsync(load, res, (size_t)len * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return epb_fail(h, EPB_ERR_CUDA, "epb_load_profile: %s", cudaGetErrorString(e));
  // load = push_per_field * load; load(ng+1:st-ng) += cells of one slab across the other axes
  int64_t other = 1;
  for (int q = 0; q < c.ndims; q++)
    if (q != axis) other *= c.n_global[q];
  for (int i = 0; i < len; i++) {
    load[i] *= EPB_PUSH_PER_FIELD;   // shared_data.F90:821
    if (i >= NG && i < len - NG) load[i] += other;
  }
  return EPB_OK;
}

int epb_allreduce_sum_f64(epb_handle *h, const double *src, double *dst, int n) {
  EPB_NCCL(h, ncclAllReduce(src, dst, n, ncclDouble, ncclSum, (ncclComm_t)h->nccl, h->stream));
  return EPB_OK;
}

extern "C" int epb_global_count(epb_handle *h, int is, int64_t *n) {
  if (!h || is < 0 || is >= (int)h->sp.size() || !n) return EPB_ERR_ARG;
  long long local = h->sp[is].n;
  if (h->sp[is].slots) {
    int rcs = epb_slots_count(h, is, &local);
    if (rcs) return rcs;
  }
  if (h->cfg.nranks <= 1 || !h->nccl) { *n = local; return EPB_OK; }
  long long *d = (long long *)h->d_scratch;
  EPB_CUDA(h, cudaMemcpyAsync(d, &local, sizeof local, cudaMemcpyHostToDevice, h->stream));
  EPB_NCCL(h, ncclAllReduce(d, d + 1, 1, ncclInt64, ncclSum, (ncclComm_t)h->nccl, h->stream));
  long long out = 0;
  EPB_CUDA(h, cudaMemcpyAsync(&out, d + 1, sizeof out, cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  *n = out;
  return EPB_OK;
}

int epb_halo_exchange(epb_handle *h, int f0, int nf, bool add) {
  const epb_config &c = h->cfg;
  for (int d = 0; d < c.ndims; d++) {
    const int nm = nbr1(c, d, -1), np = nbr1(c, d, +1);
    bool lo_ok, hi_ok;  // do we take part in an exchange across our low / high face
    if (!add) {
      lo_ok = (!c.is_boundary[2 * d] || c.bc_field[2 * d] == EPB_BC_PERIODIC) && nm >= 0;
      hi_ok = (!c.is_boundary[2 * d + 1] || c.bc_field[2 * d + 1] == EPB_BC_PERIODIC) && np >= 0;
    } else {
      lo_ok = !(c.is_boundary[2 * d] && bc_allspecies(h, 2 * d) != EPB_BC_PERIODIC) && nm >= 0;
      hi_ok = !(c.is_boundary[2 * d + 1] && bc_allspecies(h, 2 * d + 1) != EPB_BC_PERIODIC) && np >= 0;
    }
    const bool self = (nm == c.rank || nm < 0) && (np == c.rank || np < 0);
    if (self) {
      // copy: pass 0 fills the high ghosts (received from proc_max), pass 1 the low ghosts
      // add : pass 0 adds into the low interior (received from neighbour -1), pass 1 the high interior
      if (!add) {
        if (hi_ok) epb_halo_local(h, f0, nf, false, d, 0);
        if (lo_ok) epb_halo_local(h, f0, nf, false, d, 1);
      } else {
        if (lo_ok) epb_halo_local(h, f0, nf, true, d, 0);
        if (hi_ok) epb_halo_local(h, f0, nf, true, d, 1);
      }
      continue;
    }
    if (!h->nccl) return epb_fail(h, EPB_ERR_NCCL, "rank has remote neighbours but epb_set_comm was not called");
    ncclComm_t comm = (ncclComm_t)h->nccl;
    int lo[3], ext[3];
    for (int q = 0; q < 3; q++) { lo[q] = q < c.ndims ? 1 - NG : 1; ext[q] = h->sz[q]; }
    ext[d] = NG;
    const size_t strip = (size_t)ext[0] * ext[1] * ext[2] * nf;
    int rc = ensure_buf(h, &h->sendbuf, &h->sendbuf_elems, 2 * strip);
    if (rc) return rc;
    rc = ensure_buf(h, &h->recvbuf, &h->recvbuf_elems, 2 * strip);
    if (rc) return rc;
    double *sA = h->sendbuf, *sB = h->sendbuf + strip;   // A goes to nm, B goes to np
    double *rA = h->recvbuf, *rB = h->recvbuf + strip;   // rA comes from np, rB comes from nm
    const int n = c.n[d];
    int l2[3] = {lo[0], lo[1], lo[2]};
    if (lo_ok) { l2[d] = add ? 1 - NG : 1; run_pack(h, f0, nf, l2, ext, sA, 0); }
    if (hi_ok) { l2[d] = add ? n + 1 : n + 1 - NG; run_pack(h, f0, nf, l2, ext, sB, 0); }
    EPB_NCCL(h, ncclGroupStart());
    if (lo_ok) EPB_NCCL(h, ncclSend(sA, strip, ncclDouble, nm, comm, h->stream));
    if (hi_ok) EPB_NCCL(h, ncclSend(sB, strip, ncclDouble, np, comm, h->stream));
    if (hi_ok) EPB_NCCL(h, ncclRecv(rA, strip, ncclDouble, np, comm, h->stream));
    if (lo_ok) EPB_NCCL(h, ncclRecv(rB, strip, ncclDouble, nm, comm, h->stream));
    EPB_NCCL(h, ncclGroupEnd());
    h->launches++;
    if (!add) {
      if (hi_ok) { l2[d] = n + 1; run_pack(h, f0, nf, l2, ext, rA, 1); }
      if (lo_ok) { l2[d] = 1 - NG; run_pack(h, f0, nf, l2, ext, rB, 1); }
    } else {
      // np's low ghosts add into our high interior; nm's high ghosts add into our low interior
      if (lo_ok) { l2[d] = 1; run_pack(h, f0, nf, l2, ext, rB, 2); }
      if (hi_ok) { l2[d] = n + 1 - NG; run_pack(h, f0, nf, l2, ext, rA, 2); }
    }
  }
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

// particle_bcs tail: remove leavers, exchange with up to 26 neighbours, append arrivals
// Particle migration of a slot-layout species with no host round trip (see k_ppack_dev): the leavers sit in the
// mover buffer, flagged, the outbox lists their indices per direction; arrivals are appended to the same buffer and
// placed by k_deliver.  Overflows of any fixed capacity set bits of the device error word, which the next
// host-visible call (epb_global_count, epb_step_scalars_async, a download) reports as EPB_ERR_CAPACITY.
static int particle_exchange_async(epb_handle *h, int is) {
  const epb_config &c = h->cfg;
  SpeciesDev &S = h->sp[is];
  const int nd = c.ndims, nv = nd + 4;
  size_t off[28];
  off[0] = 0;
  for (int q = 0; q < 27; q++) off[q + 1] = off[q] + ((size_t)h->xcap[q] + 1) * nv;
  int rc = ensure_buf(h, &h->sendbuf, &h->sendbuf_elems, off[27]);
  if (rc) return rc;
  rc = ensure_buf(h, &h->recvbuf, &h->recvbuf_elems, off[27]);
  if (rc) return rc;
  double *const *arr = S.mbuf[S.mcur];
  for (int q = 0; q < 27; q++) {
    const int to = c.neighbour[q];
    if (q == 13 || to < 0 || to == c.rank || h->xcap[q] <= 0) continue;
    PPackDevOp O;
    for (int k = 0; k < 7; k++) O.src[k] = arr[k];
    O.nv = nv; O.nd = nd;
    O.idx = h->out_idx + (size_t)q * h->out_cap;
    O.count_dev = h->out_count + q;
    O.xcap = h->xcap[q]; O.out_cap = h->out_cap;
    O.buf = h->sendbuf + off[q];
    O.err = h->d_err;
    k_ppack_dev<<<nblk((size_t)std::min(h->xcap[q], 1 << 18)), 256, 0, h->stream>>>(O);
    h->launches++;
  }
  ncclComm_t comm = (ncclComm_t)h->nccl;
  EPB_NCCL(h, ncclGroupStart());
  for (int q = 0; q < 27; q++) {
    if (q == 13 || h->xcap[q] <= 0) continue;
    const int to = c.neighbour[q], from = c.neighbour[26 - q];
    const size_t len = ((size_t)h->xcap[q] + 1) * nv;
    if (to >= 0 && to != c.rank) EPB_NCCL(h, ncclSend(h->sendbuf + off[q], len, ncclDouble, to, comm, h->stream));
    if (from >= 0 && from != c.rank) EPB_NCCL(h, ncclRecv(h->recvbuf + off[q], len, ncclDouble, from, comm, h->stream));
  }
  EPB_NCCL(h, ncclGroupEnd());
  h->launches++;
  // append arrivals in the reference's direction order (boundary.F90:1436-1446)
  for (int q = 0; q < 27; q++) {
    const int from = c.neighbour[26 - q];
    if (q == 13 || from < 0 || from == c.rank || h->xcap[q] <= 0) continue;
    PUnpackDevOp O;
    for (int k = 0; k < 7; k++) O.dst[k] = arr[k];
    O.nv = nv; O.nd = nd;
    O.buf = h->recvbuf + off[q];
    O.xcap = h->xcap[q];
    O.mcount = S.mcount + S.mcur;
    O.flag = S.mflag[S.mcur];
    O.cap = (int)S.mcap;
    O.err = h->d_err;
    k_punpack_dev<<<nblk((size_t)std::min(h->xcap[q], 1 << 18)), 256, 0, h->stream>>>(O);
    k_bump_count_dev<<<1, 1, 0, h->stream>>>(S.mcount + S.mcur, h->recvbuf + off[q]);
    h->launches += 2;
  }
  EPB_CUDA(h, cudaMemsetAsync(h->out_count, 0, 27 * sizeof(int), h->stream));
  EPB_CUDA(h, cudaGetLastError());
  return EPB_OK;
}

int epb_particle_exchange(epb_handle *h, int is) {
  const epb_config &c = h->cfg;
  SpeciesDev &S = h->sp[is];
  if (h->out_cap == 0 || S.cfg.immobile) return EPB_OK;
  const int nd = c.ndims, nv = nd + 4;
  static const int force_sync = epb_env("EPB_EXCHANGE_SYNC") ? atoi(epb_env("EPB_EXCHANGE_SYNC")) : 0;
  if (S.slots && c.nranks > 1 && h->nccl && !force_sync) return particle_exchange_async(h, is);
  int *cnt = h->h_counts;
  EPB_CUDA(h, cudaMemcpyAsync(cnt, h->out_count, 27 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  long long gone_total = 0;
  // An overflowing outbox is reported AFTER the exchange has completed with what fits: the neighbours have
  // already posted their sends / receives against this rank, and an early return here would leave them
  // blocked for ever; the error then surfaces on this rank while the others stay consistent.
  int overflow_q = -1, overflow_n = 0;
  for (int q = 0; q < 27; q++) {
    if (cnt[q] > h->out_cap) { overflow_q = q; overflow_n = cnt[q]; cnt[q] = h->out_cap; }
    gone_total += cnt[q];
  }
  const bool remote = c.nranks > 1 && h->nccl;
  int sendc[27], recvc[27];
  size_t send_off[27], recv_off[27];
  size_t send_tot = 0, recv_tot = 0;
  for (int q = 0; q < 27; q++) { sendc[q] = (q == 13) ? 0 : cnt[q]; recvc[q] = 0; }
  if (remote) {
    // count exchange (partlist.F90:850): for direction q we send to neighbour(q) and receive
    // from neighbour(-q) what that rank sends in direction q
    int *d_send = h->d_scratch + 64, *d_recv = h->d_scratch + 128;
    EPB_CUDA(h, cudaMemcpyAsync(d_send, h->out_count, 27 * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    EPB_CUDA(h, cudaMemsetAsync(d_recv, 0, 27 * sizeof(int), h->stream));
    ncclComm_t comm = (ncclComm_t)h->nccl;
    EPB_NCCL(h, ncclGroupStart());
    for (int q = 0; q < 27; q++) {
      if (q == 13) continue;
      const int to = c.neighbour[q], from = c.neighbour[26 - q];
      if (to >= 0 && to != c.rank) EPB_NCCL(h, ncclSend(d_send + q, 1, ncclInt32, to, comm, h->stream));
      if (from >= 0 && from != c.rank) EPB_NCCL(h, ncclRecv(d_recv + q, 1, ncclInt32, from, comm, h->stream));
    }
    EPB_NCCL(h, ncclGroupEnd());
    h->launches++;
    EPB_CUDA(h, cudaMemcpyAsync(recvc, d_recv, 27 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    EPB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  for (int q = 0; q < 27; q++) {
    send_off[q] = send_tot; send_tot += (size_t)sendc[q] * nv;
    recv_off[q] = recv_tot; recv_tot += (size_t)recvc[q] * nv;
  }
  const long long n_recv = (long long)(recv_tot / nv);
  if (gone_total == 0 && n_recv == 0) return EPB_OK;
  // the receive side cannot refuse either (the senders are committed): arrivals beyond the capacity are dropped
  // by the unpack below and the error is returned once the exchange is complete
  bool cap_fail = false;
  if (!S.slots && S.n - gone_total + n_recv > S.cap) cap_fail = true;
  double *const *arr = S.slots ? S.mbuf[S.mcur] : S.buf[S.cur];
  if (remote && (send_tot || recv_tot)) {
    int rc = ensure_buf(h, &h->sendbuf, &h->sendbuf_elems, send_tot);
    if (rc) return rc;
    rc = ensure_buf(h, &h->recvbuf, &h->recvbuf_elems, recv_tot);
    if (rc) return rc;
    for (int q = 0; q < 27; q++) {
      if (!sendc[q]) continue;
      PPackOp O;
      for (int k = 0; k < 7; k++) O.src[k] = arr[k];
      O.nv = nv; O.nd = nd;
      O.idx = h->out_idx + (size_t)q * h->out_cap;
      O.count = sendc[q];
      O.buf = h->sendbuf + send_off[q];
      k_ppack<<<nblk((size_t)sendc[q]), 256, 0, h->stream>>>(O);
      h->launches++;
    }
    ncclComm_t comm = (ncclComm_t)h->nccl;
    EPB_NCCL(h, ncclGroupStart());
    for (int q = 0; q < 27; q++) {
      if (q == 13) continue;
      const int to = c.neighbour[q], from = c.neighbour[26 - q];
      if (sendc[q] && to >= 0 && to != c.rank)
        EPB_NCCL(h, ncclSend(h->sendbuf + send_off[q], (size_t)sendc[q] * nv, ncclDouble, to, comm, h->stream));
      if (recvc[q] && from >= 0 && from != c.rank)
        EPB_NCCL(h, ncclRecv(h->recvbuf + recv_off[q], (size_t)recvc[q] * nv, ncclDouble, from, comm, h->stream));
    }
    EPB_NCCL(h, ncclGroupEnd());
    h->launches++;
  }
  // compaction: holes below n_new are filled with the survivors of the tail [n_new, n_old)
  // (slot columns: the leavers sit in the mover buffer, flagged, and k_deliver skips them)
  if (gone_total > 0 && !S.slots) {
    const long long n_old = S.n, n_new = S.n - gone_total;
    int *ctr = h->d_scratch;  // [0] mover count, [1] fill cursor
    EPB_CUDA(h, cudaMemsetAsync(ctr, 0, 2 * sizeof(int), h->stream));
    k_tail_movers<<<nblk((size_t)gone_total), 256, 0, h->stream>>>(S.gone, n_new, n_old, h->movers, ctr);
    h->launches++;
    for (int q = 0; q < 27; q++) {
      if (!cnt[q]) continue;
      FillOp F;
      for (int k = 0; k < 7; k++) F.a[k] = arr[k];
      F.gone = S.gone;
      F.idx = h->out_idx + (size_t)q * h->out_cap;
      F.count = cnt[q];
      F.n_new = n_new;
      F.movers = h->movers;
      F.key = S.info_valid ? S.key : nullptr;
      F.rank = S.rank;
      F.cursor = ctr + 1;
      k_fill_holes<<<nblk((size_t)cnt[q]), 256, 0, h->stream>>>(F);
      h->launches++;
    }
    S.n = n_new;
    if (S.n_sorted > S.n) S.n_sorted = S.n;
  }
  // append arrivals in the reference's direction order (boundary.F90:1436-1446)
  for (int q = 0; q < 27; q++) {
    if (!recvc[q]) continue;
    PUnpackOp O;
    memset(&O, 0, sizeof O);
    for (int k = 0; k < 7; k++) O.dst[k] = arr[k];
    O.nv = nv; O.nd = nd;
    O.first = S.n;
    O.key = (!S.slots && S.info_valid) ? S.key : nullptr;
    O.count = recvc[q];
    if (!S.slots && S.n + recvc[q] > S.cap) O.count = (int)std::max<long long>(0, S.cap - S.n);
    O.buf = h->recvbuf + recv_off[q];
    if (S.slots) {
      O.first_dev = S.mcount + S.mcur;
      O.flag = S.mflag[S.mcur];
      O.cap = (int)S.mcap;
      O.err = h->d_err;
    }
    if (O.count > 0) {
      k_punpack<<<nblk((size_t)O.count), 256, 0, h->stream>>>(O);
      h->launches++;
    }
    if (S.slots) { k_bump_count<<<1, 1, 0, h->stream>>>(S.mcount + S.mcur, recvc[q]); h->launches++; }
    else S.n += O.count;
  }
  EPB_CUDA(h, cudaMemsetAsync(h->out_count, 0, 27 * sizeof(int), h->stream));
  EPB_CUDA(h, cudaGetLastError());
  if (overflow_q >= 0)
    return epb_fail(h, EPB_ERR_CAPACITY, "species %d: %d particles leave in direction %d > outbox capacity %d", is, overflow_n, overflow_q, h->out_cap);
  if (cap_fail)
    return epb_fail(h, EPB_ERR_CAPACITY, "species %d: more particles after migration than the capacity %lld", is, S.cap);
  return EPB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Dynamic load balancing, device half: the data movement of balance_workload
// ---------------------------------------------------------------------------------------------------------
namespace {

struct Box { int lo[3], hi[3]; };   // global cell indices, inclusive

// the cells rank `r` owns in decomposition D, widened by the ghost cells at the physical domain edges
// (redistribute_field_2d: ng0 / ng1, balance.F90:1357-1362)
Box owned_box(const epb_decomp &D, const epb_config &c, int r) {
  Box b;
  const int npx = D.nproc[0], npy = c.ndims >= 2 ? D.nproc[1] : 1;
  const int co[3] = {r % npx, (r / npx) % npy, r / (npx * npy)};
  for (int d = 0; d < 3; d++) {
    if (d >= c.ndims) { b.lo[d] = b.hi[d] = 1; continue; }
    b.lo[d] = D.cell_min[d][co[d]];
    b.hi[d] = D.cell_max[d][co[d]];
    if (b.lo[d] == 1) b.lo[d] -= NG;
    if (b.hi[d] == c.n_global[d]) b.hi[d] += NG;
  }
  return b;
}
bool intersect(const Box &a, const Box &b, Box &o) {
  for (int d = 0; d < 3; d++) {
    o.lo[d] = std::max(a.lo[d], b.lo[d]);
    o.hi[d] = std::min(a.hi[d], b.hi[d]);
    if (o.lo[d] > o.hi[d]) return false;
  }
  return true;
}
size_t box_cells(const Box &b) {
  size_t n = 1;
  for (int d = 0; d < 3; d++) n *= (size_t)(b.hi[d] - b.lo[d] + 1);
  return n;
}
// pack / unpack of a global box on a handle's arrays (local index = global index - (first owned cell - 1))
void run_box(epb_handle *h, const int gmin[3], int f0, int nf, const Box &b, double *buf, int mode, cudaStream_t st) {
  PackOp B;
  B.nf = nf;
  B.nd = h->cfg.ndims;
  for (int q = 0; q < nf; q++) B.f[q] = h->f(f0 + q);
  for (int q = 0; q < 3; q++) {
    B.sz[q] = h->sz[q];
    B.lo[q] = q < h->cfg.ndims ? b.lo[q] - (gmin[q] - 1) : 1;
    B.ext[q] = b.hi[q] - b.lo[q] + 1;
  }
  B.buf = buf;
  B.mode = mode;
  k_pack<<<nblk(box_cells(b)), 256, 0, st>>>(B);
  h->launches++;
}

// get_particle_processor (balance.F90:2095-2151): the processor coordinate whose [minpos, maxpos) holds the position
struct DestOp {
  const double *x[3];
  long long n;
  int nd, nproc[3];
  double minpos[3][32], maxpos[3][32];
  int *dest;
  int *count;      // [nranks]
  int *err;
  int drop;        // moving window, x_min ranks: particles with x < drop_below are not kept (remove_particles, window.F90:324-345)
  double drop_below;
};
__global__ void __launch_bounds__(256) k_dest(const __grid_constant__ DestOp D) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < D.n; i += (long long)gridDim.x * blockDim.x) {
    if (D.drop && D.x[0][i] < D.drop_below) { D.dest[i] = -1; continue; }
    int co[3] = {0, 0, 0};
    bool ok = true;
    for (int d = 0; d < D.nd; d++) {
      const double p = D.x[d][i];
      int found = -1;
      for (int ip = 0; ip < D.nproc[d]; ip++)
        if (p >= D.minpos[d][ip] && p < D.maxpos[d][ip]) { found = ip; break; }
      if (found < 0) ok = false;
      co[d] = found;
    }
    int r = -1;
    if (ok) {
      r = (co[2] * (D.nd >= 2 ? D.nproc[1] : 1) + co[1]) * D.nproc[0] + co[0];
      atomicAdd(&D.count[r], 1);
    } else {
      atomicOr(D.err, 4);   // "Unlocatable particle"
    }
    D.dest[i] = r;
  }
}
struct RouteOp {
  const double *src[7];
  const int *dest;
  const int *offset;   // [nranks] first slot of every destination in buf
  int *cursor;         // [nranks]
  long long n;
  int nd, nv;
  double *buf;         // AoS, pack_particle order, grouped by destination rank
};
__global__ void __launch_bounds__(256) k_route(const __grid_constant__ RouteOp R) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R.n; i += (long long)gridDim.x * blockDim.x) {
    const int r = R.dest[i];
    if (r < 0) continue;
    const int slot = R.offset[r] + atomicAdd(&R.cursor[r], 1);
    double *o = R.buf + (size_t)slot * R.nv;
    int q = 0;
    for (int d = 0; d < R.nd; d++) o[q++] = R.src[d][i];
    for (int d = 3; d < 7; d++) o[q++] = R.src[d][i];
  }
}
__global__ void k_absmax(const double *a, size_t n, unsigned long long *out) {
  double m = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmax(m, fabs(a[i]));
  if (m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// shift_fields on an x_max rank after the arrays have moved one cell to the left (window.F90:126-143, epoch1d
// :114-127, epoch3d :141-160): the outermost ghost layer keeps what it held (shift_field's loop stops one short), the
// incoming cell takes the boundary snapshots and its neighbours are averaged.  One thread per transverse point.
struct WindowFixOp {
  double *f[9];
  const double *snap;   // [6][plane] of the x_max side
  int nd, sz[3], nx;
  size_t plane;
};
__global__ void __launch_bounds__(256) k_window_fix(const __grid_constant__ WindowFixOp O) {
  const int ey_ = O.nd >= 2 ? O.sz[1] : 1, ez_ = O.nd >= 3 ? O.sz[2] : 1;
  const size_t total = (size_t)ey_ * ez_;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % ey_) + 1 - NG, k = (int)(t / ey_) + 1 - NG;
    const int jj = O.nd >= 2 ? j : 1, kk = O.nd >= 3 ? k : 1;
    const int nx = O.nx;
    size_t o = (size_t)(nx + NG - 1);                 // cell (nx, j, k); x is the fastest index: cell nx + a is at o + a
    if (O.nd >= 2) o += (size_t)O.sz[0] * (size_t)(jj + NG - 1);
    if (O.nd >= 3) o += (size_t)O.sz[0] * (size_t)O.sz[1] * (size_t)(kk + NG - 1);
    for (int q = 0; q < 9; q++) O.f[q][o + NG] = O.f[q][o + NG - 1];
#define SN(q) O.snap[(size_t)(q) * O.plane + t]
    double *ex = O.f[EPB_EX], *ey = O.f[EPB_EY], *ez = O.f[EPB_EZ], *bx = O.f[EPB_BX], *by = O.f[EPB_BY], *bz = O.f[EPB_BZ];
    ex[o] = SN(EPB_EX);
    ex[o + 1] = SN(EPB_EX);
    ey[o + 1] = SN(EPB_EY);
    ez[o + 1] = SN(EPB_EZ);
    ex[o - 1] = 0.5 * (ex[o - 2] + ex[o]);
    ey[o] = 0.5 * (ey[o - 1] + ey[o + 1]);
    ez[o] = 0.5 * (ez[o - 1] + ez[o + 1]);
    bx[o + 1] = SN(EPB_BX);
    by[o] = SN(EPB_BY);
    bz[o] = SN(EPB_BZ);
    bx[o] = 0.5 * (bx[o - 1] + bx[o + 1]);
    by[o - 1] = 0.5 * (by[o - 2] + by[o]);
    bz[o - 1] = 0.5 * (bz[o - 2] + bz[o]);
#undef SN
  }
}

// what makes a redistribution a window shift: the x geometry of the new window
struct WindowArgs { double x_grid_min; };

int redistribute_impl(epb_handle *oh, const epb_decomp *od, const epb_decomp *nd_, const epb_config *ncfg,
                      const epb_species *nsp, epb_handle **out, const WindowArgs *win);

}  // namespace

extern "C" int epb_redistribute(epb_handle *oh, const epb_decomp *od, const epb_decomp *nd_, const epb_config *ncfg,
                                const epb_species *nsp, epb_handle **out) {
  return redistribute_impl(oh, od, nd_, ncfg, nsp, out, nullptr);
}

// One cell of shift_window (housekeeping/window.F90:62-94) plus the particle_bcs that follows it (:383-385), as a
// redistribution onto the same decomposition one cell further along x: every field cell goes to the rank that owns
// it in the new window (shift_field + field_bc), the incoming cell of the x_max ranks is fixed up (k_window_fix),
// particles left of the new x_min are dropped on the x_min ranks (remove_particles) and the others go to the rank
// whose [x_min_local, x_max_local) of the NEW grid holds them.  insert_particles is the host's (epb_append_species
// afterwards).  Collective; *out replaces old_h.
extern "C" int epb_shift_window(epb_handle *oh, const epb_decomp *d, const epb_config *ncfg, const epb_species *nsp,
                                double x_grid_min, epb_handle **out) {
  if (!oh || !d || !ncfg || !out) return EPB_ERR_ARG;
  for (int s = 0; s < 2; s++)
    if (oh->cfg.bc_field[s] == EPB_BC_PERIODIC || ncfg->bc_field[s] == EPB_BC_PERIODIC)
      return epb_fail(oh, EPB_ERR_UNSUPPORTED, "epb_shift_window: the window moves along a non-periodic x only");
  for (int q = 0; q < 3; q++)
    if (ncfg->n[q] != oh->cfg.n[q] || ncfg->n_global[q] != oh->cfg.n_global[q])
      return epb_fail(oh, EPB_ERR_ARG, "epb_shift_window: the decomposition must stay as it is");
  WindowArgs w;
  w.x_grid_min = x_grid_min;
  return redistribute_impl(oh, d, d, ncfg, nsp, out, &w);
}

namespace {

int redistribute_impl(epb_handle *oh, const epb_decomp *od, const epb_decomp *nd_, const epb_config *ncfg,
                      const epb_species *nsp, epb_handle **out, const WindowArgs *win) {
  if (!oh || !od || !nd_ || !ncfg || !out) return EPB_ERR_ARG;
  *out = nullptr;
  const epb_config &oc = oh->cfg;
  const int nd = oc.ndims, nranks = oc.nranks, me = oc.rank;
  if (ncfg->ndims != nd || ncfg->nranks != nranks || ncfg->rank != me || ncfg->n_species != oc.n_species)
    return epb_fail(oh, EPB_ERR_ARG, "epb_redistribute: the new config describes another run");
  int nprod = 1;
  for (int d = 0; d < nd; d++) {
    if (od->nproc[d] != nd_->nproc[d] || od->nproc[d] > 32)
      return epb_fail(oh, EPB_ERR_ARG, "epb_redistribute: processor grid must stay the same (<= 32 per axis)");
    nprod *= od->nproc[d];
  }
  if (nprod != nranks || nranks > 64) return epb_fail(oh, EPB_ERR_ARG, "epb_redistribute: nproc does not match nranks (<= 64)");
  if (nranks > 1 && !oh->nccl) return epb_fail(oh, EPB_ERR_NCCL, "epb_redistribute: epb_set_comm was not called");
  if (oh->cpml) return epb_fail(oh, EPB_ERR_UNSUPPORTED, "epb_redistribute: the CPML auxiliary arrays are not re-cut (balance.F90:600-700 remaps them)");
  EPB_CUDA(oh, cudaStreamSynchronize(oh->stream));
  // The snapshots setup_field_boundaries took of the initial fields on laser / outflow faces are not moved: refuse
  // the (rare) deck that has non-zero ones instead of silently zeroing them
  {
    unsigned long long *d_m = (unsigned long long *)(oh->d_scratch + 768);
    EPB_CUDA(oh, cudaMemsetAsync(d_m, 0, sizeof *d_m, oh->stream));
    k_absmax<<<nblk(12 * oh->plane), 256, 0, oh->stream>>>(oh->snap, 12 * oh->plane, d_m);
    for (int a = 1; a < nd; a++)
      if (oh->snapA[a]) k_absmax<<<nblk(12 * oh->planeA[a]), 256, 0, oh->stream>>>(oh->snapA[a], 12 * oh->planeA[a], d_m);
    unsigned long long m = 0;
    EPB_CUDA(oh, cudaMemcpyAsync(&m, d_m, sizeof m, cudaMemcpyDeviceToHost, oh->stream));
    EPB_CUDA(oh, cudaStreamSynchronize(oh->stream));
    if (m != 0)
      return epb_fail(oh, EPB_ERR_UNSUPPORTED, "epb_redistribute: non-zero initial fields on a laser / outflow boundary (their snapshots are not re-cut)");
  }
  // EPB_DEBUG=1 EPB_REDIST_TIMING=1: wall time of the phases on stderr (each one closed by a device synchronize)
  static const bool timing = epb_env("EPB_REDIST_TIMING") && atoi(epb_env("EPB_REDIST_TIMING"));
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    cudaDeviceSynchronize();
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[epb redistribute%s rank %d] %-10s %8.3f ms\n", win ? " (window)" : "", me, what,
            std::chrono::duration<double, std::milli>(t - t_last).count());
    t_last = t;
  };
  lap("checks");
  epb_handle *nh = nullptr;
  int rc = epb_create(ncfg, nsp, &nh);
  if (rc) return epb_fail(oh, rc, "epb_redistribute: epb_create for the new decomposition failed");
  for (int is = 0; is < (int)oh->sp.size() && !rc; is++)   // columns as deep as the ones they replace
    if (oh->sp[is].slots && nh->sp[is].slots) rc = epb_slots_ensure_rows(nh, is, oh->sp[is].R);
  if (rc) { oh->err = nh->err; epb_destroy(nh); return rc; }
  lap("create");
  nh->nccl = oh->nccl;   // the communicator moves to the new state (same ranks, same Cartesian topology)
  nh->time_push = oh->time_push;
  ncclComm_t comm = (ncclComm_t)nh->nccl;
  cudaStream_t st = nh->stream;
  auto fail = [&](int code) { nh->nccl = nullptr; oh->err = nh->err; epb_destroy(nh); return code; };

  // ---- fields: every cell goes from its old owner to its new owner --------------------------------------------
  {
    int omin[3] = {1, 1, 1}, nmin[3] = {1, 1, 1};
    {
      const int npx = od->nproc[0], npy = nd >= 2 ? od->nproc[1] : 1;
      const int co[3] = {me % npx, (me / npx) % npy, me / (npx * npy)};
      for (int d = 0; d < nd; d++) { omin[d] = od->cell_min[d][co[d]]; nmin[d] = nd_->cell_min[d][co[d]]; }
    }
    // window shift: in the numbering of the old window the new one owns the same cells one further along x
    const int xs = win ? 1 : 0;
    nmin[0] += xs;
    auto new_box = [&](int p) {
      Box b = owned_box(*nd_, oc, p);
      b.lo[0] += xs; b.hi[0] += xs;
      return b;
    };
    const Box my_old = owned_box(*od, oc, me), my_new = new_box(me);
    std::vector<Box> sbox(nranks), rbox(nranks);
    std::vector<char> shas(nranks, 0), rhas(nranks, 0);
    std::vector<size_t> soff(nranks + 1, 0), roff(nranks + 1, 0);
    for (int p = 0; p < nranks; p++) {
      shas[p] = intersect(my_old, new_box(p), sbox[p]);
      rhas[p] = intersect(my_new, owned_box(*od, oc, p), rbox[p]);
      soff[p + 1] = soff[p] + (shas[p] ? 3 * box_cells(sbox[p]) : 0);
      roff[p + 1] = roff[p] + (rhas[p] ? 3 * box_cells(rbox[p]) : 0);
    }
    rc = ensure_buf(nh, &nh->sendbuf, &nh->sendbuf_elems, soff[nranks]);
    if (!rc) rc = ensure_buf(nh, &nh->recvbuf, &nh->recvbuf_elems, roff[nranks]);
    if (rc) return fail(rc);
    for (int f0 = 0; f0 < 9; f0 += 3) {   // E, B, J: three components per message
      for (int p = 0; p < nranks; p++)
        if (shas[p]) run_box(oh, omin, f0, 3, sbox[p], nh->sendbuf + soff[p], 0, st);
      if (nranks > 1) {
        ncclGroupStart();
        for (int p = 0; p < nranks; p++) {
          if (p == me) continue;
          if (shas[p]) ncclSend(nh->sendbuf + soff[p], soff[p + 1] - soff[p], ncclDouble, p, comm, st);
          if (rhas[p]) ncclRecv(nh->recvbuf + roff[p], roff[p + 1] - roff[p], ncclDouble, p, comm, st);
        }
        ncclResult_t r = ncclGroupEnd();
        if (r != ncclSuccess) { epb_fail(nh, EPB_ERR_NCCL, "epb_redistribute (fields): %s", ncclGetErrorString(r)); return fail(EPB_ERR_NCCL); }
        nh->launches++;
      }
      for (int p = 0; p < nranks; p++) {
        if (!rhas[p]) continue;
        // what stays on this rank is copied through the send buffer (same box on both sides)
        const double *src = (p == me) ? nh->sendbuf + soff[p] : nh->recvbuf + roff[p];
        run_box(nh, nmin, f0, 3, rbox[p], const_cast<double *>(src), 1, st);
      }
      rc = epb_halo_exchange(nh, f0, 3, false);   // do_field_mpi_with_lengths (remap_field, balance.F90:1082)
      if (rc) return fail(rc);
    }
    if (win && ncfg->is_boundary[1]) {
      WindowFixOp W;
      for (int q = 0; q < 9; q++) W.f[q] = nh->f(q);
      W.snap = nh->snap + (size_t)6 * nh->plane;
      W.nd = nd;
      for (int q = 0; q < 3; q++) W.sz[q] = nh->sz[q];
      W.nx = ncfg->n[0];
      W.plane = nh->plane;
      k_window_fix<<<nblk(nh->plane), 256, 0, st>>>(W);
      nh->launches++;
    }
  }

  lap("fields");
  // ---- particles: distribute_particles ------------------------------------------------------------------------
  const int nv = nd + 4;
  const long long CH = 4 << 20;
  int *d_dest = nullptr, *d_cnt = nullptr;   // d_cnt: [nranks + 1] counts + "more chunks" flag, then [nranks][nranks + 1], offsets, cursors
  if (cudaMalloc(&d_dest, (size_t)CH * sizeof(int)) != cudaSuccess ||
      cudaMalloc(&d_cnt, (size_t)(nranks + 1) * (nranks + 4) * sizeof(int)) != cudaSuccess) {
    cudaFree(d_dest);
    epb_fail(nh, EPB_ERR_CUDA, "epb_redistribute: scratch allocation failed");
    return fail(EPB_ERR_CUDA);
  }
  int *d_all = d_cnt + (nranks + 1), *d_off = d_all + (nranks + 1) * nranks, *d_cur = d_off + (nranks + 1);
  auto cleanup = [&]() { cudaFree(d_dest); cudaFree(d_cnt); };
  DestOp D;
  memset(&D, 0, sizeof D);
  D.nd = nd;
  for (int d = 0; d < nd; d++) {
    D.nproc[d] = nd_->nproc[d];
    const double dx = oc.dx[d];
    // setup.F90:169,180; the moving window accumulates its own x_grid_min (window.F90:74)
    const double x_grid_min = (win && d == 0) ? win->x_grid_min : oc.gmin[d] + dx / 2.0;
    const int np = nd_->nproc[d];
    for (int ip = 0; ip < np; ip++) {
      const double gmin_ip = x_grid_min + (double)(nd_->cell_min[d][ip] - 1) * dx;   // x_grid_mins(iproc)
      const double gmax_ip = x_grid_min + (double)(nd_->cell_max[d][ip] - 1) * dx;   // x_grid_maxs(iproc)
      D.minpos[d][ip] = ip == 0 ? gmin_ip - dx * (0.5 + 3.0) : gmin_ip - dx * 0.5;    // png = 3
      D.maxpos[d][ip] = ip == np - 1 ? gmax_ip + dx * (0.5 + 3.0) : gmax_ip + dx * 0.5;
    }
  }
  D.dest = d_dest;
  D.count = d_cnt;
  D.drop = (win && oc.is_boundary[0]) ? 1 : 0;
  D.drop_below = ncfg->gmin[0];
  if (!nh->d_err) {
    if (cudaMalloc(&nh->d_err, sizeof(int)) != cudaSuccess) { cleanup(); return fail(EPB_ERR_CUDA); }
    cudaMemsetAsync(nh->d_err, 0, sizeof(int), st);
  }
  D.err = nh->d_err;
  std::vector<int> hall((size_t)(nranks + 1) * nranks);
  for (int is = 0; is < (int)oh->sp.size(); is++) {
    SpeciesIter I;
    rc = epb_species_iter_begin(oh, is, I);
    if (rc) { cleanup(); return fail(rc); }
    bool mine_done = false;
    for (;;) {
      double *sa[7];
      long long m = 0;
      if (!mine_done) {
        rc = epb_species_iter_next(oh, is, I, CH, sa, &m);   // gathers on oh->stream
        if (rc) { cleanup(); return fail(rc); }
        if (m == 0) mine_done = true;
        cudaStreamSynchronize(oh->stream);
      }
      cudaMemsetAsync(d_cnt, 0, (size_t)(nranks + 1) * sizeof(int), st);
      if (m > 0) {
        for (int d = 0; d < 3; d++) D.x[d] = sa[d];
        D.n = m;
        k_dest<<<nblk((size_t)m), 256, 0, st>>>(D);
        nh->launches++;
      }
      const int more = mine_done ? 0 : 1;
      cudaMemcpyAsync(d_cnt + nranks, &more, sizeof(int), cudaMemcpyHostToDevice, st);
      if (nranks > 1) {
        ncclResult_t r = ncclAllGather(d_cnt, d_all, nranks + 1, ncclInt32, comm, st);
        if (r != ncclSuccess) { cleanup(); epb_fail(nh, EPB_ERR_NCCL, "epb_redistribute (counts): %s", ncclGetErrorString(r)); return fail(EPB_ERR_NCCL); }
      } else {
        cudaMemcpyAsync(d_all, d_cnt, (size_t)(nranks + 1) * sizeof(int), cudaMemcpyDeviceToDevice, st);
      }
      cudaMemcpyAsync(hall.data(), d_all, hall.size() * sizeof(int), cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) { cleanup(); epb_fail(nh, EPB_ERR_CUDA, "epb_redistribute: %s", cudaGetErrorString(cudaGetLastError())); return fail(EPB_ERR_CUDA); }
      bool any_more = false;
      for (int p = 0; p < nranks; p++) any_more = any_more || hall[(size_t)p * (nranks + 1) + nranks];
      // send counts = my row, receive counts = my column
      std::vector<int> hoff(nranks + 1, 0), roffp(nranks + 1, 0);
      for (int p = 0; p < nranks; p++) {
        hoff[p + 1] = hoff[p] + hall[(size_t)me * (nranks + 1) + p];
        roffp[p + 1] = roffp[p] + (p == me ? 0 : hall[(size_t)p * (nranks + 1) + me]);
      }
      if (hoff[nranks] > 0 || roffp[nranks] > 0) {
        rc = ensure_buf(nh, &nh->sendbuf, &nh->sendbuf_elems, (size_t)hoff[nranks] * nv);
        if (!rc) rc = ensure_buf(nh, &nh->recvbuf, &nh->recvbuf_elems, (size_t)roffp[nranks] * nv);
        if (rc) { cleanup(); return fail(rc); }
        if (m > 0) {
          cudaMemcpyAsync(d_off, hoff.data(), (size_t)(nranks + 1) * sizeof(int), cudaMemcpyHostToDevice, st);
          cudaMemsetAsync(d_cur, 0, (size_t)(nranks + 1) * sizeof(int), st);
          RouteOp R;
          for (int q = 0; q < 7; q++) R.src[q] = sa[q];
          R.dest = d_dest; R.offset = d_off; R.cursor = d_cur; R.n = m; R.nd = nd; R.nv = nv; R.buf = nh->sendbuf;
          k_route<<<nblk((size_t)m), 256, 0, st>>>(R);
          nh->launches++;
        }
        if (nranks > 1) {
          ncclGroupStart();
          for (int p = 0; p < nranks; p++) {
            if (p == me) continue;
            const int sc = hoff[p + 1] - hoff[p], rcnt = roffp[p + 1] - roffp[p];
            if (sc) ncclSend(nh->sendbuf + (size_t)hoff[p] * nv, (size_t)sc * nv, ncclDouble, p, comm, st);
            if (rcnt) ncclRecv(nh->recvbuf + (size_t)roffp[p] * nv, (size_t)rcnt * nv, ncclDouble, p, comm, st);
          }
          ncclResult_t r = ncclGroupEnd();
          if (r != ncclSuccess) { cleanup(); epb_fail(nh, EPB_ERR_NCCL, "epb_redistribute (particles): %s", ncclGetErrorString(r)); return fail(EPB_ERR_NCCL); }
          nh->launches++;
        }
        // what stays here, then what arrived
        rc = epb_species_insert_aos(nh, is, nh->sendbuf + (size_t)hoff[me] * nv, hoff[me + 1] - hoff[me]);
        if (!rc) rc = epb_species_insert_aos(nh, is, nh->recvbuf, roffp[nranks]);
        if (rc) { cleanup(); return fail(rc); }
        // the staging arrays of the old handle are reused by the next chunk: finish with them first
        cudaStreamSynchronize(st);
      }
      if (!any_more) break;
    }
  }
  cleanup();
  {
    int e = 0;
    cudaMemcpyAsync(&e, nh->d_err, sizeof e, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess || e) {
      epb_fail(nh, e ? EPB_ERR_CAPACITY : EPB_ERR_CUDA, e & 4 ? "epb_redistribute: unlocatable particle (balance.F90:2145)" : "epb_redistribute: device error %d", e);
      return fail(e ? EPB_ERR_CAPACITY : EPB_ERR_CUDA);
    }
  }
  oh->nccl = nullptr;   // moved
  lap("particles");
  {
    int rcx = epb_agree_exchange_caps(nh);
    if (rcx) return rcx;
  }
  epb_destroy(oh);
  lap("destroy");
  *out = nh;
  return EPB_OK;
}

}  // namespace
