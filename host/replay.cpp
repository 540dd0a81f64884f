// replay.cpp -- a compiled host on the reference's side of the C ABI (missing #8).
//
// The image has no Fortran compiler, so fortran/epoch_b200_mod.F90 cannot be built here.  This program
// is the same host logic in C++: it holds the particles the way EPOCH does -- one heap node per
// particle in a doubly linked list per species (TYPE particle / particle_list, shared_data.F90:93-171;
// create_allocated_partlist, partlist.F90:89-113) -- and drives libepoch_b200.so through
// include/epoch_b200.h with exactly the call sequence of the shim:
//
//   b200_attach   : epb_abi_info check, epb_create, [epb_set_comm], b200_upload, epb_init_boundaries
//   b200_upload   : six epb_upload_field; per species: walk the list, pack_particle every node into one
//                   buffer (partlist.F90:414-486), epb_upload_species
//   the PIC loop  : epb_fields_half, epb_push, epb_current_finish, [time], epb_fields_final
//                   (epoch2d.F90:211,216,250,265)
//   b200_download : nine epb_download_field; per species: epb_species_count, epb_download_species,
//                   destroy_partlist + create_allocated_partlist + unpack_particle per node
//
// Input: a state file written by tests/test_host_replay.py (or bench.py): the epb_config and epb_species
// structs as raw bytes, the step count, the six field arrays, the packed particles of every species.
// Output: a result file (nine fields + the particles, read back out of the linked lists) and one JSON
// line with the wall time of every phase, which bench.py reports as `e2e_full`.
//
// Build: g++ -O2 -std=c++17 -I../include replay.cpp -o replay -L../epoch_b200 -lepoch_b200 -Wl,-rpath,...
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "epoch_b200.h"

namespace {

constexpr int c_ndims_max = 3;

// TYPE particle (shared_data.F90:93-142, default build: no optional members)
struct particle {
  double part_pos[c_ndims_max];
  double part_p[3];
  double weight;
  particle *next, *prev;
};
// TYPE particle_list (shared_data.F90:159-171)
struct particle_list {
  particle *head = nullptr, *tail = nullptr;
  int64_t count = 0;
};

void create_allocated_partlist(particle_list &l, int64_t n) {  // partlist.F90:89-113: one ALLOCATE per particle
  l.head = l.tail = nullptr;
  l.count = 0;
  for (int64_t i = 0; i < n; i++) {
    particle *p = new particle;
    p->next = nullptr;
    p->prev = l.tail;
    if (l.tail) l.tail->next = p; else l.head = p;
    l.tail = p;
    l.count++;
  }
}
void destroy_partlist(particle_list &l) {  // partlist.F90:348-366
  particle *cur = l.head;
  while (cur) {
    particle *nx = cur->next;
    delete cur;
    cur = nx;
  }
  l.head = l.tail = nullptr;
  l.count = 0;
}
inline void pack_particle(double *a, const particle *p, int nd) {  // partlist.F90:414-486
  int c = 0;
  for (int d = 0; d < nd; d++) a[c++] = p->part_pos[d];
  for (int d = 0; d < 3; d++) a[c++] = p->part_p[d];
  a[c++] = p->weight;
}
inline void unpack_particle(const double *a, particle *p, int nd) {  // partlist.F90:490-564
  int c = 0;
  for (int d = 0; d < nd; d++) p->part_pos[d] = a[c++];
  for (int d = 0; d < 3; d++) p->part_p[d] = a[c++];
  p->weight = a[c++];
}

struct Host {  // the slice of shared_data the path touches
  epb_config cfg;
  std::vector<epb_species> species;
  std::vector<particle_list> lists;            // species_list(:)%attached_list
  std::vector<double> f[EPB_NFIELD];           // ex .. jz, (1-ng:nx+ng, ...)
  size_t fsize = 0;
  int nvar = 0;
  epb_handle *b200 = nullptr;
};

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void b200_check(Host &H, int rc, const char *what) {  // abort_code(c_err_generic_error), utilities.f90:261-281
  if (rc == 0) return;
  std::fprintf(stderr, "replay: %s failed with code %d: %s\n", what, rc, H.b200 ? epb_last_error(H.b200) : "");
  std::exit(2);
}

void b200_upload(Host &H, double *t_pack, double *t_api) {
  for (int q = 0; q < 6; q++) b200_check(H, epb_upload_field(H.b200, q, H.f[q].data()), "epb_upload_field");
  const int nd = H.cfg.ndims;
  for (size_t is = 0; is < H.lists.size(); is++) {
    const int64_t npart = H.lists[is].count;
    double t0 = now();
    std::vector<double> buf((size_t)std::max<int64_t>(npart * H.nvar, 1));
    int64_t ipart = 0;
    for (const particle *cur = H.lists[is].head; cur; cur = cur->next) {
      pack_particle(buf.data() + ipart * H.nvar, cur, nd);
      ipart++;
    }
    double t1 = now();
    b200_check(H, epb_upload_species(H.b200, (int)is, npart, buf.data()), "epb_upload_species");
    b200_check(H, epb_synchronize(H.b200), "epb_synchronize");
    double t2 = now();
    *t_pack += t1 - t0;
    *t_api += t2 - t1;
  }
}

void b200_attach(Host &H, double *t_pack, double *t_api) {
  int32_t info[4];
  b200_check(H, epb_abi_info(info), "epb_abi_info");
  if (info[0] != (int32_t)sizeof(epb_config) || info[1] != (int32_t)sizeof(epb_species) || info[2] != H.cfg.ng) {
    std::fprintf(stderr, "replay: ABI mismatch\n");
    std::exit(2);
  }
  b200_check(H, epb_create(&H.cfg, H.species.data(), &H.b200), "epb_create");
  b200_upload(H, t_pack, t_api);
  b200_check(H, epb_init_boundaries(H.b200), "epb_init_boundaries");
}

void b200_download(Host &H, bool with_particles, double *t_unpack, double *t_api) {
  for (int q = 0; q < EPB_NFIELD; q++) b200_check(H, epb_download_field(H.b200, q, H.f[q].data()), "epb_download_field");
  if (!with_particles) return;
  const int nd = H.cfg.ndims;
  for (size_t is = 0; is < H.lists.size(); is++) {
    double t0 = now();
    int64_t npart = 0;
    b200_check(H, epb_species_count(H.b200, (int)is, &npart), "epb_species_count");
    std::vector<double> buf((size_t)std::max<int64_t>(npart * H.nvar, 1));
    b200_check(H, epb_download_species(H.b200, (int)is, npart, buf.data()), "epb_download_species");
    double t1 = now();
    destroy_partlist(H.lists[is]);
    create_allocated_partlist(H.lists[is], npart);
    int64_t ipart = 0;
    for (particle *cur = H.lists[is].head; cur; cur = cur->next) {
      unpack_particle(buf.data() + ipart * H.nvar, cur, nd);
      ipart++;
    }
    double t2 = now();
    *t_api += t1 - t0;
    *t_unpack += t2 - t1;
  }
}

bool rd(FILE *f, void *p, size_t n) { return std::fread(p, 1, n, f) == n; }
bool wr(FILE *f, const void *p, size_t n) { return std::fwrite(p, 1, n, f) == n; }

}  // namespace

int main(int argc, char **argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: replay <state file> <result file | -> [--no-particle-download]\n");
    return 1;
  }
  const bool skip_pd = argc > 3 && std::string(argv[3]) == "--no-particle-download";
  Host H;
  FILE *fi = std::fopen(argv[1], "rb");
  if (!fi) { std::perror(argv[1]); return 1; }
  char magic[8];
  int32_t hdr[4];   // sizeof(epb_config), sizeof(epb_species), n_species, nsteps
  if (!rd(fi, magic, 8) || std::memcmp(magic, "EPBRPLY1", 8) != 0 || !rd(fi, hdr, sizeof hdr)) {
    std::fprintf(stderr, "replay: bad state file\n");
    return 1;
  }
  if (hdr[0] != (int32_t)sizeof(epb_config) || hdr[1] != (int32_t)sizeof(epb_species)) {
    std::fprintf(stderr, "replay: the state file was written for another ABI (%d/%d vs %zu/%zu)\n", hdr[0], hdr[1],
                 sizeof(epb_config), sizeof(epb_species));
    return 1;
  }
  const int nsp = hdr[2], nsteps = hdr[3];
  H.species.resize(std::max(1, nsp));
  if (!rd(fi, &H.cfg, sizeof H.cfg) || !rd(fi, H.species.data(), sizeof(epb_species) * nsp)) return 1;
  const int nd = H.cfg.ndims;
  H.nvar = nd + 4;
  H.fsize = 1;
  for (int d = 0; d < nd; d++) H.fsize *= (size_t)(H.cfg.n[d] + 2 * H.cfg.ng);
  for (int q = 0; q < EPB_NFIELD; q++) H.f[q].assign(H.fsize, 0.0);
  for (int q = 0; q < 6; q++)
    if (!rd(fi, H.f[q].data(), H.fsize * sizeof(double))) return 1;
  // the loader's job (auto_load, helper.F90:95): the particles arrive packed and are put on the lists
  H.lists.resize(nsp);
  double t_load = now();
  int64_t ntotal = 0;
  for (int is = 0; is < nsp; is++) {
    int64_t n = 0;
    if (!rd(fi, &n, sizeof n)) return 1;
    create_allocated_partlist(H.lists[is], n);
    std::vector<double> chunk((size_t)H.nvar * (1 << 20));
    particle *cur = H.lists[is].head;
    for (int64_t i0 = 0; i0 < n; i0 += (1 << 20)) {
      const int64_t m = std::min<int64_t>(1 << 20, n - i0);
      if (!rd(fi, chunk.data(), (size_t)m * H.nvar * sizeof(double))) return 1;
      for (int64_t i = 0; i < m; i++, cur = cur->next) unpack_particle(chunk.data() + i * H.nvar, cur, nd);
    }
    ntotal += n;
  }
  std::fclose(fi);
  t_load = now() - t_load;

  double t_pack = 0, t_up = 0, t_unpack = 0, t_down = 0;
  double t0 = now();
  b200_attach(H, &t_pack, &t_up);
  double t_attach = now() - t0;

  t0 = now();
  for (int step = 0; step < nsteps; step++) {   // epoch2d.F90:190-268 without lasers / diagnostics
    b200_check(H, epb_fields_half(H.b200), "epb_fields_half");
    b200_check(H, epb_push(H.b200), "epb_push");
    b200_check(H, epb_current_finish(H.b200), "epb_current_finish");
    b200_check(H, epb_fields_final(H.b200), "epb_fields_final");
  }
  b200_check(H, epb_synchronize(H.b200), "epb_synchronize");
  double t_steps = now() - t0;

  t0 = now();
  b200_download(H, !skip_pd, &t_unpack, &t_down);
  double t_download = now() - t0;
  const int64_t launches = epb_launch_count(H.b200);

  if (std::string(argv[2]) != "-") {
    FILE *fo = std::fopen(argv[2], "wb");
    if (!fo) { std::perror(argv[2]); return 1; }
    wr(fo, "EPBRSLT1", 8);
    for (int q = 0; q < EPB_NFIELD; q++) wr(fo, H.f[q].data(), H.fsize * sizeof(double));
    for (int is = 0; is < nsp; is++) {
      const int64_t n = H.lists[is].count;
      wr(fo, &n, sizeof n);
      std::vector<double> chunk((size_t)H.nvar * (1 << 20));
      int64_t k = 0;
      for (const particle *cur = H.lists[is].head; cur; cur = cur->next) {
        pack_particle(chunk.data() + k * H.nvar, cur, nd);
        if (++k == (1 << 20)) { wr(fo, chunk.data(), (size_t)k * H.nvar * sizeof(double)); k = 0; }
      }
      if (k) wr(fo, chunk.data(), (size_t)k * H.nvar * sizeof(double));
    }
    std::fclose(fo);
  }
  int64_t nend = 0;
  for (auto &l : H.lists) nend += l.count;
  std::printf("{\"particles\": %lld, \"particles_end\": %lld, \"steps\": %d, \"attach_s\": %.6f, \"list_to_packed_s\": %.6f, "
              "\"upload_api_s\": %.6f, \"steps_s\": %.6f, \"download_s\": %.6f, \"download_api_s\": %.6f, "
              "\"packed_to_list_s\": %.6f, \"loader_s\": %.6f, \"bytes_per_particle\": %d, \"gpu_launches\": %lld}\n",
              (long long)ntotal, (long long)nend, nsteps, t_attach, t_pack, t_up, t_steps, t_download, t_down, t_unpack,
              t_load, H.nvar * 8, (long long)launches);
  b200_check(H, epb_destroy(H.b200), "epb_destroy");
  for (auto &l : H.lists) destroy_partlist(l);
  return 0;
}
