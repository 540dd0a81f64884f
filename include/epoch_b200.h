/* epoch_b200.h — C ABI of the B200-native PIC hot path (drop-in boundary).
 *
 * EPOCH has no plugin/FFI layer: PROGRAM pic calls argument-less module
 * procedures that work on shared_data globals (epoch2d/src/epoch2d.F90:211,216,
 * 250,265).  "Drop-in" therefore means a Fortran shim whose procedure bodies call
 * these entry points through ISO_C_BINDING (fortran/epoch_b200_mod.F90,
 * INTEGRATION.md).  Every entry point names the reference routine it replaces.
 *
 * Conventions
 *  - all functions return 0 on success, non-zero on error (epb_last_error());
 *    EPB_ERR_UNSUPPORTED is returned by epb_create for configurations the device
 *    path does not implement, instead of silently diverging (the shim maps any
 *    non-zero code to abort_code(c_err_generic_error), utilities.f90:261-281).
 *  - host arrays are the Fortran allocatables passed with C_LOC: column-major,
 *    full extent incl. ghost cells, field(1-ng:nx+ng [,1-ng:ny+ng [,1-ng:nz+ng]])
 *    (mpi_routines.F90:379-388).  The library copies; it never keeps a host pointer.
 *  - particle blocks use the pack_particle wire layout (partlist.F90:414-486,
 *    default flags): nvar = ndims + 4 doubles per particle:
 *    pos(1..ndims), p(1..3), weight.
 *  - one host thread <-> one handle <-> one GPU <-> one rank.
 */
#ifndef EPOCH_B200_H
#define EPOCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct epb_handle epb_handle;

enum {
  EPB_OK = 0,
  EPB_ERR_ARG = 1,
  EPB_ERR_CUDA = 2,
  EPB_ERR_UNSUPPORTED = 3,
  EPB_ERR_CAPACITY = 4,
  EPB_ERR_NCCL = 5,
};

/* boundary codes, identical to constants.F90:75-90 */
enum {
  EPB_BC_PERIODIC = 1,
  EPB_BC_OTHER = 2,
  EPB_BC_SIMPLE_LASER = 3,
  EPB_BC_SIMPLE_OUTFLOW = 4,
  EPB_BC_OPEN = 5,
  EPB_BC_ZERO_GRADIENT = 7,
  EPB_BC_CLAMP = 8,
  EPB_BC_REFLECT = 9,
  EPB_BC_CONDUCT = 10,
  EPB_BC_THERMAL = 11,
  EPB_BC_CPML_LASER = 12,
  EPB_BC_CPML_OUTFLOW = 13,
};

/* field ids for upload/download */
enum { EPB_EX = 0, EPB_EY, EPB_EZ, EPB_BX, EPB_BY, EPB_BZ, EPB_JX, EPB_JY, EPB_JZ, EPB_NFIELD };

/* Mirror of the shared_data globals the hot path reads (shared_data.F90:484-683).
 * All values are the caller's: the library derives nothing that would change
 * bit-level bookkeeping (x_min_local etc. come from utilities.f90:343-421). */
typedef struct epb_config {
  int32_t ndims;            /* c_ndims: 1, 2 or 3 */
  int32_t n[3];             /* nx, ny, nz of this rank */
  int32_t n_global[3];      /* nx_global ... */
  int32_t ng;               /* ghost cells; must be 5 (triangle shape: png+2) */
  int32_t bc_field[6];      /* bc_field(c_bd_x_min..c_bd_z_max) after setup_boundaries */
  int32_t is_boundary[6];   /* x_min_boundary, x_max_boundary, ... */
  int32_t neighbour[27];    /* neighbour(ix,iy,iz) at [(iz+1)*9+(iy+1)*3+(ix+1)], -1 = MPI_PROC_NULL */
  int32_t rank, nranks;
  int32_t n_species;
  int32_t strict_fp;        /* 1: kernels built without FMA contraction (bit-level parity build) */
  int32_t sort_interval;    /* steps between on-GPU counting sorts; 0 = library default (2 or 8, by kernel) */
  int32_t field_order;      /* 0 or 2, 4, 6: finite-difference order of the Yee solver (fields.f90:32-46) */
  int32_t maxwell_solver;   /* c_maxwell_solver_* (constants.F90:173-180): 0 yee; -1 custom, 2..4 lehe_x/y/z, 5 cowan, 6 pukhov: extended B stencil, order 2 (fields.f90:51-100, epoch3d :53-162, epoch1d :48-62) */
  int32_t smooth_its;       /* smooth_currents: smooth_its passes (0 = off), current_smooth.F90:50-141 */
  int32_t smooth_comp_its;  /* smooth_compensation: 0 or 1 */
  int32_t smooth_strides;   /* up to 4 strides (1..5), one per nibble, low nibble first; 0 = stride 1 */
  int32_t hc_push;          /* -DHC_PUSH: Higuera-Cary gamma in the momentum rotation (particles.F90:386-398); fills the
                               former padding word before dx, sizeof(epb_config) is unchanged */
  double dx[3];             /* dx, dy, dz */
  double dt;
  double grid_min_local[3]; /* x_grid_min_local ... (cell centre of local cell 1) */
  double min_local[3];      /* x_min_local ... */
  double max_local[3];      /* x_max_local ... */
  double gmin[3], gmax[3];  /* x_min, x_max ... (global domain) */
  double min_outer[3];      /* x_min_outer ... (utilities.f90:367-369) */
  double max_outer[3];
  double stencil[15];       /* as set_maxwell_solver leaves them: alphax..z, betaxy, betaxz, betayx, betayz, betazx, betazy, gammax..z, deltax..z */
  /* CPML boundaries (bc_field 12 cpml_laser / 13 cpml_outflow; boundary.F90:1479-2025).  n, n_global, grid_min_local,
   * min_local / max_local (with the cpml offsets of utilities.f90:364-365) and min_outer / max_outer are the
   * caller's as ever; the library restates set_cpml_helpers from these four numbers and n_global_min. */
  double cpml_kappa_max, cpml_a_max, cpml_sigma_max;   /* shared_data.F90:451 */
  int32_t cpml_thickness;   /* cells; 0 unless a field boundary is a CPML (mpi_routines.F90:285) */
  int32_t n_global_min[3];  /* nx_global_min ...: global index of this rank's first cell */
} epb_config;

/* Mirror of TYPE particle_species (shared_data.F90:194-285), hot-path members only */
typedef struct epb_species {
  double charge;            /* C */
  double mass;              /* kg */
  int32_t bc_particle[6];   /* after setup_particle_boundary (boundary.F90:99-139) */
  int32_t zero_current;     /* tracer species: pushed, no current */
  int32_t immobile;
  int64_t capacity;         /* device slots to reserve for this species on this rank */
} epb_species;

/* -- lifetime ----------------------------------------------------------------
 * replaces: mpi_initialise's ALLOCATE of ex..jz (mpi_routines.F90:379-388) and
 * the per-species particle lists (partlist.F90:89-113). */
int epb_create(const epb_config *cfg, const epb_species *species, epb_handle **out);
int epb_destroy(epb_handle *h);
const char *epb_last_error(const epb_handle *h);
const char *epb_version(void);
/* ABI self-description for the binding side (the Fortran shim checks it once at start-up):
 * out[0] = sizeof(epb_config), out[1] = sizeof(epb_species), out[2] = ng, out[3] = EPB_NFIELD */
int epb_abi_info(int32_t out[4]);
/* run every kernel on this cudaStream_t (default: a stream owned by the handle) */
int epb_set_stream(epb_handle *h, void *cuda_stream);
int epb_synchronize(epb_handle *h);

/* -- multi-GPU ----------------------------------------------------------------
 * replaces: the Cartesian communicator (mpi_routines.F90:179-275).  id is the
 * 128-byte ncclUniqueId produced by epb_nccl_unique_id on rank 0 and broadcast
 * by the host (MPI_BCAST on the Fortran side, torch.distributed in the harness). */
int epb_nccl_unique_id(void *id128);
int epb_set_comm(epb_handle *h, const void *id128);

/* -- state transfer ------------------------------------------------------------ */
int epb_upload_field(epb_handle *h, int field, const double *host);
int epb_download_field(epb_handle *h, int field, double *host);
/* the same as a dump that overlaps the following steps (io/diagnostics.F90 writes field dumps while
 * nothing else runs; here the array is snapshotted on the device and leaves over a second stream):
 * host must be page-locked and stay untouched until epb_wait_downloads returns */
int epb_download_field_async(epb_handle *h, int field, double *host);
int epb_wait_downloads(epb_handle *h);
int epb_upload_species(epb_handle *h, int ispecies, int64_t n, const double *packed);
int epb_download_species(epb_handle *h, int ispecies, int64_t n, double *packed);
/* append_partlist(species%attached_list, ...): particles the host creates in mid-run (run_injectors,
 * injectors.F90:150-330; the moving window's insert_particles, window.F90:182-320) join the species */
int epb_append_species(epb_handle *h, int ispecies, int64_t n, const double *packed);
int epb_species_count(epb_handle *h, int ispecies, int64_t *n);   /* attached_list%count */
/* device-side loader for the bench: npart_per_cell particles per cell, uniform
 * density, Maxwellian momenta (stands in for auto_load, helper.F90:95, whose
 * KISS stream is host-serial; parity runs upload the oracle's particles instead).
 * npart_per_cell < 0: |npart_per_cell| per cell ON AVERAGE, each particle's cell drawn at random (Poisson counts
 * per cell: the state a thermal plasma relaxes to; bench.py's "mixed_state" figure) */
int epb_load_uniform(epb_handle *h, int ispecies, int32_t npart_per_cell, double density,
                     const double temp_k[3], const double drift[3], uint64_t seed);
/* particles-per-cell as calc_ppc defines it (io/calc_df.F90:761-808); out(nx,ny,nz) int32 */
int epb_cell_counts(epb_handle *h, int ispecies, int32_t *out);
/* device pointers of the field arrays, for harness-side reductions (may be NULL-checked) */
int epb_field_device_ptr(epb_handle *h, int field, void **dptr);

/* -- laser / outflow boundary sources -------------------------------------------
 * side = c_bd_x_min .. c_bd_z_max - 1: 0 x_min, 1 x_max, 2 y_min, 3 y_max, 4 z_min, 5 z_max.
 * source1/source2 are the arrays of laser.f90:347-352 (x faces), :479-500 (y faces), epoch3d
 * laser.f90 (z faces) on the local plane of that face: the two transverse axes in axis order,
 * (0:n) each, lower axis fastest; evaluated by the host each step (the time profile is a deck
 * expression, laser.f90:159-176). */
int epb_set_laser_source(epb_handle *h, int side, const double *source1, const double *source2);

/* Thermal particle boundaries (bc_particle = EPB_BC_THERMAL; boundary.F90:1104-1148 and its copies for the other
 * faces): the wall temperature species%ext_temp_<side> (shared_data.F90:255-256) of boundary `side` (0 x_min .. 5
 * z_max) as (plane, 3) doubles -- the transverse axes in axis order with their ghost cells (1-ng:n+ng), lower axis
 * fastest, then the three momentum components; 1D: three numbers.  Must be set before the first push of a rank that
 * owns such a wall.  Re-emission draws from counter-based per-particle streams (the reference: the rank's serial
 * KISS stream), so parity with the reference is statistical there. */
int epb_set_boundary_temperature(epb_handle *h, int ispecies, int side, const double *temp);

/* -- the hot path ---------------------------------------------------------------- */
/* epoch2d.F90:144-162: setup_field_boundaries snapshots, setup_bc_lists + particle_bcs,
 * efield_bcs, bfield_final_bcs with dt/2 */
int epb_init_boundaries(epb_handle *h);
/* update_eb_fields_half (fields.f90:533-559) */
int epb_fields_half(epb_handle *h);
/* push_particles (particles.F90:28-650) incl. particle_bcs (boundary.F90:1029-1462) */
int epb_push(epb_handle *h);
/* current_finish (current_smooth.F90:29-45) */
int epb_current_finish(epb_handle *h);
/* update_eb_fields_final (fields.f90:563-582) */
int epb_fields_final(epb_handle *h);
/* counting sort of every species by cell (supersedes reorder_particles_to_grid,
 * split_particle.F90:29-77); also runs automatically every sort_interval pushes */
int epb_sort(epb_handle *h);
/* update_particle_count (partlist.F90:984-1003): global count of a species */
int epb_global_count(epb_handle *h, int ispecies, int64_t *n);

/* -- device-side diagnostics (feed output_routines without downloading the state) ----
 * calc_total_energy_sum (io/calc_df.F90:1321-1417): out[0] = 0.5*eps0*sum(E^2)*dV,
 * out[1] = 0.5/mu0*sum(B^2)*dV over the local interior; kinetic: sum w*(gamma-1)*m*c^2 */
int epb_field_energy(epb_handle *h, double out[2]);
/* the scalars the host looks at after every step, without stopping it: host (page-locked, 3 + n_species doubles)
 * receives [0] the E-field energy, [1] the B-field energy (as above, summed over the ranks like
 * calc_total_energy_sum's MPI_ALLREDUCE), [2 + is] the global count of species is (update_particle_count; exact,
 * an integer below 2^53), [2 + n_species] the device error word (non-zero: a capacity overflow lost particles).
 * Evaluated in stream order; *ticket identifies the request, epb_wait_scalars blocks until its numbers are in host.
 * At most four requests may be outstanding. */
int epb_step_scalars_async(epb_handle *h, double *host, int64_t *ticket);
int epb_wait_scalars(epb_handle *h, int64_t ticket);
int epb_kinetic_energy(epb_handle *h, int ispecies, double *out);
/* calc_number_density (io/calc_df.F90:689-757), calc_charge_density (:608-685), calc_mass_density (:35-110)
 * of species ispecies (-1: sum over all species, tracers left out) incl. calc_boundary (ghost-cell sums with
 * the particle boundary codes, across ranks) and the zero-gradient ghost fill: host receives the array at
 * full extent (1-ng:nx+ng, ...), like the data_array the reference's output routines pass */
enum { EPB_MOMENT_NUMBER_DENSITY = 0, EPB_MOMENT_CHARGE_DENSITY = 1, EPB_MOMENT_MASS_DENSITY = 2,
       /* calc_ekbar (io/calc_df.F90:116-221): mean kinetic energy per cell, J */
       EPB_MOMENT_EKBAR = 3,
       /* calc_temperature (:877-1128), K: all momentum components (dof 3), or direction x / y / z (dof 1) */
       EPB_MOMENT_TEMPERATURE = 4, EPB_MOMENT_TEMPERATURE_X = 5, EPB_MOMENT_TEMPERATURE_Y = 6, EPB_MOMENT_TEMPERATURE_Z = 7,
       /* calc_ekflux (:415-557), W/m^2-like flux of kinetic energy through -x, +x, -y, +y, -z, +z */
       EPB_MOMENT_EKFLUX_XM = 8, EPB_MOMENT_EKFLUX_XP = 9, EPB_MOMENT_EKFLUX_YM = 10, EPB_MOMENT_EKFLUX_YP = 11,
       EPB_MOMENT_EKFLUX_ZM = 12, EPB_MOMENT_EKFLUX_ZP = 13,
       /* calc_average_momentum (:1239-1317), direction x / y / z */
       EPB_MOMENT_AVERAGE_PX = 14, EPB_MOMENT_AVERAGE_PY = 15, EPB_MOMENT_AVERAGE_PZ = 16,
       /* calc_per_species_current (:1132-1235), direction x / y / z */
       EPB_MOMENT_JX = 17, EPB_MOMENT_JY = 18, EPB_MOMENT_JZ = 19,
       /* calc_average_weight (:811-873): nearest cell, no ghost-cell sums */
       EPB_MOMENT_AVERAGE_WEIGHT = 20,
       /* calc_poynt_flux (:561-604): (E x B) / mu0 at the cell centres, direction x / y / z; ispecies ignored,
        * ghost cells of the result are zero (the reference leaves them undefined) */
       EPB_MOMENT_POYNT_FLUX_X = 21, EPB_MOMENT_POYNT_FLUX_Y = 22, EPB_MOMENT_POYNT_FLUX_Z = 23 };
int epb_calc_moment(epb_handle *h, int kind, int ispecies, double *host);

/* -- load balancing (host half in the binding: calculate_breaks, balance.F90:1948-2091) ---------------
 * get_load_x / get_load_y / get_load_z (balance.F90:1766-1844, epoch3d :2247-2362): the load profile of one axis
 * over the GLOBAL cells, load(1-ng : n_global+ng) as int64: push_per_field (5) per particle of any species in
 * that cell column, summed over the ranks, plus one field column per interior cell */
int epb_load_profile(epb_handle *h, int axis, int64_t *load);

/* balance_workload's data movement (housekeeping/balance.F90:93-300): redistribute_domain / redistribute_fields
 * (:383-436, :436-1060; redistribute_field_2d :1282) and distribute_particles (:2156-2211).  Collective over the ranks.
 * The host has re-cut the slabs (calculate_breaks on epb_load_profile) and describes the old and the new tensor-product
 * decomposition -- cell_min / cell_max are cell_x_min(1:nprocx) ... of mpi_routines.F90:317-351, 1-based inclusive global
 * cells per processor coordinate; rank = (cz * nprocy + cy) * nprocx + cx -- together with the config of THIS rank in the
 * new decomposition.  The library builds the new device state, sends every field cell (interior, plus the ghost cells of
 * the physical domain edges, as redistribute_field_2d does; the other ghost cells are refilled by the halo exchange) and
 * every particle (get_particle_processor, balance.F90:2095-2151) to the rank that owns it now, hands the communicator
 * over and destroys the old handle.  *out replaces old_h, which must not be used again. */
typedef struct epb_decomp {
  int32_t nproc[3];
  const int32_t *cell_min[3];
  const int32_t *cell_max[3];
} epb_decomp;
int epb_redistribute(epb_handle *old_h, const epb_decomp *old_d, const epb_decomp *new_d, const epb_config *new_cfg,
                     const epb_species *new_species, epb_handle **out);

/* -- moving window (housekeeping/window.F90) -------------------------------------------------------------------
 * One cell of shift_window (:62-94) and the setup_bc_lists / particle_bcs that moving_window runs after it (:383-385).
 * The caller has moved its grid by one cell along x the way window.F90:73-86 does (x_grid_min = x_global(1) + dx,
 * xb_min, x_min, x_max accumulated; dx unchanged) and passes the config of THIS rank in the new window -- same
 * decomposition `d`, same extents, grid_min_local / min_local / max_local / gmin / gmax / min_outer / max_outer of
 * the new grid, the boundary conditions to use from now on (bc_*_after_move) -- and the new x_grid_min.  The library
 * builds the new device state: every field array one cell to the left with the ghost cells refilled from the
 * neighbours (shift_field + field_bc), the incoming cell of the x_max ranks fixed up (:126-143), particles left of
 * the new x_min removed on the x_min ranks (remove_particles :324-345), every other particle on the rank whose
 * [x_min_local, x_max_local) of the new grid holds it.  insert_particles (:182-320) consumes the host's random
 * numbers and deck expressions and stays the host's: hand the new plasma over with epb_append_species afterwards.
 * Collective.  *out replaces old_h.  Not for CPML runs or non-zero boundary snapshots (as epb_redistribute); laser
 * sources and wall temperatures are per-step / per-handle inputs and have to be set again. */
int epb_shift_window(epb_handle *old_h, const epb_decomp *d, const epb_config *new_cfg, const epb_species *new_species,
                     double x_grid_min, epb_handle **out);

/* -- binary collisions ---------------------------------------------------------------------------------------
 * particle_collisions (physics_packages/collisions.F90:86-214), called by PROGRAM pic after push_particles on the
 * steps where MODULO(step, coll_n_step) == coll_n_step - 1 (epoch2d.F90:219-236).  The per-cell lists the reference
 * builds with reorder_particles_to_grid are the device's own cell-resident layout, so nothing is re-sorted.
 * coll_pairs(n_species, n_species): the deck's user_factor for every species pair (row-major, upper triangle read;
 * <= 0: the pair does not collide).  coulomb_log > 0: fixed value; <= 0: coulomb_log_auto (calc_coulomb_log :1288).
 * use_nanbu: 1 Nanbu / Perez (the reference's default), 0 Sentoku-Kemp.  The random numbers come from counter-based
 * per-pair streams derived from seed (the reference's single KISS stream is inherently serial). */
typedef struct epb_collisions {
  int32_t n_species;
  int32_t coll_n_step;
  int32_t use_nanbu;
  int32_t reserved;
  double coulomb_log;
  uint64_t seed;
  const double *coll_pairs;
} epb_collisions;
int epb_collide(epb_handle *h, const epb_collisions *c);
/* test hook: the pair operator alone on explicit pairs and random numbers (tests/test_collisions.py) */
int epb_collide_pairs_test(int n, double *p1, double *p2, const double *w1, const double *w2, const double *ran,
                           const double *env, int *done);

/* -- instrumentation ---------------------------------------------------------------
 * kernel launch counter since creation (bench.py's gpu_launches), and CUDA-event
 * timing of the push/deposit kernel alone: average ms per launch since the last reset */
int64_t epb_launch_count(const epb_handle *h);
int epb_push_kernel_ms(epb_handle *h, double *avg_ms, int64_t *launches, int reset);

#ifdef __cplusplus
}
#endif
#endif /* EPOCH_B200_H */
