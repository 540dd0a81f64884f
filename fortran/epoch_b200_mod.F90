! epoch_b200_mod.F90 -- ISO_C_BINDING shim between EPOCH's Fortran host and libepoch_b200.so.
!
! SOURCE ONLY: the build image has no Fortran compiler (SURVEY.md section 0), so this file is
! delivered as the binding a maintainer adds to epoch{1,2,3}d/src/; everything below the C ABI
! is verified through the ctypes harness (epoch_b200/lib.py), which binds the same symbols
! with the same struct layout (checked at start-up through epb_abi_info).
!
! How it drops in (INTEGRATION.md has the step list):
!   * PROGRAM pic keeps its call order (epoch2d.F90:211,216,250,265).  The four routines it calls
!     get the one-line bodies at the bottom of this file (compile with -DEPOCH_B200):
!       update_eb_fields_half  -> epb_fields_half      (fields.f90:533)
!       push_particles         -> epb_push             (particles.F90:28, incl. particle_bcs)
!       current_finish         -> epb_current_finish   (housekeeping/current_smooth.F90:29)
!       update_eb_fields_final -> epb_fields_final     (fields.f90:563)
!   * b200_attach is called once after set_dt / before setup_bc_lists (epoch2d.F90:139-146): it
!     creates the device state from shared_data and uploads fields and particles.
!   * b200_download is called from output_routines when a dump is due (io/diagnostics.F90:206)
!     and before any host-side package that walks species_list(:)%attached_list.
!
! c_ndims = 2 is written out; the 1D/3D trees differ only in array ranks.

MODULE epoch_b200_mod

  USE, INTRINSIC :: ISO_C_BINDING
  USE shared_data
  USE partlist

  IMPLICIT NONE

  ! struct epb_config (include/epoch_b200.h)
  TYPE, BIND(C) :: epb_config
    INTEGER(C_INT32_T) :: ndims
    INTEGER(C_INT32_T) :: n(3)
    INTEGER(C_INT32_T) :: n_global(3)
    INTEGER(C_INT32_T) :: ng
    INTEGER(C_INT32_T) :: bc_field(6)
    INTEGER(C_INT32_T) :: is_boundary(6)
    INTEGER(C_INT32_T) :: neighbour(27)
    INTEGER(C_INT32_T) :: rank, nranks
    INTEGER(C_INT32_T) :: n_species
    INTEGER(C_INT32_T) :: strict_fp
    INTEGER(C_INT32_T) :: sort_interval
    INTEGER(C_INT32_T) :: field_order
    INTEGER(C_INT32_T) :: maxwell_solver
    INTEGER(C_INT32_T) :: smooth_its
    INTEGER(C_INT32_T) :: smooth_comp_its
    INTEGER(C_INT32_T) :: smooth_strides
    INTEGER(C_INT32_T) :: hc_push
    REAL(C_DOUBLE) :: dx(3)
    REAL(C_DOUBLE) :: dt
    REAL(C_DOUBLE) :: grid_min_local(3)
    REAL(C_DOUBLE) :: min_local(3)
    REAL(C_DOUBLE) :: max_local(3)
    REAL(C_DOUBLE) :: gmin(3), gmax(3)
    REAL(C_DOUBLE) :: min_outer(3)
    REAL(C_DOUBLE) :: max_outer(3)
    REAL(C_DOUBLE) :: stencil(15)
    REAL(C_DOUBLE) :: cpml_kappa_max, cpml_a_max, cpml_sigma_max
    INTEGER(C_INT32_T) :: cpml_thickness
    INTEGER(C_INT32_T) :: n_global_min(3)
  END TYPE epb_config

  ! struct epb_decomp (include/epoch_b200.h): cell_x_min(1:nprocx) ... of mpi_routines.F90:317-351 per axis
  TYPE, BIND(C) :: epb_decomp
    INTEGER(C_INT32_T) :: nproc(3)
    TYPE(C_PTR) :: cell_min(3)
    TYPE(C_PTR) :: cell_max(3)
  END TYPE epb_decomp

  ! struct epb_species
  TYPE, BIND(C) :: epb_species
    REAL(C_DOUBLE) :: charge
    REAL(C_DOUBLE) :: mass
    INTEGER(C_INT32_T) :: bc_particle(6)
    INTEGER(C_INT32_T) :: zero_current
    INTEGER(C_INT32_T) :: immobile
    INTEGER(C_INT64_T) :: capacity
  END TYPE epb_species

  INTEGER, PARAMETER :: epb_ex = 0, epb_ey = 1, epb_ez = 2, epb_bx = 3, epb_by = 4, &
      epb_bz = 5, epb_jx = 6, epb_jy = 7, epb_jz = 8

  INTERFACE
    FUNCTION epb_abi_info(info) BIND(C, NAME='epb_abi_info') RESULT(rc)
      IMPORT :: C_INT, C_INT32_T
      INTEGER(C_INT32_T) :: info(4)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_create(cfg, species, handle) BIND(C, NAME='epb_create') RESULT(rc)
      IMPORT :: C_INT, C_PTR, epb_config, epb_species
      TYPE(epb_config), INTENT(IN) :: cfg
      TYPE(epb_species), INTENT(IN) :: species(*)
      TYPE(C_PTR), INTENT(OUT) :: handle
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_destroy(handle) BIND(C, NAME='epb_destroy') RESULT(rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_last_error(handle) BIND(C, NAME='epb_last_error') RESULT(msg)
      IMPORT :: C_PTR
      TYPE(C_PTR), VALUE :: handle
      TYPE(C_PTR) :: msg
    END FUNCTION
    FUNCTION epb_nccl_unique_id(id) BIND(C, NAME='epb_nccl_unique_id') RESULT(rc)
      IMPORT :: C_INT, C_CHAR
      CHARACTER(KIND=C_CHAR) :: id(128)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_set_comm(handle, id) BIND(C, NAME='epb_set_comm') RESULT(rc)
      IMPORT :: C_INT, C_PTR, C_CHAR
      TYPE(C_PTR), VALUE :: handle
      CHARACTER(KIND=C_CHAR) :: id(128)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_upload_field(handle, field, host) BIND(C, NAME='epb_upload_field') RESULT(rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT), VALUE :: field
      TYPE(C_PTR), VALUE :: host
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_download_field(handle, field, host) BIND(C, NAME='epb_download_field') RESULT(rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT), VALUE :: field
      TYPE(C_PTR), VALUE :: host
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_upload_species(handle, ispecies, n, packed) &
        BIND(C, NAME='epb_upload_species') RESULT(rc)
      IMPORT :: C_INT, C_INT64_T, C_PTR, C_DOUBLE
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT), VALUE :: ispecies
      INTEGER(C_INT64_T), VALUE :: n
      REAL(C_DOUBLE), INTENT(IN) :: packed(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_shift_window(handle, d, cfg, species, x_grid_min_new, new_handle) &
        BIND(C, NAME='epb_shift_window') RESULT(rc)
      IMPORT :: C_INT, C_PTR, C_DOUBLE, epb_config, epb_species, epb_decomp
      TYPE(C_PTR), VALUE :: handle
      TYPE(epb_decomp), INTENT(IN) :: d
      TYPE(epb_config), INTENT(IN) :: cfg
      TYPE(epb_species), INTENT(IN) :: species(*)
      REAL(C_DOUBLE), VALUE :: x_grid_min_new
      TYPE(C_PTR), INTENT(OUT) :: new_handle
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_append_species(handle, ispecies, n, packed) &
        BIND(C, NAME='epb_append_species') RESULT(rc)
      IMPORT :: C_INT, C_INT64_T, C_PTR, C_DOUBLE
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT), VALUE :: ispecies
      INTEGER(C_INT64_T), VALUE :: n
      REAL(C_DOUBLE), INTENT(IN) :: packed(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_download_species(handle, ispecies, n, packed) &
        BIND(C, NAME='epb_download_species') RESULT(rc)
      IMPORT :: C_INT, C_INT64_T, C_PTR, C_DOUBLE
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT), VALUE :: ispecies
      INTEGER(C_INT64_T), VALUE :: n
      REAL(C_DOUBLE), INTENT(OUT) :: packed(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_species_count(handle, ispecies, n) BIND(C, NAME='epb_species_count') RESULT(rc)
      IMPORT :: C_INT, C_INT64_T, C_PTR
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT), VALUE :: ispecies
      INTEGER(C_INT64_T), INTENT(OUT) :: n
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_global_count(handle, ispecies, n) BIND(C, NAME='epb_global_count') RESULT(rc)
      IMPORT :: C_INT, C_INT64_T, C_PTR
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT), VALUE :: ispecies
      INTEGER(C_INT64_T), INTENT(OUT) :: n
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_set_laser_source(handle, side, source1, source2) &
        BIND(C, NAME='epb_set_laser_source') RESULT(rc)
      IMPORT :: C_INT, C_PTR, C_DOUBLE
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT), VALUE :: side
      REAL(C_DOUBLE), INTENT(IN) :: source1(*), source2(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_init_boundaries(handle) BIND(C, NAME='epb_init_boundaries') RESULT(rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_fields_half(handle) BIND(C, NAME='epb_fields_half') RESULT(rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_push(handle) BIND(C, NAME='epb_push') RESULT(rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_current_finish(handle) BIND(C, NAME='epb_current_finish') RESULT(rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION epb_fields_final(handle) BIND(C, NAME='epb_fields_final') RESULT(rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT) :: rc
    END FUNCTION
    ! load(1-ng : n_global+ng) of get_load_x/y/z (balance.F90:1766-1844); axis = 0, 1, 2
    FUNCTION epb_load_profile(handle, axis, load) BIND(C, NAME='epb_load_profile') RESULT(rc)
      IMPORT :: C_INT, C_PTR, C_INT64_T
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT), VALUE :: axis
      INTEGER(C_INT64_T), INTENT(OUT) :: load(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
    ! kind: 0 number density, 1 charge density, 2 mass density, 3 ekbar, 4 temperature, 5..7 temperature x/y/z,
    ! 8..13 ekflux -x,+x,-y,+y,-z,+z, 14..16 average px,py,pz, 17..19 species current jx,jy,jz, 20 average weight,
    ! 21..23 Poynting flux x,y,z;
    ! ispecies = -1: all species
    FUNCTION epb_calc_moment(handle, kind, ispecies, host) BIND(C, NAME='epb_calc_moment') RESULT(rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE :: handle
      INTEGER(C_INT), VALUE :: kind, ispecies
      TYPE(C_PTR), VALUE :: host
      INTEGER(C_INT) :: rc
    END FUNCTION
  END INTERFACE

  TYPE(C_PTR), SAVE :: b200 = C_NULL_PTR
  LOGICAL, SAVE :: b200_fields_on_host = .TRUE.
  ! device record: pos(1..ndims), p(1..3), weight
  INTEGER, PARAMETER :: b200_nv = c_ndims + 4
#ifdef PER_SPECIES_WEIGHT
  INTEGER, PARAMETER :: b200_has_weight = 0
#else
  INTEGER, PARAMETER :: b200_has_weight = 1
#endif

CONTAINS

  SUBROUTINE b200_check(rc)

    INTEGER(C_INT), INTENT(IN) :: rc

    ! the reference aborts through abort_code (utilities.f90:261-281)
    IF (rc /= 0) THEN
      IF (rank == 0) PRINT *, '*** ERROR *** epoch_b200 returned code ', rc
      errcode = c_err_generic_error
      CALL abort_code(errcode)
    END IF

  END SUBROUTINE b200_check



  ! Build the device state from shared_data (after set_dt, epoch2d.F90:139).
  SUBROUTINE b200_attach

    TYPE(epb_config) :: cfg
    TYPE(epb_species), ALLOCATABLE :: sp(:)
    CHARACTER(KIND=C_CHAR) :: id(128)
    INTEGER(C_INT32_T) :: info(4)
    INTEGER :: ierr

    CALL b200_check(epb_abi_info(info))
    IF (info(1) /= C_SIZEOF(cfg)) CALL b200_check(1_C_INT)
    ! The device layout is pos(1..ndims), p(1..3), weight: builds whose pack_particle carries anything else
    ! (-DPER_SPECIES_WEIGHT, -DPER_PARTICLE_CHARGE_MASS, -DPARTICLE_ID, ..., partlist.F90:43-85) are refused
    ! -DPER_SPECIES_WEIGHT builds carry no weight per particle (shared_data.F90:101): b200_pack / b200_unpack below
    ! write species%weight into the device's weight column and drop it again on the way back
    IF (nvar /= c_ndims + 3 + b200_has_weight) THEN
      IF (rank == 0) PRINT *, '*** ERROR *** epoch_b200: unsupported particle layout, nvar =', nvar
      CALL abort_code(c_err_generic_error)
    END IF
#if defined(PARTICLE_SHAPE_TOPHAT) || defined(PARTICLE_SHAPE_BSPLINE3)
    IF (rank == 0) PRINT *, '*** ERROR *** epoch_b200: only the default (triangle) particle shape is supported'
    CALL abort_code(c_err_generic_error)
#endif

    CALL b200_fill_config(cfg, sp)
    CALL b200_check(epb_create(cfg, sp, b200))
    DEALLOCATE(sp)

    IF (nproc > 1) THEN
      IF (rank == 0) CALL b200_check(epb_nccl_unique_id(id))
      CALL MPI_BCAST(id, 128, MPI_CHARACTER, 0, comm, ierr)
      CALL b200_check(epb_set_comm(b200, id))
    END IF

    CALL b200_upload
    ! setup_bc_lists + particle_bcs + efield_bcs + bfield_final_bcs(dt/2), epoch2d.F90:144-162
    CALL b200_push_laser_sources
    CALL b200_check(epb_init_boundaries(b200))

  END SUBROUTINE b200_attach



  ! The mirror of the shared_data globals the path reads, as they are NOW (b200_attach; again after the moving window
  ! or the balancer has changed the grid or the decomposition).
  SUBROUTINE b200_fill_config(cfg, sp)

    TYPE(epb_config), INTENT(OUT) :: cfg
    TYPE(epb_species), ALLOCATABLE, INTENT(OUT) :: sp(:)
    INTEGER :: ispecies, i, ix, iy

    cfg%ndims = c_ndims
    cfg%n = (/ nx, ny, 1 /)
    cfg%n_global = (/ nx_global, ny_global, 1 /)
    cfg%ng = ng
    cfg%bc_field = c_bc_periodic
    cfg%bc_field(1:2*c_ndims) = bc_field(1:2*c_ndims)
    cfg%is_boundary = 0
    IF (x_min_boundary) cfg%is_boundary(1) = 1
    IF (x_max_boundary) cfg%is_boundary(2) = 1
    IF (y_min_boundary) cfg%is_boundary(3) = 1
    IF (y_max_boundary) cfg%is_boundary(4) = 1
    cfg%neighbour = -1
    DO iy = -1, 1
      DO ix = -1, 1
        ! neighbour(ix,iy) at [(iz+1)*9 + (iy+1)*3 + (ix+1)], iz = 0; MPI_PROC_NULL -> -1
        i = 9 + (iy + 1) * 3 + (ix + 1) + 1
        IF (neighbour(ix,iy) /= MPI_PROC_NULL) cfg%neighbour(i) = neighbour(ix,iy)
      END DO
    END DO
    cfg%rank = rank
    cfg%nranks = nproc
    cfg%n_species = n_species
    cfg%strict_fp = 1
    cfg%sort_interval = 0   ! library default
    cfg%smooth_its = 0
    cfg%smooth_comp_its = 0
    cfg%smooth_strides = 0
#ifdef HC_PUSH
    cfg%hc_push = 1
#else
    cfg%hc_push = 0
#endif
    IF (smooth_currents) THEN
      cfg%smooth_its = smooth_its
      cfg%smooth_comp_its = smooth_comp_its
      IF (ALLOCATED(smooth_strides)) THEN
        DO i = 1, MIN(4, SIZE(smooth_strides))
          cfg%smooth_strides = IOR(cfg%smooth_strides, ISHFT(smooth_strides(i), 4 * (i - 1)))
        END DO
      END IF
    END IF
    cfg%field_order = field_order
    cfg%maxwell_solver = maxwell_solver   ! c_maxwell_solver_* (constants.F90); 0 = yee
    cfg%dx = (/ dx, dy, 1.0_num /)
    cfg%dt = dt
    cfg%grid_min_local = (/ x_grid_min_local, y_grid_min_local, 0.0_num /)
    cfg%min_local = (/ x_min_local, y_min_local, 0.0_num /)
    cfg%max_local = (/ x_max_local, y_max_local, 0.0_num /)
    cfg%gmin = (/ x_min, y_min, 0.0_num /)
    cfg%gmax = (/ x_max, y_max, 0.0_num /)
    cfg%min_outer = (/ x_min_outer, y_min_outer, 0.0_num /)
    cfg%max_outer = (/ x_max_outer, y_max_outer, 0.0_num /)
    cfg%stencil = 0.0_num   ! fields.f90 module variables (epoch2d has no z / gamma terms)
    cfg%stencil(1:2) = (/ alphax, alphay /)
    cfg%stencil(3) = 1.0_num
    cfg%stencil(4) = betaxy
    cfg%stencil(6) = betayx
    cfg%stencil(13:14) = (/ deltax, deltay /)
    ! CPML: cpml_thickness is already 0 when no field boundary is cpml_laser / cpml_outflow (mpi_routines.F90:285);
    ! the library restates set_cpml_helpers (boundary.F90:1479-1770) from these numbers
    cfg%cpml_thickness = cpml_thickness
    cfg%cpml_kappa_max = cpml_kappa_max
    cfg%cpml_a_max = cpml_a_max
    cfg%cpml_sigma_max = cpml_sigma_max
    cfg%n_global_min = (/ nx_global_min, ny_global_min, 1 /)

    ALLOCATE(sp(n_species))
    DO ispecies = 1, n_species
      sp(ispecies)%charge = species_list(ispecies)%charge
      sp(ispecies)%mass = species_list(ispecies)%mass
      sp(ispecies)%bc_particle = c_bc_periodic
      sp(ispecies)%bc_particle(1:2*c_ndims) = species_list(ispecies)%bc_particle(1:2*c_ndims)
      sp(ispecies)%zero_current = MERGE(1, 0, species_list(ispecies)%zero_current)
      sp(ispecies)%immobile = MERGE(1, 0, species_list(ispecies)%immobile)
      ! head-room for migration; the library reports EPB_ERR_CAPACITY if it is exceeded
      sp(ispecies)%capacity = species_list(ispecies)%attached_list%count * 3 / 2 + 65536
    END DO

  END SUBROUTINE b200_fill_config



  ! housekeeping/window.F90, one pass of shift_window's loop (:69-92).  insert_particles keeps building its
  ! append_list (it consumes the deck expressions and the random stream) but no longer appends it on the host; the
  ! grid update (:73-86, setup_grid_x) stays as it is; then this routine replaces remove_particles and shift_fields
  ! (:88-91) and the setup_bc_lists / particle_bcs of moving_window (:384-385), and the new plasma follows with
  ! b200_append.  species%count of the host lists is stale while the device owns the particles: size the capacities
  ! from update_particle_count's numbers in a long run.
  SUBROUTINE b200_shift_window

    TYPE(epb_config) :: cfg
    TYPE(epb_species), ALLOCATABLE :: sp(:)
    TYPE(epb_decomp) :: d
    TYPE(C_PTR) :: new_handle
    INTEGER(C_INT32_T), TARGET :: cxmin(nprocx), cxmax(nprocx), cymin(nprocy), cymax(nprocy)
    INTEGER(C_INT32_T), TARGET :: one(1)

    CALL b200_fill_config(cfg, sp)   ! x_grid_min_local, x_min_local, x_min ... of the window that has just moved
    cxmin = cell_x_min(1:nprocx)
    cxmax = cell_x_max(1:nprocx)
    cymin = cell_y_min(1:nprocy)
    cymax = cell_y_max(1:nprocy)
    one = 1
    d%nproc = (/ nprocx, nprocy, 1 /)
    d%cell_min = (/ C_LOC(cxmin), C_LOC(cymin), C_LOC(one) /)
    d%cell_max = (/ C_LOC(cxmax), C_LOC(cymax), C_LOC(one) /)
    CALL b200_check(epb_shift_window(b200, d, cfg, sp, x_grid_min, new_handle))
    b200 = new_handle   ! the old device state is gone; laser sources are handed over every step anyway
    DEALLOCATE(sp)

  END SUBROUTINE b200_shift_window



  ! Host arrays / particle lists -> device (also after any host package changed them).
  SUBROUTINE b200_upload

    REAL(num), ALLOCATABLE, TARGET :: buf(:)
    TYPE(particle), POINTER :: cur
    INTEGER(i8) :: npart, ipart
    INTEGER :: ispecies

    CALL b200_check(epb_upload_field(b200, epb_ex, C_LOC(ex)))
    CALL b200_check(epb_upload_field(b200, epb_ey, C_LOC(ey)))
    CALL b200_check(epb_upload_field(b200, epb_ez, C_LOC(ez)))
    CALL b200_check(epb_upload_field(b200, epb_bx, C_LOC(bx)))
    CALL b200_check(epb_upload_field(b200, epb_by, C_LOC(by)))
    CALL b200_check(epb_upload_field(b200, epb_bz, C_LOC(bz)))

    DO ispecies = 1, n_species
      npart = species_list(ispecies)%attached_list%count
      ALLOCATE(buf(MAX(npart * b200_nv, 1_i8)))
      cur => species_list(ispecies)%attached_list%head
      ipart = 0
      DO WHILE (ASSOCIATED(cur))
        CALL b200_pack(buf(ipart*b200_nv+1:(ipart+1)*b200_nv), cur, species_list(ispecies))
        ipart = ipart + 1
        cur => cur%next
      END DO
      CALL b200_check(epb_upload_species(b200, ispecies - 1, npart, buf))
      DEALLOCATE(buf)
    END DO

  END SUBROUTINE b200_upload



  ! One particle in the device's record layout: the wire layout of pack_particle (partlist.F90:414-486) for the default
  ! build; for -DPER_SPECIES_WEIGHT the species' weight fills the weight column.
  SUBROUTINE b200_pack(rec, cur, species)

    REAL(num), INTENT(OUT) :: rec(:)
    TYPE(particle), POINTER :: cur
    TYPE(particle_species), INTENT(IN) :: species

#ifdef PER_SPECIES_WEIGHT
    rec(1:c_ndims) = cur%part_pos(1:c_ndims)
    rec(c_ndims+1:c_ndims+3) = cur%part_p(1:3)
    rec(c_ndims+4) = species%weight
#else
    CALL pack_particle(rec, cur)
#endif

  END SUBROUTINE b200_pack



  SUBROUTINE b200_unpack(rec, cur)

    REAL(num), INTENT(IN) :: rec(:)
    TYPE(particle), POINTER :: cur

#ifdef PER_SPECIES_WEIGHT
    cur%part_pos(1:c_ndims) = rec(1:c_ndims)
    cur%part_p(1:3) = rec(c_ndims+1:c_ndims+3)
#else
    CALL unpack_particle(rec, cur)
#endif

  END SUBROUTINE b200_unpack



  ! Particles the host created in mid-run (run_injectors, injectors.F90:150-330; the moving window's insert_particles):
  ! call with the list that was appended to species_list(ispecies)%attached_list, right after append_partlist.
  SUBROUTINE b200_append(ispecies, list)

    INTEGER, INTENT(IN) :: ispecies
    TYPE(particle_list), INTENT(IN) :: list
    REAL(num), ALLOCATABLE, TARGET :: buf(:)
    TYPE(particle), POINTER :: cur
    INTEGER(i8) :: ipart

    IF (list%count <= 0) RETURN
    ALLOCATE(buf(list%count * b200_nv))
    cur => list%head
    ipart = 0
    DO WHILE (ASSOCIATED(cur) .AND. ipart < list%count)
      CALL b200_pack(buf(ipart*b200_nv+1:(ipart+1)*b200_nv), cur, species_list(ispecies))
      ipart = ipart + 1
      cur => cur%next
    END DO
    CALL b200_check(epb_append_species(b200, ispecies - 1, ipart, buf))
    DEALLOCATE(buf)

  END SUBROUTINE b200_append



  ! Device -> host arrays and lists (before output_routines or a host-side package).
  SUBROUTINE b200_download(with_particles)

    LOGICAL, INTENT(IN) :: with_particles
    REAL(num), ALLOCATABLE, TARGET :: buf(:)
    TYPE(particle), POINTER :: cur
    INTEGER(C_INT64_T) :: npart
    INTEGER(i8) :: ipart
    INTEGER :: ispecies

    CALL b200_check(epb_download_field(b200, epb_ex, C_LOC(ex)))
    CALL b200_check(epb_download_field(b200, epb_ey, C_LOC(ey)))
    CALL b200_check(epb_download_field(b200, epb_ez, C_LOC(ez)))
    CALL b200_check(epb_download_field(b200, epb_bx, C_LOC(bx)))
    CALL b200_check(epb_download_field(b200, epb_by, C_LOC(by)))
    CALL b200_check(epb_download_field(b200, epb_bz, C_LOC(bz)))
    CALL b200_check(epb_download_field(b200, epb_jx, C_LOC(jx)))
    CALL b200_check(epb_download_field(b200, epb_jy, C_LOC(jy)))
    CALL b200_check(epb_download_field(b200, epb_jz, C_LOC(jz)))
    IF (.NOT. with_particles) RETURN

    DO ispecies = 1, n_species
      CALL b200_check(epb_species_count(b200, ispecies - 1, npart))
      ALLOCATE(buf(MAX(npart * b200_nv, 1_i8)))
      CALL b200_check(epb_download_species(b200, ispecies - 1, npart, buf))
      CALL destroy_partlist(species_list(ispecies)%attached_list)
      CALL create_allocated_partlist(species_list(ispecies)%attached_list, npart)
      cur => species_list(ispecies)%attached_list%head
      ipart = 0
      DO WHILE (ASSOCIATED(cur))
        CALL b200_unpack(buf(ipart*b200_nv+1:(ipart+1)*b200_nv), cur)
        ipart = ipart + 1
        cur => cur%next
      END DO
      DEALLOCATE(buf)
    END DO

  END SUBROUTINE b200_download



  ! The deck expressions of the laser blocks are evaluated on the host exactly as
  ! outflow_bcs_x_min/x_max do (laser.f90:338-357, 418-437) and the two source lines
  ! are handed to the device boundary kernel.
  SUBROUTINE b200_push_laser_sources

    REAL(num), ALLOCATABLE :: source1(:), source2(:)
    TYPE(laser_block), POINTER :: current
    REAL(num) :: t_env, base
    INTEGER :: side, i, nt
    LOGICAL :: on_edge

    ! side = c_bd_x_min - 1 .. c_bd_y_max - 1 (epoch2d): the transverse extent is ny on the x
    ! faces and nx on the y faces (laser.f90:338-357, :479-500)
    DO side = 0, 3
      SELECT CASE (side)
        CASE (0); on_edge = x_min_boundary; nt = ny
        CASE (1); on_edge = x_max_boundary; nt = ny
        CASE (2); on_edge = y_min_boundary; nt = nx
        CASE (3); on_edge = y_max_boundary; nt = nx
      END SELECT
      IF (.NOT. on_edge) CYCLE
      IF (.NOT. (add_laser(side + 1) .OR. bc_field(side + 1) == c_bc_simple_outflow)) CYCLE
      ALLOCATE(source1(0:nt), source2(0:nt))
      source1 = 0.0_num
      source2 = 0.0_num
      IF (add_laser(side + 1)) THEN
        current => lasers
        DO WHILE (ASSOCIATED(current))
          IF (current%boundary == side + 1 &
              .AND. time >= current%t_start .AND. time <= current%t_end) THEN
            IF (current%use_phase_function) CALL laser_update_phase(current)
            IF (current%use_profile_function) CALL laser_update_profile(current)
            t_env = laser_time_profile(current) * current%amp
            DO i = 0, nt
              base = t_env * current%profile(i) &
                  * SIN(current%current_integral_phase + current%phase(i))
              source1(i) = source1(i) + base * COS(current%pol_angle)
              source2(i) = source2(i) + base * SIN(current%pol_angle)
            END DO
          END IF
          current => current%next
        END DO
      END IF
      CALL b200_check(epb_set_laser_source(b200, side, source1, source2))
      DEALLOCATE(source1, source2)
    END DO

  END SUBROUTINE b200_push_laser_sources

END MODULE epoch_b200_mod


#ifdef EPOCH_B200
! Replacement bodies (each goes into the module that owns the routine):

!   MODULE fields
!     SUBROUTINE update_eb_fields_half
!       CALL b200_check(epb_fields_half(b200))
!     END SUBROUTINE
!     SUBROUTINE update_eb_fields_final
!       CALL update_laser_omegas                 ! laser.f90:253-269 (host state)
!       CALL b200_push_laser_sources             ! sources at time = (k + 1/2) dt
!       CALL b200_check(epb_fields_final(b200))
!     END SUBROUTINE
!
!   MODULE particles
!     SUBROUTINE push_particles
!       CALL b200_check(epb_push(b200))          ! zero J, push, deposit, particle_bcs, migration
!     END SUBROUTINE
!
!   MODULE current_smooth
!     SUBROUTINE current_finish
!       CALL b200_check(epb_current_finish(b200))
!     END SUBROUTINE
!
!   MODULE calc_df                               ! io/calc_df.F90:608-757: no particle download for a density dump
!     SUBROUTINE calc_number_density(data_array, current_species, direction)
!       CALL b200_check(epb_calc_moment(b200, 0, MAX(current_species, 0) - 1, C_LOC(data_array)))
!     END SUBROUTINE                             ! calc_charge_density: kind 1, calc_mass_density: kind 2
!
!   MODULE partlist
!     SUBROUTINE update_particle_count           ! partlist.F90:984-1003
!       DO ispecies = 1, n_species
!         CALL b200_check(epb_global_count(b200, ispecies - 1, species_list(ispecies)%count))
!       END DO
!     END SUBROUTINE
#endif
