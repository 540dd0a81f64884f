"""Host-side mirror of the hot-path procedures PROGRAM pic calls, bound to the CUDA library.

`Simulation` owns one epb handle = one GPU = one decomposition rank.  Method names follow
the reference routines: fields_half (update_eb_fields_half, fields.f90:533), push
(push_particles, particles.F90:28), current_finish (current_smooth.F90:29), fields_final
(update_eb_fields_final, fields.f90:563).  Any error raises; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import lib as _lib
from .deck import NG, Deck


class EpbError(RuntimeError):
    pass


def _neighbour_table(deck: Deck, rank: int, periods):
    """mpi_routines.F90:256-273"""
    nd = deck.ndims
    np_ = [max(1, deck.nproc[d]) if d < nd else 1 for d in range(3)]
    co = deck.rank_coords(rank)
    out = [-1] * 27
    for iz in (-1, 0, 1):
        for iy in (-1, 0, 1):
            for ix in (-1, 0, 1):
                t = [co[0] + ix, co[1] + iy, co[2] + iz]
                ok = True
                for d in range(3):
                    if d >= nd:
                        if t[d] != 0:
                            ok = False
                        continue
                    if t[d] < 0 or t[d] >= np_[d]:
                        if not periods[d]:
                            ok = False
                        else:
                            t[d] %= np_[d]
                if ok:
                    out[(iz + 1) * 9 + (iy + 1) * 3 + (ix + 1)] = (t[2] * np_[1] + t[1]) * np_[0] + t[0]
    return out


def rank_geometry(deck: Deck, rank: int):
    """Per-rank grid scalars, operation-for-operation as utilities.f90:343-421."""
    nd = deck.ndims
    n, g = deck.local_extent(rank)
    co = deck.rank_coords(rank)
    geo = dict(n=n, gmin=g, coords=co, is_bnd=[0] * 6, grid_min_local=[0.0] * 3, min_local=[0.0] * 3,
               max_local=[0.0] * 3, min_outer=[0.0] * 3, max_outer=[0.0] * 3)
    png = 3
    t = deck.cpml_t()
    for d in range(nd):
        npd = max(1, deck.nproc[d])
        mins, maxs = deck.cell_ranges(d)
        dx = deck.dx(d)
        geo["is_bnd"][2 * d] = int(co[d] == 0)
        geo["is_bnd"][2 * d + 1] = int(co[d] == npd - 1)
        geo["grid_min_local"][d] = float(deck.x_global(d, mins[co[d]]))
        hdx = 0.5 * dx
        geo["min_local"][d] = geo["grid_min_local"][d] - hdx
        geo["max_local"][d] = float(deck.x_global(d, maxs[co[d]] + 1)) - hdx
        # utilities.f90:364-365: a CPML layer is not part of the particle domain (offsets of boundary.F90:1528-1543,
        # :1590-1605)
        off_min, off_max = cpml_offsets(deck, d, mins[co[d]], maxs[co[d]])
        geo["min_local"][d] = geo["min_local"][d] + off_min * dx
        geo["max_local"][d] = geo["max_local"][d] - off_max * dx
        shift = float((1 + png + t) // 2)
        geo["min_outer"][d] = deck.xmin[d] - shift * dx
        geo["max_outer"][d] = deck.xmax[d] + shift * dx
    return geo


def cpml_offsets(deck: Deck, d: int, gmin: int, gmax: int):
    """cpml_x_min_offset / cpml_x_max_offset of a rank that owns global cells gmin..gmax of axis d."""
    t = deck.cpml_t()
    ng_ = deck.ncells(d)
    off_min = off_max = 0
    if t and deck.bc_field[2 * d] in ("cpml_laser", "cpml_outflow") and gmin <= t:
        off_min = t - gmin + 1 if gmax >= t else t
    if t and deck.bc_field[2 * d + 1] in ("cpml_laser", "cpml_outflow") and gmax >= ng_ - t + 1:
        off_max = t - ng_ + gmax if gmin <= ng_ - t + 1 else t
    return off_min, off_max


def build_config(deck: Deck, rank: int, strict_fp: bool, sort_interval: int, capacity_factor: float,
                 min_capacity: int, capacities: Optional[Sequence[int]] = None):
    """(epb_config, epb_species[], geometry) of one rank: the mirror of the shared_data globals b200_attach copies
    (fortran/epoch_b200_mod.F90), from the deck and its decomposition (incl. slabs re-cut by the balancer)."""
    nd = deck.ndims
    geo = rank_geometry(deck, rank)
    bcf = deck.bc_codes()
    # setup_boundaries normalisation (boundary.F90:44-57)
    for i in range(6):
        if bcf[i] in (2, 9):
            bcf[i] = 8
        if bcf[i] == 5:
            bcf[i] = 4
    periods = []
    for d in range(3):
        per = d < nd and bcf[2 * d] == 1
        for s in deck.species:
            if d < nd and deck.species_bc_codes(s)[2 * d] == 1:
                per = True
        periods.append(per)
    cfg = _lib.Config()
    cfg.ndims = nd
    for d in range(3):
        cfg.n[d] = geo["n"][d]
        cfg.n_global[d] = deck.ncells(d) if d < nd else 1
        cfg.n_global_min[d] = geo["gmin"][d]
        cfg.dx[d] = deck.dx(d) if d < nd else 1.0
        cfg.grid_min_local[d] = geo["grid_min_local"][d]
        cfg.min_local[d] = geo["min_local"][d]
        cfg.max_local[d] = geo["max_local"][d]
        cfg.gmin[d] = deck.xmin[d] if d < nd else 0.0
        cfg.gmax[d] = deck.xmax[d] if d < nd else 0.0
        cfg.min_outer[d] = geo["min_outer"][d]
        cfg.max_outer[d] = geo["max_outer"][d]
    cfg.ng = NG
    for i in range(6):
        cfg.bc_field[i] = bcf[i]
        cfg.is_boundary[i] = geo["is_bnd"][i]
    for i, v in enumerate(_neighbour_table(deck, rank, periods)):
        cfg.neighbour[i] = v
    cfg.rank = rank
    cfg.nranks = deck.nranks()
    cfg.n_species = len(deck.species)
    cfg.strict_fp = int(strict_fp)
    cfg.sort_interval = sort_interval
    cfg.dt = deck.dt()
    cfg.field_order = int(deck.field_order)
    cfg.maxwell_solver = deck.maxwell_solver_code()
    cfg.hc_push = int(getattr(deck, "hc_push", False))
    cfg.cpml_thickness = deck.cpml_t()
    cfg.cpml_kappa_max, cfg.cpml_a_max, cfg.cpml_sigma_max = deck.cpml_kappa_max, deck.cpml_a_max, deck.cpml_sigma_max
    if deck.smooth_currents:
        cfg.smooth_its = int(deck.smooth_iterations)
        cfg.smooth_comp_its = 1 if deck.smooth_compensation else 0
        cfg.smooth_strides = sum(int(v) << (4 * i) for i, v in enumerate(deck.smooth_strides))
    st = deck.stencil()
    for i, k in enumerate(deck.STENCIL_KEYS):
        cfg.stencil[i] = st[k]
    ncell = geo["n"][0] * geo["n"][1] * geo["n"][2]
    sp = (_lib.SpeciesCfg * max(1, len(deck.species)))()
    for i, s in enumerate(deck.species):
        sp[i].charge, sp[i].mass = s.charge, s.mass
        for k, b in enumerate(deck.species_bc_codes(s)):
            # setup_particle_boundary (boundary.F90:108-122)
            if b in (2, 10):
                b = 9
            if b in (3, 4, 12, 13):
                b = 5
            sp[i].bc_particle[k] = b
        sp[i].zero_current = int(s.zero_current)
        sp[i].immobile = int(s.immobile)
        sp[i].capacity = max(min_capacity, int(capacity_factor * s.npart_per_cell * ncell))
        if capacities is not None:
            sp[i].capacity = max(min_capacity, int(capacities[i]))
    return cfg, sp, geo


def _decomp(deck: Deck):
    """struct epb_decomp of a deck's decomposition (keeps the int arrays alive on the returned object)."""
    d = _lib.Decomp()
    keep = []
    for a in range(3):
        d.nproc[a] = max(1, deck.nproc[a]) if a < deck.ndims else 1
        mins, maxs = deck.cell_ranges(a) if a < deck.ndims else ([1], [1])
        amin, amax = (C.c_int32 * len(mins))(*mins), (C.c_int32 * len(maxs))(*maxs)
        keep += [amin, amax]
        d.cell_min[a] = C.cast(amin, C.POINTER(C.c_int32))
        d.cell_max[a] = C.cast(amax, C.POINTER(C.c_int32))
    d._keep = keep
    return d


class Simulation:
    def __init__(self, deck: Deck, rank: int = 0, strict_fp: bool = True, sort_interval: int = 1,
                 capacity_factor: float = 1.5, min_capacity: int = 4096, stream: Optional[int] = None):
        self.deck = deck
        self.rank = rank
        self.L = _lib.load()
        nd = deck.ndims
        self._build_args = dict(strict_fp=strict_fp, sort_interval=sort_interval, capacity_factor=capacity_factor,
                                min_capacity=min_capacity)
        cfg, sp, geo = build_config(deck, rank, **self._build_args)
        self.geo = geo
        self._h = C.c_void_p()
        rc = self.L.epb_create(C.byref(cfg), sp, C.byref(self._h))
        if rc != 0:
            raise EpbError(f"epb_create failed with code {rc} (see stderr)")
        self.cfg = cfg
        self.species_cfg = sp
        self.nd = nd
        self.shape = tuple((geo["n"][d] + 2 * NG) if d < nd else 1 for d in (2, 1, 0))
        self._stream = stream
        if stream is not None:
            self._chk(self.L.epb_set_stream(self._h, C.c_void_p(stream)))

    def rebalance(self, cuts: dict, capacities: Optional[Sequence[int]] = None):
        """balance_workload's data movement (balance.F90:93-300) for slabs the host has re-cut: `cuts` maps an axis
        to (mins, maxs) as calculate_breaks returns them.  Collective: every rank calls it with the same cuts.
        Fields and particles move to their new owners on the device (epb_redistribute); this object then describes
        the rank's new sub-domain."""
        import copy
        new_deck = copy.copy(self.deck)
        new_cuts = dict(self.deck.cuts or {})
        new_cuts.update({int(a): (list(v[0]), list(v[1])) for a, v in cuts.items()})
        new_deck.cuts = new_cuts
        cfg, sp, geo = build_config(new_deck, self.rank, capacities=capacities, **self._build_args)
        od, nd_ = _decomp(self.deck), _decomp(new_deck)
        out = C.c_void_p()
        rc = self.L.epb_redistribute(self._h, C.byref(od), C.byref(nd_), C.byref(cfg), sp, C.byref(out))
        if rc != 0:
            self._chk(rc)
        self._h = out
        self.deck, self.cfg, self.species_cfg, self.geo = new_deck, cfg, sp, geo
        self.shape = tuple((geo["n"][d] + 2 * NG) if d < self.nd else 1 for d in (2, 1, 0))
        if self._stream is not None:
            self._chk(self.L.epb_set_stream(self._h, C.c_void_p(self._stream)))

    def shift_window(self, cells: int = 1, inserted: Optional[Sequence[np.ndarray]] = None,
                     capacities: Optional[Sequence[int]] = None):
        """One cell of the moving window (window.F90:62-94, :383-385) on the device: epb_shift_window, then the new
        plasma.  The driver has already moved the deck's grid (Deck.shift_window_geometry).  `inserted`: per species
        the particles insert_particles created on this rank ((n, ndims + 4), pack_particle order; None or empty on
        ranks that are not on the x_max edge) -- drawing them is the host's job (deck expressions, KISS stream)."""
        if int(cells) != 1:
            raise EpbError("Simulation.shift_window moves one cell per call (dt * window_v_x / dx < 1 under the CFL limit)")
        cfg, sp, geo = build_config(self.deck, self.rank, capacities=capacities, **self._build_args)
        d = _decomp(self.deck)
        out = C.c_void_p()
        rc = self.L.epb_shift_window(self._h, C.byref(d), C.byref(cfg), sp, C.c_double(self.deck.grid_min(0)), C.byref(out))
        if rc != 0:
            self._chk(rc)
        self._h = out
        self.cfg, self.species_cfg, self.geo = cfg, sp, geo
        if self._stream is not None:
            self._chk(self.L.epb_set_stream(self._h, C.c_void_p(self._stream)))
        if inserted is not None:
            for isp, arr in enumerate(inserted):
                if arr is not None and len(arr):
                    self.append_species(isp, arr)

    # ------------------------------------------------------------------
    def _chk(self, rc):
        if rc != 0:
            msg = self.L.epb_last_error(self._h)
            raise EpbError(f"epoch_b200 error {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_h", None):
            self.L.epb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state transfer ---------------------------------------------------
    def upload_field(self, name: str, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(self.shape)
        self._chk(self.L.epb_upload_field(self._h, _lib.FIELD_NAMES.index(name), a.ctypes.data))

    def download_field(self, name: str):
        a = np.empty(self.shape, dtype=np.float64)
        self._chk(self.L.epb_download_field(self._h, _lib.FIELD_NAMES.index(name), a.ctypes.data))
        return a

    def interior(self, name: str):
        a = self.download_field(name)
        sl = tuple(slice(NG, -NG) if a.shape[ax] > 1 else slice(None) for ax in range(3))
        return a[sl]

    def upload_species(self, isp: int, packed):
        p = np.ascontiguousarray(packed, dtype=np.float64)
        self._chk(self.L.epb_upload_species(self._h, isp, p.shape[0], p.ctypes.data))

    def append_species(self, isp: int, packed):
        """Particles the host creates in mid-run (injectors, the moving window's insertions) join the species."""
        p = np.ascontiguousarray(packed, dtype=np.float64)
        self._chk(self.L.epb_append_species(self._h, isp, p.shape[0], p.ctypes.data))

    def count(self, isp: int) -> int:
        n = C.c_int64()
        self._chk(self.L.epb_species_count(self._h, isp, C.byref(n)))
        return n.value

    def global_count(self, isp: int) -> int:
        n = C.c_int64()
        self._chk(self.L.epb_global_count(self._h, isp, C.byref(n)))
        return n.value

    def download_species(self, isp: int):
        n = self.count(isp)
        out = np.empty((n, self.nd + 4), dtype=np.float64)
        if n:
            self._chk(self.L.epb_download_species(self._h, isp, n, out.ctypes.data))
        return out

    def cell_counts(self, isp: int):
        n = self.geo["n"]
        out = np.zeros((n[2], n[1], n[0]), dtype=np.int32)
        self._chk(self.L.epb_cell_counts(self._h, isp, out.ctypes.data))
        return out

    def load_uniform(self, isp: int, seed: int = 12345, mixed: bool = False):
        """Device-side loader.  mixed: every particle's cell is drawn at random (Poisson counts per cell, the state
        a thermal plasma relaxes to) instead of exactly npart_per_cell per cell."""
        s = self.deck.species[isp]
        t = (C.c_double * 3)(*s.temp)
        d = (C.c_double * 3)(*s.drift)
        ppc = int(s.npart_per_cell)
        self._chk(self.L.epb_load_uniform(self._h, isp, -ppc if mixed else ppc, s.density, t, d, seed))

    def set_comm(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._chk(self.L.epb_set_comm(self._h, buf))

    @staticmethod
    def nccl_unique_id() -> bytes:
        L = _lib.load()
        buf = C.create_string_buffer(128)
        if L.epb_nccl_unique_id(buf) != 0:
            raise EpbError("ncclGetUniqueId failed")
        return buf.raw

    # -- backend interface used by deck.run --------------------------------
    def set_laser_source(self, local_rank: int, side: int, s1, s2):
        s1 = np.ascontiguousarray(s1, dtype=np.float64)
        s2 = np.ascontiguousarray(s2, dtype=np.float64)
        self._chk(self.L.epb_set_laser_source(self._h, side, s1.ctypes.data, s2.ctypes.data))

    def init(self): self._chk(self.L.epb_init_boundaries(self._h))
    def fields_half(self): self._chk(self.L.epb_fields_half(self._h))
    def push(self): self._chk(self.L.epb_push(self._h))
    def current_finish(self): self._chk(self.L.epb_current_finish(self._h))
    def fields_final(self): self._chk(self.L.epb_fields_final(self._h))
    def sort(self): self._chk(self.L.epb_sort(self._h))
    def synchronize(self): self._chk(self.L.epb_synchronize(self._h))

    def set_boundary_temperature(self, isp: int, side: int, temp_k):
        """species%ext_temp_<side> of a thermal particle boundary: a (3,) temperature [K] applied to the whole face, or
        the full (3, plane) array (transverse axes with ghost cells, lower axis fastest)."""
        plane = 1
        for d in range(self.nd):
            if d != side // 2:
                plane *= self.geo["n"][d] + 2 * NG
        t = np.asarray(temp_k, dtype=np.float64)
        if t.size == 3:
            t = np.repeat(t.reshape(3, 1), plane, axis=1)
        t = np.ascontiguousarray(t.reshape(3, plane))
        self._chk(self.L.epb_set_boundary_temperature(self._h, isp, side, t.ctypes.data))

    def collide(self, coll_pairs, coulomb_log: float = 0.0, use_nanbu: bool = True, coll_n_step: int = 1,
                seed: int = 7842432):
        """particle_collisions (physics_packages/collisions.F90:86-214): coll_pairs[i][j] = user_factor of the species
        pair (<= 0: no collisions); coulomb_log <= 0 means coulomb_log = auto.  PROGRAM pic calls it after
        push_particles on collision steps (epoch2d.F90:219-236)."""
        n = len(self.deck.species)
        cp = np.ascontiguousarray(np.asarray(coll_pairs, dtype=np.float64).reshape(n, n))
        c = _lib.Collisions()
        c.n_species, c.coll_n_step, c.use_nanbu = n, int(coll_n_step), int(use_nanbu)
        c.coulomb_log, c.seed = float(coulomb_log), int(seed)
        c.coll_pairs = cp.ctypes.data_as(C.POINTER(C.c_double))
        self._chk(self.L.epb_collide(self._h, C.byref(c)))

    def step(self):
        """One pass of the hot path with no boundary sources (uniform-plasma benchmark step)."""
        self.fields_half()
        self.push()
        self.current_finish()
        self.fields_final()

    def field_energy(self):
        """(electric, magnetic) field energy of the local interior, calc_df.F90:1321-1417"""
        out = (C.c_double * 2)()
        self._chk(self.L.epb_field_energy(self._h, out))
        return out[0], out[1]

    def kinetic_energy(self, isp: int) -> float:
        out = C.c_double()
        self._chk(self.L.epb_kinetic_energy(self._h, isp, C.byref(out)))
        return out.value

    def load_profile(self, axis: int):
        """get_load_x/y/z (balance.F90:1766-1844) over the global cells of `axis`, summed over the ranks; feed it to
        deck.calculate_breaks."""
        out = np.zeros(self.deck.ncells(axis) + 2 * NG, dtype=np.int64)
        self._chk(self.L.epb_load_profile(self._h, axis, out.ctypes.data))
        return out

    MOMENTS = {"number_density": 0, "charge_density": 1, "mass_density": 2, "ekbar": 3, "temperature": 4,
               "temperature_x": 5, "temperature_y": 6, "temperature_z": 7,
               "ekflux_xm": 8, "ekflux_xp": 9, "ekflux_ym": 10, "ekflux_yp": 11, "ekflux_zm": 12, "ekflux_zp": 13,
               "average_px": 14, "average_py": 15, "average_pz": 16, "jx": 17, "jy": 18, "jz": 19, "average_weight": 20,
               "poynt_flux_x": 21, "poynt_flux_y": 22, "poynt_flux_z": 23}

    def moment(self, kind: str, isp: int = -1):
        """calc_number_density / calc_charge_density / calc_mass_density (io/calc_df.F90) computed on the device;
        isp = -1 sums all species.  Returns the array with ghost cells, like download_field."""
        a = np.empty(self.shape, dtype=np.float64)
        self._chk(self.L.epb_calc_moment(self._h, self.MOMENTS[kind], isp, a.ctypes.data))
        return a

    def download_field_into(self, name: str, host_ptr: int):
        """D2H of one field array (with ghosts) into caller-owned (e.g. pinned) host memory."""
        self._chk(self.L.epb_download_field(self._h, _lib.FIELD_NAMES.index(name), C.c_void_p(host_ptr)))

    def download_field_async(self, name: str, host_ptr: int):
        """Field dump that overlaps the following steps; host_ptr must be page-locked and must not be read
        before wait_downloads() returns."""
        self._chk(self.L.epb_download_field_async(self._h, _lib.FIELD_NAMES.index(name), C.c_void_p(host_ptr)))

    def wait_downloads(self):
        self._chk(self.L.epb_wait_downloads(self._h))

    def write_replay_state(self, path: str, nsteps: int, fields: dict, particles: Sequence):
        """State file for host/replay.cpp (the compiled host that replays EPOCH's call sequence through the C ABI):
        this rank's epb_config / epb_species structs as raw bytes, the step count, ex..bz and the packed particles."""
        with open(path, "wb") as f:
            f.write(b"EPBRPLY1")
            f.write(np.array([C.sizeof(self.cfg), C.sizeof(_lib.SpeciesCfg), len(self.deck.species), nsteps],
                             dtype=np.int32).tobytes())
            f.write(bytes(self.cfg))
            for i in range(len(self.deck.species)):
                f.write(bytes(self.species_cfg[i]))
            for name in _lib.FIELD_NAMES[:6]:
                a = fields.get(name)
                a = np.zeros(self.shape) if a is None else np.ascontiguousarray(a, dtype=np.float64).reshape(self.shape)
                f.write(a.tobytes())
            for p in particles:
                p = np.ascontiguousarray(p, dtype=np.float64)
                f.write(np.array([p.shape[0]], dtype=np.int64).tobytes())
                f.write(p.tobytes())

    def read_replay_result(self, path: str):
        """(fields dict incl. jx..jz, [packed particles per species]) written by host/replay.cpp"""
        nvar = self.nd + 4
        with open(path, "rb") as f:
            assert f.read(8) == b"EPBRSLT1"
            n = int(np.prod(self.shape))
            fields = {name: np.frombuffer(f.read(8 * n), dtype=np.float64).reshape(self.shape)
                      for name in _lib.FIELD_NAMES}
            parts = []
            for _ in self.deck.species:
                k = int(np.frombuffer(f.read(8), dtype=np.int64)[0])
                parts.append(np.frombuffer(f.read(8 * k * nvar), dtype=np.float64).reshape(k, nvar))
        return fields, parts

    def launch_count(self) -> int:
        return int(self.L.epb_launch_count(self._h))

    def step_scalars_async(self, host_ptr: int) -> int:
        """Field energies and global particle counts of the state the stream has reached, copied to the page-locked
        array at host_ptr (3 + n_species doubles) without stopping the host; returns the ticket for wait_scalars."""
        t = C.c_int64()
        self._chk(self.L.epb_step_scalars_async(self._h, host_ptr, C.byref(t)))
        return t.value

    def wait_scalars(self, ticket: int):
        self._chk(self.L.epb_wait_scalars(self._h, ticket))

    def push_kernel_ms(self, reset: int = 0):
        ms = C.c_double(); n = C.c_int64()
        self._chk(self.L.epb_push_kernel_ms(self._h, C.byref(ms), C.byref(n), reset))
        return ms.value, n.value
