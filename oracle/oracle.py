"""ctypes wrapper around oracle/liboracle.so (the CPU restatement of EPOCH's hot path).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from epoch_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Config(C.Structure):
    _fields_ = [
        ("ndims", C.c_int), ("n_global", C.c_int * 3), ("nproc", C.c_int * 3),
        ("xmin", C.c_double * 3), ("xmax", C.c_double * 3), ("bc_field", C.c_int * 6),
        ("dt", C.c_double), ("n_species", C.c_int), ("seed", C.c_int),
        ("field_order", C.c_int), ("maxwell_solver", C.c_int),
        ("st_alpha", C.c_double * 3), ("st_beta", C.c_double * 6), ("st_gamma", C.c_double * 3),
        ("st_delta", C.c_double * 3),
        ("smooth_its", C.c_int), ("smooth_comp_its", C.c_int), ("smooth_nstrides", C.c_int),
        ("smooth_strides", C.c_int * 4), ("force_mixed", C.c_int), ("hc_push", C.c_int),
        ("cpml_thickness", C.c_int), ("cpml_kappa_max", C.c_double), ("cpml_a_max", C.c_double),
        ("cpml_sigma_max", C.c_double),
    ]


class _Species(C.Structure):
    _fields_ = [
        ("charge", C.c_double), ("mass", C.c_double), ("bc_particle", C.c_int * 6),
        ("npart_per_cell", C.c_double), ("density", C.c_double),
        ("box_lo", C.c_double * 3), ("box_hi", C.c_double * 3),
        ("temp", C.c_double * 3), ("drift", C.c_double * 3),
        ("zero_current", C.c_int), ("immobile", C.c_int),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "epoch_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(_Config), C.POINTER(_Species)]
        L.orc_field.restype = C.POINTER(C.c_double)
        L.orc_field.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_field_size.restype = C.c_int64
        L.orc_field_size.argtypes = [C.c_void_p, C.c_int]
        L.orc_species_count.restype = C.c_int64
        L.orc_species_count.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_dx.restype = C.c_double
        L.orc_dx.argtypes = [C.c_void_p, C.c_int]
        L.orc_nranks.argtypes = [C.c_void_p]
        for name in ("orc_destroy", "orc_auto_load", "orc_init", "orc_fields_half", "orc_fields_final",
                     "orc_push", "orc_push_only", "orc_particle_bcs", "orc_setup_bc_lists",
                     "orc_current_finish", "orc_efield_bcs"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.orc_bfield_bcs.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_particles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_set_particles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]
        L.orc_set_laser_source.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_cell_counts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_rank_info.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8
        L.orc_outer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_kiss.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_calc_moment.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_calc_moment.restype = None
        L.orc_set_boundary_temperature.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        L.orc_set_boundary_temperature.restype = None
        L.orc_shift_window.argtypes = [C.c_void_p, C.c_int]
        L.orc_window_inserted_count.restype = C.c_int64
        L.orc_window_inserted_count.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_window_inserted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_window_inserted.restype = None
        L.orc_window_clear_inserted.argtypes = [C.c_void_p]
        L.orc_window_clear_inserted.restype = None
        L.orc_window_geometry.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_window_geometry.restype = None
        L.orc_collide.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.orc_collide.restype = None
        L.orc_collide_pairs_test.argtypes = [C.c_int] + [C.c_void_p] * 7
        L.orc_coulomb_log.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.orc_coulomb_log.restype = None
        L.orc_collide_pairs_test.restype = None
        _LIB = L
    return _LIB


FIELD_NAMES = ("ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz")


class Oracle:
    """In-process multi-rank oracle driven with the same call sequence as the GPU backend."""

    def __init__(self, deck):
        self.deck = deck
        L = lib()
        cfg = _Config()
        cfg.ndims = deck.ndims
        for d in range(3):
            cfg.n_global[d] = deck.n[d] if d < deck.ndims else 1
            cfg.nproc[d] = deck.nproc[d] if d < deck.ndims else 1
            cfg.xmin[d] = deck.xmin[d] if d < deck.ndims else 0.0
            cfg.xmax[d] = deck.xmax[d] if d < deck.ndims else 1.0
        for i, b in enumerate(deck.bc_codes()):
            cfg.bc_field[i] = b
        cfg.dt = deck.dt()
        cfg.n_species = len(deck.species)
        cfg.seed = deck.seed
        cfg.field_order = int(getattr(deck, "field_order", 2))
        cfg.maxwell_solver = deck.maxwell_solver_code() if hasattr(deck, "maxwell_solver_code") else 0
        st = deck.stencil()
        for i, ax in enumerate("xyz"):
            cfg.st_alpha[i] = st["alpha" + ax]
            cfg.st_gamma[i] = st["gamma" + ax]
            cfg.st_delta[i] = st["delta" + ax]
        for i, k in enumerate(("betaxy", "betaxz", "betayx", "betayz", "betazx", "betazy")):
            cfg.st_beta[i] = st[k]
        cfg.force_mixed = int(getattr(deck, "force_mixed_bc", False))
        cfg.hc_push = int(getattr(deck, "hc_push", False))
        cfg.cpml_thickness = int(getattr(deck, "cpml_thickness", 6))
        cfg.cpml_kappa_max = float(getattr(deck, "cpml_kappa_max", 20.0))
        cfg.cpml_a_max = float(getattr(deck, "cpml_a_max", 0.15))
        cfg.cpml_sigma_max = float(getattr(deck, "cpml_sigma_max", 0.7))
        if getattr(deck, "smooth_currents", False):
            cfg.smooth_its = int(deck.smooth_iterations)
            cfg.smooth_comp_its = 1 if deck.smooth_compensation else 0
            cfg.smooth_nstrides = len(deck.smooth_strides)
            for i, v in enumerate(deck.smooth_strides):
                cfg.smooth_strides[i] = int(v)
        sp = (_Species * max(1, len(deck.species)))()
        for i, s in enumerate(deck.species):
            sp[i].charge, sp[i].mass = s.charge, s.mass
            for k, b in enumerate(deck.species_bc_codes(s)):
                sp[i].bc_particle[k] = b
            sp[i].npart_per_cell = s.npart_per_cell
            sp[i].density = s.density
            for d in range(3):
                sp[i].box_lo[d], sp[i].box_hi[d] = s.box_lo[d], s.box_hi[d]
                sp[i].temp[d], sp[i].drift[d] = s.temp[d], s.drift[d]
            sp[i].zero_current = int(s.zero_current)
            sp[i].immobile = int(s.immobile)
        self._h = C.c_void_p(L.orc_create(C.byref(cfg), sp))
        self.nranks = L.orc_nranks(self._h)
        self.nd = deck.ndims

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    # -- info ---------------------------------------------------------------
    def rank_info(self, rk):
        n = (C.c_int * 3)(); g = (C.c_int * 3)(); co = (C.c_int * 3)()
        ib = (C.c_int * 6)(); nb = (C.c_int * 27)()
        gm = (C.c_double * 3)(); mn = (C.c_double * 3)(); mx = (C.c_double * 3)()
        lib().orc_rank_info(self._h, rk, n, g, co, ib, nb, gm, mn, mx)
        return dict(n=list(n), gmin=list(g), coords=list(co), is_bnd=list(ib), neighbour=list(nb),
                    grid_min_local=list(gm), min_local=list(mn), max_local=list(mx))

    def outer(self):
        a = (C.c_double * 3)(); b = (C.c_double * 3)()
        lib().orc_outer(self._h, a, b)
        return list(a), list(b)

    def field_shape(self, rk):
        n = self.rank_info(rk)["n"]
        # numpy shape (z, y, x) with ghosts on active dims
        return tuple((n[d] + 2 * 5) if d < self.nd else 1 for d in (2, 1, 0))

    def field(self, rk, name):
        """Live numpy view (with ghost cells), indexed [k][j][i] (x fastest)."""
        which = FIELD_NAMES.index(name)
        sz = lib().orc_field_size(self._h, rk)
        p = lib().orc_field(self._h, rk, which)
        return np.ctypeslib.as_array(p, shape=(sz,)).reshape(self.field_shape(rk))

    def interior(self, rk, name):
        a = self.field(rk, name)
        sl = tuple(slice(5, -5) if a.shape[ax] > 1 else slice(None) for ax in range(3))
        return a[sl]

    def count(self, rk, isp):
        return lib().orc_species_count(self._h, rk, isp)

    def get_particles(self, rk, isp):
        n = self.count(rk, isp)
        out = np.empty((n, self.nd + 4), dtype=np.float64)
        if n:
            lib().orc_get_particles(self._h, rk, isp, out.ctypes.data)
        return out

    def set_particles(self, rk, isp, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        lib().orc_set_particles(self._h, rk, isp, arr.shape[0], arr.ctypes.data)

    MOMENTS = {"number_density": 0, "charge_density": 1, "mass_density": 2, "ekbar": 3, "temperature": 4,
               "temperature_x": 5, "temperature_y": 6, "temperature_z": 7,
               "ekflux_xm": 8, "ekflux_xp": 9, "ekflux_ym": 10, "ekflux_yp": 11, "ekflux_zm": 12, "ekflux_zp": 13,
               "average_px": 14, "average_py": 15, "average_pz": 16, "jx": 17, "jy": 18, "jz": 19, "average_weight": 20,
               "poynt_flux_x": 21, "poynt_flux_y": 22, "poynt_flux_z": 23}

    def moment(self, rk, kind, isp=-1):
        """calc_number_density / calc_charge_density / calc_mass_density (io/calc_df.F90) of species isp
        (-1: all species); returns a copy of the work array of rank rk, ghost cells included."""
        L = lib()
        L.orc_calc_moment(self._h, self.MOMENTS[kind], isp)
        sz = L.orc_field_size(self._h, rk)
        p = L.orc_field(self._h, rk, 9)
        return np.ctypeslib.as_array(p, shape=(sz,)).reshape(self.field_shape(rk)).copy()

    def cell_counts(self, rk, isp):
        n = self.rank_info(rk)["n"]
        out = np.zeros((n[2], n[1], n[0]), dtype=np.int32)
        lib().orc_cell_counts(self._h, rk, isp, out.ctypes.data)
        return out

    # -- backend interface (epoch_b200.deck.run) ----------------------------
    def set_laser_source(self, rk, side, s1, s2):
        s1 = np.ascontiguousarray(s1, dtype=np.float64)
        s2 = np.ascontiguousarray(s2, dtype=np.float64)
        lib().orc_set_laser_source(self._h, rk, side, s1.ctypes.data, s2.ctypes.data)

    def auto_load(self): lib().orc_auto_load(self._h)
    def init(self): lib().orc_init(self._h)
    def fields_half(self): lib().orc_fields_half(self._h)
    def fields_final(self): lib().orc_fields_final(self._h)
    def set_boundary_temperature(self, rk, isp, side, temp_k):
        """ext_temp_<side> of a thermal particle boundary on rank rk: (3,) or (3, plane) [K]"""
        n = self.rank_info(rk)["n"]
        plane = 1
        for d in range(self.deck.ndims):
            if d != side // 2:
                plane *= n[d] + 10
        t = np.asarray(temp_k, dtype=np.float64)
        if t.size == 3:
            t = np.repeat(t.reshape(3, 1), plane, axis=1)
        t = np.ascontiguousarray(t.reshape(3, plane))
        lib().orc_set_boundary_temperature(self._h, rk, isp, side, t.ctypes.data, t.size)

    def collide(self, coll_pairs, coulomb_log=0.0, use_nanbu=True, coll_n_step=1):
        """particle_collisions (physics_packages/collisions.F90:86-214) on every rank, with the rank's KISS stream"""
        n = len(self.deck.species)
        cp = np.ascontiguousarray(np.asarray(coll_pairs, dtype=np.float64).reshape(n, n))
        lib().orc_collide(self._h, int(coll_n_step), int(use_nanbu), float(coulomb_log), cp.ctypes.data)

    # -- moving window (housekeeping/window.F90) -------------------------------
    def shift_window(self, cells):
        """shift_window(cells) + setup_bc_lists + particle_bcs (window.F90:383-385).  The particles insert_particles
        created are kept per rank and species until window_clear_inserted (window_inserted)."""
        if lib().orc_shift_window(self._h, int(cells)) != 0:
            raise RuntimeError("the oracle's moving window is restated for non-periodic x without CPML")
        self._window_shifts = getattr(self, "_window_shifts", 0) + int(cells)
        # the driver moves the deck's grid (Deck.shift_window_geometry, once per cell, before this call): both
        # restatements of window.F90:73-86 must agree to the bit
        if getattr(self.deck, "window_shifts", 0) == self._window_shifts:
            g = self.window_geometry()
            assert (self.deck.window_grid_min, self.deck.window_xb_min, self.deck.xmin[0], self.deck.xmax[0]) == g, \
                "host-side window geometry differs from the oracle's"

    def window_inserted(self, rk, isp):
        n = lib().orc_window_inserted_count(self._h, rk, isp)
        out = np.empty((n, self.nd + 4), dtype=np.float64)
        if n:
            lib().orc_window_inserted(self._h, rk, isp, out.ctypes.data)
        return out

    def window_clear_inserted(self):
        lib().orc_window_clear_inserted(self._h)

    def window_geometry(self):
        g = (C.c_double * 4)()
        lib().orc_window_geometry(self._h, g)
        return tuple(g)

    def push(self): lib().orc_push(self._h)
    def push_only(self): lib().orc_push_only(self._h)
    def particle_bcs(self): lib().orc_particle_bcs(self._h)
    def current_finish(self): lib().orc_current_finish(self._h)


def kiss(seed, n):
    out = np.empty(n, dtype=np.float64)
    lib().orc_kiss(seed, n, out.ctypes.data)
    return out
