"""Host-side description of a run and the PROGRAM pic time loop.

This mirrors, on the host, the pieces of EPOCH that stay on the Fortran side of
the drop-in boundary: grid/timestep set-up (housekeeping/setup.F90:162-204,
:633-711), domain decomposition arithmetic (housekeeping/mpi_routines.F90:
317-351), laser source evaluation (laser.f90:253-269,343-352) and the main loop
call order (epoch2d.F90:144-268).  It is not a deck parser: the `.deck` surface
stays EPOCH's own; tests and the bench construct `Deck` objects directly.

A backend (the CUDA library behind `epoch_b200.pic.Simulation`, or the CPU
oracle in tests) only has to provide: init(), fields_half(), push(),
current_finish(), fields_final(), set_laser_source(local_rank, side, s1, s2).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

# constants.F90:192-199
pi = 3.141592653589793238462643383279503
q0 = 1.602176565e-19
m0 = 9.10938291e-31
c = 2.99792458e8
kb = 1.3806488e-23
epsilon0 = 8.854187817620389850536563031710750e-12
ev = q0
micron = 1e-6
femto = 1e-15

# constants.F90:75-90
BC = {
    "periodic": 1, "other": 2, "simple_laser": 3, "simple_outflow": 4, "open": 5,
    "zero_gradient": 7, "clamp": 8, "reflect": 9, "conduct": 10, "thermal": 11,
    "cpml_laser": 12, "cpml_outflow": 13,
}

NG = 5  # triangle shape: png = 3, ng = png + 2 (constants.F90:549-559)


def gauss(x, x0, w):
    """parser/evaluator_blocks.F90:998-1001"""
    return np.exp(-(((x - x0) / w) ** 2))


@dataclass
class Species:
    name: str
    charge: float            # C
    mass: float              # kg
    npart_per_cell: float = 0
    density: float = 0.0     # m^-3, uniform inside box
    temp: Sequence[float] = (0.0, 0.0, 0.0)    # K
    drift: Sequence[float] = (0.0, 0.0, 0.0)   # kg m/s
    box_lo: Sequence[float] = (-1e300, -1e300, -1e300)
    box_hi: Sequence[float] = (1e300, 1e300, 1e300)
    bc_particle: Optional[Sequence[str]] = None  # default: follows the field BCs
    zero_current: bool = False
    immobile: bool = False


@dataclass
class Laser:
    boundary: str                    # 'x_min' | 'x_max'
    amp: float
    omega: float
    pol_angle: float = 0.0
    t_start: float = 0.0
    t_end: Optional[float] = None    # default: deck t_end (laser.f90:40)
    profile: Optional[Callable] = None   # f(y, z) -> array, default 1
    phase: Optional[Callable] = None     # f(y, z) -> array, default 0
    t_profile: Optional[Callable] = None  # f(time) -> float, default 1

    @staticmethod
    def amp_from_intensity_w_cm2(i_w_cm2: float) -> float:
        """deck/deck_laser_block.f90:132-136"""
        return math.sqrt(i_w_cm2 / (c * epsilon0 / 2.0)) * 100.0


@dataclass
class Deck:
    ndims: int
    n: Sequence[int]
    xmin: Sequence[float]
    xmax: Sequence[float]
    bc_field: Sequence[str]                 # 2*ndims names, x_min,x_max,y_min,...
    species: List[Species] = field(default_factory=list)
    lasers: List[Laser] = field(default_factory=list)
    t_end: float = 0.0
    nsteps: int = -1
    dt_multiplier: float = 0.95
    nproc: Sequence[int] = (1, 1, 1)
    seed: int = 7842432
    dt_snapshot: float = -1.0
    field_order: int = 2                    # control block field_order: 2, 4 or 6 (fields.f90:32-46)
    maxwell_solver: str = "yee"             # yee | lehe_x | lehe_y | pukhov | custom (2D, order 2)
    stencil_custom: Optional[dict] = None   # custom solver: betaxy, betayx, deltax, deltay, dt
    smooth_currents: bool = False           # control block (deck_control_block.F90:343,456-471)
    smooth_iterations: int = 1
    smooth_compensation: bool = False
    smooth_strides: Sequence[int] = (1,)    # 'auto' = (1, 2, 3, 4)
    hc_push: bool = False                   # build flag -DHC_PUSH (Makefile:264): Higuera-Cary rotation, particles.F90:386-398
    cuts: Optional[dict] = None             # axis -> (mins, maxs): slabs re-cut by the load balancer (balance.F90:383-436)
    # boundaries block (deck_boundaries_block.f90:143-152; defaults setup.F90:80-83); only read when some field
    # boundary is cpml_laser / cpml_outflow
    cpml_thickness: int = 6
    cpml_kappa_max: float = 20.0
    cpml_a_max: float = 0.15
    cpml_sigma_max: float = 0.7
    # window block (deck_window_block.f90): move_window, window_v_x, window_start_time, window_stop_time.  The
    # boundary conditions after the move are the ones before it here (bc_x_min_after_move ... default to them).
    move_window: bool = False
    window_v_x: float = 0.0
    window_start_time: float = 0.0
    window_stop_time: float = 1.0e300
    # state of the moving window (window.F90:62-94): x_grid_min and xb_min as shift_window has accumulated them
    # (None: the grid of setup_grid); dx is frozen at its setup_grid value the first time the window moves
    window_grid_min: Optional[float] = None
    window_xb_min: Optional[float] = None
    window_dx: Optional[float] = None
    window_shifts: int = 0

    # -- grid (setup.F90:162-204) ------------------------------------------
    def cpml_t(self) -> int:
        """cpml_thickness as the code sees it: 0 unless some field boundary is a CPML (mpi_routines.F90:285)."""
        return int(self.cpml_thickness) if any(b in ("cpml_laser", "cpml_outflow") for b in self.bc_field) else 0

    def ncells(self, d: int) -> int:
        """nx_global after mpi_routines.F90:295-296: EVERY axis grows by 2 cpml_thickness cells as soon as one
        boundary is a CPML; `n` keeps the deck's own cell counts."""
        return int(self.n[d]) + 2 * self.cpml_t()

    def dx(self, d: int) -> float:
        # setup.F90:168: length_x / REAL(nx_global - 2 * cpml_thickness); not recomputed when the window moves
        if d == 0 and self.window_dx is not None:
            return self.window_dx
        return (self.xmax[d] - self.xmin[d]) / float(self.n[d])

    def grid_min(self, d: int) -> float:
        # setup.F90:169,180: x_grid_min = x_min - dx * cpml_thickness, then shifted to the cell centre
        if d == 0 and self.window_grid_min is not None:
            return self.window_grid_min
        return (self.xmin[d] - self.dx(d) * self.cpml_t()) + self.dx(d) / 2.0

    def shift_window_geometry(self) -> None:
        """The grid part of one pass of shift_window's loop (window.F90:73-86): x_grid_min = x_global(1) + dx,
        xb_min = xb_global(1) + dx, x_min = xb_min + dx * cpml_thickness, x_max = xb_global(nx_global+1) - dx *
        cpml_thickness, each accumulated in floating point exactly like that; dx and length_x stay."""
        dx = self.dx(0)
        t = self.cpml_t()
        if self.window_dx is None:
            xb = self.xmin[0] - dx * t                      # setup.F90:169,176
            self.window_grid_min = xb + dx / 2.0
            self.window_xb_min = xb
            self.window_dx = dx
            self.xmin, self.xmax = list(self.xmin), list(self.xmax)
        self.window_grid_min = self.window_grid_min + dx
        self.window_xb_min = self.window_xb_min + dx
        self.xmin[0] = self.window_xb_min + dx * t
        self.xmax[0] = (self.window_xb_min + float(self.ncells(0) + 1 - 1) * dx) - dx * t
        self.window_shifts += 1

    def x_global(self, d: int, i):
        """Cell-centre coordinate of global cell i (1-based), setup.F90:188."""
        return self.grid_min(d) + (np.asarray(i, dtype=np.float64) - 1.0) * self.dx(d)

    def bc_codes(self) -> List[int]:
        out = [BC["periodic"]] * 6
        for i, name in enumerate(self.bc_field):
            out[i] = BC[name]
        return out

    def species_bc_codes(self, s: Species) -> List[int]:
        names = s.bc_particle if s.bc_particle is not None else self.bc_field
        out = [BC["periodic"]] * 6
        for i, name in enumerate(names):
            out[i] = BC[name]
        return out

    def any_open(self) -> bool:
        # boundary.F90:41-57
        return any(b in ("simple_laser", "simple_outflow", "open") for b in self.bc_field)   # not the CPML codes

    # -- timestep (setup.F90:577-711; 1D :574-606; 3D :700-746) --------------
    def dt_plasma_frequency(self) -> float:
        min_dt = 1000000.0
        k_max = 2.0 * pi / min(self.dx(d) for d in range(self.ndims))
        for s in self.species:
            fac1 = q0 ** 2 / s.mass / epsilon0
            fac2 = 3.0 * k_max ** 2 * kb / s.mass
            omega2 = fac1 * s.density + fac2 * max(s.temp)
            if omega2 <= 1e-50:
                continue
            omega = math.sqrt(omega2)
            if 2.0 * pi / omega < min_dt:
                min_dt = 2.0 * pi / omega
        return min_dt / 2.0

    def dt_laser(self) -> float:
        v = 1.7976931348623157e308
        for l in self.lasers:
            v = min(v, 2.0 * pi / l.omega)
        return v / 2.0

    def dt(self) -> float:
        d = [self.dx(i) for i in range(self.ndims)]
        if self.ndims == 1:
            solver = d[0] / c
        elif self.ndims == 2:
            solver = d[0] * d[1] / math.sqrt(d[0] ** 2 + d[1] ** 2) / c
        else:
            solver = d[0] * d[1] * d[2] / math.sqrt(
                (d[0] * d[1]) ** 2 + (d[1] * d[2]) ** 2 + (d[2] * d[0]) ** 2) / c
        if self.maxwell_solver == "yee":
            # cfl is a function of field_order (fields.f90:38-44)
            cfl = {2: 1.0, 4: 6.0 / 7.0, 6: 120.0 / 149.0}[self.field_order]
            dt = cfl * solver
        elif self.ndims == 3 and self.maxwell_solver in ("lehe_x", "lehe_y", "lehe_z"):
            a = {"lehe_x": 0, "lehe_y": 1, "lehe_z": 2}[self.maxwell_solver]   # epoch3d setup.F90:707-715
            o1, o2 = [d[i] for i in range(3) if i != a]
            dt = min(d[a], o1 * o2 / math.sqrt(o1 ** 2 + o2 ** 2)) / c
        else:
            dt = min(d) / c        # setup.F90:645-649 (Lehe, Pukhov), epoch3d :717-721 (Cowan), epoch1d :579-581
        if self.any_open():
            dt = min(dt, solver)
        if self.maxwell_solver == "custom":
            dt = float(self.stencil_custom["dt"])
        dtp = self.dt_plasma_frequency()
        if dtp > 1e-50:
            dt = min(dt, dtp)
        dtl = self.dt_laser()
        if dtl > 1e-50:
            dt = min(dt, dtl)
        if self.maxwell_solver == "custom" and self.dt_multiplier < 1.0:
            return dt          # setup.F90:657-668: dt_multiplier is overridden to 1 for the custom solver
        return self.dt_multiplier * dt

    def maxwell_solver_code(self) -> int:
        return {"custom": -1, "yee": 0, "lehe_x": 2, "lehe_y": 3, "lehe_z": 4, "cowan": 5,
                "pukhov": 6}[self.maxwell_solver]  # constants.F90:173-180

    STENCIL_KEYS = ("alphax", "alphay", "alphaz", "betaxy", "betaxz", "betayx", "betayz", "betazx", "betazy",
                    "gammax", "gammay", "gammaz", "deltax", "deltay", "deltaz")

    def stencil(self) -> dict:
        """set_maxwell_solver: epoch1d fields.f90:48-62, epoch2d :51-86, epoch3d :53-162."""
        o = {k: 0.0 for k in self.STENCIL_KEYS}
        o["alphax"] = o["alphay"] = o["alphaz"] = 1.0
        ms, nd = self.maxwell_solver, self.ndims
        if ms == "yee":
            return o
        ok = {1: ("lehe_x", "custom"), 2: ("lehe_x", "lehe_y", "pukhov", "custom"),
              3: ("lehe_x", "lehe_y", "lehe_z", "cowan", "pukhov", "custom")}[nd]
        if ms not in ok:
            raise NotImplementedError(f"maxwell_solver {ms} does not exist in epoch{nd}d")
        d = [self.dx(i) for i in range(nd)] + [0.0] * (3 - nd)
        dx, dy, dz = d
        dt = self.dt()

        def lehe_delta(h):
            r = h / (c * dt)
            return 0.25 * (1.0 - r ** 2 * math.sin(0.5 * pi / r) ** 2)

        if ms == "custom":
            for k in self.STENCIL_KEYS[3:]:
                o[k] = float(self.stencil_custom.get(k, 0.0))
        elif ms == "lehe_x":
            if nd >= 2:
                o["betaxy"] = 0.125 * (dx / dy) ** 2
                o["betayx"] = 0.125
            if nd == 3:
                o["betaxz"] = 0.125 * (dx / dz) ** 2
                o["betazx"] = 0.125
            o["deltax"] = lehe_delta(dx)
        elif ms == "lehe_y":
            o["betayx"] = 0.125 * (dy / dx) ** 2
            o["betaxy"] = 0.125
            if nd == 3:
                o["betayz"] = 0.125 * (dy / dz) ** 2
                o["betazy"] = 0.125
            o["deltay"] = lehe_delta(dy)
        elif ms == "lehe_z":
            o["betazx"] = 0.125 * (dz / dx) ** 2
            o["betazy"] = 0.125 * (dz / dy) ** 2
            o["betaxz"] = 0.125
            o["betayz"] = 0.125
            o["deltaz"] = lehe_delta(dz)
        elif ms == "pukhov":
            delta = min(d[:nd])
            o["betayx"] = 0.125 * (delta / dx) ** 2
            o["betaxy"] = 0.125 * (delta / dy) ** 2
            if nd == 3:
                o["betaxz"] = 0.125 * (delta / dz) ** 2
                o["betazx"] = o["betayx"]
                o["betazy"] = o["betaxy"]
                o["betayz"] = o["betaxz"]
        elif ms == "cowan":
            delta = min(dx, dy, dz)
            c1, c2, c3 = (delta / dx) ** 2, (delta / dy) ** 2, (delta / dz) ** 2
            cx1 = 1.0 / (c1 * c2 + c2 * c3 + c1 * c3)
            cx2 = 1.0 - c1 * c2 * c3 * cx1
            o["betayx"] = 0.125 * c1 * cx2
            o["betaxy"] = 0.125 * c2 * cx2
            o["betaxz"] = 0.125 * c3 * cx2
            o["betazx"], o["betazy"], o["betayz"] = o["betayx"], o["betaxy"], o["betaxz"]
            o["gammax"] = c2 * c3 * (0.0625 - 0.125 * c2 * c3 * cx1)
            o["gammay"] = c1 * c3 * (0.0625 - 0.125 * c1 * c3 * cx1)
            o["gammaz"] = c1 * c2 * (0.0625 - 0.125 * c1 * c2 * cx1)
        # alpha = 1 - 2 beta - 2 beta' - 4 gamma - 3 delta, term by term as in the reference
        o["alphax"] = 1.0 - 2.0 * o["betaxy"] - 2.0 * o["betaxz"] - 4.0 * o["gammax"] - 3.0 * o["deltax"]
        o["alphay"] = 1.0 - 2.0 * o["betayx"] - 2.0 * o["betayz"] - 4.0 * o["gammay"] - 3.0 * o["deltay"]
        o["alphaz"] = 1.0 - 2.0 * o["betazx"] - 2.0 * o["betazy"] - 4.0 * o["gammaz"] - 3.0 * o["deltaz"]
        return o

    # -- decomposition (mpi_routines.F90:317-351) ---------------------------
    def cell_ranges(self, d: int):
        if self.cuts and d in self.cuts:
            return list(self.cuts[d][0]), list(self.cuts[d][1])
        npd = max(1, self.nproc[d]) if d < self.ndims else 1
        ng_ = self.ncells(d) if d < self.ndims else 1
        n0 = ng_ // npd
        nxp = (n0 + 1) * npd - ng_ if n0 * npd != ng_ else npd
        mins, maxs = [], []
        for i in range(1, nxp + 1):
            mins.append((i - 1) * n0 + 1)
            maxs.append(i * n0)
        for i in range(nxp + 1, npd + 1):
            mins.append(nxp * n0 + (i - nxp - 1) * (n0 + 1) + 1)
            maxs.append(nxp * n0 + (i - nxp) * (n0 + 1))
        return mins, maxs

    def nranks(self) -> int:
        r = 1
        for d in range(self.ndims):
            r *= max(1, self.nproc[d])
        return r

    def rank_coords(self, rank: int):
        """MPI_CART row-major with dims=(nprocz,nprocy,nprocx): x fastest
        (mpi_routines.F90:186-187,239-245)."""
        npx = max(1, self.nproc[0])
        npy = max(1, self.nproc[1]) if self.ndims >= 2 else 1
        return (rank % npx, (rank // npx) % npy, rank // (npx * npy))

    def local_extent(self, rank: int):
        """(n_local[3], global_min[3]) of a rank."""
        co = self.rank_coords(rank)
        n, g = [1, 1, 1], [1, 1, 1]
        for d in range(self.ndims):
            mins, maxs = self.cell_ranges(d)
            g[d] = mins[co[d]]
            n[d] = maxs[co[d]] - mins[co[d]] + 1
        return n, g

    # -- laser sources (laser.f90:253-269, 338-357) --------------------------
    def laser_sources(self, rank: int, side: int, time: float):
        """source1/source2 on the local plane of boundary `side` (0 x_min, 1 x_max, 2 y_min, ...): the two
        transverse axes in axis order, (0:n) each, lower axis fastest.  Profile / phase callables get the
        two transverse coordinates in that order ((y, z) on an x face, (x, z) on a y face, (x, y) on a z face)."""
        n, g = self.local_extent(rank)
        axis = side // 2
        tr = [d for d in range(3) if d != axis]
        coords = []
        for d in tr:
            if d < self.ndims:
                coords.append(self.x_global(d, np.arange(0, n[d] + 1) + g[d] - 1))
            else:
                coords.append(np.zeros(1))
        U, V = np.meshgrid(coords[0], coords[1], indexing="xy")  # shape (n_v+1, n_u+1): u fastest
        s1 = np.zeros_like(U)
        s2 = np.zeros_like(U)
        name = ("x", "y", "z")[axis] + ("_min" if side % 2 == 0 else "_max")
        for l in self.lasers:
            if l.boundary != name:
                continue
            t_end = self.t_end if l.t_end is None else l.t_end
            if not (time >= l.t_start and time <= t_end):
                continue
            integral_phase = l.omega * time
            prof = np.ones_like(U) if l.profile is None else np.broadcast_to(l.profile(U, V), U.shape)
            ph = np.zeros_like(U) if l.phase is None else np.broadcast_to(l.phase(U, V), U.shape)
            t_env = (1.0 if l.t_profile is None else float(l.t_profile(time))) * l.amp
            base = t_env * prof * np.sin(integral_phase + ph)
            s1 = s1 + base * math.cos(l.pol_angle)
            s2 = s2 + base * math.sin(l.pol_angle)
        return np.ascontiguousarray(s1.ravel()), np.ascontiguousarray(s2.ravel())

    def has_boundary_source(self, side: int) -> bool:
        return side < 2 * self.ndims and self.bc_field[side] in ("simple_laser", "simple_outflow", "open", "cpml_laser")


class DumpClock:
    """Time-based dump decision, io/diagnostics.F90:1218-1343 (dt_snapshot only)."""

    def __init__(self, dt_snapshot: float):
        self.dt_snapshot = dt_snapshot
        self.time_prev = 0.0
        self.first = True

    def due(self, time: float, last: bool) -> bool:
        dump = False
        if self.first:
            dump = True
            self.first = False
        if last:
            dump = True
        if self.dt_snapshot > 0:
            t0 = self.time_prev + self.dt_snapshot
            if time >= t0:
                while True:
                    t0 = self.time_prev + self.dt_snapshot
                    if t0 > time:
                        break
                    self.time_prev = t0
                dump = True
        return dump


def run(deck: Deck, backend, local_ranks: Sequence[int], on_dump: Optional[Callable] = None,
        max_steps: Optional[int] = None):
    """PROGRAM pic main loop (epoch2d.F90:144-268).  Returns (step, time).

    `local_ranks` are the decomposition ranks this backend instance owns
    (all of them for the in-process oracle, one for a GPU process).
    """
    dt = deck.dt()
    time = 0.0
    step = 0
    clock = DumpClock(deck.dt_snapshot)

    def push_sources(t):
        for lr, rank in enumerate(local_ranks):
            for side in range(2 * deck.ndims):
                if deck.has_boundary_source(side):
                    s1, s2 = deck.laser_sources(rank, side, t)
                    backend.set_laser_source(lr, side, s1, s2)

    # epoch2d.F90:158-162: dt halved, time advanced, bfield_final_bcs, dt restored
    time = time + dt / 2.0
    push_sources(time)
    backend.init()
    if on_dump is not None and clock.due(time, False):
        on_dump(step, time)
    nsteps = deck.nsteps if max_steps is None else max_steps
    window_started, window_shift_fraction = False, 0.0

    def moving_window(t):
        # window.F90:350-392; the shift itself (shift_window, setup_bc_lists, particle_bcs) is the backend's
        nonlocal window_started, window_shift_fraction
        if not deck.move_window:
            return
        if not window_started and deck.window_start_time <= t < deck.window_stop_time:
            window_shift_fraction = 0.0
            window_started = True
        if window_started:
            if t >= deck.window_stop_time or deck.window_v_x <= 0.0:
                return
            window_shift_fraction = window_shift_fraction + dt * deck.window_v_x / deck.dx(0)
            cells = int(np.floor(window_shift_fraction))
            if cells > 0:
                for _ in range(cells):
                    deck.shift_window_geometry()
                backend.shift_window(cells)
                window_shift_fraction = window_shift_fraction - float(cells)

    while True:
        backend.fields_half()
        backend.push()
        backend.current_finish()
        step += 1
        time = time + dt / 2.0
        if (nsteps >= 0 and step >= nsteps) or (deck.t_end > 0 and time >= deck.t_end):
            break
        if on_dump is not None and clock.due(time, False):
            on_dump(step, time)
        time = time + dt / 2.0
        push_sources(time)
        backend.fields_final()
        moving_window(time)
    if on_dump is not None:
        on_dump(step, time)
    return step, time


# ---------------------------------------------------------------------------------------------------------
# Load balancing, host half: the slab boundaries EPOCH's balancer chooses from a load profile.
# The device half (histogram of the resident particles, remap of fields and particles) is not built yet; these
# functions are what it will call, restated from housekeeping/balance.F90 so that a re-cut domain matches EPOCH's.
# ---------------------------------------------------------------------------------------------------------
PUSH_PER_FIELD = 5          # shared_data.F90:821
NCELL_MIN = (3 + 1) // 2 + 1  # constants.F90:562 with png = 3 (triangle shape)


def load_profile(cell_index, n_global: int, n_other_global: int, ng: int = NG):
    """get_load_x / get_load_y (balance.F90:1766-1844) from the particles' global cell indices along one axis.

    cell_index: FLOOR((pos - x_grid_min) / dx + 1.5) of every particle of every species (1-based global cell, ghost
    cells of an open boundary included), already summed over the ranks.  Returns load(1-ng : n_global+ng) as a
    numpy int64 array: push_per_field particles' worth per particle, plus one field column (n_other_global cells)
    per interior cell."""
    load = np.zeros(n_global + 2 * ng, dtype=np.int64)
    idx = np.asarray(cell_index, dtype=np.int64) + ng - 1     # Fortran index `cell + ng` of load(1:), 0-based here
    np.add.at(load, idx, 1)
    load *= PUSH_PER_FIELD
    load[ng:ng + n_global] += n_other_global
    return load


def calculate_breaks(load, nproc: int, ng: int = NG, ncell_min: int = NCELL_MIN):
    """calculate_breaks (balance.F90:1948-2091): split a load profile load(1-ng : sz+ng) into nproc slabs.

    Returns (mins, maxs), 1-based inclusive cell ranges like cell_x_min / cell_x_max.  Control flow as in the
    reference: ideal load per slab, greedy cuts at the nearer side of the cell that crosses it, at least
    ncell_min cells per slab, then single-cell perturbations of every cut while the max-min spread improves
    (the EXITs leave the loop over the cuts, not the iteration), then the two sanity sweeps."""
    load = np.asarray(load, dtype=np.int64)
    sz = load.shape[0] - 2 * ng
    ld = lambda i: int(load[i + ng - 1])                       # load(i), Fortran index
    seg = lambda i0, i1: int(load[i0 + ng - 1:i1 + ng].sum())  # SUM(load(i0:i1))
    mins = [1] * nproc
    maxs = [sz] * nproc            # maxs[proc - 1] = maxs(proc)
    if nproc < 2:
        return mins, maxs
    ideal = int(math.floor(float(seg(1, sz)) / nproc + 0.5))
    proc, old, total = 0, 1, 0
    for idim in range(1, sz + 1):
        total_old = total
        total = total + ld(idim)
        if total >= ideal:
            proc += 1
            maxs[proc - 1] = idim - 1 if ideal - total_old < total - ideal else idim
            nextra = old - maxs[proc - 1] + ncell_min
            if nextra > 0:
                maxs[proc - 1] += nextra
            if proc == nproc - 1:
                break
            old = maxs[proc - 1]
            total = total - ideal

    def sweep_back():
        o = sz
        for p in range(nproc - 1, 0, -1):
            if o - maxs[p - 1] < ncell_min:
                maxs[p - 1] = o - ncell_min
            o = maxs[p - 1]

    def spread():
        lmax, lmin, i0 = -1, None, 1
        for p in range(1, nproc + 1):
            i1 = maxs[p - 1]
            v = seg(i0, i1) if i1 >= i0 else 0
            lmax = max(lmax, v)
            lmin = v if lmin is None else min(lmin, v)
            i0 = i1 + 1
        return lmax, lmin

    sweep_back()
    # load_var_best = HUGE(1): the default-INTEGER huge (2^31 - 1) assigned to an INTEGER(i8), balance.F90:2006 --
    # a spread of 2^31 - 1 or more therefore rejects every perturbation and the loop exits after one iteration
    best = 2 ** 31 - 1
    lmax = lmin = None             # undefined in the reference until a perturbation was possible
    for _ in range(1000):
        for i in range(1, nproc):
            left = False
            for sign in (-1, +1):
                old_maxs = maxs[i - 1]
                if sign < 0:
                    o = 0 if i == 1 else maxs[i - 2]
                    ok = old_maxs - o - 1 >= ng
                else:
                    o = maxs[i]
                    ok = o - old_maxs - 1 >= ng
                if ok:
                    maxs[i - 1] = old_maxs + sign
                    lmax, lmin = spread()
                    if lmax - lmin < best:
                        left = True
                        break
                    maxs[i - 1] = old_maxs
            if left:
                break
        if lmax is not None and lmax - lmin < best:
            best = lmax - lmin
        else:
            break
    sweep_back()
    o = 0
    for p in range(1, nproc):
        if maxs[p - 1] - o < ncell_min:
            maxs[p - 1] = o + ncell_min
        o = maxs[p - 1]
    for p in range(2, nproc + 1):
        mins[p - 1] = maxs[p - 2] + 1
    return mins, maxs
