"""Moving window (housekeeping/window.F90): shift_window / shift_fields / insert_particles /
remove_particles and the particle_bcs that follows.

The reference holds no numbers for a moving-window run (parity unpinned).  The oracle's restatement is
held to what the routines are built for -- the arrays move one cell, the grid moves with them to the bit on the
host and in the oracle, a decomposed run equals the one-rank run, a pulse followed at c stays put in the window, the
plasma stays uniform -- and the CUDA path (epb_shift_window) is held to the oracle."""
import numpy as np
import pytest

from epoch_b200 import deck as D
from oracle.oracle import Oracle
from tests import decks

FIELDS = ("ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz")


def window_deck(ndims, n, ppc=4, nproc=(1, 1, 1), laser=True, plasma=True, temp_k=1.0e7, nsteps=12):
    """Thermal plasma (or vacuum) with a laser on x_min, open x, periodic y / z, window moving at c from t = 0."""
    dk = decks.thermal(ndims, n, ppc=ppc, temp_k=temp_k, nproc=nproc, nsteps=nsteps,
                       bc=["simple_laser" if laser else "open", "open"] + ["periodic"] * (2 * (ndims - 1)))
    if not plasma:
        dk.species = []
    for s in dk.species:
        s.bc_particle = ["open", "open"] + ["periodic"] * (2 * (ndims - 1))
    if laser:
        lam = 8.0 * dk.dx(0)
        w0 = 0.25 * (dk.xmax[1] - dk.xmin[1]) if ndims >= 2 else 1.0
        yc = 0.5 * (dk.xmax[1] + dk.xmin[1]) if ndims >= 2 else 0.0
        zc = 0.5 * (dk.xmax[2] + dk.xmin[2]) if ndims >= 3 else 0.0
        t0 = 6.0 * lam / D.c / 8.0
        dk.lasers = [D.Laser("x_min", D.Laser.amp_from_intensity_w_cm2(1.0e14), 2 * D.pi * D.c / lam,
                             profile=lambda y, z: D.gauss(np.sqrt((y - yc) ** 2 + (z - zc) ** 2), 0.0, w0),
                             t_profile=lambda t: D.gauss(t, t0, 0.4 * t0), t_end=1.0)]
    dk.move_window = True
    dk.window_v_x = D.c
    dk.window_start_time = 0.0
    return dk


def _assemble(o, dk, name):
    nd = dk.ndims
    shape = tuple(dk.n[d] if d < nd else 1 for d in (2, 1, 0))
    full = np.zeros(shape)
    for rk in range(o.nranks):
        info = o.rank_info(rk)
        a = o.interior(rk, name)
        lo = [info["gmin"][d] - 1 if d < nd else 0 for d in range(3)]
        full[lo[2]:lo[2] + a.shape[0], lo[1]:lo[1] + a.shape[1], lo[0]:lo[0] + a.shape[2]] = a
    return full


def test_shift_moves_arrays_and_grid_by_one_cell():
    dk = window_deck(2, (20, 12), laser=False, plasma=False)
    o = Oracle(dk)
    o.init()
    rng = np.random.default_rng(3)
    for name in FIELDS:
        o.field(0, name)[...] = rng.standard_normal(o.field_shape(0))
    before = {name: o.field(0, name).copy() for name in FIELDS}
    g0 = (dk.grid_min(0), dk.xmin[0], dk.xmax[0], dk.dx(0))
    dk.shift_window_geometry()
    o.shift_window(1)                      # asserts that the deck's grid equals the oracle's to the bit
    assert dk.dx(0) == g0[3]
    assert dk.grid_min(0) == g0[0] + g0[3] and dk.xmin[0] == g0[1] + g0[3]
    nx = dk.n[0]
    for name in FIELDS:
        new, old = o.field(0, name), before[name]
        # everything left of the cells shift_fields fixes up (nx-1 .. nx+1) is the old array one cell on
        hi = 5 + nx - 2 if name in FIELDS[:6] else 5 + nx + 4
        assert np.array_equal(new[:, 5:-5, :hi], old[:, 5:-5, 1:hi + 1]), name
    # the incoming cell: zero snapshots, averaged neighbours (window.F90:130-141)
    ex = o.field(0, "ex")
    assert np.all(ex[0, :, 5 + nx - 1] == 0.0) and np.all(ex[0, :, 5 + nx] == 0.0)
    assert np.array_equal(ex[0, :, 5 + nx - 2], 0.5 * (ex[0, :, 5 + nx - 3] + ex[0, :, 5 + nx - 1]))


@pytest.mark.parametrize("ndims,n,nproc", [(1, (96,), (3, 1, 1)), (2, (48, 16), (2, 2, 1)), (3, (24, 8, 8), (2, 1, 2))])
def test_decomposed_window_run_equals_single_rank(ndims, n, nproc):
    """Vacuum + laser (the inserted plasma depends on the per-rank random streams, the fields do not)."""
    res = []
    for np_ in ((1, 1, 1), nproc):
        dk = window_deck(ndims, n, nproc=np_, plasma=False, nsteps=20)
        o = Oracle(dk)
        D.run(dk, o, list(range(o.nranks)))
        assert dk.window_shifts >= 8
        res.append({name: _assemble(o, dk, name) for name in FIELDS[:6]})
    assert max(np.abs(res[0][k]).max() for k in ("ey", "bz")) > 0
    for name in FIELDS[:6]:
        assert np.array_equal(res[0][name], res[1][name]), name


def test_pulse_followed_at_c_stays_in_the_window():
    """2D vacuum: a pulse launched from x_min before the window starts keeps its place in the moving box while the
    box's x_min advances by one dx per shift."""
    dk = window_deck(2, (96, 8), plasma=False, nsteps=400)
    lam = 8.0 * dk.dx(0)
    dk.window_start_time = 16.0 * lam / D.c / 8.0      # the pulse (centred at 6 lam / 8 c) is inside by then
    o = Oracle(dk)
    xmin0 = dk.xmin[0]
    cent = []

    def dump(step, t):
        ey = o.interior(0, "ey")[0]
        e2 = (ey ** 2).sum(axis=0)
        if e2.sum() > 0 and dk.window_shifts > 0:
            cent.append((dk.window_shifts, float((np.arange(dk.n[0]) * e2).sum() / e2.sum())))

    dk.dt_snapshot = 10 * dk.dt()
    D.run(dk, o, [0], dump, max_steps=120)
    assert dk.window_shifts >= 60
    assert abs(dk.xmin[0] - (xmin0 + dk.window_shifts * dk.dx(0))) <= 1e-12 * dk.dx(0) * dk.window_shifts
    c = np.array(cent)
    assert len(c) >= 4
    # centroid (in cells of the window) moves by far less than the window itself did
    assert np.ptp(c[:, 1]) <= 0.1 * (c[-1, 0] - c[0, 0])


def test_plasma_stays_uniform_under_the_window():
    dk = window_deck(2, (24, 12), ppc=8, laser=False, nsteps=30)
    o = Oracle(dk)
    o.auto_load()
    n0 = o.count(0, 0)
    D.run(dk, o, [0])
    assert dk.window_shifts >= 15
    cc = o.cell_counts(0, 0)[0]
    assert abs(o.count(0, 0) - n0) <= 0.05 * n0
    assert abs(cc.mean() - 8.0) <= 0.4
    # the last inserted particles sit in the last cell of the window, with the deck's weight
    ins = o.window_inserted(0, 0)
    assert len(ins) == dk.window_shifts * 8 * dk.n[1]
    last = ins[-8 * dk.n[1]:]
    assert np.all(last[:, 0] >= dk.xmax[0] - dk.dx(0) * (1 + 1e-9)) and np.all(last[:, 0] < dk.xmax[0] * (1 + 1e-12))
    w = dk.species[0].density * dk.dx(0) * dk.dx(1) / 8.0
    assert np.allclose(ins[:, -1], w, rtol=1e-12)


class _Stream:
    """The rank's KISS stream with random_box_muller on top (random_generator.f90:112-173: polar method, the second
    deviate of a pair is kept for the next call), written independently of the oracle's C++ from the Fortran."""
    def __init__(self, seed, n=200000):
        from oracle.oracle import kiss
        self.u, self.i, self.cached = kiss(seed, n), 0, None

    def random(self):
        v = self.u[self.i]
        self.i += 1
        return float(v)

    def box_muller(self, stdev, mu):
        if self.cached is not None:
            r, self.cached = self.cached * stdev + mu, None
            return r
        while True:
            r1, r2 = 2.0 * self.random() - 1.0, 2.0 * self.random() - 1.0
            w = r1 * r1 + r2 * r2
            if 2.2250738585072014e-308 < w < 1.0:
                break
        w = np.sqrt((-2.0 * np.log(w)) / w)
        self.cached = r2 * w
        return r1 * w * stdev + mu


def test_insert_particles_against_an_independent_restatement():
    """epoch2d window.F90:182-320 once more, in Python straight from the Fortran (order of the random draws, the
    y weights, the weight formula), on the rank's KISS stream: every inserted particle equal to the oracle's bit
    for bit, for two species (the second one with a fractional particle count per cell)."""
    dk = window_deck(2, (12, 6), ppc=3, laser=False, temp_k=2.0e7)
    dk.species.append(D.Species("proton", D.q0, 1836.2 * D.m0, npart_per_cell=2.5, density=3.0e24,
                                temp=(1.0e6, 2.0e6, 3.0e6), drift=(1.0e-24, 0.0, -2.0e-24),
                                bc_particle=["open", "open", "periodic", "periodic"]))
    o = Oracle(dk)           # no auto_load: the stream is untouched (setup.F90:566-571: seed + rank, 1000 draws)
    o.init()
    x_grid_max = float(dk.x_global(0, dk.n[0]))
    dx, dy = dk.dx(0), dk.dx(1)
    dk.shift_window_geometry()
    o.shift_window(1)
    g = _Stream(dk.seed + 0)
    for isp, s in enumerate(dk.species):
        npc = int(np.floor(s.npart_per_cell))
        frac = s.npart_per_cell - npc
        x0 = x_grid_max + 0.5 * dx
        want = []
        for iy in range(1, dk.n[1] + 1):
            n_frac = 0
            if frac > 0.0 and g.random() < frac:
                n_frac = 1
            wdata = dx * dy / (npc + n_frac)
            for _ in range(npc + n_frac):
                cf = 0.5 - g.random()
                x = x0 + g.random() * dx
                y = float(dk.x_global(1, iy)) - cf * dy
                gy = [0.5 * (0.25 + cf * cf + cf), 0.75 - cf * cf, 0.5 * (0.25 + cf * cf - cf)]
                p = []
                for i in range(3):
                    t = d = 0.0
                    for k in range(3):
                        t = t + gy[k] * s.temp[i]
                        d = d + gy[k] * s.drift[i]
                    p.append((t, d))
                p = [g.box_muller(np.sqrt(t * D.kb * s.mass), d) for t, d in p]
                wl = 0.0
                for k in range(3):
                    wl = wl + gy[k] * s.density
                want.append([x, y] + p + [wl * wdata])
        got = o.window_inserted(0, isp)
        assert got.shape == (len(want), 6), (isp, got.shape, len(want))
        assert np.array_equal(got, np.array(want)), isp


# ---------------------------------------------------------------------------------------------------------
# CUDA path against the oracle: epb_shift_window + epb_append_species
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("ndims,n", [(1, (64,)), (2, (40, 24)), (3, (20, 9, 8))])
def test_window_matches_oracle_gpu(ndims, n):
    from tests.gpu_util import FIELDS as GF, make_pair, rel_l2
    dk = window_deck(ndims, n, ppc=4, nsteps=14)
    o, sim = make_pair(dk, strict=True)

    class Both:
        """deck.run backend that drives the oracle and the device side by side"""
        def set_laser_source(self, lr, side, s1, s2):
            o.set_laser_source(0, side, s1, s2); sim.set_laser_source(0, side, s1, s2)
        def init(self): o.init(); sim.init()
        def fields_half(self): o.fields_half(); sim.fields_half()
        def push(self): o.push(); sim.push()
        def current_finish(self): o.current_finish(); sim.current_finish()
        def fields_final(self): o.fields_final(); sim.fields_final()
        def shift_window(self, cells):
            o.window_clear_inserted()
            o.shift_window(cells)
            sim.shift_window(cells, [o.window_inserted(0, isp) for isp in range(len(dk.species))])

    D.run(dk, Both(), [0])
    assert dk.window_shifts >= 6
    for name in GF:
        assert rel_l2(sim.download_field(name), o.field(0, name)) <= 1e-12, name
    for isp in range(len(dk.species)):
        assert sim.count(isp) == o.count(0, isp)
        assert np.array_equal(sim.cell_counts(isp), o.cell_counts(0, isp))
