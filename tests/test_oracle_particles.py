"""Oracle particle half: the reference holds no golden numbers for push/deposit
(SURVEY.md §4), so the restatement is checked through the property particles.F90:30-34
claims: the Esirkepov deposit satisfies d(rho)/dt + div(J) = 0 exactly on the grid."""
import numpy as np
import pytest

from epoch_b200 import deck as D
from oracle.oracle import Oracle
from tests import decks


def _rho(dk, pos, w, q):
    """Charge density with the triangle shape at cell centres (calc_charge_density,
    io/calc_df.F90:608-685), periodic wrap; pos are positions at a half time level."""
    nd = dk.ndims
    n = [dk.n[d] for d in range(nd)]
    rho = np.zeros(n[::-1])
    cells, gs = [], []
    for d in range(nd):
        r = (pos[:, d] - dk.grid_min(d)) / dk.dx(d)
        cx = np.floor(r + 0.5)
        f = cx - r
        cells.append(cx.astype(np.int64))
        gs.append([0.5 * (0.25 + f * f + f), 0.75 - f * f, 0.5 * (0.25 + f * f - f)])
    vol = np.prod([dk.dx(d) for d in range(nd)])
    import itertools
    for offs in itertools.product((-1, 0, 1), repeat=nd):
        wgt = q * w / vol
        idx = []
        for d in range(nd):
            wgt = wgt * gs[d][offs[d] + 1]
            idx.append((cells[d] + offs[d]) % n[d])
        np.add.at(rho, tuple(idx[::-1]), wgt)
    return rho


def _half_positions(dk, p, mass, sign):
    """x(t -/+ dt/2) from the stored x(t), p(t): x + sign * u c dt/2 / gamma (particles.F90:297-302)."""
    nd = dk.ndims
    u = p[:, nd:nd + 3] / (mass * D.c)
    gamma = np.sqrt((u ** 2).sum(axis=1) + 1.0)
    return p[:, :nd] + sign * u[:, :nd] * (D.c * dk.dt() / 2.0) / gamma[:, None]


@pytest.mark.parametrize("ndims,n", [(1, (32,)), (2, (16, 12)), (3, (8, 7, 6))])
def test_charge_conservation(ndims, n):
    dk = decks.thermal(ndims, n, ppc=6, temp_k=5.0e8, nsteps=3)  # hot: many cell crossings
    o = Oracle(dk)
    o.auto_load()
    o.init()
    s = dk.species[0]
    rng = np.random.default_rng(1)
    for name in ("ex", "ey", "ez", "bx", "by", "bz"):
        a = o.field(0, name)
        a[...] = rng.normal(size=a.shape) * (1e9 if name[0] == "e" else 3.0)
    o.lib = None
    for _ in range(3):
        before = o.get_particles(0, 0)
        o.push_only()
        after = o.get_particles(0, 0)
        o.current_finish()          # folds ghost J back (periodic) before particle_bcs reorders
        # x(t+dt/2) from the old state, x(t+3dt/2) from the new state
        rho0 = _rho(dk, _half_positions(dk, before, s.mass, +1), before[:, -1], s.charge)
        rho1 = _rho(dk, _half_positions(dk, after, s.mass, +1), after[:, -1], s.charge)
        div = np.zeros_like(rho0)
        for d, name in zip(range(ndims), ("jx", "jy", "jz")):
            j = o.interior(0, name)
            ax = 2 - d
            div += (j - np.roll(j, 1, axis=ax)).reshape(rho0.shape) / dk.dx(d)
        resid = (rho1 - rho0) / dk.dt() + div
        scale = np.abs(div).max()
        assert scale > 0
        assert np.abs(resid).max() < 1e-11 * scale
        o.particle_bcs()
        assert o.count(0, 0) == before.shape[0]


def test_loader_counts_and_weights():
    dk = decks.thermal(2, (12, 10), ppc=5, nproc=(2, 2, 1))
    o = Oracle(dk)
    o.auto_load()
    total = 0
    for r in range(o.nranks):
        cc = o.cell_counts(r, 0)
        assert (cc == 5).all()      # npart_per_cell in every valid cell (helper.F90:556-583)
        p = o.get_particles(r, 0)
        total += p.shape[0]
        # weight = density * dx*dy / npart_in_cell for a uniform plasma (helper.F90:711-770)
        assert np.allclose(p[:, -1], dk.species[0].density * dk.dx(0) * dk.dx(1) / 5, rtol=1e-12)
    assert total == 12 * 10 * 5


def test_decomposed_matches_single_rank_fields():
    """Same particles pushed on 1 rank and on 2x2 ranks give the same J to round-off."""
    dk1 = decks.thermal(2, (16, 12), ppc=4, temp_k=2.0e8)
    dk4 = decks.thermal(2, (16, 12), ppc=4, temp_k=2.0e8, nproc=(2, 2, 1))
    o1, o4 = Oracle(dk1), Oracle(dk4)
    o1.auto_load()
    allp = o1.get_particles(0, 0)
    for r in range(4):
        info = o4.rank_info(r)
        m = np.ones(allp.shape[0], bool)
        for d in range(2):
            m &= (allp[:, d] >= info["min_local"][d]) & (allp[:, d] < info["max_local"][d])
        o4.set_particles(r, 0, allp[m])
    for o in (o1, o4):
        o.init()
        for _ in range(4):
            o.fields_half(); o.push(); o.current_finish(); o.fields_final()
    assert sum(o4.count(r, 0) for r in range(4)) == o1.count(0, 0)
    full = o1.interior(0, "ey")[0]
    for r in range(4):
        info = o4.rank_info(r)
        x0, y0 = info["gmin"][0] - 1, info["gmin"][1] - 1
        loc = o4.interior(r, "ey")[0]
        ref = full[y0:y0 + loc.shape[0], x0:x0 + loc.shape[1]]
        assert np.allclose(loc, ref, rtol=0, atol=1e-9 * np.abs(full).max())


@pytest.mark.parametrize("bc", ["reflect", "periodic"])
@pytest.mark.parametrize("nproc", [(1, 1, 1), (2, 2, 1)])
def test_per_species_current_bcs_are_consistent(bc, nproc):
    """c_bc_mixed path (current_bcs per species + particle_clear_bcs, particles.F90:645, boundary.F90:783-804):
    forced on a deck whose species agree, it must give the fields of the ordinary path (interior cells; the
    ghost cells of J are cleared in the mixed path only) to summation-order round-off."""
    from epoch_b200 import deck as D
    from tests import decks

    def run(force):
        dk = decks.thermal(2, (24, 20), ppc=5, temp_k=4.0e8, bc=bc, two_species=True, nproc=nproc)
        dk.force_mixed_bc = force
        o = Oracle(dk)
        o.auto_load()
        D.run(dk, o, list(range(o.nranks)), None, max_steps=6)
        return o
    a, b = run(False), run(True)
    for f in ("ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz"):
        for r in range(a.nranks):
            x, y = a.interior(r, f), b.interior(r, f)
            assert np.max(np.abs(x - y)) <= 1e-13 * max(np.max(np.abs(x)), 1e-300), (f, r)


# ---------------------------------------------------------------------------------------------------------
# Exact identities of the Boris scheme (particles.F90:382-428): independent of any restatement
# ---------------------------------------------------------------------------------------------------------
def _one_push_uniform(ndims, e, b, temp_k=2.0e9):
    n = {1: (16,), 2: (12, 10), 3: (8, 7, 6)}[ndims]
    dk = decks.thermal(ndims, n, ppc=4, temp_k=temp_k)
    dk.species[0].zero_current = True          # the fields stay uniform
    o = Oracle(dk)
    o.auto_load()
    o.init()
    for k, name in enumerate(("ex", "ey", "ez")):
        o.field(0, name)[...] = e[k]
    for k, name in enumerate(("bx", "by", "bz")):
        o.field(0, name)[...] = b[k]
    p0 = o.get_particles(0, 0)[:, ndims:ndims + 3]
    o.push_only()
    p1 = o.get_particles(0, 0)[:, ndims:ndims + 3]
    return dk, p0, p1


@pytest.mark.parametrize("ndims", [1, 2, 3])
def test_boris_uniform_e_is_exact(ndims):
    """B = 0: the two half kicks add up to p(t + dt) = p(t) + q E dt, whatever gamma is."""
    e = np.array([3.0e10, -2.0e10, 1.0e10])
    dk, p0, p1 = _one_push_uniform(ndims, e, np.zeros(3))
    dp = dk.species[0].charge * e * dk.dt()
    assert np.abs(p1 - p0 - dp).max() <= 2e-14 * np.abs(p1).max()


@pytest.mark.parametrize("ndims", [1, 2, 3])
def test_boris_rotation_angle_is_exact(ndims):
    """E = 0: |p| is conserved to round-off and p_perp turns about B by exactly 2 atan(q B dt / (2 gamma m)),
    p_parallel stays (Birdsall & Langdon's tan(theta/2) identity of the rotation :413-423)."""
    b = np.array([120.0, -260.0, 310.0])
    dk, p0, p1 = _one_push_uniform(ndims, np.zeros(3), b)
    s = dk.species[0]
    bn = b / np.linalg.norm(b)
    assert np.abs(np.linalg.norm(p1, axis=1) / np.linalg.norm(p0, axis=1) - 1.0).max() <= 1e-14
    par0, par1 = p0 @ bn, p1 @ bn
    assert np.abs(par1 - par0).max() <= 2e-14 * np.abs(p0).max()
    perp0, perp1 = p0 - np.outer(par0, bn), p1 - np.outer(par1, bn)
    gamma = np.sqrt(1.0 + (p0 ** 2).sum(axis=1) / (s.mass * D.c) ** 2)
    theta = 2.0 * np.arctan(s.charge * np.linalg.norm(b) * dk.dt() / (2.0 * gamma * s.mass))
    # signed angle from perp0 to perp1 about bn; dp/dt = q v x B turns p about B by -q|B|/(gamma m) t
    sin_a = np.einsum("ij,ij->i", np.cross(perp0, perp1), bn[None, :])
    cos_a = np.einsum("ij,ij->i", perp0, perp1)
    ang = np.arctan2(sin_a, cos_a)
    assert np.abs(ang + theta).max() <= 1e-12
    assert np.abs(theta).min() > 1e-3          # a visible turn, not a null test


@pytest.mark.parametrize("ndims,n", [(1, (32,)), (2, (16, 12)), (3, (8, 7, 6))])
def test_total_current_identity(ndims, n):
    """The box total of the deposited current, which the continuity equation cannot see (it fixes div J only):
    along a gridded axis  sum(J_a) dV = sum_p q w (x_a(t + 3dt/2) - x_a(t + dt/2)) / dt  (Esirkepov's weights
    telescope), along an ignorable axis  sum(J_a) dV = sum_p q w v_a  (particles.F90:573, epoch1d :489-506)."""
    dk = decks.thermal(ndims, n, ppc=6, temp_k=5.0e8, drift=(4.0e-23, -2.0e-23, 3.0e-23))
    o = Oracle(dk)
    o.auto_load()
    o.init()
    s = dk.species[0]
    rng = np.random.default_rng(2)
    for name in ("ex", "ey", "ez", "bx", "by", "bz"):
        a = o.field(0, name)
        a[...] = rng.normal(size=a.shape) * (1e9 if name[0] == "e" else 3.0)
    before = o.get_particles(0, 0)
    o.push_only()
    after = o.get_particles(0, 0)
    o.current_finish()
    dv = np.prod([dk.dx(d) for d in range(ndims)])
    w = after[:, -1]
    mom = after[:, ndims:ndims + 3]
    gamma = np.sqrt(1.0 + (mom ** 2).sum(axis=1) / (s.mass * D.c) ** 2)
    x_a = _half_positions(dk, before, s.mass, +1)      # x(t + dt/2) from the old state
    x_b = _half_positions(dk, after, s.mass, +1)       # x(t + 3dt/2) from the new state
    # `after` positions are wrapped into the box by particle_bcs only later (push_only): no unwrapping needed
    for a, name in enumerate(("jx", "jy", "jz")):
        total = o.interior(0, name).sum() * dv
        if a < ndims:
            want = s.charge * (w * (x_b[:, a] - x_a[:, a])).sum() / dk.dt()
        else:
            want = s.charge * (w * mom[:, a] / (gamma * s.mass)).sum()
        assert abs(total / want - 1.0) <= 1e-11, (name, total, want)


def _coords(dk, o, name):
    """Coordinates of every array element (ghost cells included) of field `name`, Yee-staggered as in
    setup.F90:124-134; arrays are indexed [k][j][i]."""
    stag = {"ex": (0,), "ey": (1,), "ez": (2,), "bx": (1, 2), "by": (0, 2), "bz": (0, 1)}[name]
    shape = o.field(0, name).shape
    out = []
    for d in range(3):
        m = shape[2 - d]
        if d < dk.ndims:
            x = dk.grid_min(d) + (np.arange(m) - 5) * dk.dx(d) + (0.5 * dk.dx(d) if d in stag else 0.0)
        else:
            x = np.zeros(m)
        sh = [1, 1, 1]
        sh[2 - d] = m
        out.append(x.reshape(sh))
    return out


@pytest.mark.parametrize("ndims,n", [(1, (24,)), (2, (12, 10)), (3, (8, 7, 6))])
def test_gather_is_exact_for_linear_fields(ndims, n):
    """The quadratic B-spline gather (e_part.inc / b_part.inc) reproduces a field that is linear in space exactly,
    at the position x(t + dt/2) and with each component's own Yee staggering: with B = 0 the push must give
    dp = q E(x_half) dt per particle; with E = 0 and one linear B component the turn about that axis must be
    2 atan(q B(x_half) dt / 2 gamma m)."""
    dk = decks.thermal(ndims, n, ppc=3, temp_k=1.0e9)
    s = dk.species[0]
    s.zero_current = True
    L = [dk.xmax[d] - dk.xmin[d] for d in range(ndims)]
    rng = np.random.default_rng(5)

    def linear(o, name, a, b):
        X = _coords(dk, o, name)
        v = a + sum(b[d] * X[d] / L[d] for d in range(ndims))
        o.field(0, name)[...] = np.broadcast_to(v, o.field(0, name).shape)

    def at(pos, a, b):
        return a + sum(b[d] * pos[:, d] / L[d] for d in range(ndims))

    # E
    o = Oracle(dk)
    o.auto_load()
    o.init()
    coef = {}
    for name in ("ex", "ey", "ez"):
        coef[name] = (rng.normal() * 1e10, rng.normal(size=3) * 1e10)
        linear(o, name, *coef[name])
    before = o.get_particles(0, 0)
    o.push_only()
    after = o.get_particles(0, 0)
    xh = _half_positions(dk, before, s.mass, +1)
    for k, name in enumerate(("ex", "ey", "ez")):
        dp = after[:, ndims + k] - before[:, ndims + k]
        want = s.charge * at(xh, *coef[name]) * dk.dt()
        assert np.abs(dp - want).max() <= 1e-12 * np.abs(want).max(), name
    # B, one component at a time
    for m, name in enumerate(("bx", "by", "bz")):
        o = Oracle(dk)
        o.auto_load()
        o.init()
        a, b = 200.0, rng.normal(size=3) * 60.0
        linear(o, name, a, b)
        before = o.get_particles(0, 0)
        o.push_only()
        after = o.get_particles(0, 0)
        xh = _half_positions(dk, before, s.mass, +1)
        p0, p1 = before[:, ndims:ndims + 3], after[:, ndims:ndims + 3]
        gamma = np.sqrt(1.0 + (p0 ** 2).sum(axis=1) / (s.mass * D.c) ** 2)
        theta = 2.0 * np.arctan(s.charge * at(xh, a, b) * dk.dt() / (2.0 * gamma * s.mass))
        i, j = (m + 1) % 3, (m + 2) % 3
        ang = np.arctan2(p0[:, i] * p1[:, j] - p0[:, j] * p1[:, i], p0[:, i] * p1[:, i] + p0[:, j] * p1[:, j])
        assert np.abs(ang + theta).max() <= 1e-11, name
        assert np.abs(p1[:, m] - p0[:, m]).max() <= 1e-14 * np.abs(p0).max()


# ---------------------------------------------------------------------------------------------------------
# particle_bcs (boundary.F90:1029-1462) on hand-placed particles
# ---------------------------------------------------------------------------------------------------------
def _bc_deck(bc):
    dk = decks.thermal(1, (16,), ppc=1, temp_k=0.0, bc=bc, length=16.0e-6)
    dk.species[0].zero_current = True
    return dk


def _place(o, xs, pxs):
    p = o.get_particles(0, 0)[:len(xs)].copy()
    p[:, 0] = xs
    p[:, 1] = pxs
    p[:, 2:4] = 0.0
    o.set_particles(0, 0, p)
    return p


def test_reflecting_wall_mirrors_position_and_momentum():
    dk = _bc_deck("reflect")
    o = Oracle(dk)
    o.auto_load()
    o.init()
    dx, dt = dk.dx(0), dk.dt()
    v = 0.5 * D.c
    px = D.m0 * v / np.sqrt(1 - 0.25)
    # 0.3 dx from either wall, moving towards it: crosses by v dt - 0.3 dx
    p = _place(o, [0.3 * dx, 16e-6 - 0.3 * dx, 8e-6], [-px, +px, px])
    o.push()
    q = o.get_particles(0, 0)
    assert q.shape[0] == 3
    q = q[np.argsort(q[:, 0])]
    over = v * dt - 0.3 * dx
    assert over > 0
    assert np.isclose(q[0, 0], over, rtol=1e-12) and np.isclose(q[0, 1], +px, rtol=1e-14)              # 2 x_min - pos
    assert np.isclose(q[2, 0], 16e-6 - over, rtol=1e-12) and np.isclose(q[2, 1], -px, rtol=1e-14)      # 2 x_max - pos
    assert np.isclose(q[1, 0], 8e-6 + v * dt, rtol=1e-13)


def test_periodic_wrap_and_open_deletion_threshold():
    # periodic: leaves through x_max, re-enters at x_min + overshoot
    dk = _bc_deck("periodic")
    o = Oracle(dk)
    o.auto_load()
    o.init()
    dx, dt = dk.dx(0), dk.dt()
    v = 0.5 * D.c
    px = D.m0 * v / np.sqrt(1 - 0.25)
    _place(o, [16e-6 - 0.3 * dx], [px])
    o.push()
    q = o.get_particles(0, 0)
    assert q.shape[0] == 1 and np.isclose(q[0, 0], v * dt - 0.3 * dx, rtol=1e-10)
    # open: a particle stays alive (and on this rank) until it is beyond x_min_outer = x_min - ((1 + png) / 2) dx
    # = x_min - 2 dx (utilities.f90:367-369, integer division), then it is deleted
    dk = _bc_deck("open")
    o = Oracle(dk)
    o.auto_load()
    o.init()
    dx, dt = dk.dx(0), dk.dt()
    _place(o, [0.3 * dx], [-px])
    steps_inside = int(np.floor((0.3 * dx + 2.0 * dx) / (v * dt)))   # pushes after which it is still >= x_min_outer
    for s in range(steps_inside):
        o.push()
        assert o.count(0, 0) == 1, s
    assert o.get_particles(0, 0)[0, 0] < 0.0                         # outside the domain, not yet deleted
    o.push()
    assert o.count(0, 0) == 0
