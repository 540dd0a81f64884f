"""Thermal particle boundaries (boundary.F90:1104-1148 and its copies for the other faces,
particle_temperature.F90:388-460): a particle that passes the outer edge of a thermal wall comes back with a
momentum drawn from the wall's temperature -- flux-weighted (Rayleigh) along the normal, Maxwellian along the wall.

The reference draws from the rank's serial KISS stream, the device from counter-based per-particle streams, so oracle
and device are compared on the distributions they must both produce."""
import math

import numpy as np
import pytest

from epoch_b200 import deck as D
from oracle import oracle as O
from tests import decks

T_WALL = 2.0e7   # K, colder than the 1e8 K plasma so that the walls visibly cool it


def _deck(ndims, n, ppc=12):
    # fields: clamped (a deck writes bc_x_min_field = reflect); particles: thermal walls on every face
    dk = decks.thermal(ndims, n, ppc=ppc, temp_k=1.0e8, bc="reflect")
    dk.species[0].bc_particle = ["thermal"] * (2 * ndims)
    return dk


def _walls(dk):
    return [(s, (T_WALL, 2 * T_WALL, 3 * T_WALL)) for s in range(2 * dk.ndims)]


def test_thermal_wall_oracle_distribution():
    """Oracle, particle_bcs alone (2D): particles placed beyond the outer edges of thermal walls come back mirrored
    about the edge, moving inwards, with <p_n^2> = 2 m k T_n along the wall normal (flux-weighted: a Rayleigh deviate)
    and <p_t^2> = m k T_t along the wall; particles that have not reached the outer edge are left alone."""
    dk = _deck(2, (16, 12), ppc=400)
    o = O.Oracle(dk)
    o.auto_load()
    for side, t in _walls(dk):
        o.set_boundary_temperature(0, 0, side, t)
    p = o.get_particles(0, 0).copy()
    n = p.shape[0]
    dx = dk.dx(0)
    rng = np.random.default_rng(4)
    third = n // 3
    # first third beyond x_max_outer, second beyond x_min_outer, the rest between x_max and x_max_outer (not re-emitted)
    p[:third, 0] = dk.xmax[0] + (2.0 + rng.random(third) * 0.3) * dx
    p[third:2 * third, 0] = dk.xmin[0] - (2.0 + rng.random(third) * 0.3) * dx
    p[2 * third:, 0] = dk.xmax[0] + rng.random(n - 2 * third) * 1.9 * dx
    o.set_particles(0, 0, p)
    O.lib().orc_setup_bc_lists(o._h)
    o.particle_bcs()
    q = o.get_particles(0, 0)
    assert q.shape == p.shape
    assert np.array_equal(q[2 * third:], p[2 * third:])                      # inside the outer edge: untouched
    hi, lo = q[:third], q[third:2 * third]
    assert np.allclose(hi[:, 0], 2 * (dk.xmax[0] + 2 * dx) - p[:third, 0], rtol=0, atol=1e-20)
    assert np.allclose(lo[:, 0], 2 * (dk.xmin[0] - 2 * dx) - p[third:2 * third, 0], rtol=0, atol=1e-20)
    assert (hi[:, 2] < 0).all() and (lo[:, 2] > 0).all()
    mk = D.m0 * D.kb
    for h in (hi, lo):
        assert abs(np.mean(h[:, 2] ** 2) / (2 * mk * T_WALL) - 1.0) < 0.05
        assert abs(np.mean(h[:, 3] ** 2) / (mk * 2 * T_WALL) - 1.0) < 0.05
        assert abs(np.mean(h[:, 4] ** 2) / (mk * 3 * T_WALL) - 1.0) < 0.05
        assert abs(np.mean(h[:, 3])) < 0.05 * math.sqrt(mk * 2 * T_WALL)


@pytest.mark.gpu
@pytest.mark.parametrize("ndims,n", [(1, (64,)), (2, (32, 24)), (3, (16, 8, 6))])
def test_thermal_walls_gpu(ndims, n):
    """Device vs oracle with thermal walls on every face: identical particle counts (nobody is lost), all particles
    inside the outer edges, and the same cooling of the plasma by the colder walls within the statistical scatter of
    the two random streams; J and the fields stay finite and comparable in size."""
    from epoch_b200.pic import Simulation
    dk = _deck(ndims, n)
    o = O.Oracle(dk)
    o.auto_load()
    p0 = o.get_particles(0, 0).copy()
    sim = Simulation(dk, strict_fp=True, sort_interval=2, capacity_factor=2.0)
    sim.upload_species(0, p0)
    for side, t in _walls(dk):
        o.set_boundary_temperature(0, 0, side, t)
        sim.set_boundary_temperature(0, side, t)
    o.init(); sim.init()
    nsteps = 60
    for _ in range(nsteps):
        o.fields_half(); sim.fields_half()
        o.push(); sim.push()
        o.current_finish(); sim.current_finish()
        o.fields_final(); sim.fields_final()
    a, b = sim.download_species(0), o.get_particles(0, 0)
    assert a.shape == b.shape == p0.shape
    shift = 2.0                                           # x_min_outer = x_min - 2 dx (png = 3)
    for d in range(ndims):
        lo, hi = dk.xmin[d] - shift * dk.dx(d), dk.xmax[d] + shift * dk.dx(d)
        for q in (a, b):
            assert (q[:, d] >= lo - 0.5 * dk.dx(d)).all() and (q[:, d] <= hi + 0.5 * dk.dx(d)).all()
    ke = lambda q: np.mean(np.sum(q[:, ndims:ndims + 3] ** 2, axis=1))
    ke0, kea, keb = ke(p0), ke(a), ke(b)
    assert kea < 0.97 * ke0 and keb < 0.97 * ke0           # the walls are colder than the plasma
    assert abs(kea / keb - 1.0) < 0.05, (kea / ke0, keb / ke0)
    for f in ("jx", "ex"):
        x, y = sim.download_field(f), o.field(0, f)
        assert np.isfinite(x).all()
        assert 0.5 < (np.abs(x).mean() + 1e-300) / (np.abs(y).mean() + 1e-300) < 2.0


@pytest.mark.gpu
@pytest.mark.parametrize("ndims,n", [(2, (16, 96)), (3, (16, 16, 12))])
def test_thermal_wall_distribution_gpu(ndims, n):
    """The device's re-emission alone: a tenth of a nearly cold plasma is uploaded beyond the outer edges of the x
    walls, is pushed (it only moves further out) and comes back with the wall's flux-weighted / Maxwellian momenta,
    pointing inwards; the rest of the plasma stays cold."""
    from epoch_b200.pic import Simulation
    dk = _deck(ndims, n, ppc=400 if ndims == 2 else 100)
    o = O.Oracle(dk)
    o.auto_load()
    p = o.get_particles(0, 0).copy()
    npart = p.shape[0]
    dx = dk.dx(0)
    rng = np.random.default_rng(4)
    rng.shuffle(p, axis=0)
    k = npart // 20
    p[:k, 0] = dk.xmax[0] + (2.05 + rng.random(k) * 0.25) * dx
    p[k:2 * k, 0] = dk.xmin[0] - (2.05 + rng.random(k) * 0.25) * dx
    p[:, ndims:ndims + 3] *= 1e-3                      # nearly at rest: the push barely moves them
    sim = Simulation(dk, strict_fp=True, sort_interval=2, capacity_factor=2.0)
    sim.upload_species(0, p)
    for side, t in _walls(dk):
        sim.set_boundary_temperature(0, side, t)
    sim.init()
    sim.fields_half(); sim.push()
    q = sim.download_species(0)
    assert q.shape == p.shape
    mk = D.m0 * D.kb
    hot = np.sum(q[:, ndims:ndims + 3] ** 2, axis=1) > (0.02 ** 2) * mk * 1.0e8     # cold plasma: 1e-3 sqrt(m k 1e8)
    assert hot.sum() == 2 * k
    q = q[hot]
    hi, lo = q[q[:, 0] > 0.5 * (dk.xmin[0] + dk.xmax[0])], q[q[:, 0] < 0.5 * (dk.xmin[0] + dk.xmax[0])]
    assert hi.shape[0] == k and (hi[:, ndims] < 0).all() and (lo[:, ndims] > 0).all()
    assert (hi[:, 0] <= dk.xmax[0] + 2 * dx).all() and (lo[:, 0] >= dk.xmin[0] - 2 * dx).all()
    for h in (hi, lo):
        assert abs(np.mean(h[:, ndims] ** 2) / (2 * mk * T_WALL) - 1.0) < 0.05
        assert abs(np.mean(h[:, ndims + 1] ** 2) / (mk * 2 * T_WALL) - 1.0) < 0.05
        assert abs(np.mean(h[:, ndims + 2] ** 2) / (mk * 3 * T_WALL) - 1.0) < 0.05


def test_thermal_reemission_equals_an_independent_restatement():
    """The thermal branch of particle_bcs (boundary.F90:1104-1148, :1190-1234) with flux_momentum_from_temperature and
    momentum_from_temperature (particle_temperature.F90:388-460) once more, in Python from the Fortran on the rank's
    KISS stream: every re-emitted particle equal to the oracle's bit for bit (order of the draws: two deviates for the
    Rayleigh normal component, then the two tangential ones; the wall temperature interpolated with the triangle
    weights in y; the position mirrored about the OUTER edge)."""
    from tests.test_window import _Stream
    dk = _deck(2, (16, 12), ppc=1)
    dk.species[0].bc_particle = ["thermal", "thermal", "periodic", "periodic"]
    dk.bc_field = ["reflect", "reflect", "periodic", "periodic"]
    o = O.Oracle(dk)          # no auto_load: the stream is untouched (seed + rank, 1000 draws)
    for side, t in _walls(dk)[:2]:
        o.set_boundary_temperature(0, 0, side, t)
    rng = np.random.default_rng(12)
    n = 600
    dx, dy = dk.dx(0), dk.dx(1)
    p = np.zeros((n, 6))
    p[:, 1] = dk.xmin[1] + rng.random(n) * (dk.xmax[1] - dk.xmin[1])
    p[:, 2:5] = rng.standard_normal((n, 3)) * 1.0e-23
    p[:, 5] = 1.0
    # a third beyond x_max_outer, a third beyond x_min_outer, a third between the wall and its outer edge, shuffled
    kind = rng.integers(0, 3, n)
    p[:, 0] = np.where(kind == 0, dk.xmax[0] + (2.0 + rng.random(n) * 0.3) * dx,
                       np.where(kind == 1, dk.xmin[0] - (2.0 + rng.random(n) * 0.3) * dx,
                                dk.xmax[0] + rng.random(n) * 1.9 * dx))
    o.set_particles(0, 0, p)
    O.lib().orc_setup_bc_lists(o._h)
    o.particle_bcs()
    got = o.get_particles(0, 0)
    mo, xo = o.outer()
    info = o.rank_info(0)
    g = _Stream(dk.seed + 0)
    m = dk.species[0].mass
    want = p.copy()
    for P in want:
        part_pos = float(P[0])
        for sgn, beyond, outer, tw in ((-1, part_pos < mo[0], mo[0], _walls(dk)[0][1]),
                                       (+1, part_pos >= xo[0], xo[0], _walls(dk)[1][1])):
            if not beyond:
                continue
            cell_y_r = (float(P[1]) - info["grid_min_local"][1]) / dy
            cell_y = math.floor(cell_y_r + 0.5)
            cf = float(cell_y) - cell_y_r
            cf2 = cf * cf
            gy = (0.5 * (0.25 + cf2 + cf), 0.75 - cf2, 0.5 * (0.25 + cf2 - cf))
            temp = []
            for i in range(3):
                t = 0.0
                for k in range(3):
                    t = t + gy[k] * tw[i]
                temp.append(t)
            direction = -float(sgn)
            mom1 = g.box_muller(math.sqrt(temp[0] * D.kb * m), 0.0)
            mom2 = g.box_muller(math.sqrt(temp[0] * D.kb * m), 0.0)
            P[2] = direction * math.sqrt(mom1 * mom1 + mom2 * mom2)
            P[3] = g.box_muller(math.sqrt(temp[1] * D.kb * m), 0.0)
            P[4] = g.box_muller(math.sqrt(temp[2] * D.kb * m), 0.0)
            P[0] = 2.0 * outer - part_pos
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert (kind == 0).sum() > 100 and (kind == 1).sum() > 100
