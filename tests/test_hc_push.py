"""Higuera-Cary momentum rotation (-DHC_PUSH, particles.F90:386-398; 1D :345-357, 3D :423-435).

The reference holds no numbers for it.  The oracle's lines are pinned twice: (1) against an independent
numpy restatement of particles.F90:382-428 on uniform fields, and (2) on what defines the scheme (Higuera &
Cary, Phys. Plasmas 24, 052104): a particle that moves with the E x B drift velocity feels no force at ANY
time step, which the Boris rotation satisfies only to O((w_c dt)^2).

Reference quirk, restated as it is: `beta = alpha * bx_part` (:391-393) uses the gathered field BEFORE the
shape normalisation (`fac` = (1/2)^ndims for the triangle shape is folded into cmratio, :162,255, not into
bx_part), so with the default shape the reference's HC gamma is computed from 2^ndims x B and the drift
property only holds for `fac = 1` (tophat).  test_hc_drift_property shows both: the formula with
alpha*fac keeps the drift to round-off, the reference's form (= the oracle = the device kernel) does not.
The CUDA path (push_generic<ND, true>) is held to the oracle bit for bit."""
import numpy as np
import pytest

from epoch_b200 import deck as D
from oracle.oracle import Oracle
from tests import decks


def _rotate(u, e_part, b_part, cmratio, ccmratio, alpha, hc):
    """particles.F90:382-428 for arrays of particles; e_part/b_part are the un-normalised gathers."""
    um = u + cmratio * e_part
    if hc:
        g2 = (um ** 2).sum(axis=1) + 1.0
        beta = alpha * b_part
        beta2 = (beta ** 2).sum(axis=1)
        sigma = g2 - beta2
        bdu = (beta * um).sum(axis=1)
        gamma = np.sqrt(0.5 * (sigma + np.sqrt(sigma ** 2 + 4.0 * (beta2 + bdu ** 2))))
    else:
        gamma = np.sqrt((um ** 2).sum(axis=1) + 1.0)
    t = b_part * (ccmratio / gamma)[:, None]
    t2 = t ** 2
    tau = 1.0 / (1.0 + t2.sum(axis=1))
    tx, ty, tz = t[:, 0], t[:, 1], t[:, 2]
    ux, uy, uz = um[:, 0], um[:, 1], um[:, 2]
    upx = ((1.0 + t2[:, 0] - t2[:, 1] - t2[:, 2]) * ux + 2.0 * ((tx * ty + tz) * uy + (tx * tz - ty) * uz)) * tau
    upy = ((1.0 - t2[:, 0] + t2[:, 1] - t2[:, 2]) * uy + 2.0 * ((ty * tz + tx) * uz + (ty * tx - tz) * ux)) * tau
    upz = ((1.0 - t2[:, 0] - t2[:, 1] + t2[:, 2]) * uz + 2.0 * ((tz * tx + ty) * ux + (tz * ty - tx) * uy)) * tau
    return np.stack([upx, upy, upz], axis=1) + cmratio * e_part


def _constants(dk, s):
    fac = 0.5 ** dk.ndims
    dt = dk.dt()
    cmratio = s.charge * (0.5 * dt * fac) / (D.c * s.mass)
    return fac, dt, cmratio, D.c * cmratio, 0.5 * s.charge * dt / s.mass


@pytest.mark.parametrize("ndims", [1, 2, 3])
@pytest.mark.parametrize("hc", [False, True])
def test_oracle_rotation_against_numpy_restatement(ndims, hc):
    n = {1: (16,), 2: (12, 10), 3: (8, 7, 6)}[ndims]
    dk = decks.thermal(ndims, n, ppc=3, temp_k=3.0e9)
    dk.species[0].zero_current = True
    dk.hc_push = hc
    s = dk.species[0]
    o = Oracle(dk)
    o.auto_load()
    o.init()
    e = np.array([3.0e10, -2.0e10, 1.0e10])
    b = np.array([150.0, -80.0, 220.0])
    for k, name in enumerate(("ex", "ey", "ez")):
        o.field(0, name)[...] = e[k]
    for k, name in enumerate(("bx", "by", "bz")):
        o.field(0, name)[...] = b[k]
    p0 = o.get_particles(0, 0)
    o.push_only()
    p1 = o.get_particles(0, 0)
    fac, dt, cmratio, ccmratio, alpha = _constants(dk, s)
    u0 = p0[:, ndims:ndims + 3] / (D.c * s.mass)
    u1 = _rotate(u0, np.tile(e / fac, (len(u0), 1)), np.tile(b / fac, (len(u0), 1)), cmratio, ccmratio, alpha, hc)
    got = p1[:, ndims:ndims + 3] / (D.c * s.mass)
    assert np.abs(got - u1).max() <= 1e-12 * np.abs(u1).max()


@pytest.mark.parametrize("ndims", [1, 2, 3])
def test_hc_drift_property(ndims):
    gamma_d, b0 = 5.0, 2000.0
    dk = decks.thermal(ndims, (8,) * ndims, ppc=1, temp_k=0.0, length=8.0e-3)
    s = dk.species[0]
    fac, dt, cmratio, ccmratio, alpha = _constants(dk, s)
    vd = D.c * np.sqrt(1.0 - 1.0 / gamma_d ** 2)
    u0 = np.array([[gamma_d * vd / D.c, 0.0, 0.0]])
    e = np.array([[0.0, vd * b0, 0.0]]) / fac      # B = b0 z, v = vd x  =>  E = -v x B = vd b0 y
    b = np.array([[0.0, 0.0, b0]]) / fac
    assert abs(s.charge) * b0 / (gamma_d * s.mass) * dt > 0.3      # Boris is visibly off at this w_c dt
    scale = abs(u0[0, 0])
    # the scheme itself (normalised field in beta): no force on the drifting particle
    assert np.abs(_rotate(u0, e, b, cmratio, ccmratio, alpha * fac, True) - u0).max() <= 1e-13 * scale
    # Boris: spurious force
    assert np.abs(_rotate(u0, e, b, cmratio, ccmratio, alpha, False) - u0).max() >= 1e-3 * scale
    # the reference's form (un-normalised field in beta): exact only when fac == 1
    ref = np.abs(_rotate(u0, e, b, cmratio, ccmratio, alpha, True) - u0).max()
    assert ref >= 1e-3 * scale


@pytest.mark.parametrize("ndims", [1, 2])
def test_hc_equals_boris_without_b(ndims):
    """beta = 0  =>  sigma = gamma^2 and the HC gamma is the Boris gamma (to rounding of the two square roots)."""
    out = []
    for hc in (False, True):
        dk = decks.thermal(ndims, (24,) if ndims == 1 else (12, 10), ppc=4, temp_k=2.0e9)
        dk.species[0].zero_current = True
        dk.hc_push = hc
        o = Oracle(dk)
        o.auto_load()
        o.init()
        rng = np.random.default_rng(3)
        for name in ("ex", "ey", "ez"):
            a = o.field(0, name)
            a[...] = rng.normal(size=a.shape) * 1e11
        for _ in range(3):
            o.push_only()
            o.particle_bcs()
        out.append(o.get_particles(0, 0))
    a, b = out
    assert a.shape == b.shape
    assert np.all(np.abs(a - b) <= 1e-13 * np.max(np.abs(a), axis=0))


def test_hc_differs_from_boris_in_general():
    out = []
    for hc in (False, True):
        dk = decks.thermal(2, (12, 10), ppc=4, temp_k=2.0e9)
        dk.hc_push = hc
        o = Oracle(dk)
        o.auto_load()
        o.init()
        rng = np.random.default_rng(3)
        for name in ("ex", "ey", "ez", "bx", "by", "bz"):
            a = o.field(0, name)
            a[...] = rng.normal(size=a.shape) * (1e11 if name[0] == "e" else 300.0)
        o.push_only()
        out.append(o.get_particles(0, 0))
    a, b = out
    assert np.abs(a[:, 2:5] - b[:, 2:5]).max() > 1e-6 * np.abs(a[:, 2:5]).max()


# ---------------------------------------------------------------------------------------------------------
# CUDA path against the oracle
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("ndims,n", [(1, (64,)), (2, (32, 24)), (3, (10, 9, 8))])
@pytest.mark.parametrize("sort_interval", [1, 3])
def test_hc_push_matches_oracle_gpu(ndims, n, sort_interval):
    from tests.gpu_util import FIELDS, make_pair, rel_l2, run_both, set_random_fields, sorted_rows
    dk = decks.thermal(ndims, n, ppc=6, temp_k=2.0e9, two_species=(ndims == 2))
    dk.hc_push = True
    o, sim = make_pair(dk, strict=True, sort_interval=sort_interval)
    set_random_fields(o, sim, dk, seed=5, e_amp=1e10, b_amp=300.0)
    # one push from identical state: bit-exact particle state
    o.push(); sim.push()
    for isp in range(len(dk.species)):
        a, b = sorted_rows(sim.download_species(isp)), sorted_rows(o.get_particles(0, isp))
        assert np.array_equal(a, b)
    o2, sim2 = make_pair(dk, strict=True, sort_interval=sort_interval)
    run_both(dk, o2, sim2, 8)
    for name in FIELDS:
        assert rel_l2(sim2.download_field(name), o2.field(0, name)) <= 1e-12, name
    for isp in range(len(dk.species)):
        assert sim2.count(isp) == o2.count(0, isp)
        assert np.array_equal(sim2.cell_counts(isp), o2.cell_counts(0, isp))


@pytest.mark.gpu
def test_hc_push_performance_build_gpu():
    """strict_fp = 0 (FMA contraction allowed): particle state within 1e-13 of each column's scale after one push."""
    from tests.gpu_util import make_pair, set_random_fields, sorted_rows
    dk = decks.thermal(2, (32, 24), ppc=6, temp_k=2.0e9)
    dk.hc_push = True
    o, sim = make_pair(dk, strict=False, sort_interval=2)
    set_random_fields(o, sim, dk, seed=5, e_amp=1e10, b_amp=300.0)
    o.push(); sim.push()
    a, b = sorted_rows(sim.download_species(0)), sorted_rows(o.get_particles(0, 0))
    assert a.shape == b.shape
    assert np.all(np.abs(a - b) <= 1e-13 * np.max(np.abs(b), axis=0))
