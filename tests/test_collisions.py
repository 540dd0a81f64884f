"""Binary collisions (BASELINE config 5): physics_packages/collisions.F90 on the device's cell-resident
layout (epb_collide) and in the CPU oracle (oracle/collisions_oracle.inc).

The reference holds no numbers for collisions (no test deck switches them on), so the pins are what the operators must
satisfy -- momentum and energy conservation pair by pair, the classical isotropisation rate of a temperature
anisotropy -- and, between device and oracle, pair-by-pair agreement on identical random numbers."""
import ctypes as C
import math

import numpy as np
import pytest

from epoch_b200 import deck as D
from oracle import oracle as O
from tests.gpu_util import sorted_rows

M0, Q0, CL = D.m0, D.q0, D.c


def _pairs(n, m1, m2, t_ev, seed):
    rng = np.random.default_rng(seed)
    p1 = rng.normal(size=(n, 3)) * math.sqrt(m1 * t_ev * Q0)
    p2 = rng.normal(size=(n, 3)) * math.sqrt(m2 * t_ev * Q0)
    return p1, p2, rng.random((n, 4))


def _env(m1, m2, q1, q2, nanbu, inter, dens=1e28, loglam=10.0, dt=1e-16, factor=1.0 / 64e28 * 64, np_=1.0):
    # SK: nu * factor * np * dt; NP: s_fac = cell_fac * loglam / (4 pi eps0^2 c^4) with cell_fac = n^2 dt factor dV
    dv = 1e-16
    fac = 1.0 / (32.0 * dens * dv / 64.0)        # ~ user_factor / sum(min(w)) for 32 pairs of weight n dV / 64
    cell_fac = dens * dens * dt * fac * dv
    s_fac = cell_fac * loglam / (4.0 * math.pi * D.epsilon0 ** 2 * CL ** 4)
    pi_fac = (4.0 * math.pi / 3.0) ** (1.0 / 3.0)
    if inter:
        s_fac_prime, sp_den = cell_fac * pi_fac, max(m1, m2) * dens ** (2.0 / 3.0)
    else:
        s_fac_prime, sp_den = cell_fac * pi_fac / dens ** (2.0 / 3.0), max(m1, m2)
    npart = dens * dv
    return np.array([m1, m2, q1, q2, dens, loglam, fac / (1.0 if inter else 2.0), npart, dt, s_fac, s_fac_prime, sp_den,
                     float(inter), float(nanbu)])


def _run_pairs(fn, p1, p2, w1, w2, ran, env):
    a, b = p1.copy(), p2.copy()
    done = np.zeros(p1.shape[0], dtype=np.int32)
    fn(p1.shape[0], a.ctypes.data, b.ctypes.data, w1.ctypes.data, w2.ctypes.data, ran.ctypes.data, env.ctypes.data,
       done.ctypes.data)
    return a, b, done


@pytest.mark.parametrize("nanbu", [0, 1])
@pytest.mark.parametrize("inter", [0, 1])
def test_pair_operators_conserve_momentum_and_energy(nanbu, inter):
    """Every scattered pair of equal weights keeps its total momentum and energy (the scattering is a rotation of
    the relative momentum in the centre-of-momentum frame): to 1e-12 for both operators, like and unlike masses."""
    O.build()
    m1, m2 = M0, (1836.2 * M0 if inter else M0)
    q1, q2 = -Q0, (Q0 if inter else -Q0)
    n = 4096
    p1, p2, ran = _pairs(n, m1, m2, 5000.0, 11)          # 5 keV: mildly relativistic
    w = np.full(n, 3.0e10)
    a, b, done = _run_pairs(O.lib().orc_collide_pairs_test, p1, p2, w, w, ran, _env(m1, m2, q1, q2, nanbu, inter))
    assert done.all()
    e = lambda p, m: CL * np.sqrt(np.sum(p * p, axis=1) + (m * CL) ** 2)
    scale = np.abs(p1).max() + np.abs(p2).max()
    assert np.abs((a + b) - (p1 + p2)).max() <= 1e-12 * scale
    et0, et1 = e(p1, m1) + e(p2, m2), e(a, m1) + e(b, m2)
    assert np.abs(et1 / et0 - 1.0).max() <= 1e-13
    moved = np.abs(a - p1).max(axis=1) > 1e-6 * np.abs(p1).max()
    assert moved.mean() > 0.9                             # and it does scatter


def _aniso_deck(n=(16, 16), ppc=200, dens=1.0e28):
    dx = 1.0e-8
    sp = [D.Species("electron", -Q0, M0, npart_per_cell=ppc, density=dens, temp=(0.0, 0.0, 0.0))]
    return D.Deck(2, list(n), [0.0, 0.0], [dx * n[0], dx * n[1]], ["periodic"] * 4, species=sp)


def _aniso_particles(o, t_par_ev=150.0, t_perp_ev=75.0, seed=5):
    p = o.get_particles(0, 0)
    rng = np.random.default_rng(seed)
    p[:, 2] = rng.normal(size=p.shape[0]) * math.sqrt(M0 * t_par_ev * Q0)
    p[:, 3] = rng.normal(size=p.shape[0]) * math.sqrt(M0 * t_perp_ev * Q0)
    p[:, 4] = rng.normal(size=p.shape[0]) * math.sqrt(M0 * t_perp_ev * Q0)
    return p


def _temps_ev(p):
    t = (p[:, 2:5] ** 2).mean(axis=0) / M0 / Q0
    return t[0], 0.5 * (t[1] + t[2])


def _nu_iso(dens, t_ev, loglam):
    """small-anisotropy limit of the NRL isotropisation rate (SI): dT_perp/dt = -nu (T_perp - T_par)"""
    e2 = Q0 ** 2 / (4.0 * math.pi * D.epsilon0)
    return 8.0 * math.sqrt(math.pi) / 15.0 * e2 ** 2 * dens * loglam / (math.sqrt(M0) * (t_ev * Q0) ** 1.5)


@pytest.mark.parametrize("nanbu", [1, 0])
def test_temperature_isotropisation_oracle(nanbu):
    """A bi-Maxwellian electron plasma relaxes towards isotropy at the classical rate: d(T_par - T_perp)/dt =
    -3 nu (T_par - T_perp).  Collisions only (no push); total momentum and energy are kept to round-off."""
    dk = _aniso_deck()
    o = O.Oracle(dk)
    o.auto_load()
    p = _aniso_particles(o)
    o.set_particles(0, 0, p)
    loglam, dens = 10.0, dk.species[0].density
    tpar0, tperp0 = _temps_ev(p)
    tmean = (tpar0 + 2 * tperp0) / 3.0
    nu = _nu_iso(dens, tmean, loglam)
    nsteps = int(0.5 / (3.0 * nu * dk.dt()))
    pt0, e0 = p[:, 2:5].sum(axis=0), np.sqrt((p[:, 2:5] ** 2).sum(axis=1) + (M0 * CL) ** 2).sum()
    for _ in range(nsteps):
        o.collide([[1.0]], coulomb_log=loglam, use_nanbu=bool(nanbu))
    q = o.get_particles(0, 0)
    assert np.array_equal(q[:, :2], p[:, :2]) and np.array_equal(q[:, 5], p[:, 5])
    pt1, e1 = q[:, 2:5].sum(axis=0), np.sqrt((q[:, 2:5] ** 2).sum(axis=1) + (M0 * CL) ** 2).sum()
    assert np.abs(pt1 - pt0).max() <= 1e-9 * np.abs(q[:, 2:5]).sum() and abs(e1 / e0 - 1.0) < 1e-12
    tpar1, tperp1 = _temps_ev(q)
    decay = (tpar1 - tperp1) / (tpar0 - tperp0)
    want = math.exp(-3.0 * nu * nsteps * dk.dt())
    assert 0.0 < decay < 1.0
    if nanbu:
        # finite anisotropy (A = -1/2) and the operator's own accuracy: within 25 % of the small-anisotropy e-folding
        assert abs(math.log(decay) / math.log(want) - 1.0) < 0.25, (decay, want)
    else:
        # Sentoku-Kemp as the reference codes it relaxes like particles several times faster than the classical rate
        # (measured here: ~5x) -- the "unusual behaviour" its own deck reader warns about when it announces that Nanbu
        # is the default now (deck_collision_block.F90:120-131).  Restated as it is; only the direction is asserted.
        assert decay < want, (decay, want)


def test_coulomb_log_auto_oracle():
    """coulomb_log = auto (calc_coulomb_log, collisions.F90:1288-1316) on a hot dense plasma: ln(Lambda) of a few,
    and the run relaxes; two species so that inter-species pairs are exercised too."""
    dk = _aniso_deck(n=(8, 8), ppc=60)
    dk.species.append(D.Species("proton", Q0, 1836.2 * M0, npart_per_cell=60, density=1.0e28, temp=(1.0e6,) * 3))
    o = O.Oracle(dk)
    o.auto_load()
    p = _aniso_particles(o)
    o.set_particles(0, 0, p)
    e = lambda: sum(np.sqrt((o.get_particles(0, i)[:, 2:5] ** 2).sum(axis=1) + (m * CL) ** 2).sum() * CL
                    for i, m in ((0, M0), (1, 1836.2 * M0)))
    e0 = e()
    t0 = _temps_ev(p)
    for _ in range(40):
        o.collide([[1.0, 1.0], [0.0, 1.0]], coulomb_log=0.0, use_nanbu=True)
    t1 = _temps_ev(o.get_particles(0, 0))
    assert abs(e() / e0 - 1.0) < 1e-12
    assert (t1[0] - t1[1]) < 0.97 * (t0[0] - t0[1])    # ln(Lambda) ~ 2.5 here: four times slower than with the fixed 10


# ---- device ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("nanbu", [0, 1])
@pytest.mark.parametrize("inter", [0, 1])
def test_pair_operators_match_oracle_gpu(nanbu, inter):
    """The device's pair operator against the oracle's on identical pairs and random numbers: the arithmetic is the
    same expression tree (both built without FMA contraction); the math-library calls (log, sin, cos, exp, sinh,
    acos, pow) may differ in the last place, hence 1e-12 of the momentum scale instead of bit equality."""
    from epoch_b200 import lib as L
    lib = L.load()
    m1, m2 = M0, (1836.2 * M0 if inter else M0)
    q1, q2 = -Q0, (Q0 if inter else -Q0)
    n = 20000
    p1, p2, ran = _pairs(n, m1, m2, 2000.0, 23)
    rng = np.random.default_rng(2)
    w1 = np.where(rng.random(n) < 0.5, 3.0e10, 1.0e10)     # unequal weights: SK correction / NP rejection paths
    w2 = np.full(n, 3.0e10)
    env = _env(m1, m2, q1, q2, nanbu, inter)
    a0, b0, d0 = _run_pairs(O.lib().orc_collide_pairs_test, p1, p2, w1, w2, ran, env)
    a1, b1, d1 = _run_pairs(lib.epb_collide_pairs_test, p1, p2, w1, w2, ran, env)
    assert np.array_equal(d0, d1)
    assert np.abs(a1 - a0).max() <= 1e-12 * np.abs(p1).max()
    assert np.abs(b1 - b0).max() <= 1e-12 * np.abs(p2).max()


@pytest.mark.gpu
@pytest.mark.parametrize("nanbu", [1, 0])
def test_temperature_isotropisation_gpu(nanbu):
    """The same relaxation on the device (epb_collide on the slot columns, per-pair counter-based streams): conserved
    totals, the classical rate, and the oracle's decay within the statistical scatter of two independent runs."""
    from epoch_b200.pic import Simulation
    dk = _aniso_deck()
    o = O.Oracle(dk)
    o.auto_load()
    p = _aniso_particles(o)
    o.set_particles(0, 0, p)
    sim = Simulation(dk, strict_fp=True, sort_interval=2, capacity_factor=2.0)
    sim.upload_species(0, p)
    loglam, dens = 10.0, dk.species[0].density
    tpar0, tperp0 = _temps_ev(p)
    nu = _nu_iso(dens, (tpar0 + 2 * tperp0) / 3.0, loglam)
    nsteps = int(0.5 / (3.0 * nu * dk.dt()))
    for _ in range(nsteps):
        o.collide([[1.0]], coulomb_log=loglam, use_nanbu=bool(nanbu))
        sim.collide([[1.0]], coulomb_log=loglam, use_nanbu=bool(nanbu), seed=99)
    q = sim.download_species(0)
    assert q.shape == p.shape
    srt = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]
    assert np.array_equal(srt(q)[:, :2], srt(p)[:, :2])                      # positions untouched
    pt0, pt1 = p[:, 2:5].sum(axis=0), q[:, 2:5].sum(axis=0)
    e = lambda a: np.sqrt((a[:, 2:5] ** 2).sum(axis=1) + (M0 * CL) ** 2).sum()
    assert np.abs(pt1 - pt0).max() <= 1e-9 * np.abs(q[:, 2:5]).sum() and abs(e(q) / e(p) - 1.0) < 1e-12
    tpar1, tperp1 = _temps_ev(q)
    decay = (tpar1 - tperp1) / (tpar0 - tperp0)
    want = math.exp(-3.0 * nu * nsteps * dk.dt())
    assert 0.0 < decay < 1.0
    if nanbu:
        assert abs(math.log(decay) / math.log(want) - 1.0) < 0.25, (decay, want)
    to = _temps_ev(o.get_particles(0, 0))
    decay_o = (to[0] - to[1]) / (tpar0 - tperp0)
    # two independent random streams over 51 200 particles: the remaining anisotropies agree to a few per cent of
    # the initial one (the device's pairing also depends on the order in which the upload's atomics filled the
    # columns, so its result scatters by ~0.01 from run to run; seen: 0.070 (oracle) against 0.075 ... 0.103)
    assert abs(decay - decay_o) < 0.05, (decay, decay_o)


@pytest.mark.gpu
def test_collisions_inside_the_pic_loop_gpu():
    """PROGRAM pic's order (epoch2d.F90:211-250): fields_half, push, collide, current_finish, fields_final -- two
    species, coulomb_log = auto, both operators' inter-species pairs; the particle count and the total energy
    (kinetic + field) are kept while the electron anisotropy relaxes."""
    from epoch_b200.pic import Simulation
    dk = _aniso_deck(n=(32, 16), ppc=40)
    dk.species.append(D.Species("proton", Q0, 1836.2 * M0, npart_per_cell=40, density=1.0e28, temp=(1.0e6,) * 3))
    o = O.Oracle(dk)
    o.auto_load()
    p = _aniso_particles(o)
    sim = Simulation(dk, strict_fp=False, sort_interval=2, capacity_factor=2.0)
    sim.upload_species(0, p)
    sim.upload_species(1, o.get_particles(0, 1))
    sim.init()
    n0 = [sim.count(0), sim.count(1)]
    tot = lambda: sum(sim.field_energy()) + sim.kinetic_energy(0) + sim.kinetic_energy(1)
    e0 = tot()
    for _ in range(30):
        sim.fields_half(); sim.push()
        sim.collide([[1.0, 1.0], [0.0, 1.0]], coulomb_log=0.0, use_nanbu=True)
        sim.current_finish(); sim.fields_final()
    assert [sim.count(0), sim.count(1)] == n0
    assert abs(tot() / e0 - 1.0) < 2e-2
    t1 = _temps_ev(sim.download_species(0))
    t0 = _temps_ev(p)
    assert (t1[0] - t1[1]) < 0.98 * (t0[0] - t0[1])


# ---------------------------------------------------------------------------------------------------------
# A second restatement of the Nanbu / Perez pair (collisions.F90:984-1101 = :516-633), scalar Python written from
# the Fortran independently of oracle/collisions_oracle.inc, against the oracle's pair operator bit for bit.
# The reference holds no numbers for collisions ("parity unpinned"): a slip in the long chain of formulas would
# have to be made twice to pass.  sin / cos / acos / log / exp / sinh are the same libm on both sides.
# ---------------------------------------------------------------------------------------------------------
def _pair_np_python(p1_in, p2_in, w1, w2, ran, m1, m2, q1, q2, s_fac, s_fac_prime, sp_den, inter):
    c, m0_, eps, c_tiny = CL, M0, 2.220446049250313e-16, 2.2250738585072014e-308
    dot = lambda a, b: a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
    p1 = [float(v) / c for v in p1_in]
    p2 = [float(v) / c for v in p2_in]
    p1_norm = [v / m0_ for v in p1]
    p2_norm = [v / m0_ for v in p2]
    if dot(p1_norm, p1_norm) < eps and dot(p2_norm, p2_norm) < eps:
        return None
    vc = [a - b for a, b in zip(p1_norm, p2_norm)]
    if dot(vc, vc) < eps:
        return None
    p1_norm = [v / m1 for v in p1]
    gm1 = math.sqrt(dot(p1_norm, p1_norm) + 1.0) * m1
    p2_norm = [v / m2 for v in p2]
    gm2 = math.sqrt(dot(p2_norm, p2_norm) + 1.0) * m2
    gm = gm1 + gm2
    v1 = [v / gm1 for v in p1]
    v2 = [v / gm2 for v in p2]
    vc = [(a + b) / gm for a, b in zip(p1, p2)]
    vc_sq = dot(vc, vc)
    gamma_rel_inv = math.sqrt(1.0 - vc_sq)
    gc = 1.0 / gamma_rel_inv
    gc_m1_vc = (gc - 1.0) / vc_sq
    t = (gc_m1_vc * dot(vc, v1) - gc) * gm1
    p3 = [a + t * b for a, b in zip(p1, vc)]
    v_sq = dot(vc, v1)
    gm3 = (1.0 - v_sq) * gc * gm1
    v_sq = dot(vc, v2)
    gm4 = (1.0 - v_sq) * gc * gm2
    p_mag2 = dot(p3, p3)
    p_mag = math.sqrt(p_mag2)
    q12 = q1 * q2
    fac = q12 * q12 * s_fac / (gm1 * gm2)
    t1 = gm3 * gm4 / p_mag2 + 1.0
    s12 = fac * gc * p_mag * c / gm * (t1 * t1)
    v_rel = gm * p_mag * c / (gm3 * gm4 * gc)
    s_prime = s_fac_prime * (m1 + m2) * v_rel / sp_den
    s12 = min(s12, s_prime)
    ran1 = float(ran[0])
    ran2 = float(ran[1]) * 2.0 * math.pi
    branch = 0 if s12 < 0.1 else 1 if s12 < 3.0 else 2 if s12 < 6.0 else 3
    if s12 < 0.1:
        cosp = 1.0 + s12 * math.log(max(ran1, 5e-9))
    elif s12 < 3.0:
        a_inv = 0.0056958 + (0.9560202 + (-0.508139 + (0.47913906 + (-0.12788975 + 0.02389567
                 * s12) * s12) * s12) * s12) * s12
        a = 1.0 / a_inv
        cosp = a_inv * math.log(math.exp(-a) + 2.0 * ran1 * math.sinh(a))
    elif s12 < 6.0:
        a = 3.0 * math.exp(-s12)
        cosp = math.log(math.exp(-a) + 2.0 * ran1 * math.sinh(a)) / a
    else:
        cosp = 2.0 * ran1 - 1.0
    cosp = max(min(cosp, 1.0), -1.0)
    sinp = math.sin(math.acos(cosp))
    p_perp2 = p3[0] * p3[0] + p3[1] * p3[1]
    p_perp = math.sqrt(p_perp2)
    p_tot = math.sqrt(p_perp2 + p3[2] * p3[2])
    p_perp_inv = 1.0 / (p_perp + c_tiny)
    mat = [[p3[0] * p3[2] * p_perp_inv, -p3[1] * p_tot * p_perp_inv, p3[0]],
           [p3[1] * p3[2] * p_perp_inv, p3[0] * p_tot * p_perp_inv, p3[1]],
           [-p_perp, 0.0, p3[2]]]
    sinp_cos = sinp * math.cos(ran2)
    sinp_sin = sinp * math.sin(ran2)
    p3 = [mat[i][0] * sinp_cos + mat[i][1] * sinp_sin + mat[i][2] * cosp for i in range(3)]
    p4 = [-v for v in p3]
    # inter_collisions_np :1087-1096 rejects by the weight ratio with a third random number; intra_collisions_np
    # :627-632 always updates both particles
    ran1 = float(ran[2]) if inter else -1.0
    out1, out2 = [float(v) for v in p1_in], [float(v) for v in p2_in]
    if ran1 < w2 / w1:
        t = gc_m1_vc * dot(vc, p3) + gm3 * gc
        out1 = [(a + t * b) * c for a, b in zip(p3, vc)]
    if ran1 < w1 / w2:
        t = gc_m1_vc * dot(vc, p4) + gm4 * gc
        out2 = [(a + t * b) * c for a, b in zip(p4, vc)]
    return out1, out2, branch


@pytest.mark.parametrize("inter", [0, 1])
def test_nanbu_pair_equals_an_independent_restatement_bit_for_bit(inter):
    O.build()
    m1, m2 = M0, (1836.2 * M0 if inter else M0)
    q1, q2 = -Q0, (Q0 if inter else -Q0)
    n = 3000
    p1, p2, ran = _pairs(n, m1, m2, 3000.0, 31 + inter)
    rng = np.random.default_rng(8)
    w1 = np.where(rng.random(n) < 0.5, 3.0e10, 1.0e10)     # unequal weights: the rejection branches
    w2 = np.full(n, 2.0e10)
    branches = [0, 0, 0, 0]
    for loglam, dt in ((10.0, 1e-16), (10.0, 3e-14), (10.0, 4e-13), (10.0, 1e-11)):   # s12 < 0.1, < 3, < 6, >= 6
        env = _env(m1, m2, q1, q2, 1, inter, loglam=loglam, dt=dt)
        a, b, done = _run_pairs(O.lib().orc_collide_pairs_test, p1, p2, w1, w2, ran, env)
        for i in range(n):
            r = _pair_np_python(p1[i], p2[i], w1[i], w2[i], ran[i], m1, m2, q1, q2, env[9], env[10], env[11], inter)
            assert (r is not None) == bool(done[i])
            if r is not None:
                assert np.array_equal(np.array(r[0]), a[i]) and np.array_equal(np.array(r[1]), b[i]), (dt, i)
                branches[r[2]] += 1
    assert min(branches) >= 200, branches          # every branch of the inversion was taken many times


# The same for the Sentoku-Kemp pair (collisions.F90:751-880 = :296-430) with coll_freq (:1119-1142), new_coords
# (:1189-1220) and weighted_particles_correction (:1146-1185).
def _new_coords_python(v):
    c_tiny = 2.2250738585072014e-308
    vmag = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    vtrans = math.sqrt(v[1] * v[1] + v[2] * v[2])
    if vtrans > c_tiny:
        c1 = [x / vmag for x in v]
        c2 = [0.0 / vtrans, v[2] / vtrans, -v[1] / vtrans]
        den = vmag * vtrans
        c3 = [vtrans * vtrans / den, -(v[0] * v[1]) / den, -(v[0] * v[2]) / den]
        return c1, c2, c3
    return [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]


def _pair_sk_python(p1_in, p2_in, w1, w2, ran, m1, m2, q1, q2, dens, log_lambda, factor, np_, dt_coll):
    c, eps, c_tiny, huge = CL, 2.220446049250313e-16, 2.2250738585072014e-308, 1.7976931348623157e308
    cc, mc0, pi, eps0 = CL * CL, 2.73092429345209278e-22, math.pi, D.epsilon0
    dot = lambda a, b: a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
    rnd = iter(float(r) for r in ran)
    p1, p2 = [float(v) for v in p1_in], [float(v) for v in p2_in]
    wr = w1 / w2
    p1_norm, p2_norm = [v / mc0 for v in p1], [v / mc0 for v in p2]
    if dot(p1_norm, p1_norm) < eps and dot(p2_norm, p2_norm) < eps:
        return None
    vc = [a - b for a, b in zip(p1_norm, p2_norm)]
    if dot(vc, vc) < eps:
        return None
    e1 = c * math.sqrt(dot(p1, p1) + (m1 * c) * (m1 * c))
    e2 = c * math.sqrt(dot(p2, p2) + (m2 * c) * (m2 * c))
    vc = [(a + b) * cc / (e1 + e2) for a, b in zip(p1, p2)]
    vc_sq = dot(vc, vc)
    vc_sq_cc = vc_sq / cc
    gamma_rel2 = 1.0 / (1.0 - vc_sq_cc)
    gamma_rel = math.sqrt(gamma_rel2)
    gamma_rel_m1 = gamma_rel2 * vc_sq_cc / (gamma_rel + 1.0)
    p1_vc, p2_vc = dot(p1, vc), dot(p2, vc)
    tvar = p1_vc * gamma_rel_m1 / (vc_sq + c_tiny)
    t = tvar - gamma_rel * e1 / cc
    p3 = [a + b * t for a, b in zip(p1, vc)]
    tvar = p2_vc * gamma_rel_m1 / (vc_sq + c_tiny)
    t = tvar - gamma_rel * e2 / cc
    p4 = [a + b * t for a, b in zip(p2, vc)]
    p3_mag = math.sqrt(dot(p3, p3))
    e3 = gamma_rel * (e1 - p1_vc)
    e4 = gamma_rel * (e2 - p2_vc)
    v3 = [a * cc / e3 for a in p3]
    v4 = [a * cc / e4 for a in p4]
    tvar = 1.0 - (dot(v3, v4) / cc)
    vr = [(a - b) / tvar for a, b in zip(v3, v4)]
    vrabs = math.sqrt(dot(vr, vr))
    # coll_freq
    mu = (m1 * m2) / (m1 + m2)
    nu = 0.0
    if vrabs > 0.0:
        q12 = q1 * q2
        numerator = q12 * q12 * dens * log_lambda
        denominator = 4.0 * pi * (eps0 * eps0) * (mu * mu) * (vrabs * vrabs * vrabs)
        if not (denominator <= 0.0 or math.frexp(numerator)[1] - math.frexp(denominator)[1] >= 1024):
            nu = numerator / denominator
    nu = min(nu * factor * np_ * dt_coll, 0.02)
    c1, c2, c3 = _new_coords_python(vr)
    ran1 = (1.0 - 1.0e-10) * next(rnd) + 0.5e-10
    ran2 = 2.0 * pi * next(rnd)
    delta = math.sqrt(-2.0 * nu * math.log(ran1)) * math.sin(ran2)
    ran2 = 2.0 * pi * next(rnd)
    sin_theta = 2.0 * delta / (1.0 + delta * delta)
    cos_theta = (1.0 - delta * delta) / (1.0 + delta * delta)
    vcr = v3 if m1 > m2 else v4
    vcr2 = dot(vcr, vcr)
    gamma_rel_r = 1.0 / math.sqrt(1.0 - (vcr2 / cc))
    denominator = gamma_rel_r * (cos_theta - math.sqrt(vcr2) / max(vrabs, c_tiny))
    if abs(denominator) > math.sqrt(c_tiny):
        tan_theta_cm = sin_theta / denominator
        tan_theta_cm2 = tan_theta_cm * tan_theta_cm
    else:
        tan_theta_cm = tan_theta_cm2 = huge
    sin_theta = tan_theta_cm / math.sqrt(1.0 + tan_theta_cm2)
    cos_theta = 1.0 / math.sqrt(1.0 + tan_theta_cm2)
    cr, sr = math.cos(ran2), math.sin(ran2)
    p3 = [p3_mag * (a * cos_theta + b * sin_theta * cr + d * sin_theta * sr) for a, b, d in zip(c1, c2, c3)]
    p4 = [-v for v in p3]
    tvar = dot(p3, vc) * gamma_rel_m1 / vc_sq
    t = tvar + gamma_rel * e3 / cc
    p5 = [a + b * t for a, b in zip(p3, vc)]
    tvar = dot(p4, vc) * gamma_rel_m1 / vc_sq
    t = tvar + gamma_rel * e4 / cc
    p6 = [a + b * t for a, b in zip(p4, vc)]
    e5 = c * math.sqrt(dot(p5, p5) + (m1 * c) * (m1 * c))
    e6 = c * math.sqrt(dot(p6, p6) + (m2 * c) * (m2 * c))

    def correction(wtr, p, p_scat, en, en_scat, mass):
        en_after = (1.0 - wtr) * en + wtr * en_scat
        p_after = [(1.0 - wtr) * a + wtr * b for a, b in zip(p, p_scat)]
        p_mag = math.sqrt(dot(p_after, p_after))
        gamma_en = en_after / (mass * cc)
        pm = p_mag / mass / c
        gamma_p = math.sqrt(1.0 + pm * pm)
        if gamma_p < gamma_en:
            delta_p = mass * c * math.sqrt(gamma_en * gamma_en - gamma_p * gamma_p)
            _, d2, d3 = _new_coords_python(p_after)
            phi = 2.0 * pi * next(rnd)
            cp, sp = math.cos(phi), math.sin(phi)
            return [a + delta_p * (b * cp + d * sp) for a, b, d in zip(p_after, d2, d3)]
        return p_scat

    if wr > 1.0 + 2.0 * eps:
        p5 = correction(w2 / w1, p1, p5, e1, e5, m1)
    elif wr < 1.0 - 2.0 * eps:
        p6 = correction(w1 / w2, p2, p6, e2, e6, m2)
    return p5, p6


@pytest.mark.parametrize("inter", [0, 1])
def test_sentoku_kemp_pair_equals_an_independent_restatement_bit_for_bit(inter):
    O.build()
    m1, m2 = M0, (1836.2 * M0 if inter else M0)
    q1, q2 = -Q0, (Q0 if inter else -Q0)
    n = 3000
    p1, p2, ran = _pairs(n, m1, m2, 3000.0, 41 + inter)
    rng = np.random.default_rng(9)
    w1 = np.where(rng.random(n) < 0.34, 3.0e10, np.where(rng.random(n) < 0.5, 2.0e10, 1.0e10))   # wr > 1, = 1, < 1
    w2 = np.full(n, 2.0e10)
    env = _env(m1, m2, q1, q2, 0, inter)
    a, b, done = _run_pairs(O.lib().orc_collide_pairs_test, p1, p2, w1, w2, ran, env)
    assert done.all()
    for i in range(n):
        r = _pair_sk_python(p1[i], p2[i], w1[i], w2[i], ran[i], m1, m2, q1, q2, env[4], env[5], env[6], env[7], env[8])
        assert np.array_equal(np.array(r[0]), a[i]) and np.array_equal(np.array(r[1]), b[i]), i
    assert len(set(w1)) == 3


def test_coulomb_log_equals_an_independent_restatement():
    """calc_coulomb_log (collisions.F90:1288-1316) in Python from the Fortran against the oracle's, bit for bit, over
    the floors (100 eV, 100 q0 J), the density cut-off, the classical and the quantum branch of bmin, and the floor
    of 1 on the result."""
    O.build()
    rng = np.random.default_rng(6)
    n = 20000
    a = np.empty((n, 7))
    a[:, 0] = 10.0 ** rng.uniform(-19, -12, n)          # ekbar1 [J]: below and above 100 eV
    a[:, 1] = 10.0 ** rng.uniform(0, 6, n)              # temp2 [eV]
    a[:, 2] = 10.0 ** rng.uniform(-1, 32, n)            # densities, some below the cut-off of 1 m^-3
    a[:, 3] = 10.0 ** rng.uniform(-1, 32, n)
    a[:, 4] = -Q0
    a[:, 5] = np.where(rng.random(n) < 0.5, Q0, -Q0) * rng.integers(1, 20, n)
    a[:, 6] = np.where(rng.random(n) < 0.5, M0, 1836.2 * M0)
    out = np.empty(n)
    O.lib().orc_coulomb_log(n, np.ascontiguousarray(a).ctypes.data, out.ctypes.data)
    h_bar, cc = 1.054571725336289397963133257349698e-34, CL * CL
    kinds = set()
    for i in range(n):
        ekbar1, temp2, dens1, dens2, q1, q2, m1 = (float(v) for v in a[i])
        local_ekbar1 = max(ekbar1, 100.0 * Q0)
        local_temp2 = max(temp2, 100.0)
        if dens1 <= 1.0 or dens2 <= 1.0:
            want, kind = 1.0, "thin"
        else:
            bmax = math.sqrt(D.epsilon0 * Q0 * local_temp2 / (abs(q2) * Q0 * dens2))
            b0 = abs(q1 * q2) / (8.0 * math.pi * D.epsilon0 * local_ekbar1)
            gamm = (local_ekbar1 / (m1 * cc)) + 1.0
            dB = 2.0 * math.pi * h_bar / (math.sqrt(gamm * gamm - 1.0) * m1 * CL)
            bmin = max(b0, dB)
            want = max(1.0, math.log(bmax / bmin))
            kind = ("floor" if want == 1.0 else "classical" if b0 > dB else "quantum")
        kinds.add(kind)
        assert out[i] == want, (i, a[i], out[i], want)
    assert kinds == {"thin", "floor", "classical", "quantum"}


def test_collision_step_equals_an_independent_restatement():
    """A whole collision step of one species on one rank, fixed Coulomb logarithm, once more in Python from the
    Fortran: reorder_particles_to_grid (split_particle.F90:29-77), calc_coll_number_density (:1320-1363), the
    Durstenfeld shuffle of every cell list on the rank's KISS stream (:1224-1284, cells in iy, ix order), then
    intra_collisions_np per cell (:446-646): the pair count with the odd particle wrapping round to the head of the
    list, the weight-sum factor, the per-cell constants, the pairs in list order drawing two random numbers each (none
    for a pair that is skipped).  The oracle's particles after collide() must be the same set, bit for bit."""
    from tests.test_window import _Stream
    O.build()
    n = (5, 4)
    dx = 1.0e-8
    ppc_mean = 7
    sp = [D.Species("electron", -Q0, M0, npart_per_cell=ppc_mean, density=1.0e28, temp=(0.0, 0.0, 0.0))]
    dk = D.Deck(2, list(n), [0.0, 0.0], [dx * n[0], dx * n[1]], ["periodic"] * 4, species=sp)
    o = O.Oracle(dk)                    # no auto_load: the KISS stream is untouched
    o.init()
    rng = np.random.default_rng(21)
    npart = ppc_mean * n[0] * n[1]
    p = np.zeros((npart, 6))
    p[:, 0] = rng.random(npart) * dx * n[0]          # uneven counts per cell, some of them odd
    p[:, 1] = rng.random(npart) * dx * n[1]
    p[:, 2:5] = rng.standard_normal((npart, 3)) * math.sqrt(M0 * 300.0 * Q0) * np.array([1.0, 0.3, 0.3])
    p[:, 5] = 1.0e28 * dx * dx / ppc_mean * rng.choice([1.0, 2.0], npart)
    o.set_particles(0, 0, p)
    log_lambda, user_factor = 10.0, 1.0
    o.collide([[user_factor]], coulomb_log=log_lambda, use_nanbu=True)
    got = o.get_particles(0, 0)

    info = o.rank_info(0)
    gmin = info["grid_min_local"]
    dt_coll = dk.dt() * 1.0
    g = _Stream(dk.seed + 0)
    cells = {}
    for P in p.copy():
        cx = math.floor((P[0] - gmin[0]) / dx + 1.5)
        cy = math.floor((P[1] - gmin[1]) / dx + 1.5)
        cells.setdefault((cx, cy), []).append(P)
    idx = 1.0 / dx / dx
    dens = {}
    for k, v in cells.items():      # summed in list order BEFORE the shuffle, then scaled (no sum(): it compensates)
        d = 0.0
        for P in v:
            d = d + float(P[5])
        dens[k] = d * idx
    order = [(ix, iy) for iy in range(1, n[1] + 1) for ix in range(1, n[0] + 1)]
    for k in order:                                   # shuffle_particle_list_random
        lst = cells.get(k, [])
        if len(lst) <= 2:
            continue
        for i in range(len(lst), 1, -1):
            sw = math.floor(i * g.random()) + 1
            lst[i - 1], lst[sw - 1] = lst[sw - 1], lst[i - 1]
    pi4_eps2_c4 = 4.0 * math.pi * (D.epsilon0 * D.epsilon0) * ((CL * CL) * (CL * CL))
    pi_fac = (4.0 * math.pi / 3.0) ** (1.0 / 3.0)

    class Draw:                                       # ran1, ran2 of a pair, drawn when the pair first asks
        def __init__(self):
            self.v = []

        def __getitem__(self, i):
            while len(self.v) <= i:
                self.v.append(g.random())
            return self.v[i]

    odd = 0
    for k in order:                                   # intra_collisions_np
        lst = cells.get(k, [])
        icount = len(lst)
        if icount <= 1:
            continue
        pcount = icount // 2 + icount % 2
        odd += icount % 2
        nxt = lambda i: (i + 1) % icount              # the list is circular while the routine runs
        factor, cur = 0.0, 0
        for _ in range(pcount):
            imp = nxt(cur)
            factor = factor + min(float(lst[cur][5]), float(lst[imp][5]))
            cur = nxt(imp)
        factor = user_factor / factor / 2.0
        d = dens[k]
        cell_fac = d * d * dt_coll * factor * dx * dx
        s_fac = cell_fac * log_lambda / pi4_eps2_c4
        dens_23 = d ** (2.0 / 3.0)
        s_fac_prime = cell_fac * pi_fac / dens_23
        cur = 0
        for _ in range(pcount):
            imp = nxt(cur)
            r = _pair_np_python(lst[cur][2:5], lst[imp][2:5], float(lst[cur][5]), float(lst[imp][5]), Draw(),
                                M0, M0, -Q0, -Q0, s_fac, s_fac_prime, max(M0, M0), 0)
            if r is not None:
                lst[cur][2:5], lst[imp][2:5] = r[0], r[1]
            cur = nxt(imp)
    want = np.array([P for k in order for P in cells.get(k, [])])
    assert odd >= 3
    assert got.shape == want.shape
    assert np.array_equal(sorted_rows(got), sorted_rows(want))
    moved = np.abs(sorted_rows(got)[:, 2:5] - sorted_rows(p)[:, 2:5]).max(axis=1) > 0
    assert moved.mean() > 0.5                      # and the step did scatter most particles
