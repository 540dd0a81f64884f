"""Pins the oracle on every golden value the reference's tests hold for the hot path
(SURVEY.md §8c): sum(Ey^2) / sum(Ex^2) of the laser decks, np.isclose default rtol=1e-5."""
import os

import numpy as np
import pytest

from epoch_b200 import deck as D
from oracle.oracle import Oracle, kiss
from tests import decks


def _run(dk, fld, max_dumps=None):
    o = Oracle(dk)
    res = []

    class Stop(Exception):
        pass

    def dump(step, t):
        res.append(sum(float(np.sum(o.interior(r, fld) ** 2)) for r in range(o.nranks)))
        if max_dumps is not None and len(res) >= max_dumps:
            raise Stop

    try:
        D.run(dk, o, list(range(o.nranks)), dump)
    except Stop:
        pass
    return res


def test_laser1d_golden():
    # epoch1d/tests/test_laser.py:70-80
    res = _run(decks.laser1d(), "ey")
    assert len(res) == 8
    assert res[0] == 0.0
    assert np.isclose(res[1], 1.38636e+23)
    assert np.isclose(res[3], 1.40618e+23)
    assert np.isclose(res[7], 6.90067e+17)


def test_laser1d_golden_two_ranks():
    res = _run(decks.laser1d(nproc=(2, 1, 1)), "ey")
    assert np.isclose(res[1], 1.38636e+23) and np.isclose(res[7], 6.90067e+17)


def test_laser2d_golden():
    # epoch2d/tests/test_laser.py:70-77
    res = _run(decks.laser2d(), "ey")
    assert res[0] == 0.0
    assert np.isclose(res[1], 7.55007e+25)
    assert np.isclose(res[2], 1.51319e+26)


def test_laser2d_golden_decomposed():
    res = _run(decks.laser2d(nproc=(2, 3, 1)), "ey", max_dumps=2)
    assert np.isclose(res[1], 7.55007e+25)


def test_laser3d_golden_first_dump():
    # epoch3d/tests/test_laser.py:70-74 (dump 0002 is covered by the slow test)
    res = _run(decks.laser3d(), "ex", max_dumps=2)
    assert res[0] == 0.0
    assert np.isclose(res[1], 3.89491e+25)


@pytest.mark.parametrize("face,fld", [("y", "ey"), ("z", "ez")])
def test_laser3d_golden_on_other_faces(face, fld):
    """The same golden value with the laser on y_min / z_min (axes permuted): pins outflow_bcs_{y,z}_*."""
    res = _run(decks.laser3d_face(face), fld, max_dumps=2)
    assert res[0] == 0.0
    assert np.isclose(res[1], 3.89491e+25)


@pytest.mark.slow
def test_laser3d_golden_full():
    res = _run(decks.laser3d(nproc=(2, 2, 2)), "ex")
    assert np.isclose(res[1], 3.89491e+25)
    assert np.isclose(res[2], 7.78759e+25)


def test_kiss_stream():
    # random_generator.f90:45-78: uniform in [0,1), reproducible, rank-dependent seed
    a = kiss(7842432, 1000)
    b = kiss(7842432, 1000)
    c_ = kiss(7842433, 1000)
    assert np.array_equal(a, b) and not np.array_equal(a, c_)
    assert a.min() >= 0.0 and a.max() < 1.0
    assert abs(a.mean() - 0.5) < 0.05


# epoch2d/tests/custom_stencils/{optimized,optimized_symm,optimized_xaxis}/input.deck: stencil blocks
CUSTOM_STENCILS = {
    "optimized": dict(dt=0.9082126568805592, betaxy=0.04075757835916255, betayx=0.04075757835916255,
                      deltax=-0.04142032920970152, deltay=-0.20827814817872584, vg=1.0490493627815458),
    "optimized_symm": dict(dt=0.8988685682513151, betaxy=0.01862292597327679, betayx=0.01862292597327679,
                           deltax=-0.04155873453935287, deltay=-0.04155873453935287, vg=1.044753207834214),
    "optimized_xaxis": dict(dt=0.956632159129662, betaxy=0.025096871992206993, betayx=0.025096871992206993,
                            deltax=-0.017744324957063393, deltay=-0.0009692545471922645, vg=1.0197513694119302),
}


def custom_stencil_deck(name):
    """The reference's custom-stencil test deck as it is: 240 x 80 cells over 24 um x 24 um, simple_laser on x_min,
    open on x_max, periodic in y, lambda = 0.5 um, gauss(y, 0, 4 um) x gauss(t, 8 fs, 1.8 fs), t_end = 75 fs,
    dumps every 25 fs, maxwell_solver = custom with dt from the stencil block and dt_multiplier = 1."""
    st = CUSTOM_STENCILS[name]
    nx, ny = 240, 80
    L = 12 * D.micron
    dx = 2 * L / nx
    lam = 0.5 * D.micron
    las = D.Laser("x_min", D.Laser.amp_from_intensity_w_cm2(1.0e15), 2 * D.pi * D.c / lam,
                  profile=lambda y, z: D.gauss(y, 0, 4 * D.micron),
                  t_profile=lambda t: D.gauss(t, 8 * D.femto, 1.8 * D.femto))
    custom = dict(betaxy=st["betaxy"], betayx=st["betayx"], deltax=st["deltax"], deltay=st["deltay"],
                  dt=st["dt"] * dx / D.c)
    return D.Deck(2, [nx, ny], [-L, -L], [L, L], ["simple_laser", "open", "periodic", "periodic"], lasers=[las],
                  t_end=75 * D.femto, dt_snapshot=25 * D.femto, dt_multiplier=1.0,
                  maxwell_solver="custom", stencil_custom=custom)


@pytest.mark.parametrize("name", sorted(CUSTOM_STENCILS))
def test_custom_stencil_group_velocity_golden(name):
    """epoch2d/tests/test_custom_stencils.py:36-45 (golden vg per stencil), :62-72 (xt2: centroid of Ey^2 over the
    whole box), :158-176: the slope of the centroid over dumps 1..3 equals vg to rtol 0.003 (the reference's own
    runs: 0.0008 / 0.0020 / 0.0008)."""
    dk = custom_stencil_deck(name)
    o = Oracle(dk)
    tx = []
    x = dk.grid_min(0) + np.arange(dk.n[0]) * dk.dx(0)       # grid_mid of Ey along x: cell centres

    def dump(step, t):
        ey = o.interior(0, "ey").reshape(dk.n[1], dk.n[0])
        b = float(np.sum(ey ** 2))
        if b > 0 and t > 0:
            tx.append((t, float(np.sum(x[None, :] * ey ** 2) / b)))

    D.run(dk, o, [0], dump)
    assert len(tx) == 3
    tx = np.array(tx)
    vg_sim = np.polyfit(tx[:, 0], tx[:, 1], 1)[0]
    assert np.isclose(vg_sim, CUSTOM_STENCILS[name]["vg"] * D.c, rtol=0.003), (name, vg_sim / D.c)
    if name == "optimized":
        # the number the reference binary itself printed for this deck (test_custom_stencils.py:170, "optimized
        # 314241436.846"; the deck the reference's test runs by default): reproduced to all twelve digits.  The
        # comments for the two other decks (312578029.167, 305472651.829) predate their current coefficients: the
        # oracle gives 312921089.277 and 305670732.614, still inside the asserted rtol.
        assert np.isclose(vg_sim, 314241436.846, rtol=2e-12, atol=0), vg_sim


def custom_stencil_deck_1d(name):
    """epoch1d/tests/custom_stencils/{optimized,lehe_custom,lehe_x}/input.deck: 240 cells over 24 um, simple_laser /
    open, lambda = 0.5 um, gauss(t, 8 fs, 1.8 fs) until 14 fs, t_end = 75 fs, dumps every 12 fs."""
    nx = 240
    L = 12 * D.micron
    dx = 2 * L / nx
    las = D.Laser("x_min", D.Laser.amp_from_intensity_w_cm2(1.0e15), 2 * D.pi * D.c / (0.5 * D.micron),
                  t_profile=lambda t: D.gauss(t, 8 * D.femto, 1.8 * D.femto), t_end=14 * D.femto)
    kw = dict(lasers=[las], t_end=75 * D.femto, dt_snapshot=12 * D.femto)
    if name == "lehe_x":
        return D.Deck(1, [nx], [-L], [L], ["simple_laser", "open"], maxwell_solver="lehe_x", **kw)
    delta = {"optimized": -0.013364149548965119, "lehe_custom": -0.025303094265254511}[name]
    return D.Deck(1, [nx], [-L], [L], ["simple_laser", "open"], maxwell_solver="custom", dt_multiplier=1.0,
                  stencil_custom=dict(deltax=delta, dt=0.95 * dx / D.c), **kw)


@pytest.mark.parametrize("name,recorded,vg_theory", [
    ("optimized", 301440080.113, 1.0062495084969005 * D.c),
    ("lehe_custom", 310055314.605, None),
    ("lehe_x", 310055314.605, None)])
def test_custom_stencil_1d_reproduces_the_reference_binary(name, recorded, vg_theory):
    """epoch1d/tests/test_custom_stencils.py:121-135: slope of the Ey^2 centroid over dumps 1..7 against the
    theoretical group velocity (rtol 0.006) - and against the numbers the reference binary printed for these
    three decks (:127-129), which the oracle reproduces to all twelve digits."""
    dk = custom_stencil_deck_1d(name)
    o = Oracle(dk)
    tx = []
    x = dk.grid_min(0) + np.arange(dk.n[0]) * dk.dx(0)

    def dump(step, t):
        ey = o.interior(0, "ey").reshape(-1)
        b = float(np.sum(ey ** 2))
        if b > 0 and t > 0:
            tx.append((t, float(np.sum(x * ey ** 2) / b)))

    D.run(dk, o, [0], dump)
    assert len(tx) == 7
    tx = np.array(tx)
    vg_sim = np.polyfit(tx[:, 0], tx[:, 1], 1)[0]
    assert np.isclose(vg_sim, recorded, rtol=5e-12, atol=0), (name, vg_sim)
    if vg_theory is None:   # test_custom_stencils.py:53: vg_lehe with dt = 0.95 dx / c
        dx = dk.dx(0)
        k_l = 2 * np.pi / (0.5 * D.micron)
        vg_theory = D.c * (1.0 + 2.0 * (1.0 - 0.95) * (k_l * dx / 2.0) ** 2)
    assert np.isclose(vg_sim, vg_theory, rtol=0.006)


@pytest.mark.slow
@pytest.mark.skipif(not os.environ.get("EPB_RUN_SLOW"), reason="136 s on one core: set EPB_RUN_SLOW=1 (passes, see DESIGN.md 2)")
def test_custom_stencil_3d_reproduces_the_reference_binary():
    """epoch3d/tests/custom_stencils/optimized/input.deck (240 x 80 x 80 cells), epoch3d/tests/
    test_custom_stencils.py:84, :207-213: vg = 1.0713226616321112 c to rtol 0.003, and the number the reference
    binary printed, 320424292.475, to all twelve digits (136 s on one core, hence `slow`)."""
    st = dict(dt=0.8661145061674279, betaxy=0.03739407994958153, betaxz=0.03739457628015302,
              betayx=0.03739407994958153, betayz=0.016829085115575494, betazx=0.03739457628015302,
              betazy=0.016829085115575494, deltax=-0.06126939775259932, deltay=-0.209906582294693,
              deltaz=-0.20990862719271147)
    nx, ny = 240, 80
    L = 12 * D.micron
    st["dt"] = st["dt"] * (2 * L / nx) / D.c
    las = D.Laser("x_min", D.Laser.amp_from_intensity_w_cm2(1.0e15), 2 * D.pi * D.c / (0.5 * D.micron),
                  profile=lambda y, z: D.gauss(np.sqrt(y * y + z * z), 0, 4 * D.micron),
                  t_profile=lambda t: D.gauss(t, 8 * D.femto, 1.8 * D.femto))
    dk = D.Deck(3, [nx, ny, ny], [-L] * 3, [L] * 3, ["simple_laser", "open"] + ["periodic"] * 4, lasers=[las],
                t_end=75 * D.femto, dt_snapshot=25 * D.femto, dt_multiplier=1.0, maxwell_solver="custom",
                stencil_custom=st)
    o = Oracle(dk)
    tx = []
    x = dk.grid_min(0) + np.arange(nx) * dk.dx(0)

    def dump(step, t):
        ey = o.interior(0, "ey")
        b = float(np.sum(ey ** 2))
        if b > 0 and t > 0:
            tx.append((t, float(np.sum(x[None, None, :] * ey ** 2) / b)))

    D.run(dk, o, [0], dump)
    tx = np.array(tx)
    vg_sim = np.polyfit(tx[:, 0], tx[:, 1], 1)[0]
    assert np.isclose(vg_sim, 1.0713226616321112 * D.c, rtol=0.003)
    assert np.isclose(vg_sim, 320424292.475, rtol=5e-12, atol=0), vg_sim
