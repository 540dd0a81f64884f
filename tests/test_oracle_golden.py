"""Pins the oracle on every golden value the reference's tests hold for the hot path
(SURVEY.md §8c): sum(Ey^2) / sum(Ex^2) of the laser decks, np.isclose default rtol=1e-5."""
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
