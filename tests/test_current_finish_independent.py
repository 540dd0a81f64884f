"""current_finish (housekeeping/current_smooth.F90:29-45 without smoothing) once more, in numpy straight from
boundary.F90 -- particle_reflection_bcs :534-630 with its asymmetric folds (1 .. ng-1 at the lower wall, 1 .. ng at
the upper one; the component normal to the wall changes sign and pairs i with -i instead of 1-i),
particle_periodic_bcs :634-751 (ghost strips of the whole transverse extent, x before y before z) and field_bc(j)
:145-315 -- against the oracle on random currents, bit for bit, one rank, every mix of reflecting and periodic
walls.  SURVEY.md 8 lists these folds among the places where a natural rewrite silently diverges."""
import itertools

import numpy as np
import pytest

from oracle.oracle import Oracle
from tests import decks

NG = 5


def _ax(a, d, nd):
    """view with Fortran dimension d (0 = x) first; arrays are stored [z][y][x]"""
    return np.moveaxis(a, 2 - d, 0)


def current_finish_numpy(j3, n, nd, bc):
    """j3: [jx, jy, jz] arrays with ghosts (modified in place); n: cells per axis; bc: 'reflect' / 'periodic' per side"""
    ng = NG
    for comp, a in enumerate(j3):
        # particle_reflection_bcs(array, ng, flip_direction = comp + 1)
        for d in range(nd):
            v = _ax(a, d, nd)
            nn = n[d]
            F = lambda i: i + ng - 1          # Fortran index -> 0-based
            if bc[2 * d] == "reflect":
                for i in range(1, ng):
                    if comp == d:
                        v[F(i)] = v[F(i)] - v[F(-i)]
                        v[F(-i)] = 0.0
                    else:
                        v[F(i)] = v[F(i)] + v[F(1 - i)]
                        v[F(1 - i)] = 0.0
            if bc[2 * d + 1] == "reflect":
                for i in range(1, ng + 1):
                    if comp == d:
                        v[F(nn - i)] = v[F(nn - i)] - v[F(nn + i)]
                        v[F(nn + i)] = 0.0
                    else:
                        v[F(nn + 1 - i)] = v[F(nn + 1 - i)] + v[F(nn + i)]
                        v[F(nn + i)] = 0.0
        # particle_periodic_bcs: one rank, so a periodic axis receives its own strips and any other axis zeros
        for d in range(nd):
            v = _ax(a, d, nd)
            nn = n[d]
            F = lambda i: i + ng - 1
            per = bc[2 * d] == "periodic"
            temp = v[F(nn + 1):F(nn + ng) + 1].copy() if per else 0.0
            v[F(1):F(ng) + 1] = v[F(1):F(ng) + 1] + temp
            temp = v[F(1 - ng):F(0) + 1].copy() if per else 0.0
            v[F(nn + 1 - ng):F(nn) + 1] = v[F(nn + 1 - ng):F(nn) + 1] + temp
    for a in j3:
        # field_bc(j, jng): ghost cells of a periodic axis from the other end, x then y then z
        for d in range(nd):
            if bc[2 * d] != "periodic":
                continue
            v = _ax(a, d, nd)
            nn = n[d]
            F = lambda i: i + ng - 1
            v[F(nn + 1):F(nn + ng) + 1] = v[F(1):F(ng) + 1]
            v[F(1 - ng):F(0) + 1] = v[F(nn + 1 - ng):F(nn) + 1]
    return j3


@pytest.mark.parametrize("ndims,n", [(1, (17,)), (2, (13, 11)), (3, (9, 8, 7))])
def test_current_finish_equals_an_independent_restatement(ndims, n):
    for kinds in itertools.product(("reflect", "periodic"), repeat=ndims):
        bc = [k for k in kinds for _ in range(2)]
        dk = decks.thermal(ndims, n, ppc=1, bc=bc)
        o = Oracle(dk)
        o.init()
        rng = np.random.default_rng(hash(kinds) % 1000)
        j3 = []
        for name in ("jx", "jy", "jz"):
            a = o.field(0, name)
            a[...] = rng.standard_normal(a.shape)
            j3.append(a.copy())
        current_finish_numpy(j3, list(n) + [1] * (3 - ndims), ndims, bc)
        o.current_finish()
        for name, mine in zip(("jx", "jy", "jz"), j3):
            assert np.array_equal(o.field(0, name), mine), (kinds, name)
