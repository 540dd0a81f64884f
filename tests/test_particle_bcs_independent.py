"""particle_bcs (boundary.F90:1029-1462; the same text per axis in the three trees) once more, in Python from the
Fortran, for one rank: the candidate rule of the push (strict < / > against the local bounds, particles.F90:451-454),
the classification (< x_min_local, >= x_max_local), reflection about x_min / x_max with the momentum flipped, the
periodic shift by length_x, removal beyond x_min_outer / x_max_outer on open walls with the band in between left
alone.  After one push of a hot plasma the oracle's particles must be exactly this set, bit for bit."""
import itertools

import numpy as np
import pytest

from oracle.oracle import Oracle
from tests import decks
from tests.gpu_util import sorted_rows


def particle_bcs_python(p, nd, bc, xmin, xmax, min_local, max_local, min_outer, max_outer):
    out = []
    for P in p:
        P = [float(v) for v in P]
        # push_particles: only particles strictly outside the local bounds are looked at
        if not any(P[d] < min_local[d] or P[d] > max_local[d] for d in range(nd)):
            out.append(P)
            continue
        gone = False
        for d in range(nd):
            part_pos = P[d]
            length = xmax[d] - xmin[d]
            if part_pos < min_local[d]:
                b = bc[2 * d]
                if b == "reflect":
                    P[d] = 2.0 * xmin[d] - part_pos
                    P[nd + d] = -P[nd + d]
                elif b == "periodic":
                    P[d] = part_pos - (-1) * length
                elif part_pos < min_outer[d]:
                    gone = True
            if part_pos >= max_local[d]:
                b = bc[2 * d + 1]
                if b == "reflect":
                    P[d] = 2.0 * xmax[d] - part_pos
                    P[nd + d] = -P[nd + d]
                elif b == "periodic":
                    P[d] = part_pos - (+1) * length
                elif part_pos >= max_outer[d]:
                    gone = True
        if not gone:
            out.append(P)
    return np.array(out).reshape(-1, p.shape[1])


@pytest.mark.parametrize("ndims,n", [(1, (12,)), (2, (7, 6)), (3, (5, 4, 4))])
def test_particle_bcs_equals_an_independent_restatement(ndims, n):
    n_touched = n_seen = 0
    for kinds in itertools.product(("reflect", "periodic", "open"), repeat=ndims):
        bc = [k for k in kinds for _ in range(2)]
        dk = decks.thermal(ndims, n, ppc=12 if ndims < 3 else 6, temp_k=2.0e10, bc=bc)   # v_th dt / dx ~ 0.5: many leave
        o = Oracle(dk)
        o.auto_load()
        o.init()
        info = o.rank_info(0)
        mo, xo = o.outer()
        for step in range(3):
            o.push_only()
            before = o.get_particles(0, 0)
            want = particle_bcs_python(before, ndims, bc, dk.xmin, dk.xmax, info["min_local"], info["max_local"], mo, xo)
            o.particle_bcs()
            got = o.get_particles(0, 0)
            assert got.shape == want.shape, (kinds, step, got.shape, want.shape)
            assert np.array_equal(sorted_rows(got), sorted_rows(want)), (kinds, step)
            touched = (before[:, :ndims] < np.array(info["min_local"][:ndims])).any(axis=1) | \
                      (before[:, :ndims] >= np.array(info["max_local"][:ndims])).any(axis=1)
            n_touched += int(touched.sum())
            n_seen += len(touched)
            o.current_finish()
    assert n_touched >= 8 and n_touched > 0.005 * n_seen          # the cases are not idle: particles do cross the walls
