"""auto_load for a uniform-in-a-box species (user_interaction/helper.F90:371-808: load_particles :540-583,
setup_particle_density :680-757 with include/particle_to_grid.inc and triangle/gxfac.inc, then the x / y / z momentum
passes of setup_particle_temperature, particle_temperature.F90:30-81, :388-398) once more, in Python from the Fortran
on the rank's KISS stream, against the oracle's loader: every particle bit for bit.  The slab starts and ends inside
the box, so the density-map fall-back of the weight interpolation (cells outside the slab take the nearer valid
cell's density) is exercised."""
import math

import numpy as np

from epoch_b200 import deck as D
from oracle.oracle import Oracle
from tests import decks
from tests.test_window import _Stream


def test_loader_equals_an_independent_restatement():
    dk = decks.thermal(2, (12, 9), ppc=5, temp_k=3.0e7, drift=(2.0e-24, 0.0, -1.0e-24))
    s = dk.species[0]
    s.temp = (3.0e7, 1.0e7, 2.0e7)
    dx, dy = dk.dx(0), dk.dx(1)
    s.box_lo = (2.6 * dx, -1e300, -1e300)      # cells 4 .. 9 of 12 hold plasma
    s.box_hi = (9.4 * dx, 1e300, 1e300)
    o = Oracle(dk)
    o.auto_load()
    got = o.get_particles(0, 0)
    nx, ny, ng = dk.n[0], dk.n[1], 5
    g = _Stream(dk.seed + 0)
    xc = lambda i: float(dk.x_global(0, i))
    yc = lambda j: float(dk.x_global(1, j))
    gminx, gminy = xc(1), yc(1)
    # density(1-ng:nx+ng, 1-ng:ny+ng) from the deck function, field_bc (periodic: the ghost cells are images)
    density = {}
    for j in range(1 - ng, ny + ng + 1):
        for i in range(1 - ng, nx + ng + 1):
            gi = (i - 1) % nx + 1
            v = s.density if s.box_lo[0] <= xc(gi) < s.box_hi[0] else 0.0
            density[(i, j)] = v if v >= 2.220446049250313e-16 else 0.0
    dmap = {k: v > 0.0 for k, v in density.items()}
    parts = []
    for iy in range(1, ny + 1):
        for ix in range(1, nx + 1):
            if not dmap[(ix, iy)]:
                continue
            for _ in range(int(s.npart_per_cell)):
                x = xc(ix) + (g.random() - 0.5) * dx
                y = yc(iy) + (g.random() - 0.5) * dy
                parts.append([x, y, 0.0, 0.0, 0.0, 0.0])

    def to_grid(P):
        cell_x_r = (P[0] - gminx) / dx
        cell_y_r = (P[1] - gminy) / dy
        cell_x, cell_y = math.floor(cell_x_r + 0.5), math.floor(cell_y_r + 0.5)
        cfx, cfy = float(cell_x) - cell_x_r, float(cell_y) - cell_y_r
        cx2, cy2 = cfx * cfx, cfy * cfy
        gx = {-1: 0.5 * (0.25 + cx2 + cfx), 0: 0.75 - cx2, 1: 0.5 * (0.25 + cx2 - cfx)}
        gy = {-1: 0.5 * (0.25 + cy2 + cfy), 0: 0.75 - cy2, 1: 0.5 * (0.25 + cy2 - cfy)}
        return cell_x + 1, cell_y + 1, gx, gy

    tz = lambda a: int(a / 2)
    count = {}
    for P in parts:
        cell_x, cell_y, gx, gy = to_grid(P)
        wdata = 0.0
        for isuby in (-1, 0, 1):
            i, j = cell_x, cell_y + isuby
            if not dmap[(i, j)]:
                j = cell_y + tz(isuby)
            for isubx in (-1, 0, 1):
                i = cell_x + isubx
                if not dmap[(i, j)]:
                    i = cell_x + tz(isubx)
                wdata = wdata + gx[isubx] * gy[isuby] * density[(i, j)]
        P[5] = wdata
        count[(cell_x, cell_y)] = count.get((cell_x, cell_y), 0) + 1
    wdata = dx * dy
    for P in parts:
        cell_x = math.floor((P[0] - gminx) / dx + 1.5)
        cell_y = math.floor((P[1] - gminy) / dy + 1.5)
        P[5] = P[5] * wdata / count[(cell_x, cell_y)]
    for direction in range(3):
        for P in parts:
            _, _, gx, gy = to_grid(P)
            temp_local = drift_local = 0.0
            for isuby in (-1, 0, 1):
                for isubx in (-1, 0, 1):
                    temp_local = temp_local + gx[isubx] * gy[isuby] * s.temp[direction]
                    drift_local = drift_local + gx[isubx] * gy[isuby] * s.drift[direction]
            P[2 + direction] = g.box_muller(math.sqrt(temp_local * D.kb * s.mass), drift_local)
    want = np.array(parts)
    assert got.shape == want.shape == (6 * ny * 5, 6)
    assert np.array_equal(got, want)
    edge = (want[:, 0] < s.box_lo[0] + 0.5 * dx) | (want[:, 0] > s.box_hi[0] - 0.9 * dx)
    assert edge.any()
