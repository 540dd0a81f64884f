"""Oracle parity ON THE BENCH CONFIGURATION: BASELINE C2 physics -- uniform thermal
electrons, dx = dy = Debye length, periodic, 64 particles per cell from EPOCH's KISS loader -- at 512 x 512
cells (16.7 M particles, 2048 tiles of the cell-owner kernel), with the library settings bench.py uses:
performance build (strict_fp = 0) and the default sort interval, next to the parity build.  Ten steps; bars as
BASELINE.json north_star: particle counts per cell bit-exact, E/B/J within 1e-12 relative L2 (checked after
steps 1, 5 and 10).  At 64 ppc every lane of the cell-owner kernel runs >= 64 rounds through its software
pipeline and the sort machinery of the full-size run: a regime the small decks never enter.

Plus: the same deck started from the bench's device loader (epb_load_uniform) checked on moments."""
import math

import numpy as np
import pytest

from tests.gpu_util import FIELDS, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12
N, PPC = 512, 64


def _deck(n=N, ppc=PPC):
    import bench
    return bench.c2_deck(n, ppc, (1, 1))


def test_bench_configuration_matches_oracle():
    from epoch_b200.pic import Simulation
    from oracle.oracle import Oracle
    dk = _deck()
    o = Oracle(dk)
    o.auto_load()
    p0 = o.get_particles(0, 0)
    assert p0.shape[0] == N * N * PPC
    sims = {"strict": Simulation(dk, strict_fp=True, sort_interval=0, capacity_factor=1.5),
            "fast": Simulation(dk, strict_fp=False, sort_interval=0, capacity_factor=1.5)}
    for s in sims.values():
        s.upload_species(0, p0)
    del p0
    o.init()
    for s in sims.values():
        s.init()
    for step in range(1, 11):
        o.fields_half(); o.push(); o.current_finish(); o.fields_final()
        for s in sims.values():
            s.step()
        if step in (1, 5, 10):
            ref = {f: o.field(0, f) for f in FIELDS}
            cnt = o.cell_counts(0, 0)
            for nm, s in sims.items():
                for f in FIELDS:
                    e = rel_l2(s.download_field(f), ref[f])
                    assert e <= TOL, (nm, step, f, e)
                assert s.count(0) == N * N * PPC
                assert np.array_equal(s.cell_counts(0), cnt), (nm, step)
    # particle state after 10 steps: 1e-9 of each column's scale (J differs at 1e-16 from step 1 on)
    b = o.get_particles(0, 0)
    b = b[np.lexsort(tuple(b[:, k] for k in range(b.shape[1] - 1, -1, -1)))]
    for nm, s in sims.items():
        a = s.download_species(0)
        a = a[np.lexsort(tuple(a[:, k] for k in range(a.shape[1] - 1, -1, -1)))]
        assert a.shape == b.shape
        assert np.all(np.abs(a - b) <= 1e-9 * np.max(np.abs(b), axis=0)), nm


def test_bench_loader_moments():
    """bench.py's device loader (epb_load_uniform stands in for auto_load): npart_per_cell particles in every
    cell, weights that reproduce the deck density, a Maxwellian at the deck temperature; and ten steps of the
    performance build conserve the particle count and keep the moments."""
    from epoch_b200 import deck as D
    from epoch_b200.pic import Simulation
    dk = _deck(256, 64)
    sim = Simulation(dk, strict_fp=False, sort_interval=0, capacity_factor=1.2)
    sim.load_uniform(0, seed=20261017)
    n0 = 256 * 256 * 64
    assert sim.count(0) == n0
    assert np.all(sim.cell_counts(0) == 64)
    sim.init()
    ng = 5
    nd0 = sim.moment("number_density", 0)[0, ng:-ng, ng:-ng]
    assert abs(nd0.mean() / dk.species[0].density - 1.0) < 1e-12
    t0 = sim.moment("temperature", 0)[0, ng:-ng, ng:-ng].mean()
    assert abs(t0 / dk.species[0].temp[0] - 1.0) < 0.02
    ke0 = sim.kinetic_energy(0)
    for _ in range(10):
        sim.step()
    assert sim.count(0) == n0 and int(sim.cell_counts(0).sum()) == n0
    nd1 = sim.moment("number_density", 0)[0, ng:-ng, ng:-ng]
    assert abs(nd1.mean() / dk.species[0].density - 1.0) < 1e-12
    fe, fb = sim.field_energy()
    ke1 = sim.kinetic_energy(0)
    # dx = Debye length: a small part of the thermal energy moves into the fields, the total is kept
    assert abs((ke1 + fe + fb) / ke0 - 1.0) < 1e-3
    assert math.isfinite(fe) and fe > 0
