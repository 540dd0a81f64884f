"""Runs last (file name): the reference's custom-stencil decks replayed on the CUDA path against the numbers the
reference binary printed.  Added after the round's GPU minutes were spent, hence xfail(strict=False) until it has
been seen to pass on a B200; tests/test_oracle_golden.py pins the oracle on the same numbers."""
import numpy as np
import pytest

from epoch_b200 import deck as D

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(reason="added after the round's GPU minutes were spent: not yet run on a B200", strict=False)
@pytest.mark.parametrize("tree", ["1d_optimized", "1d_lehe_x", "2d_optimized"])
def test_custom_stencil_decks_reproduce_the_reference_binary_gpu(tree):
    """The reference's custom-stencil decks (simple_laser + open) on the CUDA path: the group velocities the
    reference binary printed (tests/test_oracle_golden.py pins the oracle on the same numbers).  The centroid is a
    ratio of sums over ~1e4 cells, so the device's own rounding shows at the 1e-10 level."""
    from epoch_b200.pic import Simulation
    from tests.test_oracle_golden import custom_stencil_deck, custom_stencil_deck_1d
    dk, recorded = {"1d_optimized": (custom_stencil_deck_1d("optimized"), 301440080.113),
                    "1d_lehe_x": (custom_stencil_deck_1d("lehe_x"), 310055314.605),
                    "2d_optimized": (custom_stencil_deck("optimized"), 314241436.846)}[tree]
    sim = Simulation(dk)
    tx = []
    x = dk.grid_min(0) + np.arange(dk.n[0]) * dk.dx(0)

    def dump(step, t):
        ey = sim.interior("ey").reshape(-1, dk.n[0])
        b = float(np.sum(ey ** 2))
        if b > 0 and t > 0:
            tx.append((t, float(np.sum(x[None, :] * ey ** 2) / b)))

    D.run(dk, sim, [0], dump)
    tx = np.array(tx)
    vg_sim = np.polyfit(tx[:, 0], tx[:, 1], 1)[0]
    assert np.isclose(vg_sim, recorded, rtol=1e-9, atol=0), vg_sim
