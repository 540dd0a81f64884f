"""Runs last (file name).  All of these passed on the driver's B200 box in round 1 (GPUTEST_r01: xpassed), so they
are plain tests now: the reference's custom-stencil decks replayed on the CUDA path against the numbers
the reference binary printed (tests/test_oracle_golden.py pins the oracle on the same numbers), and the device's
energy diagnostics / total energy history (tests/test_energy_history.py does it for the oracle)."""
import numpy as np
import pytest

from epoch_b200 import deck as D

pytestmark = pytest.mark.gpu


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


def test_energy_diagnostics_and_history_gpu():
    """epb_field_energy / epb_kinetic_energy (calc_total_energy_sum, io/calc_df.F90:1321-1417) against numpy on the
    same device state, and the total energy of the CUDA path alone over 300 steps (performance build)."""
    from epoch_b200.pic import Simulation
    from oracle.oracle import Oracle
    from tests import decks
    from tests.test_energy_history import energies
    dk = decks.thermal(2, (24, 24), ppc=16, temp_k=1.0e7, two_species=True)
    o = Oracle(dk)
    o.auto_load()                                   # the loader only
    sim = Simulation(dk, strict_fp=False, sort_interval=0, capacity_factor=1.5)
    for isp in range(2):
        sim.upload_species(isp, o.get_particles(0, isp))
    sim.init()
    tot = []
    for s in range(300):
        sim.step()
        if s % 30 == 29:
            fe, fb = sim.field_energy()
            ke = [sim.kinetic_energy(isp) for isp in range(2)]
            rfe, rfb, rke = energies(dk, {f: sim.interior(f) for f in ("ex", "ey", "ez", "bx", "by", "bz")},
                                     [sim.download_species(isp) for isp in range(2)])
            assert np.isclose(fe, rfe, rtol=1e-12) and np.isclose(fb, rfb, rtol=1e-12)
            assert np.allclose(ke, rke, rtol=1e-12)
            tot.append(fe + fb + sum(ke))
    tot = np.array(tot)
    assert np.abs(tot / tot[0] - 1.0).max() < 1.0e-4


@pytest.mark.parametrize("ndims,n", [(1, (96,)), (2, (40, 24)), (3, (18, 9, 7))])
def test_append_species_gpu(ndims, n):
    """epb_append_species (what the shim calls after run_injectors / insert_particles): a species uploaded in two
    halves with some steps in between the appends behaves as the oracle with the same particles -- after the append
    the device holds exactly the union, and the following steps agree in counts per cell and in the fields."""
    from epoch_b200.pic import Simulation
    from oracle.oracle import Oracle
    from tests import decks
    from tests.gpu_util import FIELDS, rel_l2, sorted_rows
    dk = decks.thermal(ndims, n, ppc=6, temp_k=2.0e8)
    o = Oracle(dk)
    o.auto_load()
    p = o.get_particles(0, 0).copy()
    half = p.shape[0] // 2
    o.set_particles(0, 0, p[:half])
    sim = Simulation(dk, strict_fp=True, sort_interval=2, capacity_factor=3.0)
    sim.upload_species(0, p[:half])
    o.init(); sim.init()

    def step():
        o.fields_half(); sim.fields_half()
        o.push(); sim.push()
        o.current_finish(); sim.current_finish()
        o.fields_final(); sim.fields_final()

    for _ in range(3):
        step()
    q = o.get_particles(0, 0).copy()
    o.set_particles(0, 0, np.concatenate([q, p[half:]]))       # append_partlist
    sim.append_species(0, p[half:])
    assert sim.count(0) == p.shape[0] == o.count(0, 0)
    a, b = sorted_rows(sim.download_species(0)), sorted_rows(o.get_particles(0, 0))
    # three steps in: the summation order of J differs between the two, so the old half agrees to round-off of each
    # column's scale; the appended half is on the device exactly as it was handed over
    assert np.all(np.abs(a - b) <= 1e-9 * np.abs(b).max(axis=0))
    rows = {r.tobytes() for r in a}
    assert all(r.tobytes() in rows for r in p[half:])
    for _ in range(4):
        step()
    assert np.array_equal(sim.cell_counts(0), o.cell_counts(0, 0))
    for name in FIELDS:
        assert rel_l2(sim.download_field(name), o.field(0, name)) <= 1e-12, name


@pytest.mark.parametrize("ndims,n", [(2, (40, 24)), (3, (18, 9, 7))])
def test_step_scalars_async_gpu(ndims, n):
    """epb_step_scalars_async (the pipelined form of update_particle_count + calc_total_energy_sum): the numbers
    that arrive in page-locked memory one step late are exactly those of the blocking calls made at that step,
    and source planes handed to epb_set_laser_source may be overwritten by the caller right after the call."""
    import torch
    from epoch_b200.pic import Simulation
    from oracle.oracle import Oracle
    from tests import decks
    dk = decks.thermal(ndims, n, ppc=6, temp_k=2.0e8, two_species=True)
    o = Oracle(dk)
    o.auto_load()
    sim = Simulation(dk, strict_fp=False, sort_interval=2, capacity_factor=2.0)
    for isp in range(2):
        sim.upload_species(isp, o.get_particles(0, isp))
    sim.init()
    bufs = [torch.zeros(8, dtype=torch.float64).pin_memory() for _ in range(2)]
    blocking, tickets = [], []
    for k in range(6):
        sim.step()
        tickets.append(sim.step_scalars_async(bufs[k % 2].data_ptr()))
        blocking.append(list(sim.field_energy()) + [float(sim.global_count(isp)) for isp in range(2)])
        if k > 0:                                    # the previous step's numbers, while this step is in flight
            sim.wait_scalars(tickets[k - 1])
        sim.wait_scalars(tickets[k])
        got = bufs[k % 2][:5].tolist()
        assert got[4] == 0.0                         # device error word
        assert got[2:4] == blocking[k][2:4]
        assert np.allclose(got[:2], blocking[k][:2], rtol=1e-13, atol=0.0)
    assert blocking[-1][2] == o.get_particles(0, 0).shape[0]


@pytest.mark.parametrize("ndims,n", [(1, (64,)), (2, (48, 32)), (3, (12, 10, 9))])
def test_load_profile_gpu(ndims, n):
    """epb_load_profile (get_load_x/y/z, balance.F90:1766-1844) against the numpy restatement in deck.py on the
    downloaded particles, then calculate_breaks on it."""
    from epoch_b200.pic import Simulation
    from oracle.oracle import Oracle
    from tests import decks
    dk = decks.thermal(ndims, n, ppc=5, temp_k=3.0e8, two_species=True)
    o = Oracle(dk)
    o.auto_load()
    sim = Simulation(dk, strict_fp=True, sort_interval=2, capacity_factor=2.0)
    for isp in range(2):
        sim.upload_species(isp, o.get_particles(0, isp))
    sim.init()
    for _ in range(3):
        sim.step()
    parts = [sim.download_species(isp) for isp in range(2)]
    for axis in range(ndims):
        cells = np.concatenate([np.floor((p[:, axis] - dk.grid_min(axis)) / dk.dx(axis) + 1.5) for p in parts])
        other = int(np.prod([n[d] for d in range(ndims) if d != axis]))
        ref = D.load_profile(cells.astype(np.int64), n[axis], other)
        got = sim.load_profile(axis)
        assert np.array_equal(got, ref), axis
        mins, maxs = D.calculate_breaks(got, 4)
        assert mins[0] == 1 and maxs[-1] == n[axis]
