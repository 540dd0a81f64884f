"""One decomposition rank of a multi-rank parity case: the CUDA path of THIS rank against the
multi-rank CPU oracle (which every process runs in full, it is deterministic).  Bars: per-cell
and per-rank particle counts bit-exact after exchange, global counts exact, E/B/J (and the
diagnostics moments) within 1e-12 relative L2 (1e-11 for the 30-step open-boundary foil deck).

Used by tests/multi_worker.py (pytest, one process per GPU) and by bench.py's untimed
`parity_check` pre-phase, so that the driver's scaling runs carry a multi-rank parity verdict.
The oracle is test infrastructure: it checks, it is never the thing measured."""
import numpy as np

from epoch_b200 import deck as D
from tests import decks
from tests.gpu_util import FIELDS, rel_l2


def make_deck(name, world):
    if name == "thermal2d_x":
        return decks.thermal(2, (48, 40), ppc=6, temp_k=3.0e8, nproc=(world, 1, 1), two_species=True), 10, 1e-12
    if name == "thermal2d_y":
        return decks.thermal(2, (40, 48), ppc=6, temp_k=3.0e8, nproc=(1, world, 1)), 10, 1e-12
    if name == "thermal2d_xy":
        return decks.thermal(2, (48, 48), ppc=5, temp_k=3.0e8, nproc=(2, world // 2, 1)), 10, 1e-12
    if name == "thermal2d_bench":   # the bench's decomposition rule (split_domain) on a hot plasma
        px = {1: 1, 2: 1, 4: 2, 8: 2}[world]
        return decks.thermal(2, (32 * px, 32 * (world // px)), ppc=8, temp_k=3.0e8, nproc=(px, world // px, 1)), 10, 1e-12
    if name == "thermal3d":
        npz = (2, 2, world // 4) if world >= 4 else (world, 1, 1)
        return decks.thermal(3, (20, 16, 12), ppc=4, temp_k=3.0e8, nproc=npz), 8, 1e-12
    if name == "thermal3d_bench":   # 2x2x2-style split of a cube
        npz = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[world]
        return decks.thermal(3, tuple(12 * k for k in npz), ppc=4, temp_k=3.0e8, nproc=npz), 8, 1e-12
    if name == "thermal1d":
        return decks.thermal(1, (128,), ppc=8, temp_k=3.0e8, nproc=(world, 1, 1)), 10, 1e-12
    if name == "reflect2d":
        return decks.thermal(2, (48, 40), ppc=5, temp_k=4.0e8, nproc=(world, 1, 1), bc="reflect"), 10, 1e-12
    if name == "foil2d":
        return decks.foil2d(n=(96, 64), nproc=(world, 1, 1), nsteps=30), 30, 1e-11
    if name == "foil2d_xy":         # C3's pinned nprocx x nprocy decomposition, down-scaled
        px = {1: 1, 2: 2, 4: 2, 8: 4}[world]
        return decks.foil2d(n=(128, 64), nproc=(px, world // px, 1), nsteps=30), 30, 1e-11
    if name == "solver2d":   # order-4 field solver + strided compensated current smoothing across ranks
        dk = decks.thermal(2, (48, 40), ppc=5, temp_k=3.0e8, nproc=(world, 1, 1))
        dk.field_order = 4
        dk.smooth_currents, dk.smooth_iterations, dk.smooth_compensation, dk.smooth_strides = True, 2, True, (1, 2)
        return dk, 8, 1e-12
    if name == "mixed2d":    # c_bc_mixed: electrons reflect, protons leave (per-species current sums across ranks)
        dk = decks.thermal(2, (48, 40), ppc=5, temp_k=4.0e8, nproc=(world, 1, 1), bc="reflect", two_species=True)
        dk.species[1].bc_particle = ["open"] * 4
        return dk, 8, 1e-12
    if name == "laser2d_y":   # laser on y_min, decomposed along y and x
        return decks.laser2d_y(nproc=(1, world, 1) if world < 4 else (2, world // 2, 1)), 40, 1e-12
    if name == "laser2d":
        return decks.laser2d(nproc=(world, 1, 1), n=64), 40, 1e-12
    raise KeyError(name)


def run_case(name, rank, world, share_id, strict=True, sort_interval=2, moments=True):
    """`share_id(id_or_None) -> id`: broadcasts rank 0's 128-byte ncclUniqueId to every rank."""
    from epoch_b200.pic import Simulation
    from oracle.oracle import Oracle
    dk, nsteps, tol = make_deck(name, world)
    assert dk.nranks() == world
    o = Oracle(dk)
    if dk.species:
        o.auto_load()
    sim = Simulation(dk, rank=rank, strict_fp=strict, sort_interval=sort_interval, capacity_factor=3.0)
    if world > 1:
        sim.set_comm(share_id(Simulation.nccl_unique_id() if rank == 0 else None))
    for isp in range(len(dk.species)):
        sim.upload_species(isp, o.get_particles(rank, isp))
    ranks = list(range(world))

    class Both:
        def set_laser_source(self, lr, side, s1, s2):
            o.set_laser_source(ranks[lr], side, s1, s2)
            if ranks[lr] == rank:
                sim.set_laser_source(0, side, s1, s2)
        def init(self): o.init(); sim.init()
        def fields_half(self): o.fields_half(); sim.fields_half()
        def push(self): o.push(); sim.push()
        def current_finish(self): o.current_finish(); sim.current_finish()
        def fields_final(self): o.fields_final(); sim.fields_final()

    D.run(dk, Both(), ranks, None, max_steps=nsteps)
    res = {"case": name, "rank": rank, "ok": True, "msgs": []}
    for f in FIELDS:
        e = rel_l2(sim.download_field(f), o.field(rank, f))
        if not e <= tol:
            res["ok"] = False
            res["msgs"].append(f"{f}: rel_l2 {e:.3e}")
    for isp in range(len(dk.species)):
        a, b = sim.count(isp), o.count(rank, isp)
        if a != b:
            res["ok"] = False
            res["msgs"].append(f"species {isp}: count {a} != oracle {b}")
        elif not np.array_equal(sim.cell_counts(isp), o.cell_counts(rank, isp)):
            res["ok"] = False
            res["msgs"].append(f"species {isp}: per-cell counts differ")
        tot = sim.global_count(isp)
        want = sum(o.count(r, isp) for r in range(world))
        if tot != want:
            res["ok"] = False
            res["msgs"].append(f"species {isp}: global count {tot} != {want}")
    # device-side diagnostics moments (collective: every rank takes part in the ghost-cell sums)
    if dk.species and moments:
        for kind in ("number_density", "charge_density", "ekbar", "temperature", "temperature_y"):
            for isp in [-1] + list(range(len(dk.species))):
                e = rel_l2(sim.moment(kind, isp), o.moment(rank, kind, isp))
                if not e <= 1e-12:
                    res["ok"] = False
                    res["msgs"].append(f"moment {kind} species {isp}: rel_l2 {e:.3e}")
    res["counts"] = [sim.count(i) for i in range(len(dk.species))]
    sim.close()
    return res
