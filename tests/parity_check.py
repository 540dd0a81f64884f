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
    if name in ("cpml2d", "cpml2d_y", "cpml3d"):   # CPML layers, laser planes and particle-domain offsets across ranks
        from tests.test_cpml import _cpml_case
        dk = _cpml_case(3 if name == "cpml3d" else 2, particles=True)
        dk.nproc = (1, world, 1) if name == "cpml2d_y" else (world, 1, 1) if world < 4 else (2, world // 2, 1)
        return dk, 40, 1e-12
    if name in ("window2d", "window3d", "window1d"):   # moving window at c: fields and particles change rank as the box moves
        from tests.test_window import window_deck
        nd = int(name[6])
        n = {1: (96,), 2: (48, 24), 3: (24, 9, 8)}[nd]
        npr = (world, 1, 1) if (world < 4 or nd == 1) else (2, world // 2, 1)
        return window_deck(nd, n, ppc=4, nproc=npr, nsteps=16), 16, 1e-12
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
        def shift_window(self, cells):
            # insert_particles is the host's (KISS stream, deck expressions): the oracle's new plasma goes to the device
            o.window_clear_inserted()
            o.shift_window(cells)
            sim.shift_window(cells, [o.window_inserted(rank, isp) for isp in range(len(dk.species))])

    D.run(dk, Both(), ranks, None, max_steps=nsteps)
    res = {"case": name, "rank": rank, "ok": True, "msgs": []}
    if dk.move_window and dk.window_shifts < 4:
        res["ok"] = False
        res["msgs"].append(f"the window moved only {dk.window_shifts} cells")
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


# ---------------------------------------------------------------------------------------------------------
# Load balancing (SURVEY.md 8 f2 / BASELINE C3): the slabs are re-cut in mid-run and fields and particles are
# redistributed on the device (epb_redistribute = balance_workload's data movement, balance.F90:93-300).
# ---------------------------------------------------------------------------------------------------------
def _assemble(deck, pieces, name=None):
    """Global interior array from per-rank interiors [(rank, array[z,y,x])] of `deck`'s decomposition."""
    shape = [deck.n[d] if d < deck.ndims else 1 for d in (2, 1, 0)]
    out = np.zeros(shape, dtype=pieces[0][1].dtype)
    for r, a in pieces:
        n, g = deck.local_extent(r)
        sl = tuple(slice(g[d] - 1, g[d] - 1 + n[d]) for d in (2, 1, 0))
        out[sl] = a
    return out


def _interior(a):
    sl = tuple(slice(5, -5) if a.shape[ax] > 1 else slice(None) for ax in range(3))
    return a[sl]


def run_rebalance_case(rank, world, share_id, allgather, k1=12, k2=10, strict=True, n=(96, 64)):
    """Foil deck (laser on x_min, particles only in the slab), `world` ranks along x.  After k1 steps the x
    slabs are re-cut with EPOCH's rule (epb_load_profile -> calculate_breaks) and the state is redistributed.
    Checks: the move loses / changes nothing (global particle multiset and global fields bit-identical, every
    particle on the rank get_particle_processor names, per-rank counts as numpy computes them from the
    positions); then k2 more steps agree with the CPU oracle, which keeps the ORIGINAL decomposition -- the
    physics does not depend on where the cuts are -- to 1e-11 on the assembled global E/B/J, with bit-exact
    global per-cell counts."""
    from epoch_b200.pic import Simulation
    from oracle.oracle import Oracle
    dk = decks.foil2d(n=n, nproc=(world, 1, 1), nsteps=k1 + k2)
    res = {"case": "rebalance2d", "rank": rank, "ok": True, "msgs": []}

    def bad(msg):
        res["ok"] = False
        res["msgs"].append(msg)

    o = Oracle(dk)
    o.auto_load()
    sim = Simulation(dk, rank=rank, strict_fp=strict, sort_interval=2, capacity_factor=3.0)
    if world > 1:
        sim.set_comm(share_id(Simulation.nccl_unique_id() if rank == 0 else None))
    nsp = len(dk.species)
    for isp in range(nsp):
        sim.upload_species(isp, o.get_particles(rank, isp))
    dt = dk.dt()
    ranks = list(range(world))
    state = {"time": 0.0}

    def sources(t):
        for side in range(2 * dk.ndims):
            if dk.has_boundary_source(side):
                for r in ranks:
                    s1, s2 = dk.laser_sources(r, side, t)
                    o.set_laser_source(r, side, s1, s2)
                s1, s2 = sim.deck.laser_sources(rank, side, t)   # the rank's CURRENT sub-domain
                sim.set_laser_source(0, side, s1, s2)

    state["time"] += dt / 2.0
    sources(state["time"])
    o.init(); sim.init()

    def step():
        o.fields_half(); sim.fields_half()
        o.push(); sim.push()
        o.current_finish(); sim.current_finish()
        state["time"] += dt
        sources(state["time"])
        o.fields_final(); sim.fields_final()

    for _ in range(k1):
        step()
    # ---- re-cut ----
    before_p = [sim.download_species(i) for i in range(nsp)]
    before_f = {f: _interior(sim.download_field(f)) for f in FIELDS}
    old_deck = sim.deck
    load = sim.load_profile(0)
    mins, maxs = D.calculate_breaks(load, world)
    res["cuts"] = [list(mins), list(maxs)]
    old_counts = [sim.count(i) for i in range(nsp)]
    sim.rebalance({0: (mins, maxs)}, capacities=[3 * sum(allgather(c)) // max(1, world) + 4096 for c in old_counts])
    after_p = [sim.download_species(i) for i in range(nsp)]
    after_f = {f: _interior(sim.download_field(f)) for f in FIELDS}
    allb = allgather((before_p, before_f))
    alla = allgather((after_p, after_f))
    from tests.gpu_util import sorted_rows
    for isp in range(nsp):
        b = sorted_rows(np.concatenate([x[0][isp] for x in allb]))
        a = sorted_rows(np.concatenate([x[0][isp] for x in alla]))
        if a.shape != b.shape or not np.array_equal(a, b):
            bad(f"species {isp}: the redistribution changed the global particle set")
        # get_particle_processor (balance.F90:2095-2151) on this rank's particles
        p = after_p[isp]
        dx = dk.dx(0)
        lo = dk.x_global(0, mins[rank]) - dx * (0.5 + (3.0 if rank == 0 else 0.0))
        hi = dk.x_global(0, maxs[rank]) + dx * (0.5 + (3.0 if rank == world - 1 else 0.0))
        if p.shape[0] and not np.all((p[:, 0] >= lo) & (p[:, 0] < hi)):
            bad(f"species {isp}: a particle sits on the wrong rank after the redistribution")
        allp = np.concatenate([x[0][isp] for x in allb])
        want = int(np.sum((allp[:, 0] >= lo) & (allp[:, 0] < hi)))
        if sim.count(isp) != want:
            bad(f"species {isp}: count {sim.count(isp)} after the re-cut, expected {want}")
    for f in FIELDS[:6]:
        gb = _assemble(old_deck, [(r, x[1][f]) for r, x in enumerate(allb)])
        ga = _assemble(sim.deck, [(r, x[1][f]) for r, x in enumerate(alla)])
        if not np.array_equal(ga, gb):
            bad(f"{f}: the redistribution changed the global field")
    # ---- carry on ----
    for _ in range(k2):
        step()
    mine = ({f: _interior(sim.download_field(f)) for f in FIELDS}, [sim.cell_counts(i) for i in range(nsp)],
            [sim.count(i) for i in range(nsp)])
    allm = allgather(mine)
    for f in FIELDS:
        g_sim = _assemble(sim.deck, [(r, x[0][f]) for r, x in enumerate(allm)])
        g_orc = _assemble(dk, [(r, _interior(o.field(r, f))) for r in ranks])
        e = rel_l2(g_sim, g_orc)
        if not e <= 1e-11:
            bad(f"{f} after the re-cut: rel_l2 {e:.3e}")
    for isp in range(nsp):
        g_sim = _assemble(sim.deck, [(r, x[1][isp]) for r, x in enumerate(allm)])
        g_orc = _assemble(dk, [(r, o.cell_counts(r, isp)) for r in ranks])
        if not np.array_equal(g_sim, g_orc):
            bad(f"species {isp}: global per-cell counts differ after the re-cut")
        tot = sim.global_count(isp)
        if tot != sum(o.count(r, isp) for r in ranks):
            bad(f"species {isp}: global count {tot}")
    res["counts_after"] = [x[2] for x in allm]
    res["counts_before"] = allgather(old_counts)
    sim.close()
    return res
