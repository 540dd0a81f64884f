#!/usr/bin/env python
"""bench.py — particle-updates/s of the PIC hot path (push + deposit + FDTD + exchange
[+ amortised sort]) on BASELINE.md config C2: epoch2d uniform thermal plasma, periodic,
4096 x 4096 cells, 64 particles per cell (1.07e9 particles) per B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun, one rank per GPU; the domain is decomposed as EPOCH's
split_domain would (mpi_routines.F90:107-138) with 4096^2 cells per rank (weak scaling).
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of EPOCH's
algorithm (oracle/) on the host cores: the reference binary is Fortran 2003 + MPI and can
not be built in this image (no Fortran compiler, no MPI), see DESIGN.md.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-updates/s (push+deposit+FDTD)"
UNIT = "particle-updates/s"


def c2_deck(n_local, ppc, nproc, temp_k=1.0e7, density=1.0e25):
    """uniform thermal electrons, dx = dy = Debye length."""
    from epoch_b200 import deck as D
    debye = math.sqrt(D.epsilon0 * D.kb * temp_k / (density * D.q0 ** 2))
    n = [n_local * nproc[0], n_local * nproc[1]]
    sp = [D.Species("electron", -D.q0, D.m0, npart_per_cell=ppc, density=density, temp=(temp_k,) * 3)]
    return D.Deck(2, n, [0.0, 0.0], [debye * n[0], debye * n[1]], ["periodic"] * 4, species=sp,
                  nproc=(nproc[0], nproc[1], 1))


def c4_deck(n_local, ppc, nproc, temp_k=1.0e7, density=1.0e25):
    """BASELINE.md C4 shape: epoch3d uniform thermal electrons, periodic, one n_local^3 block per GPU."""
    from epoch_b200 import deck as D
    debye = math.sqrt(D.epsilon0 * D.kb * temp_k / (density * D.q0 ** 2))
    n = [n_local * nproc[0], n_local * nproc[1], n_local * nproc[2]]
    sp = [D.Species("electron", -D.q0, D.m0, npart_per_cell=ppc, density=density, temp=(temp_k,) * 3)]
    return D.Deck(3, n, [0.0] * 3, [debye * k for k in n], ["periodic"] * 6, species=sp, nproc=tuple(nproc))


def split_3d(nranks):
    """8 -> (2,2,2), 4 -> (1,2,2), 2 -> (1,1,2), 1 -> (1,1,1) (epoch3d mpi_routines.F90:127-153 on a cube)."""
    return {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[nranks]


def split_2d(nranks):
    """split_domain's minimum-surface rule for a square per-rank tile (mpi_routines.F90:107-138):
    8 -> (2,4), 4 -> (2,2), 2 -> (1,2), 1 -> (1,1)."""
    best, area = (1, nranks), None
    for ix in range(1, nranks + 1):
        iy = nranks // ix
        if ix * iy != nranks:
            continue
        a = ix + iy
        if area is None or a < area:
            best, area = (ix, iy), a
    return best


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows, self.proc = [], None
        self.index = index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        mx = 0
        for r in self.rows:
            try:
                mx = max(mx, int(float(r[1])))
            except Exception:
                pass
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _oracle_sample(args):
    n, ppc, steps, seed = args
    from oracle.oracle import Oracle
    dk = c2_deck(n, ppc, (1, 1))
    dk.seed = seed
    o = Oracle(dk)
    o.auto_load()
    o.init()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.fields_half(); o.push(); o.current_finish(); o.fields_final()
    return n * n * ppc * steps, time.perf_counter() - t0


def cpu_baseline(n=192, ppc=64, steps=4, procs=1):
    """CPU restatement of EPOCH's algorithm on a bounded sample of the C2 workload."""
    import multiprocessing as mp
    from oracle import oracle as _o
    _o.build()
    t0 = time.perf_counter()
    if procs == 1:
        res = [_oracle_sample((n, ppc, steps, 7842432))]
        wall = res[0][1]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_oracle_sample, [(n, ppc, steps, 7842432 + i) for i in range(procs)])
        wall = max(r[1] for r in res)
    updates = sum(r[0] for r in res)
    return {"value": updates / wall, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{procs} x (C2 physics at {n}x{n} cells, {ppc} ppc, {steps} steps; "
                      f"oracle/epoch_oracle.cpp, g++ -O3 -ffp-contract=off); "
                      "CPU restatement of EPOCH's algorithm, not the EPOCH binary",
            "wall_s": time.perf_counter() - t0}


def e2e_full(n=512, ppc=64, steps=10):
    """The drop-in cost an EPOCH user pays around the resident loop: host/replay.cpp -- a
    compiled host that keeps the particles in EPOCH-style linked lists -- attaches (list -> pack_particle buffer ->
    device), runs `steps` steps and takes a full particle dump back into fresh lists, on a bounded C2-physics
    sample (n x n cells).  Reported per phase; `value` is updates/s with attach and dump inside the clock."""
    import tempfile
    import numpy as np
    from epoch_b200 import deck as D
    from epoch_b200.pic import Simulation
    exe = os.path.join(ROOT, "host", "replay")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
    dk = c2_deck(n, ppc, (1, 1))
    sim = Simulation(dk, strict_fp=False, sort_interval=0, capacity_factor=1.5)   # for its config structs
    rng = np.random.default_rng(7)
    npart = n * n * ppc
    spc = dk.species[0]
    dx = dk.dx(0)
    p = np.empty((npart, 6))
    cell = np.repeat(np.arange(n * n), ppc)
    p[:, 0] = dk.x_global(0, cell % n + 1) + (rng.random(npart) - 0.5) * dx
    p[:, 1] = dk.x_global(1, cell // n + 1) + (rng.random(npart) - 0.5) * dx
    p[:, 2:5] = rng.normal(size=(npart, 3)) * math.sqrt(spc.temp[0] * D.kb * spc.mass)
    p[:, 5] = spc.density * dx * dx / ppc
    with tempfile.TemporaryDirectory() as td:
        state = os.path.join(td, "state.bin")
        sim.write_replay_state(state, steps, {}, [p])
        sim.close()
        del p
        r = subprocess.run([exe, state, "-"], capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        return {"error": r.stderr[-300:]}
    info = json.loads(r.stdout.strip().splitlines()[-1])
    tot = info["attach_s"] + info["steps_s"] + info["download_s"]
    nbytes = info["particles"] * info["bytes_per_particle"]
    return {"value": info["particles"] * steps / tot, "unit": UNIT, "particles": info["particles"], "steps": steps,
            "attach_s": info["attach_s"], "list_to_packed_s": info["list_to_packed_s"], "upload_api_s": info["upload_api_s"],
            "upload_api_GBps": nbytes / max(info["upload_api_s"], 1e-9) / 1e9, "steps_s": info["steps_s"],
            "download_s": info["download_s"], "download_api_s": info["download_api_s"],
            "download_api_GBps": nbytes / max(info["download_api_s"], 1e-9) / 1e9,
            "packed_to_list_s": info["packed_to_list_s"],
            "host": "host/replay.cpp (C++, linked lists of heap nodes; the Fortran shim's call sequence)",
            "sample": f"C2 physics at {n}x{n} cells, {ppc} ppc"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    vals = []
    for i in range(args.warmup + args.steps):
        b = cpu_baseline(n=128, ppc=64, steps=2, procs=procs)
        if i >= args.warmup:
            vals.append(b)
    v = sum(b["value"] for b in vals) / len(vals)
    sample = vals[-1]["sample"]
    # The samples above are independent periodic domains, one per core: they pay no exchange.  For the record, ONE
    # domain through the multi-rank oracle (2 x 2 ranks in one process, i.e. with the halo / current-sum / particle
    # exchange of boundary.F90 but on one core) against the same domain on one rank: what decomposition costs the CPU
    # algorithm itself.
    decomposed = None
    try:
        from oracle.oracle import Oracle
        res = {}
        for name, nproc in (("one_rank", (1, 1)), ("four_ranks_2x2", (2, 2))):
            dk = c2_deck(256 // nproc[0], 64, nproc)
            o = Oracle(dk)
            o.auto_load()
            o.init()
            t0 = time.perf_counter()
            for _ in range(2):
                o.fields_half(); o.push(); o.current_finish(); o.fields_final()
            res[name] = 256 * 256 * 64 * 2 / (time.perf_counter() - t0)
        decomposed = {"unit": UNIT, "cores": 1, "sample": "C2 physics at 256x256 cells, 64 ppc, 2 steps", **res}
    except Exception as e:
        decomposed = {"error": f"{type(e).__name__}: {e}"[:200]}
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(b["wall_s"] for b in vals) / len(vals),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            # the same workload the CUDA arm reports; the CPU arm advances a bounded sample of it per step
            # (cpu_baseline.sample), there being no way to hold 1.07e9 particles x 10 steps in a few minutes of CPU
            "config": {"workload": f"epoch2d uniform thermal plasma, periodic, {args.n}x{args.n} cells per GPU, "
                                   f"{args.ppc} ppc ({args.n * args.n * args.ppc} particles per GPU), triangle shape, "
                                   "Yee order 2 (BASELINE C2)",
                       "sampled": True},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "decomposed": decomposed,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def c3_deck(n_local, ppc, nproc):
    """epoch2d laser-solid interaction (after epoch2d/example_decks/ramp.deck): simple_laser on
    x_min, simple_outflow on x_max, y periodic; an overdense electron + proton slab in the middle fifth of the
    box, nothing elsewhere -- so a plain nprocx x nprocy split leaves the outer x slabs without particles."""
    from epoch_b200 import deck as D
    lam = 1 * D.micron
    omega = 2 * D.pi * D.c / lam
    ncrit = omega ** 2 * D.epsilon0 * D.m0 / D.q0 ** 2
    n = [n_local * nproc[0], n_local * nproc[1]]
    dx = lam / 32.0
    L = [n[0] * dx, n[1] * dx]
    las = D.Laser("x_min", D.Laser.amp_from_intensity_w_cm2(1.0e19), omega,
                  profile=lambda y, z: D.gauss(y, 0, 0.25 * L[1]),
                  t_profile=lambda t: D.gauss(t, 30 * D.femto, 12 * D.femto) if t < 30 * D.femto else 1.0)
    box_lo, box_hi = (0.55 * L[0], -1e300, -1e300), (0.75 * L[0], 1e300, 1e300)
    bcp = ["open", "open", "periodic", "periodic"]
    sp = [D.Species("electron", -D.q0, D.m0, npart_per_cell=ppc, density=10 * ncrit, temp=(1.0e6,) * 3,
                    box_lo=box_lo, box_hi=box_hi, bc_particle=bcp),
          D.Species("proton", D.q0, 1836.2 * D.m0, npart_per_cell=ppc, density=10 * ncrit, temp=(1.0e6,) * 3,
                    box_lo=box_lo, box_hi=box_hi, bc_particle=bcp)]
    return D.Deck(2, n, [0.0, -L[1] / 2], [L[0], L[1] / 2], ["simple_laser", "simple_outflow", "periodic", "periodic"],
                  species=sp, lasers=[las], nproc=(nproc[0], nproc[1], 1), t_end=1.0)


def run_c3(args, world, rank, local_rank, parity, torch, dist):
    """--workload c3: K steps with EPOCH's plain split, then the slabs are re-cut (epb_load_profile ->
    calculate_breaks -> epb_redistribute) and K steps are timed again.  `value` is the balanced run."""
    import numpy as np
    from epoch_b200 import deck as D
    from epoch_b200.pic import Simulation
    nproc = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[world]
    dk = c3_deck(args.n, args.ppc, nproc)
    stream = torch.cuda.Stream()
    ncell_foil = 0.2 * dk.n[0] * dk.n[1]
    cap = int(1.6 * ncell_foil * args.ppc / world) + (1 << 20)       # per species; every rank can take an equal share
    n_loc, g_loc = dk.local_extent(rank)
    # before the re-cut the two middle x slabs hold everything
    frac = max(0.0, min(0.75 * dk.n[0], g_loc[0] - 1 + n_loc[0]) - max(0.55 * dk.n[0], g_loc[0] - 1)) / max(1, n_loc[0])
    cap0 = max(cap, int(1.1 * frac * n_loc[0] * n_loc[1] * args.ppc) + (1 << 20))
    sim = Simulation(dk, rank=rank, strict_fp=bool(args.strict), sort_interval=args.sort_interval,
                     capacity_factor=cap0 / max(1.0, args.ppc * n_loc[0] * n_loc[1]), stream=stream.cuda_stream)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(Simulation.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        sim.set_comm(bytes(idt.cpu().tolist()))
    # loader (stands in for auto_load, helper.F90:95): ppc particles in every foil cell of this rank
    rng = np.random.default_rng(20261018 + rank)
    dx = dk.dx(0)
    ix = np.arange(g_loc[0], g_loc[0] + n_loc[0])
    xc = dk.x_global(0, ix)
    in_foil = ix[(xc >= dk.species[0].box_lo[0]) & (xc < dk.species[0].box_hi[0])]
    for isp, spc in enumerate(dk.species):
        npart = len(in_foil) * n_loc[1] * args.ppc
        p = np.empty((npart, 6))
        if npart:
            cx = np.repeat(in_foil, n_loc[1] * args.ppc)
            cy = np.tile(np.repeat(np.arange(g_loc[1], g_loc[1] + n_loc[1]), args.ppc), len(in_foil))
            p[:, 0] = dk.x_global(0, cx) + (rng.random(npart) - 0.5) * dx
            p[:, 1] = dk.x_global(1, cy) + (rng.random(npart) - 0.5) * dx
            sd = math.sqrt(spc.temp[0] * D.kb * spc.mass)
            p[:, 2:5] = rng.normal(size=(npart, 3)) * sd
            p[:, 5] = spc.density * dx * dx / args.ppc
        sim.upload_species(isp, p)
        del p
    dt = dk.dt()
    state = {"t": dt / 2.0}

    def sources():
        for side in (0, 1):
            if sim.geo["is_bnd"][side]:
                s1, s2 = sim.deck.laser_sources(rank, side, state["t"])
                sim.set_laser_source(0, side, s1, s2)

    sources()
    sim.init()

    def step():
        sim.fields_half(); sim.push(); sim.current_finish()
        state["t"] += dt
        sources()
        sim.fields_final()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k):
        barrier()
        sim.push_kernel_ms(reset=1)
        l0 = sim.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(k):
                step()
            e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        pm, pn = sim.push_kernel_ms(reset=2)
        t = torch.tensor([ms, wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item()), sim.launch_count() - l0, pm, pn

    def counts():
        loc = sum(sim.count(i) for i in range(2))
        t = torch.tensor([loc], dtype=torch.int64, device="cuda")
        allc = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(allc, t)
        else:
            allc = [t]
        return [int(c.item()) for c in allc]

    for _ in range(args.warmup):
        step()
    c_before = counts()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_u, wall_u, launches_u, _, _ = timed(args.steps)
    n_total = sum(c_before)
    # ---- the balancer: EPOCH's rule on the device histogram, then the device remap ----
    t0 = time.perf_counter()
    cuts = {}
    for axis in (0, 1):
        if nproc[axis] > 1:
            cuts[axis] = D.calculate_breaks(sim.load_profile(axis), nproc[axis])
    if cuts:
        sim.rebalance(cuts, capacities=[cap, cap])
    barrier()
    rebalance_s = time.perf_counter() - t0
    c_after = counts()
    for _ in range(2):
        step()
    ms_b, wall_b, launches_b, push_ms, push_n = timed(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    n_end = sum(counts())
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        bpu = 88.0 + 120.0 / (2 * args.ppc)
        # the fullest rank's kernel: its particles x bytes / its push time (rank 0's own timing is reported)
        achieved = (c_after[0] * bpu / (push_ms * 2 * 1e-3)) / 1e9 if push_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": n_total * args.steps / (ms_b * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_b / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"epoch2d laser-solid: simple_laser on x_min (1e19 W/cm^2, 1 um), simple_outflow on x_max, "
                                   f"y periodic, 10 n_crit e-/p+ foil over x in [0.55, 0.75] of the box, {args.ppc} ppc per species, "
                                   f"{dk.n[0]}x{dk.n[1]} cells, pinned {nproc[0]}x{nproc[1]} decomposition (BASELINE C3)",
                       "decomposition": f"{nproc[0]}x{nproc[1]}", "strict_fp": int(args.strict), "particles_total": n_total,
                       "particles_end": n_end, "load_balancer": "calculate_breaks (balance.F90:1948) on epb_load_profile, "
                                                                "epb_redistribute", "cuts": {str(k): v for k, v in cuts.items()},
                       "l2_policy": "particle state per GPU exceeds the 126 MB L2"},
            "c3": {"unbalanced_value": n_total * args.steps / (ms_u * 1e-3), "unbalanced_ms_per_step": ms_u / args.steps,
                   "balanced_ms_per_step": ms_b / args.steps, "rebalance_s": rebalance_s,
                   "particles_per_rank_before": c_before, "particles_per_rank_after": c_after},
            "e2e": {"value": n_total * args.steps / wall_b, "unit": UNIT,
                    "h2d_bytes_per_step": 2 * 2 * (sim.geo["n"][1] + 1) * 8, "d2h_bytes_per_step": 0,
                    "note": "wall clock of the same steps through the C ABI incl. the per-step laser source planes from the host"},
            "gpu_launches": launches_b, "parity_check": parity,
            "parity_cases": f"thermal2d_bench, foil2d_xy on {world} rank(s) (tests/parity_check.py)",
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": None,
                         "kernel": "push_slots_2d (push+deposit), two species, rank 0 after the re-cut",
                         "bytes_per_update": bpu, "kernel_ms": push_ms, "kernel_launches": push_n,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s"},
        }
        print(json.dumps(line), flush=True)
    sim.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cells", dest="n", type=int, default=4096, help="cells per side per GPU")
    ap.add_argument("--ppc", type=int, default=None)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 (default, the bench line): 2D 4096^2 x 64 ppc per GPU; c4: 3D 384^3 x 8 ppc per GPU; "
                         "c3: 2D laser-solid (laser on x_min, overdense e-/p+ foil), pinned nprocx x nprocy, load "
                         "balancer off and on")
    # kernel-selection switches are honoured by the library only under EPB_DEBUG=1 (csrc/epb_internal.h: epb_env)
    variant = int(os.environ.get("EPB_PUSH_VARIANT", "5")) if os.environ.get("EPB_DEBUG", "0") not in ("", "0") else 5
    ap.add_argument("--sort-interval", type=int, default=int(os.environ.get("EPB_SORT_INTERVAL", "0")),
                    help="0 = the library default: 2 for the cell-owner 2D kernel, 8 otherwise")
    ap.add_argument("--strict", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-full", action="store_true",
                    help="skip the compiled-host attach / dump measurement (host/replay.cpp) at N = 1")
    ap.add_argument("--no-mixed", action="store_true",
                    help="skip the second, shorter measurement from the relaxed (Poisson) particle distribution")
    ap.add_argument("--no-parity-check", action="store_true",
                    help="skip the untimed pre-phase that checks this decomposition against the multi-rank CPU oracle")
    args = ap.parse_args()
    if args.workload == "c4":
        if args.n == 4096:
            args.n = 384
        args.ppc = args.ppc or 8
    if args.workload == "c3":
        if args.n == 4096:
            args.n = 2048
        args.ppc = args.ppc or 32
    if args.workload == "c5" and args.n == 4096:
        args.n = 2048
    args.ppc = args.ppc or 64
    if args.sort_interval <= 0:
        args.sort_interval = 2 if (args.workload == "c2" and variant in (2, 3, 4)) else 8
    if args.impl == "reference":
        return run_reference(args)
    # keep stdout to the one JSON line: NCCL prints its version banner there at VERSION level
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    from epoch_b200.pic import Simulation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; epoch_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.workload == "c4":
        nproc = split_3d(world)
        dk = c4_deck(args.n, args.ppc, nproc)
    elif args.workload == "c3":
        nproc, dk = None, None
    elif args.workload == "c5":
        # BASELINE config 5: a dense plasma with binary collisions every step (collisions.F90): the C2 deck at solid
        # density and 1e6 K, electron-electron collisions (Nanbu-Perez, coulomb_log = auto) after every push
        nproc = split_2d(world)
        dk = c2_deck(args.n, args.ppc, nproc, temp_k=1.0e6, density=1.0e28)
    else:
        nproc = split_2d(world)
        dk = c2_deck(args.n, args.ppc, nproc)
    # ---- untimed pre-phase: parity of THIS decomposition (every rank checks its own share of small decomposed
    # decks against the multi-rank CPU oracle: per-cell / per-rank / global counts bit-exact, E/B/J and the
    # moments within 1e-12 relative L2; tests/parity_check.py).  The oracle is the checker here, nothing of
    # it is timed.  The verdict travels in the JSON line so that the scaling runs carry it.
    parity = "skipped"
    if not args.no_parity_check:
        from tests.parity_check import run_case

        def share_id(uid):
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                t = torch.tensor(list(uid), dtype=torch.uint8, device="cuda")
            if world > 1:
                dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())

        cases = ["thermal3d_bench"] if args.workload == "c4" else ["thermal2d_bench", "foil2d_xy"]
        msgs = []
        for name in cases:
            for strict in (True, False):
                try:
                    r = run_case(name, rank, world, share_id, strict=strict, moments=strict)
                    if not r["ok"]:
                        msgs.append(f"{name}[strict={int(strict)}] rank {rank}: " + "; ".join(r["msgs"]))
                except Exception as e:  # a failing check must not hide behind a crash
                    msgs.append(f"{name}[strict={int(strict)}] rank {rank}: {type(e).__name__}: {e}")
        bad = torch.tensor([len(msgs)], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(bad)
        parity = "ok" if int(bad.item()) == 0 else ("FAILED: " + " | ".join(msgs) if msgs else "FAILED on another rank")
        if msgs:
            print("parity_check:", msgs, file=sys.stderr, flush=True)
    if args.workload == "c3":
        run_c3(args, world, rank, local_rank, parity, torch, dist)
        if world > 1:
            dist.destroy_process_group()
        return
    stream = torch.cuda.Stream()

    def make_sim(mixed):
        sm = Simulation(dk, rank=rank, strict_fp=bool(args.strict), sort_interval=args.sort_interval,
                        capacity_factor=1.02 if world == 1 else 1.15, stream=stream.cuda_stream)
        if world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt = torch.tensor(list(Simulation.nccl_unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(idt, 0)
            sm.set_comm(bytes(idt.cpu().tolist()))
        sm.load_uniform(0, seed=20261017, mixed=mixed)
        sm.init()
        if args.workload == "c5":   # PROGRAM pic's collision step: after push_particles, before current_finish
            def step_with_collisions(sm=sm):
                sm.fields_half()
                sm.push()
                if coll_events is not None:
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    sm.collide([[1.0]])
                    b.record(stream)
                    coll_events.append((a, b))
                else:
                    sm.collide([[1.0]])
                sm.current_finish()
                sm.fields_final()
            sm.step = step_with_collisions
        return sm

    coll_events = None

    sim = make_sim(False)
    n_local = sim.count(0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") --------------------------------------
    for _ in range(args.warmup):
        sim.step()
    sim.push_kernel_ms(reset=1)   # start timing the push kernel with CUDA events on its stream
    launches0 = sim.launch_count()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.workload == "c5":
        coll_events = []
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            sim.step()
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    collide_ms = None
    if coll_events:
        collide_ms = sum(a.elapsed_time(b) for a, b in coll_events) / len(coll_events)
    coll_events = None
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.launch_count() - launches0
    push_ms, push_n = sim.push_kernel_ms(reset=2)
    n_total = sim.global_count(0)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n_total * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with host buffers ("e2e") ------------------
    # Per step the host supplies what EPOCH's Fortran side evaluates each step (the boundary
    # source planes, laser.f90:347-352; zero for this periodic deck but copied all the same)
    # from pinned memory, and reads back the step's results: the global particle count
    # (update_particle_count), the field energies (calc_total_energy_sum) and the Ey array as
    # a field dump would (asynchronously: epb_download_field_async).  The particle state itself stays device-resident by design.
    # The e2e loop runs on a FRESH load of the same deck, so that it covers the same physical steps as the device-timed
    # loop above (warm-up steps 1..W, timed steps W+1..W+K): the plasma relaxes from exactly ppc particles per cell
    # towards Poisson counts and every step is a little slower than the one before (51.5 ms fresh, 54.8 ms relaxed at
    # C2) -- continuing on the same state would book that drift (2.9 % over 20 steps) as host overhead.
    shape, geo_n = sim.shape, sim.geo["n"]
    sim.close()
    sim = make_sim(False)
    ny1 = (geo_n[1] + 1) * (geo_n[2] + 1 if args.workload == "c4" else 1)
    src = torch.zeros(2, ny1, dtype=torch.float64).pin_memory()
    ey_host = torch.empty(shape, dtype=torch.float64).pin_memory()
    scal = [torch.zeros(8, dtype=torch.float64).pin_memory() for _ in range(2)]
    h2d = 2 * 2 * ny1 * 8
    d2h = ey_host.numel() * 8 + 4 * 8
    prof = None

    def timed_call(name, fn, *a):
        if prof is None:
            return fn(*a)
        t1 = time.perf_counter()
        r = fn(*a)
        prof[name] = prof.get(name, 0.0) + time.perf_counter() - t1
        return r

    # The host runs one step ahead of the device, as a production host would: it hands over step k's source planes
    # and enqueues step k, its scalar diagnostics and its Ey dump, and only then waits for what step k-1 sent back
    # (counts and energies in page-locked memory, the Ey array of the previous dump).  Every step's inputs and
    # results cross the bus inside the clock; nothing is skipped, the device just never waits for the host.
    tickets = []
    results = []

    def e2e_step(k, last):
        if not os.environ.get("EPB_BENCH_NO_SRC"):       # diagnosis only
            for side in (0, 1):
                timed_call("set_laser_source", sim.L.epb_set_laser_source, sim._h, side, src[0].data_ptr(), src[1].data_ptr())
        timed_call("step (enqueue)", sim.step)
        if os.environ.get("EPB_BENCH_NO_SCAL"):          # diagnosis only
            if last:
                tickets.append(sim.step_scalars_async(scal[k % 2].data_ptr()))
            return
        tickets.append(timed_call("step_scalars_async", sim.step_scalars_async, scal[k % 2].data_ptr()))
        if os.environ.get("EPB_BENCH_NO_DUMP"):      # diagnosis only: what the Ey dump costs
            pass
        elif os.environ.get("EPB_BENCH_SYNC_DUMP"):
            timed_call("download_field", sim.download_field_into, "ey", ey_host.data_ptr())
        else:
            # one host array: the previous dump must have arrived before this one may overwrite it
            timed_call("wait_downloads (previous dump)", sim.wait_downloads)
            timed_call("download_field_async", sim.download_field_async, "ey", ey_host.data_ptr())
        if len(tickets) > 1:
            timed_call("wait_scalars (previous step)", sim.wait_scalars, tickets[-2])
            results.append(scal[(k - 1) % 2][:4].tolist())

    # W untimed iterations of exactly this loop first: first-use allocations (dump staging buffer, copy stream,
    # scratch of the reductions, the staging ring of the source planes) belong to start-up, and the clock must start
    # on a device that is already running -- a B200 that has idled for the ~100 ms the page-locked allocations above
    # take needs some 30 ms of work to come back to its clocks, which a 10-step region would book as 3 ms per step
    for k in range(args.warmup):
        e2e_step(k, k == args.warmup - 1)
    sim.wait_scalars(tickets[-1])
    sim.wait_downloads()
    barrier()
    tickets.clear()
    results.clear()
    prof = {} if os.environ.get("EPB_BENCH_E2E_BREAKDOWN") else None   # host wall time of every call of the loop
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_step(k, k == args.steps - 1)
    sim.wait_scalars(tickets[-1])
    results.append(scal[(args.steps - 1) % 2][:4].tolist())
    sim.wait_downloads()
    barrier()
    e2e_s = time.perf_counter() - t0
    if not os.environ.get("EPB_BENCH_NO_SCAL") and (any(r[3] != 0.0 for r in results) or any(int(r[2]) != n_total for r in results)):
        raise RuntimeError(f"e2e: the per-step results are wrong: {results[-1]} (expected {n_total} particles)")
    if prof is not None and rank == 0:
        print("e2e breakdown (ms per step): " + ", ".join(f"{k} {1e3 * v / args.steps:.3f}" for k, v in prof.items()) +
              f"; total {1e3 * e2e_s / args.steps:.3f}", file=sys.stderr, flush=True)
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = n_total * args.steps / e2e_s

    # ---- the same steps from the relaxed state ("mixed_state") --------------------------
    # EPOCH's loader puts exactly ppc particles in every cell; a thermal plasma relaxes to Poisson counts, where a
    # warp runs as many rounds as its fullest column.  The figure above is the deck as loaded (what the reference arm
    # runs too); this one is the long-time state of the same deck, device-timed the same way.
    sim.close()   # give the device memory back (also before the compiled host attaches its own state)
    mixed_state = None
    if not args.no_mixed:
        msim = make_sim(True)
        for _ in range(3):
            msim.step()
        msim.push_kernel_ms(reset=1)
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        msteps = min(args.steps, 6)
        with torch.cuda.stream(stream):
            m0.record(stream)
            for _ in range(msteps):
                msim.step()
            m1.record(stream)
        barrier()
        mms = m0.elapsed_time(m1)
        mk_ms, _ = msim.push_kernel_ms(reset=2)
        mtot = msim.global_count(0)
        if world > 1:
            t = torch.tensor([mms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            mms = float(t.item())
        mixed_state = {"value": mtot * msteps / (mms * 1e-3), "unit": UNIT, "ms_per_step": mms / msteps, "steps": msteps,
                       "kernel_ms": mk_ms, "load": "every particle's cell drawn at random (Poisson counts per cell)"}
        msim.close()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        is3d = args.workload == "c4"
        # push+deposit kernel, particle state + 120 B/cell of E/B/J traffic
        bytes_per_update = (104.0 if is3d else 88.0) + 120.0 / args.ppc
        achieved = (n_local * bytes_per_update / (push_ms * 1e-3)) / 1e9 if push_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "push_traffic_bytes.json")
        if os.path.exists(tpath):
            try:  # ncu DRAM bytes per particle of the kernel (one --set full capture) x particles per launch
                tj = json.load(open(tpath))
                if is3d:
                    traffic = tj["push_bag_3d"]["dram_bytes_per_particle"] * n_local
                elif variant == 5:
                    traffic = tj["push_slots_2d"]["dram_bytes_per_particle"] * n_local
                elif variant in (2, 3, 4) and args.sort_interval >= 2:   # round-1 kernel: the launch mix of one sort cycle
                    bpp = tj["dram_bytes_per_particle"]
                    k = args.sort_interval
                    traffic = (bpp["fused_gather"] + (k - 2) * bpp["plain"] + bpp["emitting"]) / k * n_local
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (f"epoch3d uniform thermal plasma, periodic, {args.n}^3 cells per GPU, "
                                    f"{args.ppc} ppc ({n_local} particles per GPU), triangle shape, Yee order 2 "
                                    "(BASELINE C4 per-GPU share)") if is3d else
                                   (f"epoch2d dense plasma (1e28 m^-3, 1e6 K) with binary collisions every step (electron-electron, "
                                    f"Nanbu-Perez, coulomb_log = auto), periodic, {args.n}x{args.n} cells per GPU, {args.ppc} ppc "
                                    f"({n_local} particles per GPU) (BASELINE config 5)") if args.workload == "c5" else
                                   (f"epoch2d uniform thermal plasma, periodic, {args.n}x{args.n} cells per GPU, "
                                    f"{args.ppc} ppc ({n_local} particles per GPU), triangle shape, Yee order 2 "
                                    "(BASELINE C2)"),
                       "decomposition": "x".join(str(k) for k in nproc), "sort_interval": args.sort_interval,
                       "strict_fp": int(args.strict), "particles_total": n_total,
                       "l2_policy": "inputs (51.5 GB of particle state per GPU) far exceed the 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "parity_check": parity,
            "parity_cases": ("thermal3d_bench" if is3d else "thermal2d_bench, foil2d_xy") +
                            f" on {world} rank(s), parity and performance builds (tests/parity_check.py)",
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "kernel": "push_bag_3d (push+deposit, slot columns, warp-private deposit tiles)" if is3d else
                                   ("push_slots_2d (push+deposit, in-place slot columns)" if variant == 5 else
                                    "push_cell_2d (push+deposit)" if variant in (2, 3, 4) else "push_tiled_2d (push+deposit)"),
                         "bytes_per_update": bytes_per_update,
                         "kernel_ms": push_ms, "kernel_launches": push_n,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s"},
        }
        if not args.no_cpu_baseline and world == 1:
            b = cpu_baseline()
            b.pop("wall_s", None)
            line["cpu_baseline"] = b
        if mixed_state is not None:
            line["mixed_state"] = mixed_state
        if collide_ms is not None:
            line["collisions"] = {"kernel_ms_per_step": collide_ms, "pairs_per_s": 0.5 * n_local / (collide_ms * 1e-3),
                                  "note": "epb_collide incl. the inbox settle and the per-cell moments of coulomb_log = auto"}
        if not args.no_e2e_full and world == 1 and not is3d:
            try:
                line["e2e_full"] = e2e_full()
            except Exception as e:
                line["e2e_full"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
