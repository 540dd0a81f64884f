"""Helpers for the GPU parity tests: drive the CUDA library and the CPU oracle with the
same deck, the same initial particles and the same call sequence."""
import numpy as np

from epoch_b200 import deck as D
from oracle.oracle import Oracle

FIELDS = ("ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz")


def make_pair(dk, strict=True, sort_interval=1, load=True, rank=0):
    """Single-rank pair (oracle, Simulation) with identical initial state."""
    from epoch_b200.pic import Simulation
    o = Oracle(dk)
    if load and dk.species:
        o.auto_load()
    sim = Simulation(dk, rank=rank, strict_fp=strict, sort_interval=sort_interval, capacity_factor=2.0)
    for isp in range(len(dk.species)):
        sim.upload_species(isp, o.get_particles(rank, isp))
    return o, sim


def set_random_fields(o, sim, dk, seed=0, e_amp=1e9, b_amp=3.0, rank=0):
    rng = np.random.default_rng(seed)
    for name in FIELDS[:6]:
        a = o.field(rank, name)
        a[...] = rng.normal(size=a.shape) * (e_amp if name[0] == "e" else b_amp)
        sim.upload_field(name, a)


def rel_l2(a, b):
    den = np.sqrt(np.sum(np.asarray(b, dtype=np.float64) ** 2))
    num = np.sqrt(np.sum((np.asarray(a, dtype=np.float64) - b) ** 2))
    if den == 0.0:
        return 0.0 if num == 0.0 else np.inf
    return num / den


def sorted_rows(p):
    if p.shape[0] == 0:
        return p
    keys = tuple(p[:, k] for k in range(p.shape[1] - 1, -1, -1))
    return p[np.lexsort(keys)]


def run_both(dk, o, sim, nsteps, rank=0):
    """PROGRAM pic loop on both backends (single rank)."""
    class Both:
        def set_laser_source(self, lr, side, s1, s2):
            o.set_laser_source(rank, side, s1, s2)
            sim.set_laser_source(0, side, s1, s2)
        def init(self): o.init(); sim.init()
        def fields_half(self): o.fields_half(); sim.fields_half()
        def push(self): o.push(); sim.push()
        def current_finish(self): o.current_finish(); sim.current_finish()
        def fields_final(self): o.fields_final(); sim.fields_final()
    return D.run(dk, Both(), [rank], None, max_steps=nsteps)
