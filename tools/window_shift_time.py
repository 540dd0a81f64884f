#!/usr/bin/env python
"""Wall time of one moving-window shift (epb_shift_window) next to one PIC step, on one GPU: a thermal plasma with
open x, periodic y, 64 particles per cell.  Usage: tools/window_shift_time.py [cells_per_side]   (default 1024)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epoch_b200.pic import Simulation  # noqa: E402
from tests import decks  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dk = decks.thermal(2, (n, n), ppc=64, temp_k=1.0e7, bc=["open", "open", "periodic", "periodic"])
    for s in dk.species:
        s.bc_particle = ["open", "open", "periodic", "periodic"]
    sim = Simulation(dk, strict_fp=False, sort_interval=8, capacity_factor=1.1)
    sim.load_uniform(0, seed=7)
    sim.init()

    def step():
        sim.fields_half(); sim.push(); sim.current_finish(); sim.fields_final()

    for _ in range(3):
        step()
    sim.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    sim.synchronize()
    t_step = (time.perf_counter() - t0) / 5
    shifts = []
    for _ in range(4):
        step()
        sim.synchronize()
        t0 = time.perf_counter()
        dk.shift_window_geometry()
        sim.shift_window(1)          # no new plasma: the last column stays empty, which does not change the cost
        sim.synchronize()
        shifts.append(time.perf_counter() - t0)
    step()
    sim.synchronize()
    print(json.dumps({"cells": [n, n], "ppc": 64, "particles": sim.count(0), "ms_per_step": 1e3 * t_step,
                      "ms_per_shift": [1e3 * t for t in shifts]}))


if __name__ == "__main__":
    main()
