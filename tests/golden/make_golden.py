"""Generates the fixtures in this directory from the CPU oracle (oracle/epoch_oracle.cpp).

    python tests/golden/make_golden.py

The reference itself (Fortran 2003 + MPI) cannot be built in this image, and its tests hold
golden numbers only for the field half of the path (the laser decks; those scalars are
asserted directly in tests/test_oracle_golden.py and tests/test_gpu_parity.py).  These
fixtures pin the particle half on the oracle's output so that drift of either the oracle or
the CUDA path is caught without the other being present: each file holds a deck description,
the loaded particles (KISS stream, seed 7842432) and the complete state after `nsteps` steps
of PROGRAM pic's loop.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from epoch_b200 import deck as D  # noqa: E402
from oracle.oracle import FIELD_NAMES, Oracle  # noqa: E402
from tests import decks  # noqa: E402

CASES = {
    "thermal1d": (lambda: decks.thermal(1, (48,), ppc=6, temp_k=2.0e8), 8),
    "thermal2d": (lambda: decks.thermal(2, (24, 20), ppc=4, temp_k=2.0e8, two_species=True), 6),
    "thermal3d": (lambda: decks.thermal(3, (10, 9, 8), ppc=3, temp_k=2.0e8), 5),
    "reflect2d": (lambda: decks.thermal(2, (20, 16), ppc=4, temp_k=4.0e8, bc="reflect"), 8),
    "foil2d": (lambda: decks.foil2d(n=(64, 40), ppc=3, nsteps=20), 20),
}


def main():
    for name, (mk, nsteps) in CASES.items():
        dk = mk()
        o = Oracle(dk)
        o.auto_load()
        out = {"nsteps": np.int64(nsteps)}
        for isp in range(len(dk.species)):
            out[f"p0_{isp}"] = o.get_particles(0, isp)
        D.run(dk, o, [0], None, max_steps=nsteps)
        for f in FIELD_NAMES:
            out[f] = np.array(o.field(0, f))
        for isp in range(len(dk.species)):
            p = o.get_particles(0, isp)
            keys = tuple(p[:, k] for k in range(p.shape[1] - 1, -1, -1))
            out[f"p1_{isp}"] = p[np.lexsort(keys)]
            out[f"cc_{isp}"] = o.cell_counts(0, isp)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
