"""Generates moments2d.npz: the particle moments of io/calc_df.F90 (all kinds of epb_calc_moment) and the
Higuera-Cary push, from the CPU oracle.

    python tests/golden/make_golden_moments.py

The file holds the particles of a two-species reflecting 2D deck after 3 steps (fields included) and every
moment array of every species selection, plus the particle state one HC_PUSH step later.  It pins the oracle
against drift; the CUDA path is compared with the live oracle in tests/test_moments.py / test_hc_push.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from epoch_b200 import deck as D  # noqa: E402
from oracle.oracle import FIELD_NAMES, Oracle  # noqa: E402
from tests import decks  # noqa: E402

NSTEPS = 3


def make_deck(hc=False):
    dk = decks.thermal(2, (12, 10), ppc=3, temp_k=1.5e9, bc="reflect", two_species=True, drift=(2.0e-23, 0.0, -1.0e-23))
    dk.hc_push = hc
    return dk


def moments_of(o):
    out = {}
    for kind in Oracle.MOMENTS:
        for isp in (-1, 0, 1):
            if kind.startswith("poynt") and isp >= 0:
                continue
            out[f"m_{kind}_{isp}"] = o.moment(0, kind, isp)
    return out


def main():
    dk = make_deck()
    o = Oracle(dk)
    o.auto_load()
    D.run(dk, o, [0], None, max_steps=NSTEPS)
    out = {f"p_{isp}": o.get_particles(0, isp) for isp in range(2)}
    for f in FIELD_NAMES:
        out[f] = np.array(o.field(0, f))
    out.update(moments_of(o))
    # one Higuera-Cary push from that state
    oh = Oracle(make_deck(hc=True))
    oh.init()
    for isp in range(2):
        oh.set_particles(0, isp, out[f"p_{isp}"])
    for f in FIELD_NAMES[:6]:
        oh.field(0, f)[...] = out[f]
    oh.push()
    for isp in range(2):
        out[f"hc_{isp}"] = oh.get_particles(0, isp)
    np.savez_compressed(os.path.join(HERE, "moments2d.npz"), **out)
    print("moments2d.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
