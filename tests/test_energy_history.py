"""Long-run invariant of BASELINE.json's north star: "total energy history ... within 1 %".

calc_total_energy_sum (io/calc_df.F90:1321-1417) restated in numpy: field energy 0.5 eps0 sum(E^2) dV +
0.5/mu0 sum(B^2) dV over the interior, kinetic energy sum w (gamma - 1) m c^2 with the cancellation-free
(gamma - 1) = u^2 / (gamma + 1).  A thermal plasma at dx = Debye length keeps ~0.25 % of its energy in the
fluctuating fields; the total must stay put far below that share while the two parts exchange energy."""
import numpy as np

from epoch_b200 import deck as D
from oracle.oracle import Oracle
from tests import decks

MU0 = 4.0e-7 * np.pi


def energies(dk, fields, particles):
    """fields: name -> interior array; particles: list of packed arrays per species."""
    nd = dk.ndims
    dv = np.prod([dk.dx(d) for d in range(nd)])
    fe = 0.5 * D.epsilon0 * sum(float((fields[f] ** 2).sum()) for f in ("ex", "ey", "ez")) * dv
    fb = 0.5 / MU0 * sum(float((fields[f] ** 2).sum()) for f in ("bx", "by", "bz")) * dv
    ke = []
    for s, p in zip(dk.species, particles):
        u2 = (p[:, nd:nd + 3] ** 2).sum(axis=1) / (s.mass * D.c) ** 2
        ke.append(float((p[:, -1] * u2 / (np.sqrt(u2 + 1.0) + 1.0)).sum()) * s.mass * D.c ** 2)
    return fe, fb, ke


def test_total_energy_history_oracle():
    dk = decks.thermal(2, (24, 24), ppc=16, temp_k=1.0e7)
    o = Oracle(dk)
    o.auto_load()
    o.init()
    hist = []
    for s in range(300):
        o.fields_half(); o.push(); o.current_finish(); o.fields_final()
        if s % 30 == 29:
            fe, fb, ke = energies(dk, {f: o.interior(0, f) for f in ("ex", "ey", "ez", "bx", "by", "bz")},
                                  [o.get_particles(0, 0)])
            hist.append((fe, fb, ke[0]))
    h = np.array(hist)
    tot = h.sum(axis=1)
    assert np.abs(tot / tot[0] - 1.0).max() < 1.0e-4               # 100x inside the north star's 1 %
    share = h[:, 0] / tot
    assert share.min() > 1.0e-3 and share.max() - share.min() > 3.0e-4   # the fields do trade energy with the particles
