"""Deck builders shared by the tests: the reference's own test decks restated as
`Deck` objects (file:line cited per builder), plus small particle decks."""
import math

import numpy as np

from epoch_b200 import deck as D

lambda0 = 1.06 * D.micron
theta = D.pi / 8.0


def laser1d(nproc=(1, 1, 1)):
    """epoch1d/tests/laser/input.deck"""
    lam = 1 * D.micron
    las = D.Laser("x_min", D.Laser.amp_from_intensity_w_cm2(1.0e15), 2 * D.pi * D.c / lam,
                  t_profile=lambda t: D.gauss(t, 4 * D.femto, 4 * D.femto), t_end=14 * D.femto)
    return D.Deck(1, [200], [-4 * D.micron], [4 * D.micron], ["simple_laser", "open"], lasers=[las],
                  t_end=50 * D.femto, dt_snapshot=8 * D.femto, nproc=nproc)


def laser2d(nproc=(1, 1, 1), n=500):
    """epoch2d/tests/laser/input.deck"""
    lam = lambda0 * math.cos(theta)
    las = D.Laser("x_min", D.Laser.amp_from_intensity_w_cm2(1.0e15), 2 * D.pi * D.c / lam,
                  phase=lambda y, z: -2.0 * D.pi * y * math.tan(theta) / lambda0,
                  profile=lambda y, z: D.gauss(y, 0, 4 * D.micron))
    return D.Deck(2, [n, n], [-10 * D.micron] * 2, [10 * D.micron] * 2,
                  ["simple_laser", "open", "periodic", "periodic"], lasers=[las],
                  t_end=50 * D.femto, dt_snapshot=25 * D.femto, nproc=nproc)


def laser2d_y(nproc=(1, 1, 1), n=64, side="y_min"):
    """The 2D laser deck with the laser on a y face (outflow_bcs_y_min/max, laser.f90:462-610), oblique
    in x, circular-ish polarisation so that both source arrays are exercised."""
    lam = lambda0 * math.cos(theta)
    las = D.Laser(side, D.Laser.amp_from_intensity_w_cm2(1.0e15), 2 * D.pi * D.c / lam, pol_angle=0.6,
                  phase=lambda x, z: -2.0 * D.pi * x * math.tan(theta) / lambda0,
                  profile=lambda x, z: D.gauss(x, 0, 4 * D.micron))
    ybc = ["simple_laser", "open"] if side == "y_min" else ["open", "simple_laser"]
    return D.Deck(2, [n, n + 8], [-10 * D.micron] * 2, [10 * D.micron] * 2,
                  ["periodic", "periodic"] + ybc, lasers=[las],
                  t_end=50 * D.femto, dt_snapshot=25 * D.femto, nproc=nproc)


def laser3d(nproc=(1, 1, 1), n=140):
    """epoch3d/tests/laser/input.deck"""
    lam = lambda0 * math.cos(theta)
    las = D.Laser("x_min", D.Laser.amp_from_intensity_w_cm2(1.0e15), 2 * D.pi * D.c / lam,
                  phase=lambda y, z: -2.0 * D.pi * y * math.tan(theta) / lambda0,
                  profile=lambda y, z: D.gauss(np.sqrt(y * y + z * z), 0, 4 * D.micron))
    return D.Deck(3, [n] * 3, [-10 * D.micron] * 3, [10 * D.micron] * 3,
                  ["simple_laser", "open"] + ["periodic"] * 4, lasers=[las],
                  t_end=50 * D.femto, dt_snapshot=25 * D.femto, nproc=nproc)


def laser3d_face(face, nproc=(1, 1, 1), n=140):
    """epoch3d/tests/laser/input.deck with the axes cyclically permuted so that the laser enters through
    y_min ('y': old (x, y, z) -> new (y, z, x)) or z_min ('z': old (x, y, z) -> new (z, x, y)).  The
    physics is identical, so the reference's golden sum(Ex^2) must reappear as sum(Ey^2) resp. sum(Ez^2):
    this pins the y- and z-face laser / outflow boundaries on the reference's own numbers."""
    lam = lambda0 * math.cos(theta)
    prof = lambda u, v: D.gauss(np.sqrt(u * u + v * v), 0, 4 * D.micron)
    if face == "y":    # callables get (x, z) = (old z, old y)
        las = D.Laser("y_min", D.Laser.amp_from_intensity_w_cm2(1.0e15), 2 * D.pi * D.c / lam,
                      phase=lambda u, v: -2.0 * D.pi * v * math.tan(theta) / lambda0, profile=prof)
        bcs = ["periodic"] * 2 + ["simple_laser", "open"] + ["periodic"] * 2
    else:              # callables get (x, y) = (old y, old z)
        las = D.Laser("z_min", D.Laser.amp_from_intensity_w_cm2(1.0e15), 2 * D.pi * D.c / lam,
                      phase=lambda u, v: -2.0 * D.pi * u * math.tan(theta) / lambda0, profile=prof)
        bcs = ["periodic"] * 4 + ["simple_laser", "open"]
    return D.Deck(3, [n] * 3, [-10 * D.micron] * 3, [10 * D.micron] * 3, bcs, lasers=[las],
                  t_end=50 * D.femto, dt_snapshot=25 * D.femto, nproc=nproc)


def thermal(ndims, n, ppc=8, nproc=(1, 1, 1), temp_k=1.0e7, density=1.0e25, bc="periodic",
            drift=(0.0, 0.0, 0.0), two_species=False, length=None, nsteps=10, seed=7842432):
    """Uniform thermal plasma (BASELINE.md C2/C4 shape, down-scaled).  dx ~ Debye length."""
    debye = math.sqrt(D.epsilon0 * D.kb * temp_k / (density * D.q0 ** 2))
    dx = debye if length is None else length / n[0]
    xmin = [0.0] * ndims
    xmax = [dx * n[d] for d in range(ndims)]
    sp = [D.Species("electron", -D.q0, D.m0, npart_per_cell=ppc, density=density,
                    temp=(temp_k,) * 3, drift=drift)]
    if two_species:
        sp.append(D.Species("proton", D.q0, 1836.2 * D.m0, npart_per_cell=ppc, density=density,
                            temp=(temp_k,) * 3))
    bcs = [bc] * (2 * ndims) if isinstance(bc, str) else list(bc)
    return D.Deck(ndims, list(n), xmin, xmax, bcs, species=sp, nsteps=nsteps,
                  nproc=nproc, seed=seed)


def twostream1d(nx=400, ppc_per_beam=4, nproc=(1, 1, 1), nsteps=-1, t_end=0.15):
    """epoch1d/tests/twostream/input.deck (BASELINE.md C1 uses nx=1600, 50 ppc per beam)"""
    sp = [D.Species("Left", -D.q0, D.m0, npart_per_cell=ppc_per_beam, density=10.0,
                    temp=(273.0, 0.0, 0.0), drift=(2.5e-24, 0.0, 0.0)),
          D.Species("Right", -D.q0, D.m0, npart_per_cell=ppc_per_beam, density=10.0,
                    temp=(273.0, 0.0, 0.0), drift=(-2.5e-24, 0.0, 0.0))]
    return D.Deck(1, [nx], [0.0], [5.0e5], ["periodic", "periodic"], species=sp, t_end=t_end,
                  nsteps=nsteps, nproc=nproc)


def foil2d(n=(96, 64), ppc=4, nproc=(1, 1, 1), nsteps=20, intensity=1.0e18):
    """BASELINE.md C3 shape, down-scaled: laser on x_min, outflow on x_max, y periodic,
    overdense e-/p+ slab (after epoch2d/example_decks/ramp.deck)."""
    lam = 1 * D.micron
    omega = 2 * D.pi * D.c / lam
    ncrit = omega ** 2 * D.epsilon0 * D.m0 / D.q0 ** 2
    L = [6 * D.micron, 4 * D.micron]
    las = D.Laser("x_min", D.Laser.amp_from_intensity_w_cm2(intensity), omega,
                  profile=lambda y, z: D.gauss(y, 0, 1.0 * D.micron),
                  t_profile=lambda t: D.gauss(t, 8 * D.femto, 4 * D.femto) if t < 8 * D.femto else 1.0)
    box_lo = (1.5 * D.micron, -1e300, -1e300)
    box_hi = (3.0 * D.micron, 1e300, 1e300)
    sp = [D.Species("electron", -D.q0, D.m0, npart_per_cell=ppc, density=4 * ncrit,
                    temp=(1.0e6,) * 3, box_lo=box_lo, box_hi=box_hi,
                    bc_particle=["open", "open", "periodic", "periodic"]),
          D.Species("proton", D.q0, 1836.2 * D.m0, npart_per_cell=ppc, density=4 * ncrit,
                    temp=(1.0e6,) * 3, box_lo=box_lo, box_hi=box_hi,
                    bc_particle=["open", "open", "periodic", "periodic"])]
    return D.Deck(2, list(n), [0.0, -L[1] / 2], [L[0], L[1] / 2],
                  ["simple_laser", "simple_outflow", "periodic", "periodic"], species=sp,
                  lasers=[las], nsteps=nsteps, nproc=nproc, t_end=1.0)
