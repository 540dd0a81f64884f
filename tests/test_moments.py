"""Device-side diagnostics moments: calc_number_density, calc_charge_density,
calc_mass_density of io/calc_df.F90 (:689-757, :608-685, :35-110).

The reference holds no numbers for them.  The oracle restatement is pinned on (1) an independent numpy
deposit of the triangle weights (include/triangle/gxfac.inc), (2) the conservation the routines are built
for: sum(n) dV = sum of weights inside the domain, for periodic and reflecting boundaries, on one rank and
decomposed, (3) 1-rank vs 2x2-rank consistency.  The CUDA path (epb_calc_moment) is held to the oracle."""
import itertools

import numpy as np
import pytest

from oracle.oracle import Oracle
from tests import decks

KINDS = ("number_density", "charge_density", "mass_density")


def _interior(a):
    sl = tuple(slice(5, -5) if a.shape[ax] > 1 else slice(None) for ax in range(3))
    return a[sl]


def _numpy_density(dk, p, periodic=True):
    """Triangle-shape number density at cell centres on the global periodic grid."""
    nd = dk.ndims
    n = [dk.n[d] for d in range(nd)]
    out = np.zeros(n[::-1])
    cells, gs = [], []
    for d in range(nd):
        r = (p[:, d] - dk.grid_min(d)) / dk.dx(d)
        cx = np.floor(r + 0.5)
        f = cx - r
        cells.append(cx.astype(np.int64))
        gs.append([0.5 * (0.25 + f * f + f), 0.75 - f * f, 0.5 * (0.25 + f * f - f)])
    vol = np.prod([dk.dx(d) for d in range(nd)])
    for offs in itertools.product((-1, 0, 1), repeat=nd):
        wgt = p[:, -1] / vol
        idx = []
        for d in range(nd):
            wgt = wgt * gs[d][offs[d] + 1]
            idx.append((cells[d] + offs[d]) % n[d])
        np.add.at(out, tuple(idx[::-1]), wgt)
    return out


@pytest.mark.parametrize("ndims,n", [(1, (32,)), (2, (16, 12)), (3, (8, 7, 6))])
def test_number_density_against_numpy(ndims, n):
    dk = decks.thermal(ndims, n, ppc=5, temp_k=1.0e8)
    o = Oracle(dk)
    o.auto_load()
    o.init()
    for _ in range(2):
        o.push()
    got = _interior(o.moment(0, "number_density", 0)).reshape(n[::-1])
    ref = _numpy_density(dk, o.get_particles(0, 0))
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    s = dk.species[0]
    rho = _interior(o.moment(0, "charge_density", 0)).reshape(n[::-1])
    assert np.abs(rho - s.charge * ref).max() <= 1e-12 * np.abs(s.charge * ref).max()
    rm = _interior(o.moment(0, "mass_density", 0)).reshape(n[::-1])
    assert np.abs(rm - s.mass * ref).max() <= 1e-12 * np.abs(s.mass * ref).max()


@pytest.mark.parametrize("bc", ["periodic", "reflect"])
@pytest.mark.parametrize("nproc", [(1, 1, 1), (2, 2, 1)])
def test_total_weight_is_conserved(bc, nproc):
    dk = decks.thermal(2, (16, 12), ppc=5, temp_k=3.0e9, bc=bc, nproc=nproc, two_species=True)
    o = Oracle(dk)
    o.auto_load()
    o.init()
    for _ in range(4):
        o.push()
    vol = dk.dx(0) * dk.dx(1)
    for isp in (0, 1):
        total = sum(_interior(o.moment(rk, "number_density", isp)).sum() for rk in range(o.nranks)) * vol
        wsum = sum(o.get_particles(rk, isp)[:, -1].sum() for rk in range(o.nranks))
        assert abs(total - wsum) <= 1e-12 * wsum
    # all species: charge density of electrons + protons of equal density cancels in the mean
    rho = np.concatenate([_interior(o.moment(rk, "charge_density", -1)).ravel() for rk in range(o.nranks)])
    rho_e = np.concatenate([_interior(o.moment(rk, "charge_density", 0)).ravel() for rk in range(o.nranks)])
    assert abs(rho.sum()) <= 1e-9 * np.abs(rho_e).sum()


@pytest.mark.parametrize("bc", ["periodic", "reflect"])
def test_decomposed_equals_single_rank(bc):
    """The same particles (the loader seeds every rank differently, so they are handed over) on 1 and 2x2 ranks."""
    dk1 = decks.thermal(2, (16, 12), ppc=4, temp_k=1.0e9, bc=bc)
    o1 = Oracle(dk1)
    o1.auto_load()
    o1.init()
    p = o1.get_particles(0, 0)
    ref = _interior(o1.moment(0, "number_density", 0)).reshape(12, 16)
    dk4 = decks.thermal(2, (16, 12), ppc=4, temp_k=1.0e9, bc=bc, nproc=(2, 2, 1))
    o4 = Oracle(dk4)
    o4.init()
    full = np.zeros((12, 16))
    taken = 0
    for rk in range(o4.nranks):
        info = o4.rank_info(rk)
        sel = np.ones(len(p), dtype=bool)
        for d in range(2):
            sel &= (p[:, d] >= info["min_local"][d]) & (p[:, d] < info["max_local"][d])
        o4.set_particles(rk, 0, p[sel])
        taken += int(sel.sum())
    assert taken == len(p)
    for rk in range(o4.nranks):
        info = o4.rank_info(rk)
        a = _interior(o4.moment(rk, "number_density", 0)).reshape(info["n"][1], info["n"][0])
        x0, y0 = info["gmin"][0] - 1, info["gmin"][1] - 1
        full[y0:y0 + a.shape[0], x0:x0 + a.shape[1]] = a
    assert np.abs(full - ref).max() <= 1e-13 * np.abs(ref).max()


def test_mixed_boundaries_and_tracers():
    """c_bc_mixed takes the per-species calc_boundary path; a tracer species is left out of the species sum."""
    dk = decks.thermal(2, (16, 12), ppc=4, temp_k=3.0e9, bc="reflect", two_species=True)
    dk.species[1].bc_particle = ["open"] * 4
    dk.species[1].zero_current = True
    o = Oracle(dk)
    o.auto_load()
    o.init()
    for _ in range(3):
        o.push()
    both = o.moment(0, "number_density", -1)
    only_e = o.moment(0, "number_density", 0)
    assert np.array_equal(both, only_e)
    vol = dk.dx(0) * dk.dx(1)
    assert abs(_interior(only_e).sum() * vol - o.get_particles(0, 0)[:, -1].sum()) <= 1e-12 * o.get_particles(0, 0)[:, -1].sum()


# ---------------------------------------------------------------------------------------------------------
# CUDA path against the oracle
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("ndims,n,bc", [(1, (64,), "periodic"), (2, (32, 24), "periodic"), (2, (32, 24), "reflect"),
                                        (3, (10, 9, 8), "periodic"), (3, (10, 9, 8), "reflect")])
def test_moments_match_oracle_gpu(ndims, n, bc):
    from tests.gpu_util import make_pair, rel_l2, run_both
    dk = decks.thermal(ndims, n, ppc=5, temp_k=2.0e9, bc=bc, two_species=True)
    o, sim = make_pair(dk, strict=True, sort_interval=2)
    run_both(dk, o, sim, 3)
    for kind in KINDS:
        for isp in (-1, 0, 1):
            ref = o.moment(0, kind, isp)
            got = sim.moment(kind, isp)
            assert got.shape == ref.shape
            assert rel_l2(got, ref) <= 1e-13, (kind, isp)
            assert rel_l2(_interior(got), _interior(ref)) <= 1e-13, (kind, isp)
    # the diagnostic leaves the state alone: the next steps still match
    run_both(dk, o, sim, 2)
    assert rel_l2(sim.download_field("jx"), o.field(0, "jx")) <= 1e-12


@pytest.mark.gpu
def test_moments_mixed_boundaries_gpu():
    from tests.gpu_util import make_pair, rel_l2, run_both
    dk = decks.thermal(2, (32, 24), ppc=5, temp_k=2.0e9, bc="reflect", two_species=True)
    dk.species[1].bc_particle = ["open"] * 4
    o, sim = make_pair(dk, strict=True, sort_interval=2)
    run_both(dk, o, sim, 4)
    for isp in (-1, 0, 1):
        assert rel_l2(sim.moment("charge_density", isp), o.moment(0, "charge_density", isp)) <= 1e-13, isp


# ---------------------------------------------------------------------------------------------------------
# calc_ekbar (:116-221) and calc_temperature (:877-1128)
# ---------------------------------------------------------------------------------------------------------
def _numpy_weights(dk, p):
    nd = dk.ndims
    cells, gs = [], []
    for d in range(nd):
        r = (p[:, d] - dk.grid_min(d)) / dk.dx(d)
        cx = np.floor(r + 0.5)
        f = cx - r
        cells.append(cx.astype(np.int64))
        gs.append([0.5 * (0.25 + f * f + f), 0.75 - f * f, 0.5 * (0.25 + f * f - f)])
    return cells, gs


def _numpy_deposit(dk, p, values):
    """sum over particles of (triangle weight) * values on the periodic grid, no volume factor."""
    nd = dk.ndims
    n = [dk.n[d] for d in range(nd)]
    out = np.zeros(n[::-1])
    cells, gs = _numpy_weights(dk, p)
    for offs in itertools.product((-1, 0, 1), repeat=nd):
        wgt = np.array(values, dtype=np.float64, copy=True)
        idx = []
        for d in range(nd):
            wgt = wgt * gs[d][offs[d] + 1]
            idx.append((cells[d] + offs[d]) % n[d])
        np.add.at(out, tuple(idx[::-1]), wgt)
    return out


def _numpy_gather(dk, p, grid):
    """value of `grid` at each particle's 3^nd stencil points: list of (weight, grid value)."""
    nd = dk.ndims
    n = [dk.n[d] for d in range(nd)]
    cells, gs = _numpy_weights(dk, p)
    out = []
    for offs in itertools.product((-1, 0, 1), repeat=nd):
        wgt = np.ones(len(p))
        idx = []
        for d in range(nd):
            wgt = wgt * gs[d][offs[d] + 1]
            idx.append((cells[d] + offs[d]) % n[d])
        out.append((wgt, grid[tuple(idx[::-1])], tuple(idx[::-1])))
    return out


@pytest.mark.parametrize("ndims,n", [(1, (32,)), (2, (16, 12)), (3, (8, 7, 6))])
def test_ekbar_and_temperature_against_numpy(ndims, n):
    from epoch_b200 import deck as D
    temp_k = 5.0e8
    dk = decks.thermal(ndims, n, ppc=6, temp_k=temp_k)
    o = Oracle(dk)
    o.auto_load()
    o.init()
    o.push()
    p = o.get_particles(0, 0)
    s = dk.species[0]
    w = p[:, -1]
    mom = p[:, ndims:ndims + 3]
    u2 = ((mom / (s.mass * D.c)) ** 2).sum(axis=1)
    ek = u2 / (np.sqrt(u2 + 1.0) + 1.0) * s.mass * D.c ** 2
    ref = _numpy_deposit(dk, p, ek * w) / _numpy_deposit(dk, p, w)
    got = _interior(o.moment(0, "ekbar", 0)).reshape(n[::-1])
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    # temperature: weighted mean per cell, then un-weighted spread around the mean of each stencil cell
    pm = mom / np.sqrt(s.mass)
    cnt = _numpy_deposit(dk, p, w)
    for name, comps in (("temperature", (0, 1, 2)), ("temperature_x", (0,)), ("temperature_y", (1,)), ("temperature_z", (2,))):
        sig = np.zeros(n[::-1])
        for q in comps:
            mean = _numpy_deposit(dk, p, w * pm[:, q]) / cnt
            for wgt, mval, idx in _numpy_gather(dk, p, mean):
                np.add.at(sig, idx, wgt * (pm[:, q] - mval) ** 2)
        ref_t = sig / _numpy_deposit(dk, p, np.ones(len(p))) / D.kb / len(comps)
        got_t = _interior(o.moment(0, name, 0)).reshape(n[::-1])
        assert np.abs(got_t - ref_t).max() <= 1e-11 * np.abs(ref_t).max(), name
    # physics: the loader's Maxwellian has the deck temperature (6 ppc: cell values scatter, the mean does not)
    t_mean = _interior(o.moment(0, "temperature", 0)).mean()
    assert abs(t_mean - temp_k) <= 0.12 * temp_k


@pytest.mark.parametrize("bc", ["periodic", "reflect"])
def test_ekbar_temperature_decomposed_equals_single_rank(bc):
    dk1 = decks.thermal(2, (16, 12), ppc=4, temp_k=1.0e9, bc=bc)
    o1 = Oracle(dk1)
    o1.auto_load()
    o1.init()
    p = o1.get_particles(0, 0)
    dk4 = decks.thermal(2, (16, 12), ppc=4, temp_k=1.0e9, bc=bc, nproc=(2, 2, 1))
    o4 = Oracle(dk4)
    o4.init()
    for rk in range(o4.nranks):
        info = o4.rank_info(rk)
        sel = np.ones(len(p), dtype=bool)
        for d in range(2):
            sel &= (p[:, d] >= info["min_local"][d]) & (p[:, d] < info["max_local"][d])
        o4.set_particles(rk, 0, p[sel])
    for kind in ("ekbar", "temperature", "temperature_y"):
        ref = _interior(o1.moment(0, kind, 0)).reshape(12, 16)
        full = np.zeros((12, 16))
        for rk in range(o4.nranks):
            info = o4.rank_info(rk)
            a = _interior(o4.moment(rk, kind, 0)).reshape(info["n"][1], info["n"][0])
            x0, y0 = info["gmin"][0] - 1, info["gmin"][1] - 1
            full[y0:y0 + a.shape[0], x0:x0 + a.shape[1]] = a
        assert np.abs(full - ref).max() <= 1e-12 * np.abs(ref).max(), kind


@pytest.mark.gpu
@pytest.mark.parametrize("ndims,n,bc", [(1, (64,), "periodic"), (2, (32, 24), "periodic"), (2, (32, 24), "reflect"),
                                        # extents that are no multiple of the 16 x 8-cell tiles: padding columns of k_moment2_slots
                                        (2, (21, 13), "periodic"), (2, (19, 11), "reflect"),
                                        (3, (10, 9, 8), "periodic"), (3, (10, 9, 8), "reflect")])
def test_ekbar_temperature_match_oracle_gpu(ndims, n, bc):
    from tests.gpu_util import make_pair, rel_l2, run_both
    dk = decks.thermal(ndims, n, ppc=5, temp_k=2.0e9, bc=bc, two_species=True)
    o, sim = make_pair(dk, strict=True, sort_interval=2)
    run_both(dk, o, sim, 3)
    for kind in ("ekbar", "temperature", "temperature_x", "temperature_y", "temperature_z"):
        for isp in (-1, 0, 1):
            ref = o.moment(0, kind, isp)
            got = sim.moment(kind, isp)
            assert rel_l2(_interior(got), _interior(ref)) <= 1e-12, (kind, isp)
            assert rel_l2(got, ref) <= 1e-12, (kind, isp)


@pytest.mark.gpu
def test_ekbar_temperature_mixed_boundaries_gpu():
    from tests.gpu_util import make_pair, rel_l2, run_both
    dk = decks.thermal(2, (32, 24), ppc=5, temp_k=2.0e9, bc="reflect", two_species=True)
    dk.species[1].bc_particle = ["open"] * 4
    o, sim = make_pair(dk, strict=True, sort_interval=2)
    run_both(dk, o, sim, 4)
    for kind in ("ekbar", "temperature"):
        for isp in (-1, 0, 1):
            assert rel_l2(sim.moment(kind, isp), o.moment(0, kind, isp)) <= 1e-12, (kind, isp)


# ---------------------------------------------------------------------------------------------------------
# calc_ekflux (:415-557), calc_average_momentum (:1239-1317), calc_per_species_current (:1132-1235),
# calc_average_weight (:811-873)
# ---------------------------------------------------------------------------------------------------------
EXTRA_KINDS = ("ekflux_xm", "ekflux_xp", "ekflux_ym", "ekflux_yp", "ekflux_zm", "ekflux_zp",
               "average_px", "average_py", "average_pz", "jx", "jy", "jz", "average_weight")


@pytest.mark.parametrize("ndims,n", [(1, (32,)), (2, (16, 12)), (3, (8, 7, 6))])
def test_flux_momentum_current_weight_against_numpy(ndims, n):
    from epoch_b200 import deck as D
    dk = decks.thermal(ndims, n, ppc=6, temp_k=2.0e9, drift=(3.0e-23, -1.0e-23, 2.0e-23))
    o = Oracle(dk)
    o.auto_load()
    o.init()
    o.push()
    p = o.get_particles(0, 0)
    s = dk.species[0]
    w = p[:, -1]
    mom = p[:, ndims:ndims + 3]
    u = mom / (s.mass * D.c)
    u2 = (u ** 2).sum(axis=1)
    gamma = np.sqrt(u2 + 1.0)
    ek = u2 / (gamma + 1.0) * s.mass * D.c ** 2
    wt = _numpy_deposit(dk, p, w)
    d = [dk.dx(q) for q in range(ndims)] + [1.0] * (3 - ndims)
    if ndims == 1:
        area = [1.0, d[0], d[0]]
    elif ndims == 2:
        area = [d[1], d[0], d[0] * d[1]]
    else:
        area = [d[1] * d[2], d[0] * d[2], d[0] * d[1]]
    for a, ax in enumerate("xyz"):
        flux = D.c * area[a] * u[:, a] / gamma
        for sign, tag in ((-1, "m"), (+1, "p")):
            ref = _numpy_deposit(dk, p, ek * w * np.maximum(sign * flux, 0.0)) / wt
            got = _interior(o.moment(0, f"ekflux_{ax}{tag}", 0)).reshape(n[::-1])
            assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), (ax, tag)
        ref = _numpy_deposit(dk, p, w * mom[:, a]) / wt
        got = _interior(o.moment(0, f"average_p{ax}", 0)).reshape(n[::-1])
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), ax
        vol = np.prod(d[:ndims])
        ref = _numpy_deposit(dk, p, s.charge * w * D.c * mom[:, a] / np.sqrt((s.mass * D.c) ** 2 + (mom ** 2).sum(axis=1))) / vol
        got = _interior(o.moment(0, f"j{ax}", 0)).reshape(n[::-1])
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), ax
    # weighted back with the cell weights, the mean momentum per cell sums to the particles' total momentum
    tot = (_interior(o.moment(0, "average_px", 0)).reshape(n[::-1]) * wt).sum()
    assert abs(tot / (w * mom[:, 0]).sum() - 1.0) < 1e-12
    # average weight: every particle of this deck has the same weight
    aw = _interior(o.moment(0, "average_weight", 0))
    cnt = o.cell_counts(0, 0)
    assert np.allclose(aw[cnt.reshape(aw.shape) > 0], w[0], rtol=1e-14)
    assert np.all(aw[cnt.reshape(aw.shape) == 0] == 0.0)


def test_species_current_sums_to_deposited_current_scale():
    """jx of calc_per_species_current is the instantaneous q n v; summed over the box it equals sum(q w v) / dV."""
    from epoch_b200 import deck as D
    dk = decks.thermal(2, (16, 12), ppc=6, temp_k=1.0e9, drift=(5.0e-23, 0.0, 0.0))
    o = Oracle(dk)
    o.auto_load()
    o.init()
    p = o.get_particles(0, 0)
    s = dk.species[0]
    v = D.c * p[:, 2] / np.sqrt((s.mass * D.c) ** 2 + (p[:, 2:5] ** 2).sum(axis=1))
    total = _interior(o.moment(0, "jx", 0)).sum() * dk.dx(0) * dk.dx(1)
    assert abs(total / (s.charge * (p[:, -1] * v).sum()) - 1.0) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("ndims,n,bc", [(1, (64,), "periodic"), (2, (32, 24), "reflect"), (2, (21, 13), "periodic"),
                                        (3, (10, 9, 8), "periodic")])
def test_flux_momentum_current_weight_match_oracle_gpu(ndims, n, bc):
    from tests.gpu_util import make_pair, rel_l2, run_both
    dk = decks.thermal(ndims, n, ppc=5, temp_k=2.0e9, bc=bc, two_species=True, drift=(3.0e-23, -1.0e-23, 2.0e-23))
    o, sim = make_pair(dk, strict=True, sort_interval=2)
    run_both(dk, o, sim, 3)
    for kind in EXTRA_KINDS:
        for isp in (-1, 0, 1):
            ref = o.moment(0, kind, isp)
            got = sim.moment(kind, isp)
            assert rel_l2(got, ref) <= 1e-12, (kind, isp)
            assert rel_l2(_interior(got), _interior(ref)) <= 1e-12, (kind, isp)


@pytest.mark.gpu
def test_flux_momentum_current_weight_mixed_boundaries_gpu():
    from tests.gpu_util import make_pair, rel_l2, run_both
    dk = decks.thermal(2, (32, 24), ppc=5, temp_k=2.0e9, bc="reflect", two_species=True)
    dk.species[1].bc_particle = ["open"] * 4
    o, sim = make_pair(dk, strict=True, sort_interval=2)
    run_both(dk, o, sim, 4)
    for kind in ("ekflux_xp", "ekflux_ym", "average_pz", "jy", "average_weight"):
        for isp in (-1, 0, 1):
            assert rel_l2(sim.moment(kind, isp), o.moment(0, kind, isp)) <= 1e-12, (kind, isp)


# ---------------------------------------------------------------------------------------------------------
# calc_poynt_flux (:561-604; epoch1d :441-474; epoch3d :585-650)
# ---------------------------------------------------------------------------------------------------------
def _cc(a, ndims, stag_axes):
    """Cell-centred value on the interior: mean over the active axes the component is staggered along.
    a is indexed [k][j][i] with 5 ghost cells on active axes; axis d lives in numpy axis 2 - d."""
    out = a
    n_avg = 0
    for d in stag_axes:
        if d < ndims:
            out = out + np.roll(out, 1, axis=2 - d)
            n_avg += 1
    return _interior(out * (0.5 ** n_avg))


@pytest.mark.parametrize("ndims,n", [(1, (32,)), (2, (16, 12)), (3, (8, 7, 6))])
def test_poynt_flux_against_numpy(ndims, n):
    dk = decks.thermal(ndims, n, ppc=1)
    o = Oracle(dk)
    o.init()
    rng = np.random.default_rng(11)
    f = {}
    for name in ("ex", "ey", "ez", "bx", "by", "bz"):
        a = o.field(0, name)
        a[...] = rng.normal(size=a.shape) * (1e9 if name[0] == "e" else 3.0)
        f[name] = a.copy()
    mu0 = 4.0e-7 * np.pi
    cc = {"ex": _cc(f["ex"], ndims, (0,)), "ey": _cc(f["ey"], ndims, (1,)), "ez": _cc(f["ez"], ndims, (2,)),
          "bx": _cc(f["bx"], ndims, (1, 2)), "by": _cc(f["by"], ndims, (0, 2)), "bz": _cc(f["bz"], ndims, (0, 1))}
    ref = {"x": (cc["ey"] * cc["bz"] - cc["ez"] * cc["by"]) / mu0,
           "y": (cc["ez"] * cc["bx"] - cc["ex"] * cc["bz"]) / mu0,
           "z": (cc["ex"] * cc["by"] - cc["ey"] * cc["bx"]) / mu0}
    for ax in "xyz":
        full = o.moment(0, f"poynt_flux_{ax}", -1)
        got = _interior(full)
        assert np.abs(got - ref[ax]).max() <= 1e-13 * np.abs(ref[ax]).max(), ax
        assert np.abs(full).sum() == np.abs(got).sum()          # ghost cells zero


def test_poynt_flux_of_a_plane_wave():
    """E = E0 y, B = E0/c z  =>  S = E0^2 / (mu0 c) along +x, nothing along y and z."""
    from epoch_b200 import deck as D
    dk = decks.thermal(2, (16, 12), ppc=1)
    o = Oracle(dk)
    o.init()
    e0 = 2.0e9
    o.field(0, "ey")[...] = e0
    o.field(0, "bz")[...] = e0 / D.c
    sx = _interior(o.moment(0, "poynt_flux_x", -1))
    assert np.allclose(sx, e0 ** 2 / (4.0e-7 * np.pi * D.c), rtol=1e-14)
    assert np.all(_interior(o.moment(0, "poynt_flux_y", -1)) == 0.0)
    assert np.all(_interior(o.moment(0, "poynt_flux_z", -1)) == 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("ndims,n", [(1, (64,)), (2, (32, 24)), (3, (10, 9, 8))])
def test_poynt_flux_matches_oracle_gpu(ndims, n):
    """No atomics on this path: bit-exact."""
    from tests.gpu_util import make_pair, set_random_fields
    dk = decks.thermal(ndims, n, ppc=2)
    o, sim = make_pair(dk, strict=True)
    set_random_fields(o, sim, dk, seed=4)
    for ax in "xyz":
        assert np.array_equal(sim.moment(f"poynt_flux_{ax}", -1), o.moment(0, f"poynt_flux_{ax}", -1)), ax
