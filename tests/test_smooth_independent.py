"""smooth_array (housekeeping/current_smooth.F90:61-141, the strided compensated binomial filter of smooth_currents)
once more, in numpy from the Fortran, through current_finish on one periodic rank: the work array with its own ghost
depth, field_bc before every stride, alpha = 1/2 and beta = 1/8 -- the reference resets alpha only AFTER the
compensation pass has run and never recomputes beta, which is restated as it stands -- against the oracle bit for bit."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from tests import decks

NG = 5


def smooth_array_numpy(a, nx, ny, its, comp_its, strides):
    ng_l = max(max(strides), NG)
    alpha = 0.5
    beta = (1.0 - alpha) * 0.25
    wk = np.zeros((ny + 2 * ng_l, nx + 2 * ng_l))
    o = ng_l - NG
    wk[o:o + ny + 2 * NG, o:o + nx + 2 * NG] = a
    W = lambda i0, i1, j0, j1: wk[j0 + ng_l - 1:j1 + ng_l, i0 + ng_l - 1:i1 + ng_l]
    out = a.copy()
    for it in range(1, its + comp_its + 1):
        for c in strides:
            # field_bc(wk_array, ng_l): periodic in x, then in y (whole rows, so the corners follow)
            W(nx + 1, nx + ng_l, 1 - ng_l, ny + ng_l)[...] = W(1, ng_l, 1 - ng_l, ny + ng_l)
            W(1 - ng_l, 0, 1 - ng_l, ny + ng_l)[...] = W(nx + 1 - ng_l, nx, 1 - ng_l, ny + ng_l)
            W(1 - ng_l, nx + ng_l, ny + 1, ny + ng_l)[...] = W(1 - ng_l, nx + ng_l, 1, ng_l)
            W(1 - ng_l, nx + ng_l, 1 - ng_l, 0)[...] = W(1 - ng_l, nx + ng_l, ny + 1 - ng_l, ny)
            new = alpha * W(1, nx, 1, ny) + (W(1 - c, nx - c, 1, ny) + W(1 + c, nx + c, 1, ny)
                                              + W(1, nx, 1 - c, ny - c) + W(1, nx, 1 + c, ny + c)) * beta
            W(1, nx, 1, ny)[...] = new
        if it > its:
            alpha = float(its) * 0.5 + 1.0
    out[NG:NG + ny, NG:NG + nx] = W(1, nx, 1, ny)
    return out


@pytest.mark.parametrize("its,comp,strides", [(1, False, (1,)), (2, True, (1, 2)), (1, True, (1, 2, 3, 4)), (3, False, (2,))])
def test_smoothing_equals_an_independent_restatement(its, comp, strides):
    n = (13, 11)
    dk = decks.thermal(2, n, ppc=1)
    dk.smooth_currents, dk.smooth_iterations, dk.smooth_compensation, dk.smooth_strides = True, its, comp, strides
    o = Oracle(dk)
    o.init()
    rng = np.random.default_rng(7)
    want = []
    for name in ("jx", "jy", "jz"):
        a = o.field(0, name)
        a[...] = 0.0
        a[0, NG:-NG, NG:-NG] = rng.standard_normal((n[1], n[0]))     # ghost currents zero: current_bcs adds nothing
        # current_finish = current_bcs, field_bc(j), smooth_array: the ghost cells it sees are the periodic images
        b = a[0].copy()
        b[:, n[0] + NG:] = b[:, NG:2 * NG]; b[:, :NG] = b[:, n[0]:n[0] + NG]
        b[n[1] + NG:, :] = b[NG:2 * NG, :]; b[:NG, :] = b[n[1]:n[1] + NG, :]
        want.append(smooth_array_numpy(b, n[0], n[1], its, 1 if comp else 0, strides))
    o.current_finish()
    for name, w in zip(("jx", "jy", "jz"), want):
        got = o.field(0, name)[0]
        assert np.array_equal(got[NG:-NG, NG:-NG], w[NG:-NG, NG:-NG]), name
