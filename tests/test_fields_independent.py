"""The field half of the step once more, in numpy straight from the Fortran of epoch2d: update_e_field /
update_b_field (fields.f90:206-225, :427-439, order 2, Yee), update_eb_fields_half / _final (:533-582), efield_bcs /
bfield_bcs / bfield_final_bcs (boundary.F90:808-944) with field_clamp_zero (:473-530) and the ghost-cell copy of a
periodic axis, outflow_bcs_x_min / x_max (laser.f90:310-458) and the start-up sequence of epoch2d.F90:144-162 (dt halved
for the first bfield_final_bcs).  The reference's golden sums pin these routines to 1e-5 only (np.isclose in its
tests); against this restatement the oracle is equal bit for bit, ghost cells included, on the reference's own 2D
laser deck shape (laser on x_min, outflow on x_max, periodic y)."""
import numpy as np

from epoch_b200 import deck as D
from oracle.oracle import Oracle
from tests import decks

NG = 5
STAG = {"ex": (1, 0), "ey": (0, 1), "ez": (0, 0), "bx": (0, 1), "by": (1, 0), "bz": (1, 1)}   # staggered in x, y


class NumpyFields:
    """backend of deck.run for a vacuum run on one rank"""

    def __init__(self, dk):
        self.dk = dk
        self.nx, self.ny = dk.n
        shape = (self.ny + 2 * NG, self.nx + 2 * NG)
        self.f = {k: np.zeros(shape) for k in ("ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz")}
        self.src = {}
        self.dt = dk.dt()

    def S(self, name, i0, i1, j0, j1):
        """Fortran section name(i0:i1, j0:j1)"""
        return self.f[name][j0 + NG - 1:j1 + NG, i0 + NG - 1:i1 + NG]

    def set_laser_source(self, lr, side, s1, s2):
        self.src[side] = (np.asarray(s1, dtype=np.float64).copy(), np.asarray(s2, dtype=np.float64).copy())

    # -- ghost cells ------------------------------------------------------------------------
    def periodic(self, d):
        return self.dk.bc_field[2 * d] == "periodic"

    def field_bc(self, name):
        a, nx, ny = self.f[name], self.nx, self.ny     # a periodic axis copies from its other end, x before y
        if self.periodic(0):
            a[:, nx + NG:nx + 2 * NG] = a[:, NG:2 * NG]
            a[:, 0:NG] = a[:, nx:nx + NG]
        if self.periodic(1):
            a[ny + NG:ny + 2 * NG, :] = a[NG:2 * NG, :]
            a[0:NG, :] = a[ny:ny + NG, :]

    def clamp_zero_y(self, name, side):
        a, nn = self.f[name], self.ny
        F = lambda i: i + NG - 1
        if side == 0:
            if STAG[name][1]:
                for i in range(1, NG):
                    a[F(i - NG), :] = -a[F(NG - i), :]
                a[F(0), :] = 0.0
            else:
                for i in range(1, NG + 1):
                    a[F(i - NG), :] = -a[F(NG + 1 - i), :]
        else:
            if STAG[name][1]:
                a[F(nn), :] = 0.0
                for i in range(1, NG):
                    a[F(nn + i), :] = -a[F(nn - i), :]
            else:
                for i in range(1, NG + 1):
                    a[F(nn + i), :] = -a[F(nn + 1 - i), :]

    def clamp(self, names):
        # DO i = 1, 2*c_ndims: x_min, x_max, y_min, y_max; simple_laser and simple_outflow both clamp
        for b in range(4):
            if self.periodic(b // 2):
                continue
            for k in names:
                (self.clamp_zero_x if b < 2 else self.clamp_zero_y)(k, b % 2)

    def clamp_zero_x(self, name, side):
        a, nn = self.f[name], self.nx
        F = lambda i: a[:, i + NG - 1]
        if side == 0:
            if STAG[name][0]:
                for i in range(1, NG):
                    a[:, i - NG + NG - 1] = -F(NG - i)
                a[:, 0 + NG - 1] = 0.0
            else:
                for i in range(1, NG + 1):
                    a[:, i - NG + NG - 1] = -F(NG + 1 - i)
        else:
            if STAG[name][0]:
                a[:, nn + NG - 1] = 0.0
                for i in range(1, NG):
                    a[:, nn + i + NG - 1] = -F(nn - i)
            else:
                for i in range(1, NG + 1):
                    a[:, nn + i + NG - 1] = -F(nn + 1 - i)

    def efield_bcs(self):
        for k in ("ex", "ey", "ez"):
            self.field_bc(k)
        self.clamp(("ex", "ey", "ez"))

    def bfield_bcs(self, mpi_only):
        for k in ("bx", "by", "bz"):
            self.field_bc(k)
        if mpi_only:
            return
        self.clamp(("bx", "by", "bz"))

    # -- updates ----------------------------------------------------------------------------
    def coeffs(self):
        hdt = 0.5 * self.dt
        hdtx, hdty = hdt / self.dk.dx(0), hdt / self.dk.dx(1)
        return hdtx, hdty, hdtx * (D.c * D.c), hdty * (D.c * D.c), hdt / D.epsilon0

    # fields.f90:206-294 (E) and :427-529 (B): orders 2, 4, 6 differ in the number of difference terms only
    FD = {2: (1.0,), 4: (9.0 / 8.0, -1.0 / 24.0), 6: (75.0 / 64.0, -25.0 / 384.0, 3.0 / 640.0)}

    def update_e_field(self):
        nx, ny, S = self.nx, self.ny, self.S
        _, _, cnx, cny, fac = self.coeffs()
        cs = self.FD[int(getattr(self.dk, "field_order", 2))]
        cx, cy = [c * cnx for c in cs], [c * cny for c in cs]
        dxm = lambda name, k: S(name, k - 1, nx + k - 1, 0, ny) - S(name, -k, nx - k, 0, ny)     # f(ix+k-1) - f(ix-k)
        dym = lambda name, k: S(name, 0, nx, k - 1, ny + k - 1) - S(name, 0, nx, -k, ny - k)
        ex, ey, ez = S("ex", 0, nx, 0, ny).copy(), S("ey", 0, nx, 0, ny).copy(), S("ez", 0, nx, 0, ny).copy()
        for k in range(1, len(cs) + 1):
            ex = ex + cy[k - 1] * dym("bz", k)
            ey = ey - cx[k - 1] * dxm("bz", k)
        for k in range(1, len(cs) + 1):
            ez = ez + cx[k - 1] * dxm("by", k)
        for k in range(1, len(cs) + 1):
            ez = ez - cy[k - 1] * dym("bx", k)
        ex, ey, ez = ex - fac * S("jx", 0, nx, 0, ny), ey - fac * S("jy", 0, nx, 0, ny), ez - fac * S("jz", 0, nx, 0, ny)
        S("ex", 0, nx, 0, ny)[...], S("ey", 0, nx, 0, ny)[...], S("ez", 0, nx, 0, ny)[...] = ex, ey, ez

    def update_b_field(self):
        nx, ny, S = self.nx, self.ny, self.S
        hdtx, hdty, _, _, _ = self.coeffs()
        cs = self.FD[int(getattr(self.dk, "field_order", 2))]
        cx, cy = [c * hdtx for c in cs], [c * hdty for c in cs]
        dxp = lambda name, k: S(name, k, nx + k, 0, ny) - S(name, 1 - k, nx + 1 - k, 0, ny)       # f(ix+k) - f(ix-k+1)
        dyp = lambda name, k: S(name, 0, nx, k, ny + k) - S(name, 0, nx, 1 - k, ny + 1 - k)
        bx, by, bz = S("bx", 0, nx, 0, ny).copy(), S("by", 0, nx, 0, ny).copy(), S("bz", 0, nx, 0, ny).copy()
        if getattr(self.dk, "maxwell_solver", "yee") != "yee":
            # fields.f90:441-465: alpha / beta / delta weighted differences (order 2 only)
            st = self.dk.stencil()
            ax, ay, bxy, byx, dlx, dly = st["alphax"], st["alphay"], st["betaxy"], st["betayx"], st["deltax"], st["deltay"]
            T = lambda name, di, dj: S(name, di, nx + di, dj, ny + dj)             # f(ix + di, iy + dj)

            def ddy(name):   # the y difference, weighted across x
                return (ay * (T(name, 0, 1) - T(name, 0, 0))
                        + byx * (T(name, 1, 1) - T(name, 1, 0) + T(name, -1, 1) - T(name, -1, 0))
                        + dly * (T(name, 0, 2) - T(name, 0, -1)))

            def ddx(name):   # the x difference, weighted across y
                return (ax * (T(name, 1, 0) - T(name, 0, 0))
                        + bxy * (T(name, 1, 1) - T(name, 0, 1) + T(name, 1, -1) - T(name, 0, -1))
                        + dlx * (T(name, 2, 0) - T(name, -1, 0)))

            bx = bx - hdty * ddy("ez")
            by = by + hdtx * ddx("ez")
            bz = bz - hdtx * ddx("ey") + hdty * ddy("ex")
            S("bx", 0, nx, 0, ny)[...], S("by", 0, nx, 0, ny)[...], S("bz", 0, nx, 0, ny)[...] = bx, by, bz
            return
        for k in range(1, len(cs) + 1):
            bx = bx - cy[k - 1] * dyp("ez", k)
            by = by + cx[k - 1] * dxp("ez", k)
            bz = bz - cx[k - 1] * dxp("ey", k)
        for k in range(1, len(cs) + 1):
            bz = bz + cy[k - 1] * dyp("ex", k)
        S("bx", 0, nx, 0, ny)[...], S("by", 0, nx, 0, ny)[...], S("bz", 0, nx, 0, ny)[...] = bx, by, bz

    def outflow_bcs(self, dt):
        nx, ny, S, c = self.nx, self.ny, self.S, D.c
        dtc2 = dt * (c * c)
        lx, ly = dtc2 / self.dk.dx(0), dtc2 / self.dk.dx(1)
        sum_, diff, dt_eps = 1.0 / (lx + c), lx - c, dt / D.epsilon0
        line = lambda name, i: S(name, i, i, 0, ny)[:, 0]
        zero = np.zeros(ny + 1)                 # the boundary snapshots of a run that starts from zero fields
        # x_min, laserpos = 1
        s1, s2 = self.src.get(0, (zero, zero))
        S("bx", 0, 0, 0, ny)[:, 0] = zero
        bz0 = sum_ * (4.0 * s1 + 2.0 * (zero + c * zero) - 2.0 * line("ey", 1) + dt_eps * line("jy", 1) + diff * line("bz", 1))
        by0 = sum_ * (-4.0 * s2 - 2.0 * (zero - c * zero) + 2.0 * line("ez", 1)
                      - ly * (line("bx", 1) - S("bx", 1, 1, -1, ny - 1)[:, 0]) - dt_eps * line("jz", 1) + diff * line("by", 1))
        S("bz", 0, 0, 0, ny)[:, 0], S("by", 0, 0, 0, ny)[:, 0] = bz0, by0
        # x_max, laserpos = nx
        s1, s2 = self.src.get(1, (zero, zero))
        S("bx", nx + 1, nx + 1, 0, ny)[:, 0] = zero
        bzn = sum_ * (-4.0 * s1 - 2.0 * (zero - c * zero) + 2.0 * line("ey", nx) - dt_eps * line("jy", nx) + diff * line("bz", nx - 1))
        byn = sum_ * (4.0 * s2 + 2.0 * (zero + c * zero) - 2.0 * line("ez", nx)
                      + ly * (line("bx", nx) - S("bx", nx, nx, -1, ny - 1)[:, 0]) + dt_eps * line("jz", nx) + diff * line("by", nx - 1))
        S("bz", nx, nx, 0, ny)[:, 0], S("by", nx, nx, 0, ny)[:, 0] = bzn, byn

    def outflow_bcs_y(self, dt):
        """outflow_bcs_y_min / y_max, laser.f90:462-610"""
        nx, ny, S, c = self.nx, self.ny, self.S, D.c
        dtc2 = dt * (c * c)
        lx, ly = dtc2 / self.dk.dx(0), dtc2 / self.dk.dx(1)
        sum_, diff, dt_eps = 1.0 / (ly + c), ly - c, dt / D.epsilon0
        row = lambda name, j, i0=0, i1=None: S(name, i0, nx if i1 is None else i1, j, j)[0, :]
        zero = np.zeros(nx + 1)
        s1, s2 = self.src.get(2, (zero, zero))
        S("by", 0, nx, 0, 0)[0, :] = zero
        bx0 = sum_ * (4.0 * s1 + 2.0 * (zero + c * zero) - 2.0 * row("ez", 1) - lx * (row("by", 1) - row("by", 1, -1, nx - 1))
                      + dt_eps * row("jz", 1) + diff * row("bx", 1))
        bz0 = sum_ * (-4.0 * s2 - 2.0 * (zero - c * zero) + 2.0 * row("ex", 1) - dt_eps * row("jx", 1) + diff * row("bz", 1))
        S("bx", 0, nx, 0, 0)[0, :], S("bz", 0, nx, 0, 0)[0, :] = bx0, bz0
        s1, s2 = self.src.get(3, (zero, zero))
        S("by", 0, nx, ny + 1, ny + 1)[0, :] = zero
        bxn = sum_ * (-4.0 * s1 - 2.0 * (zero - c * zero) + 2.0 * row("ez", ny) + lx * (row("by", ny) - row("by", ny, -1, nx - 1))
                      - dt_eps * row("jz", ny) + diff * row("bx", ny - 1))
        bzn = sum_ * (4.0 * s2 + 2.0 * (zero + c * zero) - 2.0 * row("ex", ny) + dt_eps * row("jx", ny) + diff * row("bz", ny - 1))
        S("bx", 0, nx, ny, ny)[0, :], S("bz", 0, nx, ny, ny)[0, :] = bxn, bzn

    def bfield_final_bcs(self, dt):
        self.bfield_bcs(False)
        if not self.periodic(0):
            self.outflow_bcs(dt)
        if not self.periodic(1):
            self.outflow_bcs_y(dt)
        self.bfield_bcs(True)

    # -- deck.run ---------------------------------------------------------------------------
    def init(self):
        self.efield_bcs()
        self.bfield_final_bcs(self.dt / 2.0)

    def fields_half(self):
        self.update_e_field()
        self.efield_bcs()
        self.update_b_field()
        self.bfield_bcs(True)

    def push(self):
        pass

    def current_finish(self):
        pass

    def fields_final(self):
        self.update_b_field()
        self.bfield_final_bcs(self.dt)
        self.update_e_field()
        self.efield_bcs()


import pytest


@pytest.mark.parametrize("order", [2, 4, 6])
def test_field_step_equals_an_independent_restatement(order):
    res = []
    for make in (Oracle, NumpyFields):
        dk = decks.laser2d(n=48)
        dk.field_order = order
        dk.lasers[0].pol_angle = 0.4            # both source terms
        b = make(dk)
        D.run(dk, b, [0], None, max_steps=60)
        res.append(b)
    o, m = res
    assert max(np.abs(o.field(0, k)).max() for k in ("ey", "ez", "by", "bz")) > 0
    for k in ("ex", "ey", "ez", "bx", "by", "bz"):
        assert np.array_equal(o.field(0, k)[0], m.f[k]), k


class NumpyFields1D:
    """the same for epoch1d: fields.f90:150-166 and the B update of :228-237 without the kappas, laser.f90:260-392"""

    def __init__(self, dk):
        self.dk, self.nx, self.dt = dk, dk.n[0], dk.dt()
        self.f = {k: np.zeros(self.nx + 2 * NG) for k in ("ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz")}
        self.src = {}

    def S(self, name, i0, i1):
        return self.f[name][i0 + NG - 1:i1 + NG]

    def set_laser_source(self, lr, side, s1, s2):
        self.src[side] = (float(np.asarray(s1).ravel()[0]), float(np.asarray(s2).ravel()[0]))

    def clamp_zero(self, name, side):
        a, nn = self.f[name], self.nx
        F = lambda i: i + NG - 1
        if side == 0:
            if STAG[name][0]:
                for i in range(1, NG):
                    a[F(i - NG)] = -a[F(NG - i)]
                a[F(0)] = 0.0
            else:
                for i in range(1, NG + 1):
                    a[F(i - NG)] = -a[F(NG + 1 - i)]
        else:
            if STAG[name][0]:
                a[F(nn)] = 0.0
                for i in range(1, NG):
                    a[F(nn + i)] = -a[F(nn - i)]
            else:
                for i in range(1, NG + 1):
                    a[F(nn + i)] = -a[F(nn + 1 - i)]

    def efield_bcs(self):
        for side in (0, 1):
            for k in ("ex", "ey", "ez"):
                self.clamp_zero(k, side)

    def bfield_bcs(self, mpi_only):
        if mpi_only:
            return
        for side in (0, 1):
            for k in ("bx", "by", "bz"):
                self.clamp_zero(k, side)

    def update_e_field(self):
        nx, S = self.nx, self.S
        hdt = 0.5 * self.dt
        cnx, fac = (hdt / self.dk.dx(0)) * (D.c * D.c), hdt / D.epsilon0
        ex = S("ex", 0, nx) - fac * S("jx", 0, nx)
        ey = S("ey", 0, nx) - cnx * (S("bz", 0, nx) - S("bz", -1, nx - 1)) - fac * S("jy", 0, nx)
        ez = S("ez", 0, nx) + cnx * (S("by", 0, nx) - S("by", -1, nx - 1)) - fac * S("jz", 0, nx)
        S("ex", 0, nx)[...], S("ey", 0, nx)[...], S("ez", 0, nx)[...] = ex, ey, ez

    def update_b_field(self):
        nx, S = self.nx, self.S
        hdtx = 0.5 * self.dt / self.dk.dx(0)
        by = S("by", 0, nx) + hdtx * (S("ez", 1, nx + 1) - S("ez", 0, nx))
        bz = S("bz", 0, nx) - hdtx * (S("ey", 1, nx + 1) - S("ey", 0, nx))
        S("by", 0, nx)[...], S("bz", 0, nx)[...] = by, bz

    def bfield_final_bcs(self, dt):
        self.bfield_bcs(False)
        c, f, F, nx = D.c, self.f, (lambda i: i + NG - 1), self.nx
        dtc2 = dt * (c * c)
        lx = dtc2 / self.dk.dx(0)
        sum_, diff, dt_eps = 1.0 / (lx + c), lx - c, dt / D.epsilon0
        s1, s2 = self.src.get(0, (0.0, 0.0))
        f["bx"][F(0)] = 0.0
        bz0 = sum_ * (4.0 * s1 + 2.0 * (0.0 + c * 0.0) - 2.0 * f["ey"][F(1)] + dt_eps * f["jy"][F(1)] + diff * f["bz"][F(1)])
        by0 = sum_ * (-4.0 * s2 - 2.0 * (0.0 - c * 0.0) + 2.0 * f["ez"][F(1)] - dt_eps * f["jz"][F(1)] + diff * f["by"][F(1)])
        f["bz"][F(0)], f["by"][F(0)] = bz0, by0
        s1, s2 = self.src.get(1, (0.0, 0.0))
        f["bx"][F(nx + 1)] = 0.0
        bzn = sum_ * (-4.0 * s1 - 2.0 * (0.0 - c * 0.0) + 2.0 * f["ey"][F(nx)] - dt_eps * f["jy"][F(nx)] + diff * f["bz"][F(nx - 1)])
        byn = sum_ * (4.0 * s2 + 2.0 * (0.0 + c * 0.0) - 2.0 * f["ez"][F(nx)] + dt_eps * f["jz"][F(nx)] + diff * f["by"][F(nx - 1)])
        f["bz"][F(nx)], f["by"][F(nx)] = bzn, byn
        self.bfield_bcs(True)

    def init(self):
        self.efield_bcs()
        self.bfield_final_bcs(self.dt / 2.0)

    def fields_half(self):
        self.update_e_field(); self.efield_bcs(); self.update_b_field(); self.bfield_bcs(True)

    def push(self):
        pass

    def current_finish(self):
        pass

    def fields_final(self):
        self.update_b_field(); self.bfield_final_bcs(self.dt); self.update_e_field(); self.efield_bcs()


def test_field_step_1d_equals_an_independent_restatement():
    """epoch1d/tests/laser/input.deck itself (the reference's first golden-sum deck), 200 steps"""
    res = []
    for make in (Oracle, NumpyFields1D):
        dk = decks.laser1d()
        dk.lasers[0].pol_angle = 0.3
        b = make(dk)
        D.run(dk, b, [0], None, max_steps=200)
        res.append(b)
    o, m = res
    assert max(np.abs(o.field(0, k)).max() for k in ("ey", "ez", "by", "bz")) > 0
    for k in ("ex", "ey", "ez", "bx", "by", "bz"):
        assert np.array_equal(o.field(0, k)[0, 0], m.f[k]), k


class NumpyFields3D:
    """the same for epoch3d: fields.f90:312-337, :637-655 (order 2, Yee), laser.f90:350-506, boundary.F90 of that tree"""
    STAG3 = {"ex": (1, 0, 0), "ey": (0, 1, 0), "ez": (0, 0, 1), "bx": (0, 1, 1), "by": (1, 0, 1), "bz": (1, 1, 0)}

    def __init__(self, dk):
        self.dk, self.dt = dk, dk.dt()
        self.nx, self.ny, self.nz = dk.n
        shape = (self.nz + 2 * NG, self.ny + 2 * NG, self.nx + 2 * NG)
        self.f = {k: np.zeros(shape) for k in ("ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz")}
        self.src = {}

    def S(self, name, ix, iy, iz):
        """Fortran section name(ix0:ix1, iy0:iy1, iz0:iz1)"""
        sl = lambda r: slice(r[0] + NG - 1, r[1] + NG)
        return self.f[name][sl(iz), sl(iy), sl(ix)]

    def set_laser_source(self, lr, side, s1, s2):
        shp = (self.nz + 1, self.ny + 1)     # (0:ny, 0:nz), y fastest
        self.src[side] = (np.asarray(s1, dtype=np.float64).reshape(shp).copy(), np.asarray(s2, dtype=np.float64).reshape(shp).copy())

    def field_bc(self, name):
        a, ny, nz = self.f[name], self.ny, self.nz      # x not periodic; y then z: copy from the other end
        a[:, ny + NG:ny + 2 * NG, :] = a[:, NG:2 * NG, :]
        a[:, 0:NG, :] = a[:, ny:ny + NG, :]
        a[nz + NG:nz + 2 * NG, :, :] = a[NG:2 * NG, :, :]
        a[0:NG, :, :] = a[nz:nz + NG, :, :]

    def clamp_zero_x(self, name, side):
        a, nn = self.f[name], self.nx
        F = lambda i: i + NG - 1
        if side == 0:
            if self.STAG3[name][0]:
                for i in range(1, NG):
                    a[:, :, F(i - NG)] = -a[:, :, F(NG - i)]
                a[:, :, F(0)] = 0.0
            else:
                for i in range(1, NG + 1):
                    a[:, :, F(i - NG)] = -a[:, :, F(NG + 1 - i)]
        else:
            if self.STAG3[name][0]:
                a[:, :, F(nn)] = 0.0
                for i in range(1, NG):
                    a[:, :, F(nn + i)] = -a[:, :, F(nn - i)]
            else:
                for i in range(1, NG + 1):
                    a[:, :, F(nn + i)] = -a[:, :, F(nn + 1 - i)]

    def efield_bcs(self):
        for k in ("ex", "ey", "ez"):
            self.field_bc(k)
        for side in (0, 1):
            for k in ("ex", "ey", "ez"):
                self.clamp_zero_x(k, side)

    def bfield_bcs(self, mpi_only):
        for k in ("bx", "by", "bz"):
            self.field_bc(k)
        if mpi_only:
            return
        for side in (0, 1):
            for k in ("bx", "by", "bz"):
                self.clamp_zero_x(k, side)

    def update_e_field(self):
        S, X, Y, Z = self.S, (0, self.nx), (0, self.ny), (0, self.nz)
        m = lambda r: (r[0] - 1, r[1] - 1)
        hdt = 0.5 * self.dt
        cc = D.c * D.c
        cnx, cny, cnz = (hdt / self.dk.dx(0)) * cc, (hdt / self.dk.dx(1)) * cc, (hdt / self.dk.dx(2)) * cc
        fac = hdt / D.epsilon0
        ex = S("ex", X, Y, Z) + cny * (S("bz", X, Y, Z) - S("bz", X, m(Y), Z)) - cnz * (S("by", X, Y, Z) - S("by", X, Y, m(Z))) \
            - fac * S("jx", X, Y, Z)
        ey = S("ey", X, Y, Z) + cnz * (S("bx", X, Y, Z) - S("bx", X, Y, m(Z))) - cnx * (S("bz", X, Y, Z) - S("bz", m(X), Y, Z)) \
            - fac * S("jy", X, Y, Z)
        ez = S("ez", X, Y, Z) + cnx * (S("by", X, Y, Z) - S("by", m(X), Y, Z)) - cny * (S("bx", X, Y, Z) - S("bx", X, m(Y), Z)) \
            - fac * S("jz", X, Y, Z)
        S("ex", X, Y, Z)[...], S("ey", X, Y, Z)[...], S("ez", X, Y, Z)[...] = ex, ey, ez

    def update_b_field(self):
        S, X, Y, Z = self.S, (0, self.nx), (0, self.ny), (0, self.nz)
        p = lambda r: (r[0] + 1, r[1] + 1)
        hdt = 0.5 * self.dt
        hx, hy, hz = hdt / self.dk.dx(0), hdt / self.dk.dx(1), hdt / self.dk.dx(2)
        bx = S("bx", X, Y, Z) - hy * (S("ez", X, p(Y), Z) - S("ez", X, Y, Z)) + hz * (S("ey", X, Y, p(Z)) - S("ey", X, Y, Z))
        by = S("by", X, Y, Z) - hz * (S("ex", X, Y, p(Z)) - S("ex", X, Y, Z)) + hx * (S("ez", p(X), Y, Z) - S("ez", X, Y, Z))
        bz = S("bz", X, Y, Z) - hx * (S("ey", p(X), Y, Z) - S("ey", X, Y, Z)) + hy * (S("ex", X, p(Y), Z) - S("ex", X, Y, Z))
        S("bx", X, Y, Z)[...], S("by", X, Y, Z)[...], S("bz", X, Y, Z)[...] = bx, by, bz

    def bfield_final_bcs(self, dt):
        self.bfield_bcs(False)
        S, c, nx, Y, Z = self.S, D.c, self.nx, (0, self.ny), (0, self.nz)
        m = lambda r: (r[0] - 1, r[1] - 1)
        dtc2 = dt * (c * c)
        lx, ly, lz = dtc2 / self.dk.dx(0), dtc2 / self.dk.dx(1), dtc2 / self.dk.dx(2)
        sum_, diff, dt_eps = 1.0 / (lx + c), lx - c, dt / D.epsilon0
        P = lambda name, i, yy=Y, zz=Z: S(name, (i, i), yy, zz)[:, :, 0]
        zero = np.zeros((self.nz + 1, self.ny + 1))
        s1, s2 = self.src.get(0, (zero, zero))
        P("bx", 0)[...] = zero
        bz0 = sum_ * (4.0 * s1 + 2.0 * (zero + c * zero) - 2.0 * P("ey", 1) - lz * (P("bx", 1) - P("bx", 1, Y, m(Z)))
                      + dt_eps * P("jy", 1) + diff * P("bz", 1))
        by0 = sum_ * (-4.0 * s2 - 2.0 * (zero - c * zero) + 2.0 * P("ez", 1) - ly * (P("bx", 1) - P("bx", 1, m(Y), Z))
                      - dt_eps * P("jz", 1) + diff * P("by", 1))
        P("bz", 0)[...], P("by", 0)[...] = bz0, by0
        s1, s2 = self.src.get(1, (zero, zero))
        P("bx", nx + 1)[...] = zero
        bzn = sum_ * (-4.0 * s1 - 2.0 * (zero - c * zero) + 2.0 * P("ey", nx) + lz * (P("bx", nx) - P("bx", nx, Y, m(Z)))
                      - dt_eps * P("jy", nx) + diff * P("bz", nx - 1))
        byn = sum_ * (4.0 * s2 + 2.0 * (zero + c * zero) - 2.0 * P("ez", nx) + ly * (P("bx", nx) - P("bx", nx, m(Y), Z))
                      + dt_eps * P("jz", nx) + diff * P("by", nx - 1))
        P("bz", nx)[...], P("by", nx)[...] = bzn, byn
        self.bfield_bcs(True)

    def init(self):
        self.efield_bcs()
        self.bfield_final_bcs(self.dt / 2.0)

    def fields_half(self):
        self.update_e_field(); self.efield_bcs(); self.update_b_field(); self.bfield_bcs(True)

    def push(self):
        pass

    def current_finish(self):
        pass

    def fields_final(self):
        self.update_b_field(); self.bfield_final_bcs(self.dt); self.update_e_field(); self.efield_bcs()


def test_field_step_3d_equals_an_independent_restatement():
    """epoch3d/tests/laser/input.deck at 28^3 cells, 40 steps, with a second polarisation component"""
    res = []
    for make in (Oracle, NumpyFields3D):
        dk = decks.laser3d(n=28)
        dk.lasers[0].pol_angle = 0.5
        b = make(dk)
        D.run(dk, b, [0], None, max_steps=40)
        res.append(b)
    o, m = res
    assert min(np.abs(o.field(0, k)).max() for k in ("ex", "ey", "ez", "bx", "by", "bz")) > 0
    for k in ("ex", "ey", "ez", "bx", "by", "bz"):
        assert np.array_equal(o.field(0, k), m.f[k]), k


def test_field_step_y_face_laser_equals_an_independent_restatement():
    """laser on y_min, outflow on y_max, periodic x (outflow_bcs_y_min / y_max, clamp on the y walls)"""
    res = []
    for make in (Oracle, NumpyFields):
        dk = decks.laser2d_y(n=40)
        dk.lasers[0].pol_angle = 0.7
        b = make(dk)
        D.run(dk, b, [0], None, max_steps=60)
        res.append(b)
    o, m = res
    assert min(np.abs(o.field(0, k)).max() for k in ("ex", "ez", "bx", "bz")) > 0
    for k in ("ex", "ey", "ez", "bx", "by", "bz"):
        assert np.array_equal(o.field(0, k)[0], m.f[k]), k


@pytest.mark.parametrize("solver", ["lehe_x", "lehe_y", "pukhov", "custom"])
def test_extended_stencils_equal_an_independent_restatement(solver):
    """the non-Yee B update (fields.f90:441-465) with the coefficients set_maxwell_solver derives"""
    res = []
    for make in (Oracle, NumpyFields):
        dk = decks.laser2d(n=40)
        dk.maxwell_solver = solver
        if solver == "custom":
            dk.stencil_custom = dict(betaxy=0.11, betayx=0.07, deltax=-0.03, deltay=0.02, dt=0.9 * dk.dx(0) / D.c / 2 ** 0.5)
        dk.lasers[0].pol_angle = 0.4
        b = make(dk)
        D.run(dk, b, [0], None, max_steps=50)
        res.append(b)
    o, m = res
    assert min(np.abs(o.field(0, k)).max() for k in ("ey", "ez", "by", "bz")) > 0
    for k in ("ex", "ey", "ez", "bx", "by", "bz"):
        assert np.array_equal(o.field(0, k)[0], m.f[k]), (solver, k)
