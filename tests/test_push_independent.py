"""The particle half of the oracle has no reference numbers to be pinned on ("parity unpinned"), so it
is held to a SECOND restatement instead: push_particles of epoch2d (src/particles.F90:138-582 with
include/triangle/{gx,hx_dcell,e_part,b_part}.inc), written here in scalar Python straight from the Fortran --
same expression trees, IEEE doubles, no fused multiply-add, x**2 written x * x as gfortran compiles it (libm's pow
is not correctly rounded: one particle in a thousand came out one ulp off with it) -- and compared with the C++
oracle bit for bit:
positions, momenta and the three current arrays (ghost cells included, before current_finish) after one push in
random fields.  A transliteration slip would have to be made twice, in two languages, to go unnoticed."""
import math

import numpy as np

from epoch_b200 import deck as D
from oracle.oracle import Oracle
from tests import decks

NG = 5


def push_particles_2d(dk, fields, parts, charge, mass, gmin_local, hc=None):
    """One call of push_particles for one species on one rank.  fields: dict name -> array [y][x] with ghosts;
    parts: (n, 6) x y px py pz w, updated in place; returns jx, jy, jz."""
    c = D.c
    dx, dy, dt = dk.dx(0), dk.dx(1), dk.dt()
    fac = 0.5 ** 2
    idx, idy, idt = 1.0 / dx, 1.0 / dy, 1.0 / dt
    dto2 = dt / 2.0
    dtco2 = c * dto2
    dtfac = 0.5 * dt * fac
    third = 1.0 / 3.0
    idty, idtx, idxy = idt * idy * fac, idt * idx * fac, idx * idy * fac
    shape = fields["ex"].shape
    jx, jy, jz = np.zeros(shape), np.zeros(shape), np.zeros(shape)

    def at(a, i, j):              # Fortran a(i, j), lower bounds 1 - ng
        return float(a[j + NG - 1, i + NG - 1])

    part_q, part_mc = charge, c * mass
    ipart_mc = 1.0 / part_mc
    cmratio = part_q * dtfac * ipart_mc
    ccmratio = c * cmratio
    ex, ey, ez, bx, by, bz = (fields[k] for k in ("ex", "ey", "ez", "bx", "by", "bz"))
    for P in parts:
        part_weight = float(P[5])
        fcx, fcy, fcz = idty * part_weight, idtx * part_weight, idxy * part_weight
        part_x, part_y = float(P[0]) - gmin_local[0], float(P[1]) - gmin_local[1]
        part_ux, part_uy, part_uz = float(P[2]) * ipart_mc, float(P[3]) * ipart_mc, float(P[4]) * ipart_mc
        gamma_rel = math.sqrt(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0)
        root = dtco2 / gamma_rel
        part_x = part_x + part_ux * root
        part_y = part_y + part_uy * root
        cell_x_r, cell_y_r = part_x * idx, part_y * idy
        cell_x1 = math.floor(cell_x_r + 0.5)
        cell_frac_x = float(cell_x1) - cell_x_r
        cell_x1 = cell_x1 + 1
        cell_y1 = math.floor(cell_y_r + 0.5)
        cell_frac_y = float(cell_y1) - cell_y_r
        cell_y1 = cell_y1 + 1
        # gx.inc; gx, gy are (sf_min-1 : sf_max+1) = (-2 : 2), zero outside -1 .. 1
        gx, gy = {k: 0.0 for k in range(-2, 3)}, {k: 0.0 for k in range(-2, 3)}
        cf2 = cell_frac_x * cell_frac_x
        gx[-1], gx[0], gx[1] = 0.25 + cf2 + cell_frac_x, 1.5 - 2.0 * cf2, 0.25 + cf2 - cell_frac_x
        cf2 = cell_frac_y * cell_frac_y
        gy[-1], gy[0], gy[1] = 0.25 + cf2 + cell_frac_y, 1.5 - 2.0 * cf2, 0.25 + cf2 - cell_frac_y
        cell_x2 = math.floor(cell_x_r)
        cell_frac_x = float(cell_x2) - cell_x_r + 0.5
        cell_x2 = cell_x2 + 1
        cell_y2 = math.floor(cell_y_r)
        cell_frac_y = float(cell_y2) - cell_y_r + 0.5
        cell_y2 = cell_y2 + 1
        # hx_dcell.inc with dcellx = dcelly = 0
        hx, hy = {k: 0.0 for k in range(-2, 3)}, {k: 0.0 for k in range(-2, 3)}
        cf2 = cell_frac_x * cell_frac_x
        hx[-1], hx[0], hx[1] = 0.25 + cf2 + cell_frac_x, 1.5 - 2.0 * cf2, 0.25 + cf2 - cell_frac_x
        cf2 = cell_frac_y * cell_frac_y
        hy[-1], hy[0], hy[1] = 0.25 + cf2 + cell_frac_y, 1.5 - 2.0 * cf2, 0.25 + cf2 - cell_frac_y

        def gather(a, wx_, cx, wy_, cy):
            # e_part.inc / b_part.inc: rows in y, each row summed left to right inside its parentheses
            tot = None
            for j in (-1, 0, 1):
                row = wx_[-1] * at(a, cx - 1, cy + j) + wx_[0] * at(a, cx, cy + j) + wx_[1] * at(a, cx + 1, cy + j)
                tot = wy_[j] * row if tot is None else tot + wy_[j] * row
            return tot

        ex_part = gather(ex, hx, cell_x2, gy, cell_y1)
        ey_part = gather(ey, gx, cell_x1, hy, cell_y2)
        ez_part = gather(ez, gx, cell_x1, gy, cell_y1)
        bx_part = gather(bx, gx, cell_x1, hy, cell_y2)
        by_part = gather(by, hx, cell_x2, gy, cell_y1)
        bz_part = gather(bz, hx, cell_x2, hy, cell_y2)
        part_ux, part_uy, part_uz = _boris((part_ux, part_uy, part_uz), (ex_part, ey_part, ez_part),
                                           (bx_part, by_part, bz_part), cmratio, ccmratio, hc)
        part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz
        gamma_rel = math.sqrt(part_u2 + 1.0)
        igamma = 1.0 / gamma_rel
        root = dtco2 * igamma
        delta_x, delta_y = part_ux * root, part_uy * root
        part_vz = part_uz * c * igamma
        part_x = part_x + delta_x
        part_y = part_y + delta_y
        P[0], P[1] = part_x + gmin_local[0], part_y + gmin_local[1]
        P[2], P[3], P[4] = part_mc * part_ux, part_mc * part_uy, part_mc * part_uz
        # current: advance to t + 3 dt / 2
        part_x = part_x + delta_x
        part_y = part_y + delta_y
        cell_x_r, cell_y_r = part_x * idx, part_y * idy
        cell_x3 = math.floor(cell_x_r + 0.5)
        cell_frac_x = float(cell_x3) - cell_x_r
        cell_x3 = cell_x3 + 1
        cell_y3 = math.floor(cell_y_r + 0.5)
        cell_frac_y = float(cell_y3) - cell_y_r
        cell_y3 = cell_y3 + 1
        hx, hy = {k: 0.0 for k in range(-2, 3)}, {k: 0.0 for k in range(-2, 3)}
        dcellx, dcelly = cell_x3 - cell_x1, cell_y3 - cell_y1
        cf2 = cell_frac_x * cell_frac_x
        hx[dcellx - 1], hx[dcellx], hx[dcellx + 1] = 0.25 + cf2 + cell_frac_x, 1.5 - 2.0 * cf2, 0.25 + cf2 - cell_frac_x
        cf2 = cell_frac_y * cell_frac_y
        hy[dcelly - 1], hy[dcelly], hy[dcelly + 1] = 0.25 + cf2 + cell_frac_y, 1.5 - 2.0 * cf2, 0.25 + cf2 - cell_frac_y
        for k in range(-2, 3):
            hx[k] = hx[k] - gx[k]
            hy[k] = hy[k] - gy[k]
        tz = lambda a, b: int(a / b)      # Fortran integer division truncates towards zero
        xmin, xmax = -1 + tz(dcellx - 1, 2), 1 + tz(dcellx + 1, 2)
        ymin, ymax = -1 + tz(dcelly - 1, 2), 1 + tz(dcelly + 1, 2)
        fjx, fjy, fjz = fcx * part_q, fcy * part_q, fcz * part_q * part_vz
        jyh = {k: 0.0 for k in range(-2, 3)}
        for iy in range(ymin, ymax + 1):
            cy = cell_y1 + iy
            yfac1 = gy[iy] + 0.5 * hy[iy]
            yfac2 = third * hy[iy] + 0.5 * gy[iy]
            hy_iy = hy[iy]
            jxh = 0.0
            for ix in range(xmin, xmax + 1):
                cx = cell_x1 + ix
                xfac1 = gx[ix] + 0.5 * hx[ix]
                wx = hx[ix] * yfac1
                wy = hy_iy * xfac1
                wz = gx[ix] * yfac1 + hx[ix] * yfac2
                jxh = jxh - fjx * wx
                jyh[ix] = jyh[ix] - fjy * wy
                jzh = fjz * wz
                jx[cy + NG - 1, cx + NG - 1] += jxh
                jy[cy + NG - 1, cx + NG - 1] += jyh[ix]
                jz[cy + NG - 1, cx + NG - 1] += jzh
    return jx, jy, jz


def _tri(cf, shift=0):
    """gx.inc / hx_dcell.inc: the three un-normalised triangle weights at dcell-1 .. dcell+1 of a (-2:2) array"""
    w = {k: 0.0 for k in range(-2, 3)}
    cf2 = cf * cf
    w[shift - 1], w[shift], w[shift + 1] = 0.25 + cf2 + cf, 1.5 - 2.0 * cf2, 0.25 + cf2 - cf
    return w


def _boris(part_u, e_part, b_part, cmratio, ccmratio, hc=None):
    """particles.F90:382-428, identical text in the three trees; hc = (part_q, dt, part_m): the -DHC_PUSH gamma
    (Higuera-Cary, :386-398)"""
    part_ux, part_uy, part_uz = part_u
    ex_part, ey_part, ez_part = e_part
    bx_part, by_part, bz_part = b_part
    uxm = part_ux + cmratio * ex_part
    uym = part_uy + cmratio * ey_part
    uzm = part_uz + cmratio * ez_part
    if hc is None:
        gamma_rel = math.sqrt(uxm * uxm + uym * uym + uzm * uzm + 1.0)
    else:
        part_q, dt, part_m = hc
        gamma_rel = uxm * uxm + uym * uym + uzm * uzm + 1.0
        alpha = 0.5 * part_q * dt / part_m
        beta_x, beta_y, beta_z = alpha * bx_part, alpha * by_part, alpha * bz_part
        beta2 = beta_x * beta_x + beta_y * beta_y + beta_z * beta_z
        sigma = gamma_rel - beta2
        beta_dot_u = beta_x * uxm + beta_y * uym + beta_z * uzm
        gamma_rel = sigma + math.sqrt(sigma * sigma + 4.0 * (beta2 + beta_dot_u * beta_dot_u))
        gamma_rel = math.sqrt(0.5 * gamma_rel)
    root = ccmratio / gamma_rel
    taux, tauy, tauz = bx_part * root, by_part * root, bz_part * root
    taux2, tauy2, tauz2 = taux * taux, tauy * tauy, tauz * tauz
    tau = 1.0 / (1.0 + taux2 + tauy2 + tauz2)
    uxp = ((1.0 + taux2 - tauy2 - tauz2) * uxm
           + 2.0 * ((taux * tauy + tauz) * uym
           + (taux * tauz - tauy) * uzm)) * tau
    uyp = ((1.0 - taux2 + tauy2 - tauz2) * uym
           + 2.0 * ((tauy * tauz + taux) * uzm
           + (tauy * taux - tauz) * uxm)) * tau
    uzp = ((1.0 - taux2 - tauy2 + tauz2) * uzm
           + 2.0 * ((tauz * taux + tauy) * uxm
           + (tauz * tauy - taux) * uym)) * tau
    return uxp + cmratio * ex_part, uyp + cmratio * ey_part, uzp + cmratio * ez_part


def _tz(a, b):
    return int(a / b)      # Fortran integer division truncates towards zero


def push_particles_1d(dk, fields, parts, charge, mass, gmin_local, hc=None):
    """epoch1d/src/particles.F90:143-507 (+ include/triangle/*.inc of that tree).  parts: (n, 5) x px py pz w."""
    c = D.c
    dx, dt = dk.dx(0), dk.dt()
    fac = 0.5 ** 1
    idx, idt = 1.0 / dx, 1.0 / dt
    dto2 = dt / 2.0
    dtco2 = c * dto2
    dtfac = 0.5 * dt * fac
    idtf, idxf = idt * fac, idx * fac
    n = fields["ex"].shape
    jx, jy, jz = np.zeros(n), np.zeros(n), np.zeros(n)
    at = lambda a, i: float(a[i + NG - 1])
    part_q, part_mc = charge, c * mass
    ipart_mc = 1.0 / part_mc
    cmratio = part_q * dtfac * ipart_mc
    ccmratio = c * cmratio
    g3 = lambda a, w, cx: w[-1] * at(a, cx - 1) + w[0] * at(a, cx) + w[1] * at(a, cx + 1)
    for P in parts:
        part_weight = float(P[4])
        fcx, fcy = idtf * part_weight, idxf * part_weight
        part_x = float(P[0]) - gmin_local[0]
        part_ux, part_uy, part_uz = float(P[1]) * ipart_mc, float(P[2]) * ipart_mc, float(P[3]) * ipart_mc
        gamma_rel = math.sqrt(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0)
        root = dtco2 / gamma_rel
        part_x = part_x + part_ux * root
        cell_x_r = part_x * idx
        cell_x1 = math.floor(cell_x_r + 0.5)
        gx = _tri(float(cell_x1) - cell_x_r)
        cell_x1 = cell_x1 + 1
        cell_x2 = math.floor(cell_x_r)
        hx = _tri(float(cell_x2) - cell_x_r + 0.5)
        cell_x2 = cell_x2 + 1
        e_part = (g3(fields["ex"], hx, cell_x2), g3(fields["ey"], gx, cell_x1), g3(fields["ez"], gx, cell_x1))
        b_part = (g3(fields["bx"], gx, cell_x1), g3(fields["by"], hx, cell_x2), g3(fields["bz"], hx, cell_x2))
        part_ux, part_uy, part_uz = _boris((part_ux, part_uy, part_uz), e_part, b_part, cmratio, ccmratio, hc)
        part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz
        gamma_rel = math.sqrt(part_u2 + 1.0)
        root = c / gamma_rel
        delta_x = part_ux * root * dto2
        part_vy = part_uy * root
        part_vz = part_uz * root
        part_x = part_x + delta_x
        P[0] = part_x + gmin_local[0]
        P[1], P[2], P[3] = part_mc * part_ux, part_mc * part_uy, part_mc * part_uz
        part_x = part_x + delta_x
        cell_x_r = part_x * idx
        cell_x3 = math.floor(cell_x_r + 0.5)
        cell_frac_x = float(cell_x3) - cell_x_r
        cell_x3 = cell_x3 + 1
        dcellx = cell_x3 - cell_x1
        hx = _tri(cell_frac_x, dcellx)
        for k in range(-2, 3):
            hx[k] = hx[k] - gx[k]
        xmin, xmax = -1 + _tz(dcellx - 1, 2), 1 + _tz(dcellx + 1, 2)
        fjx = fcx * part_q
        fjy = fcy * part_q * part_vy
        fjz = fcy * part_q * part_vz
        jxh = 0.0
        for ix in range(xmin, xmax + 1):
            cx = cell_x1 + ix
            wx = hx[ix]
            wy = gx[ix] + 0.5 * hx[ix]
            jxh = jxh - fjx * wx
            jyh = fjy * wy
            jzh = fjz * wy
            jx[cx + NG - 1] += jxh
            jy[cx + NG - 1] += jyh
            jz[cx + NG - 1] += jzh
    return jx, jy, jz


def push_particles_3d(dk, fields, parts, charge, mass, gmin_local, hc=None):
    """epoch3d/src/particles.F90:150-650 (+ include/triangle/*.inc of that tree).  parts: (n, 7) x y z px py pz w."""
    c = D.c
    dx, dy, dz, dt = dk.dx(0), dk.dx(1), dk.dx(2), dk.dt()
    fac = 0.5 ** 3
    idx, idy, idz, idt = 1.0 / dx, 1.0 / dy, 1.0 / dz, 1.0 / dt
    dto2 = dt / 2.0
    dtco2 = c * dto2
    dtfac = 0.5 * dt * fac
    third = 1.0 / 3.0
    idtyz, idtxz, idtxy = idt * idy * idz * fac, idt * idx * idz * fac, idt * idx * idy * fac
    shape = fields["ex"].shape
    jx, jy, jz = np.zeros(shape), np.zeros(shape), np.zeros(shape)
    at = lambda a, i, j, k: float(a[k + NG - 1, j + NG - 1, i + NG - 1])
    part_q, part_mc = charge, c * mass
    ipart_mc = 1.0 / part_mc
    cmratio = part_q * dtfac * ipart_mc
    ccmratio = c * cmratio

    def gather(a, wx_, cx, wy_, cy, wz_, cz):
        # e_part.inc / b_part.inc: wz * (wy * (row) + wy * (row) + wy * (row)), planes then rows, left to right
        tot = None
        for k in (-1, 0, 1):
            plane = None
            for j in (-1, 0, 1):
                row = (wx_[-1] * at(a, cx - 1, cy + j, cz + k) + wx_[0] * at(a, cx, cy + j, cz + k)
                       + wx_[1] * at(a, cx + 1, cy + j, cz + k))
                plane = wy_[j] * row if plane is None else plane + wy_[j] * row
            tot = wz_[k] * plane if tot is None else tot + wz_[k] * plane
        return tot

    for P in parts:
        part_weight = float(P[6])
        fcx, fcy, fcz = idtyz * part_weight, idtxz * part_weight, idtxy * part_weight
        pos = [float(P[d]) - gmin_local[d] for d in range(3)]
        u = [float(P[3 + d]) * ipart_mc for d in range(3)]
        gamma_rel = math.sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + 1.0)
        root = dtco2 / gamma_rel
        pos = [pos[d] + u[d] * root for d in range(3)]
        cell_r = [pos[0] * idx, pos[1] * idy, pos[2] * idz]
        cell1, g, cell2, h = [], [], [], []
        for d in range(3):
            c1 = math.floor(cell_r[d] + 0.5)
            g.append(_tri(float(c1) - cell_r[d]))
            cell1.append(c1 + 1)
        for d in range(3):
            c2 = math.floor(cell_r[d])
            h.append(_tri(float(c2) - cell_r[d] + 0.5))
            cell2.append(c2 + 1)
        (gx, gy, gz), (hx, hy, hz) = g, h
        (x1, y1, z1), (x2, y2, z2) = cell1, cell2
        e_part = (gather(fields["ex"], hx, x2, gy, y1, gz, z1), gather(fields["ey"], gx, x1, hy, y2, gz, z1),
                  gather(fields["ez"], gx, x1, gy, y1, hz, z2))
        b_part = (gather(fields["bx"], gx, x1, hy, y2, hz, z2), gather(fields["by"], hx, x2, gy, y1, hz, z2),
                  gather(fields["bz"], hx, x2, hy, y2, gz, z1))
        part_ux, part_uy, part_uz = _boris(tuple(u), e_part, b_part, cmratio, ccmratio, hc)
        part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz
        gamma_rel = math.sqrt(part_u2 + 1.0)
        root = dtco2 / gamma_rel
        delta = [part_ux * root, part_uy * root, part_uz * root]
        pos = [pos[d] + delta[d] for d in range(3)]
        for d in range(3):
            P[d] = pos[d] + gmin_local[d]
        P[3], P[4], P[5] = part_mc * part_ux, part_mc * part_uy, part_mc * part_uz
        pos = [pos[d] + delta[d] for d in range(3)]
        cell_r = [pos[0] * idx, pos[1] * idy, pos[2] * idz]
        hh, dcell = [], []
        for d in range(3):
            c3 = math.floor(cell_r[d] + 0.5)
            cf = float(c3) - cell_r[d]
            c3 = c3 + 1
            dcell.append(c3 - cell1[d])
            w = _tri(cf, dcell[d])
            for k in range(-2, 3):
                w[k] = w[k] - g[d][k]
            hh.append(w)
        hx, hy, hz = hh
        lo = [-1 + _tz(dcell[d] - 1, 2) for d in range(3)]
        hi = [1 + _tz(dcell[d] + 1, 2) for d in range(3)]
        fjx, fjy, fjz = fcx * part_q, fcy * part_q, fcz * part_q
        jzh = {(i, j): 0.0 for i in range(-2, 3) for j in range(-2, 3)}
        for iz in range(lo[2], hi[2] + 1):
            cz = z1 + iz
            zfac1 = gz[iz] + 0.5 * hz[iz]
            zfac2 = third * hz[iz] + 0.5 * gz[iz]
            gz_iz, hz_iz = gz[iz], hz[iz]
            jyh = {i: 0.0 for i in range(-2, 3)}
            for iy in range(lo[1], hi[1] + 1):
                cy = y1 + iy
                yfac1 = gy[iy] + 0.5 * hy[iy]
                yfac2 = third * hy[iy] + 0.5 * gy[iy]
                hygz = hy[iy] * gz_iz
                hyhz = hy[iy] * hz_iz
                yzfac = gy[iy] * zfac1 + hy[iy] * zfac2
                hzyfac1 = hz_iz * yfac1
                hzyfac2 = hz_iz * yfac2
                jxh = 0.0
                for ix in range(lo[0], hi[0] + 1):
                    cx = x1 + ix
                    xfac1 = gx[ix] + 0.5 * hx[ix]
                    xfac2 = third * hx[ix] + 0.5 * gx[ix]
                    wx = hx[ix] * yzfac
                    wy = xfac1 * hygz + xfac2 * hyhz
                    wz = gx[ix] * hzyfac1 + hx[ix] * hzyfac2
                    jxh = jxh - fjx * wx
                    jyh[ix] = jyh[ix] - fjy * wy
                    jzh[(ix, iy)] = jzh[(ix, iy)] - fjz * wz
                    o = (cz + NG - 1, cy + NG - 1, cx + NG - 1)
                    jx[o] += jxh
                    jy[o] += jyh[ix]
                    jz[o] += jzh[(ix, iy)]
    return jx, jy, jz


import pytest


@pytest.mark.parametrize("hc_push", [False, True])
@pytest.mark.parametrize("ndims,n", [(1, (40,)), (2, (14, 11)), (3, (8, 7, 6))])
def test_oracle_push_equals_an_independent_restatement_bit_for_bit(ndims, n, hc_push):
    dk = decks.thermal(ndims, n, ppc=6 if ndims < 3 else 3, temp_k=4.0e9)   # hot: many particles change cell in a step
    dk.hc_push = hc_push
    o = Oracle(dk)
    o.auto_load()
    o.init()
    rng = np.random.default_rng(5)
    fields = {}
    for name in ("ex", "ey", "ez", "bx", "by", "bz"):
        a = o.field(0, name)
        a[...] = rng.standard_normal(a.shape) * (2.0e11 if name[0] == "e" else 4.0e2)
        fields[name] = {1: a[0, 0], 2: a[0], 3: a}[ndims].copy()
    p = o.get_particles(0, 0)
    p0 = p.copy()
    info = o.rank_info(0)
    s = dk.species[0]
    push = {1: push_particles_1d, 2: push_particles_2d, 3: push_particles_3d}[ndims]
    mine = push(dk, fields, p, s.charge, s.mass, info["grid_min_local"], (s.charge, dk.dt(), s.mass) if hc_push else None)
    o.push_only()
    q = o.get_particles(0, 0)
    assert q.shape == p.shape
    assert np.array_equal(q, p)
    # the case is not a trivial one: a good share of the particles changed cell (the shifted-weight branches of
    # hx_dcell.inc and the widened deposit ranges ran), in every axis and both directions
    for d in range(ndims):
        cell = lambda a: np.floor((a[:, d] - info["grid_min_local"][d]) / dk.dx(d) + 0.5)
        dc = cell(q) - cell(p0)
        assert (dc > 0).mean() > 0.03 and (dc < 0).mean() > 0.03, d
    for name, j in zip(("jx", "jy", "jz"), mine):
        theirs = o.field(0, name)
        theirs = {1: theirs[0, 0], 2: theirs[0], 3: theirs}[ndims]
        assert np.abs(theirs).max() > 0
        assert np.array_equal(theirs, j), name
