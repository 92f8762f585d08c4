"""oracle/np_restatement.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED.

A second, independent restatement of mod_tsadvc.F90 in whole-array numpy, written
from the Fortran text (not from oracle/tsadvc_oracle.c) and deliberately different in
form: every sweep is one masked array expression on its margin region, sea-only
neighbours are `where(ip_neighbour, shifted, centre)` and the coast-zeroing passes
(mod_tsadvc.F90:738-758) are replaced by the rule "a face that is not an iu/iv point
but borders a sea cell is zero".  tests/test_oracle.py demands bit equality between
this file and the C oracle on 1:ii,1:jj; that pins the C loops against transcription
slips (index, margin, operation-order) and checks the coast-zero == mask-select claim
the CUDA kernels rely on.

numpy's elementwise + - * / on float64 are IEEE round-to-nearest and never fused, so
the arithmetic is the unfused Fortran order.

Arrays are (nrows, ncols) = Fortran a(1-nbdy:idm+nbdy, 1-nbdy:jdm+nbdy), i fastest.
"""
from __future__ import annotations

import numpy as np

ONEMU = 9806.0e-12  # mod_tsadvc.F90:236,519,671,1396


def _sh(a, di, dj):
    """a(i+di, j+dj) as an array indexed by (i,j); wrapped edges are never used"""
    return np.roll(a, (-dj, -di), axis=(0, 1))


def _region(geom, margin):
    nb = geom.nbdy
    r = np.zeros((geom.nrows, geom.ncols), dtype=bool)
    r[nb - margin: nb + geom.jj + margin, nb - margin: nb + geom.ii + margin] = True
    return r


def _fmax(a, b):  # Fortran max(a,b)
    return np.where(a > b, a, b)


def _fmin(a, b):
    return np.where(a < b, a, b)


def _extrema5(f, ip):
    """max/min over the cell and its sea-only neighbours (ipim1.., bigrid.F90:322-341)"""
    c = f
    w = np.where(_sh(ip, -1, 0) != 0, _sh(f, -1, 0), c)
    e = np.where(_sh(ip, +1, 0) != 0, _sh(f, +1, 0), c)
    s = np.where(_sh(ip, 0, -1) != 0, _sh(f, 0, -1), c)
    n = np.where(_sh(ip, 0, +1) != 0, _sh(f, 0, +1), c)
    mx = _fmax(_fmax(_fmax(_fmax(c, w), e), s), n)
    mn = _fmin(_fmin(_fmin(_fmin(c, w), e), s), n)
    return mx, mn


def _upwind(fld, u, v):
    qx = np.where(u >= 0.0, _sh(fld, -1, 0), fld)
    qy = np.where(v >= 0.0, _sh(fld, 0, -1), fld)
    return u * qx, v * qy


def _faces(geom, ip, iu, iv, margin, fx_val, fy_val):
    """flux arrays after a sweep at `margin` followed by its coast-zeroing pass"""
    reg = _region(geom, margin)
    nan = np.full(ip.shape, np.nan)
    # faces bordering at least one sea cell but not iu/iv points are land faces: 0
    ucoast = (iu == 0) & ((ip != 0) | (_sh(ip, -1, 0) != 0))
    vcoast = (iv == 0) & ((ip != 0) | (_sh(ip, 0, -1) != 0))
    fx = np.where(reg & (iu != 0), fx_val, np.where(ucoast, 0.0, nan))
    fy = np.where(reg & (iv != 0), fy_val, np.where(vcoast, 0.0, nan))
    return fx, fy


def advem_fct(geom, order, fld, fldc, u, v, fco, fcn, scal, scali, dt2, ip, iu, iv):
    """advem_fct2 (:645-997) for order=2, advem_fct4 (:1370-1706) for order=4"""
    with np.errstate(all="ignore"):
        sea = ip != 0
        # S1, margin 4
        fx, fy = _upwind(fld, u, v)
        flx, fly = _faces(geom, ip, iu, iv, 4, fx, fy)
        fmx, fmn = _extrema5(fld, ip)
        # S2, margin 3
        r3 = _region(geom, 3) & sea
        flxdiv = ((_sh(flx, 1, 0) - flx) + (_sh(fly, 0, 1) - fly)) * dt2 * scali
        q = fld * (fco + ONEMU) - flxdiv
        fldlo = np.where(r3, _fmax(fmn, _fmin(fmx, q / (fcn + ONEMU))), np.nan)
        fmxlo = np.where(r3, _fmax(_fmax(fld, fldc), fldlo), np.nan)
        fmnlo = np.where(r3, _fmin(_fmin(fld, fldc), fldlo), np.nan)
        # S3, margin 3
        fhx = u * 0.5 * (fldc + _sh(fldc, -1, 0))
        fhy = v * 0.5 * (fldc + _sh(fldc, 0, -1))
        if order == 4:
            ft14, ft24 = 7.0 / 12.0, -1.0 / 12.0
            fhx4 = u * (ft14 * (fldc + _sh(fldc, -1, 0)) + ft24 * (_sh(fldc, 1, 0) + _sh(fldc, -2, 0)))
            fhy4 = v * (ft14 * (fldc + _sh(fldc, 0, -1)) + ft24 * (_sh(fldc, 0, 1) + _sh(fldc, 0, -2)))
            fhx = np.where((_sh(iu, -1, 0) == 0) | (_sh(iu, 1, 0) == 0), fhx, fhx4)
            fhy = np.where((_sh(iv, 0, -1) == 0) | (_sh(iv, 0, 1) == 0), fhy, fhy4)
        fax, fay = _faces(geom, ip, iu, iv, 3, fhx - flx, fhy - fly)
        # S4, margin 2
        r2 = _region(geom, 2) & sea
        fqmax, _ = _extrema5(fmxlo, ip)
        _, fqmin = _extrema5(fmnlo, ip)
        fax_ib = np.where(_sh(ip, 1, 0) != 0, _sh(fax, 1, 0), fax)   # fax(ib,j)  :880
        fay_jb = np.where(_sh(ip, 0, 1) != 0, _sh(fay, 0, 1), fay)   # fay(i,jb)
        zero = 0.0
        famax = _fmax(zero, fax) - _fmin(zero, fax_ib) + _fmax(zero, fay) - _fmin(zero, fay_jb)
        famin = _fmax(zero, fax_ib) - _fmin(zero, fax) + _fmax(zero, fay_jb) - _fmin(zero, fay)
        qdt2 = 1.0 / dt2
        qp = (fqmax - fldlo) * fcn * scal * qdt2
        qm = (fldlo - fqmin) * fcn * scal * qdt2
        rp = np.where(famax > 0.0, np.where(qp < famax, qp / famax, 1.0), 0.0)
        rm = np.where(famin > 0.0, np.where(qm < famin, qm / famin, 1.0), 0.0)
        rp = np.where(r2, rp, np.nan)
        rm = np.where(r2, rm, np.nan)
        # S5, margin 1
        r1 = _region(geom, 1)
        factx = np.where(fax < 0.0, _fmin(_sh(rp, -1, 0), rm), _fmin(rp, _sh(rm, -1, 0)))
        facty = np.where(fay < 0.0, _fmin(_sh(rp, 0, -1), rm), _fmin(rp, _sh(rm, 0, -1)))
        fax = np.where(r1 & (iu != 0), factx * fax, fax)
        fay = np.where(r1 & (iv != 0), facty * fay, fay)
        # S6, margin 0
        r0 = _region(geom, 0) & sea
        flxdiv = ((_sh(fax, 1, 0) - fax) + (_sh(fay, 0, 1) - fay)) * dt2 * scali
        new = _fmax(fqmin, _fmin(fqmax, fldlo - flxdiv / (fcn + ONEMU)))
        out = np.where(r0, new, fld)
        inter = dict(flx=flx, fly=fly, fldlo=fldlo, fmxlo=fmxlo, fmnlo=fmnlo, rp=rp, rm=rm,
                     fax=fax, fay=fay)
        return out, inter


def advem_mpdata(geom, fld, u, v, fco, fcn, posdef, scal, scali, dt2, ip, iu, iv):
    """advem_mpdata (:207-493)"""
    with np.errstate(all="ignore"):
        sea = ip != 0
        # M1, margin 4
        tx1 = .5 * np.abs(u) * (fld - _sh(fld, -1, 0))
        ty1 = .5 * np.abs(v) * (fld - _sh(fld, 0, -1))
        qx = np.where(u >= 0.0, _sh(fld, -1, 0), fld)
        qy = np.where(v >= 0.0, _sh(fld, 0, -1), fld)
        flx, fly = _faces(geom, ip, iu, iv, 4, u * (qx + posdef), v * (qy + posdef))
        fmx, fmn = _extrema5(fld, ip)
        fmx = fmx + posdef
        fmn = fmn + posdef
        # M2, margin 3
        r3 = _region(geom, 3) & sea
        flxdiv = ((_sh(flx, 1, 0) - flx) + (_sh(fly, 0, 1) - fly)) * dt2 * scali
        q = (fld + posdef) * (fco + ONEMU) - flxdiv
        fldlo = np.where(r3, _fmax(fmn, _fmin(fmx, q / (fcn + ONEMU))), np.nan)
        flxdiv = np.where(r3, flxdiv, np.nan)
        # M3, margin 3: only iu/iv points are rewritten, coast faces keep their zero
        r3f = _region(geom, 3)
        fco2 = fco + _sh(fco, -1, 0)
        fcn2 = fcn + _sh(fcn, -1, 0)
        fx = tx1 - u * (flxdiv + _sh(flxdiv, -1, 0)) / ((fco2 + fcn2) + ONEMU)
        fco2 = fco + _sh(fco, 0, -1)
        fcn2 = fcn + _sh(fcn, 0, -1)
        fy = ty1 - v * (flxdiv + _sh(flxdiv, 0, -1)) / ((fco2 + fcn2) + ONEMU)
        flx = np.where(r3f & (iu != 0), fx, flx)
        fly = np.where(r3f & (iv != 0), fy, fly)
        # M4, margin 2
        r2 = _region(geom, 2) & sea
        zero = 0.0
        flxdp = _fmin(zero, _sh(flx, 1, 0)) - _fmax(zero, flx)
        flxdn = _fmax(zero, _sh(flx, 1, 0)) - _fmin(zero, flx)
        flydp = _fmin(zero, _sh(fly, 0, 1)) - _fmax(zero, fly)
        flydn = _fmax(zero, _sh(fly, 0, 1)) - _fmin(zero, fly)
        rp = (fmx - fldlo) * (fcn * scal) / ((ONEMU - (flxdp + flydp)) * dt2)
        rm = (fldlo - fmn) * (fcn * scal) / ((ONEMU + (flxdn + flydn)) * dt2)
        rp = np.where(r2, rp, np.nan)
        rm = np.where(r2, rm, np.nan)
        # M5, margin 1
        r1 = _region(geom, 1)
        one = 1.0
        fx = _fmax(zero, flx) * _fmin(_fmin(one, rp), _sh(rm, -1, 0)) + \
            _fmin(zero, flx) * _fmin(_fmin(one, _sh(rp, -1, 0)), rm)
        fy = _fmax(zero, fly) * _fmin(_fmin(one, rp), _sh(rm, 0, -1)) + \
            _fmin(zero, fly) * _fmin(_fmin(one, _sh(rp, 0, -1)), rm)
        flx = np.where(r1 & (iu != 0), fx, flx)
        fly = np.where(r1 & (iv != 0), fy, fly)
        # M6, margin 0
        r0 = _region(geom, 0) & sea
        flxdiv = ((_sh(flx, 1, 0) - flx) + (_sh(fly, 0, 1) - fly)) * dt2 * scali
        new = _fmax(fmn, _fmin(fmx, fldlo - flxdiv / (fcn + ONEMU))) - posdef
        return np.where(r0, new, fld), dict(fldlo=fldlo, rp=rp, rm=rm, flx=flx, fly=fly)


def advem_pcm(geom, fld, u, v, fco, fcn, scal, scali, dt2, ip, iu, iv):
    """advem_pcm (:495-643)"""
    with np.errstate(all="ignore"):
        sea = ip != 0
        fx, fy = _upwind(fld, u, v)
        flx, fly = _faces(geom, ip, iu, iv, 1, fx, fy)
        fmx, fmn = _extrema5(fld, ip)
        r0 = _region(geom, 0) & sea
        flxdiv = ((_sh(flx, 1, 0) - flx) + (_sh(fly, 0, 1) - fly)) * dt2 * scali
        q = fld * (fco + ONEMU) - flxdiv
        return np.where(r0, _fmax(fmn, _fmin(fmx, q / (fcn + ONEMU))), fld), {}


def advem_fct2c(geom, fld, fldc, u, v, fco, scal, scali, dt2, ip, iu, iv):
    """advem_fct2c (mod_tsadvc.F90:999-1368): FCT2 with a sub-cycled low-order step (5 iterations,
    local time step dtloc, per-face elapsed time ucumdt/vcumdt), single-tile xctilr of hloc and
    fldlo after every iteration.  Whole-array form: state arrays are updated through masks."""
    epsil = 1.0e-10
    with np.errstate(all="ignore"):
        sea, su, sv = ip != 0, iu != 0, iv != 0
        shp = ip.shape
        lcalc = np.ones(shp, dtype=bool)
        flxcum, flycum = np.zeros(shp), np.zeros(shp)
        hloc, fldlo = fco.copy(), fld.copy()
        ucum, vcum = np.zeros(shp), np.zeros(shp)
        uloc, vloc, flx, fly = np.zeros(shp), np.zeros(shp), np.zeros(shp), np.zeros(shp)
        dtloc = np.zeros(shp)
        up, vp = u >= 0, v >= 0
        for _ in range(5):
            q = _fmax(_sh(u, 1, 0), 0.0) - _fmin(u, 0.0) + _fmax(_sh(v, 0, 1), 0.0) - _fmin(v, 0.0)
            dtloc = np.where(_region(geom, 5) & sea, np.where(q > 0.0, _fmin(dt2, hloc / (q * scali)), dt2), dtloc)
            r4 = _region(geom, 4)
            # x faces
            act = r4 & su & (ucum != dt2)
            dtu = _fmin(dt2 - ucum, np.where(up, _sh(dtloc, -1, 0), dtloc))
            ul = dtu * u
            fx = np.where(up, _sh(fldlo, -1, 0), fldlo) * ul
            uloc = np.where(act, ul, np.where(r4 & su, 0.0, uloc))
            flx = np.where(act, fx, np.where(r4 & su, 0.0, flx))
            ucum = np.where(act, ucum + dtu, ucum)
            flxcum = np.where(act, flxcum + flx, flxcum)
            # y faces
            act = r4 & sv & (vcum != dt2)
            dtv = _fmin(dt2 - vcum, np.where(vp, _sh(dtloc, 0, -1), dtloc))
            vl = dtv * v
            fy = np.where(vp, _sh(fldlo, 0, -1), fldlo) * vl
            vloc = np.where(act, vl, np.where(r4 & sv, 0.0, vloc))
            fly = np.where(act, fy, np.where(r4 & sv, 0.0, fly))
            vcum = np.where(act, vcum + dtv, vcum)
            flycum = np.where(act, flycum + fly, flycum)
            # cells
            go = _region(geom, 3) & sea & lcalc
            qp = hloc - (_sh(uloc, 1, 0) - uloc + _sh(vloc, 0, 1) - vloc) * scali
            new = ((epsil + hloc) * fldlo - (_sh(flx, 1, 0) - flx + _sh(fly, 0, 1) - fly) * scali) / (epsil + qp)
            fldlo = np.where(go & (qp > 0.0), new, fldlo)
            hloc = np.where(go, qp, hloc)
            more = (_sh(ucum, 1, 0) != dt2) | (ucum != dt2) | (_sh(vcum, 0, 1) != dt2) | (vcum != dt2)
            lcalc = np.where(go, more, lcalc)
            hloc = halo_single_tile(geom, hloc, 5, 5)
            fldlo = halo_single_tile(geom, fldlo, 5, 5)
        fhx = u * 0.5 * (fldc + _sh(fldc, -1, 0))
        fhy = v * 0.5 * (fldc + _sh(fldc, 0, -1))
        fax, fay = _faces(geom, ip, iu, iv, 3, fhx - flxcum / dt2, fhy - flycum / dt2)
        fqmax, fqmin = _extrema5(fldlo, ip)
        faxe, fayn = _sh(fax, 1, 0), _sh(fay, 0, 1)
        famax = _fmax(0.0, fax) - _fmin(0.0, faxe) + _fmax(0.0, fay) - _fmin(0.0, fayn)
        famin = _fmax(0.0, faxe) - _fmin(0.0, fax) + _fmax(0.0, fayn) - _fmin(0.0, fay)
        qdt2 = 1.0 / dt2
        qp = (fqmax - fldlo) * hloc * scal * qdt2
        qm = (fldlo - fqmin) * hloc * scal * qdt2
        r2 = _region(geom, 2) & sea
        rp = np.where(r2, np.where(famax > epsil, np.where(qp < famax, qp / famax, 1.0), 0.0), np.nan)
        rm = np.where(r2, np.where(famin > epsil, np.where(qm < famin, qm / famin, 1.0), 0.0), np.nan)
        fx = np.where(fax < 0.0, _fmin(_sh(rp, -1, 0), rm), _fmin(rp, _sh(rm, -1, 0))) * fax
        fy = np.where(fay < 0.0, _fmin(_sh(rp, 0, -1), rm), _fmin(rp, _sh(rm, 0, -1))) * fay
        r1 = _region(geom, 1)
        fax = np.where(r1 & su, fx, fax)
        fay = np.where(r1 & sv, fy, fay)
        flxdiv = ((_sh(fax, 1, 0) - fax) + (_sh(fay, 0, 1) - fay)) * dt2 * scali
        new = np.where(hloc > 0.0, ((epsil + hloc) * fldlo - flxdiv) / (epsil + hloc), fldlo)
        return np.where(_region(geom, 0) & sea, new, fld), {}


def prolog(geom, uflx_k, vflx_k, dp_kn, onetamas_m, delt1, scp2i, ip, margin):
    """tsadvc prolog, mod_tsadvc.F90:1905-1942: util1 = fco, util2 = fcn"""
    with np.errstate(all="ignore"):
        reg = _region(geom, margin) & (ip != 0)
        flxdiv = ((_sh(uflx_k, 1, 0) - uflx_k) + (_sh(vflx_k, 0, 1) - vflx_k)) * delt1 * scp2i
        fco = np.where(reg, _fmax(onetamas_m * dp_kn + flxdiv, 0.0), np.nan)
        fcn = np.where(reg, _fmax(onetamas_m * dp_kn, 0.0), np.nan)
        return fco, fcn


def halo_single_tile(geom, a, mh, nh, itype=1):
    """xctilr on one tile, mod_xc_sm.h:1337-1428 (nreg=2: the arctic version, :1172-1335, where
    itype selects grid and sign, mod_xc.F90:41-44), for an array (..., nrows, ncols)"""
    nb, ii, jj = geom.nbdy, geom.ii, geom.jj
    a = a.copy()
    i1, j1 = nb, nb   # numpy index of i=1 / j=1
    if geom.nreg == 2:
        grid, sgn = itype % 10, (1.0 if itype < 10 else -1.0)
        i = np.arange(1, ii + 1)
        io = (ii - (i - 1) % ii) if grid in (1, 4) else ((ii - (i - 1)) % ii + 1)
        for j in range(1, nh + 1):
            a[..., j1 - j, i1:i1 + ii] = 0.0
            jo = (jj - 1 - j) if grid in (1, 3) else (jj - j)
            src = a[..., j1 + jo - 1, :][..., i1 + io - 1]
            a[..., j1 + jj - 1 + j, i1:i1 + ii] = src if itype < 10 else sgn * src
    for j in range(1, nh + 1):
        if geom.nreg == 2:
            break
        if geom.nreg <= 2:
            a[..., j1 - j, i1:i1 + ii] = 0.0
            a[..., j1 + jj - 1 + j, i1:i1 + ii] = 0.0
        else:
            a[..., j1 - j, i1:i1 + ii] = a[..., j1 + jj - j, i1:i1 + ii]
            a[..., j1 + jj - 1 + j, i1:i1 + ii] = a[..., j1 + j - 1, i1:i1 + ii]
    rows = slice(j1 - nh, j1 + jj + nh)
    for i in range(1, mh + 1):
        if geom.nreg in (0, 4):
            a[..., rows, i1 - i] = 0.0
            a[..., rows, i1 + ii - 1 + i] = 0.0
        else:
            a[..., rows, i1 - i] = a[..., rows, i1 + ii - i]
            a[..., rows, i1 + ii - 1 + i] = a[..., rows, i1 + i - 1]
    return a


def tsadvc(cb, m, n):
    """tsadvc(m,n), hybrid coordinates (mod_tsadvc.F90:1804-2230; diffusion when temdf2>0).
    `cb` is a product-side CbArrays of host numpy arrays; returns new slot-n fields."""
    g = cb.geom
    kk = g.kdm
    mbdy = {0: 2, 1: 5, 2: 5, 4: 5}[abs(cb.advtyp)]
    nhyb = kk if cb.nhybrd < 0 else cb.nhybrd
    temp = halo_single_tile(g, cb.temp, mbdy, mbdy)
    saln = halo_single_tile(g, cb.saln, mbdy, mbdy)
    th3d = halo_single_tile(g, cb.th3d, mbdy, mbdy)
    uflx = halo_single_tile(g, cb.uflx, mbdy, mbdy, 13)   # halo_uv
    vflx = halo_single_tile(g, cb.vflx, mbdy, mbdy, 14)   # halo_vv
    tracer = halo_single_tile(g, cb.tracer, mbdy, mbdy) if cb.ntracr else None
    q2 = halo_single_tile(g, cb.q2, mbdy, mbdy) if cb.mxlmy else None
    q2l = halo_single_tile(g, cb.q2l, mbdy, mbdy) if cb.mxlmy else None
    # onetamas(:,:,m) = 1.0 (:1809), or oneta(:,:,n) when btrmas (:1806)
    oem = cb.oneta[n - 1] if cb.btrmas else np.ones((g.nrows, g.ncols))
    ip, iu, iv = cb.ip, cb.iu, cb.iv

    def adv(fld_n, fld_m, k, posdef, fco, fcn, uf=None, vf=None):
        uk = uflx[k] if uf is None else uf
        vk = vflx[k] if vf is None else vf
        a = (g, fld_n, fld_m, uk, vk, fco, fcn, cb.scp2, cb.scp2i, cb.delt1, ip, iu, iv)
        if cb.advtyp == 2 and cb.btrmas:
            return advem_fct2c(g, fld_n, fld_m, uk, vk, fco, cb.scp2, cb.scp2i, cb.delt1, ip, iu, iv)[0]
        if cb.advtyp == 2:
            return advem_fct(a[0], 2, *a[1:])[0]
        if cb.advtyp == 4:
            return advem_fct(a[0], 4, *a[1:])[0]
        if cb.advtyp == 1:
            return advem_mpdata(g, fld_n, uk, vk, fco, fcn, posdef, cb.scp2, cb.scp2i,
                                cb.delt1, ip, iu, iv)[0]
        if cb.advtyp == 0:
            return advem_pcm(g, fld_n, uk, vk, fco, fcn, cb.scp2, cb.scp2i, cb.delt1,
                             ip, iu, iv)[0]
        raise ValueError(cb.advtyp)

    usm = vsm = None
    if cb.isopyc:
        # :1859-1897 layer 1: th3d and saln run on laterally smoothed mass fluxes (0.0 off the iu/iv points
        # and outside margin mbdy-1, :1812-1813) and so does the prolog (:1930-1932); the tracers and
        # q2, q2l keep uflx(:,:,1) (:2016-2048)
        reg = _region(g, mbdy - 1)
        u1, v1 = uflx[0], vflx[0]
        vfa = np.where(_sh(iv, -1, 0) != 0, _sh(v1, -1, 0), v1)
        vfb = np.where(_sh(iv, +1, 0) != 0, _sh(v1, +1, 0), v1)
        ufa = np.where(_sh(iu, 0, -1) != 0, _sh(u1, 0, -1), u1)
        ufb = np.where(_sh(iu, 0, +1) != 0, _sh(u1, 0, +1), u1)
        with np.errstate(all="ignore"):
            usm = np.where(reg & (iu != 0), .5 * u1 + .25 * (ufa + ufb), 0.0)
            vsm = np.where(reg & (iv != 0), .5 * v1 + .25 * (vfa + vfb), 0.0)
    for k in range(kk):
        smooth = cb.isopyc and k == 0
        fco, fcn = prolog(g, usm if smooth else uflx[k], vsm if smooth else vflx[k], cb.dp[n - 1, k], oem,
                          cb.delt1, cb.scp2i, ip, mbdy - 1)
        ts = (usm, vsm) if smooth else (None, None)     # fluxes of the thermodynamic fields
        if cb.isopyc and k == 0 and not (cb.advflg == 1 and nhyb > 0):
            th3d[n - 1, k] = adv(th3d[n - 1, k], th3d[m - 1, k], k, 32.0, fco, fcn, *ts)
        if k + 1 <= nhyb:
            if cb.advflg == 0:
                temp[n - 1, k] = adv(temp[n - 1, k], temp[m - 1, k], k, 256.0, fco, fcn)
            else:
                th3d[n - 1, k] = adv(th3d[n - 1, k], th3d[m - 1, k], k, 32.0, fco, fcn)
        saln[n - 1, k] = adv(saln[n - 1, k], saln[m - 1, k], k, 0.0, fco, fcn, *ts)
        for q in range(cb.ntracr):
            pd = 256.0 if (q < len(cb.trcflg) and cb.trcflg[q] == 2) else 0.0
            tracer[q, n - 1, k] = adv(tracer[q, n - 1, k], tracer[q, m - 1, k], k, pd, fco, fcn)
        if cb.mxlmy:   # :2035-2048, layer k of q2(.., 0:kk+1, ..) is index k+1
            q2[n - 1, k + 1] = adv(q2[n - 1, k + 1], q2[m - 1, k + 1], k, 1.0, fco, fcn)
            q2l[n - 1, k + 1] = adv(q2l[n - 1, k + 1], q2l[m - 1, k + 1], k, 1.0, fco, fcn)
    if cb.temdf2 > 0.0:
        diffuse(cb, n, temp, saln, th3d, tracer)
        if cb.mxlmy:   # :2143-2146, :2180-2183
            for a in (q2, q2l):
                a[n - 1] = halo_single_tile(g, a[n - 1], 2, 2)
            for k in range(kk):
                q2[n - 1, k + 1], q2l[n - 1, k + 1] = tsdff(g, [q2[n - 1, k + 1], q2l[n - 1, k + 1]], cb.dp[n - 1, k],
                                                            cb.oneta[n - 1], cb, ip, iu, iv)
    return dict(temp=temp, saln=saln, th3d=th3d, tracer=tracer, q2=q2, q2l=q2l)


# ---- diffusion and equation of state (mod_tsadvc.F90:2138-2492, stmt_fns.h) ----------------

_C79 = {  # stmt_fns.h:53-61, :65-73, :86-96, :100-110
    1: (-1.36471E-01, 4.68181E-02, 8.07004E-01, -7.45353E-03, -2.94418E-03, 3.43570E-05, 3.48658E-05, 0.0, 0.0),
    2: (9.77093E+00, -2.26493E-02, 7.89879E-01, -6.43205E-03, -2.62983E-03, 2.75835E-05, 3.15235E-05, 0.0, 0.0),
    3: (-4.311829E-02, 5.429948E-02, 8.011774E-01, -7.641336E-03, -3.258442E-03, 3.757643E-05, 3.630361E-05,
        8.675546E-05, 3.995086E-06),
    4: (9.903308E+00, -1.618075E-02, 7.819166E-01, -6.593939E-03, -2.896464E-03, 3.038697E-05, 3.266933E-05,
        1.180109E-04, 3.399511E-06),
}
_C12 = dict(  # stmt_fns.h:129-149
    c001=-1.4627567840659594e-01, c002=6.4247392832635697e-02, c003=8.1213979591704621e-01,
    c004=-8.1321489441909698e-03, c005=4.5199845091090296e-03, c006=4.6347888132781394e-04,
    c007=5.0879498675039621e-03, c008=1.6333913018305079e-05, c009=4.3899924880543972e-06,
    c011=1.0000000000000000e+00, c012=1.0316374535350838e-02, c013=8.9521792365142522e-04,
    c014=-2.8438341552142710e-05, c015=-1.1887778959461776e-05, c016=-4.0163964812921489e-06,
    c017=1.1995545126831476e-05, c018=5.5234008384648383e-08, c019=8.4310335919950873e-09)
_C17 = dict(  # stmt_fns.h:216-242
    c001=9.9984085444849347e+02, c002=7.3471625860981584e+00, c003=-5.3211231792841769e-02,
    c004=3.6492439109814549e-04, c005=2.5880571023991390e+00, c006=6.7168282786692355e-03,
    c007=1.9203202055760151e-03, c008=1.0000000000000000e+00, c009=7.2815210113327091e-03,
    c010=-4.4787265461983921e-05, c011=3.3851002965802430e-07, c012=1.3651202389758572e-10,
    c013=1.7632126669040377e-03, c014=8.8066583251206474e-06, c015=1.8832689434804897e-10,
    c016=5.7463776745432097e-06, c017=1.4716275472242334e-09, c018=1.1798263740430364e-02,
    c019=9.8920219266399117e-08, c020=4.6996642771754730e-06, c021=2.5862187075154352e-08,
    c022=3.2921414007960662e-12, c023=6.7103246285651894e-06, c024=2.4461698007024582e-17,
    c025=9.1534417604289062e-18)


def _rpdb(sigma2):
    return np.float64(2000.0e4 if sigma2 else 0.0) * np.float64(1.0e-4)


def _c12(sigver):
    k = _C12
    r = _rpdb(sigver == 8)
    return dict(k, c101=k["c001"] + r * k["c007"], c102=k["c002"] + r * k["c008"], c103=k["c003"] + r * k["c009"],
                c111=k["c011"] + r * k["c017"], c112=k["c012"] + r * k["c018"], c113=k["c013"] + r * k["c019"])


def sig(sigver, t, s):
    """sig(t,s) of the equation of state `sigver` (stmt_fns.h:332, :368, :419-424, :503-509)"""
    t = np.asarray(t, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    with np.errstate(all="ignore"):
        if sigver in (1, 2):
            c1, c2, c3, c4, c5, c6, c7, _, _ = _C79[sigver]
            return (c1 + c3 * s + t * (c2 + c5 * s + t * (c4 + c7 * s + c6 * t)))
        if sigver in (3, 4):
            c1, c2, c3, c4, c5, c6, c7, c8, c9 = _C79[sigver]
            return (c1 + s * (c3 + s * c8) + t * (c2 + s * (c5 + s * c9) + t * (c4 + s * c7 + t * c6)))
        if sigver in (7, 8):
            k = _c12(sigver)
            n = k["c101"] + (k["c102"] + k["c004"] * t + k["c005"] * s) * t + (k["c103"] + k["c006"] * s) * s
            d = k["c111"] + (k["c112"] + k["c014"] * t + k["c015"] * s) * t + (k["c113"] + k["c016"] * s) * s
            return n * (1.0 / d)
        k = _C17
        r = _rpdb(sigver == 6)
        c101 = k["c001"] + (k["c018"] - k["c021"] * r) * r
        c103 = k["c003"] + (k["c019"] - k["c022"] * r) * r
        c105 = k["c005"] + k["c020"] * r
        c108 = k["c008"] + k["c023"] * r
        c109 = k["c009"] - k["c025"] * (r * r * r)
        c111 = k["c011"] - k["c024"] * (r * r)
        n = c101 + t * (k["c002"] + t * (c103 + t * k["c004"])) + s * (c105 - t * k["c006"] + s * k["c007"])
        d = (c108 + t * (c109 + t * (k["c010"] + t * (c111 + t * k["c012"]))) +
             s * (k["c013"] - t * (k["c014"] + t * t * k["c015"]) +
                  np.sqrt(_fmax(0.0, s)) * (k["c016"] + t * t * k["c017"])))
        return n * (1.0 / d) - 1000.0


def tofsig(sigver, r, s):
    """tofsig(r,s) (stmt_fns.h:308-323, :349-379, :441-449; 99.0 for the 17-term fit, :533)"""
    r = np.asarray(r, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    with np.errstate(all="ignore"):
        if sigver in (1, 2, 3, 4):
            c1, c2, c3, c4, c5, c6, c7, c8, c9 = _C79[sigver]
            rc6 = np.float64(1.0) / np.float64(c6)
            if sigver <= 2:
                a0, a1, a2 = (c1 + c3 * s - r) * rc6, (c2 + c5 * s) * rc6, (c4 + c7 * s) * rc6
            else:
                a0, a1, a2 = (c1 + s * (c3 + s * c8) - r) * rc6, (c2 + s * (c5 + s * c9)) * rc6, (c4 + s * c7) * rc6
            a3rd = np.float64(1.0) / np.float64(3.0)
            x = a3rd * a2
            cubq = a3rd * a1 - x * x
            cubr = a3rd * (0.5 * a1 * a2 - 1.5 * a0) - x * x * x
            cuban = a3rd * np.arctan2(np.sqrt(_fmax(0.0, -(cubq * cubq * cubq + cubr * cubr))), cubr)
            cubrl = np.sqrt(-cubq) * np.cos(cuban)
            cubim = np.sqrt(-cubq) * np.sin(cuban)
            return -cubrl + np.sqrt(np.float64(3.0)) * cubim - a3rd * a2
        if sigver in (7, 8):
            k = _c12(sigver)
            a = (k["c004"] - r * k["c014"])
            b = ((k["c102"] + k["c005"] * s) - r * (k["c112"] + k["c015"] * s))
            c = ((k["c101"] + (k["c103"] + k["c006"] * s) * s) - r * (k["c111"] + (k["c113"] + k["c016"] * s) * s))
            return (-b - np.sqrt(_fmax(0.0, b * b - 4.0 * a * c))) / (2.0 * a)
        return np.full(np.broadcast(r, s).shape, 99.0)


def _harmonc(aa, bb):
    """mod_tsadvc.F90:1770-1771"""
    a, b = _fmax(aa, 0.0), _fmax(bb, 0.0)
    return 2.0 * a * b / _fmax((a + b), 2.0 * 1.0e-20)


def tsdff(geom, flds, dp_kn, oneta_n, cb, ip, iu, iv):
    """tsdff_1x / tsdff_2x (mod_tsadvc.F90:2262-2492) for a list of fields of one layer: the
    face fluxes as whole-array expressions (0.0 on land faces: :1812-1813, geopar.F90:826-843),
    then the update on the sea points of 1:ii,1:jj"""
    with np.errstate(all="ignore"):
        h = dp_kn * oneta_n
        fu = cb.temdf2 * cb.aspux * cb.scuy * _harmonc(_sh(h, -1, 0), h)
        fv = cb.temdf2 * cb.aspvy * cb.scvx * _harmonc(_sh(h, 0, -1), h)
        factor = -cb.delt1 / (cb.scp2 * _fmax(h, 1.0e-20))
        upd = _region(geom, 0) & (ip != 0)
        out = []
        for f in flds:
            ufl = np.where(iu != 0, fu * (_sh(f, -1, 0) - f), 0.0)
            vfl = np.where(iv != 0, fv * (_sh(f, 0, -1) - f), 0.0)
            util = ((_sh(ufl, 1, 0) - ufl) + (_sh(vfl, 0, 1) - vfl)) * factor
            out.append(np.where(upd, f + util, f))
        return out


def diffuse(cb, n, temp, saln, th3d, tracer):
    """mod_tsadvc.F90:2138-2230 on slot n of the (already advected) fields, in place"""
    g = cb.geom
    kk = g.kdm
    nhyb = kk if cb.nhybrd < 0 else cb.nhybrd
    ip, iu, iv = cb.ip, cb.iu, cb.iv
    for a in (saln, temp, th3d) + ((tracer,) if cb.ntracr else ()):
        a[..., n - 1, :, :, :] = halo_single_tile(g, a[..., n - 1, :, :, :], 2, 2)
    upd = _region(g, 0) & (ip != 0)
    for k in range(kk):
        ldtemp = (k + 1 <= nhyb) and cb.temdfc > 0.0
        ldth3d = ((k + 1 <= nhyb) and cb.temdfc < 1.0) or (k == 0 and cb.isopyc)
        dpk, on = cb.dp[n - 1, k], cb.oneta[n - 1]
        T, S, H = temp[n - 1, k], saln[n - 1, k], th3d[n - 1, k]
        if ldtemp and ldth3d:
            H, T = tsdff(g, [H, T], dpk, on, cb, ip, iu, iv)
            S, = tsdff(g, [S], dpk, on, cb, ip, iu, iv)
        elif ldtemp:
            T, S = tsdff(g, [T, S], dpk, on, cb, ip, iu, iv)
        elif ldth3d:
            H, S = tsdff(g, [H, S], dpk, on, cb, ip, iu, iv)
        else:
            S, = tsdff(g, [S], dpk, on, cb, ip, iu, iv)
        if cb.ntracr:
            new = tsdff(g, [tracer[q, n - 1, k] for q in range(cb.ntracr)], dpk, on, cb, ip, iu, iv)
            for q in range(cb.ntracr):
                tracer[q, n - 1, k] = new[q]
        # :2199-2229
        with np.errstate(all="ignore"):
            if ldtemp and ldth3d:
                th3d_t = sig(cb.sigver, T, S) - cb.thbase
                Hn = (1.0 - cb.temdfc) * H + cb.temdfc * th3d_t
                Tn = tofsig(cb.sigver, Hn + cb.thbase, S)
            elif ldtemp:
                Hn, Tn = sig(cb.sigver, T, S) - cb.thbase, T
            elif ldth3d:
                Hn, Tn = H, tofsig(cb.sigver, H + cb.thbase, S)
            else:
                Hn = cb.theta[k]
                Tn = tofsig(cb.sigver, Hn + cb.thbase, S)
        temp[n - 1, k] = np.where(upd, Tn, T)
        th3d[n - 1, k] = np.where(upd, Hn, H)
        saln[n - 1, k] = S


# ---- mod_asselin.F90 ---------------------------------------------

def asselin_save(cb, m, n):
    """asselin_save (mod_asselin.F90:28-82) on a CbArrays; returns the arrays it writes"""
    g = cb.geom
    upd = _region(g, 0)
    sea = upd & (cb.ip != 0)
    oneta, onetao = cb.oneta.copy(), cb.onetao.copy()
    with np.errstate(all="ignore"):
        for t in (n, m):
            v = _fmax(cb.oneta0, 1.0 + cb.pbavg[t - 1] / cb.pbot)
            oneta[t - 1] = np.where(sea, v, oneta[t - 1])
            onetao[t - 1] = np.where(sea, oneta[t - 1], onetao[t - 1])
    out = dict(oneta=halo_single_tile(g, oneta, 6, 6), onetao=halo_single_tile(g, onetao, 6, 6))
    for name, src in (("otemp", cb.temp), ("osaln", cb.saln), ("oth3d", cb.th3d)):
        out[name] = np.where(upd, src[n - 1], getattr(cb, name))
    if cb.ntracr:
        out["otracer"] = np.where(upd, cb.tracer[:, n - 1], cb.otracer)
    if cb.mxlmy:
        for name, src in (("oq2", cb.q2), ("oq2l", cb.q2l)):
            o = getattr(cb, name).copy()
            o[1:-1] = np.where(upd, src[n - 1, 1:-1], o[1:-1])
            out[name] = o
    return out


def asselin_filter(cb, m, n):
    """asselin_filter (mod_asselin.F90:84-286): whole-array form; returns the arrays it writes"""
    g = cb.geom
    kk = g.kdm
    nhyb = kk if cb.nhybrd < 0 else cb.nhybrd
    sea = _region(g, 0) & (cb.ip != 0)
    onezm = 9806.0e-20
    ra = cb.ra2fac
    with np.errstate(all="ignore"):
        oneta = cb.oneta.copy()
        for t in (n, m):
            oneta[t - 1] = np.where(sea, _fmax(cb.oneta0, 1.0 + cb.pbavg[t - 1] / cb.pbot), oneta[t - 1])
        dp, temp, saln, th3d = cb.dp.copy(), cb.temp.copy(), cb.saln.copy(), cb.th3d.copy()
        tracer = cb.tracer.copy() if cb.ntracr else None
        q2 = cb.q2.copy() if cb.mxlmy else None
        q2l = cb.q2l.copy() if cb.mxlmy else None
        for k in range(kk):
            latemp = (k + 1 <= nhyb) and cb.advflg == 0
            lath3d = ((k + 1 <= nhyb) and cb.advflg == 1) or (k == 0 and cb.isopyc)
            dpold = cb.dpo[n - 1, k] * cb.onetao[n - 1]
            dpmid = cb.dpo[m - 1, k] * cb.onetao[m - 1]
            dpnew = cb.dp[n - 1, k] * oneta[n - 1]
            dpmidn = dpmid + 0.5 * ra * (dpold + dpnew - 2.0 * dpmid)
            dp[m - 1, k] = np.where(sea, dpmidn / oneta[m - 1], dp[m - 1, k])
            go = sea & (dpmidn > onezm)
            qd = 1.0 / dpmidn

            def ra_filter(o, fm, fn):
                smin = _fmin(_fmin(o, fm), fn)
                dpsold, dpsmid, dpsnew = dpold * (o - smin), dpmid * (fm - smin), dpnew * (fn - smin)
                return smin + (dpsmid + 0.5 * ra * (dpsold + dpsnew - 2.0 * dpsmid)) * qd
            S = ra_filter(cb.osaln[k], cb.saln[m - 1, k], cb.saln[n - 1, k])
            saln[m - 1, k] = np.where(go, S, saln[m - 1, k])
            if latemp:
                T = ra_filter(cb.otemp[k], cb.temp[m - 1, k], cb.temp[n - 1, k])
                H = sig(cb.sigver, T, S) - cb.thbase
            elif lath3d:
                H = ra_filter(cb.oth3d[k], cb.th3d[m - 1, k], cb.th3d[n - 1, k])
                T = tofsig(cb.sigver, H + cb.thbase, S)
            else:
                H = cb.theta[k]
                T = tofsig(cb.sigver, H + cb.thbase, S)
            temp[m - 1, k] = np.where(go, T, temp[m - 1, k])
            th3d[m - 1, k] = np.where(go, H, th3d[m - 1, k])
            for q in range(cb.ntracr):
                R = ra_filter(cb.otracer[q, k], cb.tracer[q, m - 1, k], cb.tracer[q, n - 1, k])
                tracer[q, m - 1, k] = np.where(go, R, tracer[q, m - 1, k])
            if cb.mxlmy:
                for arr, o, src in ((q2, cb.oq2, cb.q2), (q2l, cb.oq2l, cb.q2l)):
                    dpsold, dpsmid, dpsnew = dpold * o[k + 1], dpmid * src[m - 1, k + 1], dpnew * src[n - 1, k + 1]
                    R = (dpsmid + 0.5 * ra * (dpsold + dpsnew - 2.0 * dpsmid)) * qd
                    arr[m - 1, k + 1] = np.where(go, R, arr[m - 1, k + 1])
    return dict(oneta=oneta, dp=dp, temp=temp, saln=saln, th3d=th3d, tracer=tracer, q2=q2, q2l=q2l)


# -----------------------------------------------------------------------------------------
# cnuity(m,n) (cnuity.F90), second restatement: whole-array expressions, fluxes by the rule
# "zero unless an iu / iv point", layer coupling through a running sum of the old thicknesses.
# Scope as oracle/cnuity_oracle.inc.c (not btrmas, no interface smoothing, no open boundaries).
# -----------------------------------------------------------------------------------------
EPSIL = 1.0e-11   # mod_cb_arrays.F90:853


def cnuity(geom, st, m, n, ip, iu, iv, scuy, scvx, scp2i, depthu, depthv, pbot, delt1, ra2fac, isopyc=False, thk=None,
           mxlkta=None):
    """st: dict of arrays in the Fortran layout with halos valid to width 6 (the caller did the xctilr of
    cnuity.F90:100-107): dp, dpo (2,kk,..), u, v, dpu, dpv (2,kk,..), ubavg, vbavg (3,..), dpmixl (2,..),
    uflx, vflx, uflxav, vflxav, dpav (kk,..).  Updated in place; returns p (kk+1,..), utotn, vtotn, dpkmin.
    The halo refresh of dp(:,:,:,n) before the Robert-Asselin filter (:1400) is the caller's.
    thk: interface-depth diffusion (:745-1124), a dict with thkdf4u, thkdf4v (the coefficients at the u and v
    points), bih (True: biharmonic, thkdf4; False: Laplacian, thkdf2), nstep, scp2 and halo(a, itype), the
    xctilr of width 6 of :761-763.  mxlkta: hybrid .and. mxlkta (:1144-1324), a dict with onemm."""
    kk = geom.kdm
    dp, dpo = st["dp"], st["dpo"]
    n_, m_ = n - 1, m - 1
    shape = (geom.nrows, geom.ncols)
    R = {mg: _region(geom, mg) for mg in range(0, 7)}
    sea_p, sea_u, sea_v = ip != 0, iu != 0, iv != 0
    utotn, vtotn, util3 = np.zeros(shape), np.zeros(shape), np.zeros(shape)
    dpmold = st["dpmixl"][n_].copy()
    dpo[n_][:, R[6]] = dp[n_][:, R[6]]
    p = np.full((kk + 1,) + shape, np.nan)
    p[0] = 0.0
    dpkmin = np.full(2 * kk, np.nan)
    inner = R[0]
    with np.errstate(all="ignore"):
        for k in range(kk):
            d = dp[n_][k]
            # low-order and antidiffusive fluxes (:236-283), margin 5
            utotm = (st["u"][m_][k] + st["ubavg"][m_]) * scuy
            qu = np.where(utotm >= 0.0,
                          _fmin(_sh(d, -1, 0), _fmax(0.0, depthu - _sh(util3, -1, 0))),
                          _fmin(d, _fmax(0.0, depthu - util3)))
            fu = np.where(sea_u & R[5], utotm * qu, 0.0)
            fu2 = np.where(sea_u & R[5], utotm * st["dpu"][m_][k] - fu, 0.0)
            vtotm = (st["v"][m_][k] + st["vbavg"][m_]) * scvx
            qv = np.where(vtotm >= 0.0,
                          _fmin(_sh(d, 0, -1), _fmax(0.0, depthv - _sh(util3, 0, -1))),
                          _fmin(d, _fmax(0.0, depthv - util3)))
            fv = np.where(sea_v & R[5], vtotm * qv, 0.0)
            fv2 = np.where(sea_v & R[5], vtotm * st["dpv"][m_][k] - fv, 0.0)
            uflx_k = np.where(sea_u & R[5], fu, st["uflx"][k])
            vflx_k = np.where(sea_v & R[5], fv, st["vflx"][k])
            # low-order thickness (:293-311), margin 4
            r4 = sea_p & R[4]
            util3 = np.where(r4, util3 + d, util3)
            dlo = d - ((_sh(fu, 1, 0) - fu) + (_sh(fv, 0, 1) - fv)) * delt1 * scp2i
            d = np.where(r4, dlo, d)
            rows = np.where(r4, d, 999.0).min(axis=1)           # dpmn(j): all columns of the margin
            nb = geom.nbdy
            dpkmin[k] = min(999.0, rows[nb:nb + geom.jj].min())
            # ratios (:378-400), margin 4
            mx, mn = _extrema5(d, ip)
            mn = _fmax(0.0, mn)
            pos = lambda a: _fmax(0.0, a)   # noqa: E731
            neg = lambda a: _fmin(0.0, a)   # noqa: E731
            u1 = (mx - d) / (((pos(fu2) - neg(_sh(fu2, 1, 0))) + (pos(fv2) - neg(_sh(fv2, 0, 1))) + EPSIL) * delt1 * scp2i)
            u2 = (mn - d) / (((neg(fu2) - pos(_sh(fu2, 1, 0))) + (neg(fv2) - pos(_sh(fv2, 0, 1))) - EPSIL) * delt1 * scp2i)
            u1 = np.where(r4, u1, np.nan)
            u2 = np.where(r4, u2, np.nan)
            # limiter (:414-441), margin 3
            cu = np.where(fu2 >= 0.0, _fmin(_fmin(1.0, u1), _sh(u2, -1, 0)), _fmin(_fmin(1.0, u2), _sh(u1, -1, 0)))
            cv = np.where(fv2 >= 0.0, _fmin(_fmin(1.0, u1), _sh(u2, 0, -1)), _fmin(_fmin(1.0, u2), _sh(u1, 0, -1)))
            r3u, r3v = sea_u & R[3], sea_v & R[3]
            utotn = np.where(r3u, utotn + fu2 * (1.0 - cu), utotn)
            vtotn = np.where(r3v, vtotn + fv2 * (1.0 - cv), vtotn)
            fuc = np.where(r3u, fu2 * cu, fu)      # outside margin 3 uflux keeps the low-order value (never used)
            fvc = np.where(r3v, fv2 * cv, fv)
            uflx_k = np.where(r3u, uflx_k + fuc, uflx_k)
            vflx_k = np.where(r3v, vflx_k + fvc, vflx_k)
            # antidiffusive update (:449-469), margin 2
            r2 = sea_p & R[2]
            d = np.where(r2, d - ((_sh(fuc, 1, 0) - fuc) + (_sh(fvc, 0, 1) - fvc)) * delt1 * scp2i, d)
            p[k + 1] = np.where(r2, p[k] + d, p[k + 1])
            rows = np.where(r2, d, 999.0).min(axis=1)
            dpkmin[kk + k] = min(999.0, rows[nb:nb + geom.jj].min())
            dp[n_][k] = d
            st["uflx"][k], st["vflx"][k] = uflx_k, vflx_k
        # loop 77 (:588-683): the clipped fluxes go back in proportion to the layer thickness
        pb = p[kk].copy()
        for k in range(kk):
            d = dp[n_][k]
            r1u, r1v = sea_u & R[1], sea_v & R[1]
            qu = np.where(utotn >= 0.0, _sh(d, -1, 0) / _sh(pb, -1, 0), d / pb)
            qv = np.where(vtotn >= 0.0, _sh(d, 0, -1) / _sh(pb, 0, -1), d / pb)
            fu = np.where(r1u, utotn * qu, 0.0)
            fv = np.where(r1v, vtotn * qv, 0.0)
            st["uflx"][k] = np.where(r1u, st["uflx"][k] + fu, st["uflx"][k])
            st["vflx"][k] = np.where(r1v, st["vflx"][k] + fv, st["vflx"][k])
            r0 = sea_p & inner
            d = np.where(r0, d - ((_sh(fu, 1, 0) - fu) + (_sh(fv, 0, 1) - fv)) * delt1 * scp2i, d)
            p[k + 1] = np.where(r0, p[k] + d, p[k + 1])
            dpkmin[k] = min(999.0, np.where(r0, d, 999.0).min())
            dp[n_][k] = d
        # bottom-pressure restoring (:716-733)
        r0 = sea_p & inner
        q = pbot / p[kk]
        for k in range(kk):
            dp[n_][k] = np.where(r0, dp[n_][k] * q, dp[n_][k])
            p[k + 1] = np.where(r0, p[k] + dp[n_][k], p[k + 1])
        if isopyc:
            st["dpmixl"][n_] = np.where(r0, dp[n_][0], st["dpmixl"][n_])
        if thk is not None:
            cnuity_thkdf(geom, st, p, n, ip, iu, iv, thk["scp2"], scp2i, delt1, thk["thkdf4u"], thk["thkdf4v"],
                         thk["bih"], thk["nstep"], isopyc, thk["halo"])
        if mxlkta is not None:
            cnuity_mxlkta(geom, st, p, n, ip, iu, iv, scp2i, delt1, mxlkta["onemm"],
                          thk["thkdf4u"] if thk else None, thk["thkdf4v"] if thk else None, thk["bih"] if thk else True)
        # cumulative fluxes (:1326-1350)
        for k in range(kk):
            st["uflxav"][k] = np.where(sea_u & inner, st["uflxav"][k] + st["uflx"][k], st["uflxav"][k])
            st["vflxav"][k] = np.where(sea_v & inner, st["vflxav"][k] + st["vflx"][k], st["vflxav"][k])
            st["dpav"][k] = np.where(r0, st["dpav"][k] + dp[n_][k], st["dpav"][k])
    return p, utotn, vtotn, dpkmin, dpmold


def _extended_neighbours(ip):
    """array -> array at ipim1x / ipip1x / ipjm1x / ipjp1x (bigrid.F90:343-372): i-1 if sea; else i+1 if sea;
    otherwise i - and likewise for the other three"""
    w_sea, e_sea = _sh(ip, -1, 0) != 0, _sh(ip, 1, 0) != 0
    s_sea, n_sea = _sh(ip, 0, -1) != 0, _sh(ip, 0, 1) != 0

    def xa(a):
        return np.where(w_sea, _sh(a, -1, 0), np.where(e_sea, _sh(a, 1, 0), a))

    def xb(a):
        return np.where(e_sea, _sh(a, 1, 0), np.where(w_sea, _sh(a, -1, 0), a))

    def ya(a):
        return np.where(s_sea, _sh(a, 0, -1), np.where(n_sea, _sh(a, 0, 1), a))

    def yb(a):
        return np.where(n_sea, _sh(a, 0, 1), np.where(s_sea, _sh(a, 0, -1), a))
    return xa, xb, ya, yb


def cnuity_mxlkta(geom, st, p, n, ip, iu, iv, scp2i, delt1, onemm, thku=None, thkv=None, bih=True):
    """hybrid .and. mxlkta (cnuity.F90:1144-1324): dpmixl(:,:,n) follows the vertical excursion of the coordinates
    immediately above and below the mixed-layer base (found in the OLD thicknesses dpo(:,:,:,n)), then is
    diffused like an interface (biharmonic with thkdf4, Laplacian with thkdf2; thku None: no diffusion)"""
    kk = geom.kdm
    n_ = n - 1
    dpo = st["dpo"][n_]
    dpmx = st["dpmixl"][n_]
    sea_p, sea_u, sea_v = ip != 0, iu != 0, iv != 0
    r4 = sea_p & _region(geom, 4)
    shape = (geom.nrows, geom.ncols)
    above, below = np.zeros(shape), np.zeros(shape)
    with np.errstate(all="ignore"):
        for k in range(kk):
            above = below
            below = below + dpo[k]
            base_here = r4 & (below >= dpmx) & (above < dpmx)
            dpup = p[k] - above
            dpdn = p[k + 1] - below
            q = (below - dpmx) / np.maximum(onemm, dpo[k])
            dpmx = np.where(base_here, dpmx + (dpdn + q * (dpup - dpdn)), dpmx)
        if thku is not None:
            if bih:
                xa, xb, ya, yb = _extended_neighbours(ip)
                cell2 = sea_p & _region(geom, 2)
                t1 = np.where(cell2, dpmx - 0.5 * (xa(dpmx) + xb(dpmx)), 0.0)
                t2 = np.where(cell2, dpmx - 0.5 * (ya(dpmx) + yb(dpmx)), 0.0)
                fu = np.where(sea_u & _region(geom, 1), (delt1 * thku) * (_sh(t1, -1, 0) - t1), 0.0)
                fv = np.where(sea_v & _region(geom, 1), (delt1 * thkv) * (_sh(t2, 0, -1) - t2), 0.0)
            else:
                fu = np.where(sea_u & _region(geom, 2), (delt1 * thku) * (_sh(dpmx, -1, 0) - dpmx), 0.0)
                fv = np.where(sea_v & _region(geom, 2), (delt1 * thkv) * (_sh(dpmx, 0, -1) - dpmx), 0.0)
            cell0 = sea_p & _region(geom, 0)
            dpmx = np.where(cell0, dpmx - ((_sh(fu, 1, 0) - fu) + (_sh(fv, 0, 1) - fv)) * scp2i, dpmx)
    st["dpmixl"][n_] = dpmx


def cnuity_thkdf(geom, st, p, n, ip, iu, iv, scp2, scp2i, delt1, thku, thkv, bih, nstep, isopyc, halo):
    """biharmonic (:745-963) or Laplacian (:973-1124) thickness diffusion - literally, interface depth diffusion:
    the interfaces p(:,:,2..kk) are moved by fluxes limited so that interfaces do not intertwine; the biharmonic
    form walks through the interfaces downward or upward in alternate time steps and limits each flux against
    the one of the interface before.  Whole-array form; p, dp(:,:,:,n), uflx, vflx are updated in place."""
    kk = geom.kdm
    n_ = n - 1
    dp = st["dp"]
    onecm = 9806.0 * 0.01
    st["dpmixl"][n_] = halo(st["dpmixl"][n_], 1)
    dp[n_] = halo(dp[n_], 1)
    p[1:] = halo(p[1:], 1)
    dtinv = 1. / delt1
    iflip = nstep % 2
    sea_p, sea_u, sea_v = ip != 0, iu != 0, iv != 0
    r5, r4 = _region(geom, 5), _region(geom, 4)
    shape = (geom.nrows, geom.ncols)
    uflux, vflux = np.zeros(shape), np.zeros(shape)
    pold = p[kk].copy() if iflip == 1 else np.zeros(shape)
    xa, xb, ya, yb = _extended_neighbours(ip)

    order = range(2, kk + 1) if (not bih or iflip == 0) else range(kk, 1, -1)
    u1, u2 = np.zeros(shape), np.zeros(shape)
    with np.errstate(all="ignore"):
        for k in order:
            pk, pb = p[k - 1], p[kk]
            if bih:
                dk, dkm = dp[n_][k - 1], dp[n_][k - 2]
                a1 = pk - .5 * (xa(pk) + xb(pk))
                a2 = pk - .5 * (ya(pk) + yb(pk))
                a1 = np.where(a1 > 0.0, np.where(np.minimum(xa(dk), xb(dk)) < onecm, 0.0, a1),
                              np.where(np.minimum(xa(dkm), xb(dkm)) < onecm, 0.0, a1))
                a2 = np.where(a2 > 0.0, np.where(np.minimum(ya(dk), yb(dk)) < onecm, 0.0, a2),
                              np.where(np.minimum(ya(dkm), yb(dkm)) < onecm, 0.0, a2))
                thin = np.minimum(dk, dkm) < onecm
                cell5 = sea_p & r5
                u1 = np.where(cell5, np.where(thin, 0.0, a1), u1)
                u2 = np.where(cell5, np.where(thin, 0.0, a2), u2)
            for d, (flux, coef, util, sea_f, key) in enumerate(((uflux, thku, u1, sea_u, "uflx"),
                                                                (vflux, thkv, u2, sea_v, "vflx"))):
                di, dj = (-1, 0) if d == 0 else (0, -1)
                lo_ = lambda a: _sh(a, di, dj)   # noqa: E731  the cell on the low side of the face
                flxhi = .25 * (pb - pk) * scp2
                flxlo = -.25 * (lo_(pb) - lo_(pk)) * lo_(scp2)
                if bih:
                    if iflip == 0:   # downward k loop
                        flxhi = np.minimum(flxhi, flux + .25 * (lo_(pk) - lo_(pold)) * lo_(scp2))
                        flxlo = np.maximum(flxlo, flux - .25 * (pk - pold) * scp2)
                    else:            # upward k loop
                        flxhi = np.minimum(flxhi, flux + .25 * (pold - pk) * scp2)
                        flxlo = np.maximum(flxlo, flux - .25 * (lo_(pold) - lo_(pk)) * lo_(scp2))
                    want = (delt1 * coef) * (lo_(util) - util)
                else:
                    want = (delt1 * coef) * (lo_(pk) - pk)
                face = sea_f & r4
                new = np.where(face, np.minimum(flxhi, np.maximum(flxlo, want)), flux)
                flux[...] = new
                st[key][k - 2] = np.where(face, st[key][k - 2] + new * dtinv, st[key][k - 2])
                st[key][k - 1] = np.where(face, st[key][k - 1] - new * dtinv, st[key][k - 1])
            cell = sea_p & r4
            pold = np.where(cell, pk, pold)
            p[k - 1] = np.where(cell, pk - ((_sh(uflux, 1, 0) - uflux) + (_sh(vflux, 0, 1) - vflux)) * scp2i, pk)
        cell = sea_p & r4
        for k in range(1, kk + 1):
            p[k] = np.where(cell & (p[k] < p[k - 1]), p[k - 1], p[k])
            dp[n_][k - 1] = np.where(cell, p[k] - p[k - 1], dp[n_][k - 1])
        if isopyc:
            st["dpmixl"][n_] = np.where(cell, dp[n_][0], st["dpmixl"][n_])


def cnuity_asselin(geom, st, m, n, ip, ra2fac):
    """the Robert-Asselin tail of cnuity (:1396-1422); dp(:,:,:,n) halo valid to width 6"""
    n_, m_ = n - 1, m - 1
    r6 = (ip != 0) & _region(geom, 6)
    dp, dpo = st["dp"], st["dpo"]
    with np.errstate(all="ignore"):
        for k in range(geom.kdm):
            q = 0.5 * ra2fac * (dpo[n_][k] + dp[n_][k] - 2.0 * dp[m_][k])
            mid = dp[m_][k].copy()
            dpo[m_][k] = np.where(r6, mid, dpo[m_][k])
            dp[m_][k] = np.where(r6, mid + q, dp[m_][k])
