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


def prolog(geom, uflx_k, vflx_k, dp_kn, onetamas_m, delt1, scp2i, ip, margin):
    """tsadvc prolog, mod_tsadvc.F90:1905-1942: util1 = fco, util2 = fcn"""
    with np.errstate(all="ignore"):
        reg = _region(geom, margin) & (ip != 0)
        flxdiv = ((_sh(uflx_k, 1, 0) - uflx_k) + (_sh(vflx_k, 0, 1) - vflx_k)) * delt1 * scp2i
        fco = np.where(reg, _fmax(onetamas_m * dp_kn + flxdiv, 0.0), np.nan)
        fcn = np.where(reg, _fmax(onetamas_m * dp_kn, 0.0), np.nan)
        return fco, fcn


def halo_single_tile(geom, a, mh, nh):
    """xctilr on one tile, mod_xc_sm.h:1337-1428, for an array (..., nrows, ncols)"""
    nb, ii, jj = geom.nbdy, geom.ii, geom.jj
    a = a.copy()
    i1, j1 = nb, nb   # numpy index of i=1 / j=1
    for j in range(1, nh + 1):
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
    """tsadvc(m,n), hybrid coordinates, no diffusion (mod_tsadvc.F90:1804-2086).
    `cb` is a product-side CbArrays of host numpy arrays; returns new slot-n fields."""
    g = cb.geom
    kk = g.kdm
    mbdy = {0: 2, 1: 5, 2: 5, 4: 5}[abs(cb.advtyp)]
    nhyb = kk if cb.nhybrd < 0 else cb.nhybrd
    temp = halo_single_tile(g, cb.temp, mbdy, mbdy)
    saln = halo_single_tile(g, cb.saln, mbdy, mbdy)
    th3d = halo_single_tile(g, cb.th3d, mbdy, mbdy)
    uflx = halo_single_tile(g, cb.uflx, mbdy, mbdy)
    vflx = halo_single_tile(g, cb.vflx, mbdy, mbdy)
    tracer = halo_single_tile(g, cb.tracer, mbdy, mbdy) if cb.ntracr else None
    oem = np.ones((g.nrows, g.ncols))  # onetamas(:,:,m) = 1.0, :1809
    ip, iu, iv = cb.ip, cb.iu, cb.iv

    def adv(fld_n, fld_m, k, posdef, fco, fcn):
        a = (g, fld_n, fld_m, uflx[k], vflx[k], fco, fcn, cb.scp2, cb.scp2i, cb.delt1, ip, iu, iv)
        if cb.advtyp == 2:
            return advem_fct(a[0], 2, *a[1:])[0]
        if cb.advtyp == 4:
            return advem_fct(a[0], 4, *a[1:])[0]
        if cb.advtyp == 1:
            return advem_mpdata(g, fld_n, uflx[k], vflx[k], fco, fcn, posdef, cb.scp2, cb.scp2i,
                                cb.delt1, ip, iu, iv)[0]
        if cb.advtyp == 0:
            return advem_pcm(g, fld_n, uflx[k], vflx[k], fco, fcn, cb.scp2, cb.scp2i, cb.delt1,
                             ip, iu, iv)[0]
        raise ValueError(cb.advtyp)

    for k in range(kk):
        fco, fcn = prolog(g, uflx[k], vflx[k], cb.dp[n - 1, k], oem, cb.delt1, cb.scp2i, ip, mbdy - 1)
        if k + 1 <= nhyb:
            if cb.advflg == 0:
                temp[n - 1, k] = adv(temp[n - 1, k], temp[m - 1, k], k, 256.0, fco, fcn)
            else:
                th3d[n - 1, k] = adv(th3d[n - 1, k], th3d[m - 1, k], k, 32.0, fco, fcn)
        saln[n - 1, k] = adv(saln[n - 1, k], saln[m - 1, k], k, 0.0, fco, fcn)
        for q in range(cb.ntracr):
            pd = 256.0 if (q < len(cb.trcflg) and cb.trcflg[q] == 2) else 0.0
            tracer[q, n - 1, k] = adv(tracer[q, n - 1, k], tracer[q, m - 1, k], k, pd, fco, fcn)
    return dict(temp=temp, saln=saln, th3d=th3d, tracer=tracer)
