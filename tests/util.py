"""Shared helpers of the test-suite (TEST INFRASTRUCTURE).

`make_case` builds a synthetic single-tile CbArrays (product-side host generator,
identical bits on host and device), `run_oracle` runs the CPU oracle on a copy of it.
"""
from __future__ import annotations

import importlib

import numpy as np

pkg = importlib.import_module("hycom-src_b200")
syn = pkg.synthetic
cabi = importlib.import_module("hycom-src_b200.cabi")

REL_TOL = 1.0e-12  # BASELINE.json north_star: "within 1e-12 relative max error"


def make_cfg(itdm, jtdm, kdm, nreg=0, ntracr=0, seed=1, dx0=20000.0, delt1=3600.0):
    return syn.make_cfg(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=seed, dx0=dx0, delt1=delt1)


def make_case(itdm, jtdm, kdm, nreg=0, ntracr=0, seed=1, m=1, n=2, **scalars):
    cfg = make_cfg(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=seed)
    sea = syn.sea_mask(cfg)
    g = pkg.partition(itdm, jtdm, kdm, 1, 1, nreg)[0]
    cb = syn.build_cb_arrays(cfg, g, sea, m, n, **scalars)
    return cfg, sea, g, cb


def oracle_tile_from_cb(oracle, cb, sea=None):
    """an oracle tile loaded with the product-side CbArrays; masks come from the
    oracle's own bigrid restatement when `sea` is given (checked against cb)"""
    g = cb.geom
    ot = oracle.tile(g, cb.ntracr)
    if sea is not None:
        depth = np.zeros((g.nrows, g.ncols))
        nb = g.nbdy
        depth[nb:nb + g.jj, nb:nb + g.ii] = np.where(
            sea[g.j0:g.j0 + g.jj, g.i0:g.i0 + g.ii] != 0, 100.0, 0.0)
        ot.bigrid(depth)
    else:
        raise ValueError("sea map required")
    ot.load_cb(cb)
    return ot


def add_q2(cfg, sea, g, cb, m, n):
    """Mellor-Yamada fields q2, q2l (layers 0..kk+1, both slots) for mxlmy cases: tracer-like
    synthetic fields, NaN halos like every array tsadvc exchanges"""
    def lev_of(slot):
        return 0 if slot == n else 1
    for name, ktr in (("q2", 31), ("q2l", 32)):
        a = np.empty((2, g.kdm + 2, g.nrows, g.ncols))
        for slot in (1, 2):
            a[slot - 1] = syn.fill_host(cfg, g, sea, cabi.F_TRACER, ktr, lev_of(slot), 1, g.kdm + 2, 0)
        setattr(cb, name, a)
    cb.mxlmy = True
    return cb


def run_oracle(oracle, cb, sea, m, n):
    """CPU oracle tsadvc(m,n) on a private copy of cb; returns dict of slot-n results"""
    ot = oracle_tile_from_cb(oracle, cb, sea)
    ot.tsadvc(m, n, 1)
    out = dict(temp=ot.f64("temp").copy(), saln=ot.f64("saln").copy(), th3d=ot.f64("th3d").copy(),
               xmin=ot.f64("xmin").copy(), xmax=ot.f64("xmax").copy())
    if cb.mxlmy:
        out["q2"], out["q2l"] = ot.f64("q2").copy(), ot.f64("q2l").copy()
    if cb.ntracr:
        out["tracer"] = ot.f64("tracer").copy()
    ot.close()
    return out


def rel_err(a, b, mask):
    """max_sea|a-b| / max_sea|b| over mask"""
    d = np.abs(np.where(mask, a - b, 0.0))
    ref = np.abs(np.where(mask, b, 0.0))
    den = ref.max()
    return 0.0 if den == 0 else float(d.max() / den)


def interior_sea(cb):
    g = cb.geom
    nb = g.nbdy
    m = np.zeros((g.nrows, g.ncols), dtype=bool)
    m[nb:nb + g.jj, nb:nb + g.ii] = cb.ip[nb:nb + g.jj, nb:nb + g.ii] != 0
    return m


def make_diffusion_case(itdm, jtdm, kdm, sigver, temdfc, nreg=0, ntracr=0, nhybrd=-1, seed=3, temdf2=None,
                        **scalars):
    """a case with temdf2 > 0 (mod_tsadvc.F90:2138-2230): th3d is made consistent with the
    equation of state (sig(T,S)-thbase plus a smooth offset, so that tofsig has a root to
    find) and theta holds per-layer target densities"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import np_restatement as npr
    thbase = 34.0 if sigver % 2 == 0 else 25.0
    if temdf2 is None:
        temdf2 = 0.02          # m/s diffusion velocity (blkdat temdf2 is O(0.005..0.02))
    cfg, sea, g, cb = make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=seed, temdf2=temdf2,
                                temdfc=temdfc, sigver=sigver, thbase=thbase, nhybrd=nhybrd, **scalars)
    sv = sigver if sigver not in (5, 6) else sigver - 4   # the 17-term fit has no tofsig: seed from the 7-term
    with np.errstate(all="ignore"):
        cb.th3d = np.ascontiguousarray(npr.sig(sv, cb.temp, cb.saln) - thbase + 0.01 * np.cos(cb.temp))
    cb.th3d[~np.isfinite(cb.th3d)] = np.nan
    cb.theta = np.empty((kdm, g.nrows, g.ncols))
    for k in range(kdm):
        cb.theta[k] = np.nanmean(cb.th3d[:, k]) + 0.001 * k
    return cfg, sea, g, cb


def make_arctic_case(itdm, jtdm, kdm, ntracr=0, seed=1, m=1, n=2, **scalars):
    """nreg=2: a global grid across the arctic on one tile (periodic in i, tripole fold at the
    top).  The generator builds a periodic/closed basin; the top rows are then opened (all sea,
    hence fold-consistent) and every array that must arrive with a valid halo (dp, oneta,
    metrics) gets the arctic halo of its grid."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import np_restatement as npr
    cfg = make_cfg(itdm, jtdm, kdm, nreg=2, ntracr=ntracr, seed=seed)
    sea = syn.sea_mask(cfg)
    sea[jtdm - 8:, :] = 1
    g = pkg.partition(itdm, jtdm, kdm, 1, 1, 2)[0]
    cb = syn.build_cb_arrays(cfg, g, sea, m, n, **scalars)
    nb = g.nbdy
    cb.dp = npr.halo_single_tile(g, cb.dp, nb, nb, 1)
    cb.oneta = npr.halo_single_tile(g, cb.oneta, nb, nb, 1)
    for name, it in (("scp2", 1), ("scp2i", 1), ("scuy", 3), ("aspux", 3), ("scvx", 4), ("aspvy", 4)):
        setattr(cb, name, np.ascontiguousarray(npr.halo_single_tile(g, getattr(cb, name), nb, nb, it)))
    return cfg, sea, g, cb


def make_arctic_tiles(cfg, sea, cb1, ipr, jpr, m=1, n=2, **scalars):
    """the tiles of an arctic case: every tile is generated from the global function like any other
    tiling; the arrays that must arrive with a valid halo (dp, oneta, metrics) are windows of the
    single-tile arrays, whose halo make_arctic_case filled with the fold of their grid"""
    g1 = cb1.geom
    cbs = []
    for g in pkg.partition(g1.itdm, g1.jtdm, g1.kdm, ipr, jpr, 2):
        cb = syn.build_cb_arrays(cfg, g, sea, m, n, **scalars)
        win = (slice(g.j0, g.j0 + g.nrows), slice(g.i0, g.i0 + g.ncols))
        for name in ("dp", "oneta", "scp2", "scp2i", "scuy", "aspux", "scvx", "aspvy"):
            src = getattr(cb1, name)
            setattr(cb, name, np.ascontiguousarray(src[(Ellipsis,) + win]))
        cbs.append(cb)
    return cbs


def add_asselin(cfg, sea, g, cb, m, n, sigver=6, seed=77):
    """operands of mod_asselin.F90 on top of a case: dpo both slots (a perturbed dp), pbot/pbavg so
    that oneta = 1 + O(1e-3), the saved t-1 fields o* (a perturbed copy of slot n), dp of both slots"""
    rng = np.random.default_rng(seed)
    shp = (g.nrows, g.ncols)
    kk = g.kdm
    cb.sigver = sigver
    cb.thbase = 34.0 if sigver % 2 == 0 else 25.0
    cb.dpo = cb.dp * (1.0 + 0.02 * rng.standard_normal(cb.dp.shape))
    cb.dpo[np.abs(cb.dp) == 0.0] = 0.0                     # massless layers stay massless at t-1 and t
    cb.pbot = np.full(shp, 4.0e7) * (1.0 + 0.1 * rng.random(shp))
    cb.pbavg = 4.0e4 * rng.standard_normal((3,) + shp)
    cb.onetao = 1.0 + 1.0e-3 * rng.standard_normal((2,) + shp)
    for name, src in (("otemp", cb.temp), ("osaln", cb.saln), ("oth3d", cb.th3d)):
        setattr(cb, name, src[n - 1] * (1.0 + 1.0e-3 * rng.standard_normal(src[n - 1].shape)))
    if cb.ntracr:
        cb.otracer = cb.tracer[:, n - 1] * (1.0 + 1.0e-3 * rng.standard_normal(cb.tracer[:, n - 1].shape))
    if cb.mxlmy:
        cb.oq2 = cb.q2[n - 1] * (1.0 + 1.0e-3 * rng.standard_normal(cb.q2[n - 1].shape))
        cb.oq2l = cb.q2l[n - 1] * (1.0 + 1.0e-3 * rng.standard_normal(cb.q2l[n - 1].shape))
    if cb.theta is None:
        cb.theta = np.empty((kk, g.nrows, g.ncols))
        for k in range(kk):
            cb.theta[k] = 1.0 + 0.1 * k
    return cb


def add_cnuity(cfg, sea, g, cb, m, n, seed=91, uscale=None, thkdf=0.0, bih=True):
    """operands of cnuity(m,n) (cnuity.F90) on top of a case, as a dict of arrays in the Fortran layout.
    Halos of the arrays cnuity exchanges itself (:100-107) are NaN, like every exchanged array; pbot, depthu,
    depthv arrive with valid halos.  u, v are O(0.3 m/s) smooth fields (both signs), zero off the iu / iv
    points; dp(:,:,:,m) is a perturbed dp(:,:,:,n); pbot is the column sum of dp(m); depthu/depthv the
    shallower neighbour; dpu/dpv the thickness of layer k on the faces (dpudpv.F90's definition)."""
    rng = np.random.default_rng(seed)
    kk, shp = g.kdm, (g.nrows, g.ncols)
    nb = g.nbdy
    full = lambda fld, ktr=0, lev=0: syn.fill_host(cfg, g, sea, fld, ktr, lev, 1, kk, 1)   # noqa: E731
    dpn = syn.fill_host(cfg, g, sea, cabi.F_DP, 0, 0, 1, kk, 1)
    dpm = syn.fill_host(cfg, g, sea, cabi.F_DP, 0, 1, 1, kk, 1)
    sea_h = np.zeros(shp, dtype=bool)           # sea map on the padded tile (periodic image or land)
    js = (np.arange(g.nrows) - nb + g.j0) % g.jtdm if g.periodic_j else np.arange(g.nrows) - nb + g.j0
    is_ = (np.arange(g.ncols) - nb + g.i0) % g.itdm if g.periodic_i else np.arange(g.ncols) - nb + g.i0
    okj, oki = (js >= 0) & (js < g.jtdm), (is_ >= 0) & (is_ < g.itdm)
    sea_h[np.ix_(okj, oki)] = sea[np.ix_(js[okj], is_[oki])] != 0
    dpn = np.where(sea_h, dpn, 0.0)
    dpm = np.where(sea_h, dpm, 0.0)
    pbot = dpm.sum(axis=0)
    pbot = np.where(sea_h, np.maximum(pbot, 9806.0), 0.0)
    west = np.roll(pbot, 1, axis=1); south = np.roll(pbot, 1, axis=0)
    depthu, depthv = np.minimum(pbot, west), np.minimum(pbot, south)
    pint = np.concatenate([np.zeros((1,) + shp), np.cumsum(dpm, axis=0)])       # interfaces of dp(m)
    def face_thk(shift_axis, depth):
        lo = 0.5 * (pint + np.roll(pint, 1, axis=shift_axis))
        return np.maximum(0.0, np.minimum(depth, lo[1:]) - np.minimum(depth, lo[:-1]))
    dpu, dpv = face_thk(2, depthu), face_thk(1, depthv)
    uf = syn.fill_host(cfg, g, sea, cabi.F_UFLX, 0, 0, 1, kk, 1)
    vf = syn.fill_host(cfg, g, sea, cabi.F_VFLX, 0, 0, 1, kk, 1)
    if uscale is None:      # tiles of one case pass the single-tile value: the fields are functions of (i,j,k) only
        uscale = (0.3 / max(np.abs(uf).max(), 1e-30), 0.3 / max(np.abs(vf).max(), 1e-30))
    u = uscale[0] * uf
    v = uscale[1] * vf
    u = np.where(cb.iu != 0, u, 0.0); v = np.where(cb.iv != 0, v, 0.0)
    ub = 0.1 * u.mean(axis=0); vb = 0.1 * v.mean(axis=0)
    nanhalo = np.ones(shp, dtype=bool)
    nanhalo[nb:nb + g.jj, nb:nb + g.ii] = False

    def halo_nan(a):
        a = np.array(a, dtype=np.float64)
        a[..., nanhalo] = np.nan
        return np.ascontiguousarray(a)
    st = dict(
        dp=halo_nan(np.stack([dpm, dpn] if n == 2 else [dpn, dpm])),
        dpo=np.full((2, kk) + shp, np.nan),
        u=halo_nan(np.stack([u, u * 0.9]) if m == 1 else np.stack([u * 0.9, u])),
        v=halo_nan(np.stack([v, v * 0.9]) if m == 1 else np.stack([v * 0.9, v])),
        dpu=halo_nan(np.stack([dpu, dpu]) ), dpv=halo_nan(np.stack([dpv, dpv])),
        ubavg=halo_nan(np.stack([ub, ub, ub])), vbavg=halo_nan(np.stack([vb, vb, vb])),
        dpmixl=halo_nan(np.stack([dpn[0] * 0.5, dpn[0] * 0.6])),
        uflx=np.full((kk,) + shp, np.nan), vflx=np.full((kk,) + shp, np.nan),
        uflxav=np.zeros((kk,) + shp), vflxav=np.zeros((kk,) + shp), dpav=np.zeros((kk,) + shp),
        pbot=np.ascontiguousarray(pbot), depthu=np.ascontiguousarray(depthu), depthv=np.ascontiguousarray(depthv))
    st["_uscale"] = uscale
    if thkdf:
        # forfun.F90:2541-2568: thkdf4u = thkdf4*aspux**3*scuy (biharmonic) or thkdf2*aspux*scuy (Laplacian) at
        # the u points, 0.0 elsewhere (aspux = aspvy = 1 on the synthetic grids); a smooth variation on top
        # (a function of the global cell, wrapped like the grid, so that every tiling sees the same halos)
        ig = is_[None, :] + 0.0 * js[:, None]
        jg = js[:, None] + 0.0 * is_[None, :]
        var = 1.0 + 0.2 * np.sin(2.0 * np.pi * 3.0 * ig / g.itdm) * np.cos(2.0 * np.pi * 2.0 * jg / g.jtdm)
        st["thkdf4u"] = np.ascontiguousarray(np.where(cb.iu != 0, thkdf * var * cb.scuy, 0.0))
        st["thkdf4v"] = np.ascontiguousarray(np.where(cb.iv != 0, thkdf * var * cb.scvx, 0.0))
        st["_thkdf"] = (thkdf, bih)
    # geopar.F90:822-871: uflx, vflx are zero on the land faces that bound sea segments
    st["uflx"][:, cb.iu == 0] = 0.0
    st["vflx"][:, cb.iv == 0] = 0.0
    return st


def arctic_halos_cnuity(g, st):
    """nreg=2: the cnuity operands that must arrive with a valid halo get the tripole fold of their grid"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import np_restatement as npr
    nb = g.nbdy
    for name, it in (("pbot", 1), ("depthu", 3), ("depthv", 4), ("thkdf4u", 3), ("thkdf4v", 4)):
        if name in st:
            st[name] = np.ascontiguousarray(npr.halo_single_tile(g, st[name], nb, nb, it))
    return st


def deepen_dpmixl(st, n):
    """a mixed-layer base that wanders through the upper layers (add_cnuity keeps it inside layer 1): 30 % .. 230 %
    of the first layer, as a function of the thickness fields (so every tiling sees the same values)"""
    d = st["dp"]
    for t in (0, 1):
        frac = 0.3 + 2.0 * (np.nan_to_num(d[t][0]) % 7.0) / 7.0
        st["dpmixl"][t] = np.where(np.isnan(st["dpmixl"][t]), np.nan, frac * np.nan_to_num(d[t][0]) + 1.0)
    return st


def oracle_load_cnuity(ot, st):
    ot.cnuity_alloc()
    for name in ("dp", "dpo", "u", "v", "dpu", "dpv", "ubavg", "vbavg", "dpmixl", "uflx", "vflx", "uflxav", "vflxav",
                 "dpav", "pbot", "depthu", "depthv"):
        ot.f64(name)[...] = st[name]
    if "thkdf4u" in st:
        ot.f64("thkdf4u")[...] = st["thkdf4u"]
        ot.f64("thkdf4v")[...] = st["thkdf4v"]
        ot.set_d("thkdf4" if st["_thkdf"][1] else "thkdf2", st["_thkdf"][0])


def run_oracle_cnuity(oracle, cb, sea, st, m, n, isopyc=False, mxlkta=False):
    """CPU oracle cnuity(m,n) on a private copy; returns the arrays it updates"""
    ot = oracle_tile_from_cb(oracle, cb, sea)
    oracle_load_cnuity(ot, st)
    ot.set_i("isopyc", int(isopyc))
    ot.set_i("mxlkta", int(mxlkta))
    ot.cnuity(m, n, 1)
    out = {k: ot.f64(k).copy() for k in ("dp", "dpo", "uflx", "vflx", "p", "utotn", "vtotn", "dpkmin", "dpmixl",
                                         "uflxav", "vflxav", "dpav", "dpmold")}
    ot.close()
    return out


def run_compiled_reference_text(cb, sea, m, n, sigver=6):
    """tsadvc(m,n) of the COMPILED REFERENCE TEXT (oracle/_ref/libref_text_*.so: mod_tsadvc.F90, bigrid.F90 and xctilr of
    mod_xc_sm.h translated statement by statement to C by oracle/fortran_to_c.py where /root/reference exists) on the
    host arrays of cb, in place.  False when no library can be had on this machine (then the caller has the oracle only)."""
    import os, sys
    root = os.path.join(os.path.dirname(__file__), "..")
    sys.path.insert(0, os.path.join(root, "oracle"))
    import reference_text as rt
    import reference_text_c as rc
    g = cb.geom
    arctic = g.nreg == 2
    so = rc.RefTextC.so_path(sigver, arctic)
    if not rt.available() and not (os.path.exists(so) and os.path.exists(so[:-3] + ".json")):
        return False
    try:
        lib = rc.RefTextC(sigver, arctic)
    except (OSError, RuntimeError) as e:
        print("compiled reference text unavailable:", repr(e)[:200])
        return False
    nb = g.nbdy
    env = rt.make_env(g.ii, g.jj, g.kdm)
    depth = np.zeros((g.nrows, g.ncols))
    depth[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea != 0, 100.0, 0.0)
    lib.run(env, "bigrid", depth, 4 if g.nreg in (3, 4) else 0, *[np.zeros_like(depth) for _ in range(3)])
    assert env["nreg"] == g.nreg, (env["nreg"], g.nreg)
    rt.add_cb_arrays(env, cb)
    lib.run(env, "tsadvc", m, n)
    return True
