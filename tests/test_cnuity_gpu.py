"""cnuity(m,n) (cnuity.F90) on the device mirrors through the C ABI against the CPU
oracle, bit for bit: one tile, and ipr x jpr tiles through the library's communicator (the eight xctilr calls
of :100-107 and the one of :1400 inside the call)."""
import numpy as np
import pytest

import util
from util import pkg, syn, cabi

pytestmark = pytest.mark.gpu


def _check(ts, g, cb, ref, m, n, kdm, g1=None, isopyc=False, dpmixl=False):
    """device mirrors of one tile against the (single-tile) oracle result"""
    nb = g.nbdy
    g1 = g1 or g
    win = (slice(g.j0, g.j0 + g.nrows), slice(g.i0, g.i0 + g.ncols)) if g1 is not g else (slice(None), slice(None))
    inner = util.interior_sea(cb)
    iu_in = np.zeros_like(inner); iv_in = np.zeros_like(inner)
    iu_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iu[nb:nb + g.jj, nb:nb + g.ii] != 0
    iv_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iv[nb:nb + g.jj, nb:nb + g.ii] != 0

    def W(a):   # the tile's window of a single-tile array (..., nrows, ncols)
        out = np.full(a.shape[:-2] + (g.nrows, g.ncols), np.nan)
        src = a[(Ellipsis,) + win]
        out[..., :src.shape[-2], :src.shape[-1]] = src
        return out
    dpn, dpm = ts.download(cabi.F_DP, n), ts.download(cabi.F_DP, m)
    dpom = ts.download(cabi.F_DPO, m)
    uflx, vflx = ts.download(cabi.F_UFLX, 1), ts.download(cabi.F_VFLX, 1)
    p = ts.download(cabi.F_P, 1)
    for k in range(kdm):
        assert np.array_equal(dpn[k][inner], W(ref["dp"][n - 1, k])[inner]), ("dp.n", k)
        assert np.array_equal(dpm[k][inner], W(ref["dp"][m - 1, k])[inner]), ("dp.m", k)
        assert np.array_equal(dpom[k][inner], W(ref["dpo"][m - 1, k])[inner]), ("dpo.m", k)
        assert np.array_equal(uflx[k][iu_in], W(ref["uflx"][k])[iu_in]), ("uflx", k)
        assert np.array_equal(vflx[k][iv_in], W(ref["vflx"][k])[iv_in]), ("vflx", k)
        assert np.array_equal(p[k + 1][inner], W(ref["p"][k + 1])[inner]), ("p", k)
    assert np.array_equal(ts.download(cabi.F_UTOTN, 1)[0][iu_in], W(ref["utotn"])[iu_in])
    assert np.array_equal(ts.download(cabi.F_VTOTN, 1)[0][iv_in], W(ref["vtotn"])[iv_in])
    assert np.array_equal(ts.download(cabi.F_DPAV, 1)[0][inner], W(ref["dpav"][0])[inner])
    assert np.array_equal(ts.download(cabi.F_UFLXAV, 1)[kdm - 1][iu_in], W(ref["uflxav"][kdm - 1])[iu_in])
    if isopyc or dpmixl:
        assert np.array_equal(ts.download(cabi.F_DPMIXL, n)[0][inner], W(ref["dpmixl"][n - 1])[inner])


@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,m,n,isopyc", [
    (150, 150, 6, 0, 1, 2, False),     # BASELINE configs[0] basin
    (131, 77, 3, 3, 2, 1, False),      # doubly periodic, slots swapped
    (64, 90, 3, 1, 1, 2, True),        # periodic in i; isopyc: dpmixl(n) = dp(1,n)
    (70, 45, 2, 4, 1, 2, False),
])
def test_cnuity_device_matches_oracle(oracle, itdm, jtdm, kdm, nreg, m, n, isopyc):
    extra = dict(isopyc=True, hybrid=False, nhybrd=0) if isopyc else {}
    cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=23, m=m, n=n, nstep=3, **extra)
    st = util.add_cnuity(cfg, sea, g, cb, m, n)
    ref = util.run_oracle_cnuity(oracle, cb, sea, st, m, n, isopyc=isopyc)
    ts = pkg.Tsadvc(cb)
    ts.upload_cnuity_state(st, m, n)
    l0 = ts.launch_count
    dpkmin = ts.cnuity_device(m, n)
    assert ts.launch_count - l0 >= 4
    _check(ts, g, cb, ref, m, n, kdm, isopyc=isopyc)
    assert np.array_equal(dpkmin, ref["dpkmin"])
    ts.close()


def test_cnuity_refuses_what_is_not_built():
    cfg, sea, g, cb = util.make_case(40, 30, 2, seed=3)
    st = util.add_cnuity(cfg, sea, g, cb, 1, 2)
    ts = pkg.Tsadvc(cb)
    ts.upload_cnuity_state(st, 1, 2)
    # the interface-depth diffusion needs its coefficients, and only one of thkdf2 / thkdf4 (cnuity.F90:758)
    for kw in (dict(thkdf4=0.01), dict(thkdf2=0.01), dict(thkdf2=0.01, thkdf4=0.01)):
        with pytest.raises(cabi.TsadvcError) as e:
            ts.cnuity_device(1, 2, **kw)
        assert e.value.code == cabi.EINVAL
    cb.btrmas = True
    with pytest.raises(cabi.TsadvcError) as e:
        ts.cnuity_device(1, 2)
    assert e.value.code == cabi.EUNSUPPORTED
    ts.close()


@pytest.mark.parametrize("ipr,jpr,nreg", [(2, 2, 0), (2, 1, 3), (4, 2, 0)])
def test_cnuity_on_tiles(oracle, ipr, jpr, nreg):
    """ipr x jpr tiles through the library's communicator == the oracle on one tile"""
    from test_comm_gpu import make_tiles, run_tiles, close_tiles
    m, n = 1, 2
    itdm, jtdm, kdm = 160, 120, 3
    cfg, sea, g1, cb1 = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=23, m=m, n=n, nstep=3)
    st1 = util.add_cnuity(cfg, sea, g1, cb1, m, n)
    ref = util.run_oracle_cnuity(oracle, cb1, sea, st1, m, n)
    cbs = [syn.build_cb_arrays(cfg, g, sea, m, n, nstep=3) for g in pkg.partition(itdm, jtdm, kdm, ipr, jpr, nreg)]
    sts = [util.add_cnuity(cfg, sea, cb.geom, cb, m, n, uscale=st1["_uscale"]) for cb in cbs]
    grp, tss = make_tiles(cfg, sea, itdm, jtdm, kdm, ipr, jpr, nreg, m, n, cbs=cbs)

    def go(ts, r):
        ts.upload_cnuity_state(sts[r], m, n)
        ts.cnuity_device(m, n)
    run_tiles(tss, go)
    for ts, cb in zip(tss, cbs):
        _check(ts, cb.geom, cb, ref, m, n, kdm, g1=g1)
    close_tiles(grp, tss)


# interface-depth diffusion inside cnuity (cnuity.F90:745-1124): biharmonic in both sweep directions (nstep even:
# downward, odd: upward) and Laplacian
@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,bih,nstep,isopyc", [
    (150, 150, 6, 0, True, 4, False),
    (150, 150, 6, 0, True, 7, False),
    (64, 90, 4, 1, True, 3, True),
    (131, 77, 3, 3, False, 2, False),
    (70, 45, 4, 4, False, 5, False),
])
def test_cnuity_thickness_diffusion_matches_oracle(oracle, itdm, jtdm, kdm, nreg, bih, nstep, isopyc):
    m, n = 1, 2
    extra = dict(isopyc=True, hybrid=False, nhybrd=0) if isopyc else {}
    cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=29, m=m, n=n, nstep=nstep, **extra)
    thk = 0.01 if bih else 0.02
    st = util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=thk, bih=bih)
    ref = util.run_oracle_cnuity(oracle, cb, sea, st, m, n, isopyc=isopyc)
    ts = pkg.Tsadvc(cb)
    ts.upload_cnuity_state(st, m, n)
    l0 = ts.launch_count
    ts.cnuity_device(m, n, **({"thkdf4": thk} if bih else {"thkdf2": thk}))
    assert ts.launch_count - l0 >= 4 + (3 if bih else 2) * (kdm - 1)
    _check(ts, g, cb, ref, m, n, kdm, isopyc=isopyc)
    ts.close()


@pytest.mark.parametrize("ipr,jpr,nreg,bih,nstep", [(2, 2, 0, True, 5), (2, 1, 3, True, 2), (4, 2, 0, False, 3)])
def test_cnuity_thickness_diffusion_on_tiles(oracle, ipr, jpr, nreg, bih, nstep):
    from test_comm_gpu import make_tiles, run_tiles, close_tiles
    m, n = 1, 2
    itdm, jtdm, kdm = 160, 120, 4
    thk = 0.01 if bih else 0.02
    cfg, sea, g1, cb1 = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=23, m=m, n=n, nstep=nstep)
    st1 = util.add_cnuity(cfg, sea, g1, cb1, m, n, thkdf=thk, bih=bih)
    ref = util.run_oracle_cnuity(oracle, cb1, sea, st1, m, n)
    cbs = [syn.build_cb_arrays(cfg, g, sea, m, n, nstep=nstep) for g in pkg.partition(itdm, jtdm, kdm, ipr, jpr, nreg)]
    sts = [util.add_cnuity(cfg, sea, cb.geom, cb, m, n, uscale=st1["_uscale"], thkdf=thk, bih=bih) for cb in cbs]
    grp, tss = make_tiles(cfg, sea, itdm, jtdm, kdm, ipr, jpr, nreg, m, n, cbs=cbs)

    def go(ts, r):
        ts.upload_cnuity_state(sts[r], m, n)
        ts.cnuity_device(m, n, **({"thkdf4": thk} if bih else {"thkdf2": thk}))
    run_tiles(tss, go)
    for ts, cb in zip(tss, cbs):
        _check(ts, cb.geom, cb, ref, m, n, kdm, g1=g1)
    close_tiles(grp, tss)


# hybrid .and. mxlkta (cnuity.F90:1144-1324): dpmixl follows the coordinates around the mixed-layer base and is
# diffused like an interface
@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,mode,nstep", [
    (150, 150, 6, 0, "bih", 4),
    (64, 90, 4, 1, "lap", 3),
    (131, 77, 4, 3, None, 2),
])
def test_cnuity_mxlkta_matches_oracle(oracle, itdm, jtdm, kdm, nreg, mode, nstep):
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=31, m=m, n=n, nstep=nstep)
    thk = {"bih": 0.01, "lap": 0.02, None: 0.0}[mode]
    st = util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=thk, bih=mode != "lap")
    util.deepen_dpmixl(st, n)
    ref = util.run_oracle_cnuity(oracle, cb, sea, st, m, n, mxlkta=True)
    plain = util.run_oracle_cnuity(oracle, cb, sea, st, m, n, mxlkta=False)
    ts = pkg.Tsadvc(cb)
    ts.upload_cnuity_state(st, m, n)
    kw = {"bih": {"thkdf4": thk}, "lap": {"thkdf2": thk}, None: {}}[mode]
    ts.cnuity_device(m, n, mxlkta=True, **kw)
    _check(ts, g, cb, ref, m, n, kdm, dpmixl=True)
    inner = util.interior_sea(cb)
    assert not np.array_equal(ts.download(cabi.F_DPMIXL, n)[0][inner], plain["dpmixl"][n - 1][inner])
    ts.close()


def test_cnuity_mxlkta_on_tiles(oracle):
    from test_comm_gpu import make_tiles, run_tiles, close_tiles
    m, n = 1, 2
    itdm, jtdm, kdm, ipr, jpr, nreg, nstep = 160, 120, 4, 2, 2, 0, 5
    cfg, sea, g1, cb1 = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=23, m=m, n=n, nstep=nstep)
    st1 = util.deepen_dpmixl(util.add_cnuity(cfg, sea, g1, cb1, m, n, thkdf=0.01, bih=True), n)
    ref = util.run_oracle_cnuity(oracle, cb1, sea, st1, m, n, mxlkta=True)
    cbs = [syn.build_cb_arrays(cfg, g, sea, m, n, nstep=nstep) for g in pkg.partition(itdm, jtdm, kdm, ipr, jpr, nreg)]
    sts = [util.deepen_dpmixl(util.add_cnuity(cfg, sea, cb.geom, cb, m, n, uscale=st1["_uscale"], thkdf=0.01, bih=True), n)
           for cb in cbs]
    grp, tss = make_tiles(cfg, sea, itdm, jtdm, kdm, ipr, jpr, nreg, m, n, cbs=cbs)

    def go(ts, r):
        ts.upload_cnuity_state(sts[r], m, n)
        ts.cnuity_device(m, n, thkdf4=0.01, mxlkta=True)
    run_tiles(tss, go)
    for ts, cb in zip(tss, cbs):
        _check(ts, cb.geom, cb, ref, m, n, kdm, g1=g1, dpmixl=True)
    close_tiles(grp, tss)


@pytest.mark.parametrize("nreg,advtyp,thk", [(0, 2, 0.01), (3, 1, 0.0)])
def test_cnuity_then_tsadvc_device_chain_matches_oracle_chain(oracle, nreg, advtyp, thk):
    """the producer and its consumer back to back on the device mirrors - cnuity(m,n) leaves dp(:,:,:,n), uflx, vflx
    where tsadvc(m,n) reads them, nothing crosses PCIe in between - against the same chain in the oracle"""
    m, n = 1, 2
    itdm, jtdm, kdm = 150, 120, 5
    cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=1, seed=43, m=m, n=n, nstep=4, advtyp=advtyp)
    st = util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=thk, bih=True)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    util.oracle_load_cnuity(ot, st)
    ot.cnuity(m, n, 1)
    ot.tsadvc(m, n, 1)
    ts = pkg.Tsadvc(cb)
    ts.upload_state(m, n)
    ts.upload_cnuity_state(st, m, n)
    ts.cnuity_device(m, n, **({"thkdf4": thk} if thk else {}))
    ts.tsadvc_device(m, n)
    inner = util.interior_sea(cb)
    for fld, name in ((cabi.F_TEMP, "temp"), (cabi.F_SALN, "saln")):
        dev = ts.download(fld, n)
        for k in range(kdm):
            assert np.array_equal(dev[k][inner], ot.f64(name)[n - 1, k][inner]), (name, k)
    tr = ts.download(cabi.F_TRACER, n, ktr=1)
    for k in range(kdm):
        assert np.array_equal(tr[k][inner], ot.f64("tracer")[0, n - 1, k][inner]), ("tracer", k)
    # the advection really ran on what cnuity produced: not what tsadvc gives on the uploaded fluxes
    ot2 = util.oracle_tile_from_cb(oracle, cb, sea)
    ot2.tsadvc(m, n, 1)
    assert not np.array_equal(ot2.f64("saln")[n - 1, 0][inner], ot.f64("saln")[n - 1, 0][inner])
    ot.close(); ot2.close(); ts.close()


def test_cnuity_across_the_arctic_matches_oracle(oracle):
    """nreg=2 on one tile: every xctilr of cnuity through the tripole fold with its grid type"""
    m, n, kdm = 1, 2, 4
    cfg, sea, g, cb = util.make_arctic_case(96, 70, kdm, seed=17, m=m, n=n, nstep=4)
    st = util.arctic_halos_cnuity(g, util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=0.01, bih=True))
    ref = util.run_oracle_cnuity(oracle, cb, sea, st, m, n)
    ts = pkg.Tsadvc(cb)
    ts.upload_cnuity_state(st, m, n)
    ts.cnuity_device(m, n, thkdf4=0.01)
    _check(ts, g, cb, ref, m, n, kdm)
    ts.close()


def test_cnuity_across_the_arctic_on_tiles(oracle):
    """nreg=2 on 2 x 2 tiles: the top row exchanges the fold of every cnuity operand with its twin tiles"""
    from test_comm_gpu import make_tiles, run_tiles, close_tiles
    from test_parity_gpu import _tile_window
    m, n = 1, 2
    itdm, jtdm, kdm, ipr, jpr = 128, 70, 3, 2, 2
    cfg, sea, g1, cb1 = util.make_arctic_case(itdm, jtdm, kdm, seed=17, m=m, n=n, nstep=4)
    st1 = util.arctic_halos_cnuity(g1, util.add_cnuity(cfg, sea, g1, cb1, m, n, thkdf=0.01, bih=True))
    ref = util.run_oracle_cnuity(oracle, cb1, sea, st1, m, n)
    cbs = util.make_arctic_tiles(cfg, sea, cb1, ipr, jpr, m, n, nstep=4)
    sts = []
    for cb in cbs:   # every operand is the tile's window of the single-tile array
        g = cb.geom
        sts.append({k: (np.ascontiguousarray(_tile_window(v, g1, g, 2)) if isinstance(v, np.ndarray) else v)
                    for k, v in st1.items()})
    grp, tss = make_tiles(cfg, sea, itdm, jtdm, kdm, ipr, jpr, 2, m, n, cbs=cbs)

    def go(ts, r):
        ts.upload_cnuity_state(sts[r], m, n)
        ts.cnuity_device(m, n, thkdf4=0.01)
    run_tiles(tss, go)
    for ts, cb in zip(tss, cbs):
        _check(ts, cb.geom, cb, ref, m, n, kdm, g1=g1)
    close_tiles(grp, tss)


# the device against the REFERENCE'S OWN SOURCE TEXT: tests/golden/from_reference_text.json holds digests of what
# cnuity.F90 computes when executed as written (oracle/fortran_exec.py, tests/golden/make_reference_text_vectors.py)
# on the configurations of the tests above
import json as _json
import os as _os

import reftext_cases as _rc

_REFTEXT = _json.load(open(_os.path.join(_os.path.dirname(__file__), "golden", "from_reference_text.json")))


@pytest.mark.parametrize("name", sorted(_rc.CNUITY))
def test_cnuity_device_reproduces_the_reference_text(name):
    cfg, sea, g, cb, st, m, n, isopyc, mxlkta = _rc.cnuity_case(name)
    ts = pkg.Tsadvc(cb)
    ts.upload_cnuity_state(st, m, n)
    kw = {}
    if "_thkdf" in st:
        kw["thkdf4" if st["_thkdf"][1] else "thkdf2"] = st["_thkdf"][0]
    ts.cnuity_device(m, n, mxlkta=mxlkta, **kw)
    dp = np.stack([ts.download(cabi.F_DP, 1), ts.download(cabi.F_DP, 2)])
    d = _rc.cnuity_digest(cb, m, n, dp, ts.download(cabi.F_UFLX, 1), ts.download(cabi.F_VFLX, 1), ts.download(cabi.F_P, 1),
                          ts.download(cabi.F_DPMIXL, n)[0] if (isopyc or mxlkta) else None)
    ts.close()
    assert d == _REFTEXT[name], name
