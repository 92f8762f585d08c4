"""The library-owned transport of a multi-tile run (csrc/xc_comm.cu), on ONE GPU: every tile is a
handle of this process driven by its own host thread, the strips move between the handles'
staging buffers (the in-process transport).  The step code is exactly the one that runs under
NCCL (`bench.py --gpus N`, tools/xc_nccl_check.py): hycom_tsadvc_step_device / hycom_tsadvc_step
with the exchanges of mod_tsadvc.F90:1829-1836, :2140-2151 and :1186-1187 inside.
Reference idea: mod_pipe.F90:26-127 (1 tile vs N tiles must agree exactly)."""
import ctypes as C
import threading

import numpy as np
import pytest

import util
from util import pkg, syn, cabi

pytestmark = pytest.mark.gpu


class Group:
    def __init__(self, n):
        self.lib = cabi.load_library()
        self.h = C.c_void_p()
        assert self.lib.hycom_tsadvc_local_group_create(n, C.byref(self.h)) == 0

    def close(self):
        self.lib.hycom_tsadvc_local_group_destroy(self.h)


def run_tiles(tss, fn):
    """fn(ts, rank) on every tile concurrently (the calls are collective)"""
    err = [None] * len(tss)

    def work(r):
        try:
            fn(tss[r], r)
        except BaseException as e:  # noqa: BLE001
            err[r] = e
    th = [threading.Thread(target=work, args=(r,)) for r in range(len(tss))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None:
            raise e


def make_tiles(cfg, sea, itdm, jtdm, kdm, ipr, jpr, nreg, m, n, cbs=None, **scalars):
    grp = Group(ipr * jpr)
    tss = []
    if cbs is None:
        cbs = [syn.build_cb_arrays(cfg, g, sea, m, n, **scalars) for g in pkg.partition(itdm, jtdm, kdm, ipr, jpr, nreg)]
    for cb in cbs:
        ts = pkg.Tsadvc(cb)
        ts.comm_attach_local(grp.h)
        tss.append(ts)
    return grp, tss


def close_tiles(grp, tss):
    for ts in tss:
        ts.close()
    grp.close()


def check_tiles(tss, ref, n, ntracr, nb, names=("temp", "saln"), exact=True):
    fld_of = {"temp": cabi.F_TEMP, "saln": cabi.F_SALN, "th3d": cabi.F_TH3D}
    for ts in tss:
        g = ts.cb.geom
        sea_t = ts.cb.ip[nb:nb + g.jj, nb:nb + g.ii] != 0
        glob = (slice(None), slice(nb + g.j0, nb + g.j0 + g.jj), slice(nb + g.i0, nb + g.i0 + g.ii))
        pairs = [(fld_of[nm], 0, ref[nm][n - 1]) for nm in names]
        pairs += [(cabi.F_TRACER, q + 1, ref["tracer"][q, n - 1]) for q in range(ntracr)]
        for fld, ktr, r in pairs:
            dev = ts.download(fld, n, ktr=ktr)[:, nb:nb + g.jj, nb:nb + g.ii]
            if exact:
                assert np.array_equal(dev[:, sea_t], r[glob][:, sea_t]), (g.mproc, g.nproc, fld)
            else:
                assert np.allclose(dev[:, sea_t], r[glob][:, sea_t], rtol=1e-12, atol=0), (g.mproc, g.nproc, fld)


STEP_CASES = [
    # itdm, jtdm, kdm, ipr, jpr, nreg, ntracr, advtyp, overlap
    (150, 150, 3, 2, 1, 0, 0, 2, True),
    (150, 150, 3, 2, 2, 0, 1, 2, True),
    (150, 150, 2, 4, 2, 0, 0, 2, True),    # the 8-GPU tiling
    (131, 97, 2, 2, 2, 3, 0, 2, True),     # doubly periodic, ragged splits
    (120, 90, 2, 1, 2, 1, 1, 1, True),     # MPDATA, periodic in i wrapping onto the tile itself
    (300, 64, 2, 2, 1, 0, 0, 2, True),     # interior strips exist: interior and frame really overlap
    (300, 64, 2, 2, 1, 0, 0, 2, False),    # exchange first
    (150, 150, 2, 2, 2, 0, 0, 4, True),
    (150, 150, 2, 2, 2, 0, 1, 0, True),    # PCM, halo width 2
]


@pytest.mark.parametrize("itdm,jtdm,kdm,ipr,jpr,nreg,ntracr,advtyp,overlap", STEP_CASES)
def test_step_device_on_tiles_two_steps(oracle, itdm, jtdm, kdm, ipr, jpr, nreg, ntracr, advtyp, overlap):
    """hycom_tsadvc_step_device on ipr x jpr tiles, two leapfrog steps (1,2) then (2,1), == the oracle on
    one tile, bit for bit; the salinity range is the global one (xcminr/xcmaxr)"""
    cfg, sea, g1, cb1 = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=5, advtyp=advtyp, nstep=3)
    ot = util.oracle_tile_from_cb(oracle, cb1, sea)
    grp, tss = make_tiles(cfg, sea, itdm, jtdm, kdm, ipr, jpr, nreg, 1, 2, advtyp=advtyp, nstep=3)
    nb = g1.nbdy

    def setup(ts, r):
        ts.set_overlap(overlap)
        ts.upload_state(1, 2)
        ts.upload(cabi.F_DP, ts.cb.dp[0], 1)
    run_tiles(tss, setup)
    for step, (m, n) in enumerate(((1, 2), (2, 1))):
        ot.tsadvc(m, n, 1)
        ref = {"temp": ot.f64("temp").copy(), "saln": ot.f64("saln").copy()}
        if ntracr:
            ref["tracer"] = ot.f64("tracer").copy()
        run_tiles(tss, lambda ts, r: ts.tsadvc_device(m, n))
        check_tiles(tss, ref, n, ntracr, nb)
        for ts in tss:
            assert np.array_equal(ts.xmin, ot.f64("xmin")) and np.array_equal(ts.xmax, ot.f64("xmax"))
    ot.close()
    close_tiles(grp, tss)


def test_multi_tile_entries_refuse_without_communicator():
    """the drop-in entries never run a step with the exchanges left out (tsadvc_abi.cu:1555)"""
    cfg, sea, g1, cb1 = util.make_case(60, 40, 2, seed=3)
    cb = syn.build_cb_arrays(cfg, pkg.partition(60, 40, 2, 2, 1, 0)[0], sea, 1, 2)
    ts = pkg.Tsadvc(cb)
    with pytest.raises(cabi.TsadvcError) as e:
        ts.tsadvc(1, 2)
    assert e.value.code == cabi.EUNSUPPORTED
    ts.upload_state(1, 2)
    with pytest.raises(cabi.TsadvcError) as e:
        ts.tsadvc_device(1, 2)
    assert e.value.code == cabi.EUNSUPPORTED
    ts.close()


HOST_CASES = [
    # ipr, jpr, nreg, ntracr, advtyp, extra
    (2, 2, 0, 1, 2, {}),
    (2, 1, 3, 0, 1, {}),
    (2, 2, 0, 1, 2, {"temdf2": 0.02, "temdfc": 1.0, "sigver": 6}),     # + width-2 exchange, tsdff, EOS
    (2, 1, 3, 0, 2, {"temdf2": 0.02, "temdfc": 0.5, "sigver": 8}),
    (2, 2, 0, 1, 2, {"btrmas": True}),                                   # advem_fct2c: five in-scheme exchanges
    (1, 2, 1, 0, 2, {"btrmas": True}),
    (2, 2, 0, 0, 2, {"isopyc": True, "hybrid": False, "nhybrd": 0}),     # smoothed layer-1 fluxes read the halo
    (2, 1, 1, 2, 1, {"isopyc": True, "hybrid": False, "nhybrd": 0}),     # + tracers: two pairs of fluxes in layer 1
]


@pytest.mark.parametrize("ipr,jpr,nreg,ntracr,advtyp,extra", HOST_CASES)
def test_host_array_step_on_tiles(oracle, ipr, jpr, nreg, ntracr, advtyp, extra, monkeypatch):
    """THE drop-in entry hycom_tsadvc_step on host arrays of ipr x jpr tiles (several layer chunks), every
    exchange done by the library == the oracle on one tile"""
    monkeypatch.setenv("HYCOM_TSADVC_STEP_CHUNK", "2")
    m, n = 1, 2
    itdm, jtdm, kdm = 150, 120, 5
    diff = extra.get("temdf2", 0.0) > 0.0
    if diff:
        cfg, sea, g1, cb1 = util.make_diffusion_case(itdm, jtdm, kdm, extra["sigver"], extra["temdfc"], nreg=nreg,
                                                     ntracr=ntracr, seed=5, nstep=3, advtyp=advtyp)
        scal = dict(advtyp=advtyp, nstep=3, temdf2=cb1.temdf2, temdfc=extra["temdfc"], sigver=extra["sigver"],
                    thbase=cb1.thbase)
    else:
        cfg, sea, g1, cb1 = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=5, m=m, n=n,
                                           advtyp=advtyp, nstep=3, **extra)
        scal = dict(advtyp=advtyp, nstep=3, **extra)
    ref = util.run_oracle(oracle, cb1, sea, m, n)
    from test_parity_gpu import _tile_window
    cbs = []
    for g in pkg.partition(itdm, jtdm, kdm, ipr, jpr, nreg):
        cb = syn.build_cb_arrays(cfg, g, sea, m, n, **scal)
        if diff or extra.get("isopyc"):
            cb.th3d = np.ascontiguousarray(_tile_window(cb1.th3d, g1, g, nreg))
            if cb1.theta is not None:
                cb.theta = np.ascontiguousarray(_tile_window(cb1.theta, g1, g, nreg))
        cbs.append(cb)
    grp, tss = make_tiles(cfg, sea, itdm, jtdm, kdm, ipr, jpr, nreg, m, n, cbs=cbs)
    run_tiles(tss, lambda ts, r: ts.tsadvc(m, n))
    nb = g1.nbdy
    names = ["temp", "saln"] + (["th3d"] if diff else [])
    if extra.get("isopyc"):
        names = ["saln", "th3d"]        # th3d & saln in layer 1 on the smoothed fluxes, saln only below
    for ts in tss:   # the host arrays hold the result on 1:ii,1:jj
        g, cb = ts.cb.geom, ts.cb
        sea_t = cb.ip[nb:nb + g.jj, nb:nb + g.ii] != 0
        glob = (slice(None), slice(nb + g.j0, nb + g.j0 + g.jj), slice(nb + g.i0, nb + g.i0 + g.ii))
        for nm in names:
            got = getattr(cb, nm)[n - 1][:, nb:nb + g.jj, nb:nb + g.ii]
            assert np.array_equal(got[:, sea_t], ref[nm][n - 1][glob][:, sea_t]), (g.mproc, g.nproc, nm)
        for q in range(ntracr):
            got = cb.tracer[q, n - 1][:, nb:nb + g.jj, nb:nb + g.ii]
            assert np.array_equal(got[:, sea_t], ref["tracer"][q, n - 1][glob][:, sea_t]), (g.mproc, g.nproc, "tracer", q)
        assert np.array_equal(ts.xmin, ref["xmin"]) and np.array_equal(ts.xmax, ref["xmax"])
    close_tiles(grp, tss)


@pytest.mark.parametrize("itdm,jtdm,kdm,ipr,jpr,ntracr,advtyp,temdf2", [
    (128, 70, 2, 2, 2, 0, 2, 0.0),
    (192, 64, 2, 4, 2, 1, 2, 0.0),     # 4x2: the shifted u-grid column comes from a third tile
    (128, 70, 2, 2, 1, 0, 1, 0.0),
    (128, 70, 2, 2, 2, 1, 2, 0.02),
])
def test_arctic_step_device_on_tiles(oracle, itdm, jtdm, kdm, ipr, jpr, ntracr, advtyp, temdf2):
    """tripole fold of the top row through the library's own transport"""
    m, n = 1, 2
    extra = dict(advtyp=advtyp, nstep=3)
    if temdf2 > 0.0:
        extra.update(temdf2=temdf2, temdfc=1.0, sigver=6, thbase=34.0)
    cfg, sea, g1, cb1 = util.make_arctic_case(itdm, jtdm, kdm, ntracr=ntracr, seed=29, m=m, n=n, **extra)
    ref = util.run_oracle(oracle, cb1, sea, m, n)
    cbs = util.make_arctic_tiles(cfg, sea, cb1, ipr, jpr, m, n, **extra)
    grp, tss = make_tiles(cfg, sea, itdm, jtdm, kdm, ipr, jpr, 2, m, n, cbs=cbs)

    def go(ts, r):
        ts.upload_state(m, n)
        ts.tsadvc_device(m, n)
    run_tiles(tss, go)
    check_tiles(tss, ref, n, ntracr, g1.nbdy, names=("temp", "saln") + (("th3d",) if temdf2 > 0 else ()))
    close_tiles(grp, tss)


def np_checksum(a, ip, g, k_count):
    """numpy restatement of hycom_tsadvc_checksum: a (kdm, nrows, ncols) of ONE tile, global offsets of g"""
    def mix64(z):
        z = (z + np.uint64(0x9e3779b97f4a7c15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
        return z ^ (z >> np.uint64(31))
    nb = g.nbdy
    tot = np.uint64(0)
    with np.errstate(over="ignore"):
        for k in range(k_count):
            v = a[k, nb:nb + g.jj, nb:nb + g.ii]
            sea = ip[nb:nb + g.jj, nb:nb + g.ii] != 0
            gj, gi = np.meshgrid(np.arange(g.jj) + g.j0, np.arange(g.ii) + g.i0, indexing="ij")
            cell = (gi + g.itdm * (gj + g.jtdm * k)).astype(np.uint64)
            bits = np.where(v == 0.0, 0.0, v).view(np.uint64)
            h = mix64(bits ^ mix64(cell))
            tot = tot + h[sea].sum(dtype=np.uint64)
    return int(tot)


def test_checksum_is_tiling_invariant(oracle):
    """the PIPE_CHECK analogue: 1 tile, 2x2 tiles and the numpy restatement on the oracle's output agree"""
    m, n = 1, 2
    itdm, jtdm, kdm = 150, 120, 3
    cfg, sea, g1, cb1 = util.make_case(itdm, jtdm, kdm, seed=5, nstep=3)
    ref = util.run_oracle(oracle, cb1, sea, m, n)
    want = np_checksum(ref["saln"][n - 1], cb1.ip, g1, kdm)
    ts1 = pkg.Tsadvc(cb1)
    ts1.upload_state(m, n)
    ts1.tsadvc_device(m, n)
    assert ts1.checksum(cabi.F_SALN, n) == want
    ts1.close()
    grp, tss = make_tiles(cfg, sea, itdm, jtdm, kdm, 2, 2, 0, m, n, nstep=3)
    sums = [None] * 4

    def go(ts, r):
        ts.upload_state(m, n)
        ts.tsadvc_device(m, n)
        sums[r] = ts.checksum(cabi.F_SALN, n)
    run_tiles(tss, go)
    assert sums == [want] * 4
    part = [ts.checksum(cabi.F_SALN, n, all_tiles=False) for ts in tss]
    assert sum(part) % (1 << 64) == want and len(set(part)) == 4
    close_tiles(grp, tss)


def test_deferred_salinity_range(oracle):
    cfg, sea, g1, cb1 = util.make_case(90, 60, 2, seed=7, nstep=3)
    ref = util.run_oracle(oracle, cb1, sea, 1, 2)
    ts = pkg.Tsadvc(cb1)
    ts.upload_state(1, 2)
    ts.set_deferred_range(True)
    ts.xmin[:] = np.nan
    ts.tsadvc_device(1, 2, diag=False)        # nothing returned by the step itself
    assert np.isnan(ts.xmin).all()
    ns, xmin, xmax = ts.saln_range()
    assert ns == 3 and np.array_equal(xmin, ref["xmin"]) and np.array_equal(xmax, ref["xmax"])
    assert ts.saln_range()[0] == -1
    ts.close()
