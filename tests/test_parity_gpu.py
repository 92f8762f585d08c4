"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Bar (BASELINE.json north_star): <= 1e-12 relative
max error per field and layer, land cells bit-unchanged.  The kernels keep the
reference's operation order without FMA contraction, so the tests additionally demand
bit equality with the unfused oracle."""
import ctypes as C

import numpy as np
import pytest

import util
from util import pkg, syn, cabi, REL_TOL

pytestmark = pytest.mark.gpu


def _compare(cb, g, got, ref, n, names, exact=True):
    msk = util.interior_sea(cb)
    land = ~msk
    worst = 0.0
    for name in names:
        a, b = got[name], ref[name]
        for k in range(g.kdm):
            e = util.rel_err(a[n - 1, k], b[n - 1, k], msk)
            worst = max(worst, e)
            assert e <= REL_TOL, (name, k, e)
            if exact:
                assert np.array_equal(a[n - 1, k][msk], b[n - 1, k][msk]), (name, k, "not bit-exact")
    return worst


def _run_host_path(cb, m, n):
    """the drop-in call on host arrays; returns the updated arrays"""
    ts = pkg.Tsadvc(cb)
    before = {k: getattr(cb, k).copy() for k in ("temp", "saln", "th3d")}
    if cb.ntracr:
        before["tracer"] = cb.tracer.copy()
    ts.tsadvc(m, n)
    got = dict(temp=cb.temp, saln=cb.saln, th3d=cb.th3d, tracer=cb.tracer, xmin=ts.xmin.copy(),
               xmax=ts.xmax.copy())
    launches = ts.launch_count
    ts.close()
    return got, before, launches


CASES = [
    # itdm, jtdm, kdm, nreg, ntracr, advtyp, extra
    (150, 150, 22, 0, 0, 2, {}),                 # BASELINE configs[0]: box basin, FCT2 T+S
    (150, 150, 22, 0, 2, 1, {"trcflg": [0, 2]}), # MPDATA + tracers
    (131, 77, 3, 3, 1, 2, {}),                   # doubly periodic f-plane, odd row length
    (64, 203, 2, 1, 0, 2, {}),                   # periodic in i only, many rows (chunk seams)
    (200, 60, 2, 4, 0, 1, {}),                   # closed f-plane (periodic in j), strip seams
    (9, 8, 2, 0, 0, 2, {}),                      # tiny
    (58, 31, 4, 0, 1, 2, {"nhybrd": 2}),         # temp only in the top nhybrd layers
    (70, 45, 3, 0, 0, 2, {"advflg": 1}),         # advect th3d & S
    (70, 45, 3, 0, 0, 1, {"advflg": 1}),
    (150, 150, 5, 0, 1, 4, {}),                  # advem_fct4 (4th-order high-order flux)
    (97, 83, 3, 3, 0, 4, {}),                    # fct4, doubly periodic
    (150, 150, 5, 0, 1, 0, {}),                  # advem_pcm (donor cell), halo width 2
    (64, 90, 2, 1, 0, 0, {}),
]


@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,ntracr,advtyp,extra", CASES)
def test_tsadvc_host_path_matches_oracle(oracle, itdm, jtdm, kdm, nreg, ntracr, advtyp, extra):
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=13, m=m, n=n,
                                     advtyp=advtyp, nstep=3, **extra)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    got, before, launches = _run_host_path(cb, m, n)
    assert launches > 0
    names = ["saln"] + (["th3d"] if cb.advflg else ["temp"])
    _compare(cb, g, got, ref, n, names)
    for q in range(ntracr):
        _compare(cb, g, {"t": got["tracer"][q]}, {"t": ref["tracer"][q]}, n, ["t"])
    msk = util.interior_sea(cb)
    # land cells and the centre time level keep their bits; NaN halos of the host arrays too
    for name in names:
        a, b = got[name], before[name]
        assert np.array_equal(a[m - 1], b[m - 1], equal_nan=True), (name, "slot m modified")
        assert np.array_equal(a[n - 1][:, ~msk], b[n - 1][:, ~msk], equal_nan=True), (name, "land/halo modified")
    # fields that are not advected are untouched
    idle = "temp" if cb.advflg else "th3d"
    assert np.array_equal(got[idle], before[idle], equal_nan=True)
    # diagnostics (:2065-2094)
    assert np.array_equal(got["xmin"], ref["xmin"])
    assert np.array_equal(got["xmax"], ref["xmax"])
    # something moved
    assert not np.array_equal(got["saln"][n - 1, 0][msk], before["saln"][n - 1, 0][msk])


@pytest.mark.parametrize("m,n", [(1, 2), (2, 1)])
def test_leapfrog_slots_and_two_steps(oracle, m, n):
    """two consecutive calls with swapped slots, as HYCOM_Run does (mod_hycom.F90:2254-2257)"""
    cfg, sea, g, cb = util.make_case(90, 70, 3, nreg=0, seed=21, m=m, n=n, advtyp=2)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    ts = pkg.Tsadvc(cb)
    ts.upload_state(m, n)
    ts.upload(cabi.F_DP, cb.dp[m - 1], m)
    ot.tsadvc(m, n, 1)
    ts.tsadvc_device(m, n)
    ot.tsadvc(n, m, 1)
    ts.tsadvc_device(n, m)
    msk = util.interior_sea(cb)
    for fld, name in ((cabi.F_TEMP, "temp"), (cabi.F_SALN, "saln")):
        for slot in (1, 2):
            dev = ts.download(fld, slot)
            ref = ot.f64(name)[slot - 1]
            for k in range(g.kdm):
                assert np.array_equal(dev[k][msk], ref[k][msk]), (name, slot, k)
    ts.close()
    ot.close()


def test_device_generator_matches_host_generator():
    cfg, sea, g, cb = util.make_case(83, 61, 3, nreg=1, ntracr=1, seed=17)
    ts = pkg.Tsadvc(cb)
    syn.fill_device(ts, cfg, sea, 1, 2)
    for fld, name in ((cabi.F_TEMP, "temp"), (cabi.F_SALN, "saln")):
        for slot in (1, 2):
            assert np.array_equal(ts.download(fld, slot), getattr(cb, name)[slot - 1], equal_nan=True)
    assert np.array_equal(ts.download(cabi.F_DP, 2), cb.dp[1], equal_nan=True)
    assert np.array_equal(ts.download(cabi.F_UFLX, 1), cb.uflx, equal_nan=True)
    assert np.array_equal(ts.download(cabi.F_VFLX, 1), cb.vflx, equal_nan=True)
    assert np.array_equal(ts.download(cabi.F_TRACER, 1, ktr=1), cb.tracer[0, 0], equal_nan=True)
    ts.close()


def test_device_halo_matches_xctilr(oracle):
    """hycom_tsadvc_halo_local == xctilr of mod_xc_sm.h:1337-1428, closed and periodic"""
    for nreg in (0, 1, 3, 4):
        cfg, sea, g, cb = util.make_case(37, 29, 2, nreg=nreg, seed=3)
        ot = util.oracle_tile_from_cb(oracle, cb, sea)
        ts = pkg.Tsadvc(cb)
        ts.upload(cabi.F_SALN, cb.saln[0], 1)
        ts.upload(cabi.F_SALN, cb.saln[1], 2)
        ts._ck(ts.lib.hycom_tsadvc_halo_local(ts.h, cabi.F_SALN, 0, 0, 5, 5))
        a = ot.f64("saln")
        ot.xctilr(a, 1, 2 * g.kdm, 5, 5)
        nb = g.nbdy
        live = (slice(None), slice(nb - 5, nb + g.jj + 5), slice(nb - 5, nb + g.ii + 5))
        for slot in (1, 2):
            dev = ts.download(cabi.F_SALN, slot)
            assert np.array_equal(dev[live], a[slot - 1][live], equal_nan=True), nreg
            # cells beyond the refreshed halo are never read by the reference; the device
            # sets them to vland so that no NaN enters the recomputed aprons
            dev[live] = 0.0
            assert (dev == 0.0).all()
        ts.close()
        ot.close()


def test_error_behaviour_mirrors_xcstop():
    cfg, sea, g, cb = util.make_case(30, 30, 1, seed=1, advtyp=3)
    ts = pkg.Tsadvc(cb)
    ts.upload_state(1, 2)
    with pytest.raises(cabi.XcStop, match="advem called with advtyp"):
        ts.tsadvc_device(1, 2)
    cb.advtyp = 2
    with pytest.raises(cabi.TsadvcError):
        ts.tsadvc_device(1, 1)  # m == n
    ts.close()
    # nbdy < mbdy_advtyp -> xcstop('tsadvc') (mod_tsadvc.F90:1817-1825)
    g4 = pkg.partition(30, 30, 1, 1, 1, 0, nbdy=4)[0]
    cb4 = syn.build_cb_arrays(cfg, g4, sea, 1, 2, advtyp=2)
    ts = pkg.Tsadvc(cb4)
    ts.upload_state(1, 2)
    with pytest.raises(cabi.XcStop, match="nbdy"):
        ts.tsadvc_device(1, 2)
    ts.close()


@pytest.mark.parametrize("advtyp", [1, 2])
def test_temperature_tracer_equals_temp_on_device(advtyp):
    """PIPE_TRACER invariant (mod_pipe.F90:1517-1542) through the CUDA path"""
    cfg, sea, g, cb = util.make_case(120, 90, 4, nreg=0, ntracr=1, seed=4, advtyp=advtyp, trcflg=[2])
    cb.tracer[0] = cb.temp
    got, before, _ = _run_host_path(cb, 1, 2)
    msk = util.interior_sea(cb)
    assert np.array_equal(got["tracer"][0, 1][:, msk], got["temp"][1][:, msk])


def test_constant_field_preserved_on_device():
    cfg, sea, g, cb = util.make_case(100, 80, 2, nreg=0, seed=2, advtyp=2)
    cb.saln[:] = np.where(cb.ip != 0, 35.25, cb.saln)
    got, before, _ = _run_host_path(cb, 1, 2)
    msk = util.interior_sea(cb)
    assert (got["saln"][1][:, msk] == 35.25).all()


@pytest.mark.parametrize("advtyp", [2, 1, 4])
def test_all_sea_segments_equal_general_launch(oracle, advtyp, monkeypatch):
    """FCT2 / MPDATA / FCT4 as the launch pair (all-sea row segments on the mask-free instantiation + the rest)
    == the single general launch == the oracle, bit for bit, on a grid with wide open water"""
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(420, 360, 2, nreg=0, ntracr=1, seed=12, m=m, n=n, advtyp=advtyp)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    out = {}
    # the pair with all-sea pieces = whole bands (the default, short bands so that several fit), with maximal
    # all-sea runs cut at the band boundaries, and the single general launch
    for key, env in (("whole", {"HYCOM_TSADVC_SPLIT": "1", "HYCOM_TSADVC_SEG_WHOLE": "1", "HYCOM_TSADVC_SEG_BAND": "96"}),
                     ("runs", {"HYCOM_TSADVC_SPLIT": "1", "HYCOM_TSADVC_SEG_WHOLE": "0", "HYCOM_TSADVC_SEG_BAND": "128"}),
                     ("one", {"HYCOM_TSADVC_SPLIT": "0"})):
        for name in ("HYCOM_TSADVC_SPLIT", "HYCOM_TSADVC_SEG_WHOLE", "HYCOM_TSADVC_SEG_BAND"):
            monkeypatch.delenv(name, raising=False)
        for name, val in env.items():
            monkeypatch.setenv(name, val)
        cbx = syn.build_cb_arrays(cfg, g, sea, m, n, advtyp=advtyp)
        got, _, launches = _run_host_path(cbx, m, n)
        out[key] = (got, launches)
        _compare(cbx, g, got, ref, n, ["saln", "temp"])
        _compare(cbx, g, {"t": got["tracer"][0]}, {"t": ref["tracer"][0]}, n, ["t"])
    # the pair really ran: one marching launch more per call of run_march than the general launch alone
    assert out["whole"][1] > out["one"][1], (out["whole"][1], out["one"][1])
    assert out["runs"][1] > out["one"][1], (out["runs"][1], out["one"][1])
    for name in ("temp", "saln"):
        assert np.array_equal(out["whole"][0][name], out["one"][0][name], equal_nan=True)
        assert np.array_equal(out["runs"][0][name], out["one"][0][name], equal_nan=True)


@pytest.mark.parametrize("advtyp,ntracr", [(2, 0), (1, 1)])
def test_full_size_glb_layers_match_oracle(oracle, advtyp, ntracr):
    """BASELINE configs[1]/[2] at the full GLBb0.08 horizontal size (4500x3298) on the
    device-resident path; two of the layers are checked against the oracle (each layer
    is independent of the others, mod_tsadvc.F90:1842), the rest through the
    size-independent properties: land untouched, finite on sea, tracer bounds."""
    itdm, jtdm, kdm = 4500, 3298, 6
    m, n = 1, 2
    cfg = syn.make_cfg(itdm, jtdm, kdm, nreg=0, ntracr=ntracr, seed=1, dx0=8900.0, delt1=480.0)
    sea = syn.sea_mask(cfg)
    g = pkg.partition(itdm, jtdm, kdm, 1, 1, 0)[0]
    cb = syn.build_cb_arrays(cfg, g, sea, m, n, with_state=False, advtyp=advtyp)
    ts = pkg.Tsadvc(cb)
    syn.fill_device(ts, cfg, sea, m, n)
    ts.tsadvc_device(m, n, diag=False)
    ts.synchronize()
    # oracle on layers k0..k0+1 only: a 2-layer tile filled with the same global function
    k0, nk = 3, 2
    g2 = pkg.partition(itdm, jtdm, nk, 1, 1, 0)[0]
    cb2 = syn.build_cb_arrays(cfg, g2, sea, m, n, with_state=False, advtyp=advtyp)
    cb2.ntracr = ntracr

    def f4(fld, ktr=0, halo_mode=0):
        a = np.empty((2, nk, g.nrows, g.ncols))
        for slot in (1, 2):
            a[slot - 1] = syn.fill_host(cfg, g, sea, fld, ktr, 0 if slot == n else 1, k0, nk, halo_mode)
        return a
    cb2.temp, cb2.saln = f4(cabi.F_TEMP), f4(cabi.F_SALN)
    cb2.th3d = np.zeros_like(cb2.temp)
    cb2.dp = f4(cabi.F_DP, halo_mode=1)
    cb2.uflx = syn.fill_host(cfg, g, sea, cabi.F_UFLX, 0, 0, k0, nk, 0)
    cb2.vflx = syn.fill_host(cfg, g, sea, cabi.F_VFLX, 0, 0, k0, nk, 0)
    cb2.oneta = np.ones((2, g.nrows, g.ncols))
    if ntracr:
        cb2.tracer = np.stack([f4(cabi.F_TRACER, ktr=q + 1) for q in range(ntracr)])
    ref = util.run_oracle(oracle, cb2, sea, m, n)
    msk = util.interior_sea(cb)
    for fld, name in ((cabi.F_TEMP, "temp"), (cabi.F_SALN, "saln")):
        dev = ts.download(fld, n, k0=k0, nk=nk)
        for k in range(nk):
            e = util.rel_err(dev[k], ref[name][n - 1, k], msk)
            assert e <= REL_TOL, (name, k, e)
            assert np.array_equal(dev[k][msk], ref[name][n - 1, k][msk]), (name, k)
            # land cells inside 1:ii,1:jj keep their bits (the 2**100 sentinel)
            nb = g.nbdy
            inner = np.zeros_like(msk)
            inner[nb:nb + g.jj, nb:nb + g.ii] = True
            land = inner & ~msk
            assert np.array_equal(dev[k][land], getattr(cb2, name)[n - 1, k][land])
    if ntracr:
        dev = ts.download(cabi.F_TRACER, n, ktr=1, k0=k0, nk=nk)
        for k in range(nk):
            assert np.array_equal(dev[k][msk], ref["tracer"][0, n - 1, k][msk])
    # the same two layers against the COMPILED REFERENCE TEXT itself (oracle/_ref), when this machine has it
    import copy
    cbt = copy.deepcopy(cb2)
    if util.run_compiled_reference_text(cbt, sea, m, n):
        for fld, name in ((cabi.F_TEMP, "temp"), (cabi.F_SALN, "saln")):
            dev = ts.download(fld, n, k0=k0, nk=nk)
            assert np.array_equal(dev[:, msk], getattr(cbt, name)[n - 1][:, msk]), (name, "reference text")
        if ntracr:
            dev = ts.download(cabi.F_TRACER, n, ktr=1, k0=k0, nk=nk)
            assert np.array_equal(dev[:, msk], cbt.tracer[0, n - 1][:, msk]), "tracer vs reference text"
    # remaining layers: properties
    for k in (1, kdm):
        s = ts.download(cabi.F_SALN, n, k0=k, nk=1)[0]
        assert np.isfinite(s[msk]).all()
        assert 25.0 < s[msk].min() and s[msk].max() < 45.0
    ts.close()


# ---------------------------------------------------------------------------------------
# multi-tile: the reference's own test idea (mod_pipe.F90:26-127: 1 tile vs N tiles, exact)
# ---------------------------------------------------------------------------------------
def _exchange_in_process(tss, m, n):
    """xctilr between the tiles of one process: every tile packs on the device, a message
    leaving tile A in direction d is unpacked by tile nbr_A[d] as arriving from OPP[d]."""
    import torch
    xc = __import__("importlib").import_module("hycom-src_b200.xc")
    sends, nbrs = [], []
    for ts in tss:
        be = xc.DeviceHaloBackend(ts)
        nbr = xc.neighbors(ts.cb.geom)
        cnt = be.counts(m, n)
        mb = 2 if ts.cb.advtyp == 0 else 5      # mbdy_advtyp (mod_tsadvc.F90:24-29)
        assert cnt == xc.halo_counts(ts.cb.geom, (2 * (2 + ts.cb.ntracr) + 2) * ts.cb.geom.kdm, mb, mb)
        send = [be.alloc(c) if nbr[d] >= 0 else None for d, c in enumerate(cnt)]
        be.pack(m, n, send)
        sends.append(send)
        nbrs.append(nbr)
        ts.synchronize()
    for t, ts in enumerate(tss):
        recv = [sends[nbrs[t][d]][xc.opp_dir(ts.cb.geom, d)] if nbrs[t][d] >= 0 else None for d in range(8)]
        xc.DeviceHaloBackend(ts).unpack(m, n, recv)
        ts.synchronize()


TILINGS = [
    # itdm, jtdm, kdm, ipr, jpr, nreg, ntracr, advtyp, split
    (150, 150, 3, 2, 1, 0, 0, 2, True),    # BASELINE configs[3] tilings on the box basin
    (150, 150, 3, 2, 2, 0, 1, 2, True),
    (150, 150, 2, 4, 2, 0, 0, 2, False),
    (131, 97, 2, 2, 2, 3, 0, 2, True),     # doubly periodic, ragged splits
    (120, 90, 2, 1, 2, 1, 1, 1, True),     # MPDATA, periodic in i wrapping onto the tile itself
    (300, 64, 2, 2, 1, 0, 0, 2, True),     # tiles wide enough to have interior strips
    (150, 150, 2, 2, 2, 0, 0, 4, True),    # advem_fct4 on 2x2 tiles
    (150, 150, 2, 2, 2, 0, 1, 0, True),    # advem_pcm on 2x2 tiles (halo width 2)
]


@pytest.mark.parametrize("itdm,jtdm,kdm,ipr,jpr,nreg,ntracr,advtyp,split", TILINGS)
def test_tiling_invariance_on_device(oracle, itdm, jtdm, kdm, ipr, jpr, nreg, ntracr, advtyp, split):
    """N tiles (device pack/unpack, interior/frame split) == the oracle on 1 tile, bit for bit"""
    m, n = 1, 2
    cfg, sea, g1, cb1 = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=5, m=m, n=n,
                                       advtyp=advtyp, nstep=3)
    ref = util.run_oracle(oracle, cb1, sea, m, n)
    tiles = pkg.partition(itdm, jtdm, kdm, ipr, jpr, nreg)
    tss = []
    for g in tiles:
        cb = syn.build_cb_arrays(cfg, g, sea, m, n, advtyp=advtyp, nstep=3)
        ts = pkg.Tsadvc(cb)
        ts.upload_state(m, n)
        tss.append(ts)
    _exchange_in_process(tss, m, n)
    xmin = np.full(kdm, 999.0)
    xmax = np.full(kdm, -999.0)
    for ts in tss:
        p = ts.cb.params()
        xm, xx = ts.xmin.ctypes.data_as(C.c_void_p), ts.xmax.ctypes.data_as(C.c_void_p)
        if split:
            ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_INTERIOR, None, None))
            ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_FRAME, xm, xx))
        else:
            ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_ALL, xm, xx))
        xmin, xmax = np.minimum(xmin, ts.xmin), np.maximum(xmax, ts.xmax)   # xcminr/xcmaxr
    nb = g1.nbdy
    for ts in tss:
        g = ts.cb.geom
        sea_t = ts.cb.ip[nb:nb + g.jj, nb:nb + g.ii] != 0
        glob = (slice(None), slice(nb + g.j0, nb + g.j0 + g.jj), slice(nb + g.i0, nb + g.i0 + g.ii))
        pairs = [(cabi.F_TEMP, 0, ref["temp"][n - 1]), (cabi.F_SALN, 0, ref["saln"][n - 1])]
        pairs += [(cabi.F_TRACER, q + 1, ref["tracer"][q, n - 1]) for q in range(ntracr)]
        for fld, ktr, r in pairs:
            dev = ts.download(fld, n, ktr=ktr)[:, nb:nb + g.jj, nb:nb + g.ii]
            assert np.array_equal(dev[:, sea_t], r[glob][:, sea_t]), (g.mproc, g.nproc, fld)
        ts.close()
    assert np.array_equal(xmin, ref["xmin"]) and np.array_equal(xmax, ref["xmax"])


# ---------------------------------------------------------------------------------------
# temdf2 > 0: tsdff_1x/2x + the equation-of-state sweep (mod_tsadvc.F90:2138-2230)
# ---------------------------------------------------------------------------------------
DIFF_CASES = [
    # itdm, jtdm, kdm, nreg, ntracr, advtyp, sigver, temdfc, nhybrd
    (150, 150, 22, 0, 0, 2, 6, 1.0, -1),   # box basin, the GLB build's EOS (sigma-2, 17-term): temp & saln diffused
    (150, 150, 4, 0, 3, 2, 5, 1.0, -1),    # odd number of tracers: tsdff_2x pair + tsdff_1x
    (131, 77, 3, 3, 2, 1, 8, 0.5, -1),     # 12-term: th3d & temp diffused, combined in density space
    (64, 203, 2, 1, 0, 2, 8, 0.0, -1),     # th3d & saln diffused, temp = tofsig
    (70, 45, 4, 0, 1, 2, 7, 0.5, 2),       # exactly-isopycnal layers below nhybrd: th3d = theta
    (90, 61, 3, 0, 0, 2, 2, 1.0, -1),      # 7-term sig
    (90, 61, 3, 4, 1, 2, 4, 1.0, -1),      # 9-term sig
    (90, 61, 3, 0, 0, 2, 1, 0.5, -1),      # 7-term tofsig: atan2/cos/sin (tolerance, not bits)
    (90, 61, 3, 0, 0, 2, 3, 0.0, 2),       # 9-term tofsig
    (70, 45, 3, 0, 0, 2, 8, 0.5, -1),
]


@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,ntracr,advtyp,sigver,temdfc,nhybrd", DIFF_CASES)
def test_diffusion_host_path_matches_oracle(oracle, itdm, jtdm, kdm, nreg, ntracr, advtyp, sigver, temdfc, nhybrd):
    m, n = 1, 2
    cfg, sea, g, cb = util.make_diffusion_case(itdm, jtdm, kdm, sigver, temdfc, nreg=nreg, ntracr=ntracr,
                                               nhybrd=nhybrd, seed=17, advtyp=advtyp, nstep=3)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    got, before, launches = _run_host_path(cb, m, n)
    # device atan2/cos/sin are not glibc's to the last bit: 7/9-term tofsig within 1e-12 only
    libm = sigver <= 4 and (temdfc < 1.0 or (0 <= nhybrd < kdm))
    _compare(cb, g, got, ref, n, ["saln"])
    _compare(cb, g, got, ref, n, ["temp", "th3d"], exact=not libm)
    for q in range(ntracr):
        _compare(cb, g, {"t": got["tracer"][q]}, {"t": ref["tracer"][q]}, n, ["t"])
    msk = util.interior_sea(cb)
    for name in ("temp", "saln", "th3d"):
        a, b = got[name], before[name]
        assert np.array_equal(a[m - 1], b[m - 1], equal_nan=True), (name, "slot m modified")
        assert np.array_equal(a[n - 1][:, ~msk], b[n - 1][:, ~msk], equal_nan=True), (name, "land/halo modified")
    assert np.array_equal(got["xmin"], ref["xmin"]) and np.array_equal(got["xmax"], ref["xmax"])


def test_diffusion_needs_eos_and_metrics():
    cfg, sea, g, cb = util.make_diffusion_case(40, 30, 2, 6, 1.0)
    cb.sigver = 0
    ts = pkg.Tsadvc(cb)
    with pytest.raises(cabi.TsadvcError):
        ts.tsadvc(1, 2)
    ts.close()


@pytest.mark.parametrize("ipr,jpr,nreg,ntracr,sigver,temdfc", [(2, 2, 0, 1, 6, 1.0), (2, 1, 3, 0, 8, 0.5)])
def test_diffusion_tiling_invariance_on_device(oracle, ipr, jpr, nreg, ntracr, sigver, temdfc):
    """N tiles with the second exchange (width 2, :2140-2151) == the oracle on 1 tile"""
    import torch  # noqa: F401
    xc = __import__("importlib").import_module("hycom-src_b200.xc")
    m, n = 1, 2
    itdm, jtdm, kdm = 150, 120, 3
    cfg, sea, g1, cb1 = util.make_diffusion_case(itdm, jtdm, kdm, sigver, temdfc, nreg=nreg, ntracr=ntracr,
                                                 seed=5, nstep=3)
    ref = util.run_oracle(oracle, cb1, sea, m, n)
    nb = g1.nbdy
    tss = []
    for g in pkg.partition(itdm, jtdm, kdm, ipr, jpr, nreg):
        cb = syn.build_cb_arrays(cfg, g, sea, m, n, nstep=3, temdf2=cb1.temdf2, temdfc=temdfc, sigver=sigver,
                                 thbase=cb1.thbase)
        # the tile's window of the global th3d/theta (halo cells come from the exchange)
        cb.th3d = np.ascontiguousarray(_tile_window(cb1.th3d, g1, g, nreg))
        cb.theta = np.ascontiguousarray(_tile_window(cb1.theta, g1, g, nreg))
        ts = pkg.Tsadvc(cb)
        ts.upload_state(m, n)
        tss.append(ts)
    _exchange_in_process(tss, m, n)
    for ts in tss:
        p = ts.cb.params()
        ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_ALL, None, None))
    # second exchange, then diffusion
    sends, nbrs = [], []
    for ts in tss:
        be = xc.DeviceHaloBackend(ts)
        nbr = xc.neighbors(ts.cb.geom)
        cnt = be.diff_counts(n)
        assert cnt == xc.halo_counts(ts.cb.geom, (3 + ntracr) * kdm, 2, 2)
        send = [be.alloc(c) if nbr[d] >= 0 else None for d, c in enumerate(cnt)]
        be.diff_pack(n, send)
        sends.append(send); nbrs.append(nbr)
        ts.synchronize()
    for t, ts in enumerate(tss):
        recv = [sends[nbrs[t][d]][xc.opp_dir(ts.cb.geom, d)] if nbrs[t][d] >= 0 else None for d in range(8)]
        xc.DeviceHaloBackend(ts).diff_unpack(n, recv)
        p = ts.cb.params()
        ts._ck(ts.lib.hycom_tsadvc_diffuse_device(ts.h, m, n, C.byref(p)))
        ts.synchronize()
    for ts in tss:
        g = ts.cb.geom
        sea_t = ts.cb.ip[nb:nb + g.jj, nb:nb + g.ii] != 0
        glob = (slice(None), slice(nb + g.j0, nb + g.j0 + g.jj), slice(nb + g.i0, nb + g.i0 + g.ii))
        pairs = [(cabi.F_TEMP, 0, ref["temp"][n - 1]), (cabi.F_SALN, 0, ref["saln"][n - 1]),
                 (cabi.F_TH3D, 0, ref["th3d"][n - 1])]
        pairs += [(cabi.F_TRACER, q + 1, ref["tracer"][q, n - 1]) for q in range(ntracr)]
        for fld, ktr, r in pairs:
            dev = ts.download(fld, n, ktr=ktr)[:, nb:nb + g.jj, nb:nb + g.ii]
            assert np.array_equal(dev[:, sea_t], r[glob][:, sea_t]), (g.mproc, g.nproc, fld)
        ts.close()


def _tile_window(a, g1, g, nreg):
    """the padded window of tile g cut out of a single-tile global array (..., nrows, ncols);
    cells beyond the global array are NaN (the exchange or the closed-edge rule fills them)"""
    nb = g.nbdy
    out = np.full(a.shape[:-2] + (g.nrows, g.ncols), np.nan)
    jj0, ii0 = g.j0, g.i0           # global window starts at padded index (j0, i0)
    js = slice(jj0, min(jj0 + g.nrows, g1.nrows))
    is_ = slice(ii0, min(ii0 + g.ncols, g1.ncols))
    out[..., : js.stop - js.start, : is_.stop - is_.start] = a[..., js, is_]
    return out


# ---------------------------------------------------------------------------------------
# btrmas: advem_fct2c (mod_tsadvc.F90:999-1368), five sub-cycled low-order iterations
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,ntracr,extra", [
    (150, 150, 22, 0, 0, {}),             # box basin: three layer batches (8 + 8 + 6)
    (131, 77, 3, 3, 1, {}),               # doubly periodic: sea halo cells, margins 5..0 all matter
    (64, 203, 2, 1, 0, {}),
    (200, 60, 2, 4, 1, {}),
    (58, 31, 10, 0, 1, {"nhybrd": 4}),    # temp advected in the top layers only; batch 8 + 2
    (9, 8, 2, 0, 0, {}),
    (70, 45, 3, 0, 0, {"advflg": 1}),
])
def test_fct2c_host_path_matches_oracle(oracle, itdm, jtdm, kdm, nreg, ntracr, extra):
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=23, m=m, n=n,
                                     advtyp=2, btrmas=True, nstep=3, **extra)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    got, before, launches = _run_host_path(cb, m, n)
    names = ["saln"] + (["th3d"] if cb.advflg else ["temp"])
    _compare(cb, g, got, ref, n, names)
    for q in range(ntracr):
        _compare(cb, g, {"t": got["tracer"][q]}, {"t": ref["tracer"][q]}, n, ["t"])
    msk = util.interior_sea(cb)
    for name in names:
        a, b = got[name], before[name]
        assert np.array_equal(a[m - 1], b[m - 1], equal_nan=True), (name, "slot m modified")
        assert np.array_equal(a[n - 1][:, ~msk], b[n - 1][:, ~msk], equal_nan=True), (name, "land/halo modified")
    assert np.array_equal(got["xmin"], ref["xmin"]) and np.array_equal(got["xmax"], ref["xmax"])
    assert not np.array_equal(got["saln"][n - 1, 0][msk], before["saln"][n - 1, 0][msk])


def test_fct2c_device_path_two_steps(oracle):
    """device-resident mirrors, two leapfrog steps with swapped slots"""
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(90, 70, 3, nreg=0, seed=21, m=m, n=n, advtyp=2, btrmas=True)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    ts = pkg.Tsadvc(cb)
    ts.upload_state(m, n)
    ts.upload(cabi.F_DP, cb.dp[m - 1], m)
    ts.upload(cabi.F_ONETA, cb.oneta[m - 1], m)
    ot.tsadvc(m, n, 1)
    ts.tsadvc_device(m, n)
    ot.tsadvc(n, m, 1)
    ts.tsadvc_device(n, m)
    msk = util.interior_sea(cb)
    for fld, name in ((cabi.F_TEMP, "temp"), (cabi.F_SALN, "saln")):
        for slot in (1, 2):
            dev = ts.download(fld, slot)
            ref = ot.f64(name)[slot - 1]
            for k in range(g.kdm):
                assert np.array_equal(dev[k][msk], ref[k][msk]), (name, slot, k)
    ts.close()
    ot.close()


@pytest.mark.parametrize("ipr,jpr,nreg,ntracr,kdm", [(2, 2, 0, 1, 3), (2, 1, 3, 0, 10), (1, 2, 1, 0, 2)])
def test_fct2c_tiling_invariance_on_device(oracle, ipr, jpr, nreg, ntracr, kdm):
    """advem_fct2c on N tiles, the five in-scheme exchanges done between the tiles of one process,
    == the oracle on 1 tile, bit for bit"""
    xc = __import__("importlib").import_module("hycom-src_b200.xc")
    m, n = 1, 2
    itdm, jtdm = 150, 120
    cfg, sea, g1, cb1 = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=5, m=m, n=n,
                                       advtyp=2, btrmas=True, nstep=3)
    ref = util.run_oracle(oracle, cb1, sea, m, n)
    nb = g1.nbdy
    tss = []
    for g in pkg.partition(itdm, jtdm, kdm, ipr, jpr, nreg):
        cb = syn.build_cb_arrays(cfg, g, sea, m, n, advtyp=2, btrmas=True, nstep=3)
        ts = pkg.Tsadvc(cb)
        ts.upload_state(m, n)
        tss.append(ts)
    _exchange_in_process(tss, m, n)
    bes = [xc.DeviceHaloBackend(ts) for ts in tss]
    nbrs = [xc.neighbors(ts.cb.geom) for ts in tss]
    nbatch, per = C.c_int32(0), C.c_int32(0)
    tss[0]._ck(tss[0].lib.hycom_tsadvc_fct2c_batches(tss[0].h, C.byref(nbatch), C.byref(per)))
    assert nbatch.value == (kdm + per.value - 1) // per.value

    def stage(b, st):
        for ts in tss:
            p = ts.cb.params()
            ts._ck(ts.lib.hycom_tsadvc_fct2c_stage(ts.h, m, n, C.byref(p), b, st))

    for b in range(nbatch.value):
        stage(b, 0)
        for it in range(5):
            stage(b, 1)
            sends = []
            for t, ts in enumerate(tss):
                cnt = bes[t].fct2c_counts(m, n, b)
                layers = min(per.value, kdm - b * per.value)
                assert cnt == xc.halo_counts(ts.cb.geom, (1 + 2 + ntracr) * layers, 5, 5)
                send = [bes[t].alloc(c) if nbrs[t][d] >= 0 else None for d, c in enumerate(cnt)]
                bes[t].fct2c_pack(m, n, b, send)
                sends.append(send)
                ts.synchronize()
            for t, ts in enumerate(tss):
                recv = [sends[nbrs[t][d]][xc.opp_dir(ts.cb.geom, d)] if nbrs[t][d] >= 0 else None for d in range(8)]
                bes[t].fct2c_unpack(m, n, b, recv)
                ts.synchronize()
        stage(b, 2)
    for ts in tss:
        p = ts.cb.params()
        ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_ALL, None, None))
    for ts in tss:
        g = ts.cb.geom
        sea_t = ts.cb.ip[nb:nb + g.jj, nb:nb + g.ii] != 0
        glob = (slice(None), slice(nb + g.j0, nb + g.j0 + g.jj), slice(nb + g.i0, nb + g.i0 + g.ii))
        pairs = [(cabi.F_TEMP, 0, ref["temp"][n - 1]), (cabi.F_SALN, 0, ref["saln"][n - 1])]
        pairs += [(cabi.F_TRACER, q + 1, ref["tracer"][q, n - 1]) for q in range(ntracr)]
        for fld, ktr, r in pairs:
            dev = ts.download(fld, n, ktr=ktr)[:, nb:nb + g.jj, nb:nb + g.ii]
            assert np.array_equal(dev[:, sea_t], r[glob][:, sea_t]), (g.mproc, g.nproc, fld)
        ts.close()


# ---------------------------------------------------------------------------------------
# nreg=2: global grid across the arctic on one tile (tripole fold, mod_xc_sm.h:1172-1335)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("itdm,jtdm,kdm,ntracr,advtyp,extra", [
    (150, 97, 3, 1, 2, {}),
    (131, 77, 2, 0, 1, {}),
    (90, 64, 2, 0, 4, {}),
    (90, 64, 2, 1, 0, {}),
    (120, 70, 3, 0, 2, {"btrmas": True}),
])
def test_arctic_host_path_matches_oracle(oracle, itdm, jtdm, kdm, ntracr, advtyp, extra):
    m, n = 1, 2
    cfg, sea, g, cb = util.make_arctic_case(itdm, jtdm, kdm, ntracr=ntracr, seed=29, m=m, n=n, advtyp=advtyp,
                                            nstep=3, **extra)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    got, before, launches = _run_host_path(cb, m, n)
    _compare(cb, g, got, ref, n, ["saln", "temp"])
    for q in range(ntracr):
        _compare(cb, g, {"t": got["tracer"][q]}, {"t": ref["tracer"][q]}, n, ["t"])
    assert np.array_equal(got["xmin"], ref["xmin"]) and np.array_equal(got["xmax"], ref["xmax"])


def test_arctic_device_halo_matches_xctilr(oracle):
    """the fold per grid type: scalars (halo_ps), uflx (halo_uv), vflx (halo_vv)"""
    m, n = 1, 2
    cfg, sea, g, cb = util.make_arctic_case(64, 50, 2, seed=3, m=m, n=n)
    ts = pkg.Tsadvc(cb)
    ts.upload_state(m, n)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    for fld, name, itype in ((cabi.F_SALN, "saln", 1), (cabi.F_UFLX, "uflx", 13), (cabi.F_VFLX, "vflx", 14)):
        ts._ck(ts.lib.hycom_tsadvc_halo_local(ts.h, fld, 0, 1, 5, 5))
        dev = ts.download(fld, 1)
        a = ot.f64(name)
        a3 = a[0] if a.ndim == 4 else a
        ot.lib.orc_xctilr_type(ot.t, a3.ctypes.data_as(C.c_void_p), 1, g.kdm, 5, 5, itype)
        nb = g.nbdy
        win = (slice(None), slice(nb - 5, nb + g.jj + 5), slice(nb - 5, nb + g.ii + 5))
        assert np.array_equal(dev[win], a3[win], equal_nan=True), name
    ts.close()
    ot.close()


# nreg=2 on several tiles: the tiles of the top row exchange the fold with their twins
# (mod_xc_mp.h:4114-4662)
@pytest.mark.parametrize("ipr,jpr", [(2, 1), (2, 2), (4, 2), (1, 2)])
def test_arctic_tiles_device_exchange_matches_xctilr(oracle, ipr, jpr):
    """device pack/unpack of the folded exchange == the single-tile arctic xctilr on the global array,
    per grid type: scalars (halo_ps), uflx (halo_uv), vflx (halo_vv)"""
    m, n = 1, 2
    cfg, sea, g1, cb1 = util.make_arctic_case(96, 60, 2, seed=3, m=m, n=n)
    ot = util.oracle_tile_from_cb(oracle, cb1, sea)
    nb = g1.nbdy
    ref = {}
    for name, itype in (("saln", 1), ("temp", 1), ("uflx", 13), ("vflx", 14)):
        a = ot.f64(name)
        a3 = a.reshape((-1,) + a.shape[-2:])
        ot.lib.orc_xctilr_type(ot.t, a3.ctypes.data_as(C.c_void_p), 1, a3.shape[0], 5, 5, itype)
        ref[name] = a.copy()
    ot.close()
    tss = []
    for cb in util.make_arctic_tiles(cfg, sea, cb1, ipr, jpr, m, n):
        ts = pkg.Tsadvc(cb)
        ts.upload_state(m, n)
        tss.append(ts)
    _exchange_in_process(tss, m, n)
    for ts in tss:
        g = ts.cb.geom
        loc = (slice(None), slice(nb - 5, nb + g.jj + 5), slice(nb - 5, nb + g.ii + 5))
        glb = (slice(None), slice(g.j0 + nb - 5, g.j0 + nb + g.jj + 5), slice(g.i0 + nb - 5, g.i0 + nb + g.ii + 5))
        for fld, name, slots in ((cabi.F_SALN, "saln", (1, 2)), (cabi.F_TEMP, "temp", (1, 2)),
                                 (cabi.F_UFLX, "uflx", (1,)), (cabi.F_VFLX, "vflx", (1,))):
            for slot in slots:
                dev = ts.download(fld, slot)[loc]
                r = ref[name][slot - 1] if ref[name].ndim == 4 else ref[name]
                assert np.array_equal(dev, r[glb]), (name, slot, g.mproc, g.nproc)
        ts.close()


@pytest.mark.parametrize("itdm,jtdm,kdm,ipr,jpr,ntracr,advtyp,temdf2", [
    (128, 70, 2, 2, 2, 0, 2, 0.0),     # FCT2 on 2x2 tiles, fold in the top row
    (192, 64, 2, 4, 2, 1, 2, 0.0),     # 4x2 (the 8-GPU tiling): the shifted u-grid column comes from a third tile
    (128, 70, 2, 2, 1, 0, 1, 0.0),     # MPDATA, 2x1: NW/NE fold onto the tile itself
    (96, 80, 2, 1, 2, 0, 2, 0.0),      # 1x2: the top tile is its own twin
    (128, 70, 2, 2, 2, 1, 2, 0.02),    # with the diffusion exchange (width 2) and the EOS tail
])
def test_arctic_tiling_invariance_on_device(oracle, itdm, jtdm, kdm, ipr, jpr, ntracr, advtyp, temdf2):
    """tsadvc on N tiles of a global grid across the arctic == the oracle on the single global tile"""
    m, n = 1, 2
    xc = __import__("importlib").import_module("hycom-src_b200.xc")
    extra = dict(advtyp=advtyp, nstep=3)
    if temdf2 > 0.0:    # temp & saln diffused, th3d = sig(T,S) - thbase from the 17-term sigma-2 fit
        extra.update(temdf2=temdf2, temdfc=1.0, sigver=6, thbase=34.0)
    cfg, sea, g1, cb1 = util.make_arctic_case(itdm, jtdm, kdm, ntracr=ntracr, seed=29, m=m, n=n, **extra)
    ref = util.run_oracle(oracle, cb1, sea, m, n)
    tss = []
    nb = g1.nbdy
    for cb in util.make_arctic_tiles(cfg, sea, cb1, ipr, jpr, m, n, **extra):
        ts = pkg.Tsadvc(cb)
        ts.upload_state(m, n)
        tss.append(ts)
    _exchange_in_process(tss, m, n)
    for ts in tss:
        p = ts.cb.params()
        xm, xx = ts.xmin.ctypes.data_as(C.c_void_p), ts.xmax.ctypes.data_as(C.c_void_p)
        ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_INTERIOR, None, None))
        ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_FRAME, xm, xx))
        ts.synchronize()
    if temdf2 > 0.0:
        sends, nbrs = [], []
        for ts in tss:
            be = xc.DeviceHaloBackend(ts)
            nbr = xc.neighbors(ts.cb.geom)
            cnt = be.diff_counts(n)
            send = [be.alloc(c) if nbr[d] >= 0 else None for d, c in enumerate(cnt)]
            be.diff_pack(n, send)
            sends.append(send); nbrs.append(nbr)
            ts.synchronize()
        for t, ts in enumerate(tss):
            recv = [sends[nbrs[t][d]][xc.opp_dir(ts.cb.geom, d)] if nbrs[t][d] >= 0 else None for d in range(8)]
            xc.DeviceHaloBackend(ts).diff_unpack(n, recv)
            p = ts.cb.params()
            ts._ck(ts.lib.hycom_tsadvc_diffuse_device(ts.h, m, n, C.byref(p)))
            ts.synchronize()
    for ts in tss:
        g = ts.cb.geom
        sea_t = ts.cb.ip[nb:nb + g.jj, nb:nb + g.ii] != 0
        glob = (slice(None), slice(nb + g.j0, nb + g.j0 + g.jj), slice(nb + g.i0, nb + g.i0 + g.ii))
        pairs = [(cabi.F_TEMP, 0, ref["temp"][n - 1]), (cabi.F_SALN, 0, ref["saln"][n - 1])]
        pairs += [(cabi.F_TRACER, q + 1, ref["tracer"][q, n - 1]) for q in range(ntracr)]
        if temdf2 > 0.0:
            pairs.append((cabi.F_TH3D, 0, ref["th3d"][n - 1]))
        for fld, ktr, r in pairs:
            dev = ts.download(fld, n, ktr=ktr)[:, nb:nb + g.jj, nb:nb + g.ii]
            assert np.array_equal(dev[:, sea_t], r[glob][:, sea_t]), (g.mproc, g.nproc, fld)
        ts.close()


# ---------------------------------------------------------------------------------------
# mxlmy: q2, q2l (layers 0..kk+1) advected and diffused with the other fields
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("advtyp,nreg,diff,extra", [(2, 0, False, {}), (1, 3, False, {}), (2, 0, True, {}),
                                                    (2, 1, False, {"btrmas": True}), (4, 0, False, {"nhybrd": 2})])
def test_mxlmy_host_path_matches_oracle(oracle, advtyp, nreg, diff, extra):
    m, n = 1, 2
    kdm = 5
    if diff:
        cfg, sea, g, cb = util.make_diffusion_case(90, 61, kdm, 6, 1.0, nreg=nreg, ntracr=1, seed=31, advtyp=advtyp,
                                                   nstep=3, **extra)
    else:
        cfg, sea, g, cb = util.make_case(90, 61, kdm, nreg=nreg, ntracr=1, seed=31, m=m, n=n, advtyp=advtyp,
                                         nstep=3, **extra)
    util.add_q2(cfg, sea, g, cb, m, n)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    q2_before = cb.q2.copy()
    got, before, launches = _run_host_path(cb, m, n)
    _compare(cb, g, got, ref, n, ["saln", "temp"])
    msk = util.interior_sea(cb)
    for name in ("q2", "q2l"):
        a, b = getattr(cb, name), ref[name]
        for k in range(kdm + 2):
            assert np.array_equal(a[n - 1, k][msk], b[n - 1, k][msk]), (name, k)
        assert np.array_equal(a[m - 1], (q2_before if name == "q2" else a)[m - 1], equal_nan=True)
    assert not np.array_equal(cb.q2[n - 1, 1][msk], q2_before[n - 1, 1][msk])


@pytest.mark.parametrize("advtyp,nreg,temdf2", [(2, 0, 0.0), (1, 3, 0.0), (0, 1, 0.0), (4, 0, 0.0), (2, 0, 0.02)])
def test_isopyc_host_path_matches_oracle(oracle, advtyp, nreg, temdf2):
    """isopycnic coordinates: th3d & saln in layer 1 on smoothed fluxes, saln only below"""
    m, n = 1, 2
    if temdf2 > 0:
        cfg, sea, g, cb = util.make_diffusion_case(90, 61, 4, 8, 0.0, nreg=nreg, seed=37, advtyp=advtyp, nstep=3,
                                                   isopyc=True, hybrid=False, nhybrd=0)
    else:
        cfg, sea, g, cb = util.make_case(90, 61, 4, nreg=nreg, seed=37, m=m, n=n, advtyp=advtyp, nstep=3,
                                         isopyc=True, hybrid=False, nhybrd=0)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    got, before, launches = _run_host_path(cb, m, n)
    _compare(cb, g, got, ref, n, ["saln", "th3d"] + (["temp"] if temdf2 > 0 else []))
    if temdf2 == 0:
        assert np.array_equal(got["temp"], before["temp"], equal_nan=True)
    assert np.array_equal(got["xmin"], ref["xmin"]) and np.array_equal(got["xmax"], ref["xmax"])


@pytest.mark.parametrize("advtyp,nreg,mxlmy,temdf2", [(2, 0, True, 0.0), (1, 3, False, 0.0), (4, 0, False, 0.0),
                                                        (0, 1, True, 0.0), (2, 0, False, 0.02)])
def test_isopyc_with_tracers_matches_oracle(oracle, advtyp, nreg, mxlmy, temdf2):
    """isopyc with tracers / mxlmy (mod_tsadvc.F90:2016-2048): in layer 1 the tracers and q2, q2l are advected
    by uflx, vflx against the fco of the smoothed fluxes (:1930-1932) - the ten-array ring of the march"""
    m, n = 2, 1
    kw = dict(nreg=nreg, seed=41, advtyp=advtyp, nstep=3, isopyc=True, hybrid=False, nhybrd=0, ntracr=2,
              trcflg=[2, 0])
    if temdf2 > 0:
        cfg, sea, g, cb = util.make_diffusion_case(90, 61, 4, 8, 0.0, m=m, n=n, **kw)
    else:
        cfg, sea, g, cb = util.make_case(90, 61, 4, m=m, n=n, **kw)
    if mxlmy:
        util.add_q2(cfg, sea, g, cb, m, n)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    tr_before = cb.tracer.copy()
    got, before, launches = _run_host_path(cb, m, n)
    _compare(cb, g, got, ref, n, ["saln", "th3d"])
    msk = util.interior_sea(cb)
    for q in range(2):
        for k in range(g.kdm):
            assert np.array_equal(cb.tracer[q, n - 1, k][msk], ref["tracer"][q, n - 1, k][msk]), ("tracer", q, k)
    assert not np.array_equal(cb.tracer[1, n - 1, 0][msk], tr_before[1, n - 1, 0][msk])
    if mxlmy:
        for name in ("q2", "q2l"):
            for k in range(1, g.kdm + 1):
                assert np.array_equal(getattr(cb, name)[n - 1, k][msk], ref[name][n - 1, k][msk]), (name, k)
    assert np.array_equal(got["xmin"], ref["xmin"]) and np.array_equal(got["xmax"], ref["xmax"])


def test_isopyc_with_btrmas_is_refused():
    cfg, sea, g, cb = util.make_case(40, 30, 2, advtyp=2, isopyc=True, hybrid=False, nhybrd=0, btrmas=True)
    ts = pkg.Tsadvc(cb)
    ts.upload_state(1, 2)
    with pytest.raises(cabi.TsadvcError, match="isopyc"):
        ts.tsadvc_device(1, 2)
    ts.close()


# ---------------------------------------------------------------------------------------
# frozen vectors (tests/golden/tsadvc_golden.json): the CUDA path reproduces their bits
# ---------------------------------------------------------------------------------------
import json as _json
import os as _os

_GOLD = _json.load(open(_os.path.join(_os.path.dirname(__file__), "golden", "tsadvc_golden.json")))


@pytest.mark.parametrize("name", sorted(_GOLD))
def test_golden_vectors_on_device(name):
    import sys
    sys.path.insert(0, _os.path.join(_os.path.dirname(__file__), "golden"))
    import make_golden
    kind, kw = make_golden.CASES[name]
    cfg, sea, g, cb = make_golden.build(kind, kw)
    got, before, launches = _run_host_path(cb, 1, 2)
    assert launches > 0
    msk = util.interior_sea(cb)
    flds = dict(temp=got["temp"], saln=got["saln"], th3d=got["th3d"], tracer=got["tracer"])
    assert make_golden.digest(flds, msk, 2) == _GOLD[name], name


# ---------------------------------------------------------------------------------------
# next to tsadvc: mod_asselin.F90 on the device mirrors
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("sigver,ntracr,nreg,extra", [(6, 2, 0, {}), (8, 0, 3, {"advflg": 1}), (8, 1, 0, {"nhybrd": 1}),
                                                       (2, 0, 1, {"nhybrd": 2}),
                                                       (7, 0, 0, {"isopyc": True, "hybrid": False, "nhybrd": 0})])
def test_asselin_device_matches_oracle(oracle, sigver, ntracr, nreg, extra):
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(90, 61, 4, nreg=nreg, ntracr=ntracr, seed=41, **extra)
    if sigver == 6:
        util.add_q2(cfg, sea, g, cb, m, n)
    util.add_asselin(cfg, sea, g, cb, m, n, sigver=sigver)
    msk = util.interior_sea(cb)
    flds = [(cabi.F_TEMP, "temp", 0), (cabi.F_SALN, "saln", 0), (cabi.F_TH3D, "th3d", 0), (cabi.F_DP, "dp", 0),
            (cabi.F_ONETA, "oneta", 0)] + [(cabi.F_TRACER, "tracer", q + 1) for q in range(ntracr)]
    if cb.mxlmy:
        flds += [(cabi.F_Q2, "q2", 0), (cabi.F_Q2L, "q2l", 0)]
    libm = sigver <= 4      # tofsig of the 7/9-term fits: atan2/cos/sin, 1e-12 instead of bits

    def ref_of(ot, name, ktr, slot):
        a = ot.f64(name)
        return a[ktr - 1, slot - 1] if name == "tracer" else a[slot - 1]
    # ---- asselin_save
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    ot.asselin_save(m, n, 1)
    ts = pkg.Tsadvc(cb)
    ts.upload_asselin_state(m, n)
    ts.asselin_save_device(m, n)
    for slot in (1, 2):
        for fld, name in ((cabi.F_ONETA, "oneta"), (cabi.F_ONETAO, "onetao")):
            assert np.array_equal(ts.download(fld, slot)[0], ot.f64(name)[slot - 1], equal_nan=True), (name, slot)
    olds = [(cabi.F_OTEMP, "otemp", 0), (cabi.F_OSALN, "osaln", 0), (cabi.F_OTH3D, "oth3d", 0)]
    olds += [(cabi.F_OTRACER, "otracer", q + 1) for q in range(ntracr)]
    if cb.mxlmy:
        olds += [(cabi.F_OQ2, "oq2", 0), (cabi.F_OQ2L, "oq2l", 0)]
    for fld, name, ktr in olds:
        ref = ot.f64(name)[ktr - 1] if name == "otracer" else ot.f64(name)
        assert np.array_equal(ts.download(fld, 1, ktr=ktr)[:, msk], ref[:, msk]), name
    ts.close(); ot.close()
    # ---- asselin_filter from the same starting state
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    ot.asselin_filter(m, n)
    ts = pkg.Tsadvc(cb)
    ts.upload_asselin_state(m, n)
    ts.asselin_filter_device(m, n)
    for fld, name, ktr in flds:
        for slot in (1, 2):
            dev, ref = ts.download(fld, slot, ktr=ktr), ref_of(ot, name, ktr, slot)
            if name == "oneta":
                dev, ref = dev[0], ref
                assert np.array_equal(dev[msk], ref[msk]), (name, slot)
            elif libm and name == "temp":
                assert util.rel_err(dev, ref, msk[None]) <= REL_TOL, (name, slot)
            else:
                assert np.array_equal(dev[:, msk], ref[:, msk]), (name, slot)
    assert not np.array_equal(ts.download(cabi.F_SALN, m)[0][msk], cb.saln[m - 1, 0][msk])
    ts.close(); ot.close()


# ---------------------------------------------------------------------------------------
# the device against the REFERENCE'S OWN SOURCE TEXT: tests/golden/from_reference_text.json holds digests of what
# mod_asselin.F90 computes when executed as written (oracle/fortran_exec.py, tests/golden/make_reference_text_vectors.py)
# on the configurations of test_asselin_device_matches_oracle
# ---------------------------------------------------------------------------------------
import reftext_cases as _rc

_REFTEXT = _json.load(open(_os.path.join(_os.path.dirname(__file__), "golden", "from_reference_text.json")))


@pytest.mark.parametrize("name", sorted(_rc.ASSELIN))
def test_asselin_device_reproduces_the_reference_text(name):
    cfg, sea, g, cb, m, n = _rc.asselin_case(name)
    ts = pkg.Tsadvc(cb)
    ts.upload_asselin_state(m, n)
    ts.asselin_filter_device(m, n)
    flds = {nm: np.stack([ts.download(fld, 1), ts.download(fld, 2)])
            for fld, nm in ((cabi.F_TEMP, "temp"), (cabi.F_SALN, "saln"), (cabi.F_TH3D, "th3d"), (cabi.F_DP, "dp"))}
    if cb.ntracr:
        flds["tracer"] = np.stack([np.stack([ts.download(cabi.F_TRACER, 1, ktr=q + 1), ts.download(cabi.F_TRACER, 2, ktr=q + 1)])
                                   for q in range(cb.ntracr)])
    ts.close()
    assert _rc.asselin_digest(cb, m, flds) == _REFTEXT[name], name


def test_config1_on_device_reproduces_the_reference_text():
    """BASELINE.json configs[0] at full size (150 x 150 x 22 box basin, FCT2 T + S, one tile) through the host-array
    entry against the digest of the executed reference text"""
    import test_reference_text as T
    cfg, sea, g, cb = T.config1_case()
    got, before, launches = _run_host_path(cb, 1, 2)
    assert launches > 0
    assert T.config1_digest(cb, got["temp"], got["saln"]) == _REFTEXT[T.CONFIG1]


@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,ntracr,advtyp,extra", CASES + [(90, 64, 2, 2, 1, 2, {}), (90, 61, 3, 4, 1, 2, {"btrmas": True})])
def test_tsadvc_host_path_matches_the_compiled_reference_text(itdm, jtdm, kdm, nreg, ntracr, advtyp, extra):
    """the device against the reference text compiled (oracle/_ref), nothing in between: every scheme and driver option
    of CASES, across the arctic, and advem_fct2c"""
    import copy
    m, n = 1, 2
    if nreg == 2:
        cfg, sea, g, cb = util.make_arctic_case(itdm, jtdm, kdm, ntracr=ntracr, seed=13, m=m, n=n, advtyp=advtyp, nstep=3, **extra)
    else:
        cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=13, m=m, n=n, advtyp=advtyp, nstep=3, **extra)
    cbt = copy.deepcopy(cb)
    if not util.run_compiled_reference_text(cbt, sea, m, n):
        pytest.skip("oracle/_ref holds no compiled reference text on this machine")
    got, before, launches = _run_host_path(cb, m, n)
    assert launches > 0
    msk = util.interior_sea(cb)
    for name in ("temp", "saln", "th3d"):
        assert np.array_equal(got[name][n - 1][:, msk], getattr(cbt, name)[n - 1][:, msk]), name
    for q in range(ntracr):
        assert np.array_equal(got["tracer"][q, n - 1][:, msk], cbt.tracer[q, n - 1][:, msk]), ("tracer", q)
    assert not np.array_equal(got["saln"][n - 1, 0][msk], before["saln"][n - 1, 0][msk])
