"""The CPU oracle against the REFERENCE'S OWN SOURCE TEXT, executed here without a Fortran compiler:
oracle/fortran_exec.py translates the structured Fortran of /root/reference/bigrid.F90 and mod_tsadvc.F90
statement by statement into Python (it knows the language, not the algorithms) and runs it in IEEE double
precision; the oracle (oracle/tsadvc_oracle.c) must reproduce every result bit for bit.

  * bigrid: ip, iu, iv, iq, the sea-only neighbour indices and the segment tables ifp/ilp/isp, jfp/jlp/jsp
  * advem_pcm, advem_mpdata, advem_fct2, advem_fct4 on small basins with islands and on periodic domains

The live tests need /root/reference (present in the build container, absent on the GPU box: they skip there);
tests/golden/from_reference_text.json keeps the digests of the same runs (make_reference_text_vectors.py), which
the oracle must reproduce everywhere."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reference_text as rt  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "from_reference_text.json")

# itdm, jtdm, nreg, seed
GRIDS = [(26, 22, 0, 3), (24, 20, 1, 5), (22, 24, 3, 7), (20, 26, 4, 9)]
SCHEMES = [0, 1, 2, 4]


def build_case(itdm, jtdm, nreg, seed):
    """a small case of the usual generator + its depth array (100 m on sea) as bigrid wants it"""
    cfg, sea, g, cb = util.make_case(itdm, jtdm, 2, nreg=nreg, seed=seed)
    nb = g.nbdy
    depth = np.zeros((g.nrows, g.ncols))
    depth[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea != 0, 100.0, 0.0)
    return cfg, sea, g, cb, depth


def oracle_advem(ot, advtyp, fld, fldc, u, v, fco, fcn, posdef, scal, scali, dt2):
    P = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    rc = ot.lib.orc_advem(ot.t, advtyp, P(fld), P(fldc), P(u), P(v), P(fco), P(fcn), posdef, P(scal), P(scali), dt2, 0)
    assert rc == 0
    return fld


def advem_inputs(g, cb, seed):
    """operands of one advem call with valid halos (width nbdy): the case's fields of layer 1, halo-refreshed the
    way tsadvc does before it calls advem, and the prolog's fco, fcn"""
    import np_restatement as npr
    nb = g.nbdy
    H = lambda a, it=1: npr.halo_single_tile(g, np.array(a, dtype=np.float64), nb, nb, it)   # noqa: E731
    fld = H(cb.saln[1, 0])
    fldc = H(cb.saln[0, 0])
    u, v = H(cb.uflx[0], 13), H(cb.vflx[0], 14)
    dp = np.nan_to_num(H(cb.dp[1, 0]))
    with np.errstate(all="ignore"):
        flxdiv = ((np.roll(u, -1, 1) - u) + (np.roll(v, -1, 0) - v)) * cb.delt1 * cb.scp2i
    fco = np.maximum(dp + np.nan_to_num(flxdiv), 0.0)
    fcn = np.maximum(dp, 0.0)
    clean = lambda a: np.ascontiguousarray(np.nan_to_num(a))   # noqa: E731
    return [clean(x) for x in (fld, fldc, u, v, fco, fcn)] + [np.ascontiguousarray(cb.scp2), np.ascontiguousarray(cb.scp2i)]


def digest(a, mask):
    return hashlib.sha256(np.ascontiguousarray(a[mask]).astype("<f8").tobytes()).hexdigest()[:24]


@pytest.mark.skipif(not rt.available(), reason="the reference source tree is not on this machine")
@pytest.mark.parametrize("itdm,jtdm,nreg,seed", GRIDS)
def test_bigrid_of_the_reference_text_equals_oracle(oracle, itdm, jtdm, nreg, seed):
    cfg, sea, g, cb, depth = build_case(itdm, jtdm, nreg, seed)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    env = rt.make_env(g.ii, g.jj)
    rt.run_bigrid(env, depth.copy(), mapflg=4 if nreg in (3, 4) else 0)
    assert env["nreg"] == nreg
    nb = g.nbdy
    for name in ("ip", "iu", "iv", "iq"):
        assert np.array_equal(env[name].a, ot.i32(name)), name
    inner = np.zeros((g.nrows, g.ncols), dtype=bool)
    inner[1:-1, 1:-1] = True         # the neighbour indices are defined one line inside the array (bigrid.F90:318)
    jj_, ii_ = np.meshgrid(np.arange(g.nrows) - nb + 1, np.arange(g.ncols) - nb + 1, indexing="ij")
    for name in ("ipim1", "ipip1", "ipjm1", "ipjp1"):
        assert np.array_equal(env[name].a[inner], ot.i32(name)[inner]), name
    # segment tables: the oracle stores (ms, nrows) / (ms, ncols)
    ms = min(env["ms"], ot.get_i("ms"))
    assert np.array_equal(env["isp"].a, ot.i32("isp")) and np.array_equal(env["jsp"].a, ot.i32("jsp"))
    assert env["isp"].a.max() <= ms and env["jsp"].a.max() <= ms
    for a, b in (("ifp", "ifp"), ("ilp", "ilp"), ("jfp", "jfp"), ("jlp", "jlp")):
        assert np.array_equal(env[a].a[:ms], ot.i32(b)[:ms]), a
    ot.close()


@pytest.mark.skipif(not rt.available(), reason="the reference source tree is not on this machine")
@pytest.mark.parametrize("advtyp", SCHEMES)
@pytest.mark.parametrize("itdm,jtdm,nreg,seed", GRIDS)
def test_advem_of_the_reference_text_equals_oracle(oracle, itdm, jtdm, nreg, seed, advtyp):
    cfg, sea, g, cb, depth = build_case(itdm, jtdm, nreg, seed)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    env = rt.make_env(g.ii, g.jj)
    rt.run_bigrid(env, depth.copy(), mapflg=4 if nreg in (3, 4) else 0)
    ops = advem_inputs(g, cb, seed)
    posdef = 0.0 if advtyp != 1 else 256.0
    want = rt.run_advem(env, advtyp, ops[0].copy(), *[o.copy() for o in ops[1:6]], posdef, ops[6], ops[7], cb.delt1)
    got = oracle_advem(ot, advtyp, ops[0].copy(), *[o.copy() for o in ops[1:6]], posdef, ops[6].copy(), ops[7].copy(), cb.delt1)
    inner = util.interior_sea(cb)
    assert np.isfinite(want[inner]).all()
    assert not np.array_equal(want[inner], ops[0][inner])          # the field moved
    assert np.array_equal(got[inner], want[inner]), np.abs(got - want)[inner].max()
    gold = json.load(open(GOLD)) if os.path.exists(GOLD) else {}
    key = f"advem{advtyp}:{itdm}x{jtdm}:nreg{nreg}:seed{seed}"
    if key in gold:
        assert gold[key] == digest(want, inner), key
    ot.close()


def test_oracle_reproduces_the_digests_of_the_reference_text(oracle):
    """everywhere (also where the reference tree is absent): the committed digests of the runs above"""
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/from_reference_text.json has not been generated")
    gold = json.load(open(GOLD))
    assert len(gold) >= len(GRIDS) * len(SCHEMES)
    for itdm, jtdm, nreg, seed in GRIDS:
        cfg, sea, g, cb, depth = build_case(itdm, jtdm, nreg, seed)
        ot = util.oracle_tile_from_cb(oracle, cb, sea)
        ops = advem_inputs(g, cb, seed)
        inner = util.interior_sea(cb)
        for advtyp in SCHEMES:
            posdef = 0.0 if advtyp != 1 else 256.0
            got = oracle_advem(ot, advtyp, ops[0].copy(), *[o.copy() for o in ops[1:6]], posdef, ops[6].copy(),
                               ops[7].copy(), cb.delt1)
            key = f"advem{advtyp}:{itdm}x{jtdm}:nreg{nreg}:seed{seed}"
            assert gold[key] == digest(got, inner), key
        ot.close()


# ---------------------------------------------------------------------------------------------------------
# the driver tsadvc(m,n) itself (mod_tsadvc.F90:1708-2258): halo refresh, prolog, field selection, advem dispatch,
# salinity range, diffusion (internal procedures tsdff_1x / tsdff_2x) and the equation-of-state sweep
# (statement functions of stmt_fns.h) - the reference text against the oracle's orc_tsadvc
# ---------------------------------------------------------------------------------------------------------
DRIVER_CASES = [
    # itdm, jtdm, kdm, nreg, ntracr, advtyp, extra
    (26, 22, 3, 0, 1, 2, {}),                                    # FCT2, T + S + tracer, closed basin
    (24, 20, 2, 1, 2, 1, {"trcflg": [0, 2]}),                    # MPDATA, a temperature tracer, periodic in i
    (22, 24, 2, 3, 0, 4, {}),                                    # FCT4, doubly periodic
    (20, 26, 2, 4, 1, 0, {}),                                    # PCM
    (26, 22, 4, 0, 1, 2, {"nhybrd": 2}),                         # isopycnal layers below nhybrd: saln only
    (24, 20, 3, 1, 0, 2, {"advflg": 1}),                         # advect th3d and S
    (26, 22, 3, 0, 0, 2, {"isopyc": True, "hybrid": False, "nhybrd": 0}),   # layer 1 on smoothed fluxes
    (26, 22, 3, 0, 2, 1, {"isopyc": True, "hybrid": False, "nhybrd": 0}),   # ... with tracers on uflx, vflx
    (26, 22, 2, 0, 1, 2, {"btrmas": True}),                      # advem_fct2c: five sub-cycled iterations, onetamas
    (24, 20, 2, 1, 0, 2, {"btrmas": True}),
    (28, 24, 2, 2, 1, 2, {}),                                    # across the arctic: xctilr of mod_xc_sm.h with ARCTIC
    (28, 24, 2, 2, 0, 1, {}),
]


def _run_reference_driver(cb, sea, g, m, n, sigver=6):
    nb = g.nbdy
    depth = np.zeros((g.nrows, g.ncols))
    depth[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea != 0, 100.0, 0.0)
    env = rt.make_env(g.ii, g.jj, g.kdm)
    rt.run_bigrid(env, depth, mapflg=4 if g.nreg in (3, 4) else 0)
    rt.add_cb_arrays(env, cb)
    rt.compile_tsadvc(env, sigver)
    rt.run_tsadvc(env, m, n)
    return env


def _same(a, b, mask):
    return np.array_equal(a[..., mask], b[..., mask])


@pytest.mark.skipif(not rt.available(), reason="the reference source tree is not on this machine")
@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,ntracr,advtyp,extra", DRIVER_CASES)
def test_tsadvc_of_the_reference_text_equals_oracle(oracle, itdm, jtdm, kdm, nreg, ntracr, advtyp, extra):
    m, n = 1, 2
    if nreg == 2:
        cfg, sea, g, cb = util.make_arctic_case(itdm, jtdm, kdm, ntracr=ntracr, seed=11, m=m, n=n, advtyp=advtyp, nstep=3,
                                                **extra)
    else:
        cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=11, m=m, n=n, advtyp=advtyp,
                                         nstep=3, **extra)
    ref = util.run_oracle(oracle, cb, sea, m, n)          # the oracle works on its own copy
    before = cb.saln.copy()
    env = _run_reference_driver(cb, sea, g, m, n)          # the reference text updates cb in place
    inner = util.interior_sea(cb)
    for name in ("temp", "saln", "th3d"):
        assert _same(getattr(cb, name)[n - 1], ref[name][n - 1], inner), name
    for q in range(ntracr):
        assert _same(cb.tracer[q, n - 1], ref["tracer"][q, n - 1], inner), ("tracer", q)
    assert np.array_equal(env["xmin"].a, ref["xmin"]) and np.array_equal(env["xmax"].a, ref["xmax"])
    assert not _same(cb.saln[n - 1], before[n - 1], inner)


# diffusion + equation of state (temdf2 > 0): tsdff_1x / tsdff_2x and the sig / tofsig statement functions of the
# family the executable is compiled with (cpp: EOS_SIG0|EOS_SIG2 x EOS_7T|9T|12T|17T); mxlmy adds q2, q2l
DIFF_CASES = [
    # sigver, temdfc, nreg, ntracr, nhybrd, mxlmy
    (6, 1.0, 0, 1, -1, False),     # 17-term sigma-2, temp diffused, th3d from sig(t,s)
    (8, 0.5, 0, 0, -1, False),     # 12-term sigma-2: temp and th3d combined in density space, tofsig
    (7, 0.0, 1, 2, -1, False),     # 12-term sigma-0: th3d diffused, temp from tofsig
    (2, 1.0, 3, 0, 2, False),      # 7-term sigma-2, exactly isopycnal layers below nhybrd: th3d = theta
    (4, 1.0, 0, 1, -1, True),      # 9-term sigma-2 with Mellor-Yamada fields
    (1, 0.5, 0, 0, -1, False),     # 7-term sigma-0, temp and th3d combined: sig and tofsig
    (3, 0.0, 1, 0, -1, False),     # 9-term sigma-0, th3d diffused: tofsig
    (5, 1.0, 4, 1, -1, False),     # 17-term sigma-0
]


@pytest.mark.skipif(not rt.available(), reason="the reference source tree is not on this machine")
@pytest.mark.parametrize("sigver,temdfc,nreg,ntracr,nhybrd,mxlmy", DIFF_CASES)
def test_tsadvc_with_diffusion_of_the_reference_text_equals_oracle(oracle, sigver, temdfc, nreg, ntracr, nhybrd, mxlmy):
    m, n = 1, 2
    itdm, jtdm, kdm = 24, 20, 3
    cfg, sea, g, cb = util.make_diffusion_case(itdm, jtdm, kdm, sigver, temdfc, nreg=nreg, ntracr=ntracr, nhybrd=nhybrd,
                                               seed=13, nstep=3, m=m, n=n)
    if mxlmy:
        util.add_q2(cfg, sea, g, cb, m, n)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    env = _run_reference_driver(cb, sea, g, m, n, sigver)
    inner = util.interior_sea(cb)
    # sig of every family and tofsig of the 12-term family are bit-exact; tofsig of the 7/9-term families goes through
    # atan2 / cos, whose last bit depends on the maths library (python's libm here, the C library in the oracle)
    libm = sigver <= 4 and (temdfc < 1.0 or (0 <= nhybrd < kdm))
    for name in ("temp", "saln", "th3d"):
        a, b = getattr(cb, name)[n - 1], ref[name][n - 1]
        if libm and name != "saln":
            assert np.allclose(a[..., inner], b[..., inner], rtol=1e-12, atol=0), name
        else:
            assert _same(a, b, inner), name
    for q in range(ntracr):
        assert _same(cb.tracer[q, n - 1], ref["tracer"][q, n - 1], inner), ("tracer", q)
    if mxlmy:
        assert _same(cb.q2[n - 1, 1:-1], ref["q2"][n - 1, 1:-1], inner) and _same(cb.q2l[n - 1, 1:-1], ref["q2l"][n - 1, 1:-1], inner)


# mod_asselin.F90: asselin_save and asselin_filter of the reference text
ASSELIN_CASES = [(6, 2, {}, True), (8, 0, {"advflg": 1}, False), (2, 1, {"nhybrd": 1}, False),
                 (7, 0, {"isopyc": True, "hybrid": False, "nhybrd": 0}, False)]


@pytest.mark.skipif(not rt.available(), reason="the reference source tree is not on this machine")
@pytest.mark.parametrize("sigver,ntracr,extra,mxlmy", ASSELIN_CASES)
def test_asselin_of_the_reference_text_equals_oracle(oracle, sigver, ntracr, extra, mxlmy):
    import copy
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(26, 22, 3, nreg=0, ntracr=ntracr, seed=5, **extra)
    if mxlmy:
        util.add_q2(cfg, sea, g, cb, m, n)
    util.add_asselin(cfg, sea, g, cb, m, n, sigver=sigver)
    msk = util.interior_sea(cb)
    nb = g.nbdy
    depth = np.zeros((g.nrows, g.ncols))
    depth[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea != 0, 100.0, 0.0)

    def reference(which):
        c = copy.deepcopy(cb)
        env = rt.make_env(g.ii, g.jj, g.kdm)
        rt.run_bigrid(env, depth.copy())
        rt.add_cb_arrays(env, c)
        rt.add_asselin_arrays(env, c)
        rt.compile_asselin(env, sigver)
        env[which](m, n)
        return c, env

    # asselin_save
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    ot.asselin_save(m, n, 1)
    c, env = reference("asselin_save")
    for name in ("oneta", "onetao"):
        assert np.array_equal(ot.f64(name), getattr(c, name), equal_nan=True), name
    for name in ("otemp", "osaln", "oth3d") + (("otracer",) if ntracr else ()) + (("oq2", "oq2l") if mxlmy else ()):
        a, b = ot.f64(name), getattr(c, name)
        if name in ("oq2", "oq2l"):
            a, b = a[1:-1], b[1:-1]
        assert np.array_equal(a[..., msk], b[..., msk]), name
    ot.close()
    # asselin_filter from the same starting state
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    ot.asselin_filter(m, n)
    c, env = reference("asselin_filter")
    libm = sigver <= 4 and ("nhybrd" in extra or extra.get("advflg") == 1 or extra.get("isopyc"))
    for name in ("oneta", "dp", "temp", "saln", "th3d") + (("tracer",) if ntracr else ()) + (("q2", "q2l") if mxlmy else ()):
        a, b = ot.f64(name), getattr(c, name)
        if name in ("q2", "q2l"):
            a, b = a[:, 1:-1], b[:, 1:-1]
        if libm and name == "temp":
            assert util.rel_err(a[m - 1], b[m - 1], msk) < 1e-13, name
        else:
            assert np.array_equal(a[..., msk], b[..., msk], equal_nan=True), name
    assert not np.array_equal(c.saln[m - 1, 0][msk], cb.saln[m - 1, 0][msk])
    ot.close()


# cnuity(m,n) of cnuity.F90: loop 76 (FCT of dp), loop 77 + bottom restoring, the
# interface-depth diffusion (thkdf4 in both sweep directions, thkdf2), mxlkta, the Asselin tail
CNUITY_CASES = [
    # itdm, jtdm, kdm, nreg, m, n, isopyc, thkdf, bih, nstep, mxlkta
    (30, 24, 4, 0, 1, 2, False, 0.0, True, 3, False),      # closed basin with islands
    (26, 22, 3, 3, 2, 1, False, 0.0, True, 3, False),      # doubly periodic, slots swapped
    (24, 28, 3, 1, 1, 2, True, 0.0, True, 3, False),       # periodic in i, isopyc: dpmixl(n) = dp(1,n)
    (26, 20, 2, 4, 1, 2, False, 0.0, True, 3, False),      # closed f-plane (periodic in j)
    (30, 24, 5, 0, 1, 2, False, 0.01, True, 4, False),     # biharmonic interface-depth diffusion, downward sweep
    (30, 24, 5, 0, 1, 2, False, 0.01, True, 7, False),     # ... upward sweep
    (24, 28, 4, 1, 1, 2, True, 0.01, True, 3, True),       # ... with isopyc (mxlkta is then irrelevant)
    (26, 22, 3, 3, 1, 2, False, 0.02, False, 2, False),    # Laplacian
    (30, 24, 5, 0, 1, 2, False, 0.01, True, 4, True),      # hybrid .and. mxlkta: dpmixl follows the coordinates
    (30, 26, 3, 2, 1, 2, False, 0.0, True, 3, False),      # across the arctic
    (30, 26, 4, 2, 1, 2, False, 0.01, True, 4, False),     # ... with the interface-depth diffusion
]


@pytest.mark.skipif(not rt.available(), reason="the reference source tree is not on this machine")
@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,m,n,isopyc,thkdf,bih,nstep,mxlkta", CNUITY_CASES)
def test_cnuity_of_the_reference_text_equals_oracle(oracle, itdm, jtdm, kdm, nreg, m, n, isopyc, thkdf, bih, nstep, mxlkta):
    extra = dict(isopyc=True, hybrid=False, nhybrd=0) if isopyc else {}
    if nreg == 2:
        cfg, sea, g, cb = util.make_arctic_case(itdm, jtdm, kdm, seed=23, m=m, n=n, nstep=nstep, **extra)
        st = util.arctic_halos_cnuity(g, util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=thkdf, bih=bih))
    else:
        cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=23, m=m, n=n, nstep=nstep, **extra)
        st = util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=thkdf, bih=bih)
    if mxlkta:
        util.deepen_dpmixl(st, n)
    got = util.run_oracle_cnuity(oracle, cb, sea, st, m, n, isopyc=isopyc, mxlkta=mxlkta)     # on its own copy
    before = st["dp"].copy()
    nb = g.nbdy
    depth = np.zeros((g.nrows, g.ncols))
    depth[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea != 0, 100.0, 0.0)
    env = rt.make_env(g.ii, g.jj, g.kdm)
    rt.run_bigrid(env, depth, mapflg=4 if nreg in (3, 4) else 0)
    assert env["nreg"] == nreg
    rt.add_cb_arrays(env, cb)
    rt.add_cnuity_arrays(env, cb, st, mxlkta=mxlkta)
    rt.compile_cnuity(env)
    env["cnuity"](m, n)                                                                       # updates st in place
    inner = util.interior_sea(cb)
    sea6 = cb.ip != 0
    iu_in, iv_in = np.zeros_like(inner), np.zeros_like(inner)
    iu_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iu[nb:nb + g.jj, nb:nb + g.ii] != 0
    iv_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iv[nb:nb + g.jj, nb:nb + g.ii] != 0
    for k in range(kdm):
        assert np.array_equal(got["dp"][n - 1, k][inner], st["dp"][n - 1, k][inner]), ("dp.n", k)
        assert np.array_equal(got["dp"][m - 1, k][inner], st["dp"][m - 1, k][inner]), ("dp.m", k)
        assert np.array_equal(got["dpo"][m - 1, k][inner], st["dpo"][m - 1, k][inner]), ("dpo.m", k)
        assert np.array_equal(got["dpo"][n - 1, k][inner], st["dpo"][n - 1, k][inner]), ("dpo.n", k)
        assert np.array_equal(got["uflx"][k][iu_in], st["uflx"][k][iu_in]), ("uflx", k)
        assert np.array_equal(got["vflx"][k][iv_in], st["vflx"][k][iv_in]), ("vflx", k)
        assert np.array_equal(got["p"][k + 1][inner], env["p"].a[k + 1][inner]), ("p", k)
        assert np.array_equal(got["dpav"][k][inner], st["dpav"][k][inner]), ("dpav", k)
        assert np.array_equal(got["uflxav"][k][iu_in], st["uflxav"][k][iu_in]), ("uflxav", k)
        assert np.array_equal(got["vflxav"][k][iv_in], st["vflxav"][k][iv_in]), ("vflxav", k)
    assert np.array_equal(got["utotn"][iu_in], env["utotn"].a[iu_in])
    assert np.array_equal(got["vtotn"][iv_in], env["vtotn"].a[iv_in])
    assert np.array_equal(got["dpmixl"][..., inner], st["dpmixl"][..., inner])
    assert np.array_equal(got["dpmold"][inner], env["dpmold"].a[inner])
    assert not np.array_equal(before[n - 1, 0][inner], st["dp"][n - 1, 0][inner])
    if thkdf:
        assert not np.array_equal(before[n - 1, kdm - 1][inner], st["dp"][n - 1, kdm - 1][inner])


# ---------------------------------------------------------------------------------------------------------
# the frozen vectors of tests/golden/tsadvc_golden.json - the ones the CUDA path must reproduce in
# tests/test_parity_gpu.py::test_golden_vectors_on_device - are what the REFERENCE TEXT computes
# ---------------------------------------------------------------------------------------------------------
def _golden_cases():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    return make_golden


def reference_text_digest_of_golden_case(name):
    mg = _golden_cases()
    kind, kw = mg.CASES[name]
    cfg, sea, g, cb = mg.build(kind, kw)
    _run_reference_driver(cb, sea, g, 1, 2, int(getattr(cb, "sigver", 6)))     # updates cb in place
    flds = dict(temp=cb.temp, saln=cb.saln, th3d=cb.th3d, tracer=cb.tracer if cb.ntracr else None)
    return mg.digest(flds, util.interior_sea(cb), 2)


# BASELINE.json configs[0] at its full size: the 150 x 150 x 22 box basin, FCT2 T + S, one tile - the first case of
# tests/test_parity_gpu.py::CASES (75 s of executed reference text: generated once, not run live)
CONFIG1 = "config1:box_150x150x22_fct2"


def config1_case():
    m, n = 1, 2
    return util.make_case(150, 150, 22, nreg=0, ntracr=0, seed=13, m=m, n=n, advtyp=2, nstep=3)


def config1_digest(cb, temp, saln):
    mg = _golden_cases()
    return mg.digest(dict(temp=temp, saln=saln), util.interior_sea(cb), 2)


def reference_text_digest_of_config1():
    cfg, sea, g, cb = config1_case()
    _run_reference_driver(cb, sea, g, 1, 2)
    return config1_digest(cb, cb.temp, cb.saln)


def test_oracle_reproduces_config1_of_the_reference_text(oracle):
    gold = json.load(open(GOLD))
    cfg, sea, g, cb = config1_case()
    ref = util.run_oracle(oracle, cb, sea, 1, 2)
    assert config1_digest(cb, ref["temp"], ref["saln"]) == gold[CONFIG1]


def _reference_asselin_filter(name):
    import reftext_cases as rc
    cfg, sea, g, cb, m, n = rc.asselin_case(name)
    nb = g.nbdy
    depth = np.zeros((g.nrows, g.ncols))
    depth[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea != 0, 100.0, 0.0)
    env = rt.make_env(g.ii, g.jj, g.kdm)
    rt.run_bigrid(env, depth, mapflg=4 if g.nreg in (3, 4) else 0)
    rt.add_cb_arrays(env, cb)
    rt.add_asselin_arrays(env, cb)
    rt.compile_asselin(env, int(cb.sigver))
    env["asselin_filter"](m, n)
    return rc.asselin_digest(cb, m, dict(temp=cb.temp, saln=cb.saln, th3d=cb.th3d, dp=cb.dp, tracer=cb.tracer))


def _reference_cnuity(name):
    import reftext_cases as rc
    cfg, sea, g, cb, st, m, n, isopyc, mxlkta = rc.cnuity_case(name)
    nb = g.nbdy
    depth = np.zeros((g.nrows, g.ncols))
    depth[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea != 0, 100.0, 0.0)
    env = rt.make_env(g.ii, g.jj, g.kdm)
    rt.run_bigrid(env, depth, mapflg=4 if g.nreg in (3, 4) else 0)
    rt.add_cb_arrays(env, cb)
    rt.add_cnuity_arrays(env, cb, st, mxlkta=mxlkta)
    rt.compile_cnuity(env)
    env["cnuity"](m, n)
    return rc.cnuity_digest(cb, m, n, st["dp"], st["uflx"], st["vflx"], env["p"].a,
                            st["dpmixl"][n - 1] if (isopyc or mxlkta) else None)


def more_reference_vectors():
    """for tests/golden/make_reference_text_vectors.py: digests of the reference text on the frozen-vector cases of
    tsadvc and on the asselin / cnuity cases of tests/reftext_cases.py"""
    import reftext_cases as rc
    out = {"golden:" + name: reference_text_digest_of_golden_case(name) for name in sorted(_golden_cases().CASES)}
    out[CONFIG1] = reference_text_digest_of_config1()
    out.update({name: _reference_asselin_filter(name) for name in rc.ASSELIN})
    out.update({name: _reference_cnuity(name) for name in rc.CNUITY})
    return out


def test_oracle_reproduces_the_asselin_and_cnuity_digests_of_the_reference_text(oracle):
    """everywhere (also where the reference tree is absent)"""
    import reftext_cases as rc
    gold = json.load(open(GOLD))
    for name in rc.ASSELIN:
        cfg, sea, g, cb, m, n = rc.asselin_case(name)
        ot = util.oracle_tile_from_cb(oracle, cb, sea)
        ot.asselin_filter(m, n)
        flds = {k: ot.f64(k).copy() for k in ("temp", "saln", "th3d", "dp")}
        flds["tracer"] = ot.f64("tracer").copy() if cb.ntracr else None
        assert rc.asselin_digest(cb, m, flds) == gold[name], name
        ot.close()
    for name in rc.CNUITY:
        cfg, sea, g, cb, st, m, n, isopyc, mxlkta = rc.cnuity_case(name)
        got = util.run_oracle_cnuity(oracle, cb, sea, st, m, n, isopyc=isopyc, mxlkta=mxlkta)
        d = rc.cnuity_digest(cb, m, n, got["dp"], got["uflx"], got["vflx"], got["p"],
                             got["dpmixl"][n - 1] if (isopyc or mxlkta) else None)
        assert d == gold[name], name


@pytest.mark.skipif(not rt.available(), reason="the reference source tree is not on this machine")
@pytest.mark.parametrize("name", ["box_fct2", "periodic_mpdata_tracers", "fct2c_btrmas", "diffusion_12t_mixed", "arctic_fct2"])
def test_frozen_vectors_equal_the_reference_text_live(name):
    frozen = json.load(open(os.path.join(os.path.dirname(GOLD), "tsadvc_golden.json")))
    assert reference_text_digest_of_golden_case(name) == frozen[name]


def test_frozen_vectors_equal_the_committed_digests_of_the_reference_text():
    """everywhere: tsadvc_golden.json (checked on the device by test_golden_vectors_on_device and on the oracle by
    test_oracle.py) holds exactly the digests the reference text produced (from_reference_text.json, `golden:` keys)"""
    frozen = json.load(open(os.path.join(os.path.dirname(GOLD), "tsadvc_golden.json")))
    gold = json.load(open(GOLD))
    names = [k[len("golden:"):] for k in gold if k.startswith("golden:")]
    assert sorted(names) == sorted(frozen)
    for name in names:
        assert gold["golden:" + name] == frozen[name], name


# ---------------------------------------------------------------------------------------------------------
# the MULTI-TILE xctilr of mod_xc_mp.h (:4664-4987; ARCTIC :4114-4662) as written, one thread per tile in the
# reference's SHMEM flavour, its neighbour tables from the reference's own xcspmd text (oracle/reference_text_mp.py):
# the oracle's orc_world_xctilr - against which the product's exchange schedule and its device pack / unpack are
# tested (tests/test_oracle.py, tests/test_parity_gpu.py) - leaves the same halos, and both leave what the
# single-tile text of mod_xc_sm.h leaves on the global array
# ---------------------------------------------------------------------------------------------------------
MP_CASES = [
    # ipr, jpr, nreg, itdm, jtdm, itypes, flavour
    (2, 2, 0, 48, 36, (1,), "MPI"),            # closed
    (2, 2, 0, 48, 36, (1,), "SHMEM"),
    (4, 2, 1, 48, 36, (1,), "MPI"),            # periodic in i
    (4, 2, 1, 48, 36, (1,), "SHMEM"),
    (3, 1, 3, 45, 30, (1,), "MPI"),            # f-plane: periodic in both, jpr = 1, ragged in i
    (3, 1, 3, 45, 30, (1,), "SHMEM"),
    (2, 3, 0, 47, 41, (1,), "MPI"),            # ragged in both
    (2, 2, 2, 48, 36, (1, 13, 14, 2, 3), "MPI"),   # across the arctic: the top row exchanges the fold with its twin tiles
    (4, 2, 2, 48, 36, (1, 13, 14), "MPI"),
    (2, 1, 2, 48, 36, (1, 13, 14), "MPI"),
]


@pytest.mark.skipif(not rt.available(), reason="the reference source tree is not on this machine")
@pytest.mark.parametrize("ipr,jpr,nreg,itdm,jtdm,itypes,flavour", MP_CASES)
def test_multi_tile_xctilr_of_the_reference_text_equals_oracle(oracle, ipr, jpr, nreg, itdm, jtdm, itypes, flavour):
    import reference_text_mp as rmp
    pkg = util.pkg
    kk, mh, nh = 3, 5, 5
    rng = np.random.default_rng(4)
    tiles = pkg.partition(itdm, jtdm, kk, ipr, jpr, nreg)
    g1 = pkg.partition(itdm, jtdm, kk, 1, 1, nreg)[0]
    nb = g1.nbdy
    world = rmp.World(tiles, ipr, jpr, nreg, itdm, jtdm, kk, flavour=flavour)
    ots = [oracle.tile(g, 0) for g in tiles]
    o1 = oracle.tile(g1, 0)
    for itype in itypes:
        core = rng.standard_normal((kk, jtdm, itdm))
        core[rng.random(core.shape) < 0.1] = 0.0            # vland inside the field
        def fresh():
            out = []
            for g in tiles:
                a = np.full((kk, g.nrows, g.ncols), np.nan)
                a[:, nb:nb + g.jj, nb:nb + g.ii] = core[:, g.j0:g.j0 + g.jj, g.i0:g.i0 + g.ii]
                out.append(a)
            return out
        ref, got = fresh(), fresh()
        world.xctilr(ref, 1, kk, mh, nh, itype)                               # the reference text on the tiles
        oracle.world_xctilr(ipr, jpr, ots, got, 1, kk, mh, nh, itype)         # the oracle on the tiles
        # the reference's single-tile text on the global array
        glob = np.full((kk, g1.nrows, g1.ncols), np.nan)
        glob[:, nb:nb + jtdm, nb:nb + itdm] = core
        env = rt.make_env(g1.ii, g1.jj, kk, nreg=nreg)
        env["xctilr"](fx_array3(glob, nb), 1, kk, mh, nh, itype)
        for g, a, b in zip(tiles, ref, got):
            assert np.array_equal(a, b, equal_nan=True), (itype, g.mproc, g.nproc)     # incl. what stays untouched
            loc = a[:, nb - nh:nb + g.jj + nh, nb - mh:nb + g.ii + mh]
            assert not np.isnan(loc).any()
            win = glob[:, g.j0 + nb - nh:g.j0 + nb + g.jj + nh, g.i0 + nb - mh:g.i0 + nb + g.ii + mh]
            assert np.array_equal(loc, win), (itype, g.mproc, g.nproc, "tiling invariance of the reference text")
    for o in ots + [o1]:
        o.close()


def fx_array3(a, nb):
    import fortran_exec as fx
    return fx.FArray(a, (1 - nb, 1 - nb, 1))


def test_geopar_metrics_of_the_reference_text_equal_oracle(oracle):
    """geopar.F90:311-340 as written (cell areas, their inverses, the aspect-limited grid spacing factors)"""
    if not rt.available():
        pytest.skip("the reference source tree is not on this machine")
    import fortran_exec as fx
    cfg, sea, g, cb = util.make_case(30, 24, 1, nreg=0, seed=3)
    nb = g.nbdy
    rng = np.random.default_rng(5)
    shp = (g.nrows, g.ncols)
    sc = {n: np.ascontiguousarray(8000.0 * (0.2 + rng.random(shp) * np.where(rng.random(shp) < 0.3, 5.0, 1.0)))
          for n in ("scpx", "scpy", "scux", "scuy", "scvx", "scvy", "scqx", "scqy")}
    sc["scux"][3, 4] = 0.0                       # the max(.., epsil) guards
    sc["scpx"][5, 6] = 0.0
    env = dict(nbdy=nb, ii=g.ii, jj=g.jj, epsil=1.0e-11, aspmax=2.0)
    for n, a in sc.items():
        env[n] = fx.FArray(a, (1 - nb, 1 - nb))
    for n in ("scu2", "scv2", "scp2", "scq2", "scuxi", "scvyi", "scp2i", "scq2i", "aspux", "aspuy", "aspvx", "aspvy", "util1", "depths"):
        env[n] = fx.FArray(np.zeros(shp), (1 - nb, 1 - nb))
    fx.compile_slice(os.path.join(rt.REF, "geopar.F90"), "geopar", r"^scu2\(i,j\)=scux\(i,j\)\*scuy\(i,j\)$", r"^if\s*\(ishelf\.eq\.0\)\s*then$",
                     "metrics", env, back=2)
    with np.errstate(all="ignore"):
        env["metrics"]()
    ot = oracle.tile(g, 0)
    ot.geopar(sc["scpx"], sc["scpy"], sc["scux"], sc["scuy"], sc["scvx"], sc["scvy"])
    for n in ("scp2", "scp2i", "aspux", "aspvy"):
        assert np.array_equal(ot.f64(n), env[n].a), n
    ot.close()


# the whole path on tiles, in the reference text: bigrid and tsadvc(m,n) of every tile on its own thread, their xctilr
# the multi-tile text of mod_xc_mp.h, xcmaxr a reduction over the threads - the reference's own tiling invariance
# (mod_pipe.F90:26-127), and the per-tile results the product's tiles are compared with on the GPU
@pytest.mark.skipif(not rt.available(), reason="the reference source tree is not on this machine")
@pytest.mark.parametrize("ipr,jpr,nreg,advtyp,ntracr", [(2, 2, 0, 2, 1), (2, 1, 1, 1, 0), (2, 1, 3, 4, 0),     # (the f-plane wants jpr = 1: mod_xc_mp.h:2855)
                                                      (2, 2, 2, 2, 1), (4, 2, 2, 1, 0)])                       # across the arctic
def test_tsadvc_of_the_reference_text_on_tiles_equals_one_tile(oracle, ipr, jpr, nreg, advtyp, ntracr):
    import reference_text_mp as rmp
    pkg, syn = util.pkg, util.syn
    m, n = 1, 2
    itdm, jtdm, kdm = (36, 28, 2) if nreg != 2 else (48, 36, 2)
    if nreg == 2:
        cfg, sea, g1, cb1 = util.make_arctic_case(itdm, jtdm, kdm, ntracr=ntracr, seed=11, m=m, n=n, advtyp=advtyp, nstep=3)
        cbs = util.make_arctic_tiles(cfg, sea, cb1, ipr, jpr, m, n, advtyp=advtyp, nstep=3)
    else:
        cfg, sea, g1, cb1 = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=11, m=m, n=n, advtyp=advtyp, nstep=3)
    ref = util.run_oracle(oracle, cb1, sea, m, n)                      # one tile, the oracle
    _run_reference_driver(cb1, sea, g1, m, n)                          # one tile, the reference text (updates cb1)
    tiles = pkg.partition(itdm, jtdm, kdm, ipr, jpr, nreg)
    if nreg != 2:
        cbs = [syn.build_cb_arrays(cfg, g, sea, m, n, advtyp=advtyp, nstep=3) for g in tiles]
    envs = [rt.make_env(g.ii, g.jj, kdm) for g in tiles]
    for env, g in zip(envs, tiles):
        env.update(i0=g.i0, j0=g.j0, itdm=itdm, jtdm=jtdm)
    world = rmp.World(tiles, ipr, jpr, nreg, itdm, jtdm, kdm, envs=envs)
    nb = g1.nbdy

    def tile(r, env):
        g, cb = tiles[r], cbs[r]
        depth = np.zeros((g.nrows, g.ncols))
        depth[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea[g.j0:g.j0 + g.jj, g.i0:g.i0 + g.ii] != 0, 100.0, 0.0)
        rt.run_bigrid(env, depth, mapflg=4 if nreg in (3, 4) else 0)
        assert env["nreg"] == nreg
        for name in ("ip", "iu", "iv"):
            assert np.array_equal(env[name].a[1:-1, 1:-1], getattr(cb, name)[1:-1, 1:-1]), name
        rt.add_cb_arrays(env, cb)
        rt.compile_tsadvc(env)
        rt.run_tsadvc(env, m, n)
    world.run(tile)
    for g, cb in zip(tiles, cbs):
        sea_t = cb.ip[nb:nb + g.jj, nb:nb + g.ii] != 0
        for name in ("temp", "saln"):
            loc = getattr(cb, name)[n - 1, :, nb:nb + g.jj, nb:nb + g.ii]
            for other in (getattr(cb1, name), ref[name]):
                glb = other[n - 1, :, nb + g.j0:nb + g.j0 + g.jj, nb + g.i0:nb + g.i0 + g.ii]
                assert np.array_equal(loc[:, sea_t], glb[:, sea_t]), (name, g.mproc, g.nproc)
        for q in range(ntracr):
            loc = cb.tracer[q, n - 1, :, nb:nb + g.jj, nb:nb + g.ii]
            glb = cb1.tracer[q, n - 1, :, nb + g.j0:nb + g.j0 + g.jj, nb + g.i0:nb + g.i0 + g.ii]
            assert np.array_equal(loc[:, sea_t], glb[:, sea_t]), ("tracer", q)
