"""The REFERENCE'S SOURCE TEXT COMPILED (oracle/fortran_to_c.py -> gcc -> oracle/_ref/libref_text_*.so) against the
same text interpreted (oracle/fortran_exec.py) and against the CPU oracle:

  * on the small cases of tests/test_reference_text.py the two backends of the translator - written independently of
    each other, one emitting Python, one emitting C - must agree bit for bit on every routine of the path (bigrid,
    xctilr, every advem_*, tsadvc with diffusion and all eight equation-of-state families, asselin, cnuity);
  * at the FULL horizontal size of the headline configuration (GLBb0.08, 4500 x 3298) the compiled text and the
    oracle must agree bit for bit on a layer of FCT2 and of MPDATA with a tracer - the pin at the size the benchmark
    runs at, which the interpreter cannot reach.

Needs /root/reference and gcc (skipped on the GPU box)."""
import copy
import os
import sys

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reference_text as rt  # noqa: E402
import test_reference_text as T  # noqa: E402

pytestmark = pytest.mark.skipif(not rt.available(), reason="the reference source tree is not on this machine")


def _depth(g, sea):
    nb = g.nbdy
    d = np.zeros((g.nrows, g.ncols))
    d[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea != 0, 100.0, 0.0)
    return d


def compiled_env(g, sea, sigver=6):
    """environment with masks, neighbour indices and segment tables from the COMPILED bigrid"""
    import reference_text_c as rc
    lib = rc.RefTextC(sigver, g.nreg == 2)
    env = rt.make_env(g.ii, g.jj, g.kdm)
    d = _depth(g, sea)
    u = [np.zeros_like(d) for _ in range(3)]
    lib.run(env, "bigrid", d, 4 if g.nreg in (3, 4) else 0, *u)
    assert env["nreg"] == g.nreg
    return lib, env


@pytest.mark.parametrize("itdm,jtdm,nreg,seed", T.GRIDS)
def test_compiled_bigrid_equals_interpreted(itdm, jtdm, nreg, seed):
    cfg, sea, g, cb, depth = T.build_case(itdm, jtdm, nreg, seed)
    lib, cenv = compiled_env(g, sea)
    ienv = rt.make_env(g.ii, g.jj)
    rt.run_bigrid(ienv, depth.copy(), mapflg=4 if nreg in (3, 4) else 0)
    for name in ("ip", "iu", "iv", "iq", "ipim1", "ipip1", "ipjm1", "ipjp1", "isp", "jsp", "ifp", "ilp", "jfp", "jlp",
                 "isu", "ifu", "ilu", "jsv", "jfv", "jlv"):
        assert np.array_equal(cenv[name].a, ienv[name].a), name


def _driver(case, sigver=6, diff=None):
    m, n = 1, 2
    if diff is not None:
        sigver, temdfc, nreg, ntracr, nhybrd, mxlmy = diff

        def mk():
            cfg, sea, g, cb = util.make_diffusion_case(24, 20, 3, sigver, temdfc, nreg=nreg, ntracr=ntracr, nhybrd=nhybrd, seed=13,
                                                       nstep=3, m=m, n=n)
            if mxlmy:
                util.add_q2(cfg, sea, g, cb, m, n)
            return cfg, sea, g, cb
    else:
        itdm, jtdm, kdm, nreg, ntracr, advtyp, extra = case

        def mk():
            if nreg == 2:
                return util.make_arctic_case(itdm, jtdm, kdm, ntracr=ntracr, seed=11, m=m, n=n, advtyp=advtyp, nstep=3, **extra)
            return util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=11, m=m, n=n, advtyp=advtyp, nstep=3, **extra)
    cfg, sea, g, cb = mk()
    ienv = T._run_reference_driver(cb, sea, g, m, n, sigver)              # interpreted, updates cb
    cfg, sea, g, cb2 = mk()
    lib, env = compiled_env(g, sea, sigver)
    rt.add_cb_arrays(env, cb2)
    lib.run(env, "tsadvc", m, n)                                           # compiled, updates cb2
    inner = util.interior_sea(cb)
    for f in ("temp", "saln", "th3d"):
        assert np.array_equal(getattr(cb, f)[n - 1][..., inner], getattr(cb2, f)[n - 1][..., inner]), f
    if cb.ntracr:
        assert np.array_equal(cb.tracer[:, n - 1][..., inner], cb2.tracer[:, n - 1][..., inner])
    if getattr(cb, "mxlmy", False):
        assert np.array_equal(cb.q2[n - 1, 1:-1][..., inner], cb2.q2[n - 1, 1:-1][..., inner])
    assert np.array_equal(ienv["xmin"].a, env["xmin"].a) and np.array_equal(ienv["xmax"].a, env["xmax"].a)
    assert not np.array_equal(cb2.saln[n - 1][..., inner], mk()[3].saln[n - 1][..., inner])


@pytest.mark.parametrize("case", T.DRIVER_CASES)
def test_compiled_tsadvc_equals_interpreted(case):
    _driver(case)


@pytest.mark.parametrize("diff", T.DIFF_CASES)
def test_compiled_tsadvc_with_diffusion_equals_interpreted(diff):
    """every equation-of-state family: one compiled library per cpp configuration, like the reference executable
    (here both backends call the same libm, so tofsig of the 7/9-term fits agrees to the bit as well)"""
    _driver(None, diff=diff)


@pytest.mark.parametrize("sigver,ntracr,extra,mxlmy", T.ASSELIN_CASES)
def test_compiled_asselin_equals_interpreted(sigver, ntracr, extra, mxlmy):
    m, n = 1, 2
    cfg, sea, g, cb0 = util.make_case(26, 22, 3, nreg=0, ntracr=ntracr, seed=5, **extra)
    if mxlmy:
        util.add_q2(cfg, sea, g, cb0, m, n)
    util.add_asselin(cfg, sea, g, cb0, m, n, sigver=sigver)
    inner = util.interior_sea(cb0)
    for which in ("asselin_save", "asselin_filter"):
        ci, cc = copy.deepcopy(cb0), copy.deepcopy(cb0)
        ienv = rt.make_env(g.ii, g.jj, g.kdm)
        rt.run_bigrid(ienv, _depth(g, sea))
        rt.add_cb_arrays(ienv, ci)
        rt.add_asselin_arrays(ienv, ci)
        rt.compile_asselin(ienv, sigver)
        ienv[which](m, n)
        lib, env = compiled_env(g, sea, sigver)
        rt.add_cb_arrays(env, cc)
        rt.add_asselin_arrays(env, cc)
        lib.run(env, which, m, n)
        names = ("oneta", "onetao", "otemp", "osaln", "oth3d", "dp", "temp", "saln", "th3d") + (("tracer", "otracer") if ntracr else ()) \
            + (("q2", "q2l", "oq2", "oq2l") if mxlmy else ())
        for name in names:
            a, b = getattr(ci, name), getattr(cc, name)
            assert np.array_equal(a[..., inner], b[..., inner], equal_nan=True), (which, name)


@pytest.mark.parametrize("case", T.CNUITY_CASES)
def test_compiled_cnuity_equals_interpreted(case):
    itdm, jtdm, kdm, nreg, m, n, isopyc, thkdf, bih, nstep, mxlkta = case
    extra = dict(isopyc=True, hybrid=False, nhybrd=0) if isopyc else {}

    def mk():
        if nreg == 2:
            cfg, sea, g, cb = util.make_arctic_case(itdm, jtdm, kdm, seed=23, m=m, n=n, nstep=nstep, **extra)
            st = util.arctic_halos_cnuity(g, util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=thkdf, bih=bih))
        else:
            cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=23, m=m, n=n, nstep=nstep, **extra)
            st = util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=thkdf, bih=bih)
        if mxlkta:
            util.deepen_dpmixl(st, n)
        return cfg, sea, g, cb, st
    cfg, sea, g, cb, sti = mk()
    ienv = rt.make_env(g.ii, g.jj, g.kdm)
    rt.run_bigrid(ienv, _depth(g, sea), mapflg=4 if nreg in (3, 4) else 0)
    rt.add_cb_arrays(ienv, cb)
    rt.add_cnuity_arrays(ienv, cb, sti, mxlkta=mxlkta)
    rt.compile_cnuity(ienv)
    ienv["cnuity"](m, n)
    cfg, sea, g, cb2, stc = mk()
    lib, env = compiled_env(g, sea)
    rt.add_cb_arrays(env, cb2)
    rt.add_cnuity_arrays(env, cb2, stc, mxlkta=mxlkta)
    lib.run(env, "cnuity", m, n)
    inner = util.interior_sea(cb)
    for name in ("dp", "dpo", "uflx", "vflx", "dpmixl", "uflxav", "vflxav", "dpav"):
        assert np.array_equal(sti[name][..., inner], stc[name][..., inner], equal_nan=True), name
    for name in ("p", "utotn", "vtotn", "dpmold"):
        assert np.array_equal(ienv[name].a[..., inner], env[name].a[..., inner], equal_nan=True), name
    assert not np.array_equal(stc["dp"][n - 1, 0][inner], mk()[4]["dp"][n - 1, 0][inner])


# ---------------------------------------------------------------------------------------------------------
# the headline size: one layer of GLBb0.08 (4500 x 3298, 13.4 M sea cells), compiled reference text == oracle
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("advtyp,ntracr", [(2, 0), (1, 1)])
def test_compiled_text_equals_oracle_at_glbb008_size(oracle, advtyp, ntracr):
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(4500, 3298, 1, nreg=0, ntracr=ntracr, seed=13, m=m, n=n, advtyp=advtyp, nstep=3)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    lib, env = compiled_env(g, sea)
    rt.add_cb_arrays(env, cb)
    lib.run(env, "tsadvc", m, n)
    inner = util.interior_sea(cb)
    assert inner.sum() > 13_000_000
    for f in ("temp", "saln"):
        assert np.array_equal(getattr(cb, f)[n - 1][..., inner], ref[f][n - 1][..., inner]), f
    if ntracr:
        assert np.array_equal(cb.tracer[:, n - 1][..., inner], ref["tracer"][:, n - 1][..., inner])
    assert np.array_equal(env["xmin"].a, ref["xmin"]) and np.array_equal(env["xmax"].a, ref["xmax"])


# ---------------------------------------------------------------------------------------------------------
# stage by stage: after tsadvc(m,n) on BASELINE configs[0] (150 x 150 x 22 box) the module
# scratch of mod_tsadvc.F90:38-51 holds the intermediates of the last advem call - low-order fluxes, extrema,
# low-order solution, antidiffusive fluxes before and after the limiter, limiter ratios - exactly where the reference
# has its pipe_compare_sym hooks.  The oracle's scratch must hold the same bits, hence the same zero-flux and sign
# patterns, at every point the masks define.
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("advtyp", [2, 1, 4])
def test_intermediates_of_config1_equal_the_compiled_reference_text(oracle, advtyp):
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(150, 150, 22, nreg=0, ntracr=0, seed=13, m=m, n=n, advtyp=advtyp, nstep=3)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    ot.tsadvc(m, n, 1)
    lib, env = compiled_env(g, sea)
    rt.add_cb_arrays(env, cb)
    lib.run(env, "tsadvc", m, n)
    nb = g.nbdy
    inner = np.zeros((g.nrows, g.ncols), dtype=bool)
    inner[nb:nb + g.jj, nb:nb + g.ii] = True
    P, U, V = inner & (cb.ip != 0), inner & (cb.iu != 0), inner & (cb.iv != 0)
    stages = [("flx", U), ("fly", V), ("fmx", P), ("fmn", P), ("fldlo", P), ("flxdiv", P), ("rp", P), ("rm", P)]
    if advtyp in (2, 4):
        stages += [("fmxlo", P), ("fmnlo", P), ("fax", U), ("fay", V)]
    else:
        stages += [("tx1", U), ("ty1", V)]
    for name, msk in stages:
        a, b = ot.f64(name), env[name].a
        assert np.isfinite(b[msk]).all(), name
        assert np.array_equal(a[msk], b[msk]), name
        assert np.array_equal(np.signbit(a[msk]), np.signbit(b[msk])), (name, "sign pattern")
        assert np.array_equal(a[msk] == 0.0, b[msk] == 0.0), (name, "zero pattern")
    ot.close()


def test_compiled_cnuity_equals_oracle_on_a_large_grid(oracle):
    """cnuity(m,n) with the biharmonic interface-depth diffusion on two layers of half the GLBb0.08 extent in each
    direction (3.4 M sea cells; the full 4500 x 3298 passes too and takes 70 s): compiled reference text == oracle"""
    m, n, kdm = 1, 2, 2
    cfg, sea, g, cb = util.make_case(2250, 1649, kdm, nreg=0, seed=23, m=m, n=n, nstep=4)
    st = util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=0.01, bih=True)
    got = util.run_oracle_cnuity(oracle, cb, sea, st, m, n)
    lib, env = compiled_env(g, sea)
    rt.add_cb_arrays(env, cb)
    rt.add_cnuity_arrays(env, cb, st)
    lib.run(env, "cnuity", m, n)
    nb = g.nbdy
    inner = util.interior_sea(cb)
    iu_in, iv_in = np.zeros_like(inner), np.zeros_like(inner)
    iu_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iu[nb:nb + g.jj, nb:nb + g.ii] != 0
    iv_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iv[nb:nb + g.jj, nb:nb + g.ii] != 0
    assert inner.sum() > 3_000_000
    for name, msk in (("dp", inner), ("dpo", inner), ("uflx", iu_in), ("vflx", iv_in)):
        assert np.array_equal(got[name][..., msk], st[name][..., msk]), name
    assert np.array_equal(got["p"][1:][:, inner], env["p"].a[1:][:, inner])
    assert np.array_equal(got["utotn"][iu_in], env["utotn"].a[iu_in]) and np.array_equal(got["vtotn"][iv_in], env["vtotn"].a[iv_in])
