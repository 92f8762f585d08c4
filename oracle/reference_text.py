"""oracle/reference_text.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Executes routines of the reference's own source text (/root/reference/*.F90) through oracle/fortran_exec.py on one
tile: bigrid (masks, sea-only neighbour indices, segment tables), xctilr of mod_xc_sm.h (closed, periodic, ARCTIC),
the advection schemes of mod_tsadvc.F90 (advem_pcm, advem_mpdata, advem_fct2, advem_fct4, advem_fct2c), its driver
tsadvc(m,n) with tsdff_1x/2x and the statement functions of stmt_fns.h (all eight EOS families), asselin_save /
asselin_filter of mod_asselin.F90 and cnuity(m,n) of cnuity.F90.  Used by tests/test_reference_text.py and
tests/golden/make_reference_text_vectors.py to pin the CPU oracle against what the reference text computes.

The only things supplied from outside the reference text are the module variables (dimensions, arrays) and the
trivial communication calls of mod_xc on ONE tile: xcmaxr/xcminr (identity), xcsync/xcstop.  xctilr is the
reference's own single-tile text (mod_xc_sm.h), compiled with ARCTIC for nreg = 2.
"""
from __future__ import annotations

import os

import numpy as np

import fortran_exec as fx

REF = os.environ.get("HYCOM_REFERENCE", "/root/reference")
NBDY = 6
INT = np.int32      # integer and logical arrays are 4 bytes (no -fdefault-integer-8: config/generic-gnu-relo_one)


def available():
    return os.path.exists(os.path.join(REF, "mod_tsadvc.F90")) and os.path.exists(os.path.join(REF, "bigrid.F90"))


def _xctilr_factory(env):
    """xctilr of the reference's own mod_xc_sm.h (the single-tile mod_xc: closed / periodic :1337-1428, with the
    arctic fold :1172-1335 when the executable is compiled with ARCTIC, i.e. for nreg = 2), translated on first use.
    It reads nbdy, idm, jdm, ii, jj, nreg and vland from the module variables."""
    text = {}

    def xctilr(a, l1, ld, mh, nh, itype):
        arctic = env["nreg"] == 2
        if arctic not in text:
            sub = {k: env[k] for k in ("nbdy", "idm", "jdm", "ii", "jj")}
            sub.update(nreg=env["nreg"], vland=0.0)              # vland: mod_xc.F90 (real, save :: vland = 0.0)
            fx.compile_unit(os.path.join(REF, "mod_xc_sm.h"), "xctilr", sub, defines=("RELO",) + (("ARCTIC",) if arctic else ()),
                            skip_calls=("xctmr0", "xctmr1"))
            text[arctic] = sub
        sub = text[arctic]
        sub["nreg"] = env["nreg"]
        if a.rank == 2:
            a = fx.FArray(a.a[None], a.lo + (1,))
        elif a.lo[2] != 1:                                        # the dummy a(:,:,ld) numbers its slabs from 1
            a = fx.FArray(a.a, a.lo[:2] + (1,))
        sub["xctilr"](a, l1, ld, mh, nh, itype)
    return xctilr


def make_env(ii, jj, kdm=1, nreg=-1):
    """module variables of mod_dimensions / mod_xc / mod_cb_arrays / mod_tsadvc that the translated routines touch"""
    nb = NBDY
    b2 = ((1 - nb, ii + nb), (1 - nb, jj + nb))
    ms = max(ii, jj) // 2 + 8       # generous bound on the number of sea segments per row / column
    env = dict(idm=ii, jdm=jj, ii=ii, jj=jj, itdm=ii, jtdm=jj, i0=0, j0=0, nbdy=nb, kdm=kdm, kk=kdm, ms=ms,
               mnproc=1, lp=6, nreg=nreg, lpipe=False, ldebug_advem=False, ldebug_tsdif=False, itests=-99, jtests=-99,
               itest=-99, jtest=-99, jblk=jj, lpipe_advem=False, lconserve=False, flush_lp=1, no_flush=0,
               halo_ps=1, halo_pv=11, halo_qs=2, halo_qv=12, halo_us=3, halo_uv=13, halo_vs=4, halo_vv=14)
    for n in ("ip", "iu", "iv", "iq", "ipim1", "ipip1", "ipjm1", "ipjp1", "ipim1x", "ipip1x", "ipjm1x", "ipjp1x"):
        env[n] = fx.FArray.zeros(b2, dtype=INT)
    for n in ("allip", "alliq", "alliu", "alliv"):
        env[n] = fx.FArray.zeros(((1 - nb, jj + nb),), dtype=INT)
    for p in "pquv":
        env["is" + p] = fx.FArray.zeros(((1 - nb, jj + nb),), dtype=INT)
        env["js" + p] = fx.FArray.zeros(((1 - nb, ii + nb),), dtype=INT)
        for f in ("if", "il"):
            env[f + p] = fx.FArray.zeros(((1 - nb, jj + nb), (1, ms)), dtype=INT)
        for f in ("jf", "jl"):
            env[f + p] = fx.FArray.zeros(((1 - nb, ii + nb), (1, ms)), dtype=INT)
    # scratch of mod_tsadvc (:38-64), r_init = NaN so that nothing uninitialised goes unnoticed
    for n in ("fmx", "fmn", "flx", "fly", "fldlo", "fmxlo", "fmnlo", "fax", "fay", "rp", "rm", "flxdiv", "tx1", "ty1",
              "fldao", "fldan", "uloc", "vloc", "hloc", "dtloc", "ucumdt", "vcumdt", "flxcum", "flycum"):
        env[n] = fx.FArray.zeros(b2, fill=np.nan)
    env["lcalc"] = fx.FArray.zeros(b2, dtype=INT)
    env["mbdy_advtyp"] = fx.FArray(np.array([2, 5, 5, 0, 5], dtype=INT), (0,))   # mod_tsadvc.F90:24-29
    env["xctilr"] = _xctilr_factory(env)
    env["xcmaxr"] = lambda x: x
    env["xcminr"] = lambda x: x

    def xcstop(*a):
        raise fx.FortranStop(str(a))
    env["xcstop"] = xcstop
    env["xchalt"] = xcstop
    return env


_SKIP = ("xcsync", "xclget", "flush", "pipe_compare", "pipe_compare_sym1", "pipe_compare_sym2", "pipe_comparall",
         "mem_stat_add", "xcsum")


def compile_bigrid(env):
    path = os.path.join(REF, "bigrid.F90")
    ranks = {"indxi": (2, 2, 2, 1), "indxj": (2, 2, 2, 1)}
    for name in ("indxi", "indxj", "bigrid"):
        # (the block that prints the ip array - character handling, array sections - is left out)
        fx.compile_unit(path, name, env, skip_calls=_SKIP, inout_calls=("xcmaxr", "xcminr"), callee_ranks=ranks,
                        drop_blocks=(r"2\*nchar",))


def run_bigrid(env, depth, mapflg=0):
    """depth: (jj+2*nbdy, ii+2*nbdy) array, > 0 on sea (modified in place like the reference does); fills ip, iu,
    iv, iq, the sea-only neighbour indices and the segment tables of env"""
    nb, ii, jj = env["nbdy"], env["ii"], env["jj"]
    b2 = ((1 - nb, ii + nb), (1 - nb, jj + nb))
    lo = [b[0] for b in b2]
    d = fx.FArray(depth, lo)
    u1, u2, u3 = (fx.FArray.zeros(b2) for _ in range(3))
    if "bigrid" not in env:
        compile_bigrid(env)
    env["bigrid"](d, mapflg, u1, u2, u3)
    return env


def compile_advem(env):
    path = os.path.join(REF, "mod_tsadvc.F90")
    for name in ("advem_pcm", "advem_mpdata", "advem_fct2", "advem_fct4", "advem_fct2c"):
        if name not in env:
            fx.compile_unit(path, name, env, skip_calls=_SKIP, inout_calls=("xcmaxr", "xcminr"),
                            callee_ranks={"xctilr": (3, None, None, None, None, None)}, drop_blocks=(r"allocated",))


def run_advem(env, advtyp, fld, fldc, u, v, fco, fcn, posdef, scal, scali, dt2):
    """the reference's advem_* on (nrows, ncols) arrays; fld is updated in place"""
    nb = env["nbdy"]
    lo = (1 - nb, 1 - nb)
    compile_advem(env)
    A = lambda a: fx.FArray(a, lo)   # noqa: E731
    if advtyp == 0:
        env["advem_pcm"](A(fld), A(u), A(v), A(fco), A(fcn), A(scal), A(scali), dt2)
    elif advtyp == 1:
        env["advem_mpdata"](A(fld), A(u), A(v), A(fco), A(fcn), posdef, A(scal), A(scali), dt2)
    elif advtyp == 2:
        env["advem_fct2"](A(fld), A(fldc), A(u), A(v), A(fco), A(fcn), A(scal), A(scali), dt2)
    elif advtyp == 4:
        env["advem_fct4"](A(fld), A(fldc), A(u), A(v), A(fco), A(fcn), A(scal), A(scali), dt2)
    else:
        raise ValueError(advtyp)
    return fld


# ---------------------------------------------------------------------------------------------------------
# the driver tsadvc(m,n) itself (mod_tsadvc.F90:1708-2258) with its internal procedures tsdff_1x / tsdff_2x and the
# statement functions of stmt_fns.h
# ---------------------------------------------------------------------------------------------------------
_EOS_DEFINES = {1: ("EOS_SIG0", "EOS_7T"), 2: ("EOS_SIG2", "EOS_7T"), 3: ("EOS_SIG0", "EOS_9T"), 4: ("EOS_SIG2", "EOS_9T"),
                5: ("EOS_SIG0", "EOS_17T"), 6: ("EOS_SIG2", "EOS_17T"), 7: ("EOS_SIG0", "EOS_12T"), 8: ("EOS_SIG2", "EOS_12T")}


def add_cb_arrays(env, cb):
    """mod_cb_arrays variables that tsadvc touches, wrapping the arrays of a test case (no copies: the routine
    updates cb in place, as the reference updates its module arrays)"""
    g = cb.geom
    nb, kk = g.nbdy, g.kdm
    lo2 = (1 - nb, 1 - nb)
    W = lambda a, lo: fx.FArray(np.ascontiguousarray(a) if not a.flags["C_CONTIGUOUS"] else a, lo)   # noqa: E731
    env.update(kdm=kk, kk=kk)
    for name in ("temp", "saln", "th3d", "dp"):
        env[name] = W(getattr(cb, name), lo2 + (1, 1))
    env["uflx"], env["vflx"] = W(cb.uflx, lo2 + (1,)), W(cb.vflx, lo2 + (1,))
    env["oneta"] = W(cb.oneta, lo2 + (1,))
    env["onetamas"] = fx.FArray.zeros(((1 - nb, g.ii + nb), (1 - nb, g.jj + nb), (1, 2)), fill=np.nan)
    if cb.ntracr:
        env["tracer"] = W(cb.tracer, lo2 + (1, 1, 1))
    else:
        env["tracer"] = fx.FArray.zeros(((1 - nb, g.ii + nb), (1 - nb, g.jj + nb), (1, kk), (1, 2), (1, 1)))
    if getattr(cb, "mxlmy", False):
        env["q2"], env["q2l"] = W(cb.q2, lo2 + (0, 1)), W(cb.q2l, lo2 + (0, 1))
    else:
        z = ((1 - nb, g.ii + nb), (1 - nb, g.jj + nb), (0, kk + 1), (1, 2))
        env["q2"], env["q2l"] = fx.FArray.zeros(z), fx.FArray.zeros(z)
    if getattr(cb, "theta", None) is not None:
        env["theta"] = W(cb.theta, lo2 + (1,))
    else:
        env["theta"] = fx.FArray.zeros(((1 - nb, g.ii + nb), (1 - nb, g.jj + nb), (1, kk)))
    for name in ("scp2", "scp2i", "scuy", "scvx", "aspux", "aspvy"):
        env[name] = W(getattr(cb, name), lo2)
    b2 = ((1 - nb, g.ii + nb), (1 - nb, g.jj + nb))
    for name in ("util1", "util2", "util3", "uflux", "vflux", "uflux2", "vflux2", "sold", "told", "q2old", "q2lold"):
        env[name] = fx.FArray.zeros(b2, fill=np.nan)
    # geopar.F90:822-871: the flux scratch is zero on the land faces that bound sea segments
    for name in ("uflux", "vflux", "uflux2", "vflux2"):
        env[name].fill(0.0)
    env["trold"] = fx.FArray.zeros(b2 + ((1, max(cb.ntracr, 1)),), fill=np.nan)
    env["xmin"], env["xmax"] = fx.FArray.zeros(((1, kk),), fill=np.nan), fx.FArray.zeros(((1, kk),), fill=np.nan)
    trc = np.zeros(max(cb.ntracr, 1), dtype=INT)
    for q, v in enumerate(list(cb.trcflg)[:cb.ntracr]):
        trc[q] = v
    env["trcflg"] = fx.FArray(trc, (1,))
    env.update(advtyp=int(cb.advtyp), advflg=int(cb.advflg), btrmas=bool(cb.btrmas), nhybrd=int(cb.nhybrd if cb.nhybrd >= 0 else kk),
               hybrid=bool(cb.hybrid), isopyc=bool(cb.isopyc), mxlmy=bool(getattr(cb, "mxlmy", False)), ntracr=int(cb.ntracr),
               nstep=int(cb.nstep), diagno=bool(cb.diagno), delt1=float(cb.delt1), temdf2=float(cb.temdf2),
               temdfc=float(cb.temdfc), thbase=float(cb.thbase), onemm=float(cb.onemm), r_init=float("nan"),
               mxtrcr=max(cb.ntracr, 1), lpipe_tsadvc=False)
    return env


def compile_tsadvc(env, sigver=6):
    path = os.path.join(REF, "mod_tsadvc.F90")
    defines = ("RELO",) + _EOS_DEFINES[sigver]
    compile_advem(env)
    ranks = {"xctilr": (3, None, None, None, None, None),
             "advem": (None, 2, 2, 2, 2, 2, 2, None, 2, 2, None, None),
             "tsdff_1x": (2,), "tsdff_2x": (2, 2)}
    skip = _SKIP + ("xcminr", "xcmaxr")
    fx.compile_unit(path, "advem", env, defines=defines, skip_calls=skip, callee_ranks=ranks,
                    drop_blocks=(r"allocated", r"lconserve"))
    return fx.compile_unit(path, "tsadvc", env, defines=defines, skip_calls=skip, callee_ranks=ranks,
                           drop_blocks=(r"allocated",))


def run_tsadvc(env, m, n):
    env["tsadvc"](m, n)
    return env


# ---------------------------------------------------------------------------------------------------------
# mod_asselin.F90: asselin_save(m,n) :28-82 and asselin_filter(m,n) :84-286 (with stmt_fns.h)
# ---------------------------------------------------------------------------------------------------------
def add_asselin_arrays(env, cb):
    """module arrays of mod_cb_arrays that mod_asselin touches on top of add_cb_arrays (wrapped, not copied)"""
    g = cb.geom
    nb, kk = g.nbdy, g.kdm
    lo2 = (1 - nb, 1 - nb)
    W = lambda a, lo: fx.FArray(a, lo)   # noqa: E731
    for name in ("dpo",):
        env[name] = W(getattr(cb, name), lo2 + (1, 1))
    for name in ("onetao", "pbavg"):
        env[name] = W(getattr(cb, name), lo2 + (1,))
    env["pbot"] = W(cb.pbot, lo2)
    for name in ("otemp", "osaln", "oth3d"):
        env[name] = W(getattr(cb, name), lo2 + (1,))
    if cb.ntracr:
        env["otracer"] = W(cb.otracer, lo2 + (1, 1))
    else:
        env["otracer"] = fx.FArray.zeros(((1 - nb, g.ii + nb), (1 - nb, g.jj + nb), (1, kk), (1, 1)))
    if getattr(cb, "mxlmy", False):
        env["oq2"], env["oq2l"] = W(cb.oq2, lo2 + (0,)), W(cb.oq2l, lo2 + (0,))
    else:
        z = ((1 - nb, g.ii + nb), (1 - nb, g.jj + nb), (0, kk + 1))
        env["oq2"], env["oq2l"] = fx.FArray.zeros(z), fx.FArray.zeros(z)
    env.update(oneta0=float(cb.oneta0), ra2fac=float(cb.ra2fac), onem=9806.0)
    return env


def compile_asselin(env, sigver=6):
    path = os.path.join(REF, "mod_asselin.F90")
    defines = ("RELO",) + _EOS_DEFINES[sigver]
    ranks = {"xctilr": (3, None, None, None, None, None)}
    for name in ("asselin_save", "asselin_filter"):
        fx.compile_unit(path, name, env, defines=defines, skip_calls=_SKIP, callee_ranks=ranks)


# ---------------------------------------------------------------------------------------------------------
# cnuity(m,n) (cnuity.F90:14-1424)
# ---------------------------------------------------------------------------------------------------------
def add_cnuity_arrays(env, cb, st, mxlkta=False):
    """module variables cnuity touches on top of add_cb_arrays: `st` is the dictionary of tests/util.add_cnuity
    (arrays in the Fortran layout; wrapped, not copied, so cnuity updates them in place).  Scratch that the
    reference allocates on its first call (masku .. dpmn) and module scratch (utotm .. p) start as r_init = NaN."""
    g = cb.geom
    nb, kk = g.nbdy, g.kdm
    lo2 = (1 - nb, 1 - nb)
    b2 = ((1 - nb, g.ii + nb), (1 - nb, g.jj + nb))
    W = lambda a, lo: fx.FArray(a, lo)   # noqa: E731
    for name in ("dp", "dpo", "u", "v", "dpu", "dpv"):
        env[name] = W(st[name], lo2 + (1, 1))
    for name in ("ubavg", "vbavg", "dpmixl", "uflx", "vflx", "uflxav", "vflxav", "dpav"):
        env[name] = W(st[name], lo2 + (1,))
    for name in ("pbot", "depthu", "depthv"):
        env[name] = W(st[name], lo2)
    for name in ("thkdf4u", "thkdf4v"):
        env[name] = W(st[name], lo2) if name in st else fx.FArray.zeros(b2)
    for name in ("utotm", "vtotm", "utotn", "vtotn", "util1", "util2", "util3", "uflux", "vflux", "uflux2", "vflux2",
                 "pold", "oneta_u", "oneta_v", "dpmold", "onetacnt"):
        env[name] = fx.FArray.zeros(b2, fill=np.nan)
    # geopar.F90:822-871: the flux scratch is zero on the land faces that bound sea segments
    for name in ("uflux", "vflux", "uflux2", "vflux2", "utotm", "vtotm"):
        env[name].fill(0.0)
    for name in ("masku", "maskv", "iuopn", "ivopn"):
        env[name] = fx.FArray.zeros(b2, dtype=INT)
    env["dpmn"] = fx.FArray.zeros(((1 - nb, g.jj + nb),), fill=np.nan)
    env["p"] = fx.FArray.zeros(b2 + ((1, kk + 1),), fill=np.nan)
    env["p"].a[0] = 0.0                                        # p(:,:,1) = 0 always (geopar / inicon)
    env["wveli"] = fx.FArray.zeros(b2 + ((1, kk + 1),))         # mod_floats (synflt = .false.: never touched)
    env["onetamas"] = fx.FArray.zeros(b2 + ((1, 2),), fill=np.nan)
    thk = st.get("_thkdf")
    env.update(mxlkta=bool(mxlkta), synflt=False, wvelfl=False, onem=9806.0, onecm=98.06, qonem=1.0 / 9806.0,
               epsil=1.0e-11, ra2fac=0.125, lpipe_cnuity=False,
               thkdf4=float(thk[0]) if thk and thk[1] else 0.0, thkdf2=float(thk[0]) if thk and not thk[1] else 0.0)
    return env


def compile_cnuity(env):
    return fx.compile_unit(os.path.join(REF, "cnuity.F90"), "cnuity", env, skip_calls=_SKIP + ("xcminr", "xcmaxr"),
                           callee_ranks={"xctilr": (3, None, None, None, None, None)}, drop_blocks=(r"allocated",))
