"""oracle/reference_text.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Executes routines of the reference's own source text (/root/reference/*.F90) through oracle/fortran_exec.py on one
tile: bigrid (masks, sea-only neighbour indices, segment tables) and the advection schemes of mod_tsadvc.F90
(advem_pcm, advem_mpdata, advem_fct2, advem_fct4).  Used by tests/test_reference_text.py and
tests/golden/make_reference_text_vectors.py to pin the CPU oracle against what the reference text computes.

The only things supplied from outside the reference text are the module variables (dimensions, arrays) and the
communication calls of mod_xc on ONE tile: xctilr for a closed or periodic domain (fill with vland = 0 beyond a
closed edge, wrap across a periodic one: mod_xc_sm.h:1337-1428), xcmaxr/xcminr (identity), xcsync/xcstop.
"""
from __future__ import annotations

import os

import numpy as np

import fortran_exec as fx

REF = os.environ.get("HYCOM_REFERENCE", "/root/reference")
NBDY = 6


def available():
    return os.path.exists(os.path.join(REF, "mod_tsadvc.F90")) and os.path.exists(os.path.join(REF, "bigrid.F90"))


def _xctilr_factory(env):
    def xctilr(a, l1, ld, mh, nh, itype):
        """single tile, closed (vland = 0 outside) or periodic (wrap); `a` is a 2-D or 3-D FArray"""
        nb, ii, jj = env["nbdy"], env["ii"], env["jj"]
        per_i = env["nreg"] in (1, 3)
        per_j = env["nreg"] in (3, 4)
        if env["nreg"] == 2:
            raise NotImplementedError("arctic xctilr is not provided to the reference text")
        st = a.a if a.rank == 2 else a.a[l1 - a.lo[2]:l1 - a.lo[2] + ld]
        st = st.reshape((-1,) + st.shape[-2:])
        mh, nh = min(mh, nb), min(nh, nb)
        r0, c0 = nb, nb           # store index of (i,j) = (1,1)
        for s in st:
            # north / south
            for h in range(1, nh + 1):
                if per_j:
                    s[r0 - h, c0:c0 + ii] = s[r0 + jj - h, c0:c0 + ii]
                    s[r0 + jj - 1 + h, c0:c0 + ii] = s[r0 + h - 1, c0:c0 + ii]
                else:
                    s[r0 - h, c0:c0 + ii] = 0.0
                    s[r0 + jj - 1 + h, c0:c0 + ii] = 0.0
            # east / west over the rows just filled
            for h in range(1, mh + 1):
                rows = slice(r0 - nh, r0 + jj + nh)
                if per_i:
                    s[rows, c0 - h] = s[rows, c0 + ii - h]
                    s[rows, c0 + ii - 1 + h] = s[rows, c0 + h - 1]
                else:
                    s[rows, c0 - h] = 0.0
                    s[rows, c0 + ii - 1 + h] = 0.0
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
        env[n] = fx.FArray.zeros(b2, dtype=np.int64)
    for n in ("allip", "alliq", "alliu", "alliv"):
        env[n] = fx.FArray.zeros(((1 - nb, jj + nb),), dtype=np.int64)
    for p in "pquv":
        env["is" + p] = fx.FArray.zeros(((1 - nb, jj + nb),), dtype=np.int64)
        env["js" + p] = fx.FArray.zeros(((1 - nb, ii + nb),), dtype=np.int64)
        for f in ("if", "il"):
            env[f + p] = fx.FArray.zeros(((1 - nb, jj + nb), (1, ms)), dtype=np.int64)
        for f in ("jf", "jl"):
            env[f + p] = fx.FArray.zeros(((1 - nb, ii + nb), (1, ms)), dtype=np.int64)
    # scratch of mod_tsadvc (:38-64), r_init = NaN so that nothing uninitialised goes unnoticed
    for n in ("fmx", "fmn", "flx", "fly", "fldlo", "fmxlo", "fmnlo", "fax", "fay", "rp", "rm", "flxdiv", "tx1", "ty1",
              "fldao", "fldan", "uloc", "vloc", "hloc", "dtloc", "ucumdt", "vcumdt", "flxcum", "flycum"):
        env[n] = fx.FArray.zeros(b2, fill=np.nan)
    env["lcalc"] = fx.FArray.zeros(b2, dtype=np.int64)
    env["mbdy_advtyp"] = fx.FArray(np.array([2, 5, 5, 0, 5], dtype=np.int64), (0,))   # mod_tsadvc.F90:24-29
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
    for name in ("advem_pcm", "advem_mpdata", "advem_fct2", "advem_fct4"):
        if name not in env:
            fx.compile_unit(path, name, env, skip_calls=_SKIP, inout_calls=("xcmaxr", "xcminr"))


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
