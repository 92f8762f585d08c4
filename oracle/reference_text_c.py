"""oracle/reference_text_c.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's source text COMPILED: oracle/fortran_to_c.py turns the same routines oracle/reference_text.py
interprets (bigrid, xctilr of mod_xc_sm.h, advem_* / advem / tsadvc of mod_tsadvc.F90 with stmt_fns.h,
asselin_save / asselin_filter, cnuity) into C, gcc builds oracle/_ref/libref_text_<eos>[_arctic][_omp].so from it -
one library per equation-of-state family and ARCTIC setting, as the reference itself is one executable per cpp
configuration.  The module variables are those of the interpreter's environment (reference_text.make_env,
add_cb_arrays ...), shared, not copied.  Needs /root/reference and gcc; the libraries travel to the GPU box with the
snapshot (oracle/_ref is git-ignored, not gpurun-ignored), where bench.py's reference arm may time them."""
from __future__ import annotations

import os

import numpy as np

import fortran_exec as fx
import fortran_to_c as f2c
import reference_text as rt

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
_SKIP = rt._SKIP + ("xcminr", "xcmaxr", "xcstop", "xchalt", "xctmr0", "xctmr1")
_DROP = (r"allocated", r"lconserve", r"2\*nchar")


def full_env(ii, jj, kdm, ntracr=1):
    """an environment that names EVERY module variable the compiled routines touch (sizes are irrelevant: the library
    reads bounds at run time) - the C globals are declared from it"""
    import types
    env = rt.make_env(ii, jj, kdm)
    nb = rt.NBDY
    b2 = ((1 - nb, ii + nb), (1 - nb, jj + nb))
    Z = lambda *extra: fx.FArray.zeros(b2 + tuple(extra))   # noqa: E731
    g = types.SimpleNamespace(nbdy=nb, kdm=kdm, ii=ii, jj=jj)
    cb = types.SimpleNamespace(geom=g, ntracr=ntracr, temp=Z((1, kdm), (1, 2)).a, saln=Z((1, kdm), (1, 2)).a,
                               th3d=Z((1, kdm), (1, 2)).a, dp=Z((1, kdm), (1, 2)).a, uflx=Z((1, kdm)).a, vflx=Z((1, kdm)).a,
                               oneta=Z((1, 2)).a, tracer=np.zeros((ntracr, 2, kdm) + Z().a.shape), mxlmy=False, theta=None,
                               scp2=Z().a, scp2i=Z().a, scuy=Z().a, scvx=Z().a, aspux=Z().a, aspvy=Z().a, trcflg=[0] * ntracr,
                               advtyp=2, advflg=0, btrmas=False, nhybrd=kdm, hybrid=True, isopyc=False, nstep=1, diagno=False,
                               delt1=1.0, temdf2=0.0, temdfc=1.0, thbase=34.0, onemm=9.806, oneta0=0.01, ra2fac=0.125,
                               dpo=Z((1, kdm), (1, 2)).a, onetao=Z((1, 2)).a, pbavg=Z((1, 3)).a, pbot=Z().a,
                               otemp=Z((1, kdm)).a, osaln=Z((1, kdm)).a, oth3d=Z((1, kdm)).a,
                               otracer=np.zeros((ntracr, kdm) + Z().a.shape))
    rt.add_cb_arrays(env, cb)
    rt.add_asselin_arrays(env, cb)
    st = dict(dp=cb.dp, dpo=cb.dpo, u=Z((1, kdm), (1, 2)).a, v=Z((1, kdm), (1, 2)).a, dpu=Z((1, kdm), (1, 2)).a,
              dpv=Z((1, kdm), (1, 2)).a, ubavg=Z((1, 3)).a, vbavg=Z((1, 3)).a, dpmixl=Z((1, 2)).a, uflx=cb.uflx, vflx=cb.vflx,
              uflxav=Z((1, kdm)).a, vflxav=Z((1, kdm)).a, dpav=Z((1, kdm)).a, pbot=cb.pbot, depthu=Z().a, depthv=Z().a)
    rt.add_cnuity_arrays(env, cb, st)
    env["vland"] = 0.0
    return env


class RefTextC:
    """the compiled reference text of one cpp configuration (EOS family `sigver`, ARCTIC or not)"""

    _cache = {}

    TIMED_FLAGS = ("-O2", "-march=x86-64-v3", "-mtune=native")      # the analogue of config/xc40-gnu-relo_omp:22

    @staticmethod
    def so_path(sigver=6, arctic=False, openmp=False, tag=""):
        eos = "_".join(d.lower() for d in rt._EOS_DEFINES[sigver])
        return os.path.join(OUT, f"libref_text_{eos}{'_arctic' if arctic else ''}{'_omp' if openmp else ''}{tag}.so")

    def __init__(self, sigver=6, arctic=False, openmp=False, flags=("-O2", "-ffp-contract=off"), tag=""):
        key = (sigver, arctic, openmp, flags)
        if key in RefTextC._cache:
            self.lib = RefTextC._cache[key]
            return
        if not rt.available():       # the GPU box: what was built where the reference tree exists
            self.lib = f2c.Library.prebuilt(self.so_path(sigver, arctic, openmp, tag))
            RefTextC._cache[key] = self.lib
            return
        defines = ("RELO",) + rt._EOS_DEFINES[sigver] + (("ARCTIC",) if arctic else ())
        gen = f2c.Generator(full_env(8, 8, 2), skip=_SKIP, drop=_DROP, openmp=openmp)
        R = lambda f: os.path.join(rt.REF, f)   # noqa: E731
        gen.add(R("mod_xc_sm.h"), "xctilr", defines)
        for name in ("indxi", "indxj", "bigrid"):
            gen.add(R("bigrid.F90"), name, defines)
        for name in ("advem_pcm", "advem_mpdata", "advem_fct2", "advem_fct4", "advem_fct2c", "advem", "tsadvc"):
            gen.add(R("mod_tsadvc.F90"), name, defines)
        for name in ("asselin_save", "asselin_filter"):
            gen.add(R("mod_asselin.F90"), name, defines)
        gen.add(R("cnuity.F90"), "cnuity", defines)
        self.lib = f2c.Library(gen, self.so_path(sigver, arctic, openmp, tag), flags=flags)
        RefTextC._cache[key] = self.lib

    def run(self, env, name, *args):
        env.setdefault("vland", 0.0)
        self.lib.bind(env)
        self.lib.call(name, *args)
        self.lib.pull(env)
        return env
