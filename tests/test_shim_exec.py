"""The Fortran shim (fortran/mod_tsadvc_b200.F90: the drop-in `subroutine tsadvc(m,n)` a HYCOM build links instead of
mod_tsadvc.F90) EXECUTED without a Fortran compiler: oracle/fortran_exec.py runs its text on the module variables of
mod_cb_arrays, with the seven C entries it binds replaced by Python stand-ins that check every argument against
include/hycom_tsadvc_b200.h's contract and - for hycom_tsadvc_step - do the work with the CPU oracle.  The module
arrays must then hold what the REFERENCE'S OWN tsadvc(m,n), executed the same way, leaves in them: shim + library is a
drop-in for the reference routine.  (What the C library itself computes on the device is the business of the GPU
tests; this is the host side of the boundary)"""
import copy
import os
import sys
import types

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fortran_exec as fx  # noqa: E402
import reference_text as rt  # noqa: E402
import test_reference_text as T  # noqa: E402

pytestmark = pytest.mark.skipif(not rt.available(), reason="the reference source tree (stmt_fns.h) is not on this machine")
SHIM = os.path.join(ROOT, "fortran", "mod_tsadvc_b200.F90")


class StandIn:
    """the C ABI as the shim sees it; every call is recorded and checked"""

    def __init__(self, env, oracle, cb, sea, fail_step=0):
        self.env, self.oracle, self.cb, self.sea, self.fail_step = env, oracle, cb, sea, fail_step
        self.calls = []

    def install(self):
        e = self.env
        e.update(handle=None, c_associated=lambda h: h is not None, _newtype=self.newtype,
                 hycom_tsadvc_create=self.create, hycom_tsadvc_set_static=self.set_static, hycom_tsadvc_step=self.step,
                 hycom_tsadvc_upload=self.upload, hycom_tsadvc_download=self.download,
                 hycom_tsadvc_comm_unique_id=self.unique_id, hycom_tsadvc_comm_init=self.comm_init)

    @staticmethod
    def newtype(name):
        ns = types.SimpleNamespace(_type=name)
        if name == "tsadvc_params":
            ns.trcflg = fx.FArray(np.full(16, -1, dtype=np.int32), (1,))
        return ns

    def create(self, d, handle):
        e = self.env
        assert d._type == "tsadvc_dims" and handle is None
        for k in ("idm", "jdm", "kdm", "nbdy", "ii", "jj", "i0", "j0", "itdm", "jtdm", "nreg", "ipr", "jpr", "mproc", "nproc", "ntracr"):
            assert getattr(d, k) == e[k], k
        assert d.device == 0
        e["handle"] = object()
        self.calls.append("create")
        return 0

    def set_static(self, h, scp2, scp2i, scuy, scvx, aspux, aspvy, ip, iu, iv):
        e = self.env
        assert h is e["handle"]
        for a, n in ((scp2, "scp2"), (scp2i, "scp2i"), (scuy, "scuy"), (scvx, "scvx"), (aspux, "aspux"), (aspvy, "aspvy"),
                     (ip, "ip"), (iu, "iu"), (iv, "iv")):
            assert a is e[n], n                       # the module array itself: first element, no copy
        self.calls.append("set_static")
        return 0

    def upload(self, h, field, ktr, tlev, k0, nk, host):
        assert h is self.env["handle"]
        self.calls.append(("upload", field, ktr, tlev, k0, nk, host))
        return 0

    def download(self, h, field, ktr, tlev, k0, nk, host):
        raise AssertionError("no download expected without mxlmy")

    def unique_id(self, id_):
        raise AssertionError("single tile: no communicator")

    comm_init = unique_id

    def step(self, h, m, n, p, temp, saln, th3d, tracer, dp, uflx, vflx, oneta, xmin, xmax):
        e, cb = self.env, self.cb
        assert h is e["handle"] and p._type == "tsadvc_params"
        for k in ("advtyp", "advflg", "nhybrd", "nstep"):
            assert getattr(p, k) == e[k], k
        for k in ("btrmas", "hybrid", "isopyc", "mxlmy", "diagno"):
            assert getattr(p, k) == int(e[k]), k
        for k in ("delt1", "temdf2", "temdfc", "thbase", "onemm"):
            assert getattr(p, k) == e[k], k
        assert p.sigver == int(cb.sigver)
        assert [int(x) for x in p.trcflg.a[:cb.ntracr]] == [e["trcflg"][q + 1] for q in range(cb.ntracr)] and (p.trcflg.a[cb.ntracr:] == 0).all()
        for a, n_ in ((temp, "temp"), (saln, "saln"), (th3d, "th3d"), (dp, "dp"), (uflx, "uflx"), (vflx, "vflx"), (oneta, "oneta")):
            assert a is e[n_], n_
        assert (tracer is e["tracer"]) if cb.ntracr else (tracer is e["tr0"])
        assert xmin.a.shape == (e["kdm"],) and xmax.a.shape == (e["kdm"],)
        self.calls.append("step")
        if self.fail_step:
            return self.fail_step
        # the library's work, done by the CPU oracle on the arrays it was handed
        ref = util.run_oracle(self.oracle, cb, self.sea, m, n)
        inner = util.interior_sea(cb)
        for name in ("temp", "saln", "th3d"):
            getattr(cb, name)[n - 1][..., inner] = ref[name][n - 1][..., inner]
        if cb.ntracr:
            cb.tracer[:, n - 1][..., inner] = ref["tracer"][:, n - 1][..., inner]
        xmin.a[:], xmax.a[:] = ref["xmin"], ref["xmax"]
        return 0


def _shim_env(cb, sea, g):
    nb = g.nbdy
    depth = np.zeros((g.nrows, g.ncols))
    depth[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea != 0, 100.0, 0.0)
    env = rt.make_env(g.ii, g.jj, g.kdm)
    rt.run_bigrid(env, depth, mapflg=4 if g.nreg in (3, 4) else 0)
    rt.add_cb_arrays(env, cb)
    env.update(ipr=1, jpr=1, mproc=1, nproc=1, xmin=None, xmax=None, tr0=fx.FArray.zeros(((1, 1),)),
               xcastr=lambda a, n: None, c_null_ptr=None)
    return env


def _compile_shim(env, sigver):
    kw = dict(defines=("RELO",) + rt._EOS_DEFINES[sigver], include_dirs=(rt.REF,), skip_calls=rt._SKIP,
              extra_arrays={"xmin": 1, "xmax": 1})
    fx.compile_unit(SHIM, "b200_stop", env, **kw)
    return fx.compile_unit(SHIM, "tsadvc", env, **kw)


@pytest.mark.parametrize("case,sigver", [(T.DRIVER_CASES[0], 6), (T.DRIVER_CASES[1], 8), (T.DRIVER_CASES[4], 2), (T.DRIVER_CASES[5], 6),
                                         (T.DRIVER_CASES[2], 7)])
def test_shim_with_the_library_stood_in_by_the_oracle_equals_the_reference_routine(oracle, case, sigver):
    itdm, jtdm, kdm, nreg, ntracr, advtyp, extra = case
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=11, m=m, n=n, advtyp=advtyp, nstep=3, **extra)
    cb.sigver = sigver
    ref_cb = copy.deepcopy(cb)
    T._run_reference_driver(ref_cb, sea, g, m, n, sigver)                 # the reference's own tsadvc(m,n)
    env = _shim_env(cb, sea, g)
    lib = StandIn(env, oracle, cb, sea)
    lib.install()
    _compile_shim(env, sigver)
    env["tsadvc"](m, n)                                                    # the shim's tsadvc(m,n), first call
    assert lib.calls[:2] == ["create", "set_static"] and lib.calls[-1] == "step"
    inner = util.interior_sea(cb)
    for name in ("temp", "saln", "th3d"):
        assert np.array_equal(getattr(cb, name)[..., inner], getattr(ref_cb, name)[..., inner]), name
    if ntracr:
        assert np.array_equal(cb.tracer[..., inner], ref_cb.tracer[..., inner])
    # second call (the other leapfrog phase): no second create, the same handle
    h = env["handle"]
    env["nstep"] = 4
    env["tsadvc"](n, m)
    assert env["handle"] is h and lib.calls.count("create") == 1 and lib.calls.count("step") == 2


def test_shim_turns_error_codes_into_xcstop(oracle):
    """mod_tsadvc.F90:1817-1825, :159-166: a bad advtyp stops through xcstop('advem'), anything else through
    xcstop('tsadvc')"""
    cfg, sea, g, cb = util.make_case(20, 16, 1, seed=3)
    cb.sigver = 6
    for rc, where in ((5, "advem"), (4, "tsadvc"), (7, "tsadvc")):
        env = _shim_env(copy.deepcopy(cb), sea, g)
        said = []
        env["xcstop"] = lambda *a: said.append(a)
        lib = StandIn(env, oracle, cb, sea, fail_step=rc)
        lib.install()
        _compile_shim(env, 6)
        with pytest.raises(fx.FortranStop):
            env["tsadvc"](1, 2)
        assert len(said) == 1


def test_shim_with_diffusion_uploads_theta_once(oracle):
    """temdf2 > 0 with exactly isopycnal layers below nhybrd: theta goes to its mirror (field 8) on the first call"""
    m, n, sigver = 1, 2, 8
    cfg, sea, g, cb = util.make_diffusion_case(24, 20, 3, sigver, 1.0, nreg=0, ntracr=0, nhybrd=2, seed=13, nstep=3, m=m, n=n)
    cb.sigver = sigver
    ref_cb = copy.deepcopy(cb)
    T._run_reference_driver(ref_cb, sea, g, m, n, sigver)
    env = _shim_env(cb, sea, g)
    lib = StandIn(env, oracle, cb, sea)
    lib.install()
    _compile_shim(env, sigver)
    env["tsadvc"](m, n)
    ups = [c for c in lib.calls if isinstance(c, tuple)]
    assert len(ups) == 1 and ups[0][:6] == ("upload", 8, 0, 1, 1, g.kdm) and ups[0][6] is env["theta"]
    inner = util.interior_sea(cb)
    for name in ("temp", "saln", "th3d"):
        assert np.array_equal(getattr(cb, name)[n - 1][..., inner], getattr(ref_cb, name)[n - 1][..., inner]), name
    env["tsadvc"](n, m)
    assert len([c for c in lib.calls if isinstance(c, tuple)]) == 1


def test_shim_attaches_the_communicator_on_several_tiles(oracle):
    """ipr*jpr > 1: tile 1 draws the 128-byte id, it travels through mod_xc's own broadcast xcastr as 16 reals
    (transfer there and back) and every tile hands it to hycom_tsadvc_comm_init"""
    cfg, sea, g, cb = util.make_case(20, 16, 1, seed=3)
    cb.sigver = 6
    env = _shim_env(cb, sea, g)
    lib = StandIn(env, oracle, cb, sea)
    lib.install()
    env.update(ipr=2, jpr=1)
    seen = {}

    def unique_id(id_):
        assert id_.a.shape == (128,) and id_.a.dtype == np.uint8
        id_.a[:] = (np.arange(128) * 7 + 3) % 251
        seen["drawn"] = id_.a.copy()
        return 0

    def xcastr(a, root):
        assert a.a.shape == (16,) and a.a.dtype == np.float64 and root == 1
        seen["cast"] = a.a.copy()

    def comm_init(h, id_):
        assert h is env["handle"]
        seen["init"] = id_.a.copy()
        return 0
    env.update(hycom_tsadvc_comm_unique_id=unique_id, xcastr=xcastr, hycom_tsadvc_comm_init=comm_init)
    _compile_shim(env, 6)
    env["tsadvc"](1, 2)
    assert np.array_equal(seen["init"], seen["drawn"]) and seen["cast"].view(np.uint8).tolist() == seen["drawn"].tolist()
    assert lib.calls[:2] == ["create", "set_static"] and lib.calls[-1] == "step"


def test_shim_with_mellor_yamada_fields_fills_and_reads_their_mirrors(oracle):
    """mxlmy: q2, q2l (0:kk+1, both slots) are not arguments of hycom_tsadvc_step; the shim uploads both slots before
    the call (fields 9, 10: kdm+2 slabs from element (1-nbdy,1-nbdy,0,t)) and downloads slot n after it"""
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(26, 22, 3, nreg=0, ntracr=0, seed=11, m=m, n=n, advtyp=2, nstep=3)
    util.add_q2(cfg, sea, g, cb, m, n)
    cb.sigver = 6
    ref_cb = copy.deepcopy(cb)
    T._run_reference_driver(ref_cb, sea, g, m, n, 6)
    env = _shim_env(cb, sea, g)
    lib = StandIn(env, oracle, cb, sea)
    lib.install()
    mirrors = {}

    def addr(a):
        return a.__array_interface__["data"][0]

    def upload(h, field, ktr, tlev, k0, nk, host):
        name = {9: "q2", 10: "q2l"}[field]
        assert (ktr, k0, nk) == (0, 1, g.kdm + 2) and addr(host.a) == addr(env[name].a[tlev - 1])      # from layer 0 of slot tlev
        mirrors[(field, tlev)] = host.a[:nk * g.nrows * g.ncols].copy()
        lib.calls.append(("upload", field, tlev))
        return 0
    real_step = lib.step

    def step(*a):
        assert sorted(mirrors) == [(9, 1), (9, 2), (10, 1), (10, 2)]        # all four slabs are up before the call
        before = (cb.q2.copy(), cb.q2l.copy())
        rc = real_step(*a)
        ref = util.run_oracle(oracle, lib_cb_before, sea, m, n)
        mirrors[(9, n)], mirrors[(10, n)] = ref["q2"][n - 1].reshape(-1).copy(), ref["q2l"][n - 1].reshape(-1).copy()
        assert np.array_equal(cb.q2, before[0], equal_nan=True) and np.array_equal(cb.q2l, before[1], equal_nan=True)
        return rc

    def download(h, field, ktr, tlev, k0, nk, host):
        name = {9: "q2", 10: "q2l"}[field]
        assert tlev == n and (ktr, k0, nk) == (0, 1, g.kdm + 2) and addr(host.a) == addr(env[name].a[tlev - 1])
        host.a[:nk * g.nrows * g.ncols] = mirrors[(field, tlev)]
        lib.calls.append(("download", field, tlev))
        return 0
    lib_cb_before = copy.deepcopy(cb)
    env.update(hycom_tsadvc_upload=upload, hycom_tsadvc_step=step, hycom_tsadvc_download=download)
    kw = dict(defines=("RELO",) + rt._EOS_DEFINES[6], include_dirs=(rt.REF,), skip_calls=rt._SKIP, extra_arrays={"xmin": 1, "xmax": 1},
              callee_ranks={"hycom_tsadvc_upload": (None,) * 6 + (1,), "hycom_tsadvc_download": (None,) * 6 + (1,)})
    fx.compile_unit(SHIM, "b200_stop", env, **kw)
    fx.compile_unit(SHIM, "tsadvc", env, **kw)
    env["tsadvc"](m, n)
    order = [c for c in lib.calls if isinstance(c, tuple) or c == "step"]
    assert order[:4] == [("upload", 9, 1), ("upload", 10, 1), ("upload", 9, 2), ("upload", 10, 2)] and order[4] == "step"
    assert order[5:] == [("download", 9, n), ("download", 10, n)]
    inner = util.interior_sea(cb)
    for name in ("q2", "q2l"):
        assert np.array_equal(getattr(cb, name)[n - 1, 1:-1][..., inner], getattr(ref_cb, name)[n - 1, 1:-1][..., inner]), name
    for name in ("temp", "saln"):
        assert np.array_equal(getattr(cb, name)[n - 1][..., inner], getattr(ref_cb, name)[n - 1][..., inner]), name
