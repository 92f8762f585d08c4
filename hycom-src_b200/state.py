"""Host-side mirror of the reference interface of the path: mod_cb_arrays state
(``CbArrays``) and ``tsadvc(m, n)`` (``Tsadvc``), on top of the C ABI."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import cabi
from .cabi import Dims, Params, check, load_library
from .geometry import TileGeom

ONEM = 9806.0  # mod_cb_arrays.F90:842-846


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "arrays must be contiguous in the Fortran (i fastest) layout"
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class CbArrays:
    """The mod_cb_arrays / blkdat fields tsadvc(m,n) reads and writes.

    Shapes (i fastest == Fortran layout): 2-D (nrows, ncols); uflx, vflx
    (kdm, nrows, ncols); temp, saln, th3d, dp (2, kdm, nrows, ncols); tracer
    (ntracr, 2, kdm, nrows, ncols); oneta (2, nrows, ncols).
    """
    geom: TileGeom
    ntracr: int = 0
    # masks / metrics
    ip: np.ndarray = None
    iu: np.ndarray = None
    iv: np.ndarray = None
    scp2: np.ndarray = None
    scp2i: np.ndarray = None
    scuy: np.ndarray = None
    scvx: np.ndarray = None
    aspux: np.ndarray = None
    aspvy: np.ndarray = None
    # state
    temp: np.ndarray = None
    saln: np.ndarray = None
    th3d: np.ndarray = None
    dp: np.ndarray = None
    tracer: np.ndarray = None
    uflx: np.ndarray = None
    vflx: np.ndarray = None
    oneta: np.ndarray = None
    q2: np.ndarray = None      # (2, kdm+2, nrows, ncols): q2(:,:,0:kk+1,2), Mellor-Yamada tke (mxlmy)
    q2l: np.ndarray = None
    # mod_asselin.F90 operands: dpo (2,kdm,..), onetao (2,..), pbavg (3,..), pbot (..), otemp/osaln/oth3d
    # (kdm,..), otracer (ntracr,kdm,..), oq2/oq2l (kdm+2,..)
    dpo: np.ndarray = None
    onetao: np.ndarray = None
    pbavg: np.ndarray = None
    pbot: np.ndarray = None
    otemp: np.ndarray = None
    osaln: np.ndarray = None
    oth3d: np.ndarray = None
    otracer: np.ndarray = None
    oq2: np.ndarray = None
    oq2l: np.ndarray = None
    ra2fac: float = 0.125
    oneta0: float = 0.01
    theta: np.ndarray = None   # (kdm, nrows, ncols): isopycnic target densities - thbase
    # blkdat scalars (defaults = the benchmark configuration: FCT2, T&S, hybrid)
    advtyp: int = 2
    advflg: int = 0
    btrmas: bool = False
    nhybrd: int = -1          # -1: kdm
    hybrid: bool = True
    isopyc: bool = False
    mxlmy: bool = False
    nstep: int = 1
    diagno: bool = False
    trcflg: list = field(default_factory=list)
    delt1: float = 480.0
    temdf2: float = 0.0
    temdfc: float = 1.0
    thbase: float = 34.0
    onemm: float = ONEM * 0.001
    sigver: int = 6           # stmt_fns.h:2-22; 6 = -DEOS_SIG2 -DEOS_17T (the GLB builds)

    def params(self) -> Params:
        p = Params()
        p.advtyp, p.advflg, p.btrmas = self.advtyp, self.advflg, int(self.btrmas)
        p.nhybrd = self.geom.kdm if self.nhybrd < 0 else self.nhybrd
        p.hybrid, p.isopyc, p.mxlmy = int(self.hybrid), int(self.isopyc), int(self.mxlmy)
        p.nstep, p.diagno = self.nstep, int(self.diagno)
        for q in range(cabi.MXTRCR):
            p.trcflg[q] = self.trcflg[q] if q < len(self.trcflg) else 0
        p.delt1, p.temdf2, p.temdfc = self.delt1, self.temdf2, self.temdfc
        p.thbase, p.onemm = self.thbase, self.onemm
        p.sigver = self.sigver
        return p


class Tsadvc:
    """``tsadvc(m, n)`` of mod_tsadvc.F90 on a B200.

    ``Tsadvc(cb).tsadvc(m, n)`` is the drop-in call: same argument meaning as the
    reference (m, n = 1-based leapfrog slots; slot n holds t-1 on entry and t+1 on
    exit), operands taken from ``cb`` the way the Fortran takes them from
    mod_cb_arrays, XcStop raised where the reference calls xcstop.
    ``tsadvc_device`` is the same step on the device-resident mirrors.
    """

    def __init__(self, cb: CbArrays, device: int = 0, stream: Optional[int] = None):
        self.lib = load_library()
        self.cb = cb
        g = cb.geom
        d = Dims(idm=g.idm, jdm=g.jdm, kdm=g.kdm, nbdy=g.nbdy, ii=g.ii, jj=g.jj, i0=g.i0, j0=g.j0,
                 itdm=g.itdm, jtdm=g.jtdm, nreg=g.nreg, ipr=g.ipr, jpr=g.jpr, mproc=g.mproc,
                 nproc=g.nproc, ntracr=cb.ntracr, device=device)
        self.dims = d
        h = C.c_void_p()
        check(self.lib, None, self.lib.hycom_tsadvc_create(C.byref(d), C.byref(h)))
        self.h = h
        if stream is not None:
            self.set_stream(stream)
        self.xmin = np.full(g.kdm, np.nan)
        self.xmax = np.full(g.kdm, np.nan)
        if cb.ip is not None:
            self.set_static()

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.hycom_tsadvc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        check(self.lib, self.h, rc)

    # -- plumbing ---------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        self._ck(self.lib.hycom_tsadvc_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self.lib.hycom_tsadvc_synchronize(self.h))

    @property
    def device_bytes(self) -> int:
        return int(self.lib.hycom_tsadvc_device_bytes(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.hycom_tsadvc_launch_count(self.h))

    def set_timing(self, enable: bool = True):
        self._ck(self.lib.hycom_tsadvc_set_timing(self.h, int(enable)))

    def get_timing(self, reset: bool = False):
        """(accumulated device ms of the marching kernel, launches) since the last reset"""
        ms, nl = C.c_double(0.0), C.c_int64(0)
        self._ck(self.lib.hycom_tsadvc_get_timing(self.h, C.byref(ms), C.byref(nl), int(reset)))
        return ms.value, nl.value

    # -- multi-tile runs: the communicator lives in the handle (mod_xc's role) ---------
    def comm_init_nccl(self, dist, group=None):
        """attach an NCCL communicator: rank 0 of the torch.distributed group draws the id, the group
        broadcasts the 128 bytes (the only thing the host program contributes)"""
        import torch
        g = self.cb.geom
        rank = dist.get_rank(group)
        if rank != g.mproc - 1 + g.ipr * (g.nproc - 1):
            raise ValueError("tiles are placed row-major: rank = mproc-1 + ipr*(nproc-1)")
        buf = (C.c_char * 128)()
        if rank == 0:
            self._ck(self.lib.hycom_tsadvc_comm_unique_id(C.byref(buf)))
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        src = 0 if group is None else dist.get_global_rank(group, 0)
        dist.broadcast(t, src=src, group=group)
        buf.raw = bytes(t.cpu().numpy().tobytes())
        self._ck(self.lib.hycom_tsadvc_comm_init(self.h, C.byref(buf)))

    def comm_attach_local(self, group_handle):
        self._ck(self.lib.hycom_tsadvc_comm_attach_local(self.h, group_handle))

    def set_overlap(self, enable: bool = True):
        self._ck(self.lib.hycom_tsadvc_set_overlap(self.h, int(enable)))

    def set_deferred_range(self, enable: bool = True):
        self._ck(self.lib.hycom_tsadvc_set_deferred_range(self.h, int(enable)))

    def saln_range(self):
        """(nstep, xmin, xmax) of the last diagnostic step in deferred mode; nstep -1: nothing pending"""
        ns = C.c_int32(-1)
        self._ck(self.lib.hycom_tsadvc_saln_range(self.h, _ptr(self.xmin), _ptr(self.xmax), C.byref(ns)))
        return ns.value, self.xmin, self.xmax

    def checksum(self, fld: int, tlev: int = 1, ktr: int = 0, all_tiles: bool = True) -> int:
        """tiling-invariant checksum of the interior sea points of one mirror (PIPE_CHECK analogue)"""
        out = C.c_uint64(0)
        self._ck(self.lib.hycom_tsadvc_checksum(self.h, fld, ktr, tlev, int(all_tiles), C.byref(out)))
        return int(out.value)

    def xctilr(self, fld: int, tlev: int = 0, ktr: int = 0, mh: int = 5, nh: int = 5, itype: int = 1):
        self._ck(self.lib.hycom_tsadvc_xctilr(self.h, fld, ktr, tlev, mh, nh, itype))

    def set_static(self):
        cb = self.cb
        self._ck(self.lib.hycom_tsadvc_set_static(
            self.h, _ptr(cb.scp2), _ptr(cb.scp2i), _ptr(cb.scuy), _ptr(cb.scvx), _ptr(cb.aspux),
            _ptr(cb.aspvy), _ptr(cb.ip), _ptr(cb.iu), _ptr(cb.iv)))

    def upload(self, fld: int, host: np.ndarray, tlev: int = 1, ktr: int = 0, k0: int = 1):
        nk = host.shape[0] if host.ndim == 3 else 1
        self._ck(self.lib.hycom_tsadvc_upload(self.h, fld, ktr, tlev, k0, nk, _ptr(host)))

    def download(self, fld: int, tlev: int = 1, ktr: int = 0, k0: int = 1, nk: Optional[int] = None):
        g = self.cb.geom
        if nk is None:
            nlay = {cabi.F_ONETA: 1, cabi.F_ONETAO: 1, cabi.F_PBOT: 1, cabi.F_PBAVG: 3, cabi.F_Q2: g.kdm + 2,
                    cabi.F_Q2L: g.kdm + 2, cabi.F_OQ2: g.kdm + 2, cabi.F_OQ2L: g.kdm + 2, cabi.F_UBAVG: 3,
                    cabi.F_VBAVG: 3, cabi.F_DEPTHU: 1, cabi.F_DEPTHV: 1, cabi.F_P: g.kdm + 1, cabi.F_DPMIXL: 1,
                    cabi.F_UTOTN: 1, cabi.F_VTOTN: 1, cabi.F_DPMOLD: 1, cabi.F_THKDF4U: 1,
                    cabi.F_THKDF4V: 1}.get(fld, g.kdm)
            nk = nlay - k0 + 1
        out = np.empty((nk, g.nrows, g.ncols))
        self._ck(self.lib.hycom_tsadvc_download(self.h, fld, ktr, tlev, k0, nk, _ptr(out)))
        return out

    def upload_state(self, m: int, n: int):
        """push every operand of tsadvc(m,n) to the device mirrors"""
        cb = self.cb
        first, ff = (cb.th3d, cabi.F_TH3D) if cb.advflg else (cb.temp, cabi.F_TEMP)
        for t in (1, 2):
            self.upload(ff, first[t - 1], t)
            self.upload(cabi.F_SALN, cb.saln[t - 1], t)
            for q in range(cb.ntracr):
                self.upload(cabi.F_TRACER, cb.tracer[q, t - 1], t, ktr=q + 1)
        self.upload(cabi.F_DP, cb.dp[n - 1], n)
        self.upload(cabi.F_UFLX, cb.uflx, 1)
        self.upload(cabi.F_VFLX, cb.vflx, 1)
        if cb.mxlmy:           # q2, q2l (0:kk+1, both slots), mod_tsadvc.F90:2035-2048
            self.upload_q2()
        if cb.btrmas:          # onetamas(:,:,m) = oneta(:,:,n) (mod_tsadvc.F90:1806)
            self.upload(cabi.F_ONETA, cb.oneta[n - 1], n)
        if cb.temdf2 > 0.0:   # operands of the diffusion part (mod_tsadvc.F90:2138-2230)
            other, of = (cb.temp, cabi.F_TEMP) if cb.advflg else (cb.th3d, cabi.F_TH3D)
            self.upload(of, other[n - 1], n)
            self.upload(cabi.F_ONETA, cb.oneta[n - 1], n)
            self.upload_theta()

    def upload_q2(self):
        for t in (1, 2):
            self.upload(cabi.F_Q2, self.cb.q2[t - 1], t)
            self.upload(cabi.F_Q2L, self.cb.q2l[t - 1], t)

    def upload_theta(self):
        """theta is constant in time: pushed once, read only in exactly-isopycnal layers"""
        if self.cb.theta is not None:
            self.upload(cabi.F_THETA, self.cb.theta, 1)

    # -- next to the path: mod_asselin.F90 on the device mirrors ----------------
    def upload_asselin_state(self, m: int, n: int):
        """every operand of asselin_save / asselin_filter that tsadvc itself does not mirror"""
        cb = self.cb
        for t in (1, 2):
            self.upload(cabi.F_DPO, cb.dpo[t - 1], t)
            self.upload(cabi.F_DP, cb.dp[t - 1], t)
            self.upload(cabi.F_ONETAO, cb.onetao[t - 1], t)
            self.upload(cabi.F_ONETA, cb.oneta[t - 1], t)
            self.upload(cabi.F_TEMP, cb.temp[t - 1], t)
            self.upload(cabi.F_SALN, cb.saln[t - 1], t)
            self.upload(cabi.F_TH3D, cb.th3d[t - 1], t)
            for q in range(cb.ntracr):
                self.upload(cabi.F_TRACER, cb.tracer[q, t - 1], t, ktr=q + 1)
        self.upload(cabi.F_PBAVG, cb.pbavg, 1)
        self.upload(cabi.F_PBOT, cb.pbot, 1)
        for fld, a in ((cabi.F_OTEMP, cb.otemp), (cabi.F_OSALN, cb.osaln), (cabi.F_OTH3D, cb.oth3d)):
            self.upload(fld, a, 1)
        for q in range(cb.ntracr):
            self.upload(cabi.F_OTRACER, cb.otracer[q], 1, ktr=q + 1)
        if cb.mxlmy:
            self.upload_q2()
            self.upload(cabi.F_OQ2, cb.oq2, 1)
            self.upload(cabi.F_OQ2L, cb.oq2l, 1)
        self.upload_theta()

    def asselin_save_device(self, m: int, n: int):
        """asselin_save(m,n), mod_asselin.F90:28-82, on the device mirrors"""
        p = self.cb.params()
        self._ck(self.lib.hycom_tsadvc_asselin_save_device(self.h, m, n, C.byref(p), self.cb.oneta0))

    def asselin_filter_device(self, m: int, n: int):
        """asselin_filter(m,n), mod_asselin.F90:84-286, on the device mirrors"""
        p = self.cb.params()
        self._ck(self.lib.hycom_tsadvc_asselin_filter_device(self.h, m, n, C.byref(p), self.cb.ra2fac,
                                                            self.cb.oneta0))

    # -- upstream of the path: cnuity.F90 on the device mirrors ------------------
    _CN_NLAY = None

    def upload_cnuity_state(self, st: dict, m: int, n: int):
        """every operand of cnuity(m,n) (a dict of arrays in the Fortran layout, tests/util.add_cnuity)"""
        for t in (1, 2):
            self.upload(cabi.F_DP, st["dp"][t - 1], t)
        for fld, nm in ((cabi.F_U, "u"), (cabi.F_V, "v"), (cabi.F_DPU, "dpu"), (cabi.F_DPV, "dpv")):
            self.upload(fld, st[nm][m - 1], m)
        self.upload(cabi.F_UBAVG, st["ubavg"], 1)
        self.upload(cabi.F_VBAVG, st["vbavg"], 1)
        for t in (1, 2):
            self.upload(cabi.F_DPMIXL, st["dpmixl"][t - 1], t)
        for fld, nm in ((cabi.F_DEPTHU, "depthu"), (cabi.F_DEPTHV, "depthv"), (cabi.F_PBOT, "pbot")):
            self.upload(fld, st[nm], 1)
        for fld, nm in ((cabi.F_UFLX, "uflx"), (cabi.F_VFLX, "vflx"), (cabi.F_UFLXAV, "uflxav"), (cabi.F_VFLXAV, "vflxav"),
                        (cabi.F_DPAV, "dpav")):
            self.upload(fld, st[nm], 1)
        if "thkdf4u" in st:
            self.upload(cabi.F_THKDF4U, st["thkdf4u"], 1)
            self.upload(cabi.F_THKDF4V, st["thkdf4v"], 1)

    def cnuity_device(self, m: int, n: int, ra2fac: float = 0.125, thkdf2: float = 0.0, thkdf4: float = 0.0,
                      mxlkta: bool = False):
        """cnuity(m,n), cnuity.F90, on the device mirrors; returns dpkmin (2*kdm) when mod(nstep,3) == 0"""
        cb = self.cb
        p = cabi.CnuityParams(btrmas=int(cb.btrmas), isopyc=int(cb.isopyc), hybrid=int(cb.hybrid), mxlkta=int(mxlkta),
                              nstep=cb.nstep, pad=0, delt1=cb.delt1, ra2fac=ra2fac, thkdf2=thkdf2, thkdf4=thkdf4)
        out = np.full(2 * cb.geom.kdm, np.nan)
        self._ck(self.lib.hycom_tsadvc_cnuity_device(self.h, m, n, C.byref(p), _ptr(out)))
        return out

    # -- the path ---------------------------------------------------------
    def tsadvc(self, m: int, n: int):
        """tsadvc(m,n) on the host arrays of ``cb`` (copies in, computes, copies out)."""
        cb = self.cb
        p = cb.params()
        if cb.temdf2 > 0.0:
            self.upload_theta()
        if cb.mxlmy:      # not in the argument list of the C entry: mirrors are filled around the call
            self.upload_q2()
        self._ck(self.lib.hycom_tsadvc_step(
            self.h, m, n, C.byref(p), _ptr(cb.temp), _ptr(cb.saln), _ptr(cb.th3d), _ptr(cb.tracer),
            _ptr(cb.dp), _ptr(cb.uflx), _ptr(cb.vflx), _ptr(cb.oneta), _ptr(self.xmin),
            _ptr(self.xmax)))
        if cb.mxlmy:      # slot n back on 1:ii,1:jj, like the fields in the argument list
            g = cb.geom
            rows, cols = g.interior()
            for fld, a in ((cabi.F_Q2, cb.q2), (cabi.F_Q2L, cb.q2l)):
                a[n - 1][:, rows, cols] = self.download(fld, n, nk=g.kdm + 2)[:, rows, cols]

    def tsadvc_device(self, m: int, n: int, diag: bool = True):
        """tsadvc(m,n) on the device mirrors (no host<->device traffic)."""
        p = self.cb.params()
        xm = _ptr(self.xmin) if diag else None
        xx = _ptr(self.xmax) if diag else None
        self._ck(self.lib.hycom_tsadvc_step_device(self.h, m, n, C.byref(p), xm, xx))
