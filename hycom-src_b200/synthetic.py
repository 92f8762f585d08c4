"""Synthetic mod_cb_arrays state (include/hycom_tsadvc_synth.h): deterministic fields
of the named grid shapes, identical for every tiling and for host and device."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import cabi
from .cabi import SynthCfg, SynthTile, load_library
from .geometry import TileGeom, bigrid_masks, geopar_metrics
from .state import CbArrays

# named shapes of BASELINE.json (idm, jdm, kdm, baclin seconds, grid spacing m)
SHAPES = {
    "box": (150, 150, 22, 1800.0, 20000.0),
    "GLBb0.08": (4500, 3298, 41, 240.0, 8900.0),
    "GLBy0.04": (9000, 7055, 41, 120.0, 4450.0),
}


def make_cfg(itdm, jtdm, kdm, nreg=0, ntracr=0, seed=1, dx0=8900.0, dy0=None, delt1=480.0) -> SynthCfg:
    return SynthCfg(itdm=itdm, jtdm=jtdm, kdm=kdm, nreg=nreg, ntracr=ntracr, pad=0, seed=seed,
                    dx0=dx0, dy0=dx0 if dy0 is None else dy0, delt1=delt1)


def shape_cfg(name: str, nreg=0, ntracr=0, seed=1) -> SynthCfg:
    idm, jdm, kdm, baclin, dx = SHAPES[name]
    return make_cfg(idm, jdm, kdm, nreg=nreg, ntracr=ntracr, seed=seed, dx0=dx, delt1=2.0 * baclin)


_host_lib = None


def use_host_library(path: str) -> None:
    """take hycom_synth_sea_mask / hycom_synth_fill_host from a CUDA-free build of the same source
    (oracle/_build/libsynth_host.so): the reference arm of bench.py then maps nothing of the product"""
    global _host_lib
    lib = C.CDLL(path)
    for name in ("hycom_synth_sea_mask", "hycom_synth_fill_host"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = cabi.PROTOTYPES[name]
    _host_lib = lib


def _hostlib():
    return _host_lib if _host_lib is not None else load_library()


def sea_mask(cfg: SynthCfg) -> np.ndarray:
    lib = _hostlib()
    sea = np.zeros((cfg.jtdm, cfg.itdm), dtype=np.uint8)
    rc = lib.hycom_synth_sea_mask(C.byref(cfg), sea.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"hycom_synth_sea_mask failed: {rc}")
    return sea


def _tile(g: TileGeom) -> SynthTile:
    return SynthTile(idm=g.idm, jdm=g.jdm, nbdy=g.nbdy, ii=g.ii, jj=g.jj, i0=g.i0, j0=g.j0, pad=0)


def fill_host(cfg: SynthCfg, g: TileGeom, sea: np.ndarray, field: int, ktr: int = 0, lev: int = 0,
              k0: int = 1, nk: int = 1, halo_mode: int = 0, fill: float = np.nan) -> np.ndarray:
    lib = _hostlib()
    out = np.empty((nk, g.nrows, g.ncols))
    t = _tile(g)
    rc = lib.hycom_synth_fill_host(C.byref(cfg), C.byref(t), sea.ctypes.data_as(C.c_void_p), field,
                                   ktr, lev, k0, nk, halo_mode, fill, out.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"hycom_synth_fill_host failed: {rc}")
    return out


def static_fields(cfg: SynthCfg, g: TileGeom, sea: np.ndarray):
    """masks (bigrid) and metrics (geopar) of one tile, halos valid"""
    ip, iu, iv = bigrid_masks(sea, g)
    raw = {f: fill_host(cfg, g, sea, f, halo_mode=1)[0]
           for f in (cabi.S_SCPX, cabi.S_SCPY, cabi.S_SCUX, cabi.S_SCUY, cabi.S_SCVX, cabi.S_SCVY)}
    scp2, scp2i, aspux, aspvy = geopar_metrics(raw[cabi.S_SCPX], raw[cabi.S_SCPY], raw[cabi.S_SCUX],
                                               raw[cabi.S_SCUY], raw[cabi.S_SCVX], raw[cabi.S_SCVY])
    return dict(ip=ip, iu=iu, iv=iv, scp2=np.ascontiguousarray(scp2),
                scp2i=np.ascontiguousarray(scp2i), scuy=raw[cabi.S_SCUY], scvx=raw[cabi.S_SCVX],
                aspux=np.ascontiguousarray(aspux), aspvy=np.ascontiguousarray(aspvy), raw=raw)


def build_cb_arrays(cfg: SynthCfg, g: TileGeom, sea: np.ndarray, m: int, n: int,
                    with_state: bool = True, **scalars) -> CbArrays:
    """Host CbArrays of one tile.  Fields that tsadvc exchanges (temp, saln, th3d,
    tracer, uflx, vflx) get NaN halos (the reference's r_init, mod_dimensions.F90:
    265-289) so a missing halo update cannot go unnoticed; dp and the 2-D arrays have
    valid halos ("dp halo is up to date", mod_tsadvc.F90:1828)."""
    st = static_fields(cfg, g, sea)
    cb = CbArrays(geom=g, ntracr=cfg.ntracr, ip=st["ip"], iu=st["iu"], iv=st["iv"], scp2=st["scp2"],
                  scp2i=st["scp2i"], scuy=st["scuy"], scvx=st["scvx"], aspux=st["aspux"],
                  aspvy=st["aspvy"], delt1=cfg.delt1, **scalars)
    if not with_state:
        return cb
    kk = g.kdm
    def lev_of(slot):  # slot n holds the old level (lev 0), slot m the centre (lev 1)
        return 0 if slot == n else 1
    def f4(fld, ktr=0, halo_mode=0):
        a = np.empty((2, kk, g.nrows, g.ncols))
        for slot in (1, 2):
            a[slot - 1] = fill_host(cfg, g, sea, fld, ktr, lev_of(slot), 1, kk, halo_mode)
        return a
    cb.temp = f4(cabi.F_TEMP)
    cb.saln = f4(cabi.F_SALN)
    cb.th3d = f4(cabi.F_TH3D)
    cb.dp = f4(cabi.F_DP, halo_mode=1)
    cb.uflx = fill_host(cfg, g, sea, cabi.F_UFLX, 0, 0, 1, kk, 0)
    cb.vflx = fill_host(cfg, g, sea, cabi.F_VFLX, 0, 0, 1, kk, 0)
    if cfg.ntracr > 0:
        cb.tracer = np.empty((cfg.ntracr, 2, kk, g.nrows, g.ncols))
        for q in range(cfg.ntracr):
            cb.tracer[q] = f4(cabi.F_TRACER, ktr=q + 1)
    cb.oneta = np.empty((2, g.nrows, g.ncols))
    for slot in (1, 2):
        cb.oneta[slot - 1] = fill_host(cfg, g, sea, cabi.S_ONETA, 0, lev_of(slot), 1, 1, 1)[0]
    return cb


def fill_device(ts, cfg: SynthCfg, sea: np.ndarray, m: int, n: int, advflg: int = 0, diffusion: bool = False):
    """Generate the same state straight into the device mirrors of ``ts`` (Tsadvc).
    ``diffusion``: also the operands of the temdf2>0 part (the other thermodynamic variable
    and oneta, both slots)."""
    lib, h = ts.lib, ts.h
    ck = ts._ck
    ck(lib.hycom_tsadvc_synth_set_sea(h, C.byref(cfg), sea.ctypes.data_as(C.c_void_p)))
    nan = float("nan")
    first = cabi.F_TH3D if advflg else cabi.F_TEMP
    for slot in (1, 2):
        lev = 0 if slot == n else 1
        ck(lib.hycom_tsadvc_synth_fill(h, C.byref(cfg), first, 0, slot, lev, 0, nan))
        ck(lib.hycom_tsadvc_synth_fill(h, C.byref(cfg), cabi.F_SALN, 0, slot, lev, 0, nan))
        for q in range(cfg.ntracr):
            ck(lib.hycom_tsadvc_synth_fill(h, C.byref(cfg), cabi.F_TRACER, q + 1, slot, lev, 0, nan))
    ck(lib.hycom_tsadvc_synth_fill(h, C.byref(cfg), cabi.F_DP, 0, n, 0, 1, nan))
    ck(lib.hycom_tsadvc_synth_fill(h, C.byref(cfg), cabi.F_UFLX, 0, 1, 0, 0, nan))
    ck(lib.hycom_tsadvc_synth_fill(h, C.byref(cfg), cabi.F_VFLX, 0, 1, 0, 0, nan))
    if diffusion:
        other = cabi.F_TEMP if advflg else cabi.F_TH3D
        for slot in (1, 2):
            lev = 0 if slot == n else 1
            ck(lib.hycom_tsadvc_synth_fill(h, C.byref(cfg), other, 0, slot, lev, 0, nan))
            ck(lib.hycom_tsadvc_synth_fill(h, C.byref(cfg), cabi.S_ONETA, 0, slot, lev, 1, nan))
    ts.synchronize()
