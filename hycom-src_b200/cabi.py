"""ctypes binding of include/hycom_tsadvc_b200.h and include/hycom_tsadvc_synth.h.

This is the stub a Python host would write; the Fortran host's equivalent is
fortran/mod_tsadvc_b200.F90.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libhycom_tsadvc_b200.so"

MXTRCR = 16

F_TEMP, F_SALN, F_TH3D, F_DP, F_UFLX, F_VFLX, F_TRACER, F_ONETA, F_THETA, F_Q2, F_Q2L = range(11)
(F_DPO, F_ONETAO, F_PBAVG, F_PBOT, F_OTEMP, F_OSALN, F_OTH3D, F_OTRACER, F_OQ2, F_OQ2L) = range(11, 21)
(F_U, F_V, F_DPU, F_DPV, F_UBAVG, F_VBAVG, F_DEPTHU, F_DEPTHV, F_P, F_DPMIXL, F_UFLXAV, F_VFLXAV, F_DPAV, F_UTOTN,
 F_VTOTN, F_DPMOLD, F_THKDF4U, F_THKDF4V) = range(21, 39)
S_SCPX, S_SCPY, S_SCUX, S_SCUY, S_SCVX, S_SCVY, S_ONETA = range(10, 17)

OK, EINVAL, ECUDA, EUNSUPPORTED, ENBDY, EADVTYP, ENOMEM = range(7)
PART_ALL, PART_INTERIOR, PART_FRAME = range(3)
DIRS = ("W", "E", "S", "N", "SW", "SE", "NW", "NE")


class TsadvcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"hycom_tsadvc error {code}: {msg}")
        self.code = code
        self.msg = msg


class XcStop(TsadvcError):
    """The reference would have called xcstop('tsadvc') / xcstop('advem')."""


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "idm", "jdm", "kdm", "nbdy", "ii", "jj", "i0", "j0", "itdm", "jtdm",
        "nreg", "ipr", "jpr", "mproc", "nproc", "ntracr", "device")]


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "advtyp", "advflg", "btrmas", "nhybrd", "hybrid", "isopyc", "mxlmy",
        "nstep", "diagno")] + [
        ("trcflg", C.c_int32 * MXTRCR), ("sigver", C.c_int32),
        ("delt1", C.c_double), ("temdf2", C.c_double), ("temdfc", C.c_double),
        ("thbase", C.c_double), ("onemm", C.c_double)]


class CnuityParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("btrmas", "isopyc", "hybrid", "mxlkta", "nstep", "pad")] + [
        (n, C.c_double) for n in ("delt1", "ra2fac", "thkdf2", "thkdf4")]


class SynthCfg(C.Structure):
    _fields_ = [("itdm", C.c_int32), ("jtdm", C.c_int32), ("kdm", C.c_int32),
                ("nreg", C.c_int32), ("ntracr", C.c_int32), ("pad", C.c_int32),
                ("seed", C.c_uint64), ("dx0", C.c_double), ("dy0", C.c_double),
                ("delt1", C.c_double)]


class SynthTile(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "idm", "jdm", "nbdy", "ii", "jj", "i0", "j0", "pad")]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes): every symbol the two headers declare
PROTOTYPES = {
    "hycom_tsadvc_abi_version": (C.c_int, []),
    "hycom_tsadvc_last_error": (C.c_char_p, [_vp]),
    "hycom_tsadvc_create": (C.c_int, [C.POINTER(Dims), C.POINTER(_vp)]),
    "hycom_tsadvc_destroy": (C.c_int, [_vp]),
    "hycom_tsadvc_set_stream": (C.c_int, [_vp, _vp]),
    "hycom_tsadvc_synchronize": (C.c_int, [_vp]),
    "hycom_tsadvc_device_bytes": (C.c_int64, [_vp]),
    "hycom_tsadvc_set_static": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hycom_tsadvc_step": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params),
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hycom_tsadvc_step_device": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params), _vp, _vp]),
    "hycom_tsadvc_upload": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "hycom_tsadvc_download": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "hycom_tsadvc_device_slab": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "hycom_tsadvc_halo_local": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "hycom_tsadvc_launch_count": (C.c_int64, [_vp]),
    "hycom_tsadvc_halo_neighbors": (C.c_int, [_vp, C.POINTER(C.c_int32 * 8)]),
    "hycom_tsadvc_halo_counts": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params),
                                           C.POINTER(C.c_int64 * 8)]),
    "hycom_tsadvc_halo_pack": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params),
                                         C.POINTER(_vp * 8), _vp]),
    "hycom_tsadvc_halo_unpack": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params),
                                           C.POINTER(_vp * 8), _vp]),
    "hycom_tsadvc_step_device_part": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params),
                                                C.c_int32, _vp, _vp]),
    "hycom_tsadvc_fct2c_batches": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "hycom_tsadvc_fct2c_stage": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params), C.c_int32, C.c_int32]),
    "hycom_tsadvc_fct2c_halo_counts": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params), C.c_int32,
                                                 C.POINTER(C.c_int64 * 8)]),
    "hycom_tsadvc_fct2c_halo_pack": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params), C.c_int32,
                                               C.POINTER(_vp * 8), _vp]),
    "hycom_tsadvc_fct2c_halo_unpack": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params), C.c_int32,
                                                 C.POINTER(_vp * 8), _vp]),
    "hycom_tsadvc_diff_halo_counts": (C.c_int, [_vp, C.c_int32, C.POINTER(Params), C.POINTER(C.c_int64 * 8)]),
    "hycom_tsadvc_diff_halo_pack": (C.c_int, [_vp, C.c_int32, C.POINTER(Params), C.POINTER(_vp * 8), _vp]),
    "hycom_tsadvc_diff_halo_unpack": (C.c_int, [_vp, C.c_int32, C.POINTER(Params), C.POINTER(_vp * 8), _vp]),
    "hycom_tsadvc_diffuse_device": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params)]),
    "hycom_tsadvc_set_frame_stream": (C.c_int, [_vp, _vp]),
    "hycom_tsadvc_asselin_save_device": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params), C.c_double]),
    "hycom_tsadvc_asselin_filter_device": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(Params), C.c_double,
                                                     C.c_double]),
    "hycom_tsadvc_set_timing": (C.c_int, [_vp, C.c_int32]),
    "hycom_tsadvc_get_timing": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]),
    "hycom_tsadvc_comm_unique_id": (C.c_int, [C.POINTER(C.c_char * 128)]),
    "hycom_tsadvc_comm_init": (C.c_int, [_vp, C.POINTER(C.c_char * 128)]),
    "hycom_tsadvc_comm_version": (C.c_int, [C.POINTER(C.c_int32)]),
    "hycom_tsadvc_local_group_create": (C.c_int, [C.c_int32, C.POINTER(_vp)]),
    "hycom_tsadvc_local_group_destroy": (C.c_int, [_vp]),
    "hycom_tsadvc_comm_attach_local": (C.c_int, [_vp, _vp]),
    "hycom_tsadvc_comm_detach": (C.c_int, [_vp]),
    "hycom_tsadvc_set_overlap": (C.c_int, [_vp, C.c_int32]),
    "hycom_tsadvc_xctilr": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "hycom_tsadvc_set_deferred_range": (C.c_int, [_vp, C.c_int32]),
    "hycom_tsadvc_saln_range": (C.c_int, [_vp, _vp, _vp, C.POINTER(C.c_int32)]),
    "hycom_tsadvc_checksum": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]),
    "hycom_tsadvc_cnuity_device": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(CnuityParams), _vp]),
    "hycom_synth_sea_mask": (C.c_int, [C.POINTER(SynthCfg), _vp]),
    "hycom_synth_fill_host": (C.c_int, [C.POINTER(SynthCfg), C.POINTER(SynthTile), _vp,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_double, _vp]),
    "hycom_tsadvc_synth_set_sea": (C.c_int, [_vp, C.POINTER(SynthCfg), _vp]),
    "hycom_tsadvc_synth_fill": (C.c_int, [_vp, C.POINTER(SynthCfg), C.c_int32, C.c_int32,
                                          C.c_int32, C.c_int32, C.c_int32, C.c_double]),
    "hycom_tsadvc_synth_fill_to": (C.c_int, [_vp, C.POINTER(SynthCfg), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
}

_lib = None


def lib_path() -> str:
    return os.path.join(HERE, _LIB_NAME)


def load_library() -> C.CDLL:
    """Load the CUDA C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
            " -- the tsadvc path has no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(lib, handle, rc: int) -> None:
    if rc == OK:
        return
    msg = lib.hycom_tsadvc_last_error(handle)
    msg = msg.decode() if msg else ""
    if rc in (ENBDY, EADVTYP):
        raise XcStop(rc, msg)
    raise TsadvcError(rc, msg)
