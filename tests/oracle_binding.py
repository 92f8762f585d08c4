"""ctypes binding of oracle/tsadvc_oracle.h (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

_vp = C.c_void_p


class OracleTile:
    """one orc_tile; numpy views of its arrays in the Fortran (i fastest) layout"""

    def __init__(self, orc, geom, ntracr=0):
        self.orc, self.lib, self.geom, self.ntracr = orc, orc.lib, geom, ntracr
        g = geom
        self.t = self.lib.orc_tile_create(g.idm, g.jdm, g.kdm, g.nbdy, g.ii, g.jj, g.i0, g.j0,
                                          g.itdm, g.jtdm, g.nreg, ntracr)
        if not self.t:
            raise MemoryError("orc_tile_create")
        self.P = g.nrows * g.ncols

    def close(self):
        if self.t:
            self.lib.orc_tile_destroy(self.t)
            self.t = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    _SHAPES = {"temp": 4, "saln": 4, "th3d": 4, "dp": 4, "tracer": 5, "uflx": 3, "vflx": 3,
               "oneta": -2, "onetamas": -2, "xmin": 1, "xmax": 1, "theta": 3, "q2": 6, "q2l": 6,
               "dpo": 4, "onetao": -2, "pbavg": -3, "otemp": 3, "osaln": 3, "oth3d": 3, "otracer": 7,
               "oq2": 8, "oq2l": 8,
               "u": 4, "v": 4, "dpu": 4, "dpv": 4, "ubavg": -3, "vbavg": -3, "dpmixl": -2, "p": 9,
               "uflxav": 3, "vflxav": 3, "dpav": 3, "dpkmin": 10}

    def f64(self, name):
        g = self.geom
        p = self.lib.orc_f64(self.t, name.encode())
        if not p:
            raise KeyError(name)
        kind = self._SHAPES.get(name, 2)
        shape = {2: (g.nrows, g.ncols), 3: (g.kdm, g.nrows, g.ncols),
                 4: (2, g.kdm, g.nrows, g.ncols), 5: (self.ntracr, 2, g.kdm, g.nrows, g.ncols),
                 -2: (2, g.nrows, g.ncols), 1: (g.kdm,), 6: (2, g.kdm + 2, g.nrows, g.ncols),
                 -3: (3, g.nrows, g.ncols), 7: (self.ntracr, g.kdm, g.nrows, g.ncols),
                 8: (g.kdm + 2, g.nrows, g.ncols), 9: (g.kdm + 1, g.nrows, g.ncols), 10: (2 * g.kdm,)}[kind]
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,)).reshape(shape)

    def i32(self, name):
        g = self.geom
        p = self.lib.orc_i32(self.t, name.encode())
        if not p:
            raise KeyError(name)
        if name == "trcflg":
            shape = (16,)
        elif name in ("isp",):
            shape = (g.nrows,)
        elif name in ("jsp",):
            shape = (g.ncols,)
        elif name in ("ifp", "ilp"):
            shape = (self.get_i("ms"), g.nrows)
        elif name in ("jfp", "jlp"):
            shape = (self.get_i("ms"), g.ncols)
        else:
            shape = (g.nrows, g.ncols)
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(n,)).reshape(shape)

    def set_i(self, name, v):
        assert self.lib.orc_set_i(self.t, name.encode(), int(v)) == 0, name

    def get_i(self, name):
        return self.lib.orc_get_i(self.t, name.encode())

    def set_d(self, name, v):
        assert self.lib.orc_set_d(self.t, name.encode(), float(v)) == 0, name

    # -- calls ------------------------------------------------------------
    def bigrid(self, depth):
        d = np.ascontiguousarray(depth, dtype=np.float64)
        rc = self.lib.orc_bigrid(self.t, d.ctypes.data_as(_vp))
        if rc:
            raise RuntimeError(self.orc.last_error())
        return d

    def geopar(self, scpx, scpy, scux, scuy, scvx, scvy):
        a = [np.ascontiguousarray(x) for x in (scpx, scpy, scux, scuy, scvx, scvy)]
        self.lib.orc_geopar_metrics(self.t, *[x.ctypes.data_as(_vp) for x in a])

    def xctilr(self, arr, l1, ld, mh, nh):
        self.lib.orc_xctilr(self.t, arr.ctypes.data_as(_vp), l1, ld, mh, nh)

    def tsadvc(self, m, n, do_halo=1):
        rc = self.lib.orc_tsadvc(self.t, m, n, do_halo)
        if rc:
            raise RuntimeError(f"orc_tsadvc rc={rc}: {self.orc.last_error()}")

    def cnuity_alloc(self):
        assert self.lib.orc_cnuity_alloc(self.t) == 0

    def cnuity(self, m, n, do_halo=1):
        rc = self.lib.orc_cnuity(self.t, m, n, do_halo)
        if rc:
            raise RuntimeError(f"orc_cnuity rc={rc}: {self.orc.last_error()}")

    def asselin_save(self, m, n, do_halo=1):
        self.lib.orc_asselin_save(self.t, m, n, do_halo)

    def asselin_filter(self, m, n):
        self.lib.orc_asselin_filter(self.t, m, n)

    def load_cb(self, cb):
        """copy a product-side CbArrays (host numpy) into this oracle tile"""
        for name in ("scp2", "scp2i", "scuy", "scvx", "aspux", "aspvy", "temp", "saln", "th3d",
                     "dp", "uflx", "vflx", "oneta", "theta", "q2", "q2l", "dpo", "onetao", "pbavg", "pbot",
                     "otemp", "osaln", "oth3d", "oq2", "oq2l"):
            src = getattr(cb, name)
            if src is not None:
                self.f64(name)[...] = src
        if cb.ntracr > 0:
            self.f64("tracer")[...] = cb.tracer
            if cb.otracer is not None:
                self.f64("otracer")[...] = cb.otracer
        for name in ("advtyp", "advflg", "btrmas", "hybrid", "isopyc", "mxlmy", "nstep", "diagno", "sigver"):
            self.set_i(name, int(getattr(cb, name)))
        self.set_i("nhybrd", cb.geom.kdm if cb.nhybrd < 0 else cb.nhybrd)
        for name in ("delt1", "temdf2", "temdfc", "thbase", "onemm", "ra2fac", "oneta0"):
            self.set_d(name, getattr(cb, name))
        tf = self.i32("trcflg")
        for q, v in enumerate(cb.trcflg):
            tf[q] = v


class Oracle:
    def __init__(self, path):
        lib = C.CDLL(path)
        lib.orc_tile_create.restype = _vp
        lib.orc_tile_create.argtypes = [C.c_int] * 12
        lib.orc_tile_destroy.argtypes = [_vp]
        lib.orc_f64.restype = _vp
        lib.orc_f64.argtypes = [_vp, C.c_char_p]
        lib.orc_i32.restype = _vp
        lib.orc_i32.argtypes = [_vp, C.c_char_p]
        lib.orc_slab.restype = C.c_int64
        lib.orc_slab.argtypes = [_vp]
        lib.orc_set_i.argtypes = [_vp, C.c_char_p, C.c_int]
        lib.orc_get_i.argtypes = [_vp, C.c_char_p]
        lib.orc_set_d.argtypes = [_vp, C.c_char_p, C.c_double]
        lib.orc_xctilr.restype = None
        lib.orc_xctilr.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.orc_xctilr_type.restype = None
        lib.orc_xctilr_type.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.orc_world_xctilr.restype = None
        lib.orc_world_xctilr.argtypes = [C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp),
                                         C.c_int, C.c_int, C.c_int, C.c_int]
        lib.orc_world_xctilr_type.restype = None
        lib.orc_world_xctilr_type.argtypes = [C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp),
                                              C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.orc_bigrid.argtypes = [_vp, _vp]
        lib.orc_bigrid_stage1.argtypes = [_vp, _vp]
        lib.orc_bigrid_stage2.argtypes = [_vp]
        lib.orc_geopar_metrics.restype = None
        lib.orc_geopar_metrics.argtypes = [_vp] * 7
        lib.orc_advem.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, _vp, _vp,
                                  C.c_double, C.c_int]
        lib.orc_tsadvc.argtypes = [_vp, C.c_int, C.c_int, C.c_int]
        lib.orc_cnuity_alloc.argtypes = [_vp]
        lib.orc_cnuity.argtypes = [_vp, C.c_int, C.c_int, C.c_int]
        lib.orc_asselin_save.restype = None
        lib.orc_asselin_save.argtypes = [_vp, C.c_int, C.c_int, C.c_int]
        lib.orc_asselin_filter.restype = None
        lib.orc_asselin_filter.argtypes = [_vp, C.c_int, C.c_int]
        lib.orc_sig.restype = C.c_double
        lib.orc_sig.argtypes = [C.c_int, C.c_double, C.c_double]
        lib.orc_tofsig.restype = C.c_double
        lib.orc_tofsig.argtypes = [C.c_int, C.c_double, C.c_double]
        lib.orc_set_tap.restype = None
        lib.orc_set_tap.argtypes = [C.c_char_p, _vp]
        lib.orc_clear_taps.restype = None
        lib.orc_last_error.restype = C.c_char_p
        self.lib = lib

    def last_error(self):
        return self.lib.orc_last_error().decode()

    def sig(self, sigver, t, s):
        return self.lib.orc_sig(sigver, t, s)

    def tofsig(self, sigver, r, s):
        return self.lib.orc_tofsig(sigver, r, s)

    def tile(self, geom, ntracr=0):
        return OracleTile(self, geom, ntracr)

    def world_xctilr(self, ipr, jpr, tiles, arrays, l1, ld, mh, nh, itype=1):
        n = ipr * jpr
        T = (_vp * n)(*[t.t for t in tiles])
        A = (_vp * n)(*[a.ctypes.data_as(_vp) for a in arrays])
        self.lib.orc_world_xctilr_type(ipr, jpr, T, A, l1, ld, mh, nh, itype)
