"""hycom-src_b200: B200-native tsadvc(m,n) (HYCOM mod_tsadvc.F90) behind a C ABI.

Host-side mirror of the reference interface for this one path:

* ``CbArrays``   -- the mod_cb_arrays / mod_dimensions fields tsadvc reads, as numpy
                    arrays in the Fortran layout (i fastest), halo width nbdy.
* ``Tsadvc``     -- ``tsadvc(m, n)`` with the reference's argument meaning and error
                    behaviour (mod_tsadvc.F90:1708-1712, :1817-1825, :159-166), running
                    on the CUDA library ``libhycom_tsadvc_b200.so``.

There is no CPU fallback: if the CUDA library is missing or no GPU is visible the
calls raise.  (import under the hyphenated name with
``importlib.import_module("hycom-src_b200")``.)
"""
from __future__ import annotations

from .cabi import (  # noqa: F401
    Dims,
    Params,
    SynthCfg,
    SynthTile,
    TsadvcError,
    XcStop,
    F_TEMP, F_SALN, F_TH3D, F_DP, F_UFLX, F_VFLX, F_TRACER,
    S_SCPX, S_SCPY, S_SCUX, S_SCUY, S_SCVX, S_SCVY, S_ONETA,
    lib_path,
    load_library,
)
from .geometry import (  # noqa: F401
    TileGeom,
    bigrid_masks,
    geopar_metrics,
    partition,
)
from .state import CbArrays, Tsadvc  # noqa: F401
from .xc import XcExchange, neighbors, halo_counts  # noqa: F401
from . import synthetic  # noqa: F401

__all__ = [
    "Dims", "Params", "SynthCfg", "SynthTile", "TsadvcError", "XcStop",
    "lib_path", "load_library", "TileGeom", "bigrid_masks", "geopar_metrics",
    "partition", "CbArrays", "Tsadvc", "synthetic", "XcExchange", "neighbors", "halo_counts",
]
