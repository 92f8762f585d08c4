"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/*.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import util
from util import pkg, cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(hycom_(?:tsadvc|synth|xc)_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol(pkg):
    lib = C.CDLL(cabi.lib_path())
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/ but not exported"
    # and the ctypes binding covers the same set
    assert declared == set(cabi.PROTOTYPES), declared ^ set(cabi.PROTOTYPES)


def test_abi_version(pkg):
    lib = cabi.load_library()
    assert lib.hycom_tsadvc_abi_version() == 3


def test_struct_layout_matches_header(pkg):
    # hycom_tsadvc_dims: 17 int32; params: 9 int32 + trcflg[16] + sigver + 5 doubles
    assert C.sizeof(cabi.Dims) == 17 * 4
    assert C.sizeof(cabi.Params) == (9 + 16 + 1) * 4 + 5 * 8
    assert cabi.Params.sigver.offset == 100
    assert cabi.Params.delt1.offset == 104


def test_no_cpu_fallback_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg, sea, g, cb = util.make_case(20, 20, 1)
    with pytest.raises(cabi.TsadvcError) as e:
        pkg.Tsadvc(cb)
    assert e.value.code == cabi.ECUDA
    assert "no CPU path" in str(e.value)


def test_partition_uniform_and_ragged():
    tiles = pkg.partition(4500, 3298, 41, 4, 2, 0)
    assert len(tiles) == 8
    assert sum(t.ii for t in tiles[:4]) == 4500
    assert tiles[0].jj + tiles[4].jj == 3298
    assert all(t.idm == 1125 and t.jdm == 1649 for t in tiles)
    tiles = pkg.partition(10, 7, 1, 3, 2, 0)
    assert [t.ii for t in tiles[:3]] == [4, 3, 3]
    assert [t.i0 for t in tiles[:3]] == [0, 4, 7]
    assert tiles[0].idm == 4 and tiles[0].jdm == 4


def test_synthetic_generator_is_tiling_invariant(pkg):
    """fields are pure functions of the global (i,j,k): a 2x2 tiling sees the bits of
    the single tile (the premise of every decomposition test)"""
    syn = pkg.synthetic
    cfg = util.make_cfg(50, 38, 2, nreg=0, seed=7)
    sea = syn.sea_mask(cfg)
    g1 = pkg.partition(50, 38, 2, 1, 1, 0)[0]
    full = syn.fill_host(cfg, g1, sea, cabi.F_SALN, 0, 0, 1, 2, 1)
    nb = g1.nbdy
    for g in pkg.partition(50, 38, 2, 2, 2, 0):
        loc = syn.fill_host(cfg, g, sea, cabi.F_SALN, 0, 0, 1, 2, 1)
        a = loc[:, nb:nb + g.jj, nb:nb + g.ii]
        b = full[:, nb + g.j0:nb + g.j0 + g.jj, nb + g.i0:nb + g.i0 + g.ii]
        assert np.array_equal(a, b)


def test_sea_mask_rules(pkg):
    """closed basin: last row/column land (bigrid.F90:25-45); no sea cell with >=3 land
    neighbours (bigrid.F90:156-191)"""
    syn = pkg.synthetic
    cfg = util.make_cfg(150, 150, 1, nreg=0, seed=1)
    sea = syn.sea_mask(cfg).astype(int)
    assert sea[-1].sum() == 0 and sea[:, -1].sum() == 0
    p = np.pad(sea, 1)
    nland = 4 - (p[1:-1, :-2] + p[1:-1, 2:] + p[:-2, 1:-1] + p[2:, 1:-1])
    assert (nland[sea == 1] <= 2).all()
    assert 0.5 < sea.mean() < 0.99
