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


# ---------------------------------------------------------------------------------------------------------
# the Fortran shim (fortran/mod_tsadvc_b200.F90) against include/hycom_tsadvc_b200.h: no Fortran compiler exists in
# this image, so what can be checked without one is checked here - the source is free-form with balanced blocks, every
# bind(c) interface names an exported entry with the header's argument count, scalars by value where the header takes
# values, and the bind(c) types have the fields of the C structs in the same order
# ---------------------------------------------------------------------------------------------------------
def _shim_statements():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fortran_exec as fx
    return [st for _, st in fx.load_source(os.path.join(ROOT, "fortran", "mod_tsadvc_b200.F90"), defines=("RELO",))]


def _c_prototypes():
    text = open(os.path.join(ROOT, "include", "hycom_tsadvc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|int64_t)\s+(hycom_tsadvc_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",")]
        out[m.group(1)] = args
    return out, text


def test_fortran_shim_is_free_form_with_balanced_blocks():
    raw = open(os.path.join(ROOT, "fortran", "mod_tsadvc_b200.F90")).read().split("\n")
    assert not [l for l in raw if re.match(r"^[cC*]\s", l)], "fixed-form comment lines"
    assert not [l for l in raw if re.match(r"^     [^ 0!]", l) and not l.rstrip().endswith("&") and
                raw[raw.index(l) - 1].rstrip().endswith("&") is False and l[5] in "&+123456789"], "column-6 continuations"
    depth = dict(do=0, if_=0, unit=0, iface=0, type_=0)
    for st in _shim_statements():
        if re.match(r"^do\b", st):
            depth["do"] += 1
        elif re.match(r"^end\s*do$", st):
            depth["do"] -= 1
        elif re.match(r"^if\s*\(.*\)\s*then$", st):
            depth["if_"] += 1
        elif re.match(r"^end\s*if$", st):
            depth["if_"] -= 1
        elif re.match(r"^interface\b", st):
            depth["iface"] += 1
        elif re.match(r"^end\s*interface", st):
            depth["iface"] -= 1
        elif re.match(r"^type\s*,", st):
            depth["type_"] += 1
        elif re.match(r"^end\s*type", st):
            depth["type_"] -= 1
        elif re.match(r"^(module|subroutine|(integer\s*\(\w+\)\s*)?function)\b", st) and not st.startswith("module procedure"):
            depth["unit"] += 1
        elif re.match(r"^end(\s+(module|subroutine|function)(\s+\w+)?)?$", st):
            depth["unit"] -= 1
        assert min(depth.values()) >= 0, st
    assert all(v == 0 for v in depth.values()), depth


def test_fortran_shim_interfaces_match_the_header(pkg):
    protos, text = _c_prototypes()
    lib = C.CDLL(cabi.lib_path())
    sts = _shim_statements()
    found = 0
    for k, st in enumerate(sts):
        m = re.match(r"^integer\s*\(c_int\)\s*function\s+(\w+)\s*\((.*?)\)\s*bind\s*\(\s*c\s*,\s*name\s*=\s*''\s*\)$", st)
        if not m:
            continue
        name, fargs = m.group(1), [a.strip() for a in m.group(2).split(",")]
        found += 1
        assert name in protos, name
        assert hasattr(lib, name), name
        cargs = protos[name]
        assert len(fargs) == len(cargs), (name, fargs, cargs)
        # declarations of the dummies up to `end function`
        decl = {}
        for st2 in sts[k + 1:]:
            if re.match(r"^end\s*function", st2):
                break
            md = re.match(r"^(.*?)::(.*)$", st2)
            if md:
                for ent in re.split(r",(?![^()]*\))", md.group(2)):
                    decl[re.sub(r"\(.*\)", "", ent).strip()] = md.group(1)
        for fa, ca in zip(fargs, cargs):
            assert fa in decl, (name, fa)
            if "c_ptr" in decl[fa]:      # an opaque pointer: by value for `T *h`, by reference for `T **out`
                by_value = ca.count("*") == 1
            else:
                by_value = "*" not in ca and "[" not in ca
            assert ("value" in decl[fa]) == by_value, (name, fa, decl[fa], ca)
            if "double" in ca:
                assert "c_double" in decl[fa], (name, fa)
            if "int32_t" in ca:
                assert "c_int32_t" in decl[fa], (name, fa)
    assert found >= 7

    # the bind(c) types against the C structs, field by field
    def c_fields(struct):
        body = re.search(r"typedef\s+struct\s*(?:\w+\s*)?\{([^}]*)\}\s*%s\s*;" % struct, text, flags=re.S).group(1)
        out = []
        for ln in body.split(";"):
            ln = ln.strip()
            if not ln:
                continue
            ty, names = ln.split(None, 1)
            out += [(ty, re.sub(r"\[.*\]", "", n).strip()) for n in names.split(",")]
        return out

    def f_fields(tname):
        out, inside = [], False
        for st in sts:
            if re.match(r"^type\s*,\s*bind\s*\(c\)\s*::\s*%s$" % tname, st):
                inside = True
            elif inside and re.match(r"^end\s*type", st):
                return out
            elif inside:
                ty, names = st.split("::")
                ty = "int32_t" if "c_int32_t" in ty else "double"
                out += [(ty, re.sub(r"\(.*\)", "", n).strip()) for n in re.split(r",(?![^()]*\))", names)]
        raise KeyError(tname)
    assert f_fields("tsadvc_dims") == c_fields("hycom_tsadvc_dims")
    assert f_fields("tsadvc_params") == c_fields("hycom_tsadvc_params")
