#!/usr/bin/env python
"""Regenerates tests/golden/tsadvc_golden.json (TEST INFRASTRUCTURE).

The reference ships no golden vectors and cannot be compiled here (no Fortran compiler).  These digests freeze
the bits of the CPU oracle (oracle/tsadvc_oracle.c) on small seeded cases - and they ARE what the reference's own
source text computes on those cases: tests/golden/make_reference_text_vectors.py executes mod_tsadvc.F90 as written
(oracle/fortran_exec.py) on every case below and tests/test_reference_text.py demands that its digests
(from_reference_text.json, `golden:` keys) equal this file, live where /root/reference exists and from the committed
file everywhere.  Through tests/test_parity_gpu.py::test_golden_vectors_on_device the CUDA path must reproduce them.

    python tests/golden/make_golden.py        # rewrites the json next to this file
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), ".."))

# name -> (builder, kwargs); every case runs tsadvc(1,2)
CASES = {
    "box_fct2":      ("case", dict(itdm=150, jtdm=150, kdm=4, nreg=0, ntracr=0, seed=13, advtyp=2)),
    "periodic_mpdata_tracers": ("case", dict(itdm=131, jtdm=77, kdm=3, nreg=3, ntracr=2, seed=13, advtyp=1, trcflg=[0, 2])),
    "fct4_periodic_i": ("case", dict(itdm=64, jtdm=90, kdm=2, nreg=1, ntracr=1, seed=13, advtyp=4)),
    "pcm":           ("case", dict(itdm=70, jtdm=45, kdm=2, nreg=0, ntracr=0, seed=13, advtyp=0)),
    "fct2c_btrmas":  ("case", dict(itdm=90, jtdm=61, kdm=3, nreg=4, ntracr=1, seed=13, advtyp=2, btrmas=True)),
    "diffusion_17t": ("diff", dict(itdm=90, jtdm=61, kdm=3, sigver=6, temdfc=1.0, ntracr=1, seed=13)),
    "diffusion_12t_mixed": ("diff", dict(itdm=90, jtdm=61, kdm=3, sigver=8, temdfc=0.5, nhybrd=2, seed=13)),
    "arctic_fct2":   ("arctic", dict(itdm=90, jtdm=64, kdm=2, ntracr=0, seed=13, advtyp=2)),
    "isopyc":        ("case", dict(itdm=70, jtdm=45, kdm=3, nreg=0, seed=13, advtyp=2, isopyc=True, hybrid=False, nhybrd=0)),
}


def build(kind, kw):
    import util
    kw = dict(kw)
    if kind == "case":
        a = (kw.pop("itdm"), kw.pop("jtdm"), kw.pop("kdm"))
        return util.make_case(*a, **kw)
    if kind == "diff":
        a = (kw.pop("itdm"), kw.pop("jtdm"), kw.pop("kdm"), kw.pop("sigver"), kw.pop("temdfc"))
        return util.make_diffusion_case(*a, **kw)
    a = (kw.pop("itdm"), kw.pop("jtdm"), kw.pop("kdm"))
    return util.make_arctic_case(*a, **kw)


def digest(fields, msk, n):
    """sha256 over the sea cells of slot n of every field, layer by layer"""
    out = {}
    for name, a in fields.items():
        if a is None:
            continue
        h = hashlib.sha256()
        a = a[..., n - 1, :, :, :]
        h.update(np.ascontiguousarray(a[..., msk]).tobytes())
        out[name] = h.hexdigest()
    return out


def run_oracle_case(oracle, name):
    import util
    kind, kw = CASES[name]
    cfg, sea, g, cb = build(kind, kw)
    ref = util.run_oracle(oracle, cb, sea, 1, 2)
    msk = util.interior_sea(cb)
    flds = dict(temp=ref["temp"], saln=ref["saln"], th3d=ref["th3d"], tracer=ref.get("tracer"))
    return digest(flds, msk, 2), (cfg, sea, g, cb)


def main():
    import conftest  # noqa: F401  (path set-up)
    import oracle_binding
    root = os.path.dirname(os.path.dirname(HERE))
    import subprocess
    subprocess.run(["make", "-C", os.path.join(root, "oracle")], check=True, capture_output=True)
    orc = oracle_binding.Oracle(os.path.join(root, "oracle", "_build", "liboracle.so"))
    gold = {name: run_oracle_case(orc, name)[0] for name in CASES}
    with open(os.path.join(HERE, "tsadvc_golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print(json.dumps(gold, indent=1))


if __name__ == "__main__":
    main()
