#!/usr/bin/env python
"""I/O side of fortran/ref_driver.F90 (TEST INFRASTRUCTURE).

  ref_case.py write  <case> <dir>        inputs of a golden case (tests/golden/make_golden.py CASES) as raw
                                         little-endian files in the Fortran layout + case.txt
  ref_case.py digest <dir> <case> ...    out_*.bin of the reference -> tests/golden/from_reference.json,
                                         digests in the format of tests/golden/tsadvc_golden.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, ROOT)


def _case(name):
    import make_golden
    kind, kw = make_golden.CASES[name]
    cfg, sea, g, cb = make_golden.build(kind, kw)
    return cfg, sea, g, cb


def write(name, out):
    cfg, sea, g, cb = _case(name)
    os.makedirs(out, exist_ok=True)
    nb = g.nbdy
    depths = np.zeros((g.nrows, g.ncols))
    depths[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea[g.j0:g.j0 + g.jj, g.i0:g.i0 + g.ii] != 0, 100.0, 0.0)
    # the halo of depths as xctilr leaves it before bigrid (geopar.F90:247-256): periodic image or land
    per_i, per_j = g.nreg in (1, 2, 3), g.nreg in (3, 4)
    if per_i:
        depths[:, :nb] = depths[:, g.ii:g.ii + nb]
        depths[:, nb + g.ii:] = depths[:, nb:2 * nb]
    if per_j:
        depths[:nb, :] = depths[g.jj:g.jj + nb, :]
        depths[nb + g.jj:, :] = depths[nb:2 * nb, :]
    nhybrd = g.kdm if cb.nhybrd < 0 else cb.nhybrd
    ints = [g.itdm, g.jtdm, g.kdm, g.nreg, cb.ntracr, cb.advtyp, cb.advflg, int(cb.btrmas), nhybrd, int(cb.hybrid),
            int(cb.isopyc), int(cb.mxlmy), cb.nstep, int(cb.diagno), 1, 2]
    reals = [cb.delt1, cb.temdf2, cb.temdfc, cb.thbase, 0.0, 0.0, 0.0, 0.0]
    with open(os.path.join(out, "case.txt"), "w") as f:
        for v in ints:
            f.write(f"{v}\n")
        for v in reals:
            f.write(f"{v!r}\n")
    tf = np.zeros(max(cb.ntracr, 1), dtype="<i4")
    tf[:len(cb.trcflg)] = cb.trcflg[:cb.ntracr]
    tf.tofile(os.path.join(out, "trcflg.bin"))

    def put(fn, a):
        np.ascontiguousarray(a, dtype="<f8").tofile(os.path.join(out, fn + ".bin"))
    put("depths", depths)
    for nm in ("scp2", "scp2i", "scuy", "scvx", "aspux", "aspvy", "temp", "saln", "th3d", "dp", "uflx", "vflx", "oneta"):
        put(nm, getattr(cb, nm))
    if cb.ntracr:
        # Fortran tracer(i,j,k,t,ktr): ktr slowest = the (ntracr, 2, kdm, nrows, ncols) numpy order
        put("tracer", cb.tracer)
    if cb.theta is not None:
        put("theta", cb.theta)
    if cb.mxlmy:
        put("q2", cb.q2)
        put("q2l", cb.q2l)
    print(f"{name}: inputs in {out} (nreg {g.nreg}, sigver {cb.sigver}: build with the matching -DEOS_* flags)")


def digest(root, names):
    import make_golden
    import util
    gold = {}
    for name in names:
        cfg, sea, g, cb = _case(name)
        d = os.path.join(root, "case_" + name)
        shp = (2, g.kdm, g.nrows, g.ncols)
        flds = {nm: np.fromfile(os.path.join(d, f"out_{nm}.bin"), dtype="<f8").reshape(shp) for nm in ("temp", "saln", "th3d")}
        flds["tracer"] = (np.fromfile(os.path.join(d, "out_tracer.bin"), dtype="<f8").reshape((cb.ntracr,) + shp)
                          if cb.ntracr else None)
        gold[name] = make_golden.digest(flds, util.interior_sea(cb), 2)
    path = os.path.join(ROOT, "tests", "golden", "from_reference.json")
    with open(path, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print(json.dumps(gold, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "write":
        write(sys.argv[2], sys.argv[3])
    else:
        digest(sys.argv[2], sys.argv[3:])
