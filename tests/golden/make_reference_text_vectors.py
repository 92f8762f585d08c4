#!/usr/bin/env python
"""Writes tests/golden/from_reference_text.json: digests of what the REFERENCE'S OWN SOURCE TEXT computes
(/root/reference/mod_tsadvc.F90 and bigrid.F90 executed through oracle/fortran_exec.py, no Fortran compiler
needed) on the cases of tests/test_reference_text.py.  Run in the build container, where /root/reference exists:

    python tests/golden/make_reference_text_vectors.py

The oracle must reproduce these digests on every machine (test_oracle_reproduces_the_digests_of_the_reference_text)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import reference_text as rt  # noqa: E402
import test_reference_text as T  # noqa: E402
import util  # noqa: E402


def main():
    assert rt.available(), "the reference source tree is needed"
    out = {}
    for itdm, jtdm, nreg, seed in T.GRIDS:
        cfg, sea, g, cb, depth = T.build_case(itdm, jtdm, nreg, seed)
        env = rt.make_env(g.ii, g.jj)
        rt.run_bigrid(env, depth.copy(), mapflg=4 if nreg in (3, 4) else 0)
        ops = T.advem_inputs(g, cb, seed)
        inner = util.interior_sea(cb)
        for advtyp in T.SCHEMES:
            posdef = 0.0 if advtyp != 1 else 256.0
            want = rt.run_advem(env, advtyp, ops[0].copy(), *[o.copy() for o in ops[1:6]], posdef, ops[6], ops[7],
                                cb.delt1)
            out[f"advem{advtyp}:{itdm}x{jtdm}:nreg{nreg}:seed{seed}"] = T.digest(want, inner)
    extra = getattr(T, "more_reference_vectors", None)
    if extra:
        out.update(extra())
    json.dump(out, open(os.path.join(HERE, "from_reference_text.json"), "w"), indent=1, sort_keys=True)
    print(len(out), "digests written")


if __name__ == "__main__":
    main()
