import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util, oracle_binding
pkg = util.pkg
orc = oracle_binding.Oracle(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))
n_i, n_j, kk = [int(a) for a in sys.argv[1:4]] if len(sys.argv) > 3 else (40, 30, 2)
cfg, sea, g, cb = util.make_case(n_i, n_j, kk, nreg=0, seed=3, advtyp=2)
ref = util.run_oracle(orc, cb, sea, 1, 2)
ts = pkg.Tsadvc(cb)
ts.tsadvc(1, 2)
msk = util.interior_sea(cb)
for name in ("temp", "saln"):
    a, b = getattr(cb, name)[1], ref[name][1]
    bad = (a != b) & msk
    print(name, "mismatch cells", int(bad.sum()), "of", int(msk.sum()) * kk)
    if bad.any():
        k, r, c = np.argwhere(bad)[0]
        print(" first at k,r,c", k, r, c, a[k, r, c], b[k, r, c], "rows with errors", sorted(set(np.argwhere(bad)[:, 1]))[:20], "cols", sorted(set(np.argwhere(bad)[:, 2]))[:20])
ts.close()
print("done")
