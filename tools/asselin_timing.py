#!/usr/bin/env python
"""Device time of asselin_filter (mod_asselin.F90:84-286) on a synthetic hybrid T/S state, as
algorithmic GB/s against the measured HBM peak.  Algorithmic bytes per layer-cell (latemp layers):
read dpo(n), dpo(m), dp(n), osaln, saln(m), saln(n), otemp, temp(m), temp(n) = 72 B, write dp(m),
saln(m), temp(m), th3d(m) = 32 B  =>  104 B.   usage: python tools/asselin_timing.py [idm jdm kdm]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import util
    idm, jdm, kdm = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (2250, 1649, 41)
    cfg, sea, g, cb = util.make_case(idm, jdm, kdm, nreg=0, seed=1)
    util.add_asselin(cfg, sea, g, cb, 1, 2)
    stream = torch.cuda.Stream()
    ts = util.pkg.Tsadvc(cb, device=0, stream=stream.cuda_stream)
    ts.upload_asselin_state(1, 2)
    ts.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(3):
            ts.asselin_filter_device(1, 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        reps = 10
        for _ in range(reps):
            ts.asselin_filter_device(1, 2)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = idm * jdm * kdm * 104
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    gbs = alg / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "k_asselin_filter (+ k_asselin_oneta)", "grid": f"{idm}x{jdm}x{kdm}", "ms": ms,
                      "alg_bytes": alg, "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak}))
    ts.close()


if __name__ == "__main__":
    main()
