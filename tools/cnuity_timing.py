#!/usr/bin/env python
"""Timing of cnuity(m,n) on the GLBb0.08 state of bench.py (device mirrors, synthetic operands); used by
tools/gpu_run.sh (steps cnuity, cnuity_launches, cnuity_ncu)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    warmup = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    a = argparse.Namespace(gpus=1, steps=3, warmup=3, impl="b200", workload="GLBb0.08", advtyp=2, ntracr=0, kdm=0,
                           cpu_layers=2, e2e_steps=2, no_e2e=True, no_cpu=True, no_extra=True, no_full_kdm=True,
                           no_overlap=False, py_transport=False, sync_range=False, temdf2=0.0)
    import torch
    torch.cuda.set_device(0)
    run = bench.Run(a, "GLBb0.08", 2, 0)
    r = run.cnuity_timing(steps=steps, warmup=warmup)
    print(json.dumps(r))
    run.close()


if __name__ == "__main__":
    main()
