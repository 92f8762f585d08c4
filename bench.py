#!/usr/bin/env python
"""bench.py -- layer-cells/s of tsadvc(m,n) (T+S FCT advection) on B200, % of HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload GLBb0.08] [--advtyp 2]
                  [--ntracr 0] [--impl b200|reference]

One "step" = one tsadvc(m,n) call over all kdm layers of the named grid shape (BASELINE.json
configs[1]: GLBb0.08 4500x3298x41, T+S advem_fct2, synthetic fields).  `value` is measured with
every operand resident in HBM; `e2e` is the same call through the drop-in entry
hycom_tsadvc_step() on pinned HOST arrays, host<->device copies inside the timed region.
Under torchrun (N>1) the global grid is split into ipr x jpr tiles as mod_xc does, one rank
per GPU, halos exchanged over NVLink.  `--impl reference` times the CPU oracle (the
reference's algorithm restated in C, OpenMP over j like the reference's relo_omp build; the
Fortran itself cannot be compiled in this image) on a bounded sample.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TILINGS = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}
METRIC = "layer-cells/sec T+S FCT advection"
UNIT = "layer-cells/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def alg_bytes_per_call(idm, jdm, kk, advtyp, ntracr):
    """fp64, every array touched once.  FCT2 T+S 72 B/layer-cell
    (+24 per tracer), MPDATA 56 (+16), plus 28 B per (i,j) once per call for scp2, scp2i,
    ip, iu, iv."""
    per = (72 + 24 * ntracr) if advtyp == 2 else (56 + 16 * ntracr)
    return idm * jdm * kk * per + idm * jdm * 28


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# CPU arm: the oracle (TEST INFRASTRUCTURE; timed here only as the reported CPU baseline)
# ----------------------------------------------------------------------------------------
def cpu_oracle_rate(shape, advtyp, ntracr, nlay, calls, threads=None):
    """layer-cells/s of the CPU oracle on `nlay` layers of the named shape (bounded sample)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_binding
    import util
    pkg, syn, cabi = util.pkg, util.syn, util.cabi
    lib = os.path.join(ROOT, "oracle", "_build", "liboracle_fast.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    orc = oracle_binding.Oracle(lib)
    idm, jdm, kdm, baclin, dx = syn.SHAPES[shape]
    cores = threads or len(os.sched_getaffinity(0))
    cfg = syn.make_cfg(idm, jdm, kdm, nreg=0, ntracr=ntracr, seed=1, dx0=dx, delt1=2.0 * baclin)
    sea = syn.sea_mask(cfg)
    g = pkg.partition(idm, jdm, nlay, 1, 1, 0)[0]
    cb = syn.build_cb_arrays(cfg, g, sea, 1, 2, with_state=False, advtyp=advtyp)
    cb.ntracr = ntracr
    k0 = max(1, kdm // 2)

    def f4(fld, ktr=0, halo_mode=0):
        a = np.empty((2, nlay, g.nrows, g.ncols))
        for slot in (1, 2):
            a[slot - 1] = syn.fill_host(cfg, g, sea, fld, ktr, 0 if slot == 2 else 1, k0, nlay, halo_mode)
        return a
    cb.temp, cb.saln = f4(cabi.F_TEMP), f4(cabi.F_SALN)
    cb.th3d = np.zeros_like(cb.temp)
    cb.dp = f4(cabi.F_DP, halo_mode=1)
    cb.uflx = syn.fill_host(cfg, g, sea, cabi.F_UFLX, 0, 0, k0, nlay, 0)
    cb.vflx = syn.fill_host(cfg, g, sea, cabi.F_VFLX, 0, 0, k0, nlay, 0)
    cb.oneta = np.ones((2, g.nrows, g.ncols))
    if ntracr:
        cb.tracer = np.stack([f4(cabi.F_TRACER, ktr=q + 1) for q in range(ntracr)])
    ot = util.oracle_tile_from_cb(orc, cb, sea)
    ot.set_i("nthreads", cores)
    times = []
    for c in range(calls + 1):          # first call untimed (page faults of the scratch slabs)
        m, n = (1, 2) if c % 2 == 0 else (2, 1)
        t0 = time.perf_counter()
        ot.tsadvc(m, n, 1)
        times.append(time.perf_counter() - t0)
    ot.close()
    per_call = times[1:]
    return idm * jdm * nlay / (sum(per_call) / len(per_call)), cores, per_call


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    syn = importlib.import_module("hycom-src_b200").synthetic
    idm, jdm, kdm, _, _ = syn.SHAPES[args.workload]
    nlay = args.cpu_layers
    times_all = []
    cores = None
    t0 = time.perf_counter()
    rate, cores, per_call = cpu_oracle_rate(args.workload, args.advtyp, args.ntracr, nlay,
                                            calls=args.warmup + args.steps)
    timed = per_call[args.warmup:] if len(per_call) > args.warmup else per_call
    sec = sum(timed) / len(timed)
    value = idm * jdm * nlay / sec
    sample = (f"{nlay} of {kdm} layers of {args.workload} ({idm}x{jdm}) per step, "
              f"{len(timed)} timed tsadvc calls, C oracle -O2 -fopenmp schedule(static,jblk)")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 * kdm / nlay,
        "higher_is_better": True, "scaling": "weak" if False else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload} {idm}x{jdm}x{kdm} T+S advtyp={args.advtyp} ntracr={args.ntracr}",
                   "note": "ms_per_step extrapolated from the layer sample to kdm layers (layers are independent)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


# ----------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    pkg = importlib.import_module("hycom-src_b200")
    syn, cabi = pkg.synthetic, importlib.import_module("hycom-src_b200.cabi")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the tsadvc path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ipr, jpr = TILINGS[world]
    idm, jdm, kdm, baclin, dx = syn.SHAPES[args.workload]
    if args.kdm:            # profiling aid only (ncu replays): NOT the named config
        kdm = args.kdm
    cfg = syn.make_cfg(idm, jdm, kdm, nreg=0, ntracr=args.ntracr, seed=1, dx0=dx, delt1=2.0 * baclin)
    sea = syn.sea_mask(cfg)
    tiles = pkg.partition(idm, jdm, kdm, ipr, jpr, 0)
    g = tiles[rank]
    cb = syn.build_cb_arrays(cfg, g, sea, 1, 2, with_state=False, advtyp=args.advtyp,
                             trcflg=[0] * args.ntracr, temdf2=args.temdf2, temdfc=1.0, sigver=6)
    stream = torch.cuda.Stream()
    ts = pkg.Tsadvc(cb, device=local, stream=stream.cuda_stream)
    xc = None
    if world > 1:
        xc = pkg.XcExchange(ts, dist, compute_stream=stream)
        xc.frame_concurrent = not args.frame_serial
    # device-resident synthetic state, both leapfrog slots (dp too: the slots alternate)
    syn.fill_device(ts, cfg, sea, 1, 2, diffusion=args.temdf2 > 0.0)
    ts._ck(ts.lib.hycom_tsadvc_synth_fill(ts.h, cabi.C.byref(cfg), cabi.F_DP, 0, 1, 0, 1, float("nan")))
    ts.synchronize()

    def one_step(s):
        # HYCOM_Run: m=mod(nstep,2)+1; n=mod(nstep+1,2)+1 (mod_hycom.F90:2254-2257)
        m, n = s % 2 + 1, (s + 1) % 2 + 1
        cb.nstep = s + 1
        if xc is not None:      # pack -> NCCL send/recv -> unpack overlapped with the tile interior
            with torch.cuda.stream(stream):
                xc.tsadvc_device(m, n, diag=True, overlap=not args.no_overlap)
        else:
            ts.tsadvc_device(m, n, diag=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(args.warmup):
        one_step(s)
    barrier()
    ts.set_timing(True)
    ts.get_timing(reset=True)
    l0 = ts.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for s in range(args.warmup, args.warmup + args.steps):
            one_step(s)
        e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    march_ms, march_n = ts.get_timing(reset=True)
    ts.set_timing(False)
    launches = ts.launch_count - l0
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        # interior + frame launches of one step count as one marching pass over the tile
        t = torch.tensor([march_ms / args.steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        march_avg = float(t.item())
    else:
        march_avg = march_ms / max(march_n, 1)
    ms_step = ms_total / args.steps
    cells = idm * jdm * kdm
    value = cells / (ms_step * 1e-3)

    # roofline of the dominant kernel (k_tsadvc_march_tma; FCT2 and MPDATA run it as two launches per
    # call, the mask-free instantiation over the all-sea row segments and the general one over the
    # rest - timed together, events around the pair): algorithmic bytes of this rank's tile
    peak, peak_src = _peaks()
    alg = alg_bytes_per_call(g.ii, g.jj, kdm, args.advtyp, args.ntracr)
    achieved = alg / (march_avg * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_tsadvc_march_tma (general + all-sea launch of one call)" if
                args.advtyp in (1, 2) and os.environ.get("HYCOM_TSADVC_SPLIT", "1") != "0" else "k_tsadvc_march_tma",
                "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "alg_bytes_per_launch": alg, "kernel_ms": march_avg, "peak_source": peak_src,
                "kernel_share_of_step": march_avg / ms_step}
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tfile):
        try:
            tj = json.load(open(tfile))
            key = f"{args.workload}:advtyp{args.advtyp}:ntracr{args.ntracr}:gpus{world}"
            roofline["traffic"] = tj.get(key)
        except Exception:
            pass

    # e2e: the drop-in call on pinned host arrays (N=1: whole grid; N>1: this rank's tile)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, pkg, syn, cabi, cfg, sea, g, cb, ts, world, dist if world > 1 else None)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, cores, per_call = cpu_oracle_rate(args.workload, args.advtyp, args.ntracr,
                                                args.cpu_layers, calls=2)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_layers} of {kdm} layers of {args.workload}, 2 timed tsadvc calls after 1 "
                         f"warm-up, C oracle (gcc -O2 -fopenmp, schedule(static,jblk) over j)"}
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload} {idm}x{jdm}x{kdm} T+S advtyp={args.advtyp} ntracr={args.ntracr}"
                                   + (" (REDUCED kdm: profiling run, not a bench value)" if args.kdm else ""),
                       "tiling": f"{ipr}x{jpr}", "tile": f"{g.ii}x{g.jj}", "nreg": 0,
                       "l2": "inputs per step (%.1f GB) exceed L2 (126 MB); no flush" % (alg / 1e9),
                       "diag": "salinity min/max every 3rd step as mod_tsadvc.F90:2065",
                       "temdf2": args.temdf2},
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": int(launches),
            "clocks": clocks, "pct_of_hbm_roofline": 100.0 * alg * (1 if world == 1 else 1) /
                                                     (ms_step * 1e-3) / 1e9 / peak,
        }
        print(json.dumps(out))
    ts.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(args, pkg, syn, cabi, cfg, sea, g, cb, ts_dev, world, dist):
    """tsadvc(m,n) through hycom_tsadvc_step on pinned host arrays: H2D of every operand,
    compute, D2H of the advected fields, every step."""
    import numpy as np
    import torch
    kk = g.kdm
    P = g.nrows * g.ncols
    nt = args.ntracr

    def pinned(shape):
        t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
        return t, t.numpy()
    keep = []
    # generate on the device (fast), copy into the host arrays once: same synthetic state
    def host_from_device(fld, ktr=0):
        t, a = pinned((2, kk, g.nrows, g.ncols))
        keep.append(t)
        for slot in (1, 2):
            a[slot - 1] = ts_dev.download(fld, slot, ktr=ktr)
        return a
    cb.temp = host_from_device(cabi.F_TEMP)
    cb.saln = host_from_device(cabi.F_SALN)
    cb.dp = host_from_device(cabi.F_DP)
    t, cb.uflx = pinned((kk, g.nrows, g.ncols)); keep.append(t)
    cb.uflx[:] = ts_dev.download(cabi.F_UFLX, 1)
    t, cb.vflx = pinned((kk, g.nrows, g.ncols)); keep.append(t)
    cb.vflx[:] = ts_dev.download(cabi.F_VFLX, 1)
    if nt:
        t, cb.tracer = pinned((nt, 2, kk, g.nrows, g.ncols)); keep.append(t)
        for q in range(nt):
            for slot in (1, 2):
                cb.tracer[q, slot - 1] = ts_dev.download(cabi.F_TRACER, slot, ktr=q + 1)
    cb.th3d = None
    cb.oneta = np.ones((2, g.nrows, g.ncols))
    nadv = 2 + nt
    h2d = (2 * nadv + 3) * kk * P * 8          # fields both slots, dp(n), uflx, vflx
    d2h = nadv * kk * g.ii * g.jj * 8
    steps, warm = args.e2e_steps, 1
    times = []
    for s in range(warm + steps):
        m, n = s % 2 + 1, (s + 1) % 2 + 1
        cb.nstep = s + 1
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ts_dev.tsadvc(m, n)
        ts_dev.synchronize()
        dt = time.perf_counter() - t0
        if s >= warm:
            times.append(dt)
    sec = sum(times) / len(times)
    if dist is not None:
        t = torch.tensor([sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    idm, jdm = cfg.itdm, cfg.jtdm
    return {"value": idm * jdm * kk / sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": sec * 1e3, "steps": steps,
            "api": "hycom_tsadvc_step (pinned host arrays, Fortran layout)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="GLBb0.08")
    ap.add_argument("--advtyp", type=int, default=2)
    ap.add_argument("--ntracr", type=int, default=0)
    ap.add_argument("--kdm", type=int, default=0, help="override the layer count (profiling runs only)")
    ap.add_argument("--cpu-layers", type=int, default=2)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: exchange first, then the whole tile")
    ap.add_argument("--frame-serial", action="store_true",
                    help="N>1: launch the frame behind the interior instead of next to it (comparison)")
    ap.add_argument("--temdf2", type=float, default=0.0,
                    help="> 0: the step also runs tsdff_1x/2x + the EOS sweep (mod_tsadvc.F90:2138-2230); "
                         "the BASELINE metric is quoted without it")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    b = importlib.import_module("hycom-src_b200.build")
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        b.build_library()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
