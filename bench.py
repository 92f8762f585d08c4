#!/usr/bin/env python
"""bench.py -- layer-cells/s of tsadvc(m,n) (T+S FCT advection) on B200, % of HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload GLBb0.08] [--advtyp 2]
                  [--ntracr 0] [--impl b200|reference]

One "step" = one tsadvc(m,n) call over all kdm layers of the named grid shape (BASELINE.json
configs[1]: GLBb0.08 4500x3298x41, T+S advem_fct2, synthetic fields).  `value` is measured with
every operand resident in HBM; `e2e` is the same call through the drop-in entry
hycom_tsadvc_step() on pinned HOST arrays, host<->device copies inside the timed region.
Under torchrun (N>1) the global grid is split into ipr x jpr tiles as mod_xc does, one rank
per GPU; every step is ONE C-ABI call per rank (hycom_tsadvc_step_device / hycom_tsadvc_step):
the halo exchange over NVLink (NCCL send/recv inside the library, overlapped with the tile
interior), the salinity-range reduction and the time-level switch all happen inside it.
torch.distributed only broadcasts the 128-byte NCCL id and takes the max of the timings.
`checksum` is a tiling-invariant hash of saln(:,:,:,n) after the last timed step (the analogue
of the reference's PIPE_CHECK, mod_pipe.F90:724-757): N = 1, 2, 4, 8 must print the same value.
`--impl reference` times the CPU oracle (the reference's algorithm restated in C, OpenMP over j
like the reference's relo_omp build; the Fortran itself cannot be compiled in this image) on a
bounded sample, with inputs from a CUDA-free build of the generator: that arm maps nothing of
the product library.
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
    per = (72 + 24 * ntracr) if advtyp in (2, 4) else (56 + 16 * ntracr)
    return idm * jdm * kk * per + idm * jdm * 28


def workload_name(workload, idm, jdm, kdm, advtyp, ntracr):
    return f"{workload} {idm}x{jdm}x{kdm} T+S advtyp={advtyp} ntracr={ntracr}"


def config_block(workload, idm, jdm, kdm, advtyp, ntracr, ipr, jpr, tile, temdf2, alg=None, reduced=False):
    """the same keys in both arms (the driver compares them)"""
    return {"workload": workload_name(workload, idm, jdm, kdm, advtyp, ntracr)
            + (" (REDUCED kdm: profiling run, not a bench value)" if reduced else ""),
            "tiling": f"{ipr}x{jpr}", "tile": tile, "nreg": 0,
            "l2": "inputs per step (%.1f GB) exceed L2 (126 MB); no flush" % ((alg or 0) / 1e9),
            "diag": "salinity min/max every 3rd step as mod_tsadvc.F90:2065",
            "temdf2": temdf2}


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
def _oracle_libs():
    odir = os.path.join(ROOT, "oracle", "_build")
    need = [os.path.join(odir, n) for n in ("liboracle_fast.so", "libsynth_host.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return need


class CpuOracle:
    """`nlay` layers of the named shape in the CPU oracle, inputs from the CUDA-free generator"""

    def __init__(self, shape, advtyp, ntracr, nlay):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import numpy as np
        import oracle_binding
        import util
        pkg, syn, cabi = util.pkg, util.syn, util.cabi
        fast, synth = _oracle_libs()
        syn.use_host_library(synth)
        orc = oracle_binding.Oracle(fast)
        idm, jdm, kdm, baclin, dx = syn.SHAPES[shape]
        self.cells = idm * jdm * nlay
        cfg = syn.make_cfg(idm, jdm, kdm, nreg=0, ntracr=ntracr, seed=1, dx0=dx, delt1=2.0 * baclin)
        sea = syn.sea_mask(cfg)
        g = pkg.partition(idm, jdm, nlay, 1, 1, 0)[0]
        cb = syn.build_cb_arrays(cfg, g, sea, 1, 2, with_state=False, advtyp=advtyp)
        cb.ntracr = ntracr
        k0 = 1 if nlay == kdm else max(1, kdm // 2)

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
        self.ot = util.oracle_tile_from_cb(orc, cb, sea)
        self.cb, self.sea, self.geom = cb, sea, g
        self.calls = 0

    def run(self, calls, threads):
        """seconds of each of `calls` tsadvc(m,n) calls with `threads` OpenMP threads"""
        self.ot.set_i("nthreads", threads)
        out = []
        for _ in range(calls):
            m, n = (1, 2) if self.calls % 2 == 0 else (2, 1)
            t0 = time.perf_counter()
            self.ot.tsadvc(m, n, 1)
            out.append(time.perf_counter() - t0)
            self.calls += 1
        return out

    def close(self):
        self.ot.close()


def reference_text_rate(o, calls, threads):
    """the REFERENCE'S OWN SOURCE TEXT, compiled (oracle/fortran_to_c.py: mod_tsadvc.F90, bigrid.F90 and xctilr of
    mod_xc_sm.h translated statement by statement to C, `!$OMP PARALLEL DO ... SCHEDULE(STATIC,jblk)` carried over;
    oracle/_ref/libref_text_*_omp.so, built where /root/reference exists and shipped with the snapshot), on the
    sample the port was just timed on: seconds per tsadvc(m,n) call"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_text as rt
    import reference_text_c as rc
    so = rc.RefTextC.so_path(6, False, True)
    if not rt.available() and not (os.path.exists(so) and os.path.exists(so[:-3] + ".json")):
        raise FileNotFoundError("oracle/_ref holds no compiled reference text (build() makes it where /root/reference exists)")
    lib = rc.RefTextC(6, False, openmp=True, flags=rc.RefTextC.TIMED_FLAGS)
    import ctypes
    # (torchrun exports OMP_NUM_THREADS=1 to its ranks; the directives of the text carry no num_threads clause)
    ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
    g, cb, sea = o.geom, o.cb, o.sea
    nb = g.nbdy
    env = rt.make_env(g.ii, g.jj, g.kdm)
    env["jblk"] = (g.jj + 2 * nb + threads - 1) // threads            # mod_dimensions.F90:143
    depth = np.zeros((g.nrows, g.ncols))
    depth[nb:nb + g.jj, nb:nb + g.ii] = np.where(sea != 0, 100.0, 0.0)
    u = [np.zeros_like(depth) for _ in range(3)]
    lib.run(env, "bigrid", depth, 0, *u)
    cb.nstep = 1
    rt.add_cb_arrays(env, cb)
    out = []
    for c in range(calls + 1):
        m, n = (1, 2) if c % 2 == 0 else (2, 1)
        t0 = time.perf_counter()
        lib.run(env, "tsadvc", m, n)
        out.append(time.perf_counter() - t0)
    return out[1:]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    syn = importlib.import_module("hycom-src_b200").synthetic
    idm, jdm, kdm, _, _ = syn.SHAPES[args.workload]
    nlay = args.cpu_layers
    cores = len(os.sched_getaffinity(0))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ipr, jpr = TILINGS.get(world, (1, 1))
    o = CpuOracle(args.workload, args.advtyp, args.ntracr, nlay)
    o.run(1, cores)                                   # page faults of the scratch slabs
    per_call = o.run(args.warmup + args.steps, cores)
    timed = per_call[args.warmup:]
    sec = sum(timed) / len(timed)
    value = idm * jdm * nlay / sec
    one = o.run(2, 1)[-1]                             # relo_one analogue: one thread, same sample
    extra = {"one_thread": {"value": idm * jdm * nlay / one, "unit": UNIT, "cores": 1,
                            "sample": f"{nlay} of {kdm} layers, 1 timed call after 1 warm-up"}}
    kind, impl_note = "port", "C oracle -O2 -fopenmp schedule(static,jblk)"
    port_value = value
    try:
        tt = reference_text_rate(o, min(args.steps, 5), cores)
        tsec = sum(tt) / len(tt)
        extra["reference_text"] = {"value": idm * jdm * nlay / tsec, "unit": UNIT, "cores": cores, "kind": "reference",
                                   "ms_per_step": tsec * 1e3,
                                   "sample": f"{nlay} of {kdm} layers, {len(tt)} timed tsadvc calls after 1 warm-up: mod_tsadvc.F90 "
                                             f"as written, translated statement by statement to C with its OpenMP directives "
                                             f"(oracle/fortran_to_c.py), gcc -O2 -march=x86-64-v3 -fopenmp"}
        extra["port"] = {"value": port_value, "unit": UNIT, "cores": cores, "kind": "port"}
        if idm * jdm * nlay / tsec > value:           # the line reports the FASTER of the two CPU arms
            value, sec, kind = idm * jdm * nlay / tsec, tsec, "reference"
            impl_note = "the reference text compiled (oracle/_ref), gcc -O2 -fopenmp schedule(static,jblk)"
    except Exception as e:  # noqa: BLE001
        extra["reference_text"] = {"skipped": repr(e)[:300]}
    o.close()
    if not args.no_full_kdm:
        try:
            import psutil
            need = (2 * 2 * 2 + 4) * kdm * (idm + 12) * (jdm + 12) * 8 * 1.6      # state + the 16 scratch slabs
            if psutil.virtual_memory().available > need:
                of = CpuOracle(args.workload, args.advtyp, args.ntracr, kdm)
                of.run(1, cores)
                tf = of.run(1, cores)[0]
                of.close()
                extra["full_kdm_call"] = {"value": idm * jdm * kdm / tf, "unit": UNIT, "cores": cores,
                                          "ms": tf * 1e3, "sample": f"one timed call over all {kdm} layers after 1 warm-up"}
            else:
                extra["full_kdm_call"] = {"skipped": "not enough free host memory"}
        except Exception as e:  # noqa: BLE001
            extra["full_kdm_call"] = {"skipped": repr(e)[:200]}
    sample = (f"{nlay} of {kdm} layers of {args.workload} ({idm}x{jdm}) per step, "
              f"{len(timed)} timed tsadvc calls, {impl_note}, {cores} threads; the faster of the port and the compiled "
              f"reference text (both in extra)")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_block(args.workload, idm, jdm, kdm, args.advtyp, args.ntracr, ipr, jpr,
                               f"{idm // ipr}x{(jdm + jpr - 1) // jpr}", 0.0,
                               alg_bytes_per_call(idm // ipr, (jdm + jpr - 1) // jpr, kdm, args.advtyp, args.ntracr)),
        "step_is": f"one tsadvc call over {nlay} of the {kdm} layers (bounded sample; layers are independent): "
                   f"value = {nlay} layers' cells / ms_per_step",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "extra": extra,
    }
    print(json.dumps(out))
    return 0


# ----------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------
class Run:
    """one workload on this rank's tile: handle, device-resident synthetic state, timed steps"""

    def __init__(self, args, workload, advtyp, ntracr, temdf2=0.0, kdm_override=0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        pkg = importlib.import_module("hycom-src_b200")
        self.pkg, self.syn = pkg, pkg.synthetic
        self.cabi = importlib.import_module("hycom-src_b200.cabi")
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.workload, self.advtyp, self.ntracr, self.temdf2 = workload, advtyp, ntracr, temdf2
        self.ipr, self.jpr = TILINGS[self.world]
        idm, jdm, kdm, baclin, dx = self.syn.SHAPES[workload]
        if kdm_override:
            kdm = kdm_override
        self.idm, self.jdm, self.kdm = idm, jdm, kdm
        syn, cabi = self.syn, self.cabi
        self.cfg = syn.make_cfg(idm, jdm, kdm, nreg=0, ntracr=ntracr, seed=1, dx0=dx, delt1=2.0 * baclin)
        self.sea = syn.sea_mask(self.cfg)
        self.g = pkg.partition(idm, jdm, kdm, self.ipr, self.jpr, 0)[self.rank]
        self.cb = syn.build_cb_arrays(self.cfg, self.g, self.sea, 1, 2, with_state=False, advtyp=advtyp,
                                      trcflg=[0] * ntracr, temdf2=temdf2, temdfc=1.0, sigver=6)
        self.stream = torch.cuda.Stream()
        self.ts = pkg.Tsadvc(self.cb, device=self.local, stream=self.stream.cuda_stream)
        ts = self.ts
        self.xc = None
        if self.world > 1:
            if args.py_transport:       # comparison: torch.distributed moves the strips (round-1 path)
                self.xc = pkg.XcExchange(ts, dist, compute_stream=self.stream)
            else:                       # the library owns the communicator (mod_xc's role)
                ts.comm_init_nccl(dist)
                ts.set_overlap(not args.no_overlap)
        ts.set_deferred_range(not args.sync_range)
        # device-resident synthetic state, both leapfrog slots (dp too: the slots alternate)
        syn.fill_device(ts, self.cfg, self.sea, 1, 2, diffusion=temdf2 > 0.0)
        ts._ck(ts.lib.hycom_tsadvc_synth_fill(ts.h, cabi.C.byref(self.cfg), cabi.F_DP, 0, 1, 0, 1, float("nan")))
        ts.synchronize()
        self.nsteps = 0

    def one_step(self):
        # HYCOM_Run: m=mod(nstep,2)+1; n=mod(nstep+1,2)+1 (mod_hycom.F90:2254-2257)
        s = self.nsteps
        m, n = s % 2 + 1, (s + 1) % 2 + 1
        self.cb.nstep = s + 1
        ts = self.ts
        if self.xc is not None:
            with self.torch.cuda.stream(self.stream):
                self.xc.tsadvc_device(m, n, diag=True, overlap=not self.args.no_overlap)
        elif self.args.sync_range:
            ts.tsadvc_device(m, n, diag=True)
        else:
            # the range of the previous diagnostic step is fetched while this step is queued: the
            # step call itself never waits for the device
            ts.tsadvc_device(m, n, diag=False)
            if (s % 3) == 0:
                ts.saln_range()
        self.nsteps += 1
        return n

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxr(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, steps, warmup, sample_clocks=True):
        torch, ts = self.torch, self.ts
        for _ in range(warmup):
            self.one_step()
        self.barrier()
        ts.set_timing(True)
        ts.get_timing(reset=True)
        l0 = ts.launch_count
        sampler = ClockSampler(self.local)
        if self.rank == 0 and sample_clocks:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        n = 2
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
            for _ in range(steps):
                n = self.one_step()
            e1.record(self.stream)
        self.barrier()
        clocks = sampler.stop() if (self.rank == 0 and sample_clocks) else None
        ms_total = e0.elapsed_time(e1)
        march_ms, march_n = ts.get_timing(reset=True)
        ts.set_timing(False)
        launches = ts.launch_count - l0
        ms_step = self.maxr(ms_total) / steps
        # interior + frame launches of one step count as one marching pass over the tile
        march_avg = self.maxr(march_ms / steps)
        cells = self.idm * self.jdm * self.kdm
        peak, peak_src = _peaks()
        alg = alg_bytes_per_call(self.g.ii, self.g.jj, self.kdm, self.advtyp, self.ntracr)
        achieved = alg / (march_avg * 1e-3) / 1e9
        split = self.advtyp in (1, 2) and os.environ.get("HYCOM_TSADVC_SPLIT", "1") != "0"
        roofline = {"bound": "hbm",
                    "kernel": "k_tsadvc_march_tma" + (" (general + all-sea launch of one call)" if split else ""),
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "alg_bytes_per_launch": alg, "kernel_ms": march_avg, "peak_source": peak_src,
                    "kernel_share_of_step": march_avg / ms_step}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            roofline["traffic"] = tj.get(f"{self.workload}:advtyp{self.advtyp}:ntracr{self.ntracr}:gpus{self.world}")
        except Exception:
            pass
        # tiling-invariant hash of saln(:,:,:,n) after the last timed step (all ranks: collective)
        cks = ts.checksum(self.cabi.F_SALN, n) if self.xc is None else None
        return {"ms_step": ms_step, "value": cells / (ms_step * 1e-3), "roofline": roofline, "launches": int(launches),
                "clocks": clocks, "alg": alg, "checksum": cks, "total_steps": self.nsteps, "slot": n}

    def cnuity_timing(self, steps=3, warmup=1):
        """cnuity(m,n) (cnuity.F90, the producer of dp(n), uflx, vflx) on the same device-resident state: its
        operands are fabricated on the device from the synthetic generator (u, v of O(0.3 m/s) from the mass-flux
        fields, dpu = dpv = dp(m), no depth limiting) - a timing of the sweeps, not a parity case (those are
        tests/test_cnuity_gpu.py)."""
        torch, ts, cabi, cfg = self.torch, self.ts, self.cabi, self.cfg
        lib, C = ts.lib, cabi.C
        kk = self.kdm
        su = 1.0 / (cfg.dx0 * 150.0 * 9806.0)
        m, n = 1, 2

        def fill(gen, lev, scale, dst, tlev, nk, halo=1):
            ts._ck(lib.hycom_tsadvc_synth_fill_to(ts.h, C.byref(cfg), gen, 0, lev, halo, scale, dst, tlev, 1, nk))
        fill(cabi.F_UFLX, 0, su, cabi.F_U, m, kk)
        fill(cabi.F_VFLX, 0, su, cabi.F_V, m, kk)
        fill(cabi.F_DP, 1, 1.0, cabi.F_DPU, m, kk)
        fill(cabi.F_DP, 1, 1.0, cabi.F_DPV, m, kk)
        fill(cabi.F_UFLX, 0, 0.0, cabi.F_UBAVG, 1, 3)
        fill(cabi.F_VFLX, 0, 0.0, cabi.F_VBAVG, 1, 3)
        for dst in (cabi.F_DEPTHU, cabi.F_DEPTHV, cabi.F_PBOT):
            fill(cabi.S_ONETA, 0, 1.0e9, dst, 1, 1)
        for t in (1, 2):
            fill(cabi.F_DP, 0, 0.5, cabi.F_DPMIXL, t, 1)
        # coefficients of the interface-depth diffusion: thkdf4 = 0.01 m/s times scuy / scvx (forfun.F90:2541-2568)
        fill(cabi.S_ONETA, 0, 0.01 * cfg.dx0, cabi.F_THKDF4U, 1, 1)
        fill(cabi.S_ONETA, 0, 0.01 * cfg.dx0, cabi.F_THKDF4V, 1, 1)
        ts.synchronize()

        def timed(**kw):
            l0 = ts.launch_count
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for s in range(warmup + steps):
                if s == warmup:
                    self.barrier()
                    e0.record(self.stream)
                    l0 = ts.launch_count
                self.cb.nstep = s + 1
                ts.cnuity_device(m, n, **kw)
            e1.record(self.stream)
            self.barrier()
            return self.maxr(e0.elapsed_time(e1)) / steps, l0
        ms, l0 = timed()
        nl = int(ts.launch_count - l0)
        ms4, _ = timed(thkdf4=0.01)
        peak, _ = _peaks()
        alg = self.g.ii * self.g.jj * kk * 104
        return {"what": "cnuity(m,n) on the device mirrors (cnuity.F90), synthetic operands", "ms_per_call": ms,
                "ms_per_call_with_thkdf4": ms4,
                "value": self.idm * self.jdm * kk / (ms * 1e-3), "unit": UNIT, "steps": steps,
                "alg_bytes_per_call": alg, "alg_bytes_note": "104 B per layer-cell: read dp(n), dp(m), u, v, dpu, dpv; "
                "write dp(n), dp(m), dpo(n), dpo(m), uflx, vflx, p", "achieved_gbs": alg / (ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak, "gpu_launches": nl}

    def close(self):
        self.ts.close()


def _bind_to_gpu_numa_node(index):
    """run this rank (and first-touch its pinned host arrays) on the CPUs next to its GPU: with eight ranks
    uploading at once, host arrays on the far socket cap the per-rank H2D rate well below PCIe"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist

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
    numa = _bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    run = Run(args, args.workload, args.advtyp, args.ntracr, temdf2=args.temdf2, kdm_override=args.kdm)
    run.numa = numa
    r = run.timed(args.steps, args.warmup)
    e2e = None
    if not args.no_e2e and not args.py_transport:
        e2e = run_e2e(args, run)
    g, idm, jdm, kdm = run.g, run.idm, run.jdm, run.kdm
    cn = None
    if not args.no_extra and world == 1 and not args.kdm:
        try:
            cn = run.cnuity_timing()
        except Exception as e:  # noqa: BLE001
            cn = {"failed": repr(e)[:300]}
    run.close()
    del run
    torch.cuda.empty_cache()

    # the other configurations BASELINE.json names, each with its own roofline, as `extra`
    extra = {}
    if cn is not None:
        extra["cnuity"] = cn
    if not args.no_extra and args.workload == "GLBb0.08" and args.advtyp == 2 and args.ntracr == 0 and not args.kdm:
        todo = []
        if world == 1:
            todo.append(("config3_mpdata_8_tracers", "GLBb0.08", 1, 8))
        if world == 8:
            todo.append(("config5_GLBy0.04", "GLBy0.04", 2, 0))
        for key, wl, adv, ntr in todo:
            try:
                x = Run(args, wl, adv, ntr)
                xr = x.timed(3, 3, sample_clocks=False)
                if rank == 0:
                    extra[key] = {"workload": workload_name(wl, x.idm, x.jdm, x.kdm, adv, ntr), "steps": 3, "warmup": 3,
                                  "ms_per_step": xr["ms_step"], "value": xr["value"], "unit": UNIT,
                                  "roofline": xr["roofline"], "gpu_launches": xr["launches"],
                                  "checksum": "0x%016x" % xr["checksum"] if xr["checksum"] is not None else None,
                                  "tile": f"{x.g.ii}x{x.g.jj}"}
                x.close()
                del x
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                if rank == 0:
                    extra[key] = {"failed": repr(e)[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        o = CpuOracle(args.workload, args.advtyp, args.ntracr, args.cpu_layers)
        cores = len(os.sched_getaffinity(0))
        o.run(1, cores)
        t = o.run(2, cores)
        cpu = {"value": idm * jdm * args.cpu_layers / (sum(t) / len(t)), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_layers} of {kdm} layers of {args.workload}, 2 timed tsadvc calls after 1 "
                         f"warm-up, C oracle (gcc -O2 -fopenmp, schedule(static,jblk) over j)"}
        try:       # the reference's own text compiled (oracle/_ref), same sample; the faster of the two is the baseline
            tt = reference_text_rate(o, 2, cores)
            tv = idm * jdm * args.cpu_layers / (sum(tt) / len(tt))
            cpu["port"], cpu["reference_text"] = cpu["value"], tv
            if tv > cpu["value"]:
                cpu.update(value=tv, kind="reference",
                           sample=f"{args.cpu_layers} of {kdm} layers of {args.workload}, 2 timed tsadvc calls after 1 warm-up, "
                                  f"mod_tsadvc.F90 as written compiled through oracle/fortran_to_c.py (gcc -O2 -fopenmp, "
                                  f"schedule(static,jblk) over j)")
        except Exception as e:  # noqa: BLE001
            cpu["reference_text"] = {"skipped": repr(e)[:200]}
        o.close()
    if rank == 0:
        peak, _ = _peaks()
        out = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(args.workload, idm, jdm, kdm, args.advtyp, args.ntracr, TILINGS[world][0],
                                   TILINGS[world][1], f"{g.ii}x{g.jj}", args.temdf2, r["alg"], reduced=bool(args.kdm)),
            "roofline": r["roofline"], "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": r["launches"],
            "clocks": r["clocks"],
            "pct_of_hbm_roofline": 100.0 * r["alg"] / (r["ms_step"] * 1e-3) / 1e9 / peak,
            "checksum": {"field": f"saln(:,:,:,{r['slot']}) interior sea points, all tiles",
                         "value": "0x%016x" % r["checksum"] if r["checksum"] is not None else None,
                         "after_steps": r["total_steps"],
                         "how": "sum mod 2**64 of mix64(bits ^ mix64(global cell)): hycom_tsadvc_checksum"},
            "transport": ("torch.distributed P2P (host-owned)" if args.py_transport else
                          "NCCL send/recv inside libhycom_tsadvc_b200.so, one C-ABI call per step") if world > 1 else None,
            "extra": extra or None,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e(args, run):
    """tsadvc(m,n) through hycom_tsadvc_step on pinned host arrays: H2D of every operand, the halo
    exchange (N>1: inside the call), compute, D2H of the advected fields, every step."""
    import numpy as np
    import torch
    cabi, g, cb, ts_dev = run.cabi, run.g, run.cb, run.ts
    kk = g.kdm
    P = g.nrows * g.ncols
    nt = run.ntracr

    def pinned(shape):
        t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
        return t, t.numpy()
    keep = []
    # generate on the device (fast), copy into the host arrays once: same synthetic state
    syn = run.syn
    syn.fill_device(ts_dev, run.cfg, run.sea, 1, 2)
    ts_dev._ck(ts_dev.lib.hycom_tsadvc_synth_fill(ts_dev.h, cabi.C.byref(run.cfg), cabi.F_DP, 0, 1, 0, 1, float("nan")))

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
    ts_dev.set_deferred_range(False)
    times = []
    for s in range(warm + steps):
        m, n = s % 2 + 1, (s + 1) % 2 + 1
        cb.nstep = s + 1
        run.barrier()
        t0 = time.perf_counter()
        ts_dev.tsadvc(m, n)
        ts_dev.synchronize()
        dt = time.perf_counter() - t0
        if s >= warm:
            times.append(dt)
    sec = run.maxr(sum(times) / len(times))
    return {"value": run.idm * run.jdm * kk / sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": sec * 1e3, "steps": steps,
            "h2d_gbs_per_rank": h2d / sec / 1e9, "cpus_near_gpu": getattr(run, "numa", None),
            "api": "hycom_tsadvc_step (pinned host arrays, Fortran layout"
                   + (", NCCL halo exchange inside the call)" if run.world > 1 else ")")}


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
    ap.add_argument("--no-extra", action="store_true", help="skip the `extra` configurations (config 3 / GLBy0.04)")
    ap.add_argument("--no-full-kdm", action="store_true", help="reference arm: skip the one full-kdm call")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: exchange first, then the whole tile")
    ap.add_argument("--py-transport", action="store_true",
                    help="N>1: the round-1 path, torch.distributed moves the strips (comparison)")
    ap.add_argument("--sync-range", action="store_true",
                    help="the step waits for the salinity range every 3rd step (default: deferred fetch)")
    ap.add_argument("--temdf2", type=float, default=0.0,
                    help="> 0: the step also runs tsdff_1x/2x + the EOS sweep (mod_tsadvc.F90:2138-2230); "
                         "the BASELINE metric is quoted without it")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    b = importlib.import_module("hycom-src_b200.build")
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        b.build_library()
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
