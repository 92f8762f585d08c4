#!/usr/bin/env python
"""Parity of the multi-GPU path under real NCCL (run with torchrun, one rank per GPU):
tsadvc(m,n) on ipr x jpr tiles through the library's own communicator (hycom_tsadvc_comm_init;
pack -> ncclSend/ncclRecv -> unpack overlapped with the tile interior, all inside
hycom_tsadvc_step_device / hycom_tsadvc_step) against the CPU oracle on ONE tile, bit for bit, two
leapfrog steps, plus the tiling-invariant checksum against its numpy restatement on the oracle's
output.  XC_CHECK_PY=1: the host-owned transport (hycom-src_b200/xc.py) instead.
The oracle is test infrastructure: this script is a checker, not a product path."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle_binding
    import util
    pkg, syn, cabi = util.pkg, util.syn, util.cabi
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ipr, jpr = {2: (2, 1), 4: (2, 2), 8: (4, 2)}[world]
    if os.environ.get("XC_CHECK_TILES"):          # e.g. "4x1": another tiling of the same world size
        ipr, jpr = (int(v) for v in os.environ["XC_CHECK_TILES"].split("x"))
        assert ipr * jpr == world
    ok_all = True
    cases = [(150, 150, 4, 0, 0, 2, {}), (301, 203, 3, 3, 1, 2, {}), (180, 120, 2, 1, 1, 1, {}),
             # temdf2 > 0: second exchange (width 2) + tsdff + EOS; btrmas: advem_fct2c with five exchanges
             (150, 150, 3, 0, 1, 2, {"diffusion": (6, 1.0)}), (301, 203, 2, 3, 0, 2, {"diffusion": (8, 0.5)}),
             (150, 150, 10, 0, 1, 2, {"btrmas": True}), (301, 203, 2, 3, 0, 2, {"btrmas": True}),
             # nreg=2: the top row exchanges the tripole fold with its twin tiles (mod_xc_mp.h:4114-4662)
             (192, 120, 3, 2, 1, 2, {"arctic": True}), (256, 96, 2, 2, 0, 1, {"arctic": True}),
             (256, 96, 2, 2, 0, 2, {"arctic": True, "diffusion_arctic": True}),
             # isopyc: layer 1 on laterally smoothed fluxes, which read the halo (no interior overlap)
             (150, 150, 3, 0, 0, 2, {"isopyc": True}),
             # ... with a tracer: layer 1 advected by uflx with the prolog of the smoothed fluxes (ten-array ring)
             (150, 150, 3, 0, 1, 1, {"isopyc": True}),
             # the drop-in entry on HOST arrays: upload, exchange per layer chunk, advect, download
             (150, 150, 5, 0, 1, 2, {"host": True}), (150, 150, 3, 0, 0, 2, {"host": True, "diffusion": (6, 1.0)})]
    py_transport = os.environ.get("XC_CHECK_PY", "0") == "1"
    if py_transport:
        cases = [c for c in cases if "host" not in c[6]]
    only = os.environ.get("XC_CHECK_CASES", "")       # e.g. "arctic": run the cases carrying that key only
    if only:
        cases = [c for c in cases if only in c[6]]
    for (itdm, jtdm, kdm, nreg, ntracr, advtyp, extra) in cases:
        m, n = 1, 2
        scal = dict(advtyp=advtyp)
        if "btrmas" in extra:
            scal["btrmas"] = True
        if "isopyc" in extra:
            scal.update(isopyc=True, hybrid=False, nhybrd=0)
        if "diffusion_arctic" in extra:
            scal.update(temdf2=0.02, temdfc=1.0, sigver=6, thbase=34.0)
        if "diffusion" in extra:
            sigver, temdfc = extra["diffusion"]
            cfg, sea, g1, cb1 = util.make_diffusion_case(itdm, jtdm, kdm, sigver, temdfc, nreg=nreg, ntracr=ntracr,
                                                         seed=9, **scal)
            scal.update(temdf2=cb1.temdf2, temdfc=temdfc, sigver=sigver, thbase=cb1.thbase)
        elif "arctic" in extra:
            cfg, sea, g1, cb1 = util.make_arctic_case(itdm, jtdm, kdm, ntracr=ntracr, seed=9, m=m, n=n, **scal)
        else:
            cfg, sea, g1, cb1 = util.make_case(itdm, jtdm, kdm, nreg=nreg, ntracr=ntracr, seed=9, m=m, n=n, **scal)
        orc = oracle_binding.Oracle(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))
        ot = util.oracle_tile_from_cb(orc, cb1, sea)
        g = pkg.partition(itdm, jtdm, kdm, ipr, jpr, nreg)[rank]
        if "arctic" in extra:   # dp, oneta and the metrics arrive with the fold in their halo
            cb = util.make_arctic_tiles(cfg, sea, cb1, ipr, jpr, m, n, **scal)[rank]
        else:
            cb = syn.build_cb_arrays(cfg, g, sea, m, n, **scal)
        if "diffusion" in extra or "isopyc" in extra:   # this tile's window of the single-tile th3d/theta (both slots)
            nbd = g.nbdy
            win = (Ellipsis, slice(g.j0, g.j0 + g.nrows), slice(g.i0, g.i0 + g.ncols))
            cb.th3d = np.full((2, kdm, g.nrows, g.ncols), np.nan)
            cb.theta = np.full((kdm, g.nrows, g.ncols), np.nan)
            a = cb1.th3d[win]
            cb.th3d[..., :a.shape[-2], :a.shape[-1]] = a
            if cb1.theta is not None:
                b = cb1.theta[win]
                cb.theta[..., :b.shape[-2], :b.shape[-1]] = b
            else:
                cb.theta = None
        stream = torch.cuda.Stream()
        ts = pkg.Tsadvc(cb, device=local, stream=stream.cuda_stream)
        ts.upload_state(m, n)
        ts.upload(cabi.F_DP, cb.dp[m - 1], m)
        ts.upload(cabi.F_ONETA, cb.oneta[m - 1], m)
        ts.upload(cabi.F_ONETA, cb.oneta[n - 1], n)
        if "diffusion" in extra or "isopyc" in extra:
            ts.upload(cabi.F_TH3D, cb.th3d[m - 1], m)
            ts.upload(cabi.F_TH3D, cb.th3d[n - 1], n)
        xc = None
        if py_transport:
            xc = pkg.XcExchange(ts, dist, compute_stream=stream)
        else:
            ts.comm_init_nccl(dist)
        nb = g.nbdy
        if "host" in extra:
            os.environ["HYCOM_TSADVC_STEP_CHUNK"] = "2"
        for step, (mm, nn) in enumerate([(m, n), (n, m)]):
            cb.nstep = 3 * (step + 1)
            ot.set_i("nstep", cb.nstep) if hasattr(ot, "set_i") else None
            ot.tsadvc(mm, nn, 1)
            if "host" in extra:
                if step == 1:
                    break
                ts.tsadvc(mm, nn)
            elif xc is not None:
                with torch.cuda.stream(stream):
                    xc.tsadvc_device(mm, nn, diag=True, overlap=(step == 0))
            else:
                ts.set_overlap(step == 0)
                ts.tsadvc_device(mm, nn)
            ts.synchronize()
            if xc is None and not (np.array_equal(ts.xmin, ot.f64("xmin")) and np.array_equal(ts.xmax, ot.f64("xmax"))):
                ok_all = False
                print(f"rank {rank}: MISMATCH xmin/xmax step {step} case {(itdm, jtdm, nreg, advtyp, extra)}", flush=True)
            sea_t = cb.ip[nb:nb + g.jj, nb:nb + g.ii] != 0
            glob = (slice(None), slice(nb + g.j0, nb + g.j0 + g.jj), slice(nb + g.i0, nb + g.i0 + g.ii))
            flds = ((cabi.F_TH3D, "th3d"), (cabi.F_SALN, "saln")) if "isopyc" in extra else ((cabi.F_TEMP, "temp"), (cabi.F_SALN, "saln")) + (((cabi.F_TH3D, "th3d"),) if ("diffusion" in extra or "diffusion_arctic" in extra) else ())
            for fld, name in flds:
                src = getattr(cb, name)[nn - 1] if "host" in extra else ts.download(fld, nn)
                dev = src[:, nb:nb + g.jj, nb:nb + g.ii]
                ref = ot.f64(name)[nn - 1][glob]
                ok = np.array_equal(dev[:, sea_t], ref[:, sea_t])
                ok_all = ok_all and ok
                if not ok:
                    print(f"rank {rank}: MISMATCH {name} step {step} case {(itdm, jtdm, nreg, advtyp, extra)}", flush=True)
            for q in range(ntracr if "host" not in extra else 0):
                dev = ts.download(cabi.F_TRACER, nn, ktr=q + 1)[:, nb:nb + g.jj, nb:nb + g.ii]
                ref = ot.f64("tracer")[q, nn - 1][glob]
                ok = np.array_equal(dev[:, sea_t], ref[:, sea_t])
                ok_all = ok_all and ok
                if not ok:
                    print(f"rank {rank}: MISMATCH tracer {q} step {step} case {(itdm, jtdm, nreg, advtyp, extra)}", flush=True)
            if xc is None and "host" not in extra:   # PIPE_CHECK analogue: one number for all tiles
                import test_comm_gpu
                want = test_comm_gpu.np_checksum(ot.f64("saln")[nn - 1], cb1.ip, g1, kdm)
                got = ts.checksum(cabi.F_SALN, nn)
                if got != want:
                    ok_all = False
                    print(f"rank {rank}: CHECKSUM {got:#x} != {want:#x} case {(itdm, jtdm, nreg, advtyp, extra)}", flush=True)
        ts.close()
        ot.close()
    t = torch.tensor([1 if ok_all else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("xc_nccl_check:", "PASS" if int(t.item()) == 1 else "FAIL", f"(world {world}, tiles {ipr}x{jpr})", flush=True)
    dist.destroy_process_group()
    return 0 if int(t.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
