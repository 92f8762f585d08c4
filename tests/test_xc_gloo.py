"""The N>1 path on CPU: the product's exchange schedule (hycom-src_b200/xc.py) driven over
torch.distributed/gloo with world_size 2 and 4, numpy doing the pack/unpack.  The data are a
function of the GLOBAL (i,j,k) - the reference's own way of checking decompositions
(mod_pipe.F90:26-127) - so after xctilr every halo cell must hold the value of the global
cell it images (periodic wrap) or vland = 0 beyond a closed edge (mod_xc_sm.h:1377-1422)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from util import pkg


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _field(ig, jg, k, a):
    return 1000.0 * a + 100.0 * k + ig + 0.001 * jg


def _expected(g, a, kk, mh=5, nh=5, itype=1):
    """what xctilr must leave in the tile (interior + halo of width 5; NaN elsewhere).  nreg=2: rows
    above jtdm image the tripole fold of the global array (mod_xc_sm.h:1215-1320 / mod_xc_mp.h:4263-
    4372): mirrored in i (shifted by one column on the u grid), rows jtdm-1-j (p,u) or jtdm-j (v),
    vector fields with the sign flipped"""
    nb = g.nbdy
    out = np.full((kk, g.nrows, g.ncols), np.nan)
    grid = itype % 10
    for r in range(nb - nh, nb + g.jj + nh):
        for c in range(nb - mh, nb + g.ii + mh):
            ig, jg = g.i0 + c + 1 - nb, g.j0 + r + 1 - nb
            sgn = 1.0
            if g.periodic_i:
                ig = (ig - 1) % g.itdm + 1
            if g.periodic_j:
                jg = (jg - 1) % g.jtdm + 1
            if g.nreg == 2 and jg > g.jtdm:
                j = jg - g.jtdm
                ig = g.itdm - (ig - 1) % g.itdm if grid in (1, 4) else (g.itdm - (ig - 1)) % g.itdm + 1
                jg = g.jtdm - 1 - j if grid in (1, 3) else g.jtdm - j
                sgn = -1.0 if itype > 10 else 1.0
            inside = 1 <= ig <= g.itdm and 1 <= jg <= g.jtdm
            for k in range(kk):
                out[k, r, c] = sgn * _field(ig, jg, k, a) if inside else 0.0
    return out


def _worker(rank, world, port, itdm, jtdm, ipr, jpr, nreg, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import np_halo
        kk, narr = 2, (5 if nreg == 2 else 3)
        g = pkg.partition(itdm, jtdm, kk, ipr, jpr, nreg)[rank]
        nb = g.nbdy
        arrays = []
        for a in range(narr):
            arr = np.full((kk, g.nrows, g.ncols), np.nan)
            for k in range(kk):
                jj_, ii_ = np.meshgrid(np.arange(1, g.jj + 1), np.arange(1, g.ii + 1), indexing="ij")
                arr[k, nb:nb + g.jj, nb:nb + g.ii] = _field(g.i0 + ii_, g.j0 + jj_, k, a)
            arrays.append(arr)
        # halo_ps, halo_uv, halo_vv of tsadvc; halo_us, halo_vs as cnuity exchanges dpu, dpv (cnuity.F90:102-103)
        itypes = [1, 13, 14, 3, 4] if nreg == 2 else [1, 1, 1]
        be = np_halo.NumpyHaloBackend(g, arrays, itypes)
        ex = pkg.XcExchange(None, dist, backend=be)
        for _ in range(2):               # twice: buffers are reused
            ex.xctilr(1, 2)
        ok = True
        for a in range(narr):
            exp = _expected(g, a, kk, itype=itypes[a])
            live = ~np.isnan(exp)
            ok = ok and np.array_equal(arrays[a][live], exp[live])
            # the sixth halo line is not touched by a width-5 exchange
            ok = ok and np.isnan(arrays[a][~live]).all()
        # xcminr/xcmaxr
        lo, hi = np.array([float(rank), 5.0]), np.array([float(rank), 5.0])
        ex.xcminmax(lo, hi)
        ok = ok and lo[0] == 0.0 and hi[0] == world - 1
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _worker_staged(rank, world, port, itdm, jtdm, ipr, jpr, nreg, q):
    """the schedule of the diffusion exchange (width 2) and of the per-batch fct2c exchanges"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import np_halo
        kk = 3
        g = pkg.partition(itdm, jtdm, kk, ipr, jpr, nreg)[rank]
        nb = g.nbdy

        def make(a, layers):
            arr = np.full((layers, g.nrows, g.ncols), np.nan)
            jj_, ii_ = np.meshgrid(np.arange(1, g.jj + 1), np.arange(1, g.ii + 1), indexing="ij")
            for k in range(layers):
                arr[k, nb:nb + g.jj, nb:nb + g.ii] = _field(g.i0 + ii_, g.j0 + jj_, k, a)
            return arr
        diff = [make(a, kk) for a in (10, 11, 12)]
        batches = {0: [make(20, 2), make(21, 2)], 1: [make(22, 1), make(23, 1)]}   # 2 + 1 layers
        be = np_halo.NumpyStagedBackend(g, [], diff, batches)
        ex = pkg.XcExchange(None, dist, backend=be)
        ok = True

        def check(arrays, ids, w):
            good = True
            for arr, a in zip(arrays, ids):
                exp = _expected(g, a, arr.shape[0], w, w)
                live = ~np.isnan(exp)
                good = good and np.array_equal(arr[live], exp[live]) and np.isnan(arr[~live]).all()
            return good
        for _ in range(2):
            ex._exchange(("diff", 2), lambda: be.diff_counts(2), lambda s, st: be.diff_pack(2, s, st),
                         lambda r, st: be.diff_unpack(2, r, st))
        ok = ok and check(diff, (10, 11, 12), 2)
        for b in (0, 1):
            for _ in range(2):
                ex._exchange(("fct2c", b), lambda: be.fct2c_counts(1, 2, b), lambda s, st: be.fct2c_pack(1, 2, b, s, st),
                             lambda r, st: be.fct2c_unpack(1, 2, b, r, st))
        ok = ok and check(batches[0], (20, 21), 5) and check(batches[1], (22, 23), 5)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,itdm,jtdm,ipr,jpr,nreg", [(2, 57, 41, 2, 1, 0), (2, 44, 36, 1, 2, 3), (4, 45, 38, 2, 2, 3),
                                                          (4, 48, 38, 2, 2, 2)])
def test_staged_exchanges_over_gloo(world, itdm, jtdm, ipr, jpr, nreg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_staged, args=(r, world, port, itdm, jtdm, ipr, jpr, nreg, q))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert all(res.values()), res


CASES = [
    (2, 57, 41, 2, 1, 0),   # BASELINE configs[3] at 2 GPUs: 2x1 tiles, closed basin, ragged split
    (2, 40, 37, 1, 2, 1),   # 1x2 tiles, periodic in i: E/W wrap onto the tile itself
    (2, 44, 36, 2, 1, 3),   # 2x1 doubly periodic: both E and W neighbour are the other rank
    (4, 45, 38, 2, 2, 3),   # 2x2 doubly periodic: every neighbour pair exchanges 4 messages
    (4, 61, 33, 4, 1, 0),   # 4x1 closed
    (2, 44, 36, 2, 1, 2),   # arctic, 2x1: each tile's twin is the other rank, NW/NE fold onto itself
    (4, 48, 38, 2, 2, 2),   # arctic, 2x2: fold in the top row, closed south
    (4, 64, 30, 4, 1, 2),   # arctic, 4x1: twins 0-3 and 1-2, shifted u-grid column from a third tile
    (2, 40, 36, 1, 2, 2),   # arctic, 1x2: the top tile is its own twin
]


@pytest.mark.parametrize("world,itdm,jtdm,ipr,jpr,nreg", CASES)
def test_xctilr_over_gloo(world, itdm, jtdm, ipr, jpr, nreg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, itdm, jtdm, ipr, jpr, nreg, q))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert all(res.values()), res


def test_neighbour_tables():
    xc = __import__("importlib").import_module("hycom-src_b200.xc")
    tiles = pkg.partition(100, 80, 1, 4, 2, 0)
    t = tiles[0]                      # SW corner of a closed basin
    assert xc.neighbors(t) == [-1, 1, -1, 4, -1, -1, -1, 5]
    t = tiles[5]                      # interior column, top row
    assert xc.neighbors(t) == [4, 6, 1, -1, 0, 2, -1, -1]
    tiles = pkg.partition(100, 80, 1, 4, 2, 1)   # periodic in i
    assert xc.neighbors(tiles[0]) == [3, 1, -1, 4, -1, -1, 7, 5]
    # arctic: the top row faces its twins (N = twin, NW/NE = twins of the W/E neighbours), the fold
    # directions pair with themselves, the bottom row is closed to the south
    tiles = pkg.partition(96, 60, 1, 4, 2, 2)
    assert xc.neighbors(tiles[4]) == [7, 5, 0, 7, 3, 1, 4, 6]
    assert xc.neighbors(tiles[5]) == [4, 6, 1, 6, 0, 2, 7, 5]
    assert xc.neighbors(tiles[1]) == [0, 2, -1, 5, -1, -1, 4, 6]
    assert [xc.opp_dir(tiles[4], d) for d in range(8)] == [1, 0, 3, 3, 7, 6, 6, 7]
    assert [xc.opp_dir(tiles[1], d) for d in range(8)] == list(xc.OPP)
    for t in tiles:
        cnt, nbr = xc.halo_counts(t, 3), xc.neighbors(t)
        for d in range(8):
            if nbr[d] >= 0:
                assert cnt[d] == xc.halo_counts(tiles[nbr[d]], 3)[xc.opp_dir(t, d)]
                assert xc.neighbors(tiles[nbr[d]])[xc.opp_dir(t, d)] == t.mproc - 1 + t.ipr * (t.nproc - 1)
    # message sizes: send size of a tile == receive size of its neighbour
    tiles = pkg.partition(103, 81, 1, 4, 2, 3)
    for t in tiles:
        cnt, nbr = xc.halo_counts(t, 7), xc.neighbors(t)
        for d in range(8):
            assert cnt[d] == xc.halo_counts(tiles[nbr[d]], 7)[xc.opp_dir(t, d)]
