"""Multi-tile halo exchange of the tsadvc path: the host side of ``xctilr``.

mod_xc (mod_xc_mp.h:4664-4987) owns the communicator and moves the packed edge strips with
MPI/SHMEM.  The product path keeps the communicator INSIDE the C library (csrc/xc_comm.cu:
``Tsadvc.comm_init_nccl`` + one ``hycom_tsadvc_step_device`` call per step).  This module is the
alternative a host program uses when it wants to own the byte moving: one process drives one GPU
(= one tile, placed exactly as mod_xc places tiles: ``geometry.partition``), the edge strips are
packed and unpacked on the device by the C library (``hycom_tsadvc_halo_pack/unpack``) and travel
between GPUs as NCCL send/recv over NVLink issued by ``torch.distributed``.  All eight neighbours are addressed in one
round, and the exchange overlaps the interior of the tile:

    comm stream   : pack -> send/recv -> unpack
    compute stream: march(PART_INTERIOR) ................ wait -> march(PART_FRAME)

``XcExchange`` holds the schedule (who sends what to whom, in which order); the byte moving
is behind a small backend so that the schedule is also exercised on CPU ranks (gloo) by the
test-suite with a numpy backend.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

from . import cabi
from .geometry import TileGeom

# direction d -> (dx, dy); opposite direction
DIR_DXY = ((-1, 0), (1, 0), (0, -1), (0, 1), (-1, -1), (1, -1), (-1, 1), (1, 1))
OPP = (1, 0, 3, 2, 7, 6, 5, 4)


def arctic_fold(g: TileGeom) -> bool:
    """a tile of the top row of a multi-tile global grid across the arctic (nreg=2): its northern
    exchange is the tripole fold with the twin tiles (mod_xc_mp.h:2830, :4263-4428)"""
    return g.nreg == 2 and g.ipr * g.jpr > 1 and g.nproc == g.jpr


def opp_dir(g: TileGeom, d: int) -> int:
    """the direction in which the peer sent the message that arrives from direction d: the
    opposite one, except across the fold where N, NW, NE pair with themselves (both twins look
    north at each other)"""
    return d if arctic_fold(g) and DIR_DXY[d][1] > 0 else OPP[d]


def neighbors(g: TileGeom) -> List[int]:
    """0-based tile index (mproc-1 + ipr*(nproc-1)) of the eight neighbours, -1 at a closed
    edge (mod_xc.F90:25-31: nreg 1,3 periodic in i; nreg 3,4 periodic in j).  Across the arctic
    the top row faces its twins: idproc(m,jpr+1) = idproc(ipr+1-m,jpr) (mod_xc_mp.h:2830)."""
    out = []
    for dx, dy in DIR_DXY:
        if arctic_fold(g) and dy > 0:
            out.append((g.ipr - 1 - (g.mproc - 1 + dx) % g.ipr) + g.ipr * (g.nproc - 1))
            continue
        mp, np_ = g.mproc - 1 + dx, g.nproc - 1 + dy
        if not 0 <= mp < g.ipr:
            if not g.periodic_i:
                out.append(-1)
                continue
            mp %= g.ipr
        if not 0 <= np_ < g.jpr:
            if not g.periodic_j:
                out.append(-1)
                continue
            np_ %= g.jpr
        out.append(mp + g.ipr * np_)
    return out


def halo_counts(g: TileGeom, nslab: int, mh: int = 5, nh: int = 5) -> List[int]:
    """doubles per direction of one message of ``nslab`` slabs (fold messages: all ii columns to
    the twin, mh+1 columns north-west and mh north-east, see csrc/tsadvc_launch.h)"""
    out = []
    for dx, dy in DIR_DXY:
        w = g.ii if dx == 0 else mh
        h = g.jj if dy == 0 else nh
        if arctic_fold(g) and dy > 0 and dx < 0:
            w = mh + 1
        out.append(w * h * nslab)
    return out


class DeviceHaloBackend:
    """pack/unpack on the GPU through the C ABI; buffers are torch CUDA tensors"""
    device = True

    def __init__(self, ts):
        import torch
        self.torch = torch
        self.ts = ts
        self.dev = torch.device("cuda", ts.dims.device)
        self._tables, self._keep = {}, []

    def counts(self, m, n):
        cnt = (C.c_int64 * 8)()
        p = self.ts.cb.params()
        self.ts._ck(self.ts.lib.hycom_tsadvc_halo_counts(self.ts.h, m, n, C.byref(p), C.byref(cnt)))
        return [int(c) for c in cnt]

    def alloc(self, n):
        return self.torch.empty(n, dtype=self.torch.float64, device=self.dev)

    def _table(self, bufs):
        key = id(bufs)
        t = self._tables.get(key)
        if t is None:
            t = (C.c_void_p * 8)()
            for d in range(8):
                t[d] = None if bufs[d] is None else bufs[d].data_ptr()
            self._tables[key] = t
            self._keep.append(bufs)
        return t

    def pack(self, m, n, send, stream=None):
        p = self.ts.cb.params()
        self.ts._ck(self.ts.lib.hycom_tsadvc_halo_pack(self.ts.h, m, n, C.byref(p), C.byref(self._table(send)),
                                                      C.c_void_p(stream) if stream else None))

    def unpack(self, m, n, recv, stream=None):
        p = self.ts.cb.params()
        self.ts._ck(self.ts.lib.hycom_tsadvc_halo_unpack(self.ts.h, m, n, C.byref(p), C.byref(self._table(recv)),
                                                        C.c_void_p(stream) if stream else None))


    # -- the exchanges inside advem_fct2c (btrmas, mod_tsadvc.F90:1186-1187): hloc and fldlo of one
    # -- layer batch, width 5
    def fct2c_counts(self, m, n, batch):
        cnt = (C.c_int64 * 8)()
        p = self.ts.cb.params()
        self.ts._ck(self.ts.lib.hycom_tsadvc_fct2c_halo_counts(self.ts.h, m, n, C.byref(p), batch, C.byref(cnt)))
        return [int(c) for c in cnt]

    def fct2c_pack(self, m, n, batch, send, stream=None):
        p = self.ts.cb.params()
        self.ts._ck(self.ts.lib.hycom_tsadvc_fct2c_halo_pack(self.ts.h, m, n, C.byref(p), batch,
                                                            C.byref(self._table(send)),
                                                            C.c_void_p(stream) if stream else None))

    def fct2c_unpack(self, m, n, batch, recv, stream=None):
        p = self.ts.cb.params()
        self.ts._ck(self.ts.lib.hycom_tsadvc_fct2c_halo_unpack(self.ts.h, m, n, C.byref(p), batch,
                                                              C.byref(self._table(recv)),
                                                              C.c_void_p(stream) if stream else None))

    # -- the second exchange of tsadvc (temdf2 > 0, mod_tsadvc.F90:2140-2151): slot n, width 2
    def diff_counts(self, n):
        cnt = (C.c_int64 * 8)()
        p = self.ts.cb.params()
        self.ts._ck(self.ts.lib.hycom_tsadvc_diff_halo_counts(self.ts.h, n, C.byref(p), C.byref(cnt)))
        return [int(c) for c in cnt]

    def diff_pack(self, n, send, stream=None):
        p = self.ts.cb.params()
        self.ts._ck(self.ts.lib.hycom_tsadvc_diff_halo_pack(self.ts.h, n, C.byref(p), C.byref(self._table(send)),
                                                           C.c_void_p(stream) if stream else None))

    def diff_unpack(self, n, recv, stream=None):
        p = self.ts.cb.params()
        self.ts._ck(self.ts.lib.hycom_tsadvc_diff_halo_unpack(self.ts.h, n, C.byref(p), C.byref(self._table(recv)),
                                                             C.c_void_p(stream) if stream else None))


class XcExchange:
    """``xctilr`` of the arrays tsadvc(m,n) exchanges (mod_tsadvc.F90:1829-1836) for one tile
    per rank of a ``torch.distributed`` group, plus the overlapped ``tsadvc_device``."""

    def __init__(self, ts, dist, backend=None, group=None, rank: Optional[int] = None,
                 compute_stream=None):
        self.ts, self.dist, self.group = ts, dist, group
        self.backend = backend if backend is not None else DeviceHaloBackend(ts)
        g = ts.cb.geom if ts is not None else backend.geom
        self.geom = g
        self.rank = dist.get_rank(group) if rank is None else rank
        me = g.mproc - 1 + g.ipr * (g.nproc - 1)
        if me != self.rank:
            raise ValueError(f"rank {self.rank} holds tile {me}: tiles are placed row-major, rank = mproc-1 + ipr*(nproc-1)")
        self.nbr = neighbors(g)
        if ts is not None and hasattr(ts, "lib"):
            nb = (C.c_int32 * 8)()
            ts._ck(ts.lib.hycom_tsadvc_halo_neighbors(ts.h, C.byref(nb)))
            assert list(nb) == self.nbr, (list(nb), self.nbr)
        self._bufs = {}
        self.frame_concurrent = True
        self.comm_stream = None
        import torch
        self.torch = torch
        if getattr(self.backend, "device", False):
            self.comm_stream = torch.cuda.Stream(device=self.backend.dev)
            self.compute_stream = compute_stream

    # -- buffers ----------------------------------------------------------------------
    def _buffers(self, m, n):
        """(send, recv, ops) of an exchange with these leapfrog slots; built once and reused
        (the per-step host work is what limits the 8-GPU step, not the bytes)"""
        # the message layout depends on every scalar that selects the exchanged arrays and the halo
        # width (advtyp: mbdy 2 or 5; mxlmy: q2, q2l; advflg/isopyc/nhybrd: th3d): key on the counts
        cnt = self.backend.counts(m, n)
        key = (m, n, tuple(cnt))
        if key not in self._bufs:
            send = [self.backend.alloc(c) if self.nbr[d] >= 0 else None for d, c in enumerate(cnt)]
            recv: List = [None] * 8
            for d, c in enumerate(cnt):
                if self.nbr[d] < 0:
                    continue
                # a periodic edge that wraps onto this tile: what leaves in the opposite
                # direction is what arrives here
                recv[d] = send[opp_dir(self.geom, d)] if self.nbr[d] == self.rank else self.backend.alloc(c)
            self._bufs[key] = (send, recv, self.ops(send, recv))
        return self._bufs[key]

    def ops(self, send, recv):
        """the P2P operations of one exchange, ordered so that the k-th send of rank A to
        rank B meets the k-th receive B posts for A (NCCL matches by order, not by tag)"""
        P2POp, dist = self.dist.P2POp, self.dist
        ops = []
        for d in range(8):
            peer = self.nbr[d]
            if peer >= 0 and peer != self.rank:
                ops.append(P2POp(dist.isend, send[d], self._global(peer), group=self.group, tag=d))
        # receives in the order of the SENDER's direction d; a top-row arctic tile may hear from the
        # same d twice (d=3: from the tile below it and, folded, from its twin)
        for d, e in sorted((opp_dir(self.geom, e), e) for e in range(8)):
            peer = self.nbr[e]
            if peer >= 0 and peer != self.rank:
                ops.append(P2POp(dist.irecv, recv[e], self._global(peer), group=self.group, tag=d))
        return ops

    def _global(self, peer):
        return peer if self.group is None else self.dist.get_global_rank(self.group, peer)

    # -- the exchange -------------------------------------------------------------------
    def start(self, m, n):
        send, recv, ops = self._buffers(m, n)
        cs = self.comm_stream
        if cs is not None:
            cs.wait_stream(self._compute())
            with self.torch.cuda.stream(cs):
                self.backend.pack(m, n, send, cs.cuda_stream)
                works = self.dist.batch_isend_irecv(ops) if ops else []
        else:
            self.backend.pack(m, n, send)
            works = self.dist.batch_isend_irecv(ops) if ops else []
        return works, recv

    def finish(self, m, n, pending):
        works, recv = pending
        cs = self.comm_stream
        if cs is not None:
            with self.torch.cuda.stream(cs):
                for w in works:
                    w.wait()
                self.backend.unpack(m, n, recv, cs.cuda_stream)
            self._compute().wait_stream(cs)
        else:
            for w in works:
                w.wait()
            self.backend.unpack(m, n, recv)

    def xctilr(self, m, n):
        """blocking form: halos of every exchanged array valid to width 5 on return"""
        self.finish(m, n, self.start(m, n))

    def _compute(self):
        if self.compute_stream is None:
            raise RuntimeError("XcExchange.compute_stream is not set (the torch stream the handle runs on)")
        return self.compute_stream

    # -- tsadvc(m,n) with the exchange overlapped with the interior ----------------------
    def tsadvc_device(self, m, n, diag: bool = True, overlap: bool = True):
        ts = self.ts
        p = ts.cb.params()
        xm = ts.xmin.ctypes.data_as(C.c_void_p) if diag else None
        xx = ts.xmax.ctypes.data_as(C.c_void_p) if diag else None
        pending = self.start(m, n)
        if ts.cb.btrmas and abs(ts.cb.advtyp) == 2:
            # advem_fct2c exchanges hloc/fldlo after each of its five iterations: no interior overlap
            self.finish(m, n, pending)
            self._fct2c(m, n, p)
            ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_ALL, xm, xx))
        elif overlap and not ts.cb.isopyc:
            # (isopyc: the flux smoothing and the march of layer 1 read the halo from their first kernel on
            # and run on the handle's stream - exchange first, like advem_fct2c)
            ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_INTERIOR, None, None))
            if self.comm_stream is not None and self.frame_concurrent:
                # the frame runs on the comm stream right behind the unpack, next to the interior
                # launch (it fills that launch's tail); the handle's stream joins both afterwards
                works, recv = pending
                with self.torch.cuda.stream(self.comm_stream):
                    for w in works:
                        w.wait()
                    self.backend.unpack(m, n, recv, self.comm_stream.cuda_stream)
                ts._ck(ts.lib.hycom_tsadvc_set_frame_stream(ts.h, C.c_void_p(self.comm_stream.cuda_stream)))
                ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_FRAME, xm, xx))
                ts._ck(ts.lib.hycom_tsadvc_set_frame_stream(ts.h, None))
            else:
                self.finish(m, n, pending)
                ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_FRAME, xm, xx))
        else:
            self.finish(m, n, pending)
            ts._ck(ts.lib.hycom_tsadvc_step_device_part(ts.h, m, n, C.byref(p), cabi.PART_ALL, xm, xx))
        if diag and (ts.cb.nstep % 3 == 0 or ts.cb.diagno):
            self.xcminmax(ts.xmin, ts.xmax)
        if ts.cb.temdf2 > 0.0:   # mod_tsadvc.F90:2138-2230: second exchange (width 2), then tsdff + EOS
            self.xctilr_diff(n)
            ts._ck(ts.lib.hycom_tsadvc_diffuse_device(ts.h, m, n, C.byref(p)))

    def _exchange(self, key, counts, pack, unpack):
        """one blocking exchange of a set of device arrays (buffers and P2P ops cached per key)"""
        if key not in self._bufs:
            cnt = counts()
            send = [self.backend.alloc(c) if self.nbr[d] >= 0 else None for d, c in enumerate(cnt)]
            recv: List = [None] * 8
            for d, c in enumerate(cnt):
                if self.nbr[d] >= 0:
                    recv[d] = send[opp_dir(self.geom, d)] if self.nbr[d] == self.rank else self.backend.alloc(c)
            self._bufs[key] = (send, recv, self.ops(send, recv))
        send, recv, ops = self._bufs[key]
        cs = self.comm_stream
        if cs is not None:
            cs.wait_stream(self._compute())
            with self.torch.cuda.stream(cs):
                pack(send, cs.cuda_stream)
                for w in (self.dist.batch_isend_irecv(ops) if ops else []):
                    w.wait()
                unpack(recv, cs.cuda_stream)
            self._compute().wait_stream(cs)
        else:
            pack(send, None)
            for w in (self.dist.batch_isend_irecv(ops) if ops else []):
                w.wait()
            unpack(recv, None)

    def _fct2c(self, m, n, p):
        """advem_fct2c on several tiles: per layer batch, set-up, five iterations each followed by
        xctilr(hloc), xctilr(fldlo) (mod_tsadvc.F90:1088-1187), then the limiter and the update"""
        ts, be = self.ts, self.backend
        nbatch, per = C.c_int32(0), C.c_int32(0)
        ts._ck(ts.lib.hycom_tsadvc_fct2c_batches(ts.h, C.byref(nbatch), C.byref(per)))
        for b in range(nbatch.value):
            ts._ck(ts.lib.hycom_tsadvc_fct2c_stage(ts.h, m, n, C.byref(p), b, 0))
            for _ in range(5):
                ts._ck(ts.lib.hycom_tsadvc_fct2c_stage(ts.h, m, n, C.byref(p), b, 1))
                self._exchange(("fct2c", m, n, b, ts.cb.ntracr),
                               lambda: be.fct2c_counts(m, n, b),
                               lambda send, st: be.fct2c_pack(m, n, b, send, st),
                               lambda recv, st: be.fct2c_unpack(m, n, b, recv, st))
            ts._ck(ts.lib.hycom_tsadvc_fct2c_stage(ts.h, m, n, C.byref(p), b, 2))

    def xctilr_diff(self, n):
        """xctilr(saln|temp|th3d|tracer(:,:,:,n), 1,kk, 2,2, halo_ps) of mod_tsadvc.F90:2140-2151"""
        key = ("diff", n, self.ts.cb.ntracr)
        if key not in self._bufs:
            cnt = self.backend.diff_counts(n)
            send = [self.backend.alloc(c) if self.nbr[d] >= 0 else None for d, c in enumerate(cnt)]
            recv: List = [None] * 8
            for d, c in enumerate(cnt):
                if self.nbr[d] >= 0:
                    recv[d] = send[opp_dir(self.geom, d)] if self.nbr[d] == self.rank else self.backend.alloc(c)
            self._bufs[key] = (send, recv, self.ops(send, recv))
        send, recv, ops = self._bufs[key]
        cs = self.comm_stream
        if cs is not None:
            cs.wait_stream(self._compute())
            with self.torch.cuda.stream(cs):
                self.backend.diff_pack(n, send, cs.cuda_stream)
                for w in (self.dist.batch_isend_irecv(ops) if ops else []):
                    w.wait()
                self.backend.diff_unpack(n, recv, cs.cuda_stream)
            self._compute().wait_stream(cs)
        else:
            self.backend.diff_pack(n, send)
            for w in (self.dist.batch_isend_irecv(ops) if ops else []):
                w.wait()
            self.backend.diff_unpack(n, recv)

    def xcminmax(self, xmin, xmax):
        """xcminr / xcmaxr of the per-layer salinity range (mod_tsadvc.F90:2093-2094):
        element-wise min/max over all tiles, in place"""
        torch = self.torch
        import numpy as np
        nccl = self.dist.get_backend(self.group) == "nccl"
        dev = getattr(self.backend, "dev", "cuda") if nccl else "cpu"
        kk = xmin.shape[0]
        both = torch.from_numpy(np.concatenate([xmin, -xmax])).to(dev)   # max = -min(-x): one collective
        self.dist.all_reduce(both, op=self.dist.ReduceOp.MIN, group=self.group)
        both = both.cpu().numpy()
        xmin[:] = both[:kk]
        xmax[:] = -both[kk:]
