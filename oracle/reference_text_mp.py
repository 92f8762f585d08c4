"""oracle/reference_text_mp.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's MULTI-TILE halo exchange, executed: xctilr of mod_xc_mp.h (:4664-4987, and the ARCTIC variant
:4114-4662) runs as written on ipr x jpr tiles inside one process - one thread per tile, each with its own copy of
the module variables - in either of the reference's two flavours:
  * `-DMPI`  (what HYCOM is built with): persistent requests; supplied from outside the text are mpi_send_init /
    mpi_recv_init / mpi_startall / mpi_waitall / mpi_request_free over in-process FIFO mailboxes, nothing else;
  * `-DSHMEM`: the one-sided variant of the same routine; supplied are shmem_barrier_all (a thread barrier) and
    shmem_get64 (copy from the same symmetric buffer of the target tile).  (Its ARCTIC branch does not compile as
    written - `mod(ipr+1-mproc)` with one argument at :4506, against `mod(ipr+1-mproc,ipr)` in the MPI branch :4425 -
    so the arctic exchange is executed in the MPI flavour only.)
The tables the routine reads - idproc, idhalo, the top / bottom neighbour lists m0_top, mm_top, i0_st, ii_st,
i0_gt ... - are built by the reference's own xcspmd text (mod_xc_mp.h, from `null_tile = ...` through the tile
printout), run per tile on the tile extents that patch.input would hold (i0_pe, ii_pe, j0_pe, jj_pe).

Used by tests/test_reference_text.py to pin the oracle's multi-tile xctilr (orc_world_xctilr), against which the
product's exchange schedule and its device pack / unpack are tested."""
from __future__ import annotations

import os
import queue
import threading

import numpy as np

import fortran_exec as fx
from reference_text import REF, NBDY

_PATH = os.path.join(REF, "mod_xc_mp.h")


def _tile_env(tiles, t, ipr, jpr, nreg, itdm, jtdm, kdm, mnproc):
    nb = NBDY
    I = lambda bounds, fill=0: fx.FArray.zeros(bounds, dtype=np.int64, fill=fill)   # noqa: E731
    # (iqr, jqr, ijqr: the compile-time maxima of mod_dimensions.F90 that size the tables, larger than any tiling)
    iqr, jqr = ipr + 2, jpr + 2
    env = dict(ipr=ipr, jpr=jpr, ijpr=ipr * jpr, iqr=iqr, jqr=jqr, ijqr=iqr * jqr, nreg=nreg, itdm=itdm, jtdm=jtdm,
               nbdy=nb, kdm=kdm, idm=t.idm, jdm=t.jdm, mnproc=mnproc, npesi=ipr * jpr, lp=6, flush_lp=1, vland=0.0,
               mproc=-1, nproc=-1, null_tile=-99, mp_1st=-1, ixsum=-1, i0=-1, ii=-1, j0=-1, jj=-1,
               m0_top=-1, mm_top=-1, m0_bot=-1, mm_bot=-1)
    for n in ("i0_pe", "ii_pe", "j0_pe", "jj_pe", "i1sum", "iisum"):
        env[n] = I(((1, iqr), (1, jqr)))
    for g in tiles:     # what patch.input holds (xcspmd reads ispt = i0+1, iipe, jspt, jjpe)
        env["i0_pe"][g.mproc, g.nproc], env["ii_pe"][g.mproc, g.nproc] = g.i0, g.ii
        env["j0_pe"][g.mproc, g.nproc], env["jj_pe"][g.mproc, g.nproc] = g.j0, g.jj
    env["idproc"] = I(((0, iqr + 1), (0, jqr + 1)), -77)
    env["idproc1"] = I(((0, iqr * jqr + 1),), -77)
    env["idhalo"] = I(((1, 2),), -77)
    env["mpe_1"], env["mpe_e"] = I(((1, jqr),)), I(((1, jqr),))
    env["mpe_i"], env["npe_j"] = I(((0, itdm + 1), (0, jqr))), I(((0, jtdm + 1),))
    for n in ("i0_st", "ii_st", "i0_gt", "ii_gt", "i0_sb", "ii_sb", "i0_gb", "ii_gb"):
        env[n] = I(((1, iqr),), -77)
    # the symmetric buffers of xctilr (save, allocatable: allocated on the first call)
    env["ai"] = fx.FArray.zeros(((1, t.idm * kdm * nb + 64), (1, 4)), fill=np.nan)
    env["aj"] = fx.FArray.zeros(((1, (t.jdm + 2 * nb) * kdm * nb + 64), (1, 4)), fill=np.nan)
    env["aia"] = fx.FArray.zeros(((1, kdm * nb + 64), (1, 2)), fill=np.nan)
    # MPI flavour: the saved request tables with the values their `data` statements give them, constants of mpif.h
    env["mpireqa"], env["mpireqb"] = I(((1, 4 * iqr),), -1), I(((1, 4),), -1)
    env.update(nreqa=0, klmold=0, klnold=0, mhlold=0, nhlold=0, ityold=0, ilold=0, jlold=0, mpierr=0,
               mpi_proc_null=-2, mpi_comm_hycom=0, mpi_real8=8, mpi_statuses_ignore=0, group_1st_in_row=0)
    return env


class World:
    """ipr x jpr tiles of one global grid; `tiles` are the product's Geometry objects (hycom-src_b200/geometry.py)"""

    def __init__(self, tiles, ipr, jpr, nreg, itdm, jtdm, kdm, flavour="MPI", envs=None):
        """envs: environments to build on (reference_text.make_env of every tile, for running bigrid / tsadvc on the
        tiles: uniform tilings only); their xctilr, xcmaxr, xcminr become the multi-tile ones"""
        self.tiles, self.n = tiles, len(tiles)
        self.slots = [None] * self.n
        arctic = nreg == 2
        defines = ("RELO", flavour) + (("ARCTIC",) if arctic else ())
        self.envs = []
        self.barrier = threading.Barrier(self.n)
        self.mail = {}                       # (to, from, tag) -> FIFO of packed strips
        self.lock = threading.Lock()
        start = r"^null_tile\s*=\s*-1$" if flavour == "SHMEM" else r"^null_tile\s*=\s*mpi_proc_null$"
        for r, t in enumerate(tiles):
            assert r == (t.mproc - 1) + ipr * (t.nproc - 1)
            env = envs[r] if envs is not None else {}
            if envs is not None:
                assert (t.idm, t.jdm) == (t.ii, t.jj), "uniform tiles only"
                env["xcmaxr"] = (lambda rr: lambda x: self.allreduce(rr, x, max))(r)
                env["xcminr"] = (lambda rr: lambda x: self.allreduce(rr, x, min))(r)
            env.update(_tile_env(tiles, t, ipr, jpr, nreg, itdm, jtdm, kdm, r + 1))
            fx.compile_slice(_PATH, "xcspmd", start, r"^call xcsync\(flush_lp\)$", "xcspmd_tables", env, defines=defines,
                             skip_calls=("xcsync", "xcstop", "mpi_comm_split", "mpi_comm_free"))
            env["xcspmd_tables"]()
            assert (env["mproc"], env["nproc"], env["i0"], env["ii"], env["j0"], env["jj"]) == \
                   (t.mproc, t.nproc, t.i0, t.ii, t.j0, t.jj)
            env["shmem_barrier_all"] = self.barrier.wait
            env["shmem_get64"] = self._getter(env)
            self._mpi(env, r)
            ranks = {"shmem_get64": (1, 1, None, None), "mpi_send_init": (1, None, None, None, None, None, 1, None),
                     "mpi_recv_init": (1, None, None, None, None, None, 1, None)}
            fx.compile_unit(_PATH, "xctilr", env, defines=defines, skip_calls=("xctmr0", "xctmr1", "mem_stat_add"),
                            callee_ranks=ranks, drop_blocks=(r"allocated",))
            self.envs.append(env)

    def allreduce(self, r, x, op):
        """xcmaxr / xcminr of a scalar over the tiles"""
        self.slots[r] = x
        self.barrier.wait()
        v = op(self.slots)
        self.barrier.wait()
        return v

    def run(self, fn):
        """fn(rank, env) on every tile, one thread each"""
        errs = []

        def go(r):
            try:
                fn(r, self.envs[r])
            except BaseException as e:      # noqa: BLE001
                errs.append((r, e))
                self.barrier.abort()
        th = [threading.Thread(target=go, args=(r,)) for r in range(self.n)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0][1]

    def _box(self, to, frm, tag):
        with self.lock:
            return self.mail.setdefault((to, frm, tag), queue.Queue())

    def _mpi(self, env, me):
        """persistent point-to-point requests (mpi_send_init ... mpi_waitall) between the threads of this world"""
        reqs = []

        def init(kind):
            def f(buf, count, dtype, peer, tag, comm, req, ierr):
                reqs.append((kind, buf, count, peer, tag))
                req.a[0] = len(reqs) - 1
            return f

        def startall(n, handles, ierr):
            for h in handles.a[:n]:
                kind, buf, count, peer, tag = reqs[h]
                if kind == "send" and peer != env["mpi_proc_null"]:
                    self._box(peer, me, tag).put(buf.a[:count].copy())

        def waitall(n, handles, statuses, ierr):
            for h in handles.a[:n]:
                kind, buf, count, peer, tag = reqs[h]
                if kind == "recv" and peer != env["mpi_proc_null"]:
                    data = self._box(me, peer, tag).get(timeout=60)
                    assert len(data) == count, "message length"
                    buf.a[:count] = data
        env.update(mpi_send_init=init("send"), mpi_recv_init=init("recv"), mpi_startall=startall, mpi_waitall=waitall,
                   mpi_request_free=lambda req, ierr: None)

    def _getter(self, me):
        def shmem_get64(dest, src, n, pe):
            """dest(1:n) = src(1:n) of tile `pe`: src is a view of MY symmetric buffer; the same place in the target's"""
            for name in ("ai", "aj", "aia"):
                base = me[name].a
                off = src.a.__array_interface__["data"][0] - base.__array_interface__["data"][0]
                if 0 <= off < base.nbytes:
                    # (FArray storage is [column][row]: a(l,c) is store[c-1, l-1], contiguous along l)
                    remote = self.envs[pe][name].a.reshape(-1)
                    dest.a[:n] = remote[off // 8: off // 8 + n]
                    return
            raise ValueError("shmem_get64: the source is not a symmetric buffer")
        return shmem_get64

    def xctilr(self, arrays, l1, ld, mh, nh, itype):
        """arrays: one (nslab, nrows, ncols) numpy array per tile, updated in place"""
        errs = []

        def run(r):
            try:
                a = fx.FArray(arrays[r], (1 - NBDY, 1 - NBDY, 1))
                self.envs[r]["xctilr"](a, l1, ld, mh, nh, itype)
            except BaseException as e:      # noqa: BLE001
                errs.append((r, e))
                self.barrier.abort()
        th = [threading.Thread(target=run, args=(r,)) for r in range(self.n)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0][1]
