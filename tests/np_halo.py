"""numpy stand-in for the device pack/unpack kernels (TEST INFRASTRUCTURE): lets the
product's exchange schedule (hycom-src_b200/xc.py: neighbours, message order, periodic
self-wrap, closed edges) run on CPU ranks under gloo.  Message layout = halo.cu's:
[array][k][row][col]."""
import numpy as np
import torch

from util import pkg

xc = __import__("importlib").import_module("hycom-src_b200.xc")


def region(g, d, recv, mh=5, nh=5):
    dx, dy = xc.DIR_DXY[d]
    nb = g.nbdy
    if dx == 0:
        cs = slice(nb, nb + g.ii)
    elif dx < 0:
        cs = slice(nb - mh, nb) if recv else slice(nb, nb + mh)
    else:
        cs = slice(nb + g.ii, nb + g.ii + mh) if recv else slice(nb + g.ii - mh, nb + g.ii)
    if dy == 0:
        rs = slice(nb, nb + g.jj)
    elif dy < 0:
        rs = slice(nb - nh, nb) if recv else slice(nb, nb + nh)
    else:
        rs = slice(nb + g.jj, nb + g.jj + nh) if recv else slice(nb + g.jj - nh, nb + g.jj)
    return rs, cs


def fold_columns(g, d, mh):
    """Fortran columns a top-row arctic tile sends in a fold direction (csrc/tsadvc_launch.h)"""
    if d == 3:
        return np.arange(1, g.ii + 1)
    if d == 6:
        return np.arange(1, mh + 2)
    return np.arange(g.ii - mh + 1, g.ii + 1)


def fold_pack(g, a, d, itype, mh, nh):
    """[k][j][c] block of array a (kk, nrows, ncols) for fold direction d: rows jj-1-j (p,u grid)
    or jj-j (q,v grid), sign flipped for vector fields unless vland (mod_xc_mp.h:4263-4372)"""
    nb, grid = g.nbdy, itype % 10
    cols = fold_columns(g, d, mh) - 1 + nb
    j = np.arange(1, nh + 1)
    rows = (g.jj - 1 - j if grid in (1, 3) else g.jj - j) - 1 + nb
    blk = a[:, rows][:, :, cols]
    if itype > 10:
        blk = np.where(blk != 0.0, -blk, blk)
    return blk


def fold_unpack(g, a, d, itype, mh, nh, blk):
    """mirror: column c of the twin lands at i = ii+1+s-c in row jj+j (s=1 on the u,q grids)"""
    nb, grid = g.nbdy, itype % 10
    sh = 1 if grid in (2, 3) else 0
    col = fold_columns(g, d, mh)
    ct = col if d == 3 else col + g.ii if d == 6 else col - g.ii
    i = g.ii + 1 + sh - ct
    ok = (i >= 1 - mh) & (i <= g.ii + mh)
    rows = g.jj + np.arange(1, nh + 1) - 1 + nb
    a[:, rows[:, None], (i[ok] - 1 + nb)[None, :]] = blk[:, :, ok]


class NumpyHaloBackend:
    device = False

    def __init__(self, geom, arrays, itypes=None):
        self.geom, self.arrays = geom, arrays     # arrays: list of (kk, nrows, ncols)
        self.itypes = itypes if itypes is not None else [1] * len(arrays)

    def counts(self, m, n):
        return xc.halo_counts(self.geom, sum(a.shape[0] for a in self.arrays))

    def alloc(self, n):
        return torch.empty(n, dtype=torch.float64)

    def _is_fold(self, d):
        return xc.arctic_fold(self.geom) and xc.DIR_DXY[d][1] > 0

    def _pack(self, arrays, itypes, send, mh, nh):
        for d in range(8):
            if send[d] is None:
                continue
            if self._is_fold(d):
                send[d].copy_(torch.from_numpy(np.concatenate(
                    [fold_pack(self.geom, a, d, it, mh, nh).ravel() for a, it in zip(arrays, itypes)])))
                continue
            rs, cs = region(self.geom, d, False, mh, nh)
            send[d].copy_(torch.from_numpy(np.concatenate([a[:, rs, cs].ravel() for a in arrays])))

    def _unpack(self, arrays, itypes, recv, mh, nh):
        for d in range(8):
            if self._is_fold(d):
                buf, off = recv[d].numpy(), 0
                w = len(fold_columns(self.geom, d, mh))
                for a, it in zip(arrays, itypes):
                    n_el = a.shape[0] * nh * w
                    fold_unpack(self.geom, a, d, it, mh, nh, buf[off:off + n_el].reshape(a.shape[0], nh, w))
                    off += n_el
                continue
            rs, cs = region(self.geom, d, True, mh, nh)
            if recv[d] is None:
                for a in arrays:
                    a[:, rs, cs] = 0.0          # vland
                continue
            buf, off = recv[d].numpy(), 0
            for a in arrays:
                blk = a[:, rs, cs]
                blk[...] = buf[off:off + blk.size].reshape(blk.shape)
                off += blk.size

    def pack(self, m, n, send, stream=None):
        self._pack(self.arrays, self.itypes, send, 5, 5)

    def unpack(self, m, n, recv, stream=None):
        self._unpack(self.arrays, self.itypes, recv, 5, 5)


class NumpyStagedBackend(NumpyHaloBackend):
    """the second (diffusion, width 2) and the in-scheme (fct2c, width 5, per layer batch) exchanges of
    XcExchange with numpy arrays in place of the device mirrors"""

    def __init__(self, geom, arrays, diff_arrays, batch_arrays):
        super().__init__(geom, arrays)
        self.diff_arrays, self.batch_arrays = diff_arrays, batch_arrays

    def _xfer(self, arrays, bufs, recv, mh, nh):
        if recv:
            self._unpack(arrays, [1] * len(arrays), bufs, mh, nh)
        else:
            self._pack(arrays, [1] * len(arrays), bufs, mh, nh)

    def diff_counts(self, n):
        return xc.halo_counts(self.geom, sum(a.shape[0] for a in self.diff_arrays), 2, 2)

    def diff_pack(self, n, send, stream=None):
        self._xfer(self.diff_arrays, send, False, 2, 2)

    def diff_unpack(self, n, recv, stream=None):
        self._xfer(self.diff_arrays, recv, True, 2, 2)

    def fct2c_counts(self, m, n, batch):
        return xc.halo_counts(self.geom, sum(a.shape[0] for a in self.batch_arrays[batch]), 5, 5)

    def fct2c_pack(self, m, n, batch, send, stream=None):
        self._xfer(self.batch_arrays[batch], send, False, 5, 5)

    def fct2c_unpack(self, m, n, batch, recv, stream=None):
        self._xfer(self.batch_arrays[batch], recv, True, 5, 5)
