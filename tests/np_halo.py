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


class NumpyHaloBackend:
    device = False

    def __init__(self, geom, arrays):
        self.geom, self.arrays = geom, arrays     # arrays: list of (kk, nrows, ncols)

    def counts(self, m, n):
        return xc.halo_counts(self.geom, sum(a.shape[0] for a in self.arrays))

    def alloc(self, n):
        return torch.empty(n, dtype=torch.float64)

    def pack(self, m, n, send, stream=None):
        for d in range(8):
            if send[d] is None:
                continue
            rs, cs = region(self.geom, d, False)
            send[d].copy_(torch.from_numpy(np.concatenate([a[:, rs, cs].ravel() for a in self.arrays])))

    def unpack(self, m, n, recv, stream=None):
        for d in range(8):
            rs, cs = region(self.geom, d, True)
            if recv[d] is None:
                for a in self.arrays:
                    a[:, rs, cs] = 0.0          # vland
                continue
            buf, off = recv[d].numpy(), 0
            for a in self.arrays:
                blk = a[:, rs, cs]
                blk[...] = buf[off:off + blk.size].reshape(blk.shape)
                off += blk.size


class NumpyStagedBackend(NumpyHaloBackend):
    """the second (diffusion, width 2) and the in-scheme (fct2c, width 5, per layer batch) exchanges of
    XcExchange with numpy arrays in place of the device mirrors"""

    def __init__(self, geom, arrays, diff_arrays, batch_arrays):
        super().__init__(geom, arrays)
        self.diff_arrays, self.batch_arrays = diff_arrays, batch_arrays

    def _xfer(self, arrays, bufs, recv, mh, nh):
        for d in range(8):
            rs, cs = region(self.geom, d, recv, mh, nh)
            if not recv:
                if bufs[d] is not None:
                    bufs[d].copy_(torch.from_numpy(np.concatenate([a[:, rs, cs].ravel() for a in arrays])))
                continue
            if bufs[d] is None:
                for a in arrays:
                    a[:, rs, cs] = 0.0
                continue
            buf, off = bufs[d].numpy(), 0
            for a in arrays:
                blk = a[:, rs, cs]
                blk[...] = buf[off:off + blk.size].reshape(blk.shape)
                off += blk.size

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
