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
