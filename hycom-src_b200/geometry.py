"""Tile geometry, land/sea masks and grid metrics of the tsadvc path (host side).

Mirrors what the path needs from mod_xc (xcspmd: tile placement,
mod_xc_mp.h:2317-3288), bigrid.F90 (ip/iu/iv) and geopar.F90:311-340 (scp2, scp2i,
aspux, aspvy).  Arrays are numpy, shape (jdm+2*nbdy, idm+2*nbdy) == the Fortran
array a(1-nbdy:idm+nbdy, 1-nbdy:jdm+nbdy) with i contiguous.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass(frozen=True)
class TileGeom:
    idm: int
    jdm: int
    kdm: int
    nbdy: int
    ii: int
    jj: int
    i0: int
    j0: int
    itdm: int
    jtdm: int
    nreg: int
    ipr: int = 1
    jpr: int = 1
    mproc: int = 1   # 1-based tile column
    nproc: int = 1   # 1-based tile row

    @property
    def ncols(self) -> int:
        return self.idm + 2 * self.nbdy

    @property
    def nrows(self) -> int:
        return self.jdm + 2 * self.nbdy

    @property
    def periodic_i(self) -> bool:
        return self.nreg not in (0, 4)

    @property
    def periodic_j(self) -> bool:
        return self.nreg > 2

    def interior(self):
        """numpy slices of 1:ii, 1:jj"""
        nb = self.nbdy
        return slice(nb, nb + self.jj), slice(nb, nb + self.ii)


def partition(itdm: int, jtdm: int, kdm: int, ipr: int, jpr: int, nreg: int,
              nbdy: int = 6) -> List[TileGeom]:
    """Uniform ipr x jpr tiling (the equal-size case of patch.input,
    mod_xc_mp.h:2485-2517).  Tiles are returned row-major: index = m + ipr*n."""
    def splits(n, p):
        base, rem = divmod(n, p)
        sizes = [base + (1 if q < rem else 0) for q in range(p)]
        offs = [sum(sizes[:q]) for q in range(p)]
        return sizes, offs
    isz, ioff = splits(itdm, ipr)
    jsz, joff = splits(jtdm, jpr)
    idm, jdm = max(isz), max(jsz)   # RELO: idm/jdm = largest tile (mod_xc_mp.h:2760)
    tiles = []
    for n in range(jpr):
        for m in range(ipr):
            tiles.append(TileGeom(idm=idm, jdm=jdm, kdm=kdm, nbdy=nbdy, ii=isz[m], jj=jsz[n],
                                  i0=ioff[m], j0=joff[n], itdm=itdm, jtdm=jtdm, nreg=nreg,
                                  ipr=ipr, jpr=jpr, mproc=m + 1, nproc=n + 1))
    return tiles


def _global_index(g: TileGeom):
    """global (ig, jg) of every local cell incl. halo, with periodic wrap; -1 = outside"""
    nb = g.nbdy
    ig = g.i0 + np.arange(1 - nb, g.idm + nb + 1)
    jg = g.j0 + np.arange(1 - nb, g.jdm + nb + 1)
    if g.periodic_i:
        ig = (ig - 1) % g.itdm + 1
    else:
        ig = np.where((ig >= 1) & (ig <= g.itdm), ig, -1)
    if g.periodic_j:
        jg = (jg - 1) % g.jtdm + 1
    else:
        jg = np.where((jg >= 1) & (jg <= g.jtdm), jg, -1)
    return ig, jg


def bigrid_masks(sea: np.ndarray, g: TileGeom):
    """ip, iu, iv of one tile from the global sea map (shape (jtdm, itdm), 1 = sea).

    bigrid.F90:193-297: ip = depth>0 over the whole padded tile (halo from the
    neighbour / periodic image, land beyond closed edges), iu(i,j) = ip(i-1,j) and
    ip(i,j), iv(i,j) = ip(i,j-1) and ip(i,j); masks are only defined for
    1-nbdy..ii+nbdy (zero in the unused part of a ragged tile).
    """
    nb = g.nbdy
    if g.nreg == 2:
        return _bigrid_masks_arctic(sea, g)
    ig, jg = _global_index(g)
    nr, nc = g.nrows, g.ncols
    # one extra cell to the west/south for iu/iv at the first halo line
    def lookup(iga, jga):
        ok = (jga[:, None] >= 1) & (iga[None, :] >= 1)
        v = sea[np.clip(jga, 1, g.jtdm)[:, None] - 1, np.clip(iga, 1, g.itdm)[None, :] - 1]
        return np.where(ok, v, 0).astype(np.int32)
    ip = lookup(ig, jg)
    igw = g.i0 + np.arange(1 - nb, g.idm + nb + 1) - 1
    jgs = g.j0 + np.arange(1 - nb, g.jdm + nb + 1) - 1
    if g.periodic_i:
        igw = (igw - 1) % g.itdm + 1
    else:
        igw = np.where((igw >= 1) & (igw <= g.itdm), igw, -1)
    if g.periodic_j:
        jgs = (jgs - 1) % g.jtdm + 1
    else:
        jgs = np.where((jgs >= 1) & (jgs <= g.jtdm), jgs, -1)
    ipw = lookup(igw, jg)
    ips = lookup(ig, jgs)
    iu = (ip & ipw).astype(np.int32)
    iv = (ip & ips).astype(np.int32)
    # bigrid only fills 1-nbdy..ii+nbdy x 1-nbdy..jj+nbdy (loops :208-209, :246-247)
    live = np.zeros((nr, nc), dtype=bool)
    live[: g.jj + 2 * nb, : g.ii + 2 * nb] = True
    ip = np.where(live, ip, 0).astype(np.int32)
    iu = np.where(live, iu, 0).astype(np.int32)
    iv = np.where(live, iv, 0).astype(np.int32)
    return np.ascontiguousarray(ip), np.ascontiguousarray(iu), np.ascontiguousarray(iv)


def _bigrid_masks_arctic(sea: np.ndarray, g: TileGeom):
    """nreg=2 (global grid across the arctic): periodic in i, land south of row 1, and the rows
    above jtdm are the tripole fold ip(i,jtdm+j) = ip(itdm+1-i, jtdm-1-j) (bigrid.F90:116-131 +
    xctilr halo_ps of mod_xc_sm.h:1215-1232 / mod_xc_mp.h:4263-4281).  iu/iv follow from ip: for a
    depth array that is itself fold-consistent in row jtdm this equals the reference's
    halo_us/halo_vs updates of the interior iu/iv.  A tile is a window of the padded global map
    (the tiles of the top row see the fold in their northern halo)."""
    nb, ni, nj = g.nbdy, g.itdm, g.jtdm
    ipg = np.zeros((nj + 2 * nb + 1, ni + 2 * nb + 1), dtype=np.int32)   # one extra cell west/south
    jj = np.arange(-nb, nj + nb + 1)       # Fortran j of every row (from 1-nb-1)
    ii = np.arange(-nb, ni + nb + 1)
    for r, j in enumerate(jj):
        for_i = (ii - 1) % ni + 1           # periodic in i
        if j < 1:
            continue                        # south boundary is all land
        if j <= nj:
            ipg[r] = sea[j - 1, for_i - 1]
        else:
            jo = nj - 1 - (j - nj)
            io = ni + 1 - for_i
            ipg[r] = sea[jo - 1, io - 1]
    ip = ipg[1:, 1:]
    iu = ip & ipg[1:, :-1]
    iv = ip & ipg[:-1, 1:]
    if g.ipr * g.jpr > 1:
        win = (slice(g.j0, g.j0 + g.nrows), slice(g.i0, g.i0 + g.ncols))
        live = np.zeros((g.nrows, g.ncols), dtype=bool)     # bigrid fills 1-nbdy..ii+nbdy only
        live[: g.jj + 2 * nb, : g.ii + 2 * nb] = True
        ip, iu, iv = (np.where(live, a[win], 0) for a in (ip, iu, iv))
    return (np.ascontiguousarray(ip, dtype=np.int32), np.ascontiguousarray(iu, dtype=np.int32),
            np.ascontiguousarray(iv, dtype=np.int32))


def geopar_metrics(scpx, scpy, scux, scuy, scvx, scvy, aspmax: float = 2.0):
    """geopar.F90:311-340: cell areas, their inverses and the diffusion aspect factors."""
    epsil = 1.0e-11  # mod_cb_arrays.F90:852
    scp2 = scpx * scpy
    scp2i = 1.0 / np.maximum(scp2, epsil)
    aspux = np.minimum(np.maximum(scux, scuy), np.minimum(scux, scuy) * aspmax) / np.maximum(scux, epsil)
    aspvy = np.minimum(np.maximum(scvx, scvy), np.minimum(scvx, scvy) * aspmax) / np.maximum(scvy, epsil)
    return scp2, scp2i, aspux, aspvy
