// Host front ends of the synthetic-state generator (include/hycom_tsadvc_synth.h): pure C++ on top of
// synth.h, no CUDA.  Included by csrc/synth.cu (product library) and by oracle/synth_host.cpp (a
// CUDA-free build for the reference arm of bench.py, so that arm maps nothing of the product).
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/hycom_tsadvc_synth.h"
#include "synth.h"

namespace {

synth::Cfg to_cfg(const hycom_synth_cfg& c) {
  synth::Cfg k;
  k.itdm = c.itdm; k.jtdm = c.jtdm; k.kdm = c.kdm; k.nreg = c.nreg;
  k.ntracr = c.ntracr; k.pad = 0; k.seed = c.seed;
  k.dx0 = c.dx0; k.dy0 = c.dy0; k.delt1 = c.delt1;
  return k;
}
synth::Tile to_tile(const hycom_synth_tile& t) {
  synth::Tile k;
  k.idm = t.idm; k.jdm = t.jdm; k.nbdy = t.nbdy; k.ii = t.ii; k.jj = t.jj;
  k.i0 = t.i0; k.j0 = t.j0; k.pad = 0;
  return k;
}

}  // namespace

extern "C" {

int hycom_synth_sea_mask(const hycom_synth_cfg* cfg, uint8_t* sea) {
  if (!cfg || !sea) return 1;
  const synth::Cfg c = to_cfg(*cfg);
  const int ni = c.itdm, nj = c.jtdm;
  const bool per_i = !(c.nreg == 0 || c.nreg == 4), per_j = c.nreg > 2;
  for (int j = 1; j <= nj; ++j)
    for (int i = 1; i <= ni; ++i) {
      // a closed basin needs its last column/row land, otherwise bigrid infers
      // periodicity (bigrid.F90:25-45)
      const bool edge = (!per_i && i == ni) || (!per_j && j == nj);
      sea[(size_t)(j - 1) * ni + (i - 1)] = edge ? 0 : 1;
    }
  const long area = (long)ni * nj;
  long nisl = area / 40000;
  if (nisl < 2) nisl = 2;
  double rmax = (double)(ni < nj ? ni : nj) / 12.0;
  if (rmax > 60.0) rmax = 60.0;
  if (rmax < 4.0) rmax = 4.0;
  for (long q = 0; q < nisl; ++q) {
    const double cx = 1.0 + synth::u01(c.seed, 90, (int)q, 0, 0) * ni;
    const double cy = 1.0 + synth::u01(c.seed, 91, (int)q, 0, 0) * nj;
    const double rr = 3.0 + (rmax - 3.0) * synth::u01(c.seed, 92, (int)q, 0, 0);
    const int ilo = (int)(cx - rr) - 1, ihi = (int)(cx + rr) + 1;
    const int jlo = (int)(cy - rr) - 1, jhi = (int)(cy + rr) + 1;
    for (int j = jlo; j <= jhi; ++j)
      for (int i = ilo; i <= ihi; ++i) {
        const double dx = i - cx, dy = j - cy;
        if (dx * dx + dy * dy > rr * rr) continue;
        const int wi = synth::wrap(i, ni, per_i), wj = synth::wrap(j, nj, per_j);
        if (wi == 0 || wj == 0) continue;
        sea[(size_t)(wj - 1) * ni + (wi - 1)] = 0;
      }
  }
  // no 1-point seas or single-width inlets: bigrid aborts on 4 land neighbours
  // and warns on 3 (bigrid.F90:156-191)
  bool changed = true;
  while (changed) {
    changed = false;
    for (int j = 1; j <= nj; ++j)
      for (int i = 1; i <= ni; ++i) {
        if (!sea[(size_t)(j - 1) * ni + (i - 1)]) continue;
        int nland = 0;
        nland += !synth::is_sea(c, sea, i - 1, j);
        nland += !synth::is_sea(c, sea, i + 1, j);
        nland += !synth::is_sea(c, sea, i, j - 1);
        nland += !synth::is_sea(c, sea, i, j + 1);
        if (nland >= 3) {
          sea[(size_t)(j - 1) * ni + (i - 1)] = 0;
          changed = true;
        }
      }
  }
  return 0;
}

int hycom_synth_fill_host(const hycom_synth_cfg* cfg, const hycom_synth_tile* tile,
                          const uint8_t* sea, int32_t field, int32_t ktr, int32_t lev, int32_t k0,
                          int32_t nk, int32_t halo_mode, double fill, double* dst) {
  if (!cfg || !tile || !sea || !dst || nk < 1) return 1;
  const synth::Cfg c = to_cfg(*cfg);
  const synth::Tile t = to_tile(*tile);
  const int ncols = t.idm + 2 * t.nbdy, nrows = t.jdm + 2 * t.nbdy;
  for (int s = 0; s < nk; ++s) {
    double* d = dst + (size_t)ncols * nrows * s;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < nrows; ++r)
      for (int col = 0; col < ncols; ++col)
        d[(size_t)r * ncols + col] = synth::tile_value(c, t, sea, field, ktr, lev, col + 1 - t.nbdy,
                                                       r + 1 - t.nbdy, k0 + s, halo_mode, fill);
  }
  return 0;
}

}  // extern "C"
