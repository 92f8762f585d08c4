// Synthetic-state generator: host and device front ends of synth.h.
// Used by bench.py, __graft_entry__.smoke() and the tests to build inputs of
// the named grid shapes (there are no input decks in the reference repository).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/hycom_tsadvc_synth.h"
#include "synth.h"
#include "tsadvc_handle.h"

namespace {

__global__ void k_synth_fill(synth::Cfg c, synth::Tile t, const uint8_t* __restrict__ sea,
                             int field, int ktr, int lev, int k0, int nk, int halo_mode,
                             double fill, double* __restrict__ dst, int pitch, long slab) {
  const int ncols = t.idm + 2 * t.nbdy, nrows = t.jdm + 2 * t.nbdy;
  const long per = (long)ncols * nrows;
  const long total = per * nk;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < total;
       q += (long)gridDim.x * blockDim.x) {
    const int s = (int)(q / per);
    const long w = q - (long)s * per;
    const int r = (int)(w / ncols), col = (int)(w % ncols);
    const int i = col + 1 - t.nbdy, j = r + 1 - t.nbdy;
    dst[slab * s + (long)r * pitch + col] =
        synth::tile_value(c, t, sea, field, ktr, lev, i, j, k0 + s, halo_mode, fill);
  }
}

}  // namespace

// host front ends (hycom_synth_sea_mask, hycom_synth_fill_host): shared with oracle/synth_host.cpp, which
// builds them without CUDA for the reference arm of bench.py
#include "synth_host.inl"

extern "C" {

int hycom_tsadvc_synth_set_sea(hycom_tsadvc_handle* h, const hycom_synth_cfg* cfg,
                               const uint8_t* sea) {
  if (!h || !cfg || !sea) return HYCOM_TSADVC_EINVAL;
  const size_t n = (size_t)cfg->itdm * cfg->jtdm;
  if (cudaSetDevice(h->d.device) != cudaSuccess) return HYCOM_TSADVC_ECUDA;
  if (h->d_sea) { cudaFree(h->d_sea); h->d_sea = nullptr; }
  if (cudaMalloc((void**)&h->d_sea, n) != cudaSuccess) return HYCOM_TSADVC_ENOMEM;
  h->bytes += (int64_t)n;
  if (cudaMemcpyAsync(h->d_sea, sea, n, cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
    return HYCOM_TSADVC_ECUDA;
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return HYCOM_TSADVC_ECUDA;
  return 0;
}

int hycom_tsadvc_synth_fill(hycom_tsadvc_handle* h, const hycom_synth_cfg* cfg, int32_t field,
                            int32_t ktr, int32_t tlev, int32_t lev, int32_t halo_mode,
                            double fill) {
  if (!h || !cfg || !h->d_sea) return HYCOM_TSADVC_EINVAL;
  void* base;
  int64_t pitch;
  // the generator's oneta id (HYCOM_SYNTH_ONETA) fills the one-slab-per-slot oneta mirror
  const bool is_oneta = field == HYCOM_SYNTH_ONETA;
  int rc = hycom_tsadvc_device_slab(h, is_oneta ? HYCOM_F_ONETA : field, ktr, tlev, 1, &base, &pitch);
  if (rc) return rc;
  const synth::Cfg c = to_cfg(*cfg);
  synth::Tile t;
  t.idm = h->d.idm; t.jdm = h->d.jdm; t.nbdy = h->d.nbdy; t.ii = h->d.ii; t.jj = h->d.jj;
  t.i0 = h->d.i0; t.j0 = h->d.j0; t.pad = 0;
  k_synth_fill<<<148 * 8, 256, 0, h->stream>>>(c, t, h->d_sea, field, ktr, lev, 1, is_oneta ? 1 : h->d.kdm,
                                                halo_mode, fill, (double*)base, (int)pitch, h->slab);
  h->launches += 1;
  if (cudaGetLastError() != cudaSuccess) return HYCOM_TSADVC_ECUDA;
  return 0;
}

// values of generator field `gen` (all layers of the destination, time level `lev`), times `scale`, written into
// ANY mirror: how bench.py fabricates the operands of cnuity (u, v, dpu, dpv, ...) on the device for its timing
namespace {
__global__ void k_scale(double* a, long n, double s) {
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x) a[q] = a[q] * s;
}
}  // namespace

int hycom_tsadvc_synth_fill_to(hycom_tsadvc_handle* h, const hycom_synth_cfg* cfg, int32_t gen, int32_t ktr,
                               int32_t lev, int32_t halo_mode, double scale, int32_t dst_field, int32_t dst_tlev,
                               int32_t k0, int32_t nk) {
  if (!h || !cfg || !h->d_sea) return HYCOM_TSADVC_EINVAL;
  void* base;
  int64_t pitch;
  int rc = hycom_tsadvc_device_slab(h, dst_field, 0, dst_tlev, k0, &base, &pitch);
  if (rc) return rc;
  const synth::Cfg c = to_cfg(*cfg);
  synth::Tile t;
  t.idm = h->d.idm; t.jdm = h->d.jdm; t.nbdy = h->d.nbdy; t.ii = h->d.ii; t.jj = h->d.jj;
  t.i0 = h->d.i0; t.j0 = h->d.j0; t.pad = 0;
  k_synth_fill<<<148 * 8, 256, 0, h->stream>>>(c, t, h->d_sea, gen, ktr, lev, 1, nk, halo_mode, 0.0, (double*)base,
                                                (int)pitch, h->slab);
  if (scale != 1.0) k_scale<<<148 * 8, 256, 0, h->stream>>>((double*)base, h->slab * nk, scale);
  h->launches += 2;
  if (cudaGetLastError() != cudaSuccess) return HYCOM_TSADVC_ECUDA;
  return 0;
}

}  // extern "C"
