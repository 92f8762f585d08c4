// C ABI of the B200 tsadvc path: device mirrors of the mod_cb_arrays fields,
// host<->device copies, and the tsadvc(m,n) driver (mod_tsadvc.F90:1708-2132)
// around the marching kernels.  See include/hycom_tsadvc_b200.h.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <vector>

#include "../../include/hycom_tsadvc_b200.h"
#include "tsadvc_dev.h"
#include "tsadvc_handle.h"
#include "tsadvc_launch.h"
#include "xc_comm.h"

using namespace tsadvc;

namespace {

char g_err[512] = "";

}  // namespace

namespace tsadvc {

int fail(hycom_tsadvc_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  snprintf(g_err, sizeof g_err, "%s", buf);
  if (h) snprintf(h->err, sizeof h->err, "%s", buf);
  return code;
}

}  // namespace tsadvc

namespace {

#define CU(h, call) TSADVC_CU(h, call)

int dalloc(hycom_tsadvc_handle* h, void** p, size_t nbytes, bool zero) {
  CU(h, cudaSetDevice(h->d.device));
  cudaError_t e = cudaMalloc(p, nbytes);
  if (e != cudaSuccess)
    return fail(h, HYCOM_TSADVC_ENOMEM, "cudaMalloc(%zu) failed: %s", nbytes,
                cudaGetErrorString(e));
  h->bytes += (int64_t)nbytes;
  if (zero) CU(h, cudaMemsetAsync(*p, 0, nbytes, h->stream));
  return 0;
}

// field buffers: one zeroed guard row on each side (see tsadvc_handle.h)
int dalloc_field(hycom_tsadvc_handle* h, double** p, size_t ndoubles) {
  void* raw = nullptr;
  const size_t g = (size_t)(h->pitch > 64 ? h->pitch : 64);   // at least one 64-column window
  int rc = dalloc(h, &raw, sizeof(double) * (ndoubles + 2 * g), true);
  if (rc) return rc;
  h->raw_allocs.push_back(raw);
  *p = (double*)raw + g;
  return 0;
}

Mirror* mirror_of(hycom_tsadvc_handle* h, int field, int ktr) {
  switch (field) {
    case HYCOM_F_TEMP: return &h->temp;
    case HYCOM_F_SALN: return &h->saln;
    case HYCOM_F_TH3D: return &h->th3d;
    case HYCOM_F_DP: return &h->dp;
    case HYCOM_F_UFLX: return &h->uflx;
    case HYCOM_F_VFLX: return &h->vflx;
    case HYCOM_F_TRACER:
      if (ktr >= 1 && ktr <= h->d.ntracr) return &h->tracer[ktr - 1];
      return nullptr;
    case HYCOM_F_ONETA: return &h->oneta;
    case HYCOM_F_THETA: return &h->theta;
    case HYCOM_F_Q2: return &h->q2;
    case HYCOM_F_Q2L: return &h->q2l;
    case HYCOM_F_DPO: return &h->dpo;
    case HYCOM_F_ONETAO: return &h->onetao;
    case HYCOM_F_PBAVG: return &h->pbavg;
    case HYCOM_F_PBOT: return &h->pbot;
    case HYCOM_F_OTEMP: return &h->otemp;
    case HYCOM_F_OSALN: return &h->osaln;
    case HYCOM_F_OTH3D: return &h->oth3d;
    case HYCOM_F_OQ2: return &h->oq2;
    case HYCOM_F_OQ2L: return &h->oq2l;
    case HYCOM_F_OTRACER:
      if (ktr >= 1 && ktr <= h->d.ntracr) return &h->otracer[ktr - 1];
      return nullptr;
    case HYCOM_F_U: return &h->u;
    case HYCOM_F_V: return &h->v;
    case HYCOM_F_DPU: return &h->dpu;
    case HYCOM_F_DPV: return &h->dpv;
    case HYCOM_F_UBAVG: return &h->ubavg;
    case HYCOM_F_VBAVG: return &h->vbavg;
    case HYCOM_F_DEPTHU: return &h->depthu;
    case HYCOM_F_DEPTHV: return &h->depthv;
    case HYCOM_F_P: return &h->p;
    case HYCOM_F_DPMIXL: return &h->dpmixl;
    case HYCOM_F_UFLXAV: return &h->uflxav;
    case HYCOM_F_VFLXAV: return &h->vflxav;
    case HYCOM_F_DPAV: return &h->dpav;
    case HYCOM_F_UTOTN: return &h->utotn;
    case HYCOM_F_VTOTN: return &h->vtotn;
    case HYCOM_F_DPMOLD: return &h->dpmold;
    case HYCOM_F_THKDF4U: return &h->thkdf4u;
    case HYCOM_F_THKDF4V: return &h->thkdf4v;
  }
  return nullptr;
}
bool is3d(int field) {   // one time level only
  return field == HYCOM_F_UFLX || field == HYCOM_F_VFLX || field == HYCOM_F_THETA || field == HYCOM_F_PBAVG ||
         field == HYCOM_F_PBOT || (field >= HYCOM_F_OTEMP && field <= HYCOM_F_OQ2L) ||
         (field >= HYCOM_F_UBAVG && field <= HYCOM_F_P) || (field >= HYCOM_F_UFLXAV && field <= HYCOM_F_THKDF4V);
}
// slabs per time slot of a mirror
int nlayers_of(const hycom_tsadvc_handle* h, int field) {
  if (field == HYCOM_F_ONETA || field == HYCOM_F_ONETAO || field == HYCOM_F_PBOT || field == HYCOM_F_DEPTHU ||
      field == HYCOM_F_DEPTHV || field == HYCOM_F_DPMIXL || field == HYCOM_F_UTOTN || field == HYCOM_F_VTOTN ||
      field == HYCOM_F_DPMOLD || field == HYCOM_F_THKDF4U || field == HYCOM_F_THKDF4V)
    return 1;
  if (field == HYCOM_F_PBAVG || field == HYCOM_F_UBAVG || field == HYCOM_F_VBAVG) return 3;
  if (field == HYCOM_F_P) return h->d.kdm + 1;
  if (field == HYCOM_F_Q2 || field == HYCOM_F_Q2L || field == HYCOM_F_OQ2 || field == HYCOM_F_OQ2L)
    return h->d.kdm + 2;   // layers 0..kk+1
  return h->d.kdm;
}
// slab index of model layer k = 1 inside one time slot of a mirror
int layer1_of(int field) { return (field == HYCOM_F_Q2 || field == HYCOM_F_Q2L) ? 1 : 0; }

// device pointer of slot tlev (1,2) of a mirror, allocated on first use
int slot(hycom_tsadvc_handle* h, int field, int ktr, int tlev, double** out) {
  Mirror* mi = mirror_of(h, field, ktr);
  if (!mi) return fail(h, HYCOM_TSADVC_EINVAL, "bad field %d / ktr %d", field, ktr);
  const int s = is3d(field) ? 0 : tlev - 1;
  if (s < 0 || s > 1) return fail(h, HYCOM_TSADVC_EINVAL, "bad time slot %d", tlev);
  if (!mi->lev[s]) {
    if (field == HYCOM_F_DP || field == HYCOM_F_UFLX || field == HYCOM_F_VFLX) {
      const size_t K = (size_t)h->slab * h->d.kdm;
      int rc = dalloc_field(h, &h->flux_block, 4 * K);
      if (rc) return rc;
      h->dp.lev[0] = h->flux_block;
      h->uflx.lev[0] = h->flux_block + K;
      h->vflx.lev[0] = h->flux_block + 2 * K;
      h->dp.lev[1] = h->flux_block + 3 * K;
    } else {
      int rc = dalloc_field(h, &mi->lev[s], (size_t)h->slab * nlayers_of(h, field));
      if (rc) return rc;
    }
  }
  *out = mi->lev[s];
  return 0;
}

int spare_of(hycom_tsadvc_handle* h, Mirror* mi, double** out) {
  if (!mi->spare) {
    const bool my = mi == &h->q2 || mi == &h->q2l;
    int rc = dalloc_field(h, &mi->spare, (size_t)h->slab * (h->d.kdm + (my ? 2 : 0)));
    if (rc) return rc;
  }
  *out = mi->spare;
  return 0;
}

int up2d(hycom_tsadvc_handle* h, double** dst, const double* src) {
  if (!*dst) {
    int rc = dalloc(h, (void**)dst, sizeof(double) * (size_t)h->slab, true);
    if (rc) return rc;
  }
  CU(h, cudaMemcpy2DAsync(*dst, sizeof(double) * h->pitch, src, sizeof(double) * h->ncols,
                          sizeof(double) * h->ncols, h->nrows, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

// per-layer salinity range over sea cells with dp > onemm (mod_tsadvc.F90:2065-2084)
__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) < v) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

// SplitMix64 finaliser
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// sum over the cells tsadvc writes (M_OUT: interior sea points) of mix64(bits ^ mix64(global index)):
// independent of the summation order and of the tiling (hycom_tsadvc_checksum)
__global__ void __launch_bounds__(256) k_checksum(const double* __restrict__ fld, const uint8_t* __restrict__ mask,
                                                   long slab, int pitch, int nrows, int nbdy, int i0, int j0,
                                                   int itdm, int jtdm, int nk, unsigned long long* out) {
  const long per = (long)pitch * nrows, total = per * nk;
  unsigned long long acc = 0;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int k = (int)(t / per);
    const long q = t - (long)k * per;
    if (!(mask[q] & M_OUT)) continue;
    const int r = (int)(q / pitch), c = (int)(q - (long)r * pitch);
    const long gi = i0 + (c - nbdy), gj = j0 + (r - nbdy);   // 0-based global cell
    const unsigned long long cell = (unsigned long long)(gi + (long)itdm * (gj + (long)jtdm * k));
    const double v = fld[slab * k + q];
    const unsigned long long bits = v == 0.0 ? 0ull : (unsigned long long)__double_as_longlong(v);   // -0.0 == 0.0
    acc += mix64(bits ^ mix64(cell));
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

__global__ void k_minmax_init(double* mm, int kk) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < kk) {
    mm[k] = 999.;        // :2069
    mm[kk + k] = -999.;  // :2070
  }
}

// two cells per thread (16-byte loads), 8 independent loads in flight per thread
__global__ void __launch_bounds__(256) k_saln_minmax(const double* __restrict__ saln,
                                                      const double* __restrict__ dp,
                                                      const uint8_t* __restrict__ mask, long slab,
                                                      double onemm, double* mm, int kk) {
  const int k = blockIdx.y;
  const double2* s = reinterpret_cast<const double2*>(saln + slab * k);
  const double2* d = reinterpret_cast<const double2*>(dp + slab * k);
  const uchar2* mk2 = reinterpret_cast<const uchar2*>(mask);
  double lo = 999., hi = -999.;
  const long n2 = slab / 2;   // pitch is even
  const long stride = (long)gridDim.x * blockDim.x;
  for (long q0 = (long)blockIdx.x * blockDim.x + threadIdx.x; q0 < n2; q0 += 4 * stride) {
    double2 sv[4], dv[4];
    uchar2 mv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long q = q0 + u * stride;
      const bool ok = q < n2;
      mv[u] = ok ? __ldg(mk2 + q) : make_uchar2(0, 0);
      sv[u] = ok ? __ldg(s + q) : make_double2(0., 0.);
      dv[u] = ok ? __ldg(d + q) : make_double2(0., 0.);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if ((mv[u].x & M_OUT) && dv[u].x > onemm) { lo = lo < sv[u].x ? lo : sv[u].x; hi = hi > sv[u].x ? hi : sv[u].x; }
      if ((mv[u].y & M_OUT) && dv[u].y > onemm) { lo = lo < sv[u].y ? lo : sv[u].y; hi = hi > sv[u].y ? hi : sv[u].y; }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double l2 = __shfl_xor_sync(0xffffffffu, lo, o);
    const double h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = lo < l2 ? lo : l2;
    hi = hi > h2 ? hi : h2;
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_f64(&mm[k], lo);
    atomic_max_f64(&mm[kk + k], hi);
  }
}

}  // namespace

extern "C" {

int hycom_tsadvc_abi_version(void) { return HYCOM_TSADVC_ABI_VERSION; }

const char* hycom_tsadvc_last_error(const hycom_tsadvc_handle* h) { return h ? h->err : g_err; }

int hycom_tsadvc_create(const hycom_tsadvc_dims* dims, hycom_tsadvc_handle** out) {
  if (!dims || !out) return fail(nullptr, HYCOM_TSADVC_EINVAL, "null argument");
  const hycom_tsadvc_dims& d = *dims;
  if (d.idm < 1 || d.jdm < 1 || d.kdm < 1 || d.nbdy < 0 || d.ii < 1 || d.jj < 1 ||
      d.ii > d.idm || d.jj > d.jdm || d.ntracr < 0 || d.ntracr > HYCOM_TSADVC_MXTRCR)
    return fail(nullptr, HYCOM_TSADVC_EINVAL, "bad dimensions idm=%d jdm=%d kdm=%d nbdy=%d ii=%d jj=%d ntracr=%d",
                d.idm, d.jdm, d.kdm, d.nbdy, d.ii, d.jj, d.ntracr);
  // xcspmd's conditions for the arctic patch (mod_xc_mp.h:2614-2643, :2817-2823): ipr even or 1, and
  // every tile of the top row itdm/ipr wide (twins exchange whole, mirrored rows)
  if (d.nreg == 2 && d.ipr * d.jpr != 1) {
    if (d.ipr > 1 && d.ipr % 2 != 0)
      return fail(nullptr, HYCOM_TSADVC_EINVAL, "Error in xcspmd (arctic) - ipr must be even (ipr=%d)", d.ipr);
    // (every rank refuses an uneven split, not only the top row: the call is collective)
    if (d.itdm % d.ipr != 0 || (d.nproc == d.jpr && d.ii != d.itdm / d.ipr))
      return fail(nullptr, HYCOM_TSADVC_EINVAL, "error - arctic patch tiles should have ii = %d (ii=%d, itdm=%d, ipr=%d)",
                  d.itdm / d.ipr, d.ii, d.itdm, d.ipr);
  }
  if (d.nreg < 0 || d.nreg > 4) return fail(nullptr, HYCOM_TSADVC_EINVAL, "bad nreg %d", d.nreg);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return fail(nullptr, HYCOM_TSADVC_ECUDA, "no CUDA device: %s (this library has no CPU path)",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (d.device < 0 || d.device >= ndev)
    return fail(nullptr, HYCOM_TSADVC_EINVAL, "device %d out of range (%d devices)", d.device, ndev);
  hycom_tsadvc_handle* h = new (std::nothrow) hycom_tsadvc_handle();
  if (!h) return fail(nullptr, HYCOM_TSADVC_ENOMEM, "out of host memory");
  h->d = d;
  h->err[0] = 0;
  h->ncols = d.idm + 2 * d.nbdy;
  h->nrows = d.jdm + 2 * d.nbdy;
  h->pitch = (h->ncols + 15) & ~15;  // rows 128-byte aligned; byte planes 16-byte (TMA strides)
  h->slab = (long)h->pitch * h->nrows;
  CU(h, cudaSetDevice(d.device));
  CU(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  *out = h;
  return 0;
}

int hycom_tsadvc_destroy(hycom_tsadvc_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->d.device);
  cudaDeviceSynchronize();
  xc_detach(h);
  if (h->range_host) cudaFreeHost(h->range_host);
  if (h->ev_range) cudaEventDestroy(h->ev_range);
  if (h->ev_xc) cudaEventDestroy(h->ev_xc);
  cudaFree(h->d_cksum); cudaFree(h->d_dpkmin);
  for (void* raw : h->raw_allocs) cudaFree(raw);  // field mirrors, flux block, static block
  cudaFree(h->mask); cudaFree(h->scuy); cudaFree(h->scvx);
  for (auto& kv : h->seg_cache) { cudaFree(kv.second.d[0]); cudaFree(kv.second.d[1]); }
  cudaFree(h->aspux); cudaFree(h->aspvy); cudaFree(h->d_minmax); cudaFree(h->d_sea);
  for (auto* v : {&h->ev_pending, &h->ev_free})
    for (auto& ev : *v) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  for (cudaEvent_t e : h->ev_chunk) cudaEventDestroy(e);
  if (h->ev_frame) cudaEventDestroy(h->ev_frame);
  if (h->up_stream) cudaStreamDestroy(h->up_stream);
  if (h->down_stream) cudaStreamDestroy(h->down_stream);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int hycom_tsadvc_set_stream(hycom_tsadvc_handle* h, void* cuda_stream) {
  if (!h) return fail(nullptr, HYCOM_TSADVC_EINVAL, "null handle");
  if (h->own_stream && h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  if (cuda_stream) {
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
  } else {
    CU(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  return 0;
}

int hycom_tsadvc_set_frame_stream(hycom_tsadvc_handle* h, void* cuda_stream) {
  if (!h) return fail(nullptr, HYCOM_TSADVC_EINVAL, "null handle");
  h->frame_stream = (cudaStream_t)cuda_stream;
  return 0;
}

int hycom_tsadvc_synchronize(hycom_tsadvc_handle* h) {
  if (!h) return fail(nullptr, HYCOM_TSADVC_EINVAL, "null handle");
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int64_t hycom_tsadvc_device_bytes(const hycom_tsadvc_handle* h) { return h ? h->bytes : 0; }

int hycom_tsadvc_set_timing(hycom_tsadvc_handle* h, int32_t enable) {
  if (!h) return fail(nullptr, HYCOM_TSADVC_EINVAL, "null handle");
  h->timing = enable != 0;
  return 0;
}

int hycom_tsadvc_get_timing(hycom_tsadvc_handle* h, double* march_ms, int64_t* march_launches,
                            int32_t reset) {
  if (!h) return fail(nullptr, HYCOM_TSADVC_EINVAL, "null handle");
  CU(h, cudaSetDevice(h->d.device));
  CU(h, cudaStreamSynchronize(h->stream));
  for (auto& ev : h->ev_pending) {
    float ms = 0.f;
    CU(h, cudaEventElapsedTime(&ms, ev.first, ev.second));
    h->march_ms += ms;
    h->march_n += 1;
    h->ev_free.push_back(ev);
  }
  h->ev_pending.clear();
  if (march_ms) *march_ms = h->march_ms;
  if (march_launches) *march_launches = h->march_n;
  if (reset) { h->march_ms = 0.0; h->march_n = 0; }
  return 0;
}
int64_t hycom_tsadvc_launch_count(const hycom_tsadvc_handle* h) { return h ? h->launches : 0; }

int hycom_tsadvc_set_static(hycom_tsadvc_handle* h, const double* scp2, const double* scp2i,
                            const double* scuy, const double* scvx, const double* aspux,
                            const double* aspvy, const int32_t* ip, const int32_t* iu,
                            const int32_t* iv) {
  if (!h || !scp2 || !scp2i || !ip || !iu || !iv)
    return fail(h, HYCOM_TSADVC_EINVAL, "set_static: scp2, scp2i, ip, iu, iv are required");
  CU(h, cudaSetDevice(h->d.device));
  int rc;
  if (!h->static_block) {
    if ((rc = dalloc_field(h, &h->static_block, 3 * (size_t)h->slab))) return rc;
    h->scp2i = h->static_block;
    h->scp2 = h->static_block + h->slab;
  }
  if ((rc = up2d(h, &h->scp2, scp2))) return rc;
  if ((rc = up2d(h, &h->scp2i, scp2i))) return rc;
  if (scuy && (rc = up2d(h, &h->scuy, scuy))) return rc;
  if (scvx && (rc = up2d(h, &h->scvx, scvx))) return rc;
  if (aspux && (rc = up2d(h, &h->aspux, aspux))) return rc;
  if (aspvy && (rc = up2d(h, &h->aspvy, aspvy))) return rc;
  // pack ip/iu/iv and the sea-only neighbour flags (ipim1.. of bigrid.F90:316-341)
  // into one byte per cell
  const int nc = h->ncols, nr = h->nrows, nb = h->d.nbdy;
  std::vector<uint8_t> m((size_t)h->slab, 0);
  for (int r = 0; r < nr; ++r)
    for (int c = 0; c < nc; ++c) {
      const size_t q = (size_t)r * nc + c;
      unsigned b = 0;
      if (ip[q] != 0) b |= M_IP;
      if (iu[q] != 0) b |= M_IU;
      if (iv[q] != 0) b |= M_IV;
      if (c > 0 && ip[q - 1] != 0) b |= M_PW;
      if (c < nc - 1 && ip[q + 1] != 0) b |= M_PE;
      if (r > 0 && ip[q - nc] != 0) b |= M_PS;
      if (r < nr - 1 && ip[q + nc] != 0) b |= M_PN;
      const int i = c + 1 - nb, j = r + 1 - nb;
      if (ip[q] != 0 && i >= 1 && i <= h->d.ii && j >= 1 && j <= h->d.jj) b |= M_OUT;
      m[(size_t)r * h->pitch + c] = (uint8_t)b;
    }
  if (!h->mask && (rc = dalloc(h, (void**)&h->mask, (size_t)h->slab, true))) return rc;
  CU(h, cudaMemcpyAsync(h->mask, m.data(), (size_t)h->slab, cudaMemcpyHostToDevice, h->stream));
  h->mask_host = m;
  for (auto& kv : h->seg_cache) { cudaFree(kv.second.d[0]); cudaFree(kv.second.d[1]); }
  h->seg_cache.clear();
  {  // third plane of the static block: the mask byte in the low bits of a 64-bit word
    std::vector<uint64_t> m64((size_t)h->slab);
    for (size_t q = 0; q < m64.size(); ++q) m64[q] = m[q];
    CU(h, cudaMemcpyAsync(h->static_block + 2 * h->slab, m64.data(), sizeof(uint64_t) * m64.size(),
                          cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  CU(h, cudaStreamSynchronize(h->stream));
  h->have_static = true;
  return 0;
}

static int upload_on(hycom_tsadvc_handle* h, int32_t field, int32_t ktr, int32_t tlev, int32_t k0,
                     int32_t nk, const double* host, cudaStream_t st);

int hycom_tsadvc_upload(hycom_tsadvc_handle* h, int32_t field, int32_t ktr, int32_t tlev,
                        int32_t k0, int32_t nk, const double* host) {
  return upload_on(h, field, ktr, tlev, k0, nk, host, h ? h->stream : nullptr);
}

// `host` = the nk slabs starting at layer k0
static int upload_on(hycom_tsadvc_handle* h, int32_t field, int32_t ktr, int32_t tlev, int32_t k0,
                     int32_t nk, const double* host, cudaStream_t st) {
  if (!h || !host) return fail(h, HYCOM_TSADVC_EINVAL, "upload: null argument");
  if (k0 < 1 || nk < 1 || k0 + nk - 1 > nlayers_of(h, field))
    return fail(h, HYCOM_TSADVC_EINVAL, "upload: layers %d..%d out of 1..%d", k0, k0 + nk - 1,
                nlayers_of(h, field));
  double* base;
  int rc = slot(h, field, ktr, tlev, &base);
  if (rc) return rc;
  CU(h, cudaMemcpy2DAsync(base + h->slab * (k0 - 1), sizeof(double) * h->pitch, host,
                          sizeof(double) * h->ncols, sizeof(double) * h->ncols,
                          (size_t)h->nrows * nk, cudaMemcpyHostToDevice, st));
  return 0;
}

int hycom_tsadvc_download(hycom_tsadvc_handle* h, int32_t field, int32_t ktr, int32_t tlev,
                          int32_t k0, int32_t nk, double* host) {
  if (!h || !host) return fail(h, HYCOM_TSADVC_EINVAL, "download: null argument");
  if (k0 < 1 || nk < 1 || k0 + nk - 1 > nlayers_of(h, field))
    return fail(h, HYCOM_TSADVC_EINVAL, "download: layers %d..%d out of 1..%d", k0, k0 + nk - 1,
                nlayers_of(h, field));
  double* base;
  int rc = slot(h, field, ktr, tlev, &base);
  if (rc) return rc;
  CU(h, cudaMemcpy2DAsync(host, sizeof(double) * h->ncols, base + h->slab * (k0 - 1),
                          sizeof(double) * h->pitch, sizeof(double) * h->ncols,
                          (size_t)h->nrows * nk, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int hycom_tsadvc_device_slab(hycom_tsadvc_handle* h, int32_t field, int32_t ktr, int32_t tlev,
                             int32_t k, void** dev_ptr, int64_t* pitch) {
  if (!h || !dev_ptr) return fail(h, HYCOM_TSADVC_EINVAL, "device_slab: null argument");
  if (k < 1 || k > nlayers_of(h, field)) return fail(h, HYCOM_TSADVC_EINVAL, "device_slab: bad layer %d", k);
  double* base;
  int rc = slot(h, field, ktr, tlev, &base);
  if (rc) return rc;
  *dev_ptr = base + h->slab * (k - 1);
  if (pitch) *pitch = h->pitch;
  return 0;
}

static int halo_local_range(hycom_tsadvc_handle* h, int32_t field, int32_t ktr, int32_t tlev,
                            int32_t mh, int32_t nh, int k0, int nk);

int hycom_tsadvc_halo_local(hycom_tsadvc_handle* h, int32_t field, int32_t ktr, int32_t tlev,
                            int32_t mh, int32_t nh) {
  if (!h) return fail(nullptr, HYCOM_TSADVC_EINVAL, "null handle");
  return halo_local_range(h, field, ktr, tlev, mh, nh, 0, nlayers_of(h, field));
}

// layers k0 .. k0+nk-1 (0-based) only
static int halo_local_range(hycom_tsadvc_handle* h, int32_t field, int32_t ktr, int32_t tlev,
                            int32_t mh, int32_t nh, int k0, int nk) {
  const int nreg = h->d.nreg;
  // xctilr's itype (mod_tsadvc.F90:1829-1836): halo_ps = 1 for scalars, halo_uv = 13 / halo_vv = 14 for
  // the mass fluxes; it only matters across the arctic (nreg = 2)
  const int itype = (field == HYCOM_F_UFLX || field == HYCOM_F_U || field == HYCOM_F_UBAVG) ? 13
                    : (field == HYCOM_F_VFLX || field == HYCOM_F_V || field == HYCOM_F_VBAVG) ? 14
                    : field == HYCOM_F_DPU ? 3 : field == HYCOM_F_DPV ? 4 : 1;   // halo_uv, halo_vv, halo_us, halo_vs, halo_ps
  const int per_i = !(nreg == 0 || nreg == 4), per_j = nreg == 2 ? 100 + itype : nreg > 2;
  const int t0 = is3d(field) ? 1 : (tlev == 0 ? 1 : tlev);
  const int t1 = is3d(field) ? 1 : (tlev == 0 ? 2 : tlev);
  for (int t = t0; t <= t1; ++t) {
    double* base;
    int rc = slot(h, field, ktr, t, &base);
    if (rc) return rc;
    const int nl = nk;
    base += h->slab * k0;
    rc = launch_halo_local(base, h->slab, nl, h->pitch, h->d.nbdy, h->d.ii, h->d.jj, mh, nh,
                           per_i, per_j, h->stream);
    if (!rc)
      rc = launch_halo_outer(base, h->slab, nl, h->pitch, h->nrows, h->d.nbdy, h->d.ii, h->d.jj,
                             mh, nh, h->stream);
    h->launches += 3;
    if (rc) return fail(h, HYCOM_TSADVC_ECUDA, "halo kernel launch failed: %s",
                        cudaGetErrorString((cudaError_t)rc));
  }
  return 0;
}

}  // extern "C"

namespace {

// an advected field: layers 1..nlay of the mirror, which start at slab `koff` of a time slot
struct Adv { int field, ktr; double posdef; int nlay; int koff = 0; };

// argument checks of tsadvc/advem (mod_tsadvc.F90:1815-1825, :159-166) and the list of
// advected fields (:1855-1857, :1969-2034)
int plan_step(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params* prm,
              std::vector<Adv>& adv, int& mbdy) {
  if (!h || !prm) return fail(h, HYCOM_TSADVC_EINVAL, "step: null argument");
  if (!h->have_static) return fail(h, HYCOM_TSADVC_EINVAL, "step: set_static has not been called");
  if (m < 1 || m > 2 || n < 1 || n > 2 || m == n)
    return fail(h, HYCOM_TSADVC_EINVAL, "step: bad leapfrog slots m=%d n=%d", m, n);
  const hycom_tsadvc_params& p = *prm;
  const int aadv = abs(p.advtyp);
  if (!(aadv == 0 || aadv == 1 || aadv == 2 || aadv == 4))
    return fail(h, HYCOM_TSADVC_EADVTYP, "error: advem called with advtyp =%4d", p.advtyp);
  mbdy = (aadv == 0) ? 2 : 5;
  if (h->d.nbdy < mbdy)
    return fail(h, HYCOM_TSADVC_ENBDY,
                "error: nbdy (dimensions.h) must be at least%3d for the advection scheme indicated by advtyp",
                mbdy);
  if (p.btrmas && aadv != 2)
    return fail(h, HYCOM_TSADVC_EUNSUPPORTED, "btrmas with advtyp=%d: only advem_fct2c (advtyp=2) reads onetamas here", p.advtyp);
  // blkdat.F90:1485 isopyc = nhybrd.eq.0 .and. .not.mxl_no; :295 hybrid = nhybrd.gt.0
  if (p.isopyc && (p.nhybrd != 0 || p.hybrid))
    return fail(h, HYCOM_TSADVC_EINVAL, "isopyc needs nhybrd = 0 and hybrid = .false. (blkdat.F90:1485): nhybrd=%d hybrid=%d",
                p.nhybrd, p.hybrid);
  if (p.isopyc && p.btrmas)
    return fail(h, HYCOM_TSADVC_EUNSUPPORTED, "isopyc with btrmas: advem_fct2c on smoothed layer-1 fluxes is not built");
  if (p.temdf2 > 0.0) {
    if (p.sigver < 1 || p.sigver > 8)
      return fail(h, HYCOM_TSADVC_EINVAL, "temdf2>0 needs the equation of state: sigver=%d not in 1..8", p.sigver);
    if (!h->scuy || !h->scvx || !h->aspux || !h->aspvy)
      return fail(h, HYCOM_TSADVC_EINVAL, "temdf2>0 needs scuy, scvx, aspux, aspvy (set_static)");
  }
  const int kk = h->d.kdm;
  const int nhyb = p.nhybrd < 0 ? 0 : (p.nhybrd > kk ? kk : p.nhybrd);
  adv.clear();
  // :1855-1857  latemp = k<=nhybrd & advflg==0; lath3d = (k<=nhybrd & advflg==1) | (k==1 & isopyc)
  if (p.advflg == 0) { if (nhyb > 0) adv.push_back({HYCOM_F_TEMP, 0, 256.0, nhyb}); }
  else               { if (nhyb > 0) adv.push_back({HYCOM_F_TH3D, 0, 32.0, nhyb}); }
  if (p.isopyc && !(p.advflg != 0 && nhyb > 0)) adv.push_back({HYCOM_F_TH3D, 0, 32.0, 1});
  adv.push_back({HYCOM_F_SALN, 0, 0.0, kk});
  for (int t = 1; t <= h->d.ntracr; ++t)
    adv.push_back({HYCOM_F_TRACER, t, p.trcflg[t - 1] == 2 ? 256.0 : 0.0, kk});
  if (p.mxlmy) {   // :2035-2048, pdq2 = 1.0 (:1762)
    adv.push_back({HYCOM_F_Q2, 0, 1.0, kk, 1});
    adv.push_back({HYCOM_F_Q2L, 0, 1.0, kk, 1});
  }
  if ((int)adv.size() > kMaxFields) return fail(h, HYCOM_TSADVC_EINVAL, "too many advected fields");
  return 0;
}

// the arrays xctilr is called on at mod_tsadvc.F90:1829-1836 (th3d is exchanged by the
// reference too but not read when advflg=0 and temdf2=0: left out of the multi-tile messages)
int halo_arrays(hycom_tsadvc_handle* h, const std::vector<Adv>& adv, int mbdy, HaloArrays& a) {
  memset(&a, 0, sizeof a);
  int rc;
  for (const Adv& f : adv)
    for (int t = 1; t <= 2; ++t) {
      if ((rc = slot(h, f.field, f.ktr, t, &a.base[a.narr]))) return rc;
      a.base[a.narr++] += h->slab * f.koff;   // q2,q2l: the advected layers 1..kk
    }
  for (int q = 0; q < a.narr; ++q) a.itype[q] = 1;                  // halo_ps (:1829-1836)
  a.itype[a.narr] = 13;                                             // halo_uv
  if ((rc = slot(h, HYCOM_F_UFLX, 0, 1, &a.base[a.narr++]))) return rc;
  a.itype[a.narr] = 14;                                             // halo_vv
  if ((rc = slot(h, HYCOM_F_VFLX, 0, 1, &a.base[a.narr++]))) return rc;
  a.kk = h->d.kdm; a.slab = h->slab; a.pitch = h->pitch; a.nrows = h->nrows; a.nbdy = h->d.nbdy;
  a.ii = h->d.ii; a.jj = h->d.jj; a.mh = mbdy; a.nh = mbdy;
  a.fold = arctic_fold(h->d);
  return 0;
}

// Row segments of an FCT2 or MPDATA launch.  Every (strip, row) of the rectangles of P is assigned to a run of
// rows: runs whose staged window (32*nc columns, rows j0-3 .. j1+2) holds nothing but mask bytes 0xff
// (sea, interior, sea on all four sides) are marched by the mask-free instantiation of the kernel, the
// rest by the general one.  The marching loop works in rounds of six rows, so the all-sea runs are cut
// to a multiple of six rows (no row beyond the run is ever computed); short runs are not worth the six
// rows of pipeline fill and stay with the general kernel.  Built once per (part, nc, chunk rows).
static int march_segments(hycom_tsadvc_handle* h, const MarchParams& P, int part, int chunk_rows, bool whole_default,
                          const hycom_tsadvc_handle::SegLists** out) {
  const char* cwh = getenv("HYCOM_TSADVC_SEG_WHOLE");
  const bool whole = cwh ? atoi(cwh) != 0 : whole_default;
  const long key = ((long)whole << 44) | ((long)part << 40) | ((long)P.nc << 32) | (long)chunk_rows;   // (the band is fixed per process)
  auto it = h->seg_cache.find(key);
  if (it != h->seg_cache.end()) { *out = &it->second; return 0; }
  const int use = strip_use(P.nc), lead = strip_lead(P.nc), wid = 32 * P.nc;
  const int pitch = h->pitch, nrows = h->nrows;
  const int kMinRows = 48;
  // The 64-column windows are not 128-byte aligned: a strip whose neighbour is elsewhere in j fetches 624
  // bytes per row and array for 464 useful ones.  Pieces that start on common band boundaries let
  // neighbouring strips march the same rows at the same time and share those lines through L2.  DRAM traffic of
  // the launch pair at GLBb0.08 (algorithmic 44.2 GB; profiles/r02q-u, r03j): maximal all-sea runs cut per
  // strip 58.4 GB; cut at common 256-row bands 53.8 GB; all-sea pieces = whole 252-row bands only 50.7 GB (the
  // default of FCT2/FCT4: 27 % instead of 76 % of the rows then take the mask-free body, +0.8 % time; MPDATA, whose
  // mask-free body saves more, keeps the maximal runs: 90.2 against 94.7 ms with 8 tracers); one general launch 47.5
  const char* cbnd = getenv("HYCOM_TSADVC_SEG_BAND");
  // (maximal runs - MPDATA - are cut chunk-long: 90.9 ms with 8 tracers against 92.6 at 256-row bands, profiles/r03y)
  int band = std::max(48, std::min(cbnd ? atoi(cbnd) : (whole ? 252 : chunk_rows), chunk_rows));
  if (whole) band = band / 6 * 6;
  std::vector<MarchSeg> seg[2];
  std::vector<char> good(nrows + 8), taken(nrows);
  for (int q = 0; q < P.nrect; ++q) {
    const MarchRect& R = P.rect[q];
    const int fast_piece = (std::min(R.chunk_rows, band) / 6) * 6;
    for (int st = R.strip0; st < R.strip0 + R.nstrips; ++st) {
      const int w0 = st * use - lead;
      const bool cols_ok = w0 >= 0 && w0 + wid <= pitch;
      for (int r = 0; r < nrows; ++r) {
        bool g = cols_ok;
        if (g) {
          const uint8_t* mrow = h->mask_host.data() + (size_t)r * pitch + w0;
          for (int c = 0; c < wid && g; ++c) g = mrow[c] == 0xff;
        }
        good[r] = g;
        taken[r] = 0;
      }
      if (whole) {
        for (int a = R.row0 / band * band; a < R.row1; a += band) {
          const int lo = std::max(a, R.row0), hi = std::min(a + band, R.row1);
          bool ok = (hi - lo) % 6 == 0 && hi - lo >= kMinRows && lo - 3 >= 0 && hi + 3 <= nrows;
          for (int t = lo - 3; ok && t < hi + 3; ++t) ok = good[t];
          if (!ok) continue;
          seg[0].push_back(MarchSeg{st, lo, hi, 0});
          for (int t = lo; t < hi; ++t) taken[t] = 1;
        }
      } else if (fast_piece >= kMinRows) {
        int r = 0;
        while (r < nrows) {
          if (!good[r]) { ++r; continue; }
          int b = r;
          while (b < nrows && good[b]) ++b;                  // good rows [r, b)
          int j0 = std::max(r + 3, R.row0), j1 = std::min(b - 3, R.row1);
          int len = ((j1 - j0) / 6) * 6;
          if (len >= kMinRows) {
            // pieces end at the common band boundaries (to within six rows), so that neighbouring strips
            // march the same rows at the same time and share the 128-byte lines their windows overlap in
            for (int a = j0; a < j0 + len;) {
              int e = a + ((((a / band) + 1) * band - a) + 5) / 6 * 6;
              if (e - a < 24) e += band / 6 * 6;
              if (j0 + len - e < 24) e = j0 + len;
              e = std::min(e, j0 + len);
              seg[0].push_back(MarchSeg{st, a, e, 0});
              for (int t = a; t < e; ++t) taken[t] = 1;
              a = e;
            }
          }
          r = b;
        }
      }
      int r = R.row0;
      while (r < R.row1) {
        if (taken[r]) { ++r; continue; }
        int b = r;
        while (b < R.row1 && !taken[b] && (b == r || b % band != 0)) ++b;
        seg[1].push_back(MarchSeg{st, r, b, 0});
        r = b;
      }
    }
  }
  hycom_tsadvc_handle::SegLists L;
  for (int c = 0; c < 2; ++c) {
    // neighbouring strips of the same rows next to each other: they share their aprons through L2
    std::stable_sort(seg[c].begin(), seg[c].end(), [&](const MarchSeg& a, const MarchSeg& b) {
      const int ba = (a.j0 + 12) / band, bb = (b.j0 + 12) / band;
      return ba != bb ? ba < bb : a.strip < b.strip;
    });
    L.n[c] = (long)seg[c].size();
    if (L.n[c]) {
      CU(h, cudaMalloc(&L.d[c], sizeof(MarchSeg) * seg[c].size()));
      CU(h, cudaMemcpyAsync(L.d[c], seg[c].data(), sizeof(MarchSeg) * seg[c].size(), cudaMemcpyHostToDevice,
                            h->stream));
      CU(h, cudaStreamSynchronize(h->stream));
    }
  }
  *out = &(h->seg_cache[key] = L);
  return 0;
}

// layers k0 .. k0+nk-1 (0-based; nk < 0: all)
int run_march(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params& p,
              const std::vector<Adv>& adv, int part, int k0 = 0, int nk = -1,
              const double* u_over = nullptr, const double* v_over = nullptr,
              const double* u_prolog = nullptr, const double* v_prolog = nullptr) {
  // the frame of a multi-tile step may run on its own stream (hycom_tsadvc_set_frame_stream), next
  // to the interior launch instead of behind it: it depends on the unpacked halos only
  cudaStream_t lst = (part == HYCOM_TSADVC_PART_FRAME && h->frame_stream) ? h->frame_stream : h->stream;
  const int kk = nk < 0 ? h->d.kdm : nk, aadv = abs(p.advtyp);
  int rc;
  MarchParams P;
  memset(&P, 0, sizeof P);
  P.nfld = (int)adv.size();
  P.kk = kk;
  const char* cn = getenv("HYCOM_TSADVC_NC");
  // cells per lane (tuning knobs HYCOM_TSADVC_NC / _MINB / _CHUNK_ROWS).  Defaults measured on B200
  // (profiles/r01p_variants.txt): FCT2 and MPDATA 2 cells per lane at 2 blocks per SM (11.8 warp-
  // instructions per useful cell against 15.4 with one cell per lane), FCT4 and PCM 1 cell per lane
  P.nc = (aadv == 0 || aadv == 4 || u_prolog) ? 1 : cn ? (atoi(cn) == 2 ? 2 : 1) : 2;
  const long koff = h->slab * k0;
  for (int f = 0; f < P.nfld; ++f) {
    double *in, *ctr, *out;
    if ((rc = slot(h, adv[f].field, adv[f].ktr, n, &in))) return rc;
    if ((rc = slot(h, adv[f].field, adv[f].ktr, m, &ctr))) return rc;
    if ((rc = spare_of(h, mirror_of(h, adv[f].field, adv[f].ktr), &out))) return rc;
    const long kf = koff + h->slab * adv[f].koff;
    P.fld[f].fld = in + kf;
    P.fld[f].fldc = ctr + kf;
    P.fld[f].out = out + kf;
    P.fld[f].posdef = adv[f].posdef;
    const int nl = adv[f].nlay - k0;
    P.fld[f].nlay = nl < 0 ? 0 : (nl > kk ? kk : nl);
  }
  double *u, *v, *dpn;
  if ((rc = slot(h, HYCOM_F_UFLX, 0, 1, &u))) return rc;
  if ((rc = slot(h, HYCOM_F_VFLX, 0, 1, &v))) return rc;
  if ((rc = slot(h, HYCOM_F_DP, 0, n, &dpn))) return rc;
  P.u = u + koff; P.v = v + koff; P.dp = dpn + koff;
  if (u_over) { P.u = u_over; P.v = v_over; }   // isopyc layer 1: the smoothed fluxes, prolog included (:1930-1932)
  // isopyc layer 1, tracers and q2, q2l: advected by uflx, vflx (:2016-2048) with the fco of the smoothed fluxes
  P.u2 = u_prolog; P.v2 = v_prolog;
  P.slab = h->slab;
  P.njobs = P.nfld * kk;
  P.g.pitch = h->pitch; P.g.ncols = h->ncols; P.g.nrows = h->nrows; P.g.nbdy = h->d.nbdy;
  P.g.ii = h->d.ii; P.g.jj = h->d.jj;
  P.g.mask = h->mask; P.g.scp2 = h->scp2; P.g.scp2i = h->scp2i;
  P.g.mask64 = h->static_block + 2 * h->slab;
  P.g.delt1 = p.delt1; P.g.onemm = p.onemm;
  const char* cb = getenv("HYCOM_TSADVC_MINB");
  P.minb = cb ? atoi(cb) : (P.nc == 2 ? 2 : 3);
  if (u_prolog) P.minb = 2;
  // rows one warp marches: long chunks amortise the 6 rows of pipeline fill (1024: 0.6 %), but the
  // launch needs about ten waves of blocks to hide the tail - small tiles get shorter chunks
  const char* ce = getenv("HYCOM_TSADVC_CHUNK_ROWS");
  int chunk_rows;
  if (ce) {
    chunk_rows = atoi(ce);
  } else {
    const long per_chunk = (long)P.njobs * strip_count(h->pitch, P.nc);          // warps per chunk row-band
    const long want = 10L * 148 * P.minb * kWarpsPerBlock;
    const long nch = (want + per_chunk - 1) / per_chunk;
    chunk_rows = (int)((h->nrows + nch - 1) / (nch > 0 ? nch : 1));
    if (chunk_rows > 1024) chunk_rows = 1024;
    if (chunk_rows < 128) chunk_rows = 128;
  }
  if (chunk_rows < 8) chunk_rows = 8;

  // (strip,row) rectangles of this part.  Interior = units whose staged window (apron and
  // prefetched row included) lies inside 1..ii x 1..jj, i.e. reads no halo cell.
  const int nstrips = strip_count(h->pitch, P.nc);
  const int nb = h->d.nbdy, use = strip_use(P.nc), lead = strip_lead(P.nc), wid = 32 * P.nc;
  int s_lo = 0, s_hi = nstrips;           // interior strips [s_lo, s_hi)
  while (s_lo < nstrips && s_lo * use - lead < nb) ++s_lo;
  while (s_hi > s_lo && (s_hi - 1) * use - lead + wid > nb + h->d.ii) --s_hi;
  int r_lo = nb + 3, r_hi = nb + h->d.jj - 4;  // interior rows [r_lo, r_hi)
  if (s_lo >= s_hi || r_lo >= r_hi) { s_lo = s_hi = 0; r_lo = r_hi = 0; }
  // the two side columns of the frame are only a few strips wide: short chunks give the
  // launch enough warps to fill the machine
  auto add = [&](int strip0, int ns, int row0, int row1, int crows) {
    if (ns <= 0 || row1 <= row0) return;
    MarchRect& R = P.rect[P.nrect++];
    R.strip0 = strip0; R.nstrips = ns; R.row0 = row0; R.row1 = row1;
    R.chunk_rows = crows;
    R.nchunks = (row1 - row0 + crows - 1) / crows;
    R.unit0 = P.nunits;
    P.nunits += (long)P.njobs * ns * R.nchunks;
  };
  const bool empty_interior = (s_lo >= s_hi);
  if (part == HYCOM_TSADVC_PART_ALL || (part == HYCOM_TSADVC_PART_FRAME && empty_interior)) {
    add(0, nstrips, 0, h->nrows, chunk_rows);
  } else if (part == HYCOM_TSADVC_PART_INTERIOR) {
    add(s_lo, s_hi - s_lo, r_lo, r_hi, chunk_rows);
  } else {
    const int side = chunk_rows < 48 ? chunk_rows : 48;
    add(0, s_lo, 0, h->nrows, side);
    add(s_hi, nstrips - s_hi, 0, h->nrows, side);
    add(s_lo, s_hi - s_lo, 0, r_lo, chunk_rows);
    add(s_lo, s_hi - s_lo, r_hi, h->nrows, chunk_rows);
  }
  if (P.nunits == 0) return 0;
  std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
  if (h->timing) {
    if (!h->ev_free.empty()) { ev = h->ev_free.back(); h->ev_free.pop_back(); }
    else { CU(h, cudaEventCreate(&ev.first)); CU(h, cudaEventCreate(&ev.second)); }
    CU(h, cudaEventRecord(ev.first, lst));
  }
  // FCT2, MPDATA: the all-sea row segments go to the mask-free instantiation, the rest to the general
  // one (HYCOM_TSADVC_SPLIT=0: one general launch over the regular chunks)
  const char* cs = getenv("HYCOM_TSADVC_SPLIT");
  if ((aadv == 2 || aadv == 1 || aadv == 4) && !p.btrmas && !u_prolog && !(cs && atoi(cs) == 0) && !h->mask_host.empty()) {
    const hycom_tsadvc_handle::SegLists* L = nullptr;
    // (whole-band all-sea pieces for FCT2 / FCT4, on tiles as well: 8 x B200 on one box 2.79 ms per step with them,
    // 2.84 with maximal runs - profiles/r03v; MPDATA keeps the maximal runs)
    if ((rc = march_segments(h, P, part, chunk_rows, aadv != 1, &L))) return rc;
    rc = 0;
    for (int c = 1; c >= 0 && !rc; --c) {       // the general segments first: the long launch hides their tail
      if (!L->n[c]) continue;
      MarchParams Q = P;
      Q.seg = (const MarchSeg*)L->d[c]; Q.nseg = L->n[c]; Q.allsea = (c == 0);
      Q.nunits = (long)Q.njobs * Q.nseg;
      rc = launch_march_tma(aadv, Q, lst);
      h->launches += 1;
    }
  } else {
    rc = launch_march_tma(aadv, P, lst);
    h->launches += 1;
  }
  if (h->timing) {
    CU(h, cudaEventRecord(ev.second, lst));
    h->ev_pending.push_back(ev);
  }
  if (lst != h->stream) {   // everything after the frame (time-level switch, diagnostics) waits for it
    if (!h->ev_frame) CU(h, cudaEventCreateWithFlags(&h->ev_frame, cudaEventDisableTiming));
    CU(h, cudaEventRecord(h->ev_frame, lst));
    CU(h, cudaStreamWaitEvent(h->stream, h->ev_frame, 0));
  }
  if (rc) return fail(h, HYCOM_TSADVC_ECUDA, "march kernel launch failed: %s",
                      rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "bad scheme");
  return 0;
}

// the marching launch(es) of one call: with isopycnic coordinates layer 1 runs on the laterally
// smoothed mass fluxes (:1859-1897, :1930-1932, :1992-2005), the layers below on uflx, vflx
int advect_march(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params& p,
                 const std::vector<Adv>& adv, int part) {
  if (!p.isopyc) return run_march(h, m, n, p, adv, part);
  if (part == HYCOM_TSADVC_PART_INTERIOR) return 0;   // the smoothing reads the halo: no overlap
  int rc;
  if (!h->isopyc_flux && (rc = dalloc_field(h, &h->isopyc_flux, 2 * (size_t)h->slab))) return rc;
  double *u, *v;
  if ((rc = slot(h, HYCOM_F_UFLX, 0, 1, &u))) return rc;
  if ((rc = slot(h, HYCOM_F_VFLX, 0, 1, &v))) return rc;
  const int mbdy = abs(p.advtyp) == 0 ? 2 : 5;
  rc = launch_isopyc_smooth(u, v, h->isopyc_flux, h->isopyc_flux + h->slab, h->mask, h->pitch, h->nrows,
                            h->d.nbdy, h->d.ii, h->d.jj, mbdy - 1, h->stream);
  h->launches += 1;
  if (rc) return fail(h, HYCOM_TSADVC_ECUDA, "isopyc smoothing launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  // layer 1: the thermodynamic fields on the smoothed fluxes (:1992-2005), every other field on uflx, vflx
  // with the prolog of the smoothed ones (:1930-1932, :2016-2048)
  std::vector<Adv> thermo, other;
  for (const Adv& a : adv)
    (a.field == HYCOM_F_TEMP || a.field == HYCOM_F_TH3D || a.field == HYCOM_F_SALN ? thermo : other).push_back(a);
  double *us = h->isopyc_flux, *vs = h->isopyc_flux + h->slab;
  if ((rc = run_march(h, m, n, p, thermo, HYCOM_TSADVC_PART_ALL, 0, 1, us, vs))) return rc;
  if (!other.empty() && (rc = run_march(h, m, n, p, other, HYCOM_TSADVC_PART_ALL, 0, 1, nullptr, nullptr, us, vs))) return rc;
  if (h->d.kdm > 1 && (rc = run_march(h, m, n, p, adv, HYCOM_TSADVC_PART_ALL, 1, h->d.kdm - 1))) return rc;
  return 0;
}

// ---- advem_fct2c (advtyp=2 with btrmas, mod_tsadvc.F90:96-97, 999-1368) -------------------
// Layer batches of whole-tile kernels.  The scheme exchanges hloc and fldlo after each of its
// five iterations (:1186-1187): on a single tile the halo kernels do that, on several tiles
// the caller does (hycom_tsadvc_fct2c_stage / _halo_pack / _halo_unpack).
int fct2c_batch_layers(const hycom_tsadvc_handle* h) {
  const char* ce = getenv("HYCOM_TSADVC_FCT2C_BATCH");
  int nb = ce ? atoi(ce) : 8;
  if (nb < 1) nb = 1;
  return nb > h->d.kdm ? h->d.kdm : nb;
}

int fct2c_params(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params& p,
                 const std::vector<Adv>& adv, int batch, Fct2cParams& P) {
  const int kk = h->d.kdm, nf = (int)adv.size();
  int rc;
  const int nb = fct2c_batch_layers(h);
  if (batch < 0 || batch * nb >= kk) return fail(h, HYCOM_TSADVC_EINVAL, "fct2c: bad layer batch %d", batch);
  const long nslab = (long)(6 + 5 * nf) * nb;
  if (!h->fct2c_block || h->fct2c_slabs < nslab || h->fct2c_nb != nb) {
    if ((rc = dalloc_field(h, &h->fct2c_block, (size_t)nslab * h->slab))) return rc;
    if ((rc = dalloc(h, (void**)&h->fct2c_lcalc, (size_t)nb * h->slab, true))) return rc;
    h->fct2c_slabs = nslab; h->fct2c_nb = nb;
  }
  memset(&P, 0, sizeof P);
  for (int f = 0; f < nf; ++f) {
    double *in, *ctr, *out;
    if ((rc = slot(h, adv[f].field, adv[f].ktr, n, &in))) return rc;
    if ((rc = slot(h, adv[f].field, adv[f].ktr, m, &ctr))) return rc;
    if ((rc = spare_of(h, mirror_of(h, adv[f].field, adv[f].ktr), &out))) return rc;
    const long kf = h->slab * adv[f].koff;
    P.fld[f] = in + kf; P.fldc[f] = ctr + kf; P.out[f] = out + kf; P.nlay[f] = adv[f].nlay;
  }
  double *u, *v, *dpn, *on;
  if ((rc = slot(h, HYCOM_F_UFLX, 0, 1, &u))) return rc;
  if ((rc = slot(h, HYCOM_F_VFLX, 0, 1, &v))) return rc;
  if ((rc = slot(h, HYCOM_F_DP, 0, n, &dpn))) return rc;
  if ((rc = slot(h, HYCOM_F_ONETA, 0, n, &on))) return rc;
  P.nf = nf; P.pitch = h->pitch; P.ncols = h->ncols; P.nrows = h->nrows; P.nbdy = h->d.nbdy;
  P.ii = h->d.ii; P.jj = h->d.jj; P.slab = h->slab;
  P.mask = h->mask; P.scp2 = h->scp2; P.scp2i = h->scp2i; P.oneta = on;
  P.u = u; P.v = v; P.dp = dpn; P.dt2 = p.delt1;
  P.k0 = batch * nb;
  P.nb = (P.k0 + nb <= kk) ? nb : kk - P.k0;
  // scratch: [hloc | dtloc | ucum | vcum | uloc | vloc] of P.nb slabs each, then
  // [fldlo | flx | fly | flxcum | flycum] of nf*P.nb slabs each (a shorter last batch packs tighter)
  const long S = h->slab, L = S * P.nb, F = L * nf;
  double* b = h->fct2c_block;
  P.hloc = b; P.dtloc = b + L; P.ucum = b + 2 * L; P.vcum = b + 3 * L; P.uloc = b + 4 * L; P.vloc = b + 5 * L;
  double* pf = b + 6 * L;
  P.fldlo = pf; P.flx = pf + F; P.fly = pf + 2 * F; P.flxcum = pf + 3 * F; P.flycum = pf + 4 * F;
  P.lcalc = h->fct2c_lcalc;
  return 0;
}

// stage 0: :1072-1086; stage 1: one iteration :1090-1184 (without its xctilr); stage 2: :1202-1361
int fct2c_stage(hycom_tsadvc_handle* h, const Fct2cParams& P, int stage) {
  int rc;
  auto launch = [&](int st) -> int {
    int r2 = launch_fct2c(st, P, h->stream);
    h->launches += 1;
    if (r2) return fail(h, HYCOM_TSADVC_ECUDA, "fct2c kernel %d launch failed: %s", st,
                        r2 > 0 ? cudaGetErrorString((cudaError_t)r2) : "bad stage");
    return 0;
  };
  const long L = P.slab * P.nb;
  if (stage == 0) {   // everything but hloc, fldlo, lcalc starts from 0.0
    CU(h, cudaMemsetAsync(P.dtloc, 0, sizeof(double) * L * 5, h->stream));
    CU(h, cudaMemsetAsync(P.flx, 0, sizeof(double) * L * P.nf * 4, h->stream));
    return launch(0);
  }
  if (stage == 1) {
    if ((rc = launch(1)) || (rc = launch(2)) || (rc = launch(3))) return rc;
    return 0;
  }
  if ((rc = launch(4)) || (rc = launch(5)) || (rc = launch(6))) return rc;
  return 0;
}

// hloc and fldlo of the batch as the arrays of one exchange, halo width 5 (:1186-1187)
void fct2c_halo_arrays(const hycom_tsadvc_handle* h, const Fct2cParams& P, HaloArrays& a) {
  memset(&a, 0, sizeof a);
  a.base[a.narr++] = P.hloc;
  for (int f = 0; f < P.nf; ++f) a.base[a.narr++] = P.fldlo + (long)f * P.nb * P.slab;
  a.kk = P.nb; a.slab = h->slab; a.pitch = h->pitch; a.nrows = h->nrows; a.nbdy = h->d.nbdy;
  a.ii = h->d.ii; a.jj = h->d.jj; a.mh = 5; a.nh = 5;
  for (int q = 0; q < a.narr; ++q) a.itype[q] = 1;   // halo_ps (:1186-1187)
  a.fold = arctic_fold(h->d);
}

// the whole scheme: on a single tile the halo kernels stand for the five xctilr(hloc), xctilr(fldlo)
// of :1186-1187, on several tiles the attached communicator moves them
int run_fct2c(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params& p,
              const std::vector<Adv>& adv) {
  const int kk = h->d.kdm, nb = fct2c_batch_layers(h);
  const int nreg = h->d.nreg;
  const int per_i = !(nreg == 0 || nreg == 4), per_j = nreg == 2 ? 101 : nreg > 2;   // halo_ps
  const bool single = h->d.ipr * h->d.jpr == 1;
  int rc;
  for (int batch = 0; batch * nb < kk; ++batch) {
    Fct2cParams P;
    if ((rc = fct2c_params(h, m, n, p, adv, batch, P))) return rc;
    if ((rc = fct2c_stage(h, P, 0))) return rc;
    for (int iter = 1; iter <= 5; ++iter) {   // :1088
      if ((rc = fct2c_stage(h, P, 1))) return rc;
      if (!single) {
        HaloArrays a;
        fct2c_halo_arrays(h, P, a);
        if ((rc = xc_exchange(h, a, false, h->stream))) return rc;
        continue;
      }
      // xctilr(hloc), xctilr(fldlo): hloc and the fldlo stack are contiguous slabs
      for (int q = 0; q < 2; ++q) {
        int r2 = launch_halo_local(q ? P.fldlo : P.hloc, h->slab, q ? P.nb * P.nf : P.nb, h->pitch, h->d.nbdy,
                                   h->d.ii, h->d.jj, 5, 5, per_i, per_j, h->stream);
        h->launches += 2;
        if (r2) return fail(h, HYCOM_TSADVC_ECUDA, "halo kernel launch failed: %s", cudaGetErrorString((cudaError_t)r2));
      }
    }
    if ((rc = fct2c_stage(h, P, 2))) return rc;
  }
  return 0;
}

// slot n moves to the ping-pong buffer; salinity range diagnostics (:2065-2094)
int finish_step(hycom_tsadvc_handle* h, int32_t n, const hycom_tsadvc_params& p,
                const std::vector<Adv>& adv, double* xmin, double* xmax) {
  const int kk = h->d.kdm;
  int rc;
  for (const Adv& f : adv) {
    Mirror* mi = mirror_of(h, f.field, f.ktr);
    // layers that were not advected keep their values (:2008-2014; q2,q2l: layers 0 and kk+1)
    const int nsl = nlayers_of(h, f.field), hi = f.koff + f.nlay;
    if (f.koff > 0)
      CU(h, cudaMemcpyAsync(mi->spare, mi->lev[n - 1], sizeof(double) * (size_t)h->slab * f.koff,
                            cudaMemcpyDeviceToDevice, h->stream));
    if (hi < nsl)
      CU(h, cudaMemcpyAsync(mi->spare + h->slab * hi, mi->lev[n - 1] + h->slab * hi,
                            sizeof(double) * (size_t)h->slab * (nsl - hi),
                            cudaMemcpyDeviceToDevice, h->stream));
    double* t = mi->lev[n - 1];
    mi->lev[n - 1] = mi->spare;
    mi->spare = t;
  }
  const bool want = h->deferred_range || (xmin && xmax);
  if (want && ((p.nstep % 3 == 0) || p.diagno)) {
    double* dpn;
    if ((rc = slot(h, HYCOM_F_DP, 0, n, &dpn))) return rc;
    if (!h->d_minmax && (rc = dalloc(h, (void**)&h->d_minmax, sizeof(double) * 2 * kk, false))) return rc;
    k_minmax_init<<<(kk + 127) / 128, 128, 0, h->stream>>>(h->d_minmax, kk);
    dim3 grid(148, kk);
    k_saln_minmax<<<grid, 256, 0, h->stream>>>(h->saln.lev[n - 1], dpn, h->mask, h->slab, p.onemm,
                                               h->d_minmax, kk);
    h->launches += 2;
    // :2093-2094 xcminr, xcmaxr (a host-owned transport reduces the tile values itself)
    if (h->xc && (rc = xc_minmax(h, h->d_minmax, kk, h->stream))) return rc;
    if (!h->range_host) {
      CU(h, cudaMallocHost((void**)&h->range_host, sizeof(double) * 2 * kk));
      CU(h, cudaEventCreateWithFlags(&h->ev_range, cudaEventDisableTiming));
    }
    CU(h, cudaMemcpyAsync(h->range_host, h->d_minmax, sizeof(double) * 2 * kk, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaEventRecord(h->ev_range, h->stream));
    h->range_nstep = p.nstep;
    if (!h->deferred_range) {
      CU(h, cudaEventSynchronize(h->ev_range));
      for (int k = 0; k < kk; ++k) { xmin[k] = h->range_host[k]; xmax[k] = h->range_host[kk + k]; }
      h->range_nstep = -1;
    }
  }
  CU(h, cudaGetLastError());
  return 0;
}


inline int ffield_of(bool adv_th3d) { return adv_th3d ? HYCOM_F_TH3D : HYCOM_F_TEMP; }

// the arrays of the second exchange (mod_tsadvc.F90:2140-2151): saln, temp, th3d, tracers of
// slot n, halo width mdf = 2
int diff_halo_arrays(hycom_tsadvc_handle* h, int n, bool mxlmy, HaloArrays& a) {
  memset(&a, 0, sizeof a);
  int rc;
  if ((rc = slot(h, HYCOM_F_SALN, 0, n, &a.base[a.narr++]))) return rc;
  if ((rc = slot(h, HYCOM_F_TEMP, 0, n, &a.base[a.narr++]))) return rc;
  if ((rc = slot(h, HYCOM_F_TH3D, 0, n, &a.base[a.narr++]))) return rc;
  for (int t = 1; t <= h->d.ntracr; ++t)
    if ((rc = slot(h, HYCOM_F_TRACER, t, n, &a.base[a.narr++]))) return rc;
  if (mxlmy)   // :2143-2146 (the reference sends layers 0..kk+1; 1..kk are the ones tsdff reads)
    for (int f : {HYCOM_F_Q2, HYCOM_F_Q2L}) {
      if ((rc = slot(h, f, 0, n, &a.base[a.narr]))) return rc;
      a.base[a.narr++] += h->slab;
    }
  a.kk = h->d.kdm; a.slab = h->slab; a.pitch = h->pitch; a.nrows = h->nrows; a.nbdy = h->d.nbdy;
  a.ii = h->d.ii; a.jj = h->d.jj; a.mh = 2; a.nh = 2;
  for (int q = 0; q < a.narr; ++q) a.itype[q] = 1;   // halo_ps (:2140-2151)
  a.fold = arctic_fold(h->d);
  return 0;
}

// :2153-2229 on the device mirrors: one launch for T/S/th3d + equation of state, one for the
// tracers; slot n moves to the ping-pong buffers
int run_diffuse(hycom_tsadvc_handle* h, int32_t n, const hycom_tsadvc_params& p) {
  const int kk = h->d.kdm;
  int rc;
  double *dpn, *on;
  if ((rc = slot(h, HYCOM_F_DP, 0, n, &dpn))) return rc;
  if ((rc = slot(h, HYCOM_F_ONETA, 0, n, &on))) return rc;
  const int nhyb = p.nhybrd < 0 ? 0 : (p.nhybrd > kk ? kk : p.nhybrd);
  const bool need_theta = nhyb < kk && !(kk == 1 && p.isopyc);
  if (need_theta && !h->theta.lev[0])
    return fail(h, HYCOM_TSADVC_EINVAL, "temdf2>0 with nhybrd<kdm reads theta: upload HYCOM_F_THETA first");
  // au = temdf2*aspux*scuy, av = temdf2*aspvy*scvx: the k-independent part of the face factors
  if (!h->diff_static && (rc = dalloc_field(h, &h->diff_static, 2 * (size_t)h->slab))) return rc;
  rc = launch_diff_static(h->aspux, h->scuy, h->aspvy, h->scvx, p.temdf2, h->diff_static,
                          h->diff_static + h->slab, h->slab, h->stream);
  h->launches += 1;
  if (rc) return fail(h, HYCOM_TSADVC_ECUDA, "tsdff kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  DiffMarchParams M;
  memset(&M, 0, sizeof M);
  M.dp = dpn; M.oneta = on; M.au = h->diff_static; M.av = h->diff_static + h->slab;
  M.scp2 = h->scp2; M.mask64 = h->static_block + 2 * h->slab; M.theta = h->theta.lev[0];
  M.slab = h->slab; M.pitch = h->pitch; M.nrows = h->nrows; M.kk = kk;
  M.nstrips = (h->pitch + 1 + 29) / 30;
  // about ten waves of 4-warp blocks at 4 blocks per SM
  const long per_chunk = (long)kk * M.nstrips, want = 10L * 148 * 4 * 4;
  const long nch = (want + per_chunk - 1) / per_chunk;
  M.chunk_rows = (int)((h->nrows + nch - 1) / (nch > 0 ? nch : 1));
  if (M.chunk_rows > 1024) M.chunk_rows = 1024;
  if (M.chunk_rows < 64) M.chunk_rows = 64;
  M.nchunks = (h->nrows + M.chunk_rows - 1) / M.chunk_rows;
  M.nunits = (long)kk * M.nstrips * M.nchunks;
  M.nhybrd = nhyb; M.isopyc = p.isopyc;
  eos::fill(p.sigver, M.eosc);
  M.temdfc = p.temdfc; M.thbase = p.thbase; M.delt1 = p.delt1;
  struct Fld { int field, ktr; };
  // slot n of a field (from model layer 1) and its ping-pong buffer
  auto ptrs = [&](const Fld& f, const double** in, double** out) -> int {
    double *i0, *o0;
    int r2 = slot(h, f.field, f.ktr, n, &i0);
    if (!r2) r2 = spare_of(h, mirror_of(h, f.field, f.ktr), &o0);
    if (r2) return r2;
    const long kf = h->slab * layer1_of(f.field);
    *in = i0 + kf; *out = o0 + kf;
    return 0;
  };
  auto swap = [&](const Fld& f) -> int {
    Mirror* mi = mirror_of(h, f.field, f.ktr);
    if (layer1_of(f.field)) {   // q2,q2l: layers 0 and kk+1 travel with the buffer
      const long last = h->slab * (kk + 1);
      CU(h, cudaMemcpyAsync(mi->spare, mi->lev[n - 1], sizeof(double) * h->slab, cudaMemcpyDeviceToDevice, h->stream));
      CU(h, cudaMemcpyAsync(mi->spare + last, mi->lev[n - 1] + last, sizeof(double) * h->slab,
                            cudaMemcpyDeviceToDevice, h->stream));
    }
    double* t = mi->lev[n - 1]; mi->lev[n - 1] = mi->spare; mi->spare = t;
    return 0;
  };
  // one launch group: fields sharing the face factors
  auto group = [&](const std::vector<Fld>& g, int eos) -> int {
    M.eos = eos; M.nf = (int)g.size();
    for (int q = 0; q < M.nf; ++q)
      if ((rc = ptrs(g[q], &M.in[q], &M.out[q]))) return rc;
    int r2 = launch_tsdff_march(M, h->stream);
    h->launches += 1;
    if (r2) return fail(h, HYCOM_TSADVC_ECUDA, "tsdff kernel launch failed: %s",
                        r2 > 0 ? cudaGetErrorString((cudaError_t)r2) : "bad field group");
    for (const Fld& f : g)
      if ((rc = swap(f))) return rc;
    return 0;
  };
  // :2166-2185 + :2199-2229  temp, saln, th3d with the equation of state
  if ((rc = group({{HYCOM_F_TEMP, 0}, {HYCOM_F_SALN, 0}, {HYCOM_F_TH3D, 0}}, 1))) return rc;
  // :2180-2183 q2 & q2l; :2190-2198 tracers in tsdff_2x pairs, a last single one with tsdff_1x
  if (p.mxlmy && (rc = group({{HYCOM_F_Q2, 0}, {HYCOM_F_Q2L, 0}}, 0))) return rc;
  for (int t = 1; t <= h->d.ntracr; t += 2) {
    if (t + 1 <= h->d.ntracr) rc = group({{HYCOM_F_TRACER, t}, {HYCOM_F_TRACER, t + 1}}, 0);
    else rc = group({{HYCOM_F_TRACER, t}}, 0);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace

extern "C" {

int hycom_tsadvc_step_device_part(hycom_tsadvc_handle* h, int32_t m, int32_t n,
                                  const hycom_tsadvc_params* prm, int32_t part, double* xmin,
                                  double* xmax) {
  std::vector<Adv> adv;
  int mbdy = 0, rc;
  if ((rc = plan_step(h, m, n, prm, adv, mbdy))) return rc;
  if (part < HYCOM_TSADVC_PART_ALL || part > HYCOM_TSADVC_PART_FRAME)
    return fail(h, HYCOM_TSADVC_EINVAL, "step: bad part %d", part);
  CU(h, cudaSetDevice(h->d.device));
  // :1827-1836  xctilr of the advected fields (both slots) and the mass fluxes; "dp halo is
  // up to date".  Single tile: done here.  Multi-tile: hycom_tsadvc_halo_pack/unpack around
  // the caller's transport, before PART_ALL / PART_FRAME.
  if (h->d.ipr * h->d.jpr == 1 && part != HYCOM_TSADVC_PART_FRAME) {
    for (const Adv& a : adv)
      if ((rc = hycom_tsadvc_halo_local(h, a.field, a.ktr, 0, mbdy, mbdy))) return rc;
    if (prm->advflg == 0 && h->th3d.lev[0] && h->th3d.lev[1])
      if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_TH3D, 0, 0, mbdy, mbdy))) return rc;
    if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_UFLX, 0, 1, mbdy, mbdy))) return rc;
    if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_VFLX, 0, 1, mbdy, mbdy))) return rc;
  }
  if (prm->btrmas && abs(prm->advtyp) == 2) {   // advem_fct2c (:96-97)
    if (part == HYCOM_TSADVC_PART_INTERIOR) return 0;   // no overlap: the scheme exchanges five times itself
    // several tiles: the caller has driven hycom_tsadvc_fct2c_stage + exchanges; this call finishes the step
    if (h->d.ipr * h->d.jpr == 1 && (rc = run_fct2c(h, m, n, *prm, adv))) return rc;
  } else if ((rc = advect_march(h, m, n, *prm, adv, part))) {
    return rc;
  }
  if (part == HYCOM_TSADVC_PART_INTERIOR) return 0;
  if ((rc = finish_step(h, n, *prm, adv, xmin, xmax))) return rc;
  // :2138-2230; on a multi-tile handle the caller exchanges first (hycom_tsadvc_diff_halo_*)
  if (prm->temdf2 > 0.0 && h->d.ipr * h->d.jpr == 1) return hycom_tsadvc_diffuse_device(h, m, n, prm);
  return 0;
}

int hycom_tsadvc_diffuse_device(hycom_tsadvc_handle* h, int32_t m, int32_t n,
                                const hycom_tsadvc_params* prm) {
  std::vector<Adv> adv;
  int mbdy = 0, rc;
  if ((rc = plan_step(h, m, n, prm, adv, mbdy))) return rc;
  if (!(prm->temdf2 > 0.0)) return 0;
  CU(h, cudaSetDevice(h->d.device));
  if (h->d.ipr * h->d.jpr == 1) {   // :2140-2151, mdf = 2
    const int mdf = 2;
    if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_SALN, 0, n, mdf, mdf))) return rc;
    if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_TEMP, 0, n, mdf, mdf))) return rc;
    if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_TH3D, 0, n, mdf, mdf))) return rc;
    for (int t = 1; t <= h->d.ntracr; ++t)
      if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_TRACER, t, n, mdf, mdf))) return rc;
    if (prm->mxlmy) {
      if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_Q2, 0, n, mdf, mdf))) return rc;
      if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_Q2L, 0, n, mdf, mdf))) return rc;
    }
  } else if (h->xc) {   // the same exchange through the communicator (a host-owned transport has done it)
    HaloArrays a;
    if ((rc = diff_halo_arrays(h, n, prm->mxlmy != 0, a))) return rc;
    if ((rc = xc_exchange(h, a, false, h->stream))) return rc;
  }
  return run_diffuse(h, n, *prm);
}

int hycom_tsadvc_fct2c_batches(hycom_tsadvc_handle* h, int32_t* nbatch, int32_t* layers_per_batch) {
  if (!h || !nbatch) return fail(h, HYCOM_TSADVC_EINVAL, "fct2c_batches: null argument");
  const int nb = fct2c_batch_layers(h);
  *nbatch = (h->d.kdm + nb - 1) / nb;
  if (layers_per_batch) *layers_per_batch = nb;
  return 0;
}

int hycom_tsadvc_fct2c_stage(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params* prm,
                             int32_t batch, int32_t stage) {
  std::vector<Adv> adv;
  int mbdy = 0, rc;
  if ((rc = plan_step(h, m, n, prm, adv, mbdy))) return rc;
  if (!(prm->btrmas && abs(prm->advtyp) == 2)) return fail(h, HYCOM_TSADVC_EINVAL, "fct2c_stage: needs advtyp=2 and btrmas");
  if (stage < 0 || stage > 2) return fail(h, HYCOM_TSADVC_EINVAL, "fct2c_stage: bad stage %d", stage);
  CU(h, cudaSetDevice(h->d.device));
  Fct2cParams P;
  if ((rc = fct2c_params(h, m, n, *prm, adv, batch, P))) return rc;
  return fct2c_stage(h, P, stage);
}

static int fct2c_halo_xfer(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params* prm,
                           int32_t batch, double* const buf[8], int64_t* count, void* cuda_stream, int mode) {
  std::vector<Adv> adv;
  int mbdy = 0, rc;
  if ((rc = plan_step(h, m, n, prm, adv, mbdy))) return rc;
  CU(h, cudaSetDevice(h->d.device));
  Fct2cParams P;
  if ((rc = fct2c_params(h, m, n, *prm, adv, batch, P))) return rc;
  HaloArrays a;
  fct2c_halo_arrays(h, P, a);
  if (mode == 0) {
    for (int d = 0; d < 8; ++d) {
      count[d] = (int64_t)halo_count(a, d);
    }
    return 0;
  }
  if (!buf) return fail(h, HYCOM_TSADVC_EINVAL, "fct2c halo: null buffer table");
  HaloBufs b;
  for (int d = 0; d < 8; ++d) { b.buf[d] = buf[d]; b.count[d] = 0; }
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  rc = mode == 1 ? launch_halo_pack(a, b, st) : launch_halo_unpack(a, b, st);
  h->launches += 1;
  if (rc) return fail(h, HYCOM_TSADVC_ECUDA, "halo kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  return 0;
}

int hycom_tsadvc_fct2c_halo_counts(hycom_tsadvc_handle* h, int32_t m, int32_t n,
                                   const hycom_tsadvc_params* prm, int32_t batch, int64_t count[8]) {
  if (!count) return fail(h, HYCOM_TSADVC_EINVAL, "fct2c_halo_counts: null argument");
  return fct2c_halo_xfer(h, m, n, prm, batch, nullptr, count, nullptr, 0);
}
int hycom_tsadvc_fct2c_halo_pack(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params* prm,
                                 int32_t batch, double* const sendbuf[8], void* cuda_stream) {
  return fct2c_halo_xfer(h, m, n, prm, batch, sendbuf, nullptr, cuda_stream, 1);
}
int hycom_tsadvc_fct2c_halo_unpack(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params* prm,
                                   int32_t batch, double* const recvbuf[8], void* cuda_stream) {
  return fct2c_halo_xfer(h, m, n, prm, batch, recvbuf, nullptr, cuda_stream, 2);
}

// ---- mod_asselin.F90 on the device mirrors -------------------------------------------------
static int asselin_params(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params* prm,
                          double ra2fac, double oneta0, AsselinParams& A) {
  if (!h || !prm) return fail(h, HYCOM_TSADVC_EINVAL, "asselin: null argument");
  if (!h->have_static) return fail(h, HYCOM_TSADVC_EINVAL, "asselin: set_static has not been called");
  if (m < 1 || m > 2 || n < 1 || n > 2 || m == n)
    return fail(h, HYCOM_TSADVC_EINVAL, "asselin: bad leapfrog slots m=%d n=%d", m, n);
  if (!(oneta0 > 0.0 && oneta0 < 1.0))   // blkdat.F90:436-444
    return fail(h, HYCOM_TSADVC_EINVAL, "error - oneta0 must be above 0.0 and below 1.0");
  if (prm->sigver < 1 || prm->sigver > 8)
    return fail(h, HYCOM_TSADVC_EINVAL, "asselin: sigver=%d not in 1..8", prm->sigver);
  CU(h, cudaSetDevice(h->d.device));
  memset(&A, 0, sizeof A);
  const int kk = h->d.kdm;
  int rc;
  double *pb, *pbot;
  if ((rc = slot(h, HYCOM_F_PBAVG, 0, 1, &pb))) return rc;
  if ((rc = slot(h, HYCOM_F_PBOT, 0, 1, &pbot))) return rc;
  A.pbavg_n = pb + h->slab * (n - 1); A.pbavg_m = pb + h->slab * (m - 1); A.pbot = pbot;
  if ((rc = slot(h, HYCOM_F_ONETA, 0, n, &A.oneta_n)) || (rc = slot(h, HYCOM_F_ONETA, 0, m, &A.oneta_m))) return rc;
  if ((rc = slot(h, HYCOM_F_ONETAO, 0, n, &A.onetao_n)) || (rc = slot(h, HYCOM_F_ONETAO, 0, m, &A.onetao_m))) return rc;
  A.pitch = h->pitch; A.nrows = h->nrows; A.nbdy = h->d.nbdy; A.ii = h->d.ii; A.jj = h->d.jj; A.kk = kk;
  A.slab = h->slab; A.mask = h->mask;
  A.nhybrd = prm->nhybrd < 0 ? 0 : (prm->nhybrd > kk ? kk : prm->nhybrd);
  A.advflg = prm->advflg; A.isopyc = prm->isopyc;
  eos::fill(prm->sigver, A.eosc);
  A.ra2fac = ra2fac; A.oneta0 = oneta0; A.thbase = prm->thbase;
  return 0;
}

int hycom_tsadvc_asselin_save_device(hycom_tsadvc_handle* h, int32_t m, int32_t n,
                                     const hycom_tsadvc_params* prm, double oneta0) {
  AsselinParams A;
  int rc;
  if ((rc = asselin_params(h, m, n, prm, 0.0, oneta0, A))) return rc;
  auto launch = [&](int stage) -> int {
    int r2 = launch_asselin(stage, A, h->stream);
    h->launches += 1;
    if (r2) return fail(h, HYCOM_TSADVC_ECUDA, "asselin kernel launch failed: %s",
                        r2 > 0 ? cudaGetErrorString((cudaError_t)r2) : "bad stage");
    return 0;
  };
  if ((rc = launch(1))) return rc;   // :52-55
  // :57-75  time level t-1 of every scalar
  auto cp = [&](int ofield, int field, int ktr, long off) -> int {
    double *o, *f;
    int r2 = slot(h, ofield, ktr, 1, &o);
    if (!r2) r2 = slot(h, field, ktr, n, &f);
    if (r2) return r2;
    A.cp[A.nf].o = o + off; A.cp[A.nf].fn = f + off; ++A.nf;
    return 0;
  };
  A.nf = 0; A.kcopy = h->d.kdm;
  if ((rc = cp(HYCOM_F_OTEMP, HYCOM_F_TEMP, 0, 0)) || (rc = cp(HYCOM_F_OSALN, HYCOM_F_SALN, 0, 0)) ||
      (rc = cp(HYCOM_F_OTH3D, HYCOM_F_TH3D, 0, 0))) return rc;
  for (int t = 1; t <= h->d.ntracr; ++t)
    if ((rc = cp(HYCOM_F_OTRACER, HYCOM_F_TRACER, t, 0))) return rc;
  if (prm->mxlmy)   // oq2(i,j,k) = q2(i,j,k,n), k = 1..kk of the (0:kk+1) arrays (:67-74)
    if ((rc = cp(HYCOM_F_OQ2, HYCOM_F_Q2, 0, h->slab)) || (rc = cp(HYCOM_F_OQ2L, HYCOM_F_Q2L, 0, h->slab))) return rc;
  if ((rc = launch(2))) return rc;
  if (h->d.ipr * h->d.jpr == 1) {   // :77-78 xctilr(oneta|onetao, 1,2, 6,6, halo_ps)
    if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_ONETA, 0, 0, 6, 6))) return rc;
    if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_ONETAO, 0, 0, 6, 6))) return rc;
  } else if (h->xc) {
    HaloArrays a;
    memset(&a, 0, sizeof a);
    for (int f : {HYCOM_F_ONETA, HYCOM_F_ONETAO})
      for (int t = 1; t <= 2; ++t) {
        if ((rc = slot(h, f, 0, t, &a.base[a.narr]))) return rc;
        a.itype[a.narr++] = 1;
      }
    a.kk = 1; a.slab = h->slab; a.pitch = h->pitch; a.nrows = h->nrows; a.nbdy = h->d.nbdy;
    a.ii = h->d.ii; a.jj = h->d.jj; a.mh = std::min(6, h->d.nbdy); a.nh = a.mh;
    a.fold = arctic_fold(h->d);
    if ((rc = xc_exchange(h, a, false, h->stream))) return rc;
  }
  return 0;
}

int hycom_tsadvc_asselin_filter_device(hycom_tsadvc_handle* h, int32_t m, int32_t n,
                                       const hycom_tsadvc_params* prm, double ra2fac, double oneta0) {
  AsselinParams A;
  int rc;
  if ((rc = asselin_params(h, m, n, prm, ra2fac, oneta0, A))) return rc;
  const int kk = h->d.kdm;
  double *dpo_n, *dpo_m, *dp_n, *dp_m;
  if ((rc = slot(h, HYCOM_F_DPO, 0, n, &dpo_n)) || (rc = slot(h, HYCOM_F_DPO, 0, m, &dpo_m)) ||
      (rc = slot(h, HYCOM_F_DP, 0, n, &dp_n)) || (rc = slot(h, HYCOM_F_DP, 0, m, &dp_m))) return rc;
  A.dpo_n = dpo_n; A.dpo_m = dpo_m; A.dp_n = dp_n; A.dp_m = dp_m;
  const bool need_theta = A.nhybrd < kk && !(kk == 1 && prm->isopyc);
  if (need_theta && !h->theta.lev[0])
    return fail(h, HYCOM_TSADVC_EINVAL, "asselin_filter with nhybrd<kdm reads theta: upload HYCOM_F_THETA first");
  A.theta = h->theta.lev[0];
  auto fld = [&](int ofield, int field, int ktr) -> int {
    double *o, *fm, *fn;
    int r2 = slot(h, ofield, ktr, 1, &o);
    if (!r2) r2 = slot(h, field, ktr, m, &fm);
    if (!r2) r2 = slot(h, field, ktr, n, &fn);
    if (r2) return r2;
    A.f[A.nf].o = o; A.f[A.nf].fm = fm; A.f[A.nf].fn = fn; ++A.nf;
    return 0;
  };
  A.nf = 0;
  if ((rc = fld(HYCOM_F_OSALN, HYCOM_F_SALN, 0)) || (rc = fld(HYCOM_F_OTEMP, HYCOM_F_TEMP, 0)) ||
      (rc = fld(HYCOM_F_OTH3D, HYCOM_F_TH3D, 0))) return rc;
  for (int t = 1; t <= h->d.ntracr; ++t)
    if ((rc = fld(HYCOM_F_OTRACER, HYCOM_F_TRACER, t))) return rc;
  if (prm->mxlmy) {
    double *a, *b;
    if ((rc = slot(h, HYCOM_F_OQ2, 0, 1, &a)) || (rc = slot(h, HYCOM_F_OQ2L, 0, 1, &b))) return rc;
    A.q2_o = a; A.q2l_o = b;
    if ((rc = slot(h, HYCOM_F_Q2, 0, m, &A.q2_m)) || (rc = slot(h, HYCOM_F_Q2L, 0, m, &A.q2l_m))) return rc;
    if ((rc = slot(h, HYCOM_F_Q2, 0, n, &a)) || (rc = slot(h, HYCOM_F_Q2L, 0, n, &b))) return rc;
    A.q2_n = a; A.q2l_n = b;
  }
  for (int stage : {0, 3}) {   // :115-118 oneta, then the filter
    rc = launch_asselin(stage, A, h->stream);
    h->launches += 1;
    if (rc) return fail(h, HYCOM_TSADVC_ECUDA, "asselin kernel launch failed: %s",
                        rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "bad stage");
  }
  return 0;
}

int hycom_tsadvc_diff_halo_counts(hycom_tsadvc_handle* h, int32_t n, const hycom_tsadvc_params* prm,
                                  int64_t count[8]) {
  if (!h || !prm || !count || n < 1 || n > 2) return fail(h, HYCOM_TSADVC_EINVAL, "diff_halo_counts: bad argument");
  HaloArrays a;
  int rc;
  if ((rc = diff_halo_arrays(h, n, prm->mxlmy != 0, a))) return rc;
  for (int d = 0; d < 8; ++d) {
    count[d] = (int64_t)halo_count(a, d);
  }
  return 0;
}

static int diff_halo_xfer(hycom_tsadvc_handle* h, int32_t n, const hycom_tsadvc_params* prm,
                          double* const buf[8], void* cuda_stream, bool pack) {
  if (!h || !prm || !buf || n < 1 || n > 2) return fail(h, HYCOM_TSADVC_EINVAL, "diff_halo: bad argument");
  CU(h, cudaSetDevice(h->d.device));
  HaloArrays a;
  int rc;
  if ((rc = diff_halo_arrays(h, n, prm->mxlmy != 0, a))) return rc;
  HaloBufs b;
  for (int d = 0; d < 8; ++d) { b.buf[d] = buf[d]; b.count[d] = 0; }
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  rc = pack ? launch_halo_pack(a, b, st) : launch_halo_unpack(a, b, st);
  h->launches += 1;
  if (rc) return fail(h, HYCOM_TSADVC_ECUDA, "halo kernel launch failed: %s",
                      cudaGetErrorString((cudaError_t)rc));
  return 0;
}

int hycom_tsadvc_diff_halo_pack(hycom_tsadvc_handle* h, int32_t n, const hycom_tsadvc_params* prm,
                                double* const sendbuf[8], void* cuda_stream) {
  return diff_halo_xfer(h, n, prm, sendbuf, cuda_stream, true);
}

int hycom_tsadvc_diff_halo_unpack(hycom_tsadvc_handle* h, int32_t n, const hycom_tsadvc_params* prm,
                                  double* const recvbuf[8], void* cuda_stream) {
  return diff_halo_xfer(h, n, prm, recvbuf, cuda_stream, false);
}

// tsadvc(m,n) on the device mirrors of an ipr x jpr tile: the exchange of :1829-1836 through the attached
// communicator, overlapped with the march over the tile interior,
//     exchange stream : pack -> send/recv -> unpack -> march(frame)
//     handle stream   : march(interior) ............................ join -> time-level switch, ...
// (the frame depends on the unpacked halos only, so it runs next to the interior launch and fills its
// tail).  advem_fct2c and isopyc read halos from their first kernel on: exchange first, no overlap.
static int step_device_tiles(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params* prm,
                             double* xmin, double* xmax) {
  std::vector<Adv> adv;
  int mbdy = 0, rc;
  if ((rc = plan_step(h, m, n, prm, adv, mbdy))) return rc;
  if (!h->xc)
    return fail(h, HYCOM_TSADVC_EUNSUPPORTED,
                "tsadvc on %d x %d tiles needs a communicator (hycom_tsadvc_comm_init) or a host-driven "
                "exchange (hycom_tsadvc_halo_pack/unpack + hycom_tsadvc_step_device_part)", h->d.ipr, h->d.jpr);
  CU(h, cudaSetDevice(h->d.device));
  HaloArrays a;
  if ((rc = halo_arrays(h, adv, mbdy, a))) return rc;
  const bool fct2c = prm->btrmas && abs(prm->advtyp) == 2;
  if (fct2c || prm->isopyc || !h->overlap) {
    if ((rc = xc_exchange(h, a, true, h->stream))) return rc;
    if (fct2c) rc = run_fct2c(h, m, n, *prm, adv);
    else rc = advect_march(h, m, n, *prm, adv, HYCOM_TSADVC_PART_ALL);
    if (rc) return rc;
  } else {
    cudaStream_t cs = xc_stream(h);
    if (!h->ev_xc) CU(h, cudaEventCreateWithFlags(&h->ev_xc, cudaEventDisableTiming));
    CU(h, cudaEventRecord(h->ev_xc, h->stream));     // the strips are packed from the finished previous step
    CU(h, cudaStreamWaitEvent(cs, h->ev_xc, 0));
    if ((rc = xc_exchange(h, a, true, cs))) return rc;
    if ((rc = run_march(h, m, n, *prm, adv, HYCOM_TSADVC_PART_INTERIOR))) return rc;
    cudaStream_t keep = h->frame_stream;
    h->frame_stream = cs;
    rc = run_march(h, m, n, *prm, adv, HYCOM_TSADVC_PART_FRAME);   // the handle's stream joins behind it
    h->frame_stream = keep;
    if (rc) return rc;
  }
  if ((rc = finish_step(h, n, *prm, adv, xmin, xmax))) return rc;
  if (prm->temdf2 > 0.0) return hycom_tsadvc_diffuse_device(h, m, n, prm);   // :2138-2230
  return 0;
}

int hycom_tsadvc_step_device(hycom_tsadvc_handle* h, int32_t m, int32_t n,
                             const hycom_tsadvc_params* prm, double* xmin, double* xmax) {
  if (h && h->d.ipr * h->d.jpr > 1) return step_device_tiles(h, m, n, prm, xmin, xmax);
  return hycom_tsadvc_step_device_part(h, m, n, prm, HYCOM_TSADVC_PART_ALL, xmin, xmax);
}

int hycom_tsadvc_halo_neighbors(const hycom_tsadvc_handle* h, int32_t nbr[8]) {
  if (!h || !nbr) return fail(nullptr, HYCOM_TSADVC_EINVAL, "halo_neighbors: null argument");
  for (int d = 0; d < 8; ++d) nbr[d] = neighbour(h->d, d);
  return 0;
}

int hycom_tsadvc_halo_counts(hycom_tsadvc_handle* h, int32_t m, int32_t n,
                             const hycom_tsadvc_params* prm, int64_t count[8]) {
  std::vector<Adv> adv;
  int mbdy = 0, rc;
  if (!count) return fail(h, HYCOM_TSADVC_EINVAL, "halo_counts: null argument");
  if ((rc = plan_step(h, m, n, prm, adv, mbdy))) return rc;
  HaloArrays a;
  if ((rc = halo_arrays(h, adv, mbdy, a))) return rc;
  for (int d = 0; d < 8; ++d) {
    count[d] = (int64_t)halo_count(a, d);
  }
  return 0;
}

static int halo_xfer(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params* prm,
                     double* const buf[8], void* cuda_stream, bool pack) {
  std::vector<Adv> adv;
  int mbdy = 0, rc;
  if (!buf) return fail(h, HYCOM_TSADVC_EINVAL, "halo: null buffer table");
  if ((rc = plan_step(h, m, n, prm, adv, mbdy))) return rc;
  CU(h, cudaSetDevice(h->d.device));
  HaloArrays a;
  if ((rc = halo_arrays(h, adv, mbdy, a))) return rc;
  HaloBufs b;
  for (int d = 0; d < 8; ++d) { b.buf[d] = buf[d]; b.count[d] = 0; }
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  rc = pack ? launch_halo_pack(a, b, st) : launch_halo_unpack(a, b, st);
  h->launches += 1;
  if (!rc && !pack) { rc = launch_halo_outer_multi(a, st); h->launches += 1; }
  if (rc) return fail(h, HYCOM_TSADVC_ECUDA, "halo kernel launch failed: %s",
                      cudaGetErrorString((cudaError_t)rc));
  return 0;
}

int hycom_tsadvc_halo_pack(hycom_tsadvc_handle* h, int32_t m, int32_t n,
                           const hycom_tsadvc_params* prm, double* const sendbuf[8],
                           void* cuda_stream) {
  return halo_xfer(h, m, n, prm, sendbuf, cuda_stream, true);
}

int hycom_tsadvc_halo_unpack(hycom_tsadvc_handle* h, int32_t m, int32_t n,
                             const hycom_tsadvc_params* prm, double* const recvbuf[8],
                             void* cuda_stream) {
  return halo_xfer(h, m, n, prm, recvbuf, cuda_stream, false);
}

// Layers are independent (mod_tsadvc.F90:1842 `do k=1,kk`), so the host-array call is a
// pipeline over layer chunks on three streams: chunk c+1 is copied in while chunk c is advected
// and chunk c-1 is copied out.  PCIe is full duplex: the step costs about the host->device
// copy alone (it moves 3.5x the bytes of the way back) instead of copy-in + compute + copy-out.
int hycom_tsadvc_step(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_tsadvc_params* prm,
                      double* temp, double* saln, double* th3d, double* tracer, const double* dp,
                      const double* uflx, const double* vflx, const double* oneta, double* xmin,
                      double* xmax) {
  std::vector<Adv> adv;
  int mbdy = 0, rc;
  if ((rc = plan_step(h, m, n, prm, adv, mbdy))) return rc;
  if (h->d.ipr * h->d.jpr > 1 && !h->xc)   // never a step with the exchanges left out
    return fail(h, HYCOM_TSADVC_EUNSUPPORTED,
                "tsadvc on %d x %d tiles needs a communicator: hycom_tsadvc_comm_init", h->d.ipr, h->d.jpr);
  const int kk = h->d.kdm;
  const size_t fs = (size_t)h->ncols * h->nrows;  // Fortran slab
  const bool adv_th3d = prm->advflg != 0;
  double* first = adv_th3d ? th3d : temp;
  if (!first || !saln || !dp || !uflx || !vflx || (h->d.ntracr > 0 && !tracer) || (prm->isopyc && !th3d))
    return fail(h, HYCOM_TSADVC_EINVAL, "step: a required array is NULL");
  const bool diffuse = prm->temdf2 > 0.0;
  double* other = adv_th3d ? temp : th3d;   // the thermodynamic variable that is not advected
  const int ofield = adv_th3d ? HYCOM_F_TEMP : HYCOM_F_TH3D;
  if (diffuse && (!oneta || !other))
    return fail(h, HYCOM_TSADVC_EINVAL, "step: temdf2>0 needs oneta, temp and th3d");
  CU(h, cudaSetDevice(h->d.device));
  if (!h->up_stream) {
    CU(h, cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking));
    CU(h, cudaStreamCreateWithFlags(&h->down_stream, cudaStreamNonBlocking));
  }
  const char* ce = getenv("HYCOM_TSADVC_STEP_CHUNK");
  int chunk = ce ? atoi(ce) : 4;
  if (chunk < 1 || chunk > kk) chunk = kk;
  const bool fct2c = prm->btrmas && abs(prm->advtyp) == 2;   // advem_fct2c: its own layer batches
  if (fct2c) {
    if (!oneta) return fail(h, HYCOM_TSADVC_EINVAL, "step: btrmas needs oneta");
    chunk = kk;
  }
  if (prm->isopyc) chunk = kk;   // layer 1 runs on smoothed fluxes: advect_march splits the launch itself
  const int nchunks = (kk + chunk - 1) / chunk;
  while ((int)h->ev_chunk.size() < 2 * nchunks + 1) {
    cudaEvent_t e;
    CU(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->ev_chunk.push_back(e);
  }
  // host address of layer k0 (0-based) of slot t of a 4-D array / of tracer q
  auto h4 = [&](double* a, int t, int k0) { return a + fs * ((size_t)kk * (t - 1) + k0); };
  auto htr = [&](int q, int t, int k0) { return tracer + fs * ((size_t)kk * (2 * (size_t)(q - 1) + (t - 1)) + k0); };
  // the copies must not overtake earlier work of the handle on the mirrors they overwrite
  CU(h, cudaEventRecord(h->ev_chunk[2 * nchunks], h->stream));
  CU(h, cudaStreamWaitEvent(h->up_stream, h->ev_chunk[2 * nchunks], 0));
  CU(h, cudaStreamWaitEvent(h->down_stream, h->ev_chunk[2 * nchunks], 0));
  // make sure every mirror (and ping-pong buffer) exists before the streams fork
  for (const Adv& a : adv) {
    double* t;
    if ((rc = slot(h, a.field, a.ktr, 1, &t)) || (rc = slot(h, a.field, a.ktr, 2, &t)) ||
        (rc = spare_of(h, mirror_of(h, a.field, a.ktr), &t))) return rc;
  }
  { double* t; if ((rc = slot(h, HYCOM_F_DP, 0, n, &t))) return rc; }
  CU(h, cudaStreamSynchronize(h->stream));   // allocations zero-fill on h->stream

  const bool single = h->d.ipr * h->d.jpr == 1;
  HaloArrays xa;   // the arrays of the first exchange (:1829-1836), all layers
  if (!single && (rc = halo_arrays(h, adv, mbdy, xa))) return rc;
  const int nb = h->d.nbdy;
  auto host_of = [&](int field) { return field == HYCOM_F_SALN ? saln : field == HYCOM_F_TH3D ? th3d : temp; };
  auto resident = [](const Adv& a) { return a.field == HYCOM_F_Q2 || a.field == HYCOM_F_Q2L; };
  // D2H of layers k0..k0+nk-1 of a device buffer on 1:ii,1:jj ("valid halo 0 wide", :104)
  auto back = [&](const double* dev, double* host, int nk, cudaStream_t st) -> int {
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof cp);
    cp.srcPtr = make_cudaPitchedPtr((void*)dev, sizeof(double) * h->pitch, h->pitch, h->nrows);
    cp.dstPtr = make_cudaPitchedPtr(host, sizeof(double) * h->ncols, h->ncols, h->nrows);
    cp.srcPos = make_cudaPos(sizeof(double) * nb, nb, 0);
    cp.dstPos = make_cudaPos(sizeof(double) * nb, nb, 0);
    cp.extent = make_cudaExtent(sizeof(double) * h->d.ii, h->d.jj, nk);
    cp.kind = cudaMemcpyDeviceToHost;
    CU(h, cudaMemcpy3DAsync(&cp, st));
    return 0;
  };
  for (int c = 0; c < nchunks; ++c) {
    const int k0 = c * chunk, nk = (k0 + chunk <= kk) ? chunk : kk - k0;
    // ---- copy in (q2, q2l stay in their mirrors: the caller uploaded them)
    for (const Adv& a : adv) {
      if (k0 >= a.nlay || resident(a)) continue;   // layers below nhybrd of temp/th3d are not advected (:1855)
      const int na = (k0 + nk <= a.nlay) ? nk : a.nlay - k0;
      for (int t = 1; t <= 2; ++t) {
        double* src = a.field == HYCOM_F_TRACER ? htr(a.ktr, t, k0)
                                                : h4(host_of(a.field), t, k0);
        if ((rc = upload_on(h, a.field, a.ktr, t, k0 + 1, na, src, h->up_stream))) return rc;
      }
    }
    if ((rc = upload_on(h, HYCOM_F_DP, 0, n, k0 + 1, nk, h4((double*)dp, n, k0), h->up_stream))) return rc;
    if ((rc = upload_on(h, HYCOM_F_UFLX, 0, 1, k0 + 1, nk, uflx + fs * k0, h->up_stream))) return rc;
    if ((rc = upload_on(h, HYCOM_F_VFLX, 0, 1, k0 + 1, nk, vflx + fs * k0, h->up_stream))) return rc;
    CU(h, cudaEventRecord(h->ev_chunk[2 * c], h->up_stream));
    CU(h, cudaStreamWaitEvent(h->stream, h->ev_chunk[2 * c], 0));
    // ---- :1827-1836 halos: the layers of this chunk (layers are independent, so is their exchange)
    if (!single) {
      HaloArrays a = xa;
      for (int q = 0; q < a.narr; ++q) a.base[q] += h->slab * k0;
      a.kk = nk;
      if ((rc = xc_exchange(h, a, true, h->stream))) return rc;
    } else {
      for (const Adv& a : adv) {
        if (k0 >= a.nlay) continue;
        const int na = (k0 + nk <= a.nlay) ? nk : a.nlay - k0;
        if ((rc = halo_local_range(h, a.field, a.ktr, 0, mbdy, mbdy, k0 + a.koff, na))) return rc;
      }
      if ((rc = halo_local_range(h, HYCOM_F_UFLX, 0, 1, mbdy, mbdy, k0, nk))) return rc;
      if ((rc = halo_local_range(h, HYCOM_F_VFLX, 0, 1, mbdy, mbdy, k0, nk))) return rc;
    }
    // ---- advection of these layers into the ping-pong buffers
    if (fct2c) {   // onetamas(:,:,m) = oneta(:,:,n) (:1806)
      if ((rc = upload_on(h, HYCOM_F_ONETA, 0, n, 1, 1, oneta + fs * (n - 1), h->stream))) return rc;
      if ((rc = run_fct2c(h, m, n, *prm, adv))) return rc;
    } else if (prm->isopyc) {
      if ((rc = advect_march(h, m, n, *prm, adv, HYCOM_TSADVC_PART_ALL))) return rc;
    } else if ((rc = run_march(h, m, n, *prm, adv, HYCOM_TSADVC_PART_ALL, k0, nk))) {
      return rc;
    }
    if (diffuse) continue;   // the diffusion needs every layer's neighbours first: copy out after it
    CU(h, cudaEventRecord(h->ev_chunk[2 * c + 1], h->stream));
    CU(h, cudaStreamWaitEvent(h->down_stream, h->ev_chunk[2 * c + 1], 0));
    // ---- copy out (the new time level still sits in the ping-pong buffer)
    for (const Adv& a : adv) {
      if (k0 >= a.nlay || resident(a)) continue;
      const int na = (k0 + nk <= a.nlay) ? nk : a.nlay - k0;
      double* dst = a.field == HYCOM_F_TRACER ? htr(a.ktr, n, k0)
                                              : h4(host_of(a.field), n, k0);
      if ((rc = back(mirror_of(h, a.field, a.ktr)->spare + h->slab * k0, dst, na, h->down_stream))) return rc;
    }
  }
  // layers that were not advected must still be in the mirror the ping-pong swap retires
  for (const Adv& a : adv)
    if (a.nlay < kk && !resident(a)) {
      for (int t = 1; t <= 2; ++t) {
        double* src = a.field == HYCOM_F_TRACER ? htr(a.ktr, t, a.nlay)
                                                : h4(host_of(a.field), t, a.nlay);
        if ((rc = upload_on(h, a.field, a.ktr, t, a.nlay + 1, kk - a.nlay, src, h->stream))) return rc;
      }
    }
  if ((rc = finish_step(h, n, *prm, adv, xmin, xmax))) return rc;
  if (diffuse) {  // onetamas(:,:,n) = oneta(:,:,n) (:1805,1808); th3d/temp(:,:,:,n) (:2143-2144)
    if ((rc = upload_on(h, HYCOM_F_ONETA, 0, n, 1, 1, oneta + fs * (n - 1), h->stream))) return rc;
    bool other_advected = false;   // isopyc: th3d is advected in layer 1 and already in its mirror
    for (const Adv& a : adv) other_advected = other_advected || a.field == ofield;
    if (!other_advected && (rc = upload_on(h, ofield, 0, n, 1, kk, h4(other, n, 0), h->stream))) return rc;
    if ((rc = hycom_tsadvc_diffuse_device(h, m, n, prm))) return rc;
    auto back_all = [&](int field, int ktr, double* host_n) -> int {
      double* base;
      int r2 = slot(h, field, ktr, n, &base);
      return r2 ? r2 : back(base, host_n, kk, h->stream);
    };
    if ((rc = back_all(ffield_of(adv_th3d), 0, h4(first, n, 0)))) return rc;
    if ((rc = back_all(HYCOM_F_SALN, 0, h4(saln, n, 0)))) return rc;
    if ((rc = back_all(ofield, 0, h4(other, n, 0)))) return rc;
    for (int q = 1; q <= h->d.ntracr; ++q)
      if ((rc = back_all(HYCOM_F_TRACER, q, htr(q, n, 0)))) return rc;
  }
  CU(h, cudaStreamSynchronize(h->down_stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int hycom_tsadvc_xctilr(hycom_tsadvc_handle* h, int32_t field, int32_t ktr, int32_t tlev, int32_t mh, int32_t nh,
                        int32_t itype) {
  if (!h) return fail(nullptr, HYCOM_TSADVC_EINVAL, "null handle");
  if (h->d.ipr * h->d.jpr == 1) return hycom_tsadvc_halo_local(h, field, ktr, tlev, mh, nh);
  if (!h->xc) return fail(h, HYCOM_TSADVC_EUNSUPPORTED, "xctilr on %d x %d tiles needs a communicator", h->d.ipr, h->d.jpr);
  if (mh < 0 || nh < 0 || mh > h->d.nbdy || nh > h->d.nbdy) return fail(h, HYCOM_TSADVC_EINVAL, "xctilr: bad halo width %d,%d", mh, nh);
  CU(h, cudaSetDevice(h->d.device));
  HaloArrays a;
  memset(&a, 0, sizeof a);
  const int t0 = is3d(field) ? 1 : (tlev == 0 ? 1 : tlev), t1 = is3d(field) ? 1 : (tlev == 0 ? 2 : tlev);
  int rc;
  for (int t = t0; t <= t1; ++t) {
    if ((rc = slot(h, field, ktr, t, &a.base[a.narr]))) return rc;
    a.itype[a.narr++] = itype;
  }
  a.kk = nlayers_of(h, field); a.slab = h->slab; a.pitch = h->pitch; a.nrows = h->nrows; a.nbdy = h->d.nbdy;
  a.ii = h->d.ii; a.jj = h->d.jj; a.mh = mh; a.nh = nh;
  a.fold = arctic_fold(h->d);
  return xc_exchange(h, a, false, h->stream);
}

int hycom_tsadvc_set_deferred_range(hycom_tsadvc_handle* h, int32_t enable) {
  if (!h) return fail(nullptr, HYCOM_TSADVC_EINVAL, "null handle");
  h->deferred_range = enable != 0;
  return 0;
}

int hycom_tsadvc_saln_range(hycom_tsadvc_handle* h, double* xmin, double* xmax, int32_t* nstep) {
  if (!h || !xmin || !xmax) return fail(h, HYCOM_TSADVC_EINVAL, "saln_range: null argument");
  if (nstep) *nstep = h->range_nstep;
  if (h->range_nstep < 0 || !h->range_host) return 0;
  CU(h, cudaEventSynchronize(h->ev_range));
  const int kk = h->d.kdm;
  for (int k = 0; k < kk; ++k) { xmin[k] = h->range_host[k]; xmax[k] = h->range_host[kk + k]; }
  h->range_nstep = -1;
  return 0;
}

int hycom_tsadvc_checksum(hycom_tsadvc_handle* h, int32_t field, int32_t ktr, int32_t tlev, int32_t global,
                          uint64_t* sum) {
  if (!h || !sum) return fail(h, HYCOM_TSADVC_EINVAL, "checksum: null argument");
  if (!h->have_static) return fail(h, HYCOM_TSADVC_EINVAL, "checksum: set_static has not been called");
  CU(h, cudaSetDevice(h->d.device));
  double* base;
  int rc = slot(h, field, ktr, tlev, &base);
  if (rc) return rc;
  if (!h->d_cksum && (rc = dalloc(h, (void**)&h->d_cksum, sizeof(unsigned long long), false))) return rc;
  CU(h, cudaMemsetAsync(h->d_cksum, 0, sizeof(unsigned long long), h->stream));
  const int nk = h->d.kdm;   // the model layers 1..kdm (q2, q2l: slabs 1..kdm of 0:kdm+1)
  k_checksum<<<148 * 8, 256, 0, h->stream>>>(base + h->slab * layer1_of(field), h->mask, h->slab, h->pitch, h->nrows,
                                             h->d.nbdy, h->d.i0, h->d.j0, h->d.itdm, h->d.jtdm, nk, h->d_cksum);
  h->launches += 1;
  CU(h, cudaGetLastError());
  if (global && h->d.ipr * h->d.jpr > 1) {
    if (!h->xc) return fail(h, HYCOM_TSADVC_EUNSUPPORTED, "checksum over all tiles needs a communicator");
    if ((rc = xc_sum_u64(h, h->d_cksum, 1, h->stream))) return rc;
  }
  unsigned long long v = 0;
  CU(h, cudaMemcpyAsync(&v, h->d_cksum, sizeof v, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  *sum = (uint64_t)v;
  return 0;
}

// ---- cnuity(m,n) on the device mirrors (cnuity.F90) ------------------------------------------------
int hycom_tsadvc_cnuity_device(hycom_tsadvc_handle* h, int32_t m, int32_t n, const hycom_cnuity_params* prm,
                               double* dpkmin) {
  if (!h || !prm) return fail(h, HYCOM_TSADVC_EINVAL, "cnuity: null argument");
  if (!h->have_static) return fail(h, HYCOM_TSADVC_EINVAL, "cnuity: set_static has not been called");
  if (m < 1 || m > 2 || n < 1 || n > 2 || m == n) return fail(h, HYCOM_TSADVC_EINVAL, "cnuity: bad leapfrog slots m=%d n=%d", m, n);
  if (h->d.nbdy < 6) return fail(h, HYCOM_TSADVC_ENBDY, "error: cnuity needs nbdy >= 6 (mbdy = 6, cnuity.F90:98)");
  if (prm->btrmas) return fail(h, HYCOM_TSADVC_EUNSUPPORTED, "cnuity with btrmas (oneta_u, oneta_v, onetacnt) is not built");
  if (prm->thkdf2 != 0.0 && prm->thkdf4 != 0.0)
    return fail(h, HYCOM_TSADVC_EINVAL, "cnuity: only one of thkdf2 and thkdf4 is non-zero (cnuity.F90:758)");
  const bool thk = prm->thkdf2 != 0.0 || prm->thkdf4 != 0.0;
  if (thk && (!h->thkdf4u.lev[0] || !h->thkdf4v.lev[0]))
    return fail(h, HYCOM_TSADVC_EINVAL, "cnuity with thkdf2/thkdf4 reads thkdf4u, thkdf4v: upload HYCOM_F_THKDF4U/_THKDF4V first");
  if (!h->scuy || !h->scvx) return fail(h, HYCOM_TSADVC_EINVAL, "cnuity needs scuy, scvx (set_static)");
  CU(h, cudaSetDevice(h->d.device));
  const int kk = h->d.kdm;
  const bool single = h->d.ipr * h->d.jpr == 1;
  if (!single && !h->xc) return fail(h, HYCOM_TSADVC_EUNSUPPORTED, "cnuity on %d x %d tiles needs a communicator", h->d.ipr, h->d.jpr);
  int rc;
  CnuityParams P;
  memset(&P, 0, sizeof P);
  double *um, *vm, *dpum, *dpvm, *ub, *vb, *du, *dv, *pb, *uflx, *vflx;
  if ((rc = slot(h, HYCOM_F_DP, 0, n, &P.dp_n)) || (rc = slot(h, HYCOM_F_DP, 0, m, &P.dp_m)) ||
      (rc = slot(h, HYCOM_F_DPO, 0, n, &P.dpo_n)) || (rc = slot(h, HYCOM_F_DPO, 0, m, &P.dpo_m)) ||
      (rc = slot(h, HYCOM_F_U, 0, m, &um)) || (rc = slot(h, HYCOM_F_V, 0, m, &vm)) ||
      (rc = slot(h, HYCOM_F_DPU, 0, m, &dpum)) || (rc = slot(h, HYCOM_F_DPV, 0, m, &dpvm)) ||
      (rc = slot(h, HYCOM_F_UBAVG, 0, 1, &ub)) || (rc = slot(h, HYCOM_F_VBAVG, 0, 1, &vb)) ||
      (rc = slot(h, HYCOM_F_DEPTHU, 0, 1, &du)) || (rc = slot(h, HYCOM_F_DEPTHV, 0, 1, &dv)) ||
      (rc = slot(h, HYCOM_F_PBOT, 0, 1, &pb)) || (rc = slot(h, HYCOM_F_UFLX, 0, 1, &uflx)) ||
      (rc = slot(h, HYCOM_F_VFLX, 0, 1, &vflx)) || (rc = slot(h, HYCOM_F_P, 0, 1, &P.p)) ||
      (rc = slot(h, HYCOM_F_DPMIXL, 0, n, &P.dpmixl_n)) || (rc = slot(h, HYCOM_F_DPMOLD, 0, 1, &P.dpmold)) ||
      (rc = slot(h, HYCOM_F_UTOTN, 0, 1, &P.utotn)) || (rc = slot(h, HYCOM_F_VTOTN, 0, 1, &P.vtotn)))
    return rc;
  P.u_m = um; P.v_m = vm; P.dpu_m = dpum; P.dpv_m = dpvm;
  P.ubavg_m = ub + h->slab * (m - 1); P.vbavg_m = vb + h->slab * (m - 1);
  P.depthu = du; P.depthv = dv; P.pbot = pb; P.uflx = uflx; P.vflx = vflx;
  P.uflxav = h->uflxav.lev[0]; P.vflxav = h->vflxav.lev[0]; P.dpav = h->dpav.lev[0];
  if (!h->cnuity_scratch && (rc = dalloc_field(h, &h->cnuity_scratch, (size_t)kk * h->slab))) return rc;
  if (!h->d_dpkmin && (rc = dalloc(h, (void**)&h->d_dpkmin, sizeof(double) * 2 * kk, false))) return rc;
  P.dnew = h->cnuity_scratch;
  if (thk) {   // the interface-depth diffusion of :745-1124 and the cumulative fluxes behind it
    if (!h->thk_scratch && (rc = dalloc_field(h, &h->thk_scratch, 5 * (size_t)h->slab))) return rc;
    P.pold = h->thk_scratch; P.t1 = P.pold + h->slab; P.t2 = P.t1 + h->slab; P.fu = P.t2 + h->slab; P.fv = P.fu + h->slab;
    P.thku = h->thkdf4u.lev[0]; P.thkv = h->thkdf4v.lev[0]; P.scp2 = h->scp2;
    P.defer_av = 1;
  }
  P.dpkmin = h->d_dpkmin;
  P.pitch = h->pitch; P.nrows = h->nrows; P.nbdy = h->d.nbdy; P.ii = h->d.ii; P.jj = h->d.jj; P.kk = kk;
  P.slab = h->slab; P.mask = h->mask; P.scuy = h->scuy; P.scvx = h->scvx; P.scp2i = h->scp2i;
  P.delt1 = prm->delt1; P.ra2fac = prm->ra2fac; P.isopyc = prm->isopyc;
  // :100-107 the eight xctilr calls, width 6
  if (single) {
    if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_DPMIXL, 0, n, 6, 6)) || (rc = hycom_tsadvc_halo_local(h, HYCOM_F_DP, 0, 0, 6, 6)) ||
        (rc = hycom_tsadvc_halo_local(h, HYCOM_F_DPU, 0, m, 6, 6)) || (rc = hycom_tsadvc_halo_local(h, HYCOM_F_DPV, 0, m, 6, 6)) ||
        (rc = hycom_tsadvc_halo_local(h, HYCOM_F_U, 0, m, 6, 6)) || (rc = hycom_tsadvc_halo_local(h, HYCOM_F_V, 0, m, 6, 6)) ||
        (rc = halo_local_range(h, HYCOM_F_UBAVG, 0, 1, 6, 6, m - 1, 1)) || (rc = halo_local_range(h, HYCOM_F_VBAVG, 0, 1, 6, 6, m - 1, 1)))
      return rc;
  } else {
    // one exchange of kk-layer arrays (dp both slots, dpu, dpv, u, v) and one of the 2-D arrays
    HaloArrays a;
    memset(&a, 0, sizeof a);
    double* b3[6] = {P.dp_n, P.dp_m, dpum, dpvm, um, vm};
    const int t3[6] = {1, 1, 3, 4, 13, 14};
    for (int q = 0; q < 6; ++q) { a.base[a.narr] = b3[q]; a.itype[a.narr++] = t3[q]; }
    a.kk = kk; a.slab = h->slab; a.pitch = h->pitch; a.nrows = h->nrows; a.nbdy = h->d.nbdy;
    a.ii = h->d.ii; a.jj = h->d.jj; a.mh = 6; a.nh = 6; a.fold = arctic_fold(h->d);
    if ((rc = xc_exchange(h, a, false, h->stream))) return rc;
    HaloArrays b = a;
    b.narr = 0;
    double* b2[3] = {P.dpmixl_n, ub + h->slab * (m - 1), vb + h->slab * (m - 1)};
    const int t2[3] = {1, 13, 14};
    for (int q = 0; q < 3; ++q) { b.base[b.narr] = b2[q]; b.itype[b.narr++] = t2[q]; }
    b.kk = 1;
    if ((rc = xc_exchange(h, b, false, h->stream))) return rc;
  }
  if ((rc = launch_cnuity(0, P, h->stream))) return fail(h, HYCOM_TSADVC_ECUDA, "cnuity kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  h->launches += 3;
  if (thk) {   // :761-763 / :981-983 xctilr of dpmixl(n), dp(n), p(2:kk+1), width 6
    if (single) {
      if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_DPMIXL, 0, n, 6, 6)) || (rc = hycom_tsadvc_halo_local(h, HYCOM_F_DP, 0, n, 6, 6)) ||
          (rc = halo_local_range(h, HYCOM_F_P, 0, 1, 6, 6, 1, kk)))
        return rc;
    } else {
      HaloArrays a;
      memset(&a, 0, sizeof a);
      a.base[0] = P.dp_n; a.base[1] = P.p + h->slab; a.itype[0] = a.itype[1] = 1; a.narr = 2;
      a.kk = kk; a.slab = h->slab; a.pitch = h->pitch; a.nrows = h->nrows; a.nbdy = h->d.nbdy;
      a.ii = h->d.ii; a.jj = h->d.jj; a.mh = 6; a.nh = 6; a.fold = arctic_fold(h->d);
      if ((rc = xc_exchange(h, a, false, h->stream))) return rc;
      HaloArrays b = a;
      b.narr = 1; b.base[0] = P.dpmixl_n; b.kk = 1;
      if ((rc = xc_exchange(h, b, false, h->stream))) return rc;
    }
    const int nl = launch_cnuity_thkdf(P, prm->thkdf4 != 0.0, prm->nstep, h->stream);
    if (nl < 0) return fail(h, HYCOM_TSADVC_ECUDA, "cnuity thickness-diffusion launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    h->launches += nl;
  }
  if (prm->hybrid && prm->mxlkta) {   // :1144-1324 dpmixl follows the coordinates around its base, then diffuses
    const int nl = launch_cnuity_mxlkta(P, prm->thkdf4 != 0.0 ? 1 : prm->thkdf2 != 0.0 ? 2 : 0, h->stream);
    if (nl < 0) return fail(h, HYCOM_TSADVC_ECUDA, "cnuity dpmixl launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    h->launches += nl;
  }
  // :1400 xctilr(dp(:,:,:,n), 1,kk, 6,6, halo_ps), then the Robert-Asselin filter
  if (single) {
    if ((rc = hycom_tsadvc_halo_local(h, HYCOM_F_DP, 0, n, 6, 6))) return rc;
  } else {
    HaloArrays a;
    memset(&a, 0, sizeof a);
    a.base[0] = P.dp_n; a.itype[0] = 1; a.narr = 1;
    a.kk = kk; a.slab = h->slab; a.pitch = h->pitch; a.nrows = h->nrows; a.nbdy = h->d.nbdy;
    a.ii = h->d.ii; a.jj = h->d.jj; a.mh = 6; a.nh = 6; a.fold = arctic_fold(h->d);
    if ((rc = xc_exchange(h, a, false, h->stream))) return rc;
  }
  if ((rc = launch_cnuity(1, P, h->stream))) return fail(h, HYCOM_TSADVC_ECUDA, "cnuity kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  h->launches += 1;
  if (dpkmin && prm->nstep % 3 == 0) {   // :513-516: evaluated every third step
    CU(h, cudaMemcpyAsync(dpkmin, h->d_dpkmin, sizeof(double) * 2 * kk, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  CU(h, cudaGetLastError());
  return 0;
}

}  // extern "C"
