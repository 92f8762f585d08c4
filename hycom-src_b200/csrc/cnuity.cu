// cnuity(m,n) on the device mirrors (cnuity.F90): the continuity
// equation, flux-corrected transport of the layer thickness - the producer of dp(:,:,:,n), uflx, vflx
// that tsadvc(m,n) consumes.
//
// The reference runs loop 76 serially over the layers: layer k needs util3 = the sum of the OLD
// thicknesses above it (depth-limited donor thickness, :236-283), and utotn/vtotn/p accumulate over k.
// Those are prefix sums and reductions of quantities that do not depend on the transport itself, so
// here every sweep covers ALL layers at once:
//   init    :116-156   dpo(:,:,:,n) = dp(:,:,:,n); U3(k) = sum of the old dp above k (columns, in k order)
//   flux    :236-283   low-order fluxes uflux, vflux and antidiffusive uflux2, vflux2 (margin 5)
//   low     :293-311   dp advanced with the low-order fluxes (margin 4)
//   ratio   :378-400   util1, util2 (5-point sea-only extrema of the low-order dp)
//   limit   :414-441   clipped antidiffusive fluxes; the clipped-off part per layer (margin 3)
//   update  :449-469   dp with the clipped fluxes (margin 2)
//   column  (:437-441, :459) utotn, vtotn = sums over k IN ORDER of the clipped-off parts; p(:,:,k+1)
//   f77     :588-651   loop 77: the lost flux goes back in proportion dp/p(kk+1) (margin 1)
//   u77     :657-683   dp, p (margin 0)
//   bottom  :716-733   bottom-pressure restoring, :1326-1350 cumulative fluxes
//   asselin :1396-1422 Robert-Asselin filter of dp behind one more exchange of dp(:,:,:,n)
// Arithmetic is the Fortran's, expression by expression (-fmad=false, IEEE division); fluxes are zero
// off the iu / iv points (geopar.F90:822-871 zeroes them there once, cnuity never writes them).
// Scope: .not.btrmas, thkdf2 = thkdf4 = 0, no open-boundary faces, no Stokes drift, not (hybrid .and.
// mxlkta), not (synflt .and. wvelfl) - everything else is refused by the caller (tsadvc_abi.cu).
// These are streaming sweeps (HBM bound, about 10 passes over the 3-D state); the marching form of the
// advection kernels is the next step for loop 76.
#include <cuda_runtime.h>

#include "tsadvc_dev.h"
#include "tsadvc_launch.h"

namespace tsadvc {

namespace {

__device__ __forceinline__ double cmax(double a, double b) { return a > b ? a : b; }   // Fortran max
__device__ __forceinline__ double cmin(double a, double b) { return a < b ? a : b; }

__device__ __forceinline__ bool in_margin(const CnuityParams& P, int c, int r, int margin) {
  const int i = c + 1 - P.nbdy, j = r + 1 - P.nbdy;
  return i >= 1 - margin && i <= P.ii + margin && j >= 1 - margin && j <= P.jj + margin;
}

__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}
// min over the warp, then one atomic per warp
__device__ __forceinline__ void layer_min(double* addr, double v) {
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = v < w ? v : w;
  }
  if ((threadIdx.x & 31) == 0 && v < 999.0) atomic_min_f64(addr, v);
}

#define CN_CELL                                                                  \
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;  \
  const bool inside = c < P.pitch && r < P.nrows;                                \
  const long q = (long)r * P.pitch + c

// ---- init: columns, margin 6 -------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cn_init(const CnuityParams P) {
  CN_CELL;
  if (blockIdx.x == 0 && blockIdx.y == 0)   // per-layer minima of loops 19/14 and 15 start at 999. (:295, :451)
    for (int k = threadIdx.y * 32 + threadIdx.x; k < 2 * P.kk; k += 256) P.dpkmin[k] = 999.0;
  if (!inside || !in_margin(P, c, r, 6)) return;
  const bool acc = (P.mask[q] & M_IP) && in_margin(P, c, r, 4);   // where :300 advances util3
  double u3 = 0.0;
  P.dpmold[q] = P.dpmixl_n[q];
  for (int k = 0; k < P.kk; ++k) {
    const long qk = q + (long)k * P.slab;
    const double d = P.dp_n[qk];
    P.dpo_n[qk] = d;
    P.u3[qk] = u3;
    if (acc) u3 = u3 + d;
  }
}

// ---- low-order and antidiffusive fluxes, margin 5 ------------------------------------------------
__global__ void __launch_bounds__(256) k_cn_flux(const CnuityParams P) {
  CN_CELL;
  if (!inside) return;
  const int k = blockIdx.z;
  const long qk = q + (long)k * P.slab;
  const bool m5 = in_margin(P, c, r, 5);
  const unsigned mk = P.mask[q];
  double fu = 0.0, fu2 = 0.0, fv = 0.0, fv2 = 0.0;
  if (m5 && (mk & M_IU)) {
    const double utotm = (P.u_m[qk] + P.ubavg_m[q]) * P.scuy[q];
    double qq;
    if (utotm >= 0.0) qq = cmin(P.dpo_n[qk - 1], cmax(0.0, P.depthu[q] - P.u3[qk - 1]));
    else qq = cmin(P.dpo_n[qk], cmax(0.0, P.depthu[q] - P.u3[qk]));
    fu = utotm * qq;
    fu2 = utotm * P.dpu_m[qk] - fu;
    P.uflx[qk] = fu;
  }
  if (m5 && (mk & M_IV)) {
    const double vtotm = (P.v_m[qk] + P.vbavg_m[q]) * P.scvx[q];
    double qq;
    if (vtotm >= 0.0) qq = cmin(P.dpo_n[qk - P.pitch], cmax(0.0, P.depthv[q] - P.u3[qk - P.pitch]));
    else qq = cmin(P.dpo_n[qk], cmax(0.0, P.depthv[q] - P.u3[qk]));
    fv = vtotm * qq;
    fv2 = vtotm * P.dpv_m[qk] - fv;
    P.vflx[qk] = fv;
  }
  P.uf[qk] = fu; P.uf2[qk] = fu2; P.vf[qk] = fv; P.vf2[qk] = fv2;
}

// ---- dp -= div(uf, vf)*delt1*scp2i on sea cells of a margin; min over rows 1..jj into dpkmin[slot] -----
// src: the thickness the update starts from (dpo(n) for the low-order step, dp(n) itself afterwards)
template <bool LOW>
__global__ void __launch_bounds__(256) k_cn_update(const CnuityParams P, int margin, int minslot) {
  CN_CELL;
  const int k = blockIdx.z;
  double seen = 999.0;
  if (inside && in_margin(P, c, r, margin) && (P.mask[q] & M_IP)) {
    const long qk = q + (long)k * P.slab;
    const double d0 = LOW ? P.dpo_n[qk] : P.dp_n[qk];
    const double d = d0 - ((P.uf[qk + 1] - P.uf[qk]) + (P.vf[qk + P.pitch] - P.vf[qk])) * P.delt1 * P.scp2i[q];
    P.dp_n[qk] = d;
    const int j = r + 1 - P.nbdy;
    if (j >= 1 && j <= P.jj) seen = d;
  }
  if (minslot >= 0) layer_min(&P.dpkmin[minslot * P.kk + k], seen);
}

// ---- util1, util2, margin 4 -----------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cn_ratio(const CnuityParams P) {
  CN_CELL;
  if (!inside || !in_margin(P, c, r, 4)) return;
  const unsigned mk = P.mask[q];
  if (!(mk & M_IP)) return;
  const int k = blockIdx.z;
  const long qk = q + (long)k * P.slab;
  const double* d = P.dp_n;
  const double d0 = d[qk];
  const double d1 = (mk & M_PW) ? d[qk - 1] : d0, d2 = (mk & M_PE) ? d[qk + 1] : d0;
  const double d3 = (mk & M_PS) ? d[qk - P.pitch] : d0, d4 = (mk & M_PN) ? d[qk + P.pitch] : d0;
  double u1 = cmax(cmax(cmax(cmax(d0, d1), d2), d3), d4);
  double u2 = cmax(0.0, cmin(cmin(cmin(cmin(d0, d1), d2), d3), d4));
  const double a = P.uf2[qk], b = P.uf2[qk + 1], e = P.vf2[qk], f = P.vf2[qk + P.pitch];
  const double epsil = 1.0e-11;   // mod_cb_arrays.F90:853
  u1 = (u1 - d0) / (((cmax(0.0, a) - cmin(0.0, b)) + (cmax(0.0, e) - cmin(0.0, f)) + epsil) * P.delt1 * P.scp2i[q]);
  u2 = (u2 - d0) / (((cmin(0.0, a) - cmax(0.0, b)) + (cmin(0.0, e) - cmax(0.0, f)) - epsil) * P.delt1 * P.scp2i[q]);
  P.r1[qk] = u1; P.r2[qk] = u2;
}

// ---- limiter, margin 3: uf, vf := clipped antidiffusive flux; tn := the clipped-off part -----------------
__global__ void __launch_bounds__(256) k_cn_limit(const CnuityParams P) {
  CN_CELL;
  if (!inside) return;
  const int k = blockIdx.z;
  const long qk = q + (long)k * P.slab;
  const bool m3 = in_margin(P, c, r, 3);
  const unsigned mk = P.mask[q];
  double tu = 0.0, tv = 0.0;
  if (m3 && (mk & M_IU)) {
    const double f2 = P.uf2[qk];
    double clip;
    if (f2 >= 0.0) clip = cmin(cmin(1.0, P.r1[qk]), P.r2[qk - 1]);
    else clip = cmin(cmin(1.0, P.r2[qk]), P.r1[qk - 1]);
    tu = f2 * (1.0 - clip);
    const double fc = f2 * clip;
    P.uf[qk] = fc;
    P.uflx[qk] = P.uflx[qk] + fc;
  }
  if (m3 && (mk & M_IV)) {
    const double f2 = P.vf2[qk];
    double clip;
    if (f2 >= 0.0) clip = cmin(cmin(1.0, P.r1[qk]), P.r2[qk - P.pitch]);
    else clip = cmin(cmin(1.0, P.r2[qk]), P.r1[qk - P.pitch]);
    tv = f2 * (1.0 - clip);
    const double fc = f2 * clip;
    P.vf[qk] = fc;
    P.vflx[qk] = P.vflx[qk] + fc;
  }
  P.tnu[qk] = tu; P.tnv[qk] = tv;
}

// ---- columns: utotn, vtotn (sums over k in order) and p(:,:,k+1) after loop 76 ---------------------------
__global__ void __launch_bounds__(256) k_cn_column(const CnuityParams P) {
  CN_CELL;
  if (blockIdx.x == 0 && blockIdx.y == 0)   // loop 14 reuses dpkmin(1:kk) (:679)
    for (int k = threadIdx.y * 32 + threadIdx.x; k < P.kk; k += 256) P.dpkmin[k] = 999.0;
  if (!inside) return;
  const unsigned mk = P.mask[q];
  const bool m3 = in_margin(P, c, r, 3);
  const bool cell = in_margin(P, c, r, 2) && (mk & M_IP);
  double un = 0.0, vn = 0.0, pk = 0.0;
  P.p[q] = 0.0;
  for (int k = 0; k < P.kk; ++k) {
    const long qk = q + (long)k * P.slab;
    if (m3 && (mk & M_IU)) un = un + P.tnu[qk];
    if (m3 && (mk & M_IV)) vn = vn + P.tnv[qk];
    if (cell) { pk = pk + P.dp_n[qk]; P.p[qk + P.slab] = pk; }
  }
  P.utotn[q] = un; P.vtotn[q] = vn;
}

// ---- loop 77 fluxes, margin 1 -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cn_f77(const CnuityParams P) {
  CN_CELL;
  if (!inside) return;
  const int k = blockIdx.z;
  const long qk = q + (long)k * P.slab;
  const bool m1 = in_margin(P, c, r, 1);
  const unsigned mk = P.mask[q];
  const double* pb = P.p + (long)P.kk * P.slab;   // p(:,:,kk+1) as loop 76 left it
  double fu = 0.0, fv = 0.0;
  if (m1 && (mk & M_IU)) {
    const double un = P.utotn[q];
    const double qq = (un >= 0.0) ? P.dp_n[qk - 1] / pb[q - 1] : P.dp_n[qk] / pb[q];
    fu = un * qq;
    P.uflx[qk] = P.uflx[qk] + fu;
  }
  if (m1 && (mk & M_IV)) {
    const double vn = P.vtotn[q];
    const double qq = (vn >= 0.0) ? P.dp_n[qk - P.pitch] / pb[q - P.pitch] : P.dp_n[qk] / pb[q];
    fv = vn * qq;
    P.vflx[qk] = P.vflx[qk] + fv;
  }
  P.uf2[qk] = fu; P.vf2[qk] = fv;   // (uf, vf still feed nothing; uf2, vf2 are free: no read-write overlap with u77)
}

// ---- loop 77 update, margin 0: reads the fluxes of f77 from uf2, vf2 ------------------------------------------
__global__ void __launch_bounds__(256) k_cn_u77(const CnuityParams P) {
  CN_CELL;
  const int k = blockIdx.z;
  double seen = 999.0;
  if (inside && (P.mask[q] & M_OUT)) {
    const long qk = q + (long)k * P.slab;
    const double d = P.dp_n[qk] - ((P.uf2[qk + 1] - P.uf2[qk]) + (P.vf2[qk + P.pitch] - P.vf2[qk])) * P.delt1 * P.scp2i[q];
    P.dp_n[qk] = d;
    seen = d;
  }
  layer_min(&P.dpkmin[k], seen);
}

// ---- bottom-pressure restoring (:716-733) and cumulative fluxes (:1326-1350), columns of 1:ii,1:jj -----------
__global__ void __launch_bounds__(256) k_cn_bottom(const CnuityParams P) {
  CN_CELL;
  if (!inside) return;
  const int i = c + 1 - P.nbdy, j = r + 1 - P.nbdy;
  if (i < 1 || i > P.ii || j < 1 || j > P.jj) return;
  const unsigned mk = P.mask[q];
  if (mk & M_IP) {
    double pk = 0.0;
    for (int k = 0; k < P.kk; ++k) pk = pk + P.dp_n[q + (long)k * P.slab];   // p(kk+1) after loop 77
    const double qq = P.pbot[q] / pk;
    pk = 0.0;
    for (int k = 0; k < P.kk; ++k) {
      const long qk = q + (long)k * P.slab;
      const double d = P.dp_n[qk] * qq;
      P.dp_n[qk] = d;
      pk = pk + d;
      P.p[qk + P.slab] = pk;
      if (P.dpav) P.dpav[qk] = P.dpav[qk] + d;
    }
    if (P.isopyc) P.dpmixl_n[q] = P.dp_n[q];
  }
  if (P.uflxav && (mk & M_IU))
    for (int k = 0; k < P.kk; ++k) { const long qk = q + (long)k * P.slab; P.uflxav[qk] = P.uflxav[qk] + P.uflx[qk]; }
  if (P.vflxav && (mk & M_IV))
    for (int k = 0; k < P.kk; ++k) { const long qk = q + (long)k * P.slab; P.vflxav[qk] = P.vflxav[qk] + P.vflx[qk]; }
}

// ---- Robert-Asselin filter of dp, margin 6 (:1404-1420) ---------------------------------------------------------
__global__ void __launch_bounds__(256) k_cn_asselin(const CnuityParams P) {
  CN_CELL;
  if (!inside || !in_margin(P, c, r, 6) || !(P.mask[q] & M_IP)) return;
  const int k = blockIdx.z;
  const long qk = q + (long)k * P.slab;
  const double dpold = P.dpo_n[qk], dpmid = P.dp_m[qk], dpnew = P.dp_n[qk];
  const double qq = 0.5 * P.ra2fac * (dpold + dpnew - 2.0 * dpmid);
  P.dpo_m[qk] = dpmid;
  P.dp_m[qk] = dpmid + qq;
}

}  // namespace

// stage 0: everything up to the exchange of dp(:,:,:,n) (:1400); stage 1: the Robert-Asselin filter
int launch_cnuity(int stage, const CnuityParams& P, cudaStream_t st) {
  const dim3 block(32, 8), g2((P.pitch + 31) / 32, (P.nrows + 7) / 8), g3(g2.x, g2.y, P.kk);
  if (stage == 1) {
    k_cn_asselin<<<g3, block, 0, st>>>(P);
    return (int)cudaGetLastError();
  }
  k_cn_init<<<g2, block, 0, st>>>(P);
  k_cn_flux<<<g3, block, 0, st>>>(P);
  k_cn_update<true><<<g3, block, 0, st>>>(P, 4, 0);     // loop 19
  k_cn_ratio<<<g3, block, 0, st>>>(P);
  k_cn_limit<<<g3, block, 0, st>>>(P);
  k_cn_update<false><<<g3, block, 0, st>>>(P, 2, 1);    // loop 15
  k_cn_column<<<g2, block, 0, st>>>(P);
  k_cn_f77<<<g3, block, 0, st>>>(P);
  k_cn_u77<<<g3, block, 0, st>>>(P);                    // loop 14 (overwrites the loop-19 minima, like the Fortran)
  k_cn_bottom<<<g2, block, 0, st>>>(P);
  return (int)cudaGetLastError();
}

}  // namespace tsadvc
