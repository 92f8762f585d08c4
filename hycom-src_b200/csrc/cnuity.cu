// cnuity(m,n) on the device mirrors (cnuity.F90): the continuity
// equation, flux-corrected transport of the layer thickness - the producer of dp(:,:,:,n), uflx, vflx
// that tsadvc(m,n) consumes.
//
// The reference runs loop 76 serially over the layers: layer k needs util3 = the sum of the OLD
// thicknesses above it (depth-limited donor thickness, :236-283), and utotn/vtotn/p accumulate over k.
// Here a thread block owns a tile of the horizontal plane and walks down the layers itself:
//   k_cn_loop76  :116-511  per layer: dpo = dp; low-order and antidiffusive fluxes (margin 5); loop 19, the
//                          low-order step (margin 4); util1, util2; the limiter (margin 3: the clipped-off
//                          part joins utotn, vtotn IN k ORDER); loop 15 (margin 2); p(kk+1)
//   k_cn_loop77  :588-733  loop 77: the lost flux goes back in proportion dp/p(kk+1) (margin 1), dp (margin
//                :1326-1350 0); bottom-pressure restoring; cumulative fluxes
//   k_cn_asselin :1396-1422 Robert-Asselin filter of dp behind one more exchange of dp(:,:,:,n)
// Arithmetic is the Fortran's, expression by expression (-fmad=false, IEEE division); fluxes are zero
// off the iu / iv points (geopar.F90:822-871 zeroes them there once, cnuity never writes them).
//   k_thk_*      :745-1124 interface-depth diffusion, biharmonic (thkdf4) or Laplacian (thkdf2): three kernels per
//                          interface, behind their own exchange of dpmixl(n), dp(n), p
//   k_mxl_*      :1144-1324 hybrid .and. mxlkta: vertical advection and thickness diffusion of dpmixl
// Scope: .not.btrmas, no open-boundary faces, no Stokes drift, not (synflt .and. wvelfl) - everything else is
// refused by the caller (tsadvc_abi.cu).
// Measured at GLBb0.08 (profiles/r02x-z): ten streaming sweeps over all layers (the first version, 64
// passes over a 3-D field through nine scratch fields) 66 ms; the tile kernel with its operands loaded
// where they are used 60 ms (47 % of the warps' time waiting for them); operands of layer k+1 requested
// while layer k is finished, 2-D metrics in shared memory 44.5 ms; donor cells by select instead of
// branches, zero dividends on the fast division path 40.4 ms.
#include <cuda_runtime.h>

#include "tsadvc_dev.h"
#include "tsadvc_launch.h"
#include "march_common.cuh"

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


// =====================================================================================================
// Loop 76 (fluxes, low-order step, ratios, limiter, update) of ALL layers in one kernel:
// a block owns a tile of 58 x 26 cells, stages a window of 64 x 32 (dependency radius 3) in shared
// memory and walks down the layers, so that the prefix sum util3 of the old thicknesses and the column
// sums utotn, vtotn, p(kk+1) run in registers / shared memory in the reference's k order and no
// intermediate (u3, uf, vf, uf2, vf2, r1, r2, tnu, tnv of the sweeps above) ever reaches HBM.  The new
// thickness goes to a scratch field (the neighbouring tiles still read the old one).
// Loop 77, the bottom-pressure restoring and the cumulative fluxes follow as one column kernel that
// recomputes the four face fluxes of a cell from the 5-point stencil of that scratch field.
// HBM passes over a 3-D field: 64 (sweeps) -> about 30.
// =====================================================================================================
constexpr int CW = 64, CH = 32, CHALO = 3, CUX = CW - 2 * CHALO, CUY = CH - 2 * CHALO;
constexpr int CTY = 16, CCELL = (CW / 32) * (CH / CTY);   // block 32 x 16 threads, four cells each
constexpr int CN76_ARRAYS = 12;

__global__ void k_cn_reset(const CnuityParams P) {
  for (int k = threadIdx.x; k < 2 * P.kk; k += blockDim.x) P.dpkmin[k] = 999.0;
}

// One block per SM (192 KB of shared memory: six work arrays and the six 2-D metrics of the fluxes).  The
// 3-D operands of layer k+1 are requested into registers as soon as the fluxes of layer k are formed, so
// their latency hides behind the other four stages (per-instruction samples of the first version: 47 % of
// the warps' time waited for these loads, profiles/r02x).
__global__ void __launch_bounds__(32 * CTY, 1) k_cn_loop76(const CnuityParams P) {
  extern __shared__ double cn_smem[];
  constexpr int N = CW * CH;
  double* D = cn_smem;          // dpo, then the low-order dp
  double* U3 = D + N;           // running sum of the old thicknesses above the layer
  double* A = U3 + N;           // uflux, then util1
  double* B = A + N;            // vflux, then util2
  double* E = B + N;            // uflux2, then the clipped flux
  double* F = E + N;            // vflux2, then the clipped flux
  double* SCUY = F + N; double* UBAV = SCUY + N; double* DEPU = UBAV + N;
  double* SCVX = DEPU + N; double* VBAV = SCVX + N; double* DEPV = VBAV + N;
  const int c0 = blockIdx.x * CUX - CHALO, r0 = blockIdx.y * CUY - CHALO;
  const double epsil = 1.0e-11;   // mod_cb_arrays.F90:853
  // this thread's cells of the window: cell e sits at window column threadIdx.x + 32*(e&1), row threadIdx.y +
  // CTY*(e>>1); its flags share a word with its mask byte
  const long qb = (long)(r0 + (int)threadIdx.y) * P.pitch + (c0 + (int)threadIdx.x);
  const long qrow = (long)CTY * P.pitch;
#define CN_Q(e) (qb + 32 * ((e) & 1) + ((e) >> 1) * qrow)
#define CN_S(e) ((threadIdx.y + CTY * ((e) >> 1)) * CW + threadIdx.x + 32 * ((e) & 1))
  double sci[CCELL], un[CCELL], vn[CCELL], pk[CCELL];
  unsigned flags[CCELL];
  // stage predicates of a cell, formed once: flux faces, low-order cell, ratio cell, limiter faces, updated cell,
  // uflx / vflx written, row counts for dpkmin
  enum { VALID = 0x100, OWN = 0x200, FXF = 0x400, FYF = 0x800, SEA4 = 0x1000, LOW = 0x2000, RAT = 0x4000,
         CLX = 0x8000, CLY = 0x10000, UPD = 0x20000, WRU = 0x40000, WRV = 0x80000, JIN = 0x100000 };
#pragma unroll
  for (int e = 0; e < CCELL; ++e) {
    const int sx = threadIdx.x + 32 * (e & 1), sy = threadIdx.y + CTY * (e >> 1), s = CN_S(e);
    const int c = c0 + sx, r = r0 + sy;
    const bool valid = c >= 0 && c < P.pitch && r >= 0 && r < P.nrows;
    const long q = CN_Q(e);
    sci[e] = valid ? P.scp2i[q] : 0.0;
    un[e] = vn[e] = pk[e] = 0.0;
    unsigned f = valid ? (VALID | P.mask[q]) : 0u;
    const bool own = valid && sx >= CHALO && sx < CW - CHALO && sy >= CHALO && sy < CH - CHALO;
    if (own) f |= OWN;
    if (valid) {
      const bool m5 = in_margin(P, c, r, 5), m4 = in_margin(P, c, r, 4), m3 = in_margin(P, c, r, 3), m2 = in_margin(P, c, r, 2);
      const bool inner = sx >= 1 && sx < CW - 1 && sy >= 1 && sy < CH - 1;
      if (m5 && (f & M_IU) && sx >= 1) f |= FXF;
      if (m5 && (f & M_IV) && sy >= 1) f |= FYF;
      if (m4 && (f & M_IP)) f |= SEA4;
      if (m4 && (f & M_IP) && sx < CW - 1 && sy < CH - 1) f |= LOW;
      if (m4 && (f & M_IP) && inner) f |= RAT;
      if (m3 && (f & M_IU) && sx >= 2) f |= CLX;
      if (m3 && (f & M_IV) && sy >= 2) f |= CLY;
      if (own && m2 && (f & M_IP)) f |= UPD;
      if (own && m5 && (f & M_IU)) f |= WRU;
      if (own && m5 && (f & M_IV)) f |= WRV;
      const int j = r + 1 - P.nbdy;
      if (j >= 1 && j <= P.jj) f |= JIN;
    }
    flags[e] = f;
    U3[s] = 0.0;
    SCUY[s] = valid ? P.scuy[q] : 0.0; UBAV[s] = valid ? P.ubavg_m[q] : 0.0; DEPU[s] = valid ? P.depthu[q] : 0.0;
    SCVX[s] = valid ? P.scvx[q] : 0.0; VBAV[s] = valid ? P.vbavg_m[q] : 0.0; DEPV[s] = valid ? P.depthv[q] : 0.0;
    if ((f & OWN)) { P.dpmold[q] = P.dpmixl_n[q]; P.p[q] = 0.0; }
  }
  // the 3-D operands of the layer at hand
  double od[CCELL], ou[CCELL], ov[CCELL], odu[CCELL], odv[CCELL];
#pragma unroll
  for (int e = 0; e < CCELL; ++e) {
    const bool v = flags[e] & VALID;
    const long qk = CN_Q(e);
    od[e] = v ? P.dp_n[qk] : 0.0; ou[e] = v ? P.u_m[qk] : 0.0; ov[e] = v ? P.v_m[qk] : 0.0;
    odu[e] = v ? P.dpu_m[qk] : 0.0; odv[e] = v ? P.dpv_m[qk] : 0.0;
  }
  for (int k = 0; k < P.kk; ++k) {
    const long ko = (long)k * P.slab;
    // ---- the old thickness of the window; dpo(:,:,k,n) = dp(:,:,k,n) (:116-156)
    double fu_own[CCELL], fv_own[CCELL];
#pragma unroll
    for (int e = 0; e < CCELL; ++e) {
      D[CN_S(e)] = od[e];
      if (flags[e] & OWN) P.dpo_n[CN_Q(e) + ko] = od[e];
    }
    __syncthreads();
    // ---- low-order and antidiffusive fluxes at the west and south face of every cell (:236-283)
#pragma unroll
    for (int e = 0; e < CCELL; ++e) {
      const int s = CN_S(e);
      double fu = 0.0, fu2 = 0.0, fv = 0.0, fv2 = 0.0;
      if (flags[e] & FXF) {   // the donor cell by the sign of the transport
        const double utotm = (ou[e] + UBAV[s]) * SCUY[s];
        const int sd = (utotm >= 0.0) ? s - 1 : s;
        const double qq = cmin(D[sd], cmax(0.0, DEPU[s] - U3[sd]));
        fu = utotm * qq;
        fu2 = utotm * odu[e] - fu;
      }
      if (flags[e] & FYF) {
        const double vtotm = (ov[e] + VBAV[s]) * SCVX[s];
        const int sd = (vtotm >= 0.0) ? s - CW : s;
        const double qq = cmin(D[sd], cmax(0.0, DEPV[s] - U3[sd]));
        fv = vtotm * qq;
        fv2 = vtotm * odv[e] - fv;
      }
      A[s] = fu; E[s] = fu2; B[s] = fv; F[s] = fv2;
      fu_own[e] = fu; fv_own[e] = fv;
    }
    // the operands of the next layer: in flight while this one is finished
    if (k + 1 < P.kk) {
#pragma unroll
      for (int e = 0; e < CCELL; ++e) {
        const bool v = flags[e] & VALID;
        const long qk = CN_Q(e) + ko + P.slab;
        od[e] = v ? P.dp_n[qk] : 0.0; ou[e] = v ? P.u_m[qk] : 0.0; ov[e] = v ? P.v_m[qk] : 0.0;
        odu[e] = v ? P.dpu_m[qk] : 0.0; odv[e] = v ? P.dpv_m[qk] : 0.0;
      }
    }
    __syncthreads();
    // ---- loop 19: the low-order step (:293-311); util3 moves on to the next layer (:300)
#pragma unroll
    for (int e = 0; e < CCELL; ++e) {
      const int s = CN_S(e);
      const double d0 = D[s];
      if (flags[e] & SEA4) U3[s] = U3[s] + d0;
      if (flags[e] & LOW) D[s] = d0 - ((A[s + 1] - A[s]) + (B[s + CW] - B[s])) * P.delt1 * sci[e];
    }
    __syncthreads();
    // ---- util1, util2 (:378-400) into A, B
#pragma unroll
    for (int e = 0; e < CCELL; ++e) {
      const int s = CN_S(e);
      if (flags[e] & RAT) {
        const unsigned m = flags[e];
        const double d0 = D[s];
        const double d1 = (m & M_PW) ? D[s - 1] : d0, d2 = (m & M_PE) ? D[s + 1] : d0;
        const double d3 = (m & M_PS) ? D[s - CW] : d0, d4 = (m & M_PN) ? D[s + CW] : d0;
        double u1 = cmax(cmax(cmax(cmax(d0, d1), d2), d3), d4);
        double u2 = cmax(0.0, cmin(cmin(cmin(cmin(d0, d1), d2), d3), d4));
        const double a = E[s], b = E[s + 1], ee = F[s], f = F[s + CW];
        // (zero dividends - a cell at a local extremum - are common: div_rn takes them on its fast path)
        u1 = div_rn(u1 - d0, ((cmax(0.0, a) - cmin(0.0, b)) + (cmax(0.0, ee) - cmin(0.0, f)) + epsil) * P.delt1 * sci[e]);
        u2 = div_rn(u2 - d0, ((cmin(0.0, a) - cmax(0.0, b)) + (cmin(0.0, ee) - cmax(0.0, f)) - epsil) * P.delt1 * sci[e]);
        A[s] = u1; B[s] = u2;
      }
    }
    __syncthreads();
    // ---- the limiter (:414-441): E, F := clipped fluxes; the clipped-off part joins utotn, vtotn in k order
#pragma unroll
    for (int e = 0; e < CCELL; ++e) {
      const int s = CN_S(e);
      const long qk = CN_Q(e) + ko;
      double fcu = 0.0, fcv = 0.0;
      const bool fx = flags[e] & CLX, fy = flags[e] & CLY;
      if (fx) {
        const double f2 = E[s];
        const bool pos = f2 >= 0.0;
        const double clip = cmin(cmin(1.0, pos ? A[s] : B[s]), pos ? B[s - 1] : A[s - 1]);
        un[e] = un[e] + f2 * (1.0 - clip);
        fcu = f2 * clip;
      }
      if (fy) {
        const double f2 = F[s];
        const bool pos = f2 >= 0.0;
        const double clip = cmin(cmin(1.0, pos ? A[s] : B[s]), pos ? B[s - CW] : A[s - CW]);
        vn[e] = vn[e] + f2 * (1.0 - clip);
        fcv = f2 * clip;
      }
      E[s] = fcu; F[s] = fcv;
      if (flags[e] & WRU) P.uflx[qk] = fx ? fu_own[e] + fcu : fu_own[e];
      if (flags[e] & WRV) P.vflx[qk] = fy ? fv_own[e] + fcv : fv_own[e];
    }
    __syncthreads();
    // ---- loop 15: dp with the clipped fluxes (:449-469), into the scratch field; p(k+1)
    double seen = 999.0;
#pragma unroll
    for (int e = 0; e < CCELL; ++e) {
      if (!(flags[e] & OWN)) continue;
      const int s = CN_S(e);
      double d = D[s];
      if (flags[e] & UPD) {
        d = d - ((E[s + 1] - E[s]) + (F[s + CW] - F[s])) * P.delt1 * sci[e];
        pk[e] = pk[e] + d;
        if (flags[e] & JIN) seen = cmin(seen, d);
      }
      P.dnew[CN_Q(e) + ko] = d;
    }
    layer_min(&P.dpkmin[P.kk + k], seen);
    __syncthreads();
  }
#pragma unroll
  for (int e = 0; e < CCELL; ++e)
    if (flags[e] & OWN) {
      P.utotn[CN_Q(e)] = un[e]; P.vtotn[CN_Q(e)] = vn[e];
      if (flags[e] & UPD) P.p[CN_Q(e) + (long)P.kk * P.slab] = pk[e];
    }
}
#undef CN_Q
#undef CN_S

// ---- loop 77 (:588-683), bottom-pressure restoring (:716-733), cumulative fluxes (:1326-1350): columns ---------
__global__ void __launch_bounds__(256) k_cn_loop77(const CnuityParams P) {
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
  const bool inside = c < P.pitch && r < P.nrows;
  const long q = inside ? (long)r * P.pitch + c : 0;
  const unsigned mk = inside ? P.mask[q] : 0u;
  const bool out = mk & M_OUT;                                      // sea cell of 1..ii x 1..jj
  const bool m1 = inside && in_margin(P, c, r, 1);
  const bool fw = m1 && (mk & M_IU), fs = m1 && (mk & M_IV);        // own faces: uflx, vflx are updated there
  const bool fe = out && (P.mask[q + 1] & M_IU), fn = out && (P.mask[q + P.pitch] & M_IV);
  const double* pb = P.p + (long)P.kk * P.slab;                     // p(:,:,kk+1) as loop 76 left it
  // per face: the lost transport, the donor cell by its sign (:600-640), that column's p(kk+1) and its reciprocal
  const double un_w = fw ? P.utotn[q] : 0.0, un_e = fe ? P.utotn[q + 1] : 0.0;
  const double vn_s = fs ? P.vtotn[q] : 0.0, vn_n = fn ? P.vtotn[q + P.pitch] : 0.0;
  const long ow = (un_w >= 0.0) ? -1 : 0, oe = (un_e >= 0.0) ? 0 : 1;
  const long os = (vn_s >= 0.0) ? -(long)P.pitch : 0, on = (vn_n >= 0.0) ? 0 : (long)P.pitch;
  const double pw = fw ? pb[q + ow] : 1.0, pe = fe ? pb[q + oe] : 1.0;
  const double ps = fs ? pb[q + os] : 1.0, pn = fn ? pb[q + on] : 1.0;
  const double yw = rcp_nr(pw), ye = rcp_nr(pe), ys = rcp_nr(ps), yn = rcp_nr(pn);
  const double sci = inside ? P.scp2i[q] : 0.0;
  const int i = c + 1 - P.nbdy, j = r + 1 - P.nbdy;
  const bool col = inside && i >= 1 && i <= P.ii && j >= 1 && j <= P.jj;
  double psum = 0.0;
  for (int k = 0; k < P.kk; ++k) {
    const long qk = q + (long)k * P.slab;
    double seen = 999.0;
    if (inside) {
      const double dc = P.dnew[qk];
      double d = dc;
      double fuw = 0.0, fue = 0.0, fvs = 0.0, fvn = 0.0;
      if (fw) { fuw = un_w * div_y(P.dnew[qk + ow], pw, yw); P.uflx[qk] = P.uflx[qk] + fuw; }
      if (fs) { fvs = vn_s * div_y(P.dnew[qk + os], ps, ys); P.vflx[qk] = P.vflx[qk] + fvs; }
      if (out) {
        if (fe) fue = un_e * div_y(P.dnew[qk + oe], pe, ye);
        if (fn) fvn = vn_n * div_y(P.dnew[qk + on], pn, yn);
        d = dc - ((fue - fuw) + (fvn - fvs)) * P.delt1 * sci;
        seen = d;
      }
      P.dp_n[qk] = d;
      if (col && (mk & M_IP)) psum = psum + d;
    }
    layer_min(&P.dpkmin[k], seen);
  }
  if (!col) return;
  if (mk & M_IP) {
    const double qq = P.pbot[q] / psum;
    double pk = 0.0;
    for (int k = 0; k < P.kk; ++k) {
      const long qk = q + (long)k * P.slab;
      const double d = P.dp_n[qk] * qq;
      P.dp_n[qk] = d;
      pk = pk + d;
      P.p[qk + P.slab] = pk;
      if (P.dpav && !P.defer_av) P.dpav[qk] = P.dpav[qk] + d;
    }
    if (P.isopyc) P.dpmixl_n[q] = P.dp_n[q];
  }
  if (P.defer_av) return;
  if (P.uflxav && (mk & M_IU))
    for (int k = 0; k < P.kk; ++k) { const long qk = q + (long)k * P.slab; P.uflxav[qk] = P.uflxav[qk] + P.uflx[qk]; }
  if (P.vflxav && (mk & M_IV))
    for (int k = 0; k < P.kk; ++k) { const long qk = q + (long)k * P.slab; P.vflxav[qk] = P.vflxav[qk] + P.vflx[qk]; }
}


// =====================================================================================================
// Biharmonic (:745-963) and Laplacian (:973-1124) thickness diffusion - literally, interface depth diffusion.
// The interfaces p(:,:,2..kk) are moved one after the other (the biharmonic form limits the flux of an
// interface against the one of the interface before and walks downward or upward in alternate steps), so the
// reference's three sweeps per interface stay three kernels per interface: coupling between interfaces rules
// out the all-layers form, and the halo each sweep consumes rules out one kernel per column.  (Forming util1,
// util2 inside the flux kernel and completing one layer of uflx, vflx per interface - two kernels, 22 instead
// of 31 passes over a slab per interface - was measured no faster: 74.8 against 73.4 ms per call, r04a.)
// =====================================================================================================
__global__ void __launch_bounds__(256) k_thk_init(const CnuityParams P, int iflip) {
  CN_CELL;
  if (!inside) return;
  P.fu[q] = 0.0; P.fv[q] = 0.0; P.t1[q] = 0.0; P.t2[q] = 0.0;
  P.pold[q] = (iflip == 1 && (P.mask[q] & M_IP)) ? P.p[q + (long)P.kk * P.slab] : 0.0;
}

// util1, util2 of interface k (1-based), margin 5 (:796-833)
__global__ void __launch_bounds__(256) k_thk_util(const CnuityParams P, int k) {
  CN_CELL;
  if (!inside || !in_margin(P, c, r, 5)) return;
  const unsigned mk = P.mask[q];
  if (!(mk & M_IP)) return;
  const double onecm = 9806.0 * 0.01;   // mod_cb_arrays.F90:850
  const double* pk = P.p + (long)(k - 1) * P.slab;
  const double* dk = P.dp_n + (long)(k - 1) * P.slab;
  const double* dm = P.dp_n + (long)(k - 2) * P.slab;
  double u1 = 0.0, u2 = 0.0;
  if (!(cmin(dk[q], dm[q]) < onecm)) {
    // bigrid.F90:343-372: i-1 if sea; else i+1 if sea; otherwise i
    const long ia = (mk & M_PW) ? q - 1 : ((mk & M_PE) ? q + 1 : q);
    const long ib = (mk & M_PE) ? q + 1 : ((mk & M_PW) ? q - 1 : q);
    const long ja = (mk & M_PS) ? q - P.pitch : ((mk & M_PN) ? q + P.pitch : q);
    const long jb = (mk & M_PN) ? q + P.pitch : ((mk & M_PS) ? q - P.pitch : q);
    u1 = pk[q] - .5 * (pk[ia] + pk[ib]);
    u2 = pk[q] - .5 * (pk[ja] + pk[jb]);
    if (u1 > 0.0) { if (cmin(dk[ia], dk[ib]) < onecm) u1 = 0.0; }
    else          { if (cmin(dm[ia], dm[ib]) < onecm) u1 = 0.0; }
    if (u2 > 0.0) { if (cmin(dk[ja], dk[jb]) < onecm) u2 = 0.0; }
    else          { if (cmin(dm[ja], dm[jb]) < onecm) u2 = 0.0; }
  }
  P.t1[q] = u1; P.t2[q] = u2;
}

// the limited fluxes of interface k at the u and v points, margin 4 (:835-906, :1029-1061); uflx, vflx of the
// layers above and below take them up
template <bool BIH>
__global__ void __launch_bounds__(256) k_thk_flux(const CnuityParams P, int k, int iflip, double dtinv) {
  CN_CELL;
  if (!inside || !in_margin(P, c, r, 4)) return;
  const unsigned mk = P.mask[q];
  const double* pk = P.p + (long)(k - 1) * P.slab;
  const double* pb = P.p + (long)P.kk * P.slab;
  const long ka = q + (long)(k - 2) * P.slab, kb = q + (long)(k - 1) * P.slab;
#define THK_FACE(LO, FLUX, COEF, UTIL, OUT)                                                        \
  {                                                                                                \
    const long w = (LO);                                                                           \
    double flxhi = .25 * (pb[q] - pk[q]) * P.scp2[q];                                              \
    double flxlo = -.25 * (pb[w] - pk[w]) * P.scp2[w];                                             \
    double want;                                                                                   \
    if (BIH) {                                                                                     \
      if (iflip == 0) { /* downward k loop */                                                      \
        flxhi = cmin(flxhi, FLUX[q] + .25 * (pk[w] - P.pold[w]) * P.scp2[w]);                      \
        flxlo = cmax(flxlo, FLUX[q] - .25 * (pk[q] - P.pold[q]) * P.scp2[q]);                      \
      } else {          /* upward k loop */                                                        \
        flxhi = cmin(flxhi, FLUX[q] + .25 * (P.pold[q] - pk[q]) * P.scp2[q]);                      \
        flxlo = cmax(flxlo, FLUX[q] - .25 * (P.pold[w] - pk[w]) * P.scp2[w]);                      \
      }                                                                                            \
      want = (P.delt1 * COEF[q]) * (UTIL[w] - UTIL[q]);                                            \
    } else {                                                                                       \
      want = (P.delt1 * COEF[q]) * (pk[w] - pk[q]);                                                \
    }                                                                                              \
    const double f = cmin(flxhi, cmax(flxlo, want));                                               \
    FLUX[q] = f;                                                                                   \
    OUT[ka] = OUT[ka] + f * dtinv;                                                                 \
    OUT[kb] = OUT[kb] - f * dtinv;                                                                 \
  }
  if (mk & M_IU) THK_FACE(q - 1, P.fu, P.thku, P.t1, P.uflx)
  if (mk & M_IV) THK_FACE(q - P.pitch, P.fv, P.thkv, P.t2, P.vflx)
#undef THK_FACE
}

// pold = p(k); p(k) moves with the divergence of the limited fluxes, margin 4 (:908-922, :1063-1078)
__global__ void __launch_bounds__(256) k_thk_cell(const CnuityParams P, int k) {
  CN_CELL;
  if (!inside || !in_margin(P, c, r, 4) || !(P.mask[q] & M_IP)) return;
  double* pk = P.p + (long)(k - 1) * P.slab;
  const double v = pk[q];
  P.pold[q] = v;
  pk[q] = v - ((P.fu[q + 1] - P.fu[q]) + (P.fv[q + P.pitch] - P.fv[q])) * P.scp2i[q];
}

// interfaces may not cross; dp(:,:,:,n) from them, margin 4 (:937-962, :1092-1114); then the cumulative fluxes
// of :1326-1350 that the diffusion postponed
__global__ void __launch_bounds__(256) k_thk_final(const CnuityParams P) {
  CN_CELL;
  if (!inside) return;
  const unsigned mk = P.mask[q];
  if (in_margin(P, c, r, 4) && (mk & M_IP)) {
    double pa = P.p[q];
    for (int k = 1; k <= P.kk; ++k) {
      const long qk = q + (long)k * P.slab;
      double pbk = P.p[qk];
      if (pbk < pa) { pbk = pa; P.p[qk] = pbk; }
      P.dp_n[qk - P.slab] = pbk - pa;
      pa = pbk;
    }
    if (P.isopyc) P.dpmixl_n[q] = P.dp_n[q];
  }
  const int i = c + 1 - P.nbdy, j = r + 1 - P.nbdy;
  if (i < 1 || i > P.ii || j < 1 || j > P.jj) return;
  for (int k = 0; k < P.kk; ++k) {
    const long qk = q + (long)k * P.slab;
    if (P.uflxav && (mk & M_IU)) P.uflxav[qk] = P.uflxav[qk] + P.uflx[qk];
    if (P.vflxav && (mk & M_IV)) P.vflxav[qk] = P.vflxav[qk] + P.vflx[qk];
    if (P.dpav && (mk & M_IP)) P.dpav[qk] = P.dpav[qk] + P.dp_n[qk];
  }
}


// =====================================================================================================
// hybrid .and. mxlkta (:1144-1324): dpmixl(:,:,n) follows the vertical excursion of the coordinates immediately
// above and below the mixed-layer base (found in the OLD thicknesses), then is diffused like an interface.
// =====================================================================================================
__global__ void __launch_bounds__(256) k_mxl_vert(const CnuityParams P) {
  CN_CELL;
  if (!inside || !in_margin(P, c, r, 4) || !(P.mask[q] & M_IP)) return;
  const double onemm = 9806.0 * 0.001;   // mod_cb_arrays.F90:851
  double above = 0., below = 0., mx = P.dpmixl_n[q];
  for (int k = 0; k < P.kk; ++k) {
    const long qk = q + (long)k * P.slab;
    const double dpok = P.dpo_n[qk];
    above = below;
    below = below + dpok;
    if (below >= mx && above < mx) {
      const double dpup = P.p[qk] - above;
      const double dpdn = P.p[qk + P.slab] - below;
      const double qq = (below - mx) / cmax(onemm, dpok);
      mx = mx + (dpdn + qq * (dpup - dpdn));
    }
  }
  P.dpmixl_n[q] = mx;
}

// biharmonic: util1, util2 of dpmixl, margin 2 (:1206-1225)
__global__ void __launch_bounds__(256) k_mxl_util(const CnuityParams P) {
  CN_CELL;
  if (!inside) return;
  const unsigned mk = P.mask[q];
  double u1 = 0.0, u2 = 0.0;
  if (in_margin(P, c, r, 2) && (mk & M_IP)) {
    const double* d = P.dpmixl_n;
    const long ia = (mk & M_PW) ? q - 1 : ((mk & M_PE) ? q + 1 : q);
    const long ib = (mk & M_PE) ? q + 1 : ((mk & M_PW) ? q - 1 : q);
    const long ja = (mk & M_PS) ? q - P.pitch : ((mk & M_PN) ? q + P.pitch : q);
    const long jb = (mk & M_PN) ? q + P.pitch : ((mk & M_PS) ? q - P.pitch : q);
    u1 = d[q] - 0.5 * (d[ia] + d[ib]);
    u2 = d[q] - 0.5 * (d[ja] + d[jb]);
  }
  P.t1[q] = u1; P.t2[q] = u2;
}

// fluxes at the u and v points: biharmonic from util1, util2 (margin 1, :1227-1243), Laplacian from dpmixl
// (margin 2, :1285-1301); zero everywhere else (:1190-1203)
template <bool BIH>
__global__ void __launch_bounds__(256) k_mxl_flux(const CnuityParams P) {
  CN_CELL;
  if (!inside) return;
  const unsigned mk = P.mask[q];
  const bool in = in_margin(P, c, r, BIH ? 1 : 2);
  double fu = 0.0, fv = 0.0;
  if (in && (mk & M_IU)) fu = (P.delt1 * P.thku[q]) * (BIH ? P.t1[q - 1] - P.t1[q] : P.dpmixl_n[q - 1] - P.dpmixl_n[q]);
  if (in && (mk & M_IV)) fv = (P.delt1 * P.thkv[q]) * (BIH ? P.t2[q - P.pitch] - P.t2[q] : P.dpmixl_n[q - P.pitch] - P.dpmixl_n[q]);
  P.fu[q] = fu; P.fv[q] = fv;
}

__global__ void __launch_bounds__(256) k_mxl_update(const CnuityParams P) {   // margin 0 (:1245-1262, :1303-1320)
  CN_CELL;
  if (!inside || !(P.mask[q] & M_OUT)) return;
  P.dpmixl_n[q] = P.dpmixl_n[q] - ((P.fu[q + 1] - P.fu[q]) + (P.fv[q + P.pitch] - P.fv[q])) * P.scp2i[q];
}

}  // namespace

// stage 0: everything up to the exchange of dp(:,:,:,n) (:1400); stage 1: the Robert-Asselin filter
int launch_cnuity(int stage, const CnuityParams& P, cudaStream_t st) {
  const dim3 block(32, 8), g2((P.pitch + 31) / 32, (P.nrows + 7) / 8), g3(g2.x, g2.y, P.kk);
  if (stage == 1) {
    k_cn_asselin<<<g3, block, 0, st>>>(P);
    return (int)cudaGetLastError();
  }
  static bool attr_set = false;
  const int bytes = CN76_ARRAYS * CW * CH * (int)sizeof(double);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_cn_loop76, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  k_cn_reset<<<1, 256, 0, st>>>(P);
  const dim3 gt((P.pitch + CUX - 1) / CUX, (P.nrows + CUY - 1) / CUY);
  k_cn_loop76<<<gt, dim3(32, CTY), bytes, st>>>(P);
  k_cn_loop77<<<g2, block, 0, st>>>(P);
  return (int)cudaGetLastError();
}

int launch_cnuity_thkdf(const CnuityParams& P, int bih, int nstep, cudaStream_t st) {
  const dim3 block(32, 8), g2((P.pitch + 31) / 32, (P.nrows + 7) / 8);
  const int iflip = nstep % 2;        // :766
  const double dtinv = 1. / P.delt1;  // :765
  int n = 0;
  k_thk_init<<<g2, block, 0, st>>>(P, iflip); ++n;
  // :790 alternate between upward and downward direction in the k loop (biharmonic); :1013 k = 2,kk (Laplacian)
  const int k0 = bih ? 2 * (1 - iflip) + P.kk * iflip : 2, k1 = bih ? P.kk * (1 - iflip) + 2 * iflip : P.kk;
  const int kstep = bih ? 1 - 2 * iflip : 1;
  for (int k = k0; kstep > 0 ? k <= k1 : k >= k1; k += kstep) {
    if (bih) {
      k_thk_util<<<g2, block, 0, st>>>(P, k); ++n;
      k_thk_flux<true><<<g2, block, 0, st>>>(P, k, iflip, dtinv); ++n;
    } else {
      k_thk_flux<false><<<g2, block, 0, st>>>(P, k, iflip, dtinv); ++n;
    }
    k_thk_cell<<<g2, block, 0, st>>>(P, k); ++n;
  }
  k_thk_final<<<g2, block, 0, st>>>(P); ++n;
  return cudaGetLastError() == cudaSuccess ? n : -1;
}

// mode: 0 no diffusion of dpmixl (thkdf2 = thkdf4 = 0), 1 biharmonic, 2 Laplacian; returns launches (< 0: error)
int launch_cnuity_mxlkta(const CnuityParams& P, int mode, cudaStream_t st) {
  const dim3 block(32, 8), g2((P.pitch + 31) / 32, (P.nrows + 7) / 8);
  int n = 0;
  k_mxl_vert<<<g2, block, 0, st>>>(P); ++n;
  if (mode == 1) {
    k_mxl_util<<<g2, block, 0, st>>>(P); ++n;
    k_mxl_flux<true><<<g2, block, 0, st>>>(P); ++n;
  } else if (mode == 2) {
    k_mxl_flux<false><<<g2, block, 0, st>>>(P); ++n;
  }
  if (mode) { k_mxl_update<<<g2, block, 0, st>>>(P); ++n; }
  return cudaGetLastError() == cudaSuccess ? n : -1;
}

}  // namespace tsadvc
