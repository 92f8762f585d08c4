// Diffusion of the thermodynamic variables and tracers after advection, and the sweep that
// re-derives the non-independent variable (mod_tsadvc.F90:2138-2230: tsdff_2x :2262-2345,
// tsdff_1x :2347-2492, EOS sweep :2199-2229), fused: the reference runs, per layer and field
// pair, one sweep that stores the face fluxes uflux/vflux(/uflux2/vflux2), one sweep that
// applies their divergence, and at the end one sweep over all layers for the equation of
// state.  Here one thread owns one (i,j) column and walks the layers; per layer it forms the
// four face factors once (they depend on dp only and are shared by every field of the layer),
// diffuses every field of the launch and - for the T/S/th3d launch - applies the equation of
// state to the fresh values before the single store.
//
// HBM bound: per layer-cell it reads dp(n) and each field once (neighbours come from L1/L2)
// and writes each field once: T,S,th3d launch 7 x 8 B = 56 B, tracer launch 8 + 16 B/tracer.
//
// Arithmetic is the Fortran's, expression by expression (-fmad=false, IEEE division):
//   factor_u = temdf2*aspux(i,j)*scuy(i,j)*harmonc(dp(i-1)*onetamas(i-1), dp(i)*onetamas(i))
//   uflux(i) = factor_u*(fld(i-1)-fld(i))                       at iu points, else 0.0
//   factor   = -delt1/(scp2*max(dp*onetamas, eps_har))
//   fld     += ((uflux(i+1)-uflux(i))+(vflux(j+1)-vflux(j)))*factor
// Land faces carry 0.0: uflux/vflux are zeroed at mod_tsadvc.F90:1812-1813, uflux2/vflux2 on
// the faces bounding every sea segment by geopar.F90:826-843, and tsdff only writes iu/iv
// points.  The output goes to the ping-pong buffer (neighbours read the old values).
#include <cuda_runtime.h>

#include <cstdlib>

#include "eos.cuh"
#include "march_common.cuh"   // div_rn: IEEE round-to-nearest a/b, short sequence + exact fallback
#include "tsadvc_dev.h"
#include "tsadvc_launch.h"

namespace tsadvc {

namespace {

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

// harmonc(aa,bb) = harmonz(max(aa,0),max(bb,0)); harmonz(a,b) = 2.0*a*b/max((a+b),2.0*eps_har)
// (mod_tsadvc.F90:1770-1771)
__device__ __forceinline__ double harmonc(double aa, double bb) {
  const double eps_har = 1.0e-20;   // :1763
  const double a = dmax(aa, 0.0), b = dmax(bb, 0.0);
  return div_rn(2.0 * a * b, dmax((a + b), 2.0 * eps_har));
}

// One thread owns one (i,j) column and walks the layers: everything that does not depend on k
// (masks, temdf2*aspux*scuy, temdf2*aspvy*scvx, scp2, oneta at the five points) is loaded and
// formed once and stays in registers, so per layer the thread reads dp and the fields only.
template <bool EOS, int UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB) k_tsdff(const DiffParams P) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r = blockIdx.y * 8 + threadIdx.y;
  if (c >= P.pitch || r >= P.nrows) return;
  const long q = (long)r * P.pitch + c;
  const unsigned m = P.mask[q];
  if (!(m & M_OUT)) {   // land, halo, pad: the ping-pong slab keeps the old value
    for (int k0 = 0; k0 < P.kk; ++k0) {
      const long qk = q + (long)k0 * P.slab;
#pragma unroll 1
      for (int f = 0; f < P.nf; ++f) P.f[f].out[qk] = P.f[f].in[qk];
    }
    return;
  }
  // faces of this cell: west/south are its own iu/iv, east/north those of the neighbours
  const bool fw = m & M_IU, fs = m & M_IV;
  const bool fe = P.mask[q + 1] & M_IU, fn = P.mask[q + P.pitch] & M_IV;
  const long qw = q - 1, qe = q + 1, qs = q - P.pitch, qn = q + P.pitch;
  // temdf2*aspux(i,j)*scuy(i,j) and temdf2*aspvy(i,j)*scvx(i,j): left to right as written
  const double aw = P.temdf2 * P.aspux[q] * P.scuy[q];
  const double ae = P.temdf2 * P.aspux[qe] * P.scuy[qe];
  const double as = P.temdf2 * P.aspvy[q] * P.scvx[q];
  const double an = P.temdf2 * P.aspvy[qn] * P.scvx[qn];
  const double oc = P.oneta[q];
  const double ow = P.oneta[qw], oe = P.oneta[qe], os = P.oneta[qs], on = P.oneta[qn];
  const double scp2 = P.scp2[q];
  // every neighbour of a cell tsadvc writes exists in the slab, so all loads are unconditional
  // (issued together, read-only path) and land neighbours are removed by select afterwards
#pragma unroll UNROLL
  for (int k0 = 0; k0 < P.kk; ++k0) {
    const long ko = (long)k0 * P.slab;
    const long qk = q + ko;
    const double* dp = P.dp + ko;
    const double dc = __ldg(dp + q), dw = __ldg(dp + qw), de = __ldg(dp + qe), ds = __ldg(dp + qs),
                 dn = __ldg(dp + qn);
    const double hc = dc * oc;
    const double gw = fw ? aw * harmonc(dw * ow, hc) : 0.0;
    const double ge = fe ? ae * harmonc(hc, de * oe) : 0.0;
    const double gs = fs ? as * harmonc(ds * os, hc) : 0.0;
    const double gn = fn ? an * harmonc(hc, dn * on) : 0.0;
    const double factor = div_rn(-P.delt1, scp2 * dmax(hc, 1.0e-20));   // :2314-2315
    // one field: the divergence of the four face fluxes applied to the centre value
    auto diffuse = [&](const double* __restrict__ a, double x) {
      const double xw = __ldg(a + qw), xe = __ldg(a + qe), xs = __ldg(a + qs), xn = __ldg(a + qn);
      const double uw = fw ? gw * (xw - x) : 0.0;
      const double ue = fe ? ge * (x - xe) : 0.0;
      const double vs = fs ? gs * (xs - x) : 0.0;
      const double vn = fn ? gn * (x - xn) : 0.0;
      const double util = ((ue - uw) + (vn - vs)) * factor;
      return x + util;
    };
    if (!EOS) {
#pragma unroll 1
      for (int f = 0; f < P.nf; ++f) {
        const double* a = P.f[f].in + ko;
        P.f[f].out[qk] = diffuse(a, __ldg(a + q));
      }
      continue;
    }
    const int k = k0 + 1;
    const bool ldtemp = k <= P.nhybrd && P.temdfc > 0.0;                               // :2170
    const bool ldth3d = (k <= P.nhybrd && P.temdfc < 1.0) || (k == 1 && P.isopyc);     // :2171-2172
    // T/S/th3d launch: f = 0 temp, 1 saln, 2 th3d; which of them are diffused is :2173-2185
    double v[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const double* a = P.f[f].in + ko;
      const double x = __ldg(a + q);
      const bool on_ = f == 1 || (f == 0 ? ldtemp : ldth3d);
      const double y = diffuse(a, x);
      v[f] = on_ ? y : x;
    }
    {   // :2199-2229
      double t = v[0], s = v[1], h = v[2];
      if (ldtemp && ldth3d) {
        const double th3d_t = eos::sig(P.eosc, t, s) - P.thbase;
        h = (1.0 - P.temdfc) * h + P.temdfc * th3d_t;
        t = eos::tofsig(P.eosc, h + P.thbase, s);
      } else if (ldtemp) {
        h = eos::sig(P.eosc, t, s) - P.thbase;
      } else if (ldth3d) {
        t = eos::tofsig(P.eosc, h + P.thbase, s);
      } else {
        h = P.theta[qk];
        t = eos::tofsig(P.eosc, h + P.thbase, s);
      }
      P.f[0].out[qk] = t;
      P.f[1].out[qk] = s;
      P.f[2].out[qk] = h;
    }
  }
}

// isopycnic coordinates, layer 1: lateral smoothing of the mixed-layer mass fluxes
// (mod_tsadvc.F90:1859-1897, margin mbdy-1); faces that are not iu/iv points and everything
// outside the margin keep the 0.0 of :1812-1813
__global__ void __launch_bounds__(256) k_isopyc_smooth(const double* __restrict__ u, const double* __restrict__ v,
                                                         double* __restrict__ us, double* __restrict__ vs,
                                                         const uint8_t* __restrict__ mask, int pitch, int nrows,
                                                         int nbdy, int ii, int jj, int margin) {
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
  if (c >= pitch || r >= nrows) return;
  const long q = (long)r * pitch + c;
  const int i = c + 1 - nbdy, j = r + 1 - nbdy;
  double uo = 0.0, vo = 0.0;
  if (i >= 1 - margin && i <= ii + margin && j >= 1 - margin && j <= jj + margin) {
    const unsigned m = mask[q];
    if (m & M_IV) {
      const double vfa = (mask[q - 1] & M_IV) ? v[q - 1] : v[q];
      const double vfb = (mask[q + 1] & M_IV) ? v[q + 1] : v[q];
      vo = .5 * v[q] + .25 * (vfa + vfb);
    }
    if (m & M_IU) {
      const double ufa = (mask[q - pitch] & M_IU) ? u[q - pitch] : u[q];
      const double ufb = (mask[q + pitch] & M_IU) ? u[q + pitch] : u[q];
      uo = .5 * u[q] + .25 * (ufa + ufb);
    }
  }
  us[q] = uo;
  vs[q] = vo;
}

}  // namespace

int launch_isopyc_smooth(const double* u, const double* v, double* us, double* vs, const uint8_t* mask,
                         int pitch, int nrows, int nbdy, int ii, int jj, int margin, cudaStream_t stream) {
  const dim3 block(32, 8), grid((pitch + 31) / 32, (nrows + 7) / 8);
  k_isopyc_smooth<<<grid, block, 0, stream>>>(u, v, us, vs, mask, pitch, nrows, nbdy, ii, jj, margin);
  return (int)cudaGetLastError();
}

int launch_tsdff(const DiffParams& P, cudaStream_t stream) {
  const dim3 block(32, 8), grid((P.pitch + 31) / 32, (P.nrows + 7) / 8);
  // measured on B200 (profiles/r01q_tsdff_variants.txt): the kernel is load-latency bound, the
  // variant with the most resident warps wins (64 registers, 4 blocks per SM: 15.5 ms at GLBb0.08
  // against 17.1 with unroll 2 at 80 registers and 21.1 with unroll 2 at 103)
  static const int variant = [] { const char* e = getenv("HYCOM_TSADVC_TSDFF_VARIANT"); return e ? atoi(e) : 0; }();
  if (P.eos) {
    if (variant == 5) k_tsdff<true, 1, 5><<<grid, block, 0, stream>>>(P);
    else if (variant == 6) k_tsdff<true, 1, 6><<<grid, block, 0, stream>>>(P);
    else k_tsdff<true, 1, 4><<<grid, block, 0, stream>>>(P);
  } else {
    k_tsdff<false, 1, 4><<<grid, block, 0, stream>>>(P);
  }
  return (int)cudaGetLastError();
}

}  // namespace tsadvc
