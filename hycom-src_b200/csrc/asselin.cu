// Robert-Asselin filter of the scalar fields (mod_asselin.F90), the pointwise consumer of tsadvc's
// output in the time step:
//   asselin_save   :28-82   oneta/onetao of both slots from pbavg/pbot, time level t-1 of every
//                           scalar saved in otemp, osaln, oth3d, otracer (, oq2, oq2l)
//   asselin_filter :84-286  oneta of both slots, then per cell and layer the filter of
//                           oneta*dp*scalar at time level t ("version that exactly conserves
//                           constant salinity", :143), dp(:,:,:,m), and the dependent
//                           thermodynamic variable through the equation of state
// Pointwise and HBM bound: a hybrid T/S layer reads dpo(n), dpo(m), dp(n) and three time levels of
// saln and temp (9 x 8 B) and writes dp(m), saln(m), temp(m), th3d(m) (4 x 8 B) = 104 B per
// layer-cell, +32 B per tracer.  Arithmetic is the Fortran's, expression by expression
// (-fmad=false, IEEE division); everything is updated in place (no neighbour is read).
#include <cuda_runtime.h>

#include "eos.cuh"
#include "march_common.cuh"
#include "tsadvc_dev.h"
#include "tsadvc_launch.h"

namespace tsadvc {

namespace {

__device__ __forceinline__ double amax2(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double amin3(double a, double b, double c) {
  const double ab = a < b ? a : b;
  return ab < c ? ab : c;
}

// oneta(i,j,t) = max(oneta0, 1.0 + pbavg(i,j,t)/pbot(i,j)) on sea points of 1:ii,1:jj, t = n, m
// (:52-53 and :115-116); save additionally copies it into onetao (:54-55)
__global__ void __launch_bounds__(256) k_asselin_oneta(const AsselinParams P, int save) {
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
  if (c >= P.pitch || r >= P.nrows) return;
  const long q = (long)r * P.pitch + c;
  if (!(P.mask[q] & M_OUT)) return;
  const double pb = P.pbot[q];
  const double en = amax2(P.oneta0, 1.0 + P.pbavg_n[q] / pb);
  const double em = amax2(P.oneta0, 1.0 + P.pbavg_m[q] / pb);
  P.oneta_n[q] = en;
  P.oneta_m[q] = em;
  if (save) {
    P.onetao_n[q] = en;
    P.onetao_m[q] = em;
  }
}

// o*(i,j,k) = *(i,j,k,n) on 1:ii,1:jj, land included (:57-75)
__global__ void __launch_bounds__(256) k_asselin_copy(const AsselinParams P) {
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y, k0 = blockIdx.z;
  const int i = c + 1 - P.nbdy, j = r + 1 - P.nbdy;
  if (i < 1 || i > P.ii || j < 1 || j > P.jj) return;
  const long qk = (long)r * P.pitch + c + (long)k0 * P.slab;
  for (int f = 0; f < P.nf; ++f) P.cp[f].o[qk] = P.cp[f].fn[qk];
}

__global__ void __launch_bounds__(256) k_asselin_filter(const AsselinParams P) {
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y, k0 = blockIdx.z;
  if (c >= P.pitch || r >= P.nrows) return;
  const long q = (long)r * P.pitch + c;
  if (!(P.mask[q] & M_OUT)) return;
  const long qk = q + (long)k0 * P.slab;
  const int k = k0 + 1;
  const bool latemp = k <= P.nhybrd && P.advflg == 0;                              // :121
  const bool lath3d = (k <= P.nhybrd && P.advflg == 1) || (k == 1 && P.isopyc);    // :122-123
  const double onezm = 9806.e-20;                                                  // :93
  const double dpold = P.dpo_n[qk] * P.onetao_n[q];                                // :131-133
  const double dpmid = P.dpo_m[qk] * P.onetao_m[q];
  const double dpnew = P.dp_n[qk] * P.oneta_n[q];
  double qq = 0.5 * P.ra2fac * (dpold + dpnew - 2.0 * dpmid);
  const double dpmidn = dpmid + qq;
  P.dp_m[qk] = dpmidn / P.oneta_m[q];                                              // :136
  if (!(dpmidn > onezm)) return;                                                   // :137
  const double qdpmidn = 1.0 / dpmidn;
  // smin + (dpsmid + q)*qdpmidn with the three time levels shifted by their minimum (:144-151)
  auto filt = [&](double o, double fm, double fn) {
    const double smin = amin3(o, fm, fn);
    const double dpsold = dpold * (o - smin), dpsmid = dpmid * (fm - smin), dpsnew = dpnew * (fn - smin);
    const double w = 0.5 * P.ra2fac * (dpsold + dpsnew - 2.0 * dpsmid);
    return smin + (dpsmid + w) * qdpmidn;
  };
  // f[0] = saln, f[1] = temp, f[2] = th3d, then tracers
  const double s = filt(P.f[0].o[qk], P.f[0].fm[qk], P.f[0].fn[qk]);
  P.f[0].fm[qk] = s;
  if (latemp) {            // :169-180
    const double t = filt(P.f[1].o[qk], P.f[1].fm[qk], P.f[1].fn[qk]);
    P.f[1].fm[qk] = t;
    P.f[2].fm[qk] = eos::sig(P.eosc, t, s) - P.thbase;
  } else if (lath3d) {     // :181-192
    const double h = filt(P.f[2].o[qk], P.f[2].fm[qk], P.f[2].fn[qk]);
    P.f[2].fm[qk] = h;
    P.f[1].fm[qk] = eos::tofsig(P.eosc, h + P.thbase, s);
  } else {                 // :193-198 exactly isopycnal layer
    const double h = P.theta[qk];
    P.f[2].fm[qk] = h;
    P.f[1].fm[qk] = eos::tofsig(P.eosc, h + P.thbase, s);
  }
  for (int f = 3; f < P.nf; ++f)   // :199-225
    P.f[f].fm[qk] = filt(P.f[f].o[qk], P.f[f].fm[qk], P.f[f].fn[qk]);
  if (P.q2_o) {            // :226-237 (no minimum shift for q2, q2l)
    for (int w = 0; w < 2; ++w) {
      const double* o = w ? P.q2l_o : P.q2_o;
      double* fm = w ? P.q2l_m : P.q2_m;
      const double* fn = w ? P.q2l_n : P.q2_n;
      const long qa = qk + P.slab;   // layer k of (0:kk+1)
      const double dpsold = dpold * o[qa], dpsmid = dpmid * fm[qa], dpsnew = dpnew * fn[qa];
      qq = 0.5 * P.ra2fac * (dpsold + dpsnew - 2.0 * dpsmid);
      fm[qa] = (dpsmid + qq) * qdpmidn;
    }
  }
}

}  // namespace

// stage 0: oneta (filter), 1: oneta + onetao (save), 2: o* copies (save), 3: the filter
int launch_asselin(int stage, const AsselinParams& P, cudaStream_t stream) {
  const dim3 block(32, 8), g2((P.pitch + 31) / 32, (P.nrows + 7) / 8), g3(g2.x, g2.y, P.kk);
  switch (stage) {
    case 0: k_asselin_oneta<<<g2, block, 0, stream>>>(P, 0); break;
    case 1: k_asselin_oneta<<<g2, block, 0, stream>>>(P, 1); break;
    case 2: k_asselin_copy<<<dim3(g2.x, g2.y, P.kcopy), block, 0, stream>>>(P); break;
    case 3: k_asselin_filter<<<g3, block, 0, stream>>>(P); break;
    default: return -1;
  }
  return (int)cudaGetLastError();
}

}  // namespace tsadvc
