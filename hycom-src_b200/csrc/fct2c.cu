// advem_fct2c (mod_tsadvc.F90:999-1368), the advection scheme tsadvc uses when btrmas
// (:96-97): leapfrog FCT2 whose low-order step is sub-cycled - five iterations with a local
// time step dtloc so that no cell is drained, each ending with xctilr(hloc), xctilr(fldlo) -
// followed by the Zalesak limiter on high-order minus cumulated low-order fluxes.
//
// The five exchanges inside the scheme rule out the single-pass marching kernel: every
// iteration needs its neighbours' fresh hloc/fldlo.  The scheme is therefore a short
// sequence of whole-tile kernels over a batch of layers, all fields of a layer together:
//
//   k_c_init   hloc = fco (tsadvc prolog :1934-1937 with onetamas(:,:,m) = oneta(:,:,n)),
//              fldlo = fld, lcalc = .true.                                    (:1072-1086)
//   per iteration
//   k_c_dtloc  dtloc, margin 5                                                (:1090-1107)
//   k_c_faces  uloc,vloc,ucumdt,vcumdt (per layer: they depend on the mass fluxes only, the
//              reference recomputes them for every field) and flx,fly,flxcum,flycum per
//              field, margin 4                                                (:1109-1158)
//   k_c_cells  hloc, lcalc per layer and fldlo per field, margin 3            (:1160-1184)
//   [xctilr of hloc and fldlo, width 5: halo kernels / the caller's exchange] (:1186-1187)
//   k_c_fax    fax,fay = high-order flux - flxcum/dt2, 0 on land faces        (:1202-1245)
//   k_c_ratio  rp, rm, margin 2                                               (:1258-1306)
//   k_c_final  limited fluxes and the update of fld, margin 0                 (:1311-1361)
//
// Arithmetic: the Fortran's expression by expression (-fmad=false, IEEE division), e.g. the
// un-parenthesised uloc(i+1,j)-uloc(i,j)+vloc(i,j+1)-vloc(i,j) is evaluated left to right.
// Cells and faces outside the margins of the reference sweeps are never consumed (see the
// margin bookkeeping in DESIGN.md); they are given finite stand-ins (0.0) here where the
// reference leaves r_init.
#include <cuda_runtime.h>

#include "march_common.cuh"
#include "tsadvc_dev.h"
#include "tsadvc_launch.h"

namespace tsadvc {

namespace {

__device__ __forceinline__ double dmax2(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin2(double a, double b) { return a < b ? a : b; }

// (c, r) of this thread and whether it lies inside the sweep region 1-margin..ii+margin
struct Cell {
  int c, r;
  long q;
  bool in;
};
__device__ __forceinline__ Cell cell_of(const Fct2cParams& P, int margin) {
  Cell x;
  x.c = blockIdx.x * 32 + threadIdx.x;
  x.r = blockIdx.y * 8 + threadIdx.y;
  x.q = (long)x.r * P.pitch + x.c;
  const int i = x.c + 1 - P.nbdy, j = x.r + 1 - P.nbdy;
  x.in = x.c < P.pitch && x.r < P.nrows && i >= 1 - margin && i <= P.ii + margin && j >= 1 - margin &&
         j <= P.jj + margin;
  return x;
}

__global__ void __launch_bounds__(256) k_c_init(const Fct2cParams P) {
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y, kb = blockIdx.z;
  if (c >= P.pitch || r >= P.nrows) return;
  const long q = (long)r * P.pitch + c;
  const long qs = q + (long)kb * P.slab;             // scratch layer
  const long qk = q + (long)(P.k0 + kb) * P.slab;    // model layer
  const int i = c + 1 - P.nbdy, j = r + 1 - P.nbdy;
  const bool m4 = i >= -3 && i <= P.ii + 4 && j >= -3 && j <= P.jj + 4;
  double h = 0.0;
  if (m4 && (P.mask[q] & M_IP)) {   // :1934-1937, onetamas(:,:,m) = oneta(:,:,n) when btrmas (:1806)
    const double flxdiv = ((P.u[qk + 1] - P.u[qk]) + (P.v[qk + P.pitch] - P.v[qk])) * P.dt2 * P.scp2i[q];
    h = dmax2(P.oneta[q] * P.dp[qk] + flxdiv, 0.0);
  }
  P.hloc[qs] = h;
  P.lcalc[qs] = 1;
  for (int f = 0; f < P.nf; ++f)
    if (P.k0 + kb < P.nlay[f])
      P.fldlo[((long)f * P.nb + kb) * P.slab + q] = P.fld[f][qk];
}

__global__ void __launch_bounds__(256) k_c_dtloc(const Fct2cParams P) {
  const Cell x = cell_of(P, 5);
  const int kb = blockIdx.z;
  if (!x.in || !(P.mask[x.q] & M_IP)) return;
  const long qs = x.q + (long)kb * P.slab, qk = x.q + (long)(P.k0 + kb) * P.slab;
  // the east/north faces of a margin-5 cell may lie on the 6th halo line; such a dtloc is
  // never consumed (faces are formed at margin 4 from dtloc(i-1), dtloc(i))
  const bool e_ok = x.c + 1 < P.ncols, n_ok = x.r + 1 < P.nrows;
  const double ue = e_ok ? P.u[qk + 1] : 0.0, vn = n_ok ? P.v[qk + P.pitch] : 0.0;
  const double q = dmax2(ue, 0.0) - dmin2(P.u[qk], 0.0) + dmax2(vn, 0.0) - dmin2(P.v[qk], 0.0);
  double d = P.dt2;
  if (q > 0.0) d = dmin2(P.dt2, P.hloc[qs] / (q * P.scp2i[x.q]));
  P.dtloc[qs] = d;
}

__global__ void __launch_bounds__(256) k_c_faces(const Fct2cParams P) {
  const Cell x = cell_of(P, 4);
  const int kb = blockIdx.z;
  if (!x.in) return;
  const unsigned m = P.mask[x.q];
  const long qs = x.q + (long)kb * P.slab, qk = x.q + (long)(P.k0 + kb) * P.slab;
  const double dt2 = P.dt2;
  if (m & M_IU) {
    const double uc = P.ucum[qs];
    if (uc != dt2) {
      const double u = P.u[qk];
      const long up = (u >= 0) ? -1 : 0;                       // upwind cell
      const double dtu = dmin2(dt2 - uc, P.dtloc[qs + up]);
      const double ul = dtu * u;
      P.uloc[qs] = ul;
      P.ucum[qs] = uc + dtu;
      for (int f = 0; f < P.nf; ++f) {
        if (P.k0 + kb >= P.nlay[f]) continue;
        const long qf = ((long)f * P.nb + kb) * P.slab + x.q;
        const double fx = P.fldlo[qf + up] * ul;
        P.flx[qf] = fx;
        P.flxcum[qf] = P.flxcum[qf] + fx;
      }
    } else {
      P.uloc[qs] = 0.0;
      for (int f = 0; f < P.nf; ++f)
        if (P.k0 + kb < P.nlay[f]) P.flx[((long)f * P.nb + kb) * P.slab + x.q] = 0.0;
    }
  }
  if (m & M_IV) {
    const double vc = P.vcum[qs];
    if (vc != dt2) {
      const double v = P.v[qk];
      const long up = (v >= 0) ? -(long)P.pitch : 0;
      const double dtv = dmin2(dt2 - vc, P.dtloc[qs + up]);
      const double vl = dtv * v;
      P.vloc[qs] = vl;
      P.vcum[qs] = vc + dtv;
      for (int f = 0; f < P.nf; ++f) {
        if (P.k0 + kb >= P.nlay[f]) continue;
        const long qf = ((long)f * P.nb + kb) * P.slab + x.q;
        const double fy = P.fldlo[qf + up] * vl;
        P.fly[qf] = fy;
        P.flycum[qf] = P.flycum[qf] + fy;
      }
    } else {
      P.vloc[qs] = 0.0;
      for (int f = 0; f < P.nf; ++f)
        if (P.k0 + kb < P.nlay[f]) P.fly[((long)f * P.nb + kb) * P.slab + x.q] = 0.0;
    }
  }
}

__global__ void __launch_bounds__(256) k_c_cells(const Fct2cParams P) {
  const Cell x = cell_of(P, 3);
  const int kb = blockIdx.z;
  if (!x.in || !(P.mask[x.q] & M_IP)) return;
  const long qs = x.q + (long)kb * P.slab;
  if (!P.lcalc[qs]) return;   // lcalc from the previous iteration (:1167)
  const double epsil = 1.e-10;   // :1027
  const double sci = P.scp2i[x.q], h = P.hloc[qs];
  const double qp = h - (P.uloc[qs + 1] - P.uloc[qs] + P.vloc[qs + P.pitch] - P.vloc[qs]) * sci;
  if (qp > 0.0) {   // ":1169 it may happen that the cfl is violated, leading to a negative h"
    for (int f = 0; f < P.nf; ++f) {
      if (P.k0 + kb >= P.nlay[f]) continue;
      const long qf = ((long)f * P.nb + kb) * P.slab + x.q;
      P.fldlo[qf] = ((epsil + h) * P.fldlo[qf] -
                     (P.flx[qf + 1] - P.flx[qf] + P.fly[qf + P.pitch] - P.fly[qf]) * sci) / (epsil + qp);
    }
  }
  P.hloc[qs] = qp;
  const double dt2 = P.dt2;
  P.lcalc[qs] = P.ucum[qs + 1] != dt2 || P.ucum[qs] != dt2 || P.vcum[qs + P.pitch] != dt2 || P.vcum[qs] != dt2;
}

// fax -> flx, fay -> fly (the low-order fluxes of the last iteration are dead)
__global__ void __launch_bounds__(256) k_c_fax(const Fct2cParams P) {
  const Cell x = cell_of(P, 3);
  const int kb = blockIdx.z;
  if (!x.in) return;
  const unsigned m = P.mask[x.q];
  const long qk = x.q + (long)(P.k0 + kb) * P.slab;
  for (int f = 0; f < P.nf; ++f) {
    if (P.k0 + kb >= P.nlay[f]) continue;
    const long qf = ((long)f * P.nb + kb) * P.slab + x.q;
    const double* fc = P.fldc[f];
    double fax = 0.0, fay = 0.0;   // coast faces :1223-1245
    if (m & M_IU) fax = P.u[qk] * 0.5 * (fc[qk] + fc[qk - 1]) - P.flxcum[qf] / P.dt2;        // :1206-1207
    if (m & M_IV) fay = P.v[qk] * 0.5 * (fc[qk] + fc[qk - P.pitch]) - P.flycum[qf] / P.dt2;  // :1210-1211
    P.flx[qf] = fax;
    P.fly[qf] = fay;
  }
}

// rp -> flxcum, rm -> flycum (dead after k_c_fax)
__global__ void __launch_bounds__(256) k_c_ratio(const Fct2cParams P) {
  const Cell x = cell_of(P, 2);
  const int kb = blockIdx.z;
  if (!x.in) return;
  const unsigned m = P.mask[x.q];
  if (!(m & M_IP)) return;
  const double epsil = 1.e-10;
  const long qs = x.q + (long)kb * P.slab;
  const double qdt2 = 1.0 / P.dt2;
  const long dw = (m & M_PW) ? -1 : 0, de = (m & M_PE) ? 1 : 0;                       // ia, ib
  const long ds = (m & M_PS) ? -(long)P.pitch : 0, dn = (m & M_PN) ? P.pitch : 0;     // ja, jb
  const double hs = P.hloc[qs] , sc = P.scp2[x.q];
  for (int f = 0; f < P.nf; ++f) {
    if (P.k0 + kb >= P.nlay[f]) continue;
    const long qf = ((long)f * P.nb + kb) * P.slab + x.q;
    const double* lo = P.fldlo;
    const double c0 = lo[qf], cw = lo[qf + dw], ce = lo[qf + de], cs = lo[qf + ds], cn = lo[qf + dn];
    const double fqmax = dmax2(dmax2(dmax2(dmax2(c0, cw), ce), cs), cn);
    const double fqmin = dmin2(dmin2(dmin2(dmin2(c0, cw), ce), cs), cn);
    const double fx = P.flx[qf], fxe = P.flx[qf + 1], fy = P.fly[qf], fyn = P.fly[qf + P.pitch];
    const double famax = dmax2(0.0, fx) - dmin2(0.0, fxe) + dmax2(0.0, fy) - dmin2(0.0, fyn);   // :1274-1275
    const double famin = dmax2(0.0, fxe) - dmin2(0.0, fx) + dmax2(0.0, fyn) - dmin2(0.0, fy);   // :1276-1277
    double rp = 0.0, rm = 0.0;
    if (famax > epsil) {
      const double qp = (fqmax - c0) * hs * sc * qdt2;
      rp = qp < famax ? qp / famax : 1.0;
    }
    if (famin > epsil) {
      const double qm = (c0 - fqmin) * hs * sc * qdt2;
      rm = qm < famin ? qm / famin : 1.0;
    }
    P.flxcum[qf] = rp;
    P.flycum[qf] = rm;
  }
}

__global__ void __launch_bounds__(256) k_c_final(const Fct2cParams P) {
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y, kb = blockIdx.z;
  if (c >= P.pitch || r >= P.nrows) return;
  const long q = (long)r * P.pitch + c;
  const long qs = q + (long)kb * P.slab, qk = q + (long)(P.k0 + kb) * P.slab;
  const unsigned m = P.mask[q];
  const double epsil = 1.e-10;
  for (int f = 0; f < P.nf; ++f) {
    if (P.k0 + kb >= P.nlay[f]) continue;
    double nv = P.fld[f][qk];
    if (m & M_OUT) {
      const long qf = ((long)f * P.nb + kb) * P.slab + q;
      const double *rp = P.flxcum, *rm = P.flycum, *fax = P.flx, *fay = P.fly;
      // :1315-1332 at the four faces of the cell (land faces hold 0.0 and stay 0.0)
      auto lim = [&](long a, long b) {   // face between cell a (west/south) and cell b
        const bool x = b - a == 1;
        const double fl = x ? fax[b] : fay[b];
        const double fact = fl < 0.0 ? dmin2(rp[a], rm[b]) : dmin2(rp[b], rm[a]);
        return fact * fl;
      };
      const bool fw = m & M_IU, fs = m & M_IV;
      const bool fe = P.mask[q + 1] & M_IU, fn = P.mask[q + P.pitch] & M_IV;
      const double xw = fw ? lim(qf - 1, qf) : fax[qf];
      const double xe = fe ? lim(qf, qf + 1) : fax[qf + 1];
      const double ys = fs ? lim(qf - P.pitch, qf) : fay[qf];
      const double yn = fn ? lim(qf, qf + P.pitch) : fay[qf + P.pitch];
      const double flxdiv = ((xe - xw) + (yn - ys)) * P.dt2 * P.scp2i[q];    // :1348-1349
      const double h = P.hloc[qs], lo = P.fldlo[qf];
      nv = h > 0. ? ((epsil + h) * lo - flxdiv) / (epsil + h) : lo;          // :1350-1355
    }
    P.out[f][qk] = nv;
  }
}

}  // namespace

int launch_fct2c(int stage, const Fct2cParams& P, cudaStream_t stream) {
  const dim3 block(32, 8), grid((P.pitch + 31) / 32, (P.nrows + 7) / 8, P.nb);
  switch (stage) {
    case 0: k_c_init<<<grid, block, 0, stream>>>(P); break;
    case 1: k_c_dtloc<<<grid, block, 0, stream>>>(P); break;
    case 2: k_c_faces<<<grid, block, 0, stream>>>(P); break;
    case 3: k_c_cells<<<grid, block, 0, stream>>>(P); break;
    case 4: k_c_fax<<<grid, block, 0, stream>>>(P); break;
    case 5: k_c_ratio<<<grid, block, 0, stream>>>(P); break;
    case 6: k_c_final<<<grid, block, 0, stream>>>(P); break;
    default: return -1;
  }
  return (int)cudaGetLastError();
}

}  // namespace tsadvc
