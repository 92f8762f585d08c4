// advem_fct2 (mod_tsadvc.F90:645-997) fused with the tsadvc prolog (:1905-1942):
// one warp marches a strip of 32*NC columns along j, NC adjacent cells per lane.
//
// Software pipeline per loop iteration t (row r = r0 + t):
//   stage A, row r   : S1 upwind fluxes flx,fly (:692-707) and S3 antidiffusive fluxes
//                      fax,fay (:823-830); coast zeroing (:738-758, :835-855) by select
//   stage B, row r-1 : prolog fco,fcn (:1934-1938), S1 extrema (:708-717), S2 low-order
//                      solution fldlo and fmxlo,fmnlo (:786-795)
//   stage C, row r-2 : S4 Zalesak ratios rp,rm and fmx,fmn (:869-906)
//   stage D, row r-2 : S5 flux limiting (:926-945)
//   stage E, row r-3 : S6 update (:968-980) and store
// j-neighbours are older rows held in registers, i-neighbours come from warp shuffles.
// The row history lives in small rings indexed by (row mod 2) or (row mod 3); the loop is
// unrolled six times with the phase as a template parameter, so every ring index is a
// compile-time constant and the rings are plain registers that never have to be rotated.
//
// Instruction diet (the kernel is issue-bound, not HBM-bound: DESIGN.md section 4):
//   * sea-only neighbour selection (ipim1.. of bigrid.F90:316-341) is folded into the
//     predicate input of the DSETP of each max/min step instead of a 64-bit select;
//   * max(0,x), min(0,x) of S4 are formed as (x+|x|), (x-|x|) = twice the exact parts,
//     the factor 2 is carried through famax/famin and 2*qdt2 and cancels in the quotient
//     (scaling by 2 is exact, so every rounding is the reference's);
//   * rp/rm are min(1, q/fa) evaluated with the reference's guard fa > 0; where fa == 0
//     the reference stores 0 but that value only ever multiplies fluxes that are zero,
//     so any finite stand-in gives identical results (and saves two selects);
//   * comparisons against zero use the sign bit on the integer pipe;
//   * one reciprocal of (fcn+onemu) serves the divisions of S2 and S6;
//   * the rare-operand tests of the divisions of a stage are merged into one branch.
#pragma once
#include "march_common.cuh"

namespace tsadvc {

template <int NC>
struct RowBufT {            // one prefetched input row
  double F[NC], C[NC], U[NC], V[NC], D[NC];
  unsigned m;               // mask bytes: cell c in bits 8c..8c+7
};

template <int NC>
struct Fct2State {
  RowBufT<NC> L[2];                                 // [row&1]
  double F[3][NC];                                  // fld(:,:,n), rows r, r-1, r-2      [row%3]
  double C[2][NC];                                  // fldc, rows r, r-1                 [row&1]
  double UD[2][NC], V[2][NC], D[2][NC];             // u(i+1)-u(i), v, dp                [row&1]
  double DFLX[2][NC], FLY[2][NC];                   // flx(i+1)-flx(i), fly              [row&1]
  double FAX[3][NC], FAY[3][NC];                    // antidiffusive fluxes              [row%3]
  double LO[3][NC], FCN[3][NC], Y[3][NC];           // fldlo, fcn, 1/(fcn+onemu)         [row%3]
  double MXL[3][NC], MNL[3][NC];                    // fmxlo, fmnlo                      [row%3]
  double RP[2][NC], RM[2][NC];                      //                                   [row&1]
  double QMX[2][NC], QMN[2][NC];                    // fmx, fmn of S4                    [row&1]
  double DFAXL[2][NC], FAYL[2][NC];                 // limited fluxes                    [row&1]
  double SCI[2][NC], SC[2][NC];                     // prefetched scp2i(row), scp2(row)  [row&1]
  unsigned m1, m2, m3;                              // masks of rows r-1, r-2, r-3
};

struct MarchCtx {
  const double* __restrict__ fld;
  const double* __restrict__ fldc;
  const double* __restrict__ u;
  const double* __restrict__ v;
  const double* __restrict__ dp;
  double* __restrict__ out;
  const uint8_t* __restrict__ mask;
  const double* __restrict__ scp2;
  const double* __restrict__ scp2i;
  int pitch, nrows;
  int col, lane;
  int j0, j1;
  double dt2, qdt2x2;
  double posdef;
  bool colok;
};

// ---- NC-generic access helpers ------------------------------------------------------
template <int NC> struct Vec { double v[NC]; };

template <int NC>
__device__ __forceinline__ Vec<NC> ld_vec(const double* __restrict__ base, long off, bool ok) {
  Vec<NC> r;
  if (NC == 2) {
    const Pair p = ld_pair(base, off, ok);
    r.v[0] = p.a; r.v[NC - 1] = p.b;
  } else {
    r.v[0] = ok ? __ldg(base + off) : 0.0;
  }
  return r;
}
template <int NC>
__device__ __forceinline__ unsigned ld_mask(const uint8_t* __restrict__ mask, long off, bool ok) {
  if (!ok) return 0u;
  if (NC == 2) return (unsigned)__ldg(reinterpret_cast<const unsigned short*>(mask + off));
  return (unsigned)__ldg(mask + off);
}
// value of the west / east neighbour cell of each of the lane's cells
template <int NC>
__device__ __forceinline__ void west_of(const double (&x)[NC], double (&w)[NC]) {
  w[0] = shup(x[NC - 1]);
  if (NC == 2) w[NC - 1] = x[0];
}
template <int NC>
__device__ __forceinline__ void east_of(const double (&x)[NC], double (&e)[NC]) {
  e[NC - 1] = shdn(x[0]);
  if (NC == 2) e[0] = x[NC - 1];
}
// x(i+1) - x(i) along the strip
template <int NC>
__device__ __forceinline__ void ediff(const double (&x)[NC], double (&d)[NC]) {
  double e[NC];
  east_of<NC>(x, e);
#pragma unroll
  for (int c = 0; c < NC; ++c) d[c] = e[c] - x[c];
}

// lanes of the strip interior (the apron of 3 columns on each side is recomputed by the
// neighbouring strip): NC=2: columns w0+3..w0+60, NC=1: w0+3..w0+28
template <int NC>
__device__ __forceinline__ void store_vec(double* __restrict__ out, long off, int lane, unsigned m,
                                          const Vec<NC>& old, const double (&nv)[NC]) {
  if (NC == 2) {
    const Pair o{old.v[0], old.v[NC - 1]};
    const double n2[2] = {nv[0], nv[NC - 1]};
    store_row(out, off, lane, m, o, n2);
  } else {
    if (lane >= 3 && lane <= 28) out[off] = (m & M_OUT) ? nv[0] : old.v[0];
  }
}

template <int NC>
__device__ __forceinline__ void fct2_load_row(const MarchCtx& x, int r, RowBufT<NC>& w) {
  const bool ok = x.colok && ((unsigned)r < (unsigned)x.nrows);
  const long off = (long)r * x.pitch + x.col;
  const Vec<NC> f = ld_vec<NC>(x.fld, off, ok), cc = ld_vec<NC>(x.fldc, off, ok),
                uu = ld_vec<NC>(x.u, off, ok), vv = ld_vec<NC>(x.v, off, ok),
                dd = ld_vec<NC>(x.dp, off, ok);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    w.F[c] = f.v[c]; w.C[c] = cc.v[c]; w.U[c] = uu.v[c]; w.V[c] = vv.v[c]; w.D[c] = dd.v[c];
  }
  w.m = ld_mask<NC>(x.mask, off, ok);
}

// mx = max(mx, xa), mn = min(mn, xb) restricted to neighbours that are sea: bit `bit` of
// the mask word goes into the predicate input of both DSETPs (no separate 64-bit select
// of the neighbour value, one LOP3 for the pair)
__device__ __forceinline__ void maxmin_if(double& mx, double& mn, double xa, double xb,
                                          unsigned mword, unsigned bit) {
  asm("{\n\t.reg .pred e, p, q;\n\t.reg .b32 t;\n\t"
      "and.b32 t, %4, %5;\n\tsetp.ne.u32 e, t, 0;\n\t"
      "setp.gt.and.f64 p, %2, %0, e;\n\t"
      "setp.lt.and.f64 q, %3, %1, e;\n\t"
      "selp.f64 %0, %2, %0, p;\n\t"
      "selp.f64 %1, %3, %1, q;\n\t}"
      : "+d"(mx), "+d"(mn) : "d"(xa), "d"(xb), "r"(mword), "r"(bit));
}

// first step of a masked extremum: mx = (sea && xa > cmx) ? xa : cmx, mn likewise, into fresh
// registers (the in-place form above would first have to copy the centre value)
__device__ __forceinline__ void maxmin_first(double& mx, double& mn, double cmx, double cmn,
                                             double xa, double xb, unsigned mword, unsigned bit) {
  asm("{\n\t.reg .pred e, p, q;\n\t.reg .b32 t;\n\t"
      "and.b32 t, %6, %7;\n\tsetp.ne.u32 e, t, 0;\n\t"
      "setp.gt.and.f64 p, %4, %2, e;\n\t"
      "setp.lt.and.f64 q, %5, %3, e;\n\t"
      "selp.f64 %0, %4, %2, p;\n\t"
      "selp.f64 %1, %5, %3, q;\n\t}"
      : "=d"(mx), "=d"(mn) : "d"(cmx), "d"(cmn), "d"(xa), "d"(xb), "r"(mword), "r"(bit));
}
// min(x, 1.0) as one compare and one select (the ?: form is pattern-matched into a
// NaN-propagating minimum that costs an extra fix-up instruction)
__device__ __forceinline__ double min_one(double x) {
  double r;
  asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, 0d3FF0000000000000;\n\t"
      "selp.f64 %0, %1, 0d3FF0000000000000, p;\n\t}" : "=d"(r) : "d"(x));
  return r;
}

template <int NC, int PH, bool SAFE>
__device__ __forceinline__ void fct2_step(Fct2State<NC>& s, const MarchCtx& x, const int r, bool& bad) {
  constexpr int p2 = PH & 1, q2 = p2 ^ 1;                          // rows r (r-2), r-1 (r-3)
  constexpr int a3 = PH % 3, b3 = (PH + 2) % 3, c3 = (PH + 1) % 3;  // rows r (r-3), r-1, r-2
  const double onemu = 9806.e-12;  // :671
  const double dt2 = x.dt2;

  // ---- loads of this iteration: input row r+1, metrics of the rows in stages B, C, E,
  //      old value of the row to be stored
  //      The metrics of stages B and C are fetched one iteration ahead like the input row
  //      (their first use sits at the head of the dependency chain of the iteration);
  //      stage E's operands are consumed last and come out of L1/L2.
  fct2_load_row<NC>(x, r + 1, s.L[q2]);
  const int r3 = r - 3;
  {
    const bool okA = x.colok && ((unsigned)r < (unsigned)x.nrows);
    const bool okB = x.colok && ((unsigned)(r - 1) < (unsigned)x.nrows);
    const Vec<NC> a = ld_vec<NC>(x.scp2i, (long)r * x.pitch + x.col, okA);        // B at r+1
    const Vec<NC> b = ld_vec<NC>(x.scp2, (long)(r - 1) * x.pitch + x.col, okB);   // C at r+1
#pragma unroll
    for (int c = 0; c < NC; ++c) { s.SCI[p2][c] = a.v[c]; s.SC[p2][c] = b.v[c]; }
  }
  const bool st3 = x.colok && (r3 >= x.j0) && (r3 < x.j1);
  const long off3 = (long)r3 * x.pitch + x.col;
  const Vec<NC> sci3 = ld_vec<NC>(x.scp2i, off3, st3);
  const Vec<NC> old3 = ld_vec<NC>(x.fld, off3, st3);

  // ---- stage A: row r.  Land cells may hold anything (appendix A.3): every use of a
  //      neighbour value below is guarded by the mask, never multiplied by it
  const RowBufT<NC>& in = s.L[p2];
  const unsigned m0 = in.m;
  {
    double FW[NC], CW[NC], UE[NC], flx[NC];
    west_of<NC>(in.F, FW);
    west_of<NC>(in.C, CW);
    east_of<NC>(in.U, UE);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const unsigned mc = mk(m0, c);
      const double F = in.F[c], C = in.C[c], U = in.U[c], V = in.V[c];
      const double qx = signbit_set(U) ? F : FW[c];               // :693-697
      const double qy = signbit_set(V) ? F : s.F[b3][c];          // :700-704
      flx[c] = (mc & M_IU) ? U * qx : 0.0;
      const double fly = (mc & M_IV) ? V * qy : 0.0;
      const double fhx = U * 0.5 * (C + CW[c]);                   // :824
      const double fhy = V * 0.5 * (C + s.C[q2][c]);              // :828
      s.FAX[a3][c] = (mc & M_IU) ? fhx - flx[c] : 0.0;
      s.FAY[a3][c] = (mc & M_IV) ? fhy - fly : 0.0;
      s.FLY[p2][c] = fly;
      s.UD[p2][c] = UE[c] - U;
      s.V[p2][c] = V;
      s.D[p2][c] = in.D[c];
      s.F[a3][c] = F;
      s.C[p2][c] = C;
    }
    ediff<NC>(flx, s.DFLX[p2]);
  }

  // ---- stage B: row r-1
  {
    const unsigned m1 = s.m1;
    double Fw[NC], Fe[NC], q[NC], b[NC], y[NC], fmx[NC], fmn[NC];
    west_of<NC>(s.F[b3], Fw);
    east_of<NC>(s.F[b3], Fe);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double Fc = s.F[b3][c];
      // 5-point sea-only extrema of fld (:709-716)
      double mx = Fc, mn = Fc;
      maxmin_if(mx, mn, Fw[c], Fw[c], m1, M_PW << (8 * c));
      maxmin_if(mx, mn, Fe[c], Fe[c], m1, M_PE << (8 * c));
      maxmin_if(mx, mn, s.F[c3][c], s.F[c3][c], m1, M_PS << (8 * c));
      maxmin_if(mx, mn, s.F[a3][c], s.F[a3][c], m1, M_PN << (8 * c));
      fmx[c] = mx; fmn[c] = mn;
      // tsadvc prolog :1934-1938 (onetamas(:,:,m) = 1.0 when .not.btrmas, :1809)
      const double fdp = ((s.UD[q2][c]) + (s.V[p2][c] - s.V[q2][c])) * dt2 * s.SCI[q2][c];
      const double Dc = s.D[q2][c];
      const double fco = pos_part(Dc + fdp);
      const double fcn = pos_part(Dc);
      // :786-793
      const double flxdiv = ((s.DFLX[q2][c]) + (s.FLY[p2][c] - s.FLY[q2][c])) * dt2 * s.SCI[q2][c];
      q[c] = Fc * (fco + onemu) - flxdiv;
      b[c] = fcn + onemu;
      y[c] = SAFE ? 0.0 : rcp_nr(b[c]);
      s.FCN[b3][c] = fcn;
      s.Y[b3][c] = y[c];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double Fc = s.F[b3][c], Cc = s.C[q2][c];
      const double lo = div_flag<SAFE>(q[c], b[c], y[c], bad);
      const double l = fmax2(fmn[c], fmin2(fmx[c], lo));
      const bool g = Fc > Cc;
      s.MXL[b3][c] = fmax2(g ? Fc : Cc, l);   // :794
      s.MNL[b3][c] = fmin2(g ? Cc : Fc, l);   // :795
      s.LO[b3][c] = l;
    }
  }

  // ---- stages C and D: row r-2
  {
    const unsigned m2 = s.m2;
    double mxw[NC], mxe[NC], mnw[NC], mne[NC], faxe[NC];
    west_of<NC>(s.MXL[c3], mxw);
    east_of<NC>(s.MXL[c3], mxe);
    west_of<NC>(s.MNL[c3], mnw);
    east_of<NC>(s.MNL[c3], mne);
    east_of<NC>(s.FAX[c3], faxe);
    double qq[2 * NC], bb[2 * NC], rr[2 * NC];
    bool pos[2 * NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const unsigned mc = mk(m2, c);
      const bool pe = mc & M_PE, pn = mc & M_PN;
      // 5-point sea-only extrema of fmxlo/fmnlo (:872-879)
      double fqmax = s.MXL[c3][c], fqmin = s.MNL[c3][c];
      maxmin_if(fqmax, fqmin, mxw[c], mnw[c], m2, M_PW << (8 * c));
      maxmin_if(fqmax, fqmin, mxe[c], mne[c], m2, M_PE << (8 * c));
      maxmin_if(fqmax, fqmin, s.MXL[a3][c], s.MNL[a3][c], m2, M_PS << (8 * c));
      maxmin_if(fqmax, fqmin, s.MXL[b3][c], s.MNL[b3][c], m2, M_PN << (8 * c));
      const double faxc = s.FAX[c3][c];
      const double faxb = pe ? faxe[c] : faxc;             // fax(ib,j)  :880
      const double fayc = s.FAY[c3][c];
      const double fayb = pn ? s.FAY[b3][c] : fayc;        // fay(i,jb)  :881
      // 2*max(0,x) = x+|x|, 2*min(0,x) = x-|x| (exact); famax2 = 2*famax, famin2 = 2*famin
      const double xp = faxc + fabs(faxc), xn = faxc - fabs(faxc);
      const double bp = faxb + fabs(faxb), bn = faxb - fabs(faxb);
      const double yp = fayc + fabs(fayc), yn = fayc - fabs(fayc);
      const double ybp = fayb + fabs(fayb), ybn = fayb - fabs(fayb);
      const double famax2 = xp - bn + yp - ybn;             // :882
      const double famin2 = bp - xn + ybp - yn;             // :883
      const double lo = s.LO[c3][c], fcn = s.FCN[c3][c];
      const double qp2 = (fqmax - lo) * fcn * s.SC[q2][c] * x.qdt2x2;   // 2*qp  :885
      const double qm2 = (lo - fqmin) * fcn * s.SC[q2][c] * x.qdt2x2;   // 2*qm  :895
      pos[2 * c] = famax2 > 0.0;
      pos[2 * c + 1] = famin2 > 0.0;
      qq[2 * c] = qp2;     bb[2 * c] = pos[2 * c] ? famax2 : 1.0;
      qq[2 * c + 1] = qm2; bb[2 * c + 1] = pos[2 * c + 1] ? famin2 : 1.0;
      s.QMX[p2][c] = fqmax;                                 // :904
      s.QMN[p2][c] = fqmin;                                 // :905
    }
#pragma unroll
    for (int i = 0; i < 2 * NC; ++i)
      rr[i] = div_flag<SAFE>(qq[i], bb[i], SAFE ? 0.0 : rcp_nr(bb[i]), bad);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      // :884-903: rp = famax>0 ? (qp<famax ? qp/famax : 1) : 0 ; see the header note for fa==0
      s.RP[p2][c] = fmin2(rr[2 * c], 1.0);
      s.RM[p2][c] = fmin2(rr[2 * c + 1], 1.0);
    }
    // S5 (:926-945).  fax/fay are already zero on land faces, so no further select
    double rpw[NC], rmw[NC], faxl[NC];
    west_of<NC>(s.RP[p2], rpw);
    west_of<NC>(s.RM[p2], rmw);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double faxc = s.FAX[c3][c], fayc = s.FAY[c3][c];
      const bool ngx = signbit_set(faxc), ngy = signbit_set(fayc);
      const double fx = fmin2(ngx ? rpw[c] : s.RP[p2][c], ngx ? s.RM[p2][c] : rmw[c]);
      const double fy = fmin2(ngy ? s.RP[q2][c] : s.RP[p2][c], ngy ? s.RM[p2][c] : s.RM[q2][c]);
      faxl[c] = fx * faxc;
      s.FAYL[p2][c] = fy * fayc;
    }
    ediff<NC>(faxl, s.DFAXL[p2]);
  }

  // ---- stage E: row r-3, S6 (:968-980) and store
  if (r3 >= x.j0 && r3 < x.j1) {
    double nv[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double a = ((s.DFAXL[q2][c]) + (s.FAYL[p2][c] - s.FAYL[q2][c])) * dt2 * sci3.v[c];
      const double d = div_flag<SAFE>(a, s.FCN[a3][c] + onemu, s.Y[a3][c], bad);
      nv[c] = fmax2(s.QMN[q2][c], fmin2(s.QMX[q2][c], s.LO[a3][c] - d));
    }
    if (x.colok) store_vec<NC>(x.out, off3, x.lane, s.m3, old3, nv);
  }
  s.m3 = s.m2; s.m2 = s.m1; s.m1 = m0;
}

template <int NC>
__device__ __forceinline__ void march_ctx(MarchCtx& x, const Job& jb, const Geo& g, int w0, int j0,
                                          int j1, int lane) {
  x.fld = jb.fld; x.fldc = jb.fldc; x.u = jb.u; x.v = jb.v; x.dp = jb.dp; x.out = jb.out;
  x.mask = g.mask; x.scp2 = g.scp2; x.scp2i = g.scp2i;
  x.pitch = g.pitch; x.nrows = g.nrows;
  x.col = w0 + NC * lane; x.lane = lane;
  x.j0 = j0; x.j1 = j1;
  x.dt2 = g.delt1;
  const double qdt2 = 1.0 / g.delt1;  // :865
  x.qdt2x2 = qdt2 + qdt2;
  x.posdef = jb.posdef;
  x.colok = (unsigned)x.col < (unsigned)g.pitch;
}

template <int NC, bool SAFE>
__device__ __forceinline__ bool march_fct2_pass(const MarchCtx& x) {
  Fct2State<NC> s;
  // rows below the chunk: zeros with an all-land mask (never stored: r3 >= j0)
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      s.F[q][c] = 0.0; s.FAX[q][c] = 0.0; s.FAY[q][c] = 0.0; s.LO[q][c] = 0.0; s.FCN[q][c] = 0.0;
      s.Y[q][c] = 1.0; s.MXL[q][c] = 0.0; s.MNL[q][c] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      s.C[q][c] = 0.0; s.UD[q][c] = 0.0; s.V[q][c] = 0.0; s.D[q][c] = 0.0; s.DFLX[q][c] = 0.0;
      s.FLY[q][c] = 0.0; s.RP[q][c] = 0.0; s.RM[q][c] = 0.0; s.QMX[q][c] = 0.0; s.QMN[q][c] = 0.0;
      s.DFAXL[q][c] = 0.0; s.FAYL[q][c] = 0.0;
    }
  }
  s.m1 = s.m2 = s.m3 = 0u;
  bool bad = false;
  const int r0 = x.j0 - 3;
  fct2_load_row<NC>(x, r0, s.L[0]);
  const int niter = (x.j1 - x.j0) + 6;   // rows j0-3 .. j1+2
  for (int t = 0; t < niter; t += 6) {
    const int r = r0 + t;
    fct2_step<NC, 0, SAFE>(s, x, r, bad);
    fct2_step<NC, 1, SAFE>(s, x, r + 1, bad);
    fct2_step<NC, 2, SAFE>(s, x, r + 2, bad);
    fct2_step<NC, 3, SAFE>(s, x, r + 3, bad);
    fct2_step<NC, 4, SAFE>(s, x, r + 4, bad);
    fct2_step<NC, 5, SAFE>(s, x, r + 5, bad);
  }
  return bad;
}

// the whole chunk again with the compiler's a/b: taken by a warp only when one of its
// lanes met denormal / huge / NaN operands (never on physical data)
template <int NC>
__device__ __noinline__ void march_fct2_safe(const MarchCtx x) { march_fct2_pass<NC, true>(x); }

template <int NC>
__device__ void march_fct2(const Job& jb, const Geo& g, int w0, int j0, int j1, int lane) {
  MarchCtx x;
  march_ctx<NC>(x, jb, g, w0, j0, j1, lane);
  const bool bad = march_fct2_pass<NC, false>(x);
  if (__any_sync(TSADVC_FULLMASK, bad)) march_fct2_safe<NC>(x);
}

}  // namespace tsadvc
