// advem_fct2 (mod_tsadvc.F90:645-997) and advem_fct4 (:1370-1706) + tsadvc prolog
// (:1905-1942) as a scheme of the TMA-staged march (march_tma_common.cuh).  The two differ
// in the high-order flux of S3 only: fct4 uses the 4-point formula 7/12, -1/12 (:1534-1554)
// unless a neighbouring face is land, and therefore forms fax,fay one row later (row r-1,
// when fldc of row r is staged).
//
//   stage A, row r   : S1 upwind fluxes flx,fly (:692-707) and S3 antidiffusive fluxes
//                      fax,fay (:823-830); coast zeroing (:738-758, :835-855) by select
//   stage B, row r-1 : prolog fco,fcn (:1934-1938), S1 extrema (:708-717), S2 low-order
//                      solution fldlo and fmxlo,fmnlo (:786-795)
//   stage C, row r-2 : S4 Zalesak ratios rp,rm and fmx,fmn (:869-906)
//   stage D, row r-2 : S5 flux limiting (:926-945)
//   stage E, row r-3 : S6 update (:968-980) and store
//
// Instruction diet (the kernel is issue-bound, not HBM-bound: DESIGN.md section 4):
//   * sea-only neighbour selection (ipim1.. of bigrid.F90:316-341) is folded into the
//     predicate input of the DSETP of each max/min step instead of a 64-bit select;
//   * max(0,x), min(0,x) of S4 are formed as (x+|x|), (x-|x|) = twice the exact parts,
//     the factor 2 is carried through famax/famin and 2*qdt2 and cancels in the quotient
//     (scaling by 2 is exact, so every rounding is the reference's);
//   * rp/rm are the reference's (q < fa ? q/fa : 1) with the quotient formed unconditionally
//     (one compare and one select per ratio); where fa == 0 the reference stores 0, here 1,
//     but that value only ever multiplies fluxes that are zero, so any finite stand-in gives
//     identical results;
//   * row segments that are sea on every staged column are marched by a mask-free instantiation;
//   * every register ring has period 3 and the ring slots of the staged rows rotate at run
//     time, so the loop is unrolled three times, not six: both row bodies fit the
//     instruction cache (march_tma_common.cuh, kPeriod);
//   * comparisons against zero use the sign bit on the integer pipe;
//   * one reciprocal of (fcn+onemu) serves the divisions of S2 and S6;
//   * divisions are branch-free with a sticky flag and a whole-chunk redo (march_common.cuh).
#pragma once
#include "march_tma_common.cuh"

namespace tsadvc {

template <int NC>
struct Fct2T {                                       // computed intermediates only
  // every ring is indexed by row mod 3 (the loop is unrolled three times); the two-row rings
  // simply die one row earlier, registers are allocated by liveness
  double DFLX[3][NC], FLY[3][NC];                    // flx(i+1)-flx(i), fly              [row%3]
  double FAX[3][NC], FAY[3][NC];                     // antidiffusive fluxes              [row%3]
  double LO[3][NC], FCN[3][NC], Y[3][NC];            // fldlo, fcn, 1/(fcn+onemu)         [row%3]
  double MXL[3][NC], MNL[3][NC];                     // fmxlo, fmnlo                      [row%3]
  double RP[3][NC], RM[3][NC];                       //                                   [row%3]
  double QMX[3][NC], QMN[3][NC];                     // fmx, fmn of S4                    [row%3]
  double DFAXL[3][NC], FAYL[3][NC];                  // limited fluxes                    [row%3]
  double FLXR[3][NC];                                // flx itself (fct4 only)            [row%3]
  unsigned m1, m2, m3;                               // masks of rows r-1, r-2, r-3
};

template <int NC, int ORDER = 2, int SEA = 0, int ISO = 0>
struct Fct2Scheme {
  typedef Fct2T<NC> State;
  static constexpr bool kNeedC = true;
  static constexpr int kPeriod = 3;
  static constexpr int kLag = 3;               // the row finished in iteration r is row r-3
  static constexpr bool kNeedM = (SEA == 0);   // the mask plane is staged for the general body only
  // ISO: the prolog is built from its own pair of mass fluxes (ring arrays U2, V2; isopyc layer 1)
  static constexpr int kArrays = ISO ? 10 : (kNeedM ? 8 : 7);

  static __device__ __forceinline__ void init(State& s) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        s.FAX[q][c] = 0.0; s.FAY[q][c] = 0.0; s.LO[q][c] = 0.0; s.FCN[q][c] = 0.0;
        s.Y[q][c] = 1.0; s.MXL[q][c] = 0.0; s.MNL[q][c] = 0.0;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        s.DFLX[q][c] = 0.0; s.FLY[q][c] = 0.0; s.RP[q][c] = 0.0; s.RM[q][c] = 0.0;
        s.QMX[q][c] = 0.0; s.QMN[q][c] = 0.0; s.DFAXL[q][c] = 0.0; s.FAYL[q][c] = 0.0;
        s.FLXR[q][c] = 0.0;
      }
    }
    s.m1 = s.m2 = s.m3 = 0u;
  }

  // A marched row whose four live rows (r .. r-3) are sea, interior and surrounded by sea on all
  // 32*NC columns (every mask byte 0xff) needs none of the mask logic: no land-face zeroing, no
  // sea-only neighbour selection, no old-value select at the store: 14 % fewer instructions, all of
  // them from the saturated ALU pipe.  Results are identical by construction (every select takes
  // its "sea" branch).  The choice is made per row segment on the host, not per row in the loop: the
  // launch is split into segments of all-sea rows (SEA = 1, the mask-free body) and the rest
  // (SEA = 0), one kernel each (tsadvc_abi.cu, march_segments).  Both bodies in one loop, chosen by a
  // vote per row, thrash the instruction cache (20 % of the stall samples unrolled six times, 8 %
  // unrolled three times; profiles/r01z_*, r01zb_*) and end up slower than the general body alone.
  template <int PH, bool SAFE>
  static __device__ __forceinline__ void step(State& s, const TmaCtx& x, const RingPtr& p, const int r,
                                              const SlotOff& so, bool& bad) {
    const unsigned m0 = ld_mask_s<NC>(p, so.s0);
    body<PH, SAFE, (SEA != 0)>(s, x, p, r, so, bad, m0);
  }

  template <int PH, bool SAFE, bool ALLSEA>
  static __device__ __forceinline__ void body(State& s, const TmaCtx& x, const RingPtr& p, const int r,
                                              const SlotOff& so, bool& bad, const unsigned m0) {
  typedef Ring<NC> R;
  // ring indices of rows r (and r-3), r-1, r-2
  constexpr int a3 = PH % 3, b3 = (PH + 2) % 3, c3 = (PH + 1) % 3;
  const Off s0 = so.s0, s1 = so.s1, s2 = so.s2, s3 = so.s3;   // staged rows r..r-3
  const double onemu = 9806.e-12;  // :671
  const double dt2 = x.dt2;

  // ---- stage A: row r
  double F0[NC], F1[NC], C1[NC];
  ld_own<NC, R::F>(p, s0, F0);
  ld_own<NC, R::F>(p, s1, F1);
  ld_own<NC, R::C>(p, s1, C1);
  double V0[NC];
  ld_own<NC, R::V>(p, s0, V0);
  {
    double C0[NC], U0[NC], FW[NC], CW[NC], flx[NC];
    ld_own<NC, R::C>(p, s0, C0);
    ld_own<NC, R::U>(p, s0, U0);
    ld_west<NC, R::F>(p, s0, F0, FW);
    ld_west<NC, R::C>(p, s0, C0, CW);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const unsigned mc = mk(m0, c);
      const double F = F0[c], C = C0[c], U = U0[c], V = V0[c];
      const double qx = signbit_set(U) ? F : FW[c];               // :693-697
      const double qy = signbit_set(V) ? F : F1[c];               // :700-704
      flx[c] = (ALLSEA || (mc & M_IU)) ? U * qx : 0.0;
      const double fly = (ALLSEA || (mc & M_IV)) ? V * qy : 0.0;
      if (ORDER == 2) {
        const double fhx = U * 0.5 * (C + CW[c]);                   // :824
        const double fhy = V * 0.5 * (C + C1[c]);                   // :828
        s.FAX[a3][c] = (ALLSEA || (mc & M_IU)) ? fhx - flx[c] : 0.0;
        s.FAY[a3][c] = (ALLSEA || (mc & M_IV)) ? fhy - fly : 0.0;
      } else {
        s.FLXR[a3][c] = flx[c];
      }
      s.FLY[a3][c] = fly;
    }
    ediff<NC>(flx, s.DFLX[a3]);
  }
  if (ORDER == 4) {
    // ---- S3 of advem_fct4 for row r-1 (:1528-1558): fldc of rows r-3..r and columns i-2..i+1
    const unsigned m1 = s.m1;
    double Cc[NC], Cw[NC], Cww[NC], Ce[NC], Cn[NC], Cs[NC], Css[NC], U1[NC], V1[NC];
    ld_own<NC, R::C>(p, s1, Cc);
    ld_own<NC, R::C>(p, s0, Cn);
    ld_own<NC, R::C>(p, s2, Cs);
    ld_own<NC, R::C>(p, s3, Css);
    ld_own<NC, R::U>(p, s1, U1);
    ld_own<NC, R::V>(p, s1, V1);
    ld_west<NC, R::C>(p, s1, Cc, Cw);
    ld_east<NC, R::C>(p, s1, Cc, Ce);
    unsigned mw[NC], me[NC];
    Cww[0] = ld_at<NC, R::C>(p.w2, s1);
    if (NC == 2) Cww[NC - 1] = Cw[0];
    if (!ALLSEA) {   // (the mask plane is not staged for the mask-free body)
      mw[0] = ld_mask_at<NC>(p.w, s1);
      me[NC - 1] = ld_mask_at<NC>(p.e, s1);
      if (NC == 2) { mw[NC - 1] = mk(m1, 0); me[0] = mk(m1, 1); }
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) { mw[c] = 0xffu; me[c] = 0xffu; }
    }
    const double ft14 = 7.0 / 12.0, ft24 = -1.0 / 12.0;              // :1398-1399
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const unsigned mc = mk(m1, c);
      const double U = U1[c], V = V1[c];
      const bool lowx = !ALLSEA && (!(mw[c] & M_IU) || !(me[c] & M_IU));         // iu(i-1)==0 .or. iu(i+1)==0
      const bool lowy = !ALLSEA && (!(mk(s.m2, c) & M_IV) || !(mk(m0, c) & M_IV));
      const double fhx2 = U * 0.5 * (Cc[c] + Cw[c]);
      const double fhx4 = U * (ft14 * (Cc[c] + Cw[c]) + ft24 * (Ce[c] + Cww[c]));
      const double fhy2 = V * 0.5 * (Cc[c] + Cs[c]);
      const double fhy4 = V * (ft14 * (Cc[c] + Cs[c]) + ft24 * (Cn[c] + Css[c]));
      const double fhx = lowx ? fhx2 : fhx4;
      const double fhy = lowy ? fhy2 : fhy4;
      s.FAX[b3][c] = (ALLSEA || (mc & M_IU)) ? fhx - s.FLXR[b3][c] : 0.0;
      s.FAY[b3][c] = (ALLSEA || (mc & M_IV)) ? fhy - s.FLY[b3][c] : 0.0;
    }
  }

  // ---- stage B: row r-1
  {
    const unsigned m1 = s.m1;
    double Fw[NC], Fe[NC], F2[NC], U1[NC], UE[NC], V1[NC], D1[NC], SCI1[NC];
    ld_west<NC, R::F>(p, s1, F1, Fw);
    ld_east<NC, R::F>(p, s1, F1, Fe);
    ld_own<NC, R::F>(p, s2, F2);
    // mass fluxes of the prolog: rows r-1 (own and east face) and r
    constexpr int PU = ISO ? (int)R::U2 : (int)R::U, PV = ISO ? (int)R::V2 : (int)R::V;
    double VP0[NC];
    ld_own<NC, PU>(p, s1, U1);
    ld_east<NC, PU>(p, s1, U1, UE);
    ld_own<NC, PV>(p, s1, V1);
    if (ISO) {
      ld_own<NC, PV>(p, s0, VP0);
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) VP0[c] = V0[c];
    }
    ld_own<NC, R::D>(p, s1, D1);
    ld_own<NC, R::SCI>(p, s1, SCI1);
    double q[NC], b[NC], y[NC], fmx[NC], fmn[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double Fc = F1[c];
      // 5-point sea-only extrema of fld (:709-716)
      double mx, mn;
      if (ALLSEA) {
        // max and min of the same pair share the compare (max/min are exact: any order gives the
        // reference's value; only the sign of a zero among equal zeros may differ)
        const bool gwe = Fw[c] > Fe[c], gsn = F2[c] > F0[c];
        const double xwe = gwe ? Fw[c] : Fe[c], nwe = gwe ? Fe[c] : Fw[c];
        const double xsn = gsn ? F2[c] : F0[c], nsn = gsn ? F0[c] : F2[c];
        mx = fmax2(xwe, Fc);      mn = fmin2(nwe, Fc);
        mx = fmax2(xsn, mx);      mn = fmin2(nsn, mn);
      } else {
        maxmin_first(mx, mn, Fc, Fc, Fw[c], Fw[c], m1, M_PW << (8 * c));
        maxmin_if(mx, mn, Fe[c], Fe[c], m1, M_PE << (8 * c));
        maxmin_if(mx, mn, F2[c], F2[c], m1, M_PS << (8 * c));
        maxmin_if(mx, mn, F0[c], F0[c], m1, M_PN << (8 * c));
      }
      fmx[c] = mx; fmn[c] = mn;
      // tsadvc prolog :1934-1938 (onetamas(:,:,m) = 1.0 when .not.btrmas, :1809)
      const double fdp = ((UE[c] - U1[c]) + (VP0[c] - V1[c])) * dt2 * SCI1[c];
      const double Dc = D1[c];
      const double fco = pos_part(Dc + fdp);
      const double fcn = pos_part(Dc);
      // :786-793
      const double flxdiv = ((s.DFLX[b3][c]) + (s.FLY[a3][c] - s.FLY[b3][c])) * dt2 * SCI1[c];
      q[c] = Fc * (fco + onemu) - flxdiv;
      b[c] = fcn + onemu;
      y[c] = SAFE ? 0.0 : rcp_nr(b[c]);
      s.FCN[b3][c] = fcn;
      s.Y[b3][c] = y[c];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double Fc = F1[c], Cc = C1[c];
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
    double SC2[NC];
    ld_own<NC, R::SC>(p, s2, SC2);
    double mxw[NC], mxe[NC], mnw[NC], mne[NC], faxe[NC];
    west_of<NC>(s.MXL[c3], mxw);
    east_of<NC>(s.MXL[c3], mxe);
    west_of<NC>(s.MNL[c3], mnw);
    east_of<NC>(s.MNL[c3], mne);
    east_of<NC>(s.FAX[c3], faxe);
    double qq[2 * NC], bb[2 * NC], rr[2 * NC];
    bool lt[2 * NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const unsigned mc = mk(m2, c);
      const bool pe = ALLSEA || (mc & M_PE), pn = ALLSEA || (mc & M_PN);
      // 5-point sea-only extrema of fmxlo/fmnlo (:872-879)
      double fqmax, fqmin;
      if (ALLSEA) {
        fqmax = fmax2(mxw[c], s.MXL[c3][c]);     fqmin = fmin2(mnw[c], s.MNL[c3][c]);
        fqmax = fmax2(mxe[c], fqmax);            fqmin = fmin2(mne[c], fqmin);
        fqmax = fmax2(s.MXL[a3][c], fqmax);      fqmin = fmin2(s.MNL[a3][c], fqmin);
        fqmax = fmax2(s.MXL[b3][c], fqmax);      fqmin = fmin2(s.MNL[b3][c], fqmin);
      } else {
        maxmin_first(fqmax, fqmin, s.MXL[c3][c], s.MNL[c3][c], mxw[c], mnw[c], m2, M_PW << (8 * c));
        maxmin_if(fqmax, fqmin, mxe[c], mne[c], m2, M_PE << (8 * c));
        maxmin_if(fqmax, fqmin, s.MXL[a3][c], s.MNL[a3][c], m2, M_PS << (8 * c));
        maxmin_if(fqmax, fqmin, s.MXL[b3][c], s.MNL[b3][c], m2, M_PN << (8 * c));
      }
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
      const double qp2 = (fqmax - lo) * fcn * SC2[c] * x.qdt2x2;   // 2*qp  :885
      const double qm2 = (lo - fqmin) * fcn * SC2[c] * x.qdt2x2;   // 2*qm  :895
      // qp >= 0 and famax >= 0 (sums of non-negative parts), so qp < famax implies famax > 0
      qq[2 * c] = qp2;     bb[2 * c] = famax2;     lt[2 * c] = qp2 < famax2;
      qq[2 * c + 1] = qm2; bb[2 * c + 1] = famin2; lt[2 * c + 1] = qm2 < famin2;
      s.QMX[c3][c] = fqmax;                                 // :904
      s.QMN[c3][c] = fqmin;                                 // :905
    }
#pragma unroll
    for (int i = 0; i < 2 * NC; ++i)
      rr[i] = div_flag_if<SAFE>(qq[i], bb[i], SAFE ? 0.0 : rcp_nr(bb[i]), lt[i], bad);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      // :884-903: rp = famax>0 ? (qp<famax ? qp/famax : 1) : 0 ; see the header note for fa==0.
      // The quotient is only used where qp < famax (elsewhere it may be inf/NaN: fa == 0)
      s.RP[c3][c] = lt[2 * c] ? rr[2 * c] : 1.0;
      s.RM[c3][c] = lt[2 * c + 1] ? rr[2 * c + 1] : 1.0;
    }
    // S5 (:926-945).  fax/fay are already zero on land faces, so no further select
    double rpw[NC], rmw[NC], faxl[NC];
    west_of<NC>(s.RP[c3], rpw);
    west_of<NC>(s.RM[c3], rmw);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double faxc = s.FAX[c3][c], fayc = s.FAY[c3][c];
      const bool ngx = signbit_set(faxc), ngy = signbit_set(fayc);
      const double fx = fmin2(ngx ? rpw[c] : s.RP[c3][c], ngx ? s.RM[c3][c] : rmw[c]);
      const double fy = fmin2(ngy ? s.RP[a3][c] : s.RP[c3][c], ngy ? s.RM[c3][c] : s.RM[a3][c]);
      faxl[c] = fx * faxc;
      s.FAYL[c3][c] = fy * fayc;
    }
    ediff<NC>(faxl, s.DFAXL[c3]);
  }

  // ---- stage E: row r-3, S6 (:968-980) and store
  {
    double SCI3[NC], OLD3[NC], nv[NC];
    ld_own<NC, R::SCI>(p, s3, SCI3);
    ld_own<NC, R::F>(p, s3, OLD3);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double a = ((s.DFAXL[a3][c]) + (s.FAYL[c3][c] - s.FAYL[a3][c])) * dt2 * SCI3[c];
      const double d = div_flag<SAFE>(a, s.FCN[a3][c] + onemu, s.Y[a3][c], bad);
      nv[c] = fmax2(s.QMN[a3][c], fmin2(s.QMX[a3][c], s.LO[a3][c] - d));
    }
    if (ALLSEA) store_cells<NC>(p, nv);
    else store_row_masked<NC>(p, s.m3, OLD3, nv);
  }
  s.m3 = s.m2; s.m2 = s.m1; s.m1 = m0;
  }


};

}  // namespace tsadvc
