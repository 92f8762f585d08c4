// advem_mpdata (mod_tsadvc.F90:207-493) + tsadvc prolog (:1905-1942) as a scheme of the
// TMA-staged march (march_tma_common.cuh).
//
//   stage A, row r   : M1 (:254-271) tx1, ty1 and the upwind fluxes flx, fly; coast zeroing
//                      (:301-321) by select
//   stage B, row r-1 : prolog fco,fcn; M1 extrema (:272-281); M2 (:346-354) flxdiv, fldlo;
//                      M3 (:377-388) antidiffusive fluxes from flxdiv, fco, fcn of the cell
//                      and its west / south neighbour
//   stage C, row r-2 : M4 (:412-419) rp, rm;  M5 (:439-446) limited fluxes
//   stage E, row r-3 : M6 (:475-480) update, minus posdef, store
// fld is offset by posdef (256 for temperature, mod_tsadvc.F90:1762) so that it is positive
// definite; fldc is not used.  Land values of fld and dp are read as 0 where the stencil
// touches them (the reference never reads them; any finite stand-in gives the same sea
// results because every flux that multiplies them is coast-zeroed).
// Every register ring has period 3 and the ring slots of the staged rows rotate at run time (loop
// unrolled three times, march_tma_common.cuh kPeriod); SEA = 1 is the mask-free instantiation for
// row segments whose staged window is sea throughout (tsadvc_abi.cu, march_segments).
#pragma once
#include "march_tma_common.cuh"

namespace tsadvc {

template <int NC>
struct MpdataT {
  // all rings are indexed by row mod 3; the two-row ones die one row earlier
  double TX1[3][NC], TY1[3][NC];            // A(row r) -> B(row r, next iteration)
  double DFLX[3][NC], FLY[3][NC];           // low-order flux divergence pieces
  double FDV[3][NC], FCO[3][NC];            // flxdiv, fco of rows r-1, r-2
  double FCN[3][NC];                        // [row%3]   rows r-1, r-2, r-3
  double LO[3][NC], MX[3][NC], MN[3][NC];   // [row%3]   fldlo, fmx, fmn
  double FLX2[3][NC], FLX2E[3][NC], FLY2[3][NC];   // M3 fluxes (own face, east face)
  double RP[3][NC], RM[3][NC];
  double DFLX3[3][NC], FLY3[3][NC];         // limited fluxes
  unsigned m1, m2, m3;
};

template <int NC, int SEA = 0, int ISO = 0>
struct MpdataScheme {
  typedef MpdataT<NC> State;
  static constexpr bool kNeedC = false;
  static constexpr int kPeriod = 3;
  static constexpr int kLag = 3;
  static constexpr bool kNeedM = (SEA == 0);   // the mask plane is staged for the general body only
  // ISO: the prolog is built from its own pair of mass fluxes (ring arrays U2, V2; isopyc layer 1)
  static constexpr int kArrays = ISO ? 10 : (kNeedM ? 8 : 7);
  static constexpr bool ALLSEA = (SEA != 0);

  static __device__ __forceinline__ void init(State& s) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
      for (int q = 0; q < 3; ++q) { s.FCN[q][c] = 0.0; s.LO[q][c] = 0.0; s.MX[q][c] = 0.0; s.MN[q][c] = 0.0; }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        s.TX1[q][c] = 0.0; s.TY1[q][c] = 0.0; s.DFLX[q][c] = 0.0; s.FLY[q][c] = 0.0;
        s.FDV[q][c] = 0.0; s.FCO[q][c] = 0.0; s.FLX2[q][c] = 0.0; s.FLX2E[q][c] = 0.0;
        s.FLY2[q][c] = 0.0; s.RP[q][c] = 0.0; s.RM[q][c] = 0.0; s.DFLX3[q][c] = 0.0; s.FLY3[q][c] = 0.0;
      }
    }
    s.m1 = s.m2 = s.m3 = 0u;
  }

  // fld / dp with land cells read as zero
  template <int ARR>
  static __device__ __forceinline__ void ld_sea(const RingPtr& p, Off slot, unsigned m, double (&x)[NC]) {
    ld_own<NC, ARR>(p, slot, x);
    if (ALLSEA) return;
#pragma unroll
    for (int c = 0; c < NC; ++c) x[c] = (mk(m, c) & M_IP) ? x[c] : 0.0;
  }

  template <int PH, bool SAFE>
  static __device__ __forceinline__ void step(State& s, const TmaCtx& x, const RingPtr& p, const int r,
                                              const SlotOff& so, bool& bad) {
    typedef Ring<NC> R;
    constexpr int a3 = PH % 3, b3 = (PH + 2) % 3, c3 = (PH + 1) % 3;  // rows r (r-3), r-1, r-2
    const Off s0 = so.s0, s1 = so.s1, s2 = so.s2, s3 = so.s3;
    const double onemu = 9806.e-12;  // :236
    const double dt2 = x.dt2;
    const double posdef = x.posdef;

    // ---- stage A: row r
    const unsigned m0 = ld_mask_s<NC>(p, s0);
    double F0[NC], F1[NC], V0[NC];
    ld_sea<R::F>(p, s0, m0, F0);
    ld_sea<R::F>(p, s1, s.m1, F1);
    ld_own<NC, R::V>(p, s0, V0);
    {
      double U0[NC], FW[NC], flx[NC];
      ld_own<NC, R::U>(p, s0, U0);
      {
        // west neighbour of the first own cell, zero if land (its ip is this cell's M_PW)
        double fw[NC];
        ld_west<NC, R::F>(p, s0, F0, fw);
        FW[0] = (ALLSEA || (mk(m0, 0) & M_PW)) ? fw[0] : 0.0;
        if (NC == 2) FW[NC - 1] = F0[0];
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const unsigned mc = mk(m0, c);
        const double F = F0[c], U = U0[c], V = V0[c];
        s.TX1[a3][c] = .5 * fabs(U) * (F - FW[c]);                 // :254
        s.TY1[a3][c] = .5 * fabs(V) * (F - F1[c]);                 // :262
        const double qx = (U >= 0.0) ? FW[c] : F;                  // :255-259
        const double qy = (V >= 0.0) ? F1[c] : F;                  // :263-267
        flx[c] = (ALLSEA || (mc & M_IU)) ? U * (qx + posdef) : 0.0;
        s.FLY[a3][c] = (ALLSEA || (mc & M_IV)) ? V * (qy + posdef) : 0.0;
      }
      ediff<NC>(flx, s.DFLX[a3]);
    }

    // ---- stage B: row r-1
    {
      const unsigned m1 = s.m1;
      double Fw[NC], Fe[NC], F2[NC], U1[NC], UE[NC], V1[NC], D1[NC], SCI1[NC];
      {
        double fw[NC], fe[NC];
        ld_west<NC, R::F>(p, s1, F1, fw);
        ld_east<NC, R::F>(p, s1, F1, fe);
#pragma unroll
        for (int c = 0; c < NC; ++c) { Fw[c] = fw[c]; Fe[c] = fe[c]; }
      }
      ld_own<NC, R::F>(p, s2, F2);
      ld_own<NC, R::U>(p, s1, U1);
      ld_east<NC, R::U>(p, s1, U1, UE);
      ld_own<NC, R::V>(p, s1, V1);
      ld_sea<R::D>(p, s1, m1, D1);
      ld_own<NC, R::SCI>(p, s1, SCI1);
      // mass fluxes of the prolog (U1, V1 stay the advecting ones: M3 below)
      double UP1[NC], UPE[NC], VP1[NC], VP0[NC];
      if (ISO) {
        ld_own<NC, R::U2>(p, s1, UP1);
        ld_east<NC, R::U2>(p, s1, UP1, UPE);
        ld_own<NC, R::V2>(p, s1, VP1);
        ld_own<NC, R::V2>(p, s0, VP0);
      } else {
#pragma unroll
        for (int c = 0; c < NC; ++c) { UP1[c] = U1[c]; UPE[c] = UE[c]; VP1[c] = V1[c]; VP0[c] = V0[c]; }
      }
      double FDV[NC], FCO[NC], FCN[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double Fc = F1[c];
        // 5-point sea-only extrema of fld (:272-281), then + posdef
        double mx, mn;
        if (ALLSEA) {   // max and min of the same pair share the compare
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
        const double MX = mx + posdef, MN = mn + posdef;
        // tsadvc prolog :1934-1938
        const double fdp = ((UPE[c] - UP1[c]) + (VP0[c] - VP1[c])) * dt2 * SCI1[c];
        FCO[c] = fmax2(D1[c] + fdp, 0.0);
        FCN[c] = fmax2(D1[c], 0.0);
        // M2 :346-354
        FDV[c] = ((s.DFLX[b3][c]) + (s.FLY[a3][c] - s.FLY[b3][c])) * dt2 * SCI1[c];
        const double q = (Fc + posdef) * (FCO[c] + onemu) - FDV[c];
        const double b = FCN[c] + onemu;
        const double lo = div_flag<SAFE>(q, b, SAFE ? 0.0 : rcp_nr(b), bad);
        s.LO[b3][c] = fmax2(MN, fmin2(MX, lo));
        s.MX[b3][c] = MX; s.MN[b3][c] = MN;
        s.FCN[b3][c] = FCN[c];
      }
      // M3 :377-388 (operation order of :378-381, :390)
      double FDVW[NC], FCOW[NC], FCNW[NC];
      west_of<NC>(FDV, FDVW);
      west_of<NC>(FCO, FCOW);
      west_of<NC>(FCN, FCNW);
      double flx2[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const unsigned mc = mk(m1, c);
        const double ax = U1[c] * (FDV[c] + FDVW[c]);
        const double bx = ((FCO[c] + FCOW[c]) + (FCN[c] + FCNW[c])) + onemu;
        const double ay = V1[c] * (FDV[c] + s.FDV[c3][c]);
        const double by = ((FCO[c] + s.FCO[c3][c]) + (FCN[c] + s.FCN[c3][c])) + onemu;
        const double fx = s.TX1[b3][c] - div_flag<SAFE>(ax, bx, SAFE ? 0.0 : rcp_nr(bx), bad);
        const double fy = s.TY1[b3][c] - div_flag<SAFE>(ay, by, SAFE ? 0.0 : rcp_nr(by), bad);
        flx2[c] = (ALLSEA || (mc & M_IU)) ? fx : 0.0;
        s.FLY2[b3][c] = (ALLSEA || (mc & M_IV)) ? fy : 0.0;
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) { s.FLX2[b3][c] = flx2[c]; s.FDV[b3][c] = FDV[c]; s.FCO[b3][c] = FCO[c]; }
      east_of<NC>(flx2, s.FLX2E[b3]);
    }

    // ---- stage C: row r-2 (M4, M5).  FLX2/FLY2 of row r-2 sit in ring slot c3, of row r-1 in b3
    {
      const unsigned m2 = s.m2;
      double SC2[NC];
      ld_own<NC, R::SC>(p, s2, SC2);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double fxc = s.FLX2[c3][c], fxe = s.FLX2E[c3][c];
        const double fyc = s.FLY2[c3][c], fyn = s.FLY2[b3][c];
        const double flxdp = fmin2(0.0, fxe) - fmax2(0.0, fxc);     // :412-415
        const double flxdn = fmax2(0.0, fxe) - fmin2(0.0, fxc);
        const double flydp = fmin2(0.0, fyn) - fmax2(0.0, fyc);
        const double flydn = fmax2(0.0, fyn) - fmin2(0.0, fyc);
        const double w = s.FCN[c3][c] * SC2[c];
        const double ap = (s.MX[c3][c] - s.LO[c3][c]) * w, bp = (onemu - (flxdp + flydp)) * dt2;   // :416-417
        const double am = (s.LO[c3][c] - s.MN[c3][c]) * w, bm = (onemu + (flxdn + flydn)) * dt2;   // :418-419
        s.RP[c3][c] = div_flag<SAFE>(ap, bp, SAFE ? 0.0 : rcp_nr(bp), bad);
        s.RM[c3][c] = div_flag<SAFE>(am, bm, SAFE ? 0.0 : rcp_nr(bm), bad);
      }
      double RPW[NC], RMW[NC], flx3[NC];
      west_of<NC>(s.RP[c3], RPW);
      west_of<NC>(s.RM[c3], RMW);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const unsigned mc = mk(m2, c);
        const double fxc = s.FLX2[c3][c], fyc = s.FLY2[c3][c];
        const double RP = s.RP[c3][c], RM = s.RM[c3][c];
        const double x3 = fmax2(0.0, fxc) * fmin2(fmin2(1.0, RP), RMW[c]) +          // :439-441
                          fmin2(0.0, fxc) * fmin2(fmin2(1.0, RPW[c]), RM);
        const double y3 = fmax2(0.0, fyc) * fmin2(fmin2(1.0, RP), s.RM[a3][c]) +     // :443-445
                          fmin2(0.0, fyc) * fmin2(fmin2(1.0, s.RP[a3][c]), RM);
        flx3[c] = (ALLSEA || (mc & M_IU)) ? x3 : 0.0;
        s.FLY3[c3][c] = (ALLSEA || (mc & M_IV)) ? y3 : 0.0;
      }
      ediff<NC>(flx3, s.DFLX3[c3]);
    }

    // ---- stage E: row r-3, M6 (:475-480) and store
    {
      double SCI3[NC], OLD3[NC], nv[NC];
      ld_own<NC, R::SCI>(p, s3, SCI3);
      ld_own<NC, R::F>(p, s3, OLD3);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double flxdiv = ((s.DFLX3[a3][c]) + (s.FLY3[c3][c] - s.FLY3[a3][c])) * dt2 * SCI3[c];
        const double b = s.FCN[a3][c] + onemu;
        const double d = div_flag<SAFE>(flxdiv, b, SAFE ? 0.0 : rcp_nr(b), bad);
        const double f = fmax2(s.MN[a3][c], fmin2(s.MX[a3][c], s.LO[a3][c] - d));
        nv[c] = f - posdef;
      }
      if (ALLSEA) store_cells<NC>(p, nv);
      else store_row_masked<NC>(p, s.m3, OLD3, nv);
    }
    s.m3 = s.m2; s.m2 = s.m1; s.m1 = m0;
  }
};

}  // namespace tsadvc
