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
#pragma once
#include "march_tma_common.cuh"

namespace tsadvc {

template <int NC>
struct MpdataT {
  double TX1[2][NC], TY1[2][NC];            // [row&1]   A(row r) -> B(row r, next iteration)
  double DFLX[2][NC], FLY[2][NC];           // [row&1]   low-order flux divergence pieces
  double FDV[2][NC], FCO[2][NC];            // [row&1]   flxdiv, fco of rows r-1, r-2
  double FCN[3][NC];                        // [row%3]   rows r-1, r-2, r-3
  double LO[3][NC], MX[3][NC], MN[3][NC];   // [row%3]   fldlo, fmx, fmn
  double FLX2[2][NC], FLX2E[2][NC], FLY2[2][NC];   // [row&1]  M3 fluxes (own face, east face)
  double RP[2][NC], RM[2][NC];              // [row&1]
  double DFLX3[2][NC], FLY3[2][NC];         // [row&1]   limited fluxes
  unsigned m1, m2, m3;
};

template <int NC>
struct MpdataScheme {
  typedef MpdataT<NC> State;
  static constexpr bool kNeedC = false;
  static constexpr int kPeriod = 6;

  static __device__ __forceinline__ void init(State& s) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
      for (int q = 0; q < 3; ++q) { s.FCN[q][c] = 0.0; s.LO[q][c] = 0.0; s.MX[q][c] = 0.0; s.MN[q][c] = 0.0; }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        s.TX1[q][c] = 0.0; s.TY1[q][c] = 0.0; s.DFLX[q][c] = 0.0; s.FLY[q][c] = 0.0;
        s.FDV[q][c] = 0.0; s.FCO[q][c] = 0.0; s.FLX2[q][c] = 0.0; s.FLX2E[q][c] = 0.0;
        s.FLY2[q][c] = 0.0; s.RP[q][c] = 0.0; s.RM[q][c] = 0.0; s.DFLX3[q][c] = 0.0; s.FLY3[q][c] = 0.0;
      }
    }
    s.m1 = s.m2 = s.m3 = 0u;
  }

  // fld / dp with land cells read as zero
  template <int ARR>
  static __device__ __forceinline__ void ld_sea(const RingPtr& p, int slot, unsigned m, double (&x)[NC]) {
    ld_own<NC, ARR>(p, slot, x);
#pragma unroll
    for (int c = 0; c < NC; ++c) x[c] = (mk(m, c) & M_IP) ? x[c] : 0.0;
  }

  template <int PH, bool SAFE>
  static __device__ __forceinline__ void step(State& s, const TmaCtx& x, const RingPtr& p, const int r,
                                              bool& bad) {
    typedef Ring<NC> R;
    constexpr int p2 = PH & 1, q2 = p2 ^ 1;                           // rows r (r-2), r-1 (r-3)
    constexpr int a3 = PH % 3, b3 = (PH + 2) % 3, c3 = (PH + 1) % 3;  // rows r (r-3), r-1, r-2
    constexpr int s0 = PH % 6, s1 = (PH + 5) % 6, s2 = (PH + 4) % 6, s3 = (PH + 3) % 6;
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
        FW[0] = (mk(m0, 0) & M_PW) ? fw[0] : 0.0;
        if (NC == 2) FW[NC - 1] = F0[0];
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const unsigned mc = mk(m0, c);
        const double F = F0[c], U = U0[c], V = V0[c];
        s.TX1[p2][c] = .5 * fabs(U) * (F - FW[c]);                 // :254
        s.TY1[p2][c] = .5 * fabs(V) * (F - F1[c]);                 // :262
        const double qx = (U >= 0.0) ? FW[c] : F;                  // :255-259
        const double qy = (V >= 0.0) ? F1[c] : F;                  // :263-267
        flx[c] = (mc & M_IU) ? U * (qx + posdef) : 0.0;
        s.FLY[p2][c] = (mc & M_IV) ? V * (qy + posdef) : 0.0;
      }
      ediff<NC>(flx, s.DFLX[p2]);
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
      double FDV[NC], FCO[NC], FCN[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double Fc = F1[c];
        // 5-point sea-only extrema of fld (:272-281), then + posdef
        double mx, mn;
        maxmin_first(mx, mn, Fc, Fc, Fw[c], Fw[c], m1, M_PW << (8 * c));
        maxmin_if(mx, mn, Fe[c], Fe[c], m1, M_PE << (8 * c));
        maxmin_if(mx, mn, F2[c], F2[c], m1, M_PS << (8 * c));
        maxmin_if(mx, mn, F0[c], F0[c], m1, M_PN << (8 * c));
        const double MX = mx + posdef, MN = mn + posdef;
        // tsadvc prolog :1934-1938
        const double fdp = ((UE[c] - U1[c]) + (V0[c] - V1[c])) * dt2 * SCI1[c];
        FCO[c] = fmax2(D1[c] + fdp, 0.0);
        FCN[c] = fmax2(D1[c], 0.0);
        // M2 :346-354
        FDV[c] = ((s.DFLX[q2][c]) + (s.FLY[p2][c] - s.FLY[q2][c])) * dt2 * SCI1[c];
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
        const double ay = V1[c] * (FDV[c] + s.FDV[p2][c]);
        const double by = ((FCO[c] + s.FCO[p2][c]) + (FCN[c] + s.FCN[c3][c])) + onemu;
        const double fx = s.TX1[q2][c] - div_flag<SAFE>(ax, bx, SAFE ? 0.0 : rcp_nr(bx), bad);
        const double fy = s.TY1[q2][c] - div_flag<SAFE>(ay, by, SAFE ? 0.0 : rcp_nr(by), bad);
        flx2[c] = (mc & M_IU) ? fx : 0.0;
        s.FLY2[q2][c] = (mc & M_IV) ? fy : 0.0;
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) { s.FLX2[q2][c] = flx2[c]; s.FDV[q2][c] = FDV[c]; s.FCO[q2][c] = FCO[c]; }
      east_of<NC>(flx2, s.FLX2E[q2]);
    }

    // ---- stage C: row r-2 (M4, M5).  FLX2/FLY2 of row r-2 sit in ring slot p2, of row r-1 in q2
    {
      const unsigned m2 = s.m2;
      double SC2[NC];
      ld_own<NC, R::SC>(p, s2, SC2);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double fxc = s.FLX2[p2][c], fxe = s.FLX2E[p2][c];
        const double fyc = s.FLY2[p2][c], fyn = s.FLY2[q2][c];
        const double flxdp = fmin2(0.0, fxe) - fmax2(0.0, fxc);     // :412-415
        const double flxdn = fmax2(0.0, fxe) - fmin2(0.0, fxc);
        const double flydp = fmin2(0.0, fyn) - fmax2(0.0, fyc);
        const double flydn = fmax2(0.0, fyn) - fmin2(0.0, fyc);
        const double w = s.FCN[c3][c] * SC2[c];
        const double ap = (s.MX[c3][c] - s.LO[c3][c]) * w, bp = (onemu - (flxdp + flydp)) * dt2;   // :416-417
        const double am = (s.LO[c3][c] - s.MN[c3][c]) * w, bm = (onemu + (flxdn + flydn)) * dt2;   // :418-419
        s.RP[p2][c] = div_flag<SAFE>(ap, bp, SAFE ? 0.0 : rcp_nr(bp), bad);
        s.RM[p2][c] = div_flag<SAFE>(am, bm, SAFE ? 0.0 : rcp_nr(bm), bad);
      }
      double RPW[NC], RMW[NC], flx3[NC];
      west_of<NC>(s.RP[p2], RPW);
      west_of<NC>(s.RM[p2], RMW);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const unsigned mc = mk(m2, c);
        const double fxc = s.FLX2[p2][c], fyc = s.FLY2[p2][c];
        const double RP = s.RP[p2][c], RM = s.RM[p2][c];
        const double x3 = fmax2(0.0, fxc) * fmin2(fmin2(1.0, RP), RMW[c]) +          // :439-441
                          fmin2(0.0, fxc) * fmin2(fmin2(1.0, RPW[c]), RM);
        const double y3 = fmax2(0.0, fyc) * fmin2(fmin2(1.0, RP), s.RM[q2][c]) +     // :443-445
                          fmin2(0.0, fyc) * fmin2(fmin2(1.0, s.RP[q2][c]), RM);
        flx3[c] = (mc & M_IU) ? x3 : 0.0;
        s.FLY3[p2][c] = (mc & M_IV) ? y3 : 0.0;
      }
      ediff<NC>(flx3, s.DFLX3[p2]);
    }

    // ---- stage E: row r-3, M6 (:475-480) and store
    {
      const int r3 = r - 3;
      double SCI3[NC], OLD3[NC], nv[NC];
      ld_own<NC, R::SCI>(p, s3, SCI3);
      ld_own<NC, R::F>(p, s3, OLD3);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double flxdiv = ((s.DFLX3[q2][c]) + (s.FLY3[p2][c] - s.FLY3[q2][c])) * dt2 * SCI3[c];
        const double b = s.FCN[a3][c] + onemu;
        const double d = div_flag<SAFE>(flxdiv, b, SAFE ? 0.0 : rcp_nr(b), bad);
        const double f = fmax2(s.MN[a3][c], fmin2(s.MX[a3][c], s.LO[a3][c] - d));
        nv[c] = f - posdef;
      }
      const int col = x.w0 + NC * x.lane;
      if ((unsigned)col < (unsigned)x.pitch && r3 >= x.j0 && r3 < x.j1) {
        Vec<NC> old;
#pragma unroll
        for (int c = 0; c < NC; ++c) old.v[c] = OLD3[c];
        store_vec<NC>(x.out, (long)r3 * x.pitch + col, x.lane, s.m3, old, nv);
      }
    }
    s.m3 = s.m2; s.m2 = s.m1; s.m1 = m0;
  }
};

}  // namespace tsadvc
