// advem_pcm (mod_tsadvc.F90:495-643) + tsadvc prolog (:1905-1942) as a scheme of the
// TMA-staged march: donor-cell fluxes and 5-point extrema on row r (:531-563), the update
// fld = max(fmn, min(fmx, (fld*(fco+onemu) - flxdiv)/(fcn+onemu))) on row r-1 (:610-629).
#pragma once
#include "march_tma_common.cuh"

namespace tsadvc {

template <int NC>
struct PcmT {
  double DFLX[2][NC], FLY[2][NC];   // [row&1]
  unsigned m1;
};

template <int NC, int ISO = 0>
struct PcmScheme {
  typedef PcmT<NC> State;
  static constexpr bool kNeedC = false;
  static constexpr int kPeriod = 6;
  static constexpr int kLag = 1;               // the row finished in iteration r is row r-1
  static constexpr bool kNeedM = true;
  // ISO: the prolog is built from its own pair of mass fluxes (ring arrays U2, V2; isopyc layer 1)
  static constexpr int kArrays = ISO ? 10 : 8;

  static __device__ __forceinline__ void init(State& s) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int q = 0; q < 2; ++q) { s.DFLX[q][c] = 0.0; s.FLY[q][c] = 0.0; }
    s.m1 = 0u;
  }

  template <int PH, bool SAFE>
  static __device__ __forceinline__ void step(State& s, const TmaCtx& x, const RingPtr& p, const int r,
                                              bool& bad) {
    typedef Ring<NC, kArrays> R;
    constexpr int p2 = PH & 1, q2 = p2 ^ 1;
    const Off s0{(PH % 6) * R::SLOT}, s1{((PH + 5) % 6) * R::SLOT}, s2{((PH + 4) % 6) * R::SLOT};
    const double onemu = 9806.e-12;  // :519
    const double dt2 = x.dt2;

    // ---- row r: donor-cell fluxes (:531-547), coast zeroing (:570-590) by select
    const unsigned m0 = ld_mask_s<NC>(p, s0);
    double F0[NC], F1[NC], V0[NC];
    ld_own<NC, R::F>(p, s0, F0);
    ld_own<NC, R::F>(p, s1, F1);
    ld_own<NC, R::V>(p, s0, V0);
    {
      double U0[NC], FW[NC], flx[NC];
      ld_own<NC, R::U>(p, s0, U0);
      ld_west<NC, R::F>(p, s0, F0, FW);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const unsigned mc = mk(m0, c);
        const double qx = (U0[c] >= 0.0) ? FW[c] : F0[c];
        const double qy = (V0[c] >= 0.0) ? F1[c] : F0[c];
        flx[c] = (mc & M_IU) ? U0[c] * qx : 0.0;
        s.FLY[p2][c] = (mc & M_IV) ? V0[c] * qy : 0.0;
      }
      ediff<NC>(flx, s.DFLX[p2]);
    }

    // ---- row r-1: extrema (:548-557), prolog, update (:616-626) and store
    {
      const unsigned m1 = s.m1;
      double Fw[NC], Fe[NC], F2[NC], U1[NC], UE[NC], V1[NC], D1[NC], SCI1[NC], nv[NC];
      ld_west<NC, R::F>(p, s1, F1, Fw);
      ld_east<NC, R::F>(p, s1, F1, Fe);
      ld_own<NC, R::F>(p, s2, F2);
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
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double Fc = F1[c];
        double mx, mn;
        maxmin_first(mx, mn, Fc, Fc, Fw[c], Fw[c], m1, M_PW << (8 * c));
        maxmin_if(mx, mn, Fe[c], Fe[c], m1, M_PE << (8 * c));
        maxmin_if(mx, mn, F2[c], F2[c], m1, M_PS << (8 * c));
        maxmin_if(mx, mn, F0[c], F0[c], m1, M_PN << (8 * c));
        const double fdp = ((UE[c] - U1[c]) + (VP0[c] - V1[c])) * dt2 * SCI1[c];   // :1934-1938
        const double fco = fmax2(D1[c] + fdp, 0.0);
        const double fcn = fmax2(D1[c], 0.0);
        const double flxdiv = ((s.DFLX[q2][c]) + (s.FLY[p2][c] - s.FLY[q2][c])) * dt2 * SCI1[c];
        const double q = Fc * (fco + onemu) - flxdiv;
        const double b = fcn + onemu;
        const double lo = div_flag<SAFE>(q, b, SAFE ? 0.0 : rcp_nr(b), bad);
        nv[c] = fmax2(mn, fmin2(mx, lo));
      }
      store_row_masked<NC>(p, m1, F1, nv);
    }
    s.m1 = m0;
  }
};

}  // namespace tsadvc
