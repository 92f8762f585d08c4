// advem_fct2 (mod_tsadvc.F90:645-997) + tsadvc prolog (:1905-1942), TMA-staged marching.
//
// Same row pipeline as march_fct2.cuh (stage A row r, B row r-1, C/D row r-2, E row r-3) but
// the RAW rows (fld, fldc, uflx, vflx, dp, scp2i, scp2, masks) are not carried in registers:
// each warp owns a ring of six row slots in shared memory that the TMA engine fills
// (cp.async.bulk global->shared, SASS UBLKCP: one 256*NC-byte request per array and row,
// completion counted in bytes on one mbarrier per slot).  Row r+3 is requested at the end of iteration r, into the slot of row r-3 that
// iteration r has just finished with, so three rows are always in flight and the prefetch
// distance does not depend on the instruction scheduler.  i-neighbours of raw data are
// plain shared-memory reads at lane-1 / lane+1; j-neighbours are older slots.  Rows outside
// the slab (apron of the first/last chunk) are clamped to the nearest row; the window of the
// first/last strip may start 4 columns before / end after its row, i.e. in the neighbouring
// row or in the guard row every buffer is allocated with: real, finite data that only ever
// feeds apron lanes (dependency radius 3 < nbdy).  Nothing is predicated.  Only computed
// intermediates stay in the register rings (about 40 registers less than march_fct2.cuh).
// (The tensor-map form cp.async.bulk.tensor / UTMALDG raises "illegal instruction" on this
// pool's B200 boxes even for the CUDA programming guide's own example - tools/probe/ -
// so the rows are fetched with the descriptor-less bulk copy.)
#pragma once
#include "march_common.cuh"
#include "march_fct2.cuh"
#include "tsadvc_launch.h"

namespace tsadvc {

// One row slot = eight staged rows of 32*NC doubles: fld(n), fld(m), uflx, vflx, dp(n),
// scp2i, scp2 and the mask word plane of the static block.
template <int NC>
struct Ring {
  static constexpr int RB = 256 * NC;   // bytes of one staged row of doubles (32*NC columns)
  static constexpr int NARR = 8;
  static constexpr int SLOT = NARR * RB;
  static constexpr int NSLOT = 6;
  static constexpr int BYTES = NSLOT * SLOT;           // per warp
  static constexpr int TX = SLOT;                      // bytes the requests of one row deliver
  enum { F = 0, C = 1, U = 2, V = 3, D = 4, SCI = 5, SC = 6, MSK = 7 };
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// one lane of the (converged) warp; the compiler recognises elect.sync as the guard of a
// uniform-datapath instruction and emits UTMALDG without a per-thread serialisation loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}"
               : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// wait for the row in a slot; a request that never completes (bad descriptor) traps instead
// of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  for (int spin = 0; !mbar_try(bar, parity); ++spin)
    if (spin > (1 << 16)) __trap();
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned),
// completion on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

struct TmaCtx {
  // slabs of this (field, layer): element (row 0, column w0) of each staged array
  const double *fld, *fldc, *u, *v, *dp, *sci, *sc, *msk;
  double* __restrict__ out;
  unsigned char* ring;   // this warp's ring (generic pointer into shared memory)
  uint32_t ring_s;       // same, shared-window address
  uint32_t bar_s;        // six mbarriers of this warp
  int pitch, nrows;
  int w0;                // first staged column (even: 16-byte aligned requests)
  int lane;
  int j0, j1;
  double dt2, qdt2x2;
};

// request row r of every staged array into slot `slot` (one lane)
template <int NC>
__device__ __forceinline__ void issue_row(const TmaCtx& x, int r, int slot) {
  typedef Ring<NC> R;
  const uint32_t bar = x.bar_s + 8u * slot;
  const uint32_t dst = x.ring_s + (uint32_t)(slot * R::SLOT);
  const long off = (long)max(0, min(r, x.nrows - 1)) * x.pitch;
  mbar_expect_tx(bar, R::TX);
  bulk_g2s(dst + R::F * R::RB, x.fld + off, R::RB, bar);
  bulk_g2s(dst + R::C * R::RB, x.fldc + off, R::RB, bar);
  bulk_g2s(dst + R::U * R::RB, x.u + off, R::RB, bar);
  bulk_g2s(dst + R::V * R::RB, x.v + off, R::RB, bar);
  bulk_g2s(dst + R::D * R::RB, x.dp + off, R::RB, bar);
  bulk_g2s(dst + R::SCI * R::RB, x.sci + off, R::RB, bar);
  bulk_g2s(dst + R::SC * R::RB, x.sc + off, R::RB, bar);
  bulk_g2s(dst + R::MSK * R::RB, x.msk + off, R::RB, bar);
}

template <int NC>
struct Fct2T {                                       // computed intermediates only
  double DFLX[2][NC], FLY[2][NC];                    // flx(i+1)-flx(i), fly              [row&1]
  double FAX[3][NC], FAY[3][NC];                     // antidiffusive fluxes              [row%3]
  double LO[3][NC], FCN[3][NC], Y[3][NC];            // fldlo, fcn, 1/(fcn+onemu)         [row%3]
  double MXL[3][NC], MNL[3][NC];                     // fmxlo, fmnlo                      [row%3]
  double RP[2][NC], RM[2][NC];                       //                                   [row&1]
  double QMX[2][NC], QMN[2][NC];                     // fmx, fmn of S4                    [row&1]
  double DFAXL[2][NC], FAYL[2][NC];                  // limited fluxes                    [row&1]
  unsigned m1, m2, m3;                               // masks of rows r-1, r-2, r-3
};

// per-lane views of the ring: own columns, west neighbour of the first own column, east
// neighbour of the last own column (clamped inside the row: the clamped lanes are apron)
struct RingPtr {
  const unsigned char *c, *w, *e;
};

template <int NC, int ARR>
__device__ __forceinline__ void ld_own(const RingPtr& p, int slot, double (&x)[NC]) {
  typedef Ring<NC> R;
  const unsigned char* a = p.c + slot * R::SLOT + ARR * R::RB;
  if (NC == 2) {
    const double2 v = *reinterpret_cast<const double2*>(a);
    x[0] = v.x; x[NC - 1] = v.y;
  } else {
    x[0] = *reinterpret_cast<const double*>(a);
  }
}
template <int NC, int ARR>
__device__ __forceinline__ void ld_west(const RingPtr& p, int slot, const double (&own)[NC],
                                        double (&w)[NC]) {
  typedef Ring<NC> R;
  w[0] = *reinterpret_cast<const double*>(p.w + slot * R::SLOT + ARR * R::RB);
  if (NC == 2) w[NC - 1] = own[0];
}
template <int NC, int ARR>
__device__ __forceinline__ void ld_east(const RingPtr& p, int slot, const double (&own)[NC],
                                        double (&e)[NC]) {
  typedef Ring<NC> R;
  e[NC - 1] = *reinterpret_cast<const double*>(p.e + slot * R::SLOT + ARR * R::RB);
  if (NC == 2) e[0] = own[NC - 1];
}
template <int NC>
__device__ __forceinline__ unsigned ld_mask_s(const RingPtr& p, int slot) {
  typedef Ring<NC> R;
  const unsigned char* a = p.c + slot * R::SLOT + R::MSK * R::RB;   // low word of the mask plane
  unsigned m = *reinterpret_cast<const unsigned*>(a);
  if (NC == 2) m |= *reinterpret_cast<const unsigned*>(a + 8) << 8;
  return m;
}

template <int NC, int PH, bool SAFE>
__device__ __forceinline__ void fct2t_step(Fct2T<NC>& s, const TmaCtx& x, const RingPtr& p,
                                           const int r, const bool more, const uint32_t parity,
                                           bool& bad) {
  typedef Ring<NC> R;
  constexpr int p2 = PH & 1, q2 = p2 ^ 1;
  constexpr int a3 = PH % 3, b3 = (PH + 2) % 3, c3 = (PH + 1) % 3;
  constexpr int s0 = PH % 6, s1 = (PH + 5) % 6, s2 = (PH + 4) % 6, s3 = (PH + 3) % 6;  // rows r..r-3
  const double onemu = 9806.e-12;  // :671
  const double dt2 = x.dt2;

  mbar_wait(x.bar_s + 8u * s0, parity);   // row r has landed

  // ---- stage A: row r
  double F0[NC], F1[NC], C1[NC];
  ld_own<NC, R::F>(p, s0, F0);
  ld_own<NC, R::F>(p, s1, F1);
  ld_own<NC, R::C>(p, s1, C1);
  const unsigned m0 = ld_mask_s<NC>(p, s0);
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
      flx[c] = (mc & M_IU) ? U * qx : 0.0;
      const double fly = (mc & M_IV) ? V * qy : 0.0;
      const double fhx = U * 0.5 * (C + CW[c]);                   // :824
      const double fhy = V * 0.5 * (C + C1[c]);                   // :828
      s.FAX[a3][c] = (mc & M_IU) ? fhx - flx[c] : 0.0;
      s.FAY[a3][c] = (mc & M_IV) ? fhy - fly : 0.0;
      s.FLY[p2][c] = fly;
    }
    ediff<NC>(flx, s.DFLX[p2]);
  }

  // ---- stage B: row r-1
  {
    const unsigned m1 = s.m1;
    double Fw[NC], Fe[NC], F2[NC], U1[NC], UE[NC], V1[NC], D1[NC], SCI1[NC];
    ld_west<NC, R::F>(p, s1, F1, Fw);
    ld_east<NC, R::F>(p, s1, F1, Fe);
    ld_own<NC, R::F>(p, s2, F2);
    ld_own<NC, R::U>(p, s1, U1);
    ld_east<NC, R::U>(p, s1, U1, UE);
    ld_own<NC, R::V>(p, s1, V1);
    ld_own<NC, R::D>(p, s1, D1);
    ld_own<NC, R::SCI>(p, s1, SCI1);
    double q[NC], b[NC], y[NC], fmx[NC], fmn[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double Fc = F1[c];
      // 5-point sea-only extrema of fld (:709-716)
      double mx, mn;
      maxmin_first(mx, mn, Fc, Fc, Fw[c], Fw[c], m1, M_PW << (8 * c));
      maxmin_if(mx, mn, Fe[c], Fe[c], m1, M_PE << (8 * c));
      maxmin_if(mx, mn, F2[c], F2[c], m1, M_PS << (8 * c));
      maxmin_if(mx, mn, F0[c], F0[c], m1, M_PN << (8 * c));
      fmx[c] = mx; fmn[c] = mn;
      // tsadvc prolog :1934-1938 (onetamas(:,:,m) = 1.0 when .not.btrmas, :1809)
      const double fdp = ((UE[c] - U1[c]) + (V0[c] - V1[c])) * dt2 * SCI1[c];
      const double Dc = D1[c];
      const double fco = pos_part(Dc + fdp);
      const double fcn = pos_part(Dc);
      // :786-793
      const double flxdiv = ((s.DFLX[q2][c]) + (s.FLY[p2][c] - s.FLY[q2][c])) * dt2 * SCI1[c];
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
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const unsigned mc = mk(m2, c);
      const bool pe = mc & M_PE, pn = mc & M_PN;
      // 5-point sea-only extrema of fmxlo/fmnlo (:872-879)
      double fqmax, fqmin;
      maxmin_first(fqmax, fqmin, s.MXL[c3][c], s.MNL[c3][c], mxw[c], mnw[c], m2, M_PW << (8 * c));
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
      const double qp2 = (fqmax - lo) * fcn * SC2[c] * x.qdt2x2;   // 2*qp  :885
      const double qm2 = (lo - fqmin) * fcn * SC2[c] * x.qdt2x2;   // 2*qm  :895
      qq[2 * c] = qp2;     bb[2 * c] = (famax2 > 0.0) ? famax2 : 1.0;
      qq[2 * c + 1] = qm2; bb[2 * c + 1] = (famin2 > 0.0) ? famin2 : 1.0;
      s.QMX[p2][c] = fqmax;                                 // :904
      s.QMN[p2][c] = fqmin;                                 // :905
    }
#pragma unroll
    for (int i = 0; i < 2 * NC; ++i)
      rr[i] = div_flag<SAFE>(qq[i], bb[i], SAFE ? 0.0 : rcp_nr(bb[i]), bad);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      // :884-903: rp = famax>0 ? (qp<famax ? qp/famax : 1) : 0 ; see march_fct2.cuh for fa==0
      s.RP[p2][c] = min_one(rr[2 * c]);
      s.RM[p2][c] = min_one(rr[2 * c + 1]);
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
  {
    const int r3 = r - 3;
    double SCI3[NC], OLD3[NC], nv[NC];
    ld_own<NC, R::SCI>(p, s3, SCI3);
    ld_own<NC, R::F>(p, s3, OLD3);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double a = ((s.DFAXL[q2][c]) + (s.FAYL[p2][c] - s.FAYL[q2][c])) * dt2 * SCI3[c];
      const double d = div_flag<SAFE>(a, s.FCN[a3][c] + onemu, s.Y[a3][c], bad);
      nv[c] = fmax2(s.QMN[q2][c], fmin2(s.QMX[q2][c], s.LO[a3][c] - d));
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

  // slot s3 (row r-3) is free now: request row r+3 into it
  __syncwarp();
  if (more && elect_one()) issue_row<NC>(x, r + 3, s3);
}

template <int NC, bool SAFE>
__device__ __forceinline__ bool march_fct2_tma_pass(const TmaCtx& x, const RingPtr& p, uint32_t& round) {
  Fct2T<NC> s;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      s.FAX[q][c] = 0.0; s.FAY[q][c] = 0.0; s.LO[q][c] = 0.0; s.FCN[q][c] = 0.0;
      s.Y[q][c] = 1.0; s.MXL[q][c] = 0.0; s.MNL[q][c] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      s.DFLX[q][c] = 0.0; s.FLY[q][c] = 0.0; s.RP[q][c] = 0.0; s.RM[q][c] = 0.0;
      s.QMX[q][c] = 0.0; s.QMN[q][c] = 0.0; s.DFAXL[q][c] = 0.0; s.FAYL[q][c] = 0.0;
    }
  }
  s.m1 = s.m2 = s.m3 = 0u;
  bool bad = false;
  const int r0 = x.j0 - 3;
  const int niter = ((x.j1 - x.j0) + 6 + 5) / 6 * 6;   // rows j0-3 .. j1+2, whole rounds of six
  // rows r0-3..r0-1 are "below the chunk": zeros with an all-land mask (never stored)
  {
    typedef Ring<NC> R;
    double* z = reinterpret_cast<double*>(x.ring + 3 * R::SLOT);
    for (int i = x.lane; i < 3 * R::SLOT / 8; i += 32) z[i] = 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }
  if (elect_one()) {
    issue_row<NC>(x, r0, 0);
    issue_row<NC>(x, r0 + 1, 1);
    issue_row<NC>(x, r0 + 2, 2);
  }
  for (int t = 0; t < niter; t += 6) {
    const int r = r0 + t;
    const uint32_t par = round & 1u;
    fct2t_step<NC, 0, SAFE>(s, x, p, r, t + 3 < niter, par, bad);
    fct2t_step<NC, 1, SAFE>(s, x, p, r + 1, t + 4 < niter, par, bad);
    fct2t_step<NC, 2, SAFE>(s, x, p, r + 2, t + 5 < niter, par, bad);
    fct2t_step<NC, 3, SAFE>(s, x, p, r + 3, t + 6 < niter, par, bad);
    fct2t_step<NC, 4, SAFE>(s, x, p, r + 4, t + 7 < niter, par, bad);
    fct2t_step<NC, 5, SAFE>(s, x, p, r + 5, t + 8 < niter, par, bad);
    ++round;
  }
  return bad;
}

template <int NC>
__device__ __noinline__ void march_fct2_tma_safe(const TmaCtx x, const RingPtr p, uint32_t round) {
  march_fct2_tma_pass<NC, true>(x, p, round);
}

template <int NC>
__device__ void march_fct2_tma(const TmaCtx& x) {
  typedef Ring<NC> R;
  RingPtr p;
  const int l0 = x.lane * NC;
  p.c = x.ring + 8 * l0;
  p.w = x.ring + 8 * max(l0 - 1, 0);
  p.e = x.ring + 8 * min(l0 + NC, 32 * NC - 1);
  if (x.lane == 0) {
#pragma unroll
    for (int q = 0; q < R::NSLOT; ++q) mbar_init(x.bar_s + 8u * q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t round = 0;
  const bool bad = march_fct2_tma_pass<NC, false>(x, p, round);
  if (__any_sync(TSADVC_FULLMASK, bad)) march_fct2_tma_safe<NC>(x, p, round);
}

}  // namespace tsadvc
