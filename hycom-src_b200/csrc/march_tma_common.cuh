// TMA-staged marching: the part shared by every advection scheme (FCT2, MPDATA, ...).
//
// Row pipeline: stage A works on row r, B on row r-1, C/D on row r-2, E on row r-3.
// The RAW rows (fld, fldc, uflx, vflx, dp, scp2i, scp2, masks) are not carried in registers:
// each warp owns a ring of six row slots in shared memory that the TMA engine fills
// (cp.async.bulk global->shared, SASS UBLKCP: one 256*NC-byte request per array and row,
// completion counted in bytes on one mbarrier per slot).  Row r+3 is requested at the end of
// iteration r, into the slot of row r-3 that iteration r has just finished with, so three rows
// are always in flight and the prefetch distance does not depend on the instruction scheduler.
// The source addresses are eight slab pointers plus one warp-uniform offset that advances by one
// row per request (no per-row index arithmetic).  i-neighbours of raw data are plain shared-memory reads at lane-1 /
// lane+1; j-neighbours are older slots.  Rows outside the slab (apron of the first/last chunk)
// are clamped to the nearest row; the window of the first/last strip may start 4 columns before /
// end after its row, i.e. in the neighbouring row or in the guard row every buffer is allocated
// with: real, finite data that only ever feeds apron lanes (dependency radius 3 < nbdy).  Nothing
// is predicated.  Only computed intermediates stay in register rings indexed by (row mod 2) or
// (row mod 3).
// The finished row is stored straight from registers through one per-lane pointer that advances by
// a row per iteration (predicated 16-byte stores, no branches, no per-row 64-bit index arithmetic).
//
// What was measured on the way (profiles/r02a_*, r02c_*, r02e_*):
//  * the tensor-map form of TMA (cp.async.bulk.tensor, UTMALDG) does work on this pool - round 1's
//    "illegal instruction" came from a box whose first element is not 16-byte aligned in global
//    memory (its probe started at column 3; fp64 boxes must start at an even column, loads and
//    stores alike) - but it is the more expensive way to fetch single rows: per-instruction stall
//    samples put 19 % of the warps' time on the seven UTMALDG of a row and the uniform-register
//    hand-over behind them (635 samples per UTMALDG against 105 per UBLKCP and 57 per FSEL), and
//    the kernel was no faster than with UBLKCP although it issued 9 % fewer instructions;
//  * a TMA store of the finished row needs the strip interior to start at an even column (it
//    starts at w0+3): not possible without giving up 2 of the 58 useful columns.
#pragma once
#include "march_common.cuh"
#include "tsadvc_launch.h"

namespace tsadvc {

// One row slot = eight staged rows of 32*NC doubles: fld(n), fld(m), uflx, vflx, dp(n),
// scp2i, scp2 and the mask word plane of the static block.
template <int NC, int NA = 8>
struct Ring {
  static constexpr int RB = 256 * NC;   // bytes of one staged row of doubles (32*NC columns)
  static constexpr int NARR = NA;       // 7: the mask plane (the last array) is not staged
  static constexpr int SLOT = NARR * RB;
  static constexpr int NSLOT = 6;
  static constexpr int BYTES = NSLOT * SLOT;           // per warp
  // U2, V2: the mass fluxes of the prolog where they differ from the advecting ones (isopyc, layer 1:
  // fco is built from the smoothed fluxes, the tracers are advected by uflx, vflx; NA = 10)
  enum { F = 0, C = 1, U = 2, V = 3, D = 4, SCI = 5, SC = 6, MSK = 7, U2 = 8, V2 = 9 };
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
__device__ __forceinline__ void mbar_inval(uint32_t bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
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
// wait for the row in a slot (one try site); a request that never completes (bad descriptor)
// traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  int spin = 0;
  while (!mbar_try(bar, parity))
    if (++spin > (1 << 16)) __trap();
}
// descriptor-less global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned),
// completion on the mbarrier (SASS UBLKCP; the diffusion kernel still stages its rows this way)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
struct TmaCtx {
  // slabs of this (field, layer): element (row 0, column w0) of each staged array
  const double *fld, *fldc, *u, *v, *dp, *sci, *sc, *msk;
  const double *u2, *v2;   // prolog fluxes (ten-array ring only)
  double* out;           // the output slab (ping-pong buffer) of this layer
  unsigned char* ring;   // this warp's ring (generic pointer into shared memory)
  uint32_t ring_s;       // same, shared-window address
  uint32_t bar_s;        // six mbarriers of this warp
  int pitch, nrows;
  int w0;                // first staged column (even: 16-byte aligned requests)
  int lane;
  int j0, j1;
  double dt2, qdt2x2;
  double posdef;         // MPDATA offset (mod_tsadvc.F90:1762)
};

// the source row of the next request: one warp-uniform element offset that walks down the slabs
struct RowSrc {
  long off;              // offset of row clamp(r, 0, nrows-1) from row 0
  int r;
};
__device__ __forceinline__ RowSrc row_src(const TmaCtx& x, int r) {
  return RowSrc{(long)max(0, min(r, x.nrows - 1)) * x.pitch, r};
}

// request the row of `g` into the slot at byte offset `soff` of the ring (one lane)
template <int NC, bool NEED_C, int NA = 8>
__device__ __forceinline__ void issue_row(const TmaCtx& x, const RowSrc& g, uint32_t soff, uint32_t bar) {
  typedef Ring<NC, NA> R;
  constexpr bool NEED_M = NA >= 8;
  const uint32_t dst = x.ring_s + soff;
  const long off = g.off;
  mbar_expect_tx(bar, R::SLOT - (NEED_C ? 0 : R::RB));
  bulk_g2s(dst + R::F * R::RB, x.fld + off, R::RB, bar);
  if (NEED_C) bulk_g2s(dst + R::C * R::RB, x.fldc + off, R::RB, bar);
  bulk_g2s(dst + R::U * R::RB, x.u + off, R::RB, bar);
  bulk_g2s(dst + R::V * R::RB, x.v + off, R::RB, bar);
  bulk_g2s(dst + R::D * R::RB, x.dp + off, R::RB, bar);
  bulk_g2s(dst + R::SCI * R::RB, x.sci + off, R::RB, bar);
  bulk_g2s(dst + R::SC * R::RB, x.sc + off, R::RB, bar);
  if (NEED_M) bulk_g2s(dst + R::MSK * R::RB, x.msk + off, R::RB, bar);   // the mask-free bodies never read it
  if (NA == 10) {
    bulk_g2s(dst + R::U2 * R::RB, x.u2 + off, R::RB, bar);
    bulk_g2s(dst + R::V2 * R::RB, x.v2 + off, R::RB, bar);
  }
}
// step to the next row (every lane: the offset stays warp-uniform); rows outside the slab repeat the
// nearest one
__device__ __forceinline__ void next_row(const TmaCtx& x, RowSrc& g) {
  g.off += ((unsigned)g.r < (unsigned)(x.nrows - 1)) ? (long)x.pitch : 0L;
  g.r += 1;
}

// per-lane views of the ring: own columns, west neighbour of the first own column, east
// neighbour of the last own column (clamped inside the row: the clamped lanes are apron)
struct RingPtr {
  const unsigned char *c, *w, *e;
  const unsigned char* w2;   // two columns west of the first own column (advem_fct4)
  // output: this lane's first cell in the row being finished, and what the lane may store there
  double* outp;
  bool rowok;                // the finished row lies inside j0..j1-1 (warp-uniform)
  bool st2, stx, sty;        // both cells / only the first / only the second belong to the strip interior
};

// The finished row: new values on the cells tsadvc writes (M_OUT), the old value everywhere else
// (land, halo ring), so the ping-pong slab is complete.  Only the strip interior (columns w0+3 ..
// w0+32*NC-4, inside the slab) is written by this warp: lanes with both cells inside store 16 bytes,
// the two edge lanes 8.
template <int NC>
__device__ __forceinline__ void store_cells(const RingPtr& p, const double (&o)[NC]) {
  if (p.rowok) {
    if (NC == 2) {
      if (p.st2) asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p.outp), "d"(o[0]), "d"(o[NC - 1]) : "memory");
      if (p.stx) asm volatile("st.global.f64 [%0], %1;" ::"l"(p.outp), "d"(o[0]) : "memory");
      if (p.sty) asm volatile("st.global.f64 [%0+8], %1;" ::"l"(p.outp), "d"(o[NC - 1]) : "memory");
    } else {
      if (p.stx) asm volatile("st.global.f64 [%0], %1;" ::"l"(p.outp), "d"(o[0]) : "memory");
    }
  }
}
template <int NC>
__device__ __forceinline__ void store_row_masked(const RingPtr& p, unsigned m, const double (&old)[NC],
                                                 const double (&nv)[NC]) {
  double o[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) o[c] = (mk(m, c) & M_OUT) ? nv[c] : old[c];
  store_cells<NC>(p, o);
}

// Accessors of the staged rows; a slot is named by its byte offset in the ring (a compile-time constant
// for a scheme unrolled six times, a warp-uniform register for one unrolled three times).
struct Off { int b; };
struct SlotOff { Off s0, s1, s2, s3; };   // slots of rows r, r-1, r-2, r-3
template <int NC, int ARR>
__device__ __forceinline__ void ld_own(const RingPtr& p, Off o, double (&x)[NC]) {
  typedef Ring<NC> R;
  const unsigned char* a = p.c + o.b + ARR * R::RB;
  if (NC == 2) {
    const double2 v = *reinterpret_cast<const double2*>(a);
    x[0] = v.x; x[NC - 1] = v.y;
  } else {
    x[0] = *reinterpret_cast<const double*>(a);
  }
}
template <int NC, int ARR>
__device__ __forceinline__ void ld_west(const RingPtr& p, Off o, const double (&own)[NC], double (&w)[NC]) {
  typedef Ring<NC> R;
  w[0] = *reinterpret_cast<const double*>(p.w + o.b + ARR * R::RB);
  if (NC == 2) w[NC - 1] = own[0];
}
template <int NC, int ARR>
__device__ __forceinline__ void ld_east(const RingPtr& p, Off o, const double (&own)[NC], double (&e)[NC]) {
  typedef Ring<NC> R;
  e[NC - 1] = *reinterpret_cast<const double*>(p.e + o.b + ARR * R::RB);
  if (NC == 2) e[0] = own[NC - 1];
}
template <int NC, int ARR>
__device__ __forceinline__ double ld_at(const unsigned char* q, Off o) {
  typedef Ring<NC> R;
  return *reinterpret_cast<const double*>(q + o.b + ARR * R::RB);
}
template <int NC>
__device__ __forceinline__ unsigned ld_mask_at(const unsigned char* q, Off o) {
  typedef Ring<NC> R;
  return *reinterpret_cast<const unsigned*>(q + o.b + R::MSK * R::RB);
}
template <int NC>
__device__ __forceinline__ unsigned ld_mask_s(const RingPtr& p, Off o) {
  typedef Ring<NC> R;
  const unsigned char* a = p.c + o.b + R::MSK * R::RB;
  unsigned m = *reinterpret_cast<const unsigned*>(a);
  if (NC == 2) m |= *reinterpret_cast<const unsigned*>(a + 8) << 8;
  return m;
}

// ---------------------------------------------------------------------------------------
// the march of one (field, layer, strip, chunk) unit for a scheme S:
//   S::State          computed intermediates (register rings)
//   S::kNeedC         fldc is staged
//   S::init(State&)   rows below the chunk
//   S::kPeriod        6: S::step<PH,SAFE>(State&, ctx, ringptr, r, bad), phase PH = row mod 6, ring
//                        slots are compile-time constants;
//                     3: S::step<PH,SAFE>(State&, ctx, ringptr, r, SlotOff, bad), PH = row mod 3, the
//                        slots rotate at run time.  Half the code: the six-fold unrolled loop of
//                        FCT2 (64 KB of SASS) already missed in the instruction cache (3 % of the
//                        stall samples), and two copies of it (the all-sea and the general row body)
//                        thrash it (21 %, profiles/r01z); unrolled three times, both bodies fit.
//                        (Period 2 would need register moves: a value lives three rows.)
// ---------------------------------------------------------------------------------------
template <class S, int NC, bool SAFE>
__device__ __forceinline__ bool march_tma_pass(const TmaCtx& x, const RingPtr& p, uint32_t& round) {
  typedef Ring<NC, S::kArrays> R;
  typename S::State s;
  S::init(s);
  bool bad = false;
  const int r0 = x.j0 - 3;
  const int niter = ((x.j1 - x.j0) + 6 + 5) / 6 * 6;   // rows j0-3 .. j1+2, whole rounds of six
  const int nstore = x.j1 - x.j0;
  // rows r0-3..r0-1 are "below the chunk": zeros with an all-land mask (never stored)
  {
    double* z = reinterpret_cast<double*>(x.ring + 3 * R::SLOT);
    for (int i = x.lane; i < 3 * R::SLOT / 8; i += 32) z[i] = 0.0;
    fence_proxy_async();
    __syncwarp();
  }
  RowSrc g = row_src(x, r0);
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    if (elect_one()) issue_row<NC, S::kNeedC, S::kArrays>(x, g, q * R::SLOT, x.bar_s + 8u * q);
    next_row(x, g);
  }
  // this lane's output pointer walks down the slab one row per iteration (row r - kLag)
  RingPtr q = p;
  q.outp = x.out + ((long)(r0 - S::kLag) * x.pitch + x.w0 + NC * x.lane);
  const long ostep = x.pitch;
  // after step(row r): request row r+3 into the slot row r-3 has left (no end-of-chunk test: the three
  // rows past the chunk are drained below)
#define TSADVC_ROW_TAIL(ROW, SOFF3, BAR3)                                                      \
  __syncwarp();                                                                              \
  if (elect_one()) issue_row<NC, S::kNeedC, S::kArrays>(x, g, SOFF3, BAR3);                   \
  next_row(x, g);                                                      \
  q.outp += ostep;
#define TSADVC_ROW_HEAD(ROW) q.rowok = (unsigned)((ROW) - S::kLag - x.j0) < (unsigned)nstore;
  if constexpr (S::kPeriod == 3) {
    // The loop is unrolled three times over a ring of six slots: in one trip the rows r sit in slots
    // PH + 3*half, the older rows partly in the other half.  Two warp-uniform byte offsets (hb: this
    // half, hc: the other) that swap once per trip make every slot address "register + constant".
    uint32_t hb = 0, hc = 3 * R::SLOT;         // ring offsets of slot 0 / 3 of the two halves
    uint32_t bb = x.bar_s, bc = x.bar_s + 24;  // mbarriers of the two halves
    uint32_t par = round & 1u;                 // parity the barriers of this round complete with
#define TSADVC_PHASE3(PH)                                                                     \
  {                                                                                           \
    const SlotOff so{{(int)(hb + (PH) * R::SLOT)},                                            \
                     {(int)((PH) >= 1 ? hb + ((PH) - 1) * R::SLOT : hc + 2 * R::SLOT)},       \
                     {(int)((PH) == 2 ? hb : hc + ((PH) + 1) * R::SLOT)},                     \
                     {(int)(hc + (PH) * R::SLOT)}};                                           \
    TSADVC_ROW_HEAD(r0 + t + (PH))                                                            \
    mbar_wait(bb + 8u * (PH), par);                                                           \
    S::template step<PH, SAFE>(s, x, q, r0 + t + (PH), so, bad);                              \
    TSADVC_ROW_TAIL(r0 + t + (PH), hc + (PH) * R::SLOT, bc + 8u * (PH))                       \
  }
    for (int t = 0; t < niter; t += 3) {
      TSADVC_PHASE3(0) TSADVC_PHASE3(1) TSADVC_PHASE3(2)
      par ^= (hb != 0) ? 1u : 0u;               // the second half closes a round of six
      const uint32_t th = hb; hb = hc; hc = th;
      const uint32_t tb = bb; bb = bc; bc = tb;
    }
#undef TSADVC_PHASE3
    round += (uint32_t)(niter / 6);
  } else {
#define TSADVC_PHASE(PH)                                                                  \
  {                                                                                       \
    TSADVC_ROW_HEAD(r0 + t + (PH))                                                        \
    mbar_wait(x.bar_s + 8u * (PH), round & 1u); /* row r has landed in slot PH */          \
    S::template step<PH, SAFE>(s, x, q, r0 + t + (PH), bad);                              \
    TSADVC_ROW_TAIL(r0 + t + (PH), (((PH) + 3) % 6) * R::SLOT, x.bar_s + 8u * (((PH) + 3) % 6)) \
  }
    for (int t = 0; t < niter; t += 6) {
      TSADVC_PHASE(0) TSADVC_PHASE(1) TSADVC_PHASE(2) TSADVC_PHASE(3) TSADVC_PHASE(4) TSADVC_PHASE(5)
      ++round;
    }
#undef TSADVC_PHASE
  }
#undef TSADVC_ROW_TAIL
#undef TSADVC_ROW_HEAD
  // rows niter .. niter+2 were requested past the end of the chunk (no test in the loop): let them land
  // before the ring is reused; the three barriers are then one phase ahead of the other three
  mbar_wait(x.bar_s, round & 1u);
  mbar_wait(x.bar_s + 8, round & 1u);
  mbar_wait(x.bar_s + 16, round & 1u);
  return bad;
}

// the whole chunk again with the compiler's a/b: taken by a warp only when one of its lanes
// met denormal / huge / NaN operands (never on physical data)
template <class S, int NC>
__device__ __noinline__ void march_tma_safe(const TmaCtx x, const RingPtr p, uint32_t round) {
  march_tma_pass<S, NC, true>(x, p, round);
}

template <int NC>
__device__ __forceinline__ void ring_barriers_init(const TmaCtx& x, bool again) {
  typedef Ring<NC> R;
  if (x.lane == 0) {
#pragma unroll
    for (int q = 0; q < R::NSLOT; ++q) {
      if (again) mbar_inval(x.bar_s + 8u * q);
      mbar_init(x.bar_s + 8u * q, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}

template <class S, int NC>
__device__ void march_tma(const TmaCtx& x) {
  typedef Ring<NC> R;
  RingPtr p;
  const int l0 = x.lane * NC;
  p.c = x.ring + 8 * l0;
  p.w = x.ring + 8 * max(l0 - 1, 0);
  p.e = x.ring + 8 * min(l0 + NC, 32 * NC - 1);
  p.w2 = x.ring + 8 * max(l0 - 2, 0);
  {  // which of this lane's cells belong to the strip interior and to the slab
    const int lo = kApron, hi = 32 * NC - kApron;                  // interior columns [lo, hi) of the window
    const bool in0 = l0 >= lo && l0 < hi && (unsigned)(x.w0 + l0) < (unsigned)x.pitch;
    const bool in1 = NC == 2 && l0 + 1 >= lo && l0 + 1 < hi && (unsigned)(x.w0 + l0 + 1) < (unsigned)x.pitch;
    p.st2 = in0 && in1; p.stx = in0 && !in1; p.sty = in1 && !in0;
    p.outp = nullptr; p.rowok = false;
  }
  ring_barriers_init<NC>(x, false);
  uint32_t round = 0;
  const bool bad = march_tma_pass<S, NC, false>(x, p, round);
  if (__any_sync(TSADVC_FULLMASK, bad)) {
    ring_barriers_init<NC>(x, true);
    march_tma_safe<S, NC>(x, p, 0u);
  }
}

}  // namespace tsadvc
