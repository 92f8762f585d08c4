// TMA-staged marching: the part shared by every advection scheme (FCT2, MPDATA, ...).
//
// Row pipeline: stage A works on row r, B on row r-1, C/D on row r-2, E on row r-3.
// The RAW rows (fld, fldc, uflx, vflx, dp, scp2i, scp2, masks) are not carried in registers:
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
// intermediates stay in register rings indexed by (row mod 2) or (row mod 3); the loop is
// unrolled six times with the phase as a template parameter, so every ring and slot index
// is a compile-time constant and nothing is ever rotated.
// (The tensor-map form cp.async.bulk.tensor / UTMALDG raises "illegal instruction" on this
// pool's B200 boxes even for the CUDA programming guide's own example - tools/probe/ -
// so the rows are fetched with the descriptor-less bulk copy.)
#pragma once
#include "march_common.cuh"
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
  double posdef;         // MPDATA offset (mod_tsadvc.F90:1762)
};

// request row r of every staged array into slot `slot` (one lane)
template <int NC, bool NEED_C, bool NEED_M = true>
__device__ __forceinline__ void issue_row(const TmaCtx& x, int r, int slot) {
  typedef Ring<NC> R;
  const uint32_t bar = x.bar_s + 8u * slot;
  const uint32_t dst = x.ring_s + (uint32_t)(slot * R::SLOT);
  const long off = (long)max(0, min(r, x.nrows - 1)) * x.pitch;
  mbar_expect_tx(bar, R::TX - (NEED_C ? 0 : R::RB) - (NEED_M ? 0 : R::RB));
  bulk_g2s(dst + R::F * R::RB, x.fld + off, R::RB, bar);
  if (NEED_C) bulk_g2s(dst + R::C * R::RB, x.fldc + off, R::RB, bar);
  bulk_g2s(dst + R::U * R::RB, x.u + off, R::RB, bar);
  bulk_g2s(dst + R::V * R::RB, x.v + off, R::RB, bar);
  bulk_g2s(dst + R::D * R::RB, x.dp + off, R::RB, bar);
  bulk_g2s(dst + R::SCI * R::RB, x.sci + off, R::RB, bar);
  bulk_g2s(dst + R::SC * R::RB, x.sc + off, R::RB, bar);
  if (NEED_M) bulk_g2s(dst + R::MSK * R::RB, x.msk + off, R::RB, bar);   // the mask-free bodies never read it
}

// per-lane views of the ring: own columns, west neighbour of the first own column, east
// neighbour of the last own column (clamped inside the row: the clamped lanes are apron)
struct RingPtr {
  const unsigned char *c, *w, *e;
  const unsigned char* w2;   // two columns west of the first own column (advem_fct4)
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
// value of array ARR / mask byte at the column a per-lane pointer (p.w, p.e, p.w2) addresses
template <int NC, int ARR>
__device__ __forceinline__ double ld_at(const unsigned char* q, int slot) {
  typedef Ring<NC> R;
  return *reinterpret_cast<const double*>(q + slot * R::SLOT + ARR * R::RB);
}
template <int NC>
__device__ __forceinline__ unsigned ld_mask_at(const unsigned char* q, int slot) {
  typedef Ring<NC> R;
  return *reinterpret_cast<const unsigned*>(q + slot * R::SLOT + R::MSK * R::RB);
}
template <int NC>
__device__ __forceinline__ unsigned ld_mask_s(const RingPtr& p, int slot) {
  typedef Ring<NC> R;
  const unsigned char* a = p.c + slot * R::SLOT + R::MSK * R::RB;   // low word of the mask plane
  unsigned m = *reinterpret_cast<const unsigned*>(a);
  if (NC == 2) m |= *reinterpret_cast<const unsigned*>(a + 8) << 8;
  return m;
}


// The same accessors for a slot given by its byte offset in the ring at run time (a scheme whose
// register rings all have period 3 is unrolled three times, not six, so the slot of a row is no
// longer a compile-time constant; the offset is warp-uniform).
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
  typedef Ring<NC> R;
  typename S::State s;
  S::init(s);
  bool bad = false;
  const int r0 = x.j0 - 3;
  const int niter = ((x.j1 - x.j0) + 6 + 5) / 6 * 6;   // rows j0-3 .. j1+2, whole rounds of six
  // rows r0-3..r0-1 are "below the chunk": zeros with an all-land mask (never stored)
  {
    double* z = reinterpret_cast<double*>(x.ring + 3 * R::SLOT);
    for (int i = x.lane; i < 3 * R::SLOT / 8; i += 32) z[i] = 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }
  if (elect_one()) {
    issue_row<NC, S::kNeedC, S::kNeedM>(x, r0, 0);
    issue_row<NC, S::kNeedC, S::kNeedM>(x, r0 + 1, 1);
    issue_row<NC, S::kNeedC, S::kNeedM>(x, r0 + 2, 2);
  }
  if constexpr (S::kPeriod == 3) {
    int slot = 0;                        // slot of row r; rows r-1..r-3 sit in slot-1..slot-3 (mod 6)
    uint32_t par = round & 1u;           // parity the barrier of that slot completes next
    SlotOff so{{0}, {5 * R::SLOT}, {4 * R::SLOT}, {3 * R::SLOT}};
#define TSADVC_PHASE3(PH)                                                                 \
  {                                                                                       \
    mbar_wait(x.bar_s + 8u * slot, par);                                                  \
    S::template step<PH, SAFE>(s, x, p, r0 + t + (PH), so, bad);                          \
    __syncwarp();                                                                         \
    const int slot3 = slot >= 3 ? slot - 3 : slot + 3;   /* row r-3: free now */           \
    if (t + (PH) + 3 < niter && elect_one())                                              \
      issue_row<NC, S::kNeedC, S::kNeedM>(x, r0 + t + (PH) + 3, slot3);                              \
    so.s3 = so.s2; so.s2 = so.s1; so.s1 = so.s0;                                          \
    slot = slot == 5 ? 0 : slot + 1;                                                      \
    par ^= (slot == 0) ? 1u : 0u;                                                         \
    so.s0.b = slot * R::SLOT;                                                             \
  }
    for (int t = 0; t < niter; t += 3) { TSADVC_PHASE3(0) TSADVC_PHASE3(1) TSADVC_PHASE3(2) }
#undef TSADVC_PHASE3
    round += (uint32_t)(niter / 6);
    return bad;
  } else {
#define TSADVC_PHASE(PH)                                                                  \
  {                                                                                       \
    mbar_wait(x.bar_s + 8u * (PH), round & 1u); /* row r has landed in slot PH */          \
    S::template step<PH, SAFE>(s, x, p, r0 + t + (PH), bad);                              \
    /* the slot of row r-3 is free now: request row r+3 into it */                        \
    __syncwarp();                                                                         \
    if (t + (PH) + 3 < niter && elect_one())                                              \
      issue_row<NC, S::kNeedC, S::kNeedM>(x, r0 + t + (PH) + 3, ((PH) + 3) % 6);                     \
  }
  for (int t = 0; t < niter; t += 6) {
    TSADVC_PHASE(0) TSADVC_PHASE(1) TSADVC_PHASE(2) TSADVC_PHASE(3) TSADVC_PHASE(4) TSADVC_PHASE(5)
    ++round;
  }
#undef TSADVC_PHASE
  return bad;
  }
}

// the whole chunk again with the compiler's a/b: taken by a warp only when one of its lanes
// met denormal / huge / NaN operands (never on physical data)
template <class S, int NC>
__device__ __noinline__ void march_tma_safe(const TmaCtx x, const RingPtr p, uint32_t round) {
  march_tma_pass<S, NC, true>(x, p, round);
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
  if (x.lane == 0) {
#pragma unroll
    for (int q = 0; q < R::NSLOT; ++q) mbar_init(x.bar_s + 8u * q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t round = 0;
  const bool bad = march_tma_pass<S, NC, false>(x, p, round);
  if (__any_sync(TSADVC_FULLMASK, bad)) march_tma_safe<S, NC>(x, p, round);
}

}  // namespace tsadvc
