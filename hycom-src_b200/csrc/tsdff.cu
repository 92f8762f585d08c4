// Diffusion of the thermodynamic variables and tracers after advection, and the sweep that
// re-derives the non-independent variable (mod_tsadvc.F90:2138-2230: tsdff_2x :2262-2345,
// tsdff_1x :2347-2492, EOS sweep :2199-2229), fused: the reference runs, per layer and field
// pair, one sweep that stores the face fluxes uflux/vflux(/uflux2/vflux2), one sweep that
// applies their divergence, and at the end one sweep over all layers for the equation of
// state.  Here the face factors of a cell are formed once (they depend on dp only and are
// shared by every field of the launch group), every field of the group is diffused and - for
// the T/S/th3d group - the equation of state is applied to the fresh values before the single
// store.
//
// HBM bound by design: per layer-cell it reads dp(n) and each field once and writes each field
// once: T,S,th3d group 7 x 8 B = 56 B, a tracer pair 8 + 32 B.
//
// Arithmetic is the Fortran's, expression by expression (-fmad=false, IEEE division):
//   factor_u = temdf2*aspux(i,j)*scuy(i,j)*harmonc(dp(i-1)*onetamas(i-1), dp(i)*onetamas(i))
//   uflux(i) = factor_u*(fld(i-1)-fld(i))                       at iu points, else 0.0
//   factor   = -delt1/(scp2*max(dp*onetamas, eps_har))
//   fld     += ((uflux(i+1)-uflux(i))+(vflux(j+1)-vflux(j)))*factor
// Land faces carry 0.0: uflux/vflux are zeroed at mod_tsadvc.F90:1812-1813, uflux2/vflux2 on
// the faces bounding every sea segment by geopar.F90:826-843, and tsdff only writes iu/iv
// points.  The output goes to the ping-pong buffer (neighbours read the old values).
#include <cuda_runtime.h>

#include "eos.cuh"
#include "march_tma_common.cuh"   // div_rn (exact a/b), mbarrier / bulk-copy helpers of the TMA-staged march
#include "tsadvc_dev.h"
#include "tsadvc_launch.h"

namespace tsadvc {

namespace {

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

// harmonc(aa,bb) = harmonz(max(aa,0),max(bb,0)); harmonz(a,b) = 2.0*a*b/max((a+b),2.0*eps_har)
// (mod_tsadvc.F90:1770-1771)
__device__ __forceinline__ double harmonc(double aa, double bb) {
  const double eps_har = 1.0e-20;   // :1763
  const double a = dmax(aa, 0.0), b = dmax(bb, 0.0);
  return div_rn(2.0 * a * b, dmax((a + b), 2.0 * eps_har));
}

// isopycnic coordinates, layer 1: lateral smoothing of the mixed-layer mass fluxes
// (mod_tsadvc.F90:1859-1897, margin mbdy-1); faces that are not iu/iv points and everything
// outside the margin keep the 0.0 of :1812-1813
__global__ void __launch_bounds__(256) k_isopyc_smooth(const double* __restrict__ u, const double* __restrict__ v,
                                                         double* __restrict__ us, double* __restrict__ vs,
                                                         const uint8_t* __restrict__ mask, int pitch, int nrows,
                                                         int nbdy, int ii, int jj, int margin) {
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
  if (c >= pitch || r >= nrows) return;
  const long q = (long)r * pitch + c;
  const int i = c + 1 - nbdy, j = r + 1 - nbdy;
  double uo = 0.0, vo = 0.0;
  if (i >= 1 - margin && i <= ii + margin && j >= 1 - margin && j <= jj + margin) {
    const unsigned m = mask[q];
    if (m & M_IV) {
      const double vfa = (mask[q - 1] & M_IV) ? v[q - 1] : v[q];
      const double vfb = (mask[q + 1] & M_IV) ? v[q + 1] : v[q];
      vo = .5 * v[q] + .25 * (vfa + vfb);
    }
    if (m & M_IU) {
      const double ufa = (mask[q - pitch] & M_IU) ? u[q - pitch] : u[q];
      const double ufb = (mask[q + pitch] & M_IU) ? u[q + pitch] : u[q];
      uo = .5 * u[q] + .25 * (ufa + ufb);
    }
  }
  us[q] = uo;
  vs[q] = vo;
}


// ---------------------------------------------------------------------------------------------
// One warp owns a strip of 32 columns of one layer and
// walks along j; the rows of every operand are staged through a shared-memory ring by the TMA
// engine (cp.async.bulk + one mbarrier per slot, three rows in flight) exactly like the advection
// (march_tma_common.cuh).  Marching lets neighbouring cells share their face factors: the north
// face of row r-1 is the south face of row r (carried in a register), the east face of column i
// is the west face of column i+1 (one shuffle) - two harmonc divisions per cell instead of four -
// and every operand is read from HBM once.  Dependency radius 1: a strip yields 30 columns.
// Two cells per lane were measured (profiles/r03b-h): 18-24 % fewer instructions per cell, but with all nine
// operands staged the ring allows two warps per scheduler instead of four (5.7 ms against 3.8 at kdm=12),
// and with the five 2-D operands fetched by plain loads one row ahead ptxas copies the freshly loaded
// registers at the loop head, so every row waits out a full load latency (4.1 ms): not adopted.
//   ring slot = [fld_0 .. fld_NF-1 | dp | oneta | au | av | scp2 | mask words], 256 B each
// ---------------------------------------------------------------------------------------------
template <int NF>
struct DRing {
  static constexpr int RB = 256, NARR = NF + 6, SLOT = NARR * RB, NSLOT = 6, BYTES = NSLOT * SLOT;
  enum { DPA = NF, ON = NF + 1, AU = NF + 2, AV = NF + 3, SC = NF + 4, MSK = NF + 5 };
};
constexpr int kDiffUse = 30;   // columns a strip yields; its first staged column is 30*s - 2 (even)

// request the row at element offset `off` into the slot at byte offset `soff` of the ring (one lane)
template <int NF>
__device__ __forceinline__ void diff_issue_row(const double* const (&src)[NF + 6], long off, uint32_t ring_s,
                                               uint32_t soff, uint32_t bar) {
  typedef DRing<NF> R;
  const uint32_t dst = ring_s + soff;
  mbar_expect_tx(bar, R::SLOT);
#pragma unroll
  for (int a = 0; a < R::NARR; ++a) bulk_g2s(dst + a * R::RB, src[a] + off, R::RB, bar);
}

template <int NF, bool EOS>
__global__ void __launch_bounds__(128, 4) k_tsdff_march(const DiffMarchParams P) {
  typedef DRing<NF> R;
  extern __shared__ unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const long unit = (long)blockIdx.x * 4 + wid;
  if (unit >= P.nunits) return;
  // layer fastest: the warps that run together share the 2-D operands through L2
  const int k0 = (int)(unit % P.kk);
  const long t0 = unit / P.kk;
  const int strip = (int)(t0 % P.nstrips), chunk = (int)(t0 / P.nstrips);
  const int w0 = strip * kDiffUse - 2;
  const int j0 = chunk * P.chunk_rows, j1 = min(j0 + P.chunk_rows, P.nrows);
  const uint32_t s0 = smem_u32(smem_raw);
  const uint32_t pad = ((s0 + 127u) & ~127u) - s0;
  const unsigned char* ring = smem_raw + pad + wid * R::BYTES;
  const uint32_t ring_s = s0 + pad + wid * R::BYTES;
  const uint32_t bar_s = s0 + pad + 4 * R::BYTES + wid * 64;
  const long ko = (long)k0 * P.slab + w0;
  const double* src[R::NARR];
#pragma unroll
  for (int f = 0; f < NF; ++f) src[f] = P.in[f] + ko;
  src[R::DPA] = P.dp + ko;
  src[R::ON] = P.oneta + w0; src[R::AU] = P.au + w0; src[R::AV] = P.av + w0;
  src[R::SC] = P.scp2 + w0; src[R::MSK] = P.mask64 + w0;
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < R::NSLOT; ++q) mbar_init(bar_s + 8u * q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // step t: row r = r0+t is the NORTH row, r-1 the centre (stored when inside the chunk), r-2 the south
  // row; t = 0,1 only build the carried values for the first stored row
  const int r0 = j0 - 1;
  const int niter = (j1 - j0) + 2;
  const int nstore = j1 - j0;
  {   // rows r0-2, r0-1 ("below the chunk", slots 4 and 5): zeros, all land
    double* z = reinterpret_cast<double*>(const_cast<unsigned char*>(ring) + 4 * R::SLOT);
    for (int i = lane; i < 2 * R::SLOT / 8; i += 32) z[i] = 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }
  // the source row of the next request: one warp-uniform offset that walks down the slabs (rows outside
  // the slab repeat the nearest one)
  long goff = (long)max(0, min(r0, P.nrows - 1)) * P.pitch;
  int grow = r0;
#define TSDFF_NEXT_ROW goff += ((unsigned)grow < (unsigned)(P.nrows - 1)) ? (long)P.pitch : 0L; grow += 1;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    if (elect_one()) diff_issue_row<NF>(src, goff, ring_s, q * R::SLOT, bar_s + 8u * q);
    TSDFF_NEXT_ROW
  }
  const unsigned char* pc = ring + 8 * lane;
  const unsigned char* pw = ring + 8 * max(lane - 1, 0);
  const unsigned char* pe = ring + 8 * min(lane + 1, 31);
  auto at = [](const unsigned char* q, uint32_t so, int arr) {
    return *reinterpret_cast<const double*>(q + so + arr * R::RB);
  };
  auto mask_at = [](const unsigned char* q, uint32_t so) {
    return *reinterpret_cast<const unsigned*>(q + so + R::MSK * R::RB);
  };
  // this lane's output cell walks down the slab one row per iteration (the row finished in iteration t is
  // r0+t-1); lanes 1..30 of the window inside the slab store
  const int col = w0 + lane;
  const bool mine = lane >= 1 && lane <= kDiffUse && (unsigned)col < (unsigned)P.pitch;
  long oq = (long)k0 * P.slab + (long)(r0 - 1) * P.pitch + col;
  // carried from the previous row: dp*oneta and the mask of the centre row, its south face factor
  double hc = 0.0, gs = 0.0;
  unsigned mc = 0u;
  // the ring slots of the north, centre and south row rotate as three warp-uniform byte offsets
  uint32_t sn = 0, sc_ = 5 * R::SLOT, ss = 4 * R::SLOT;
  uint32_t bn = bar_s, par = 0;
  const int k = k0 + 1;
  const bool ldtemp = k <= P.nhybrd && P.temdfc > 0.0;                               // :2170
  const bool ldth3d = (k <= P.nhybrd && P.temdfc < 1.0) || (k == 1 && P.isopyc);     // :2171-2172
#pragma unroll 1
  for (int t = 0; t < niter; ++t) {
    mbar_wait(bn, par);
    const unsigned mn = mask_at(pc, sn);
    const double hn = at(pc, sn, R::DPA) * at(pc, sn, R::ON);
    // east and north face factors of the centre cell (west = east of the lane to the left, south =
    // north of the previous row)
    const double he = shdn(hc);
    const unsigned me = __shfl_down_sync(TSADVC_FULLMASK, mc, 1);
    const bool fe = me & M_IU, fn = mn & M_IV, fw = mc & M_IU, fs = mc & M_IV;
    const double ge = fe ? at(pe, sc_, R::AU) * harmonc(hc, he) : 0.0;
    const double gn = fn ? at(pc, sn, R::AV) * harmonc(hc, hn) : 0.0;
    const double gw = shup(ge);
    const double factor = div_rn(-P.delt1, at(pc, sc_, R::SC) * dmax(hc, 1.0e-20));   // :2314-2315
    double v[NF], old[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const double x = at(pc, sc_, f);
      const double uw = fw ? gw * (at(pw, sc_, f) - x) : 0.0;
      const double ue = fe ? ge * (x - at(pe, sc_, f)) : 0.0;
      const double vs = fs ? gs * (at(pc, ss, f) - x) : 0.0;
      const double vn = fn ? gn * (x - at(pc, sn, f)) : 0.0;
      const double util = ((ue - uw) + (vn - vs)) * factor;
      const bool on_ = !EOS || f == 1 || (f == 0 ? ldtemp : ldth3d);   // :2173-2185
      old[f] = x;
      v[f] = on_ ? x + util : x;
    }
    if (mine && (unsigned)(t - 2) < (unsigned)nstore) {
      const bool wr = mc & M_OUT;
      if (EOS) {   // :2199-2229
        double tt = v[0], s = v[1], h = v[NF - 1];
        if (wr) {
          if (ldtemp && ldth3d) {
            const double th3d_t = eos::sig(P.eosc, tt, s) - P.thbase;
            h = (1.0 - P.temdfc) * h + P.temdfc * th3d_t;
            tt = eos::tofsig(P.eosc, h + P.thbase, s);
          } else if (ldtemp) {
            h = eos::sig(P.eosc, tt, s) - P.thbase;
          } else if (ldth3d) {
            tt = eos::tofsig(P.eosc, h + P.thbase, s);
          } else {
            h = P.theta[oq];
            tt = eos::tofsig(P.eosc, h + P.thbase, s);
          }
        }
        P.out[0][oq] = wr ? tt : old[0];
        P.out[1][oq] = wr ? s : old[1];
        P.out[NF - 1][oq] = wr ? h : old[NF - 1];
      } else {
#pragma unroll
        for (int f = 0; f < NF; ++f) P.out[f][oq] = wr ? v[f] : old[f];
      }
    }
    hc = hn; gs = gn; mc = mn;
    __syncwarp();
    // row r+3 goes into the slot of row r-3 (the south row of the previous iteration); the slots rotate
    const uint32_t fill = sn >= 3 * R::SLOT ? sn - 3 * R::SLOT : sn + 3 * R::SLOT;
    const uint32_t bfill = sn >= 3 * R::SLOT ? bn - 24u : bn + 24u;
    if (t + 3 < niter && elect_one()) diff_issue_row<NF>(src, goff, ring_s, fill, bfill);
    TSDFF_NEXT_ROW
    ss = sc_; sc_ = sn;
    sn = (sn == 5 * R::SLOT) ? 0u : sn + R::SLOT;
    par ^= (sn == 0) ? 1u : 0u;               // slot 5 closes a round of six
    bn = (sn == 0) ? bar_s : bn + 8u;
    oq += P.pitch;
  }
#undef TSDFF_NEXT_ROW
}

template <int NF, bool EOS>
static int launch_diff_march_variant(const DiffMarchParams& P, cudaStream_t stream) {
  static bool attr_set = false;
  const int bytes = 4 * (DRing<NF>::BYTES + 64) + 128;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_tsdff_march<NF, EOS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long nblocks = (P.nunits + 3) / 4;
  k_tsdff_march<NF, EOS><<<(unsigned)nblocks, 128, bytes, stream>>>(P);
  return (int)cudaGetLastError();
}

// au = temdf2*aspux*scuy, av = temdf2*aspvy*scvx (left to right, as the face factors are written)
__global__ void k_diff_static(const double* __restrict__ aspux, const double* __restrict__ scuy,
                              const double* __restrict__ aspvy, const double* __restrict__ scvx, double temdf2,
                              double* __restrict__ au, double* __restrict__ av, long n) {
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x) {
    au[q] = temdf2 * aspux[q] * scuy[q];
    av[q] = temdf2 * aspvy[q] * scvx[q];
  }
}

}  // namespace

int launch_isopyc_smooth(const double* u, const double* v, double* us, double* vs, const uint8_t* mask,
                         int pitch, int nrows, int nbdy, int ii, int jj, int margin, cudaStream_t stream) {
  const dim3 block(32, 8), grid((pitch + 31) / 32, (nrows + 7) / 8);
  k_isopyc_smooth<<<grid, block, 0, stream>>>(u, v, us, vs, mask, pitch, nrows, nbdy, ii, jj, margin);
  return (int)cudaGetLastError();
}

int launch_diff_static(const double* aspux, const double* scuy, const double* aspvy, const double* scvx,
                       double temdf2, double* au, double* av, long n, cudaStream_t stream) {
  k_diff_static<<<148 * 4, 256, 0, stream>>>(aspux, scuy, aspvy, scvx, temdf2, au, av, n);
  return (int)cudaGetLastError();
}

// nf = 3 with eos (temp, saln, th3d), or 1 / 2 plain fields
int launch_tsdff_march(const DiffMarchParams& P, cudaStream_t stream) {
  if (P.nunits <= 0) return 0;
  if (P.eos && P.nf == 3) return launch_diff_march_variant<3, true>(P, stream);
  if (!P.eos && P.nf == 2) return launch_diff_march_variant<2, false>(P, stream);
  if (!P.eos && P.nf == 1) return launch_diff_march_variant<1, false>(P, stream);
  return -1;
}

}  // namespace tsadvc
