// Halo update of the device mirrors: the device-side xctilr.
//
// Single tile (mod_xc_sm.h:1337-1428): closed edges receive vland (0.0),
// periodic edges wrap; nreg=2 (mod_xc_sm.h:1172-1335): tripole fold across the arctic.  North/south lines are filled first over i=1..ii, then
// east/west columns over j=1-nhl..jj+nhl, so corners come from the second pass
// applied to the lines written by the first.
#include <cuda_runtime.h>

#include "tsadvc_launch.h"

namespace tsadvc {

// a(i,j) of the Fortran array <-> base[(j-1+nb)*pitch + (i-1+nb)]
__device__ __forceinline__ long fidx(int i, int j, int nb, int pitch) {
  return (long)(j - 1 + nb) * pitch + (i - 1 + nb);
}

__global__ void k_halo_ns(double* __restrict__ base, long slab, int nslab, int pitch, int nb,
                          int ii, int jj, int nhl, int periodic_j) {
  const long per = (long)nhl * ii;
  const long total = per * nslab;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long)gridDim.x * blockDim.x) {
    const int s = (int)(t / per);
    const long q = t - (long)s * per;
    const int j = (int)(q / ii) + 1;
    const int i = (int)(q % ii) + 1;
    double* a = base + slab * s;
    double vs = 0.0, vn = 0.0;  // vland
    if (periodic_j == 1) {
      vs = a[fidx(i, jj + 1 - j, nb, pitch)];
      vn = a[fidx(i, j, nb, pitch)];
    } else if (periodic_j >= 100) {
      // global grid that includes the arctic (mod_xc_sm.h:1172-1335): the southern boundary is
      // closed, the northern rows are the tripole fold of rows below jj, mirrored in i; the
      // mirror index and row depend on the grid (itype 1 p, 2 q, 3 u, 4 v; +10: vector, sign flips)
      const int itype = periodic_j - 100, grid = itype % 10;
      const int io = (grid == 1 || grid == 4) ? ii - (i - 1) % ii : (ii - (i - 1)) % ii + 1;
      const int jo = (grid == 1 || grid == 3) ? jj - 1 - j : jj - j;
      const double v = a[fidx(io, jo, nb, pitch)];
      vn = itype > 10 ? -v : v;
    }
    a[fidx(i, 1 - j, nb, pitch)] = vs;
    a[fidx(i, jj + j, nb, pitch)] = vn;
  }
}

__global__ void k_halo_ew(double* __restrict__ base, long slab, int nslab, int pitch, int nb,
                          int ii, int jj, int mhl, int nhl, int periodic_i) {
  const int nj = jj + 2 * nhl;
  const long per = (long)nj * mhl;
  const long total = per * nslab;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long)gridDim.x * blockDim.x) {
    const int s = (int)(t / per);
    const long q = t - (long)s * per;
    const int j = (int)(q / mhl) + 1 - nhl;
    const int i = (int)(q % mhl) + 1;
    double* a = base + slab * s;
    double vw = 0.0, ve = 0.0;
    if (periodic_i) {
      vw = a[fidx(ii + 1 - i, j, nb, pitch)];
      ve = a[fidx(i, j, nb, pitch)];
    }
    a[fidx(1 - i, j, nb, pitch)] = vw;
    a[fidx(ii + i, j, nb, pitch)] = ve;
  }
}

// Cells beyond the refreshed halo (rings mh+1..nbdy, and the unused part of a ragged
// tile) are never read by the reference and may hold r_init
// NaNs.  The marching kernels recompute an apron that touches them (results discarded),
// so they are set to vland here: finite operands keep every lane on the fast division
// path.  Nothing observable depends on these cells.
__global__ void k_halo_outer(double* __restrict__ base, long slab, int nslab, int pitch, int nrows,
                             int nb, int ii, int jj, int mhl, int nhl) {
  const int c_lo = nb - mhl, c_hi = nb + ii + mhl;   // refreshed columns [c_lo, c_hi)
  const int r_lo = nb - nhl, r_hi = nb + jj + nhl;   // refreshed rows    [r_lo, r_hi)
  const int nfull = r_lo + (nrows - r_hi);           // rows zeroed over their whole width
  const int wside = c_lo + (pitch - c_hi);           // columns zeroed in the other rows
  const long per = (long)nfull * pitch + (long)(r_hi - r_lo) * wside;
  const long total = per * nslab;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long)gridDim.x * blockDim.x) {
    const long sl = t / per;
    long q = t - sl * per;
    int r, c;
    if (q < (long)nfull * pitch) {
      r = (int)(q / pitch);
      c = (int)(q - (long)r * pitch);
      if (r >= r_lo) r += r_hi - r_lo;
    } else {
      q -= (long)nfull * pitch;
      r = r_lo + (int)(q / wside);
      c = (int)(q % wside);
      if (c >= c_lo) c += c_hi - c_lo;
    }
    base[slab * sl + (long)r * pitch + c] = 0.0;
  }
}

int launch_halo_local(double* base, long slab, int nslab, int pitch, int nbdy, int ii, int jj,
                      int mh, int nh, int periodic_i, int periodic_j, cudaStream_t stream) {
  const int mhl = max(0, min(mh, nbdy)), nhl = max(0, min(nh, nbdy));
  const int threads = 256;
  if (nhl > 0) {
    const long total = (long)nhl * ii * nslab;
    const int blocks = (int)min((total + threads - 1) / threads, (long)148 * 16);
    k_halo_ns<<<blocks, threads, 0, stream>>>(base, slab, nslab, pitch, nbdy, ii, jj, nhl,
                                              periodic_j);
  }
  if (mhl > 0) {
    const long total = (long)(jj + 2 * nhl) * mhl * nslab;
    const int blocks = (int)min((total + threads - 1) / threads, (long)148 * 16);
    k_halo_ew<<<blocks, threads, 0, stream>>>(base, slab, nslab, pitch, nbdy, ii, jj, mhl, nhl,
                                              periodic_i);
  }
  return (int)cudaGetLastError();
}

int launch_halo_outer(double* base, long slab, int nslab, int pitch, int nrows, int nbdy, int ii,
                      int jj, int mh, int nh, cudaStream_t stream) {
  const int mhl = max(0, min(mh, nbdy)), nhl = max(0, min(nh, nbdy));
  if (mhl >= nbdy && nhl >= nbdy) {
    // nothing beyond the refreshed halo unless the tile is ragged (ii < idm)
  }
  k_halo_outer<<<148 * 4, 256, 0, stream>>>(base, slab, nslab, pitch, nrows, nbdy, ii, jj, mhl, nhl);
  return (int)cudaGetLastError();
}

}  // namespace tsadvc

// ---------------------------------------------------------------------------------------
// Multi-tile exchange (mod_xc_mp.h:4664-4987).  The reference packs north/south lines,
// exchanges, then packs east/west columns INCLUDING the fresh north/south halo lines so
// that corners propagate in two hops.  Here the eight neighbours are addressed directly in
// one round (a corner comes from the diagonal tile, which is what the two hops deliver);
// directions without a neighbour (closed edge) receive vland = 0.0 like mod_xc_sm.h.
// One kernel packs every array, layer and direction of a call; one kernel unpacks.
// ---------------------------------------------------------------------------------------
namespace tsadvc {

template <bool PACK>
__global__ void k_halo_xfer(HaloArrays a, HaloBufs b) {
  long total = 0;
  long first[9];
  int w[8], h[8], c0[8], r0[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    first[d] = total;
    if (halo_is_fold(a, d)) halo_fold_region(a, d, w[d], h[d], c0[d]);   // c0: first Fortran column sent
    else halo_region(a, d, !PACK, w[d], h[d], c0[d], r0[d]);
    // a fold direction is unpacked element by element of the message; without a message (never the
    // case on a periodic top row) nothing is written
    const bool on = (PACK || halo_is_fold(a, d)) ? (b.buf[d] != nullptr) : true;
    total += on ? (long)w[d] * h[d] * a.narr * a.kk : 0;
  }
  first[8] = total;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long)gridDim.x * blockDim.x) {
    int d = 0;
#pragma unroll
    for (int q = 1; q < 8; ++q) d += (t >= first[q]) ? 1 : 0;
    // segments of skipped directions are empty, so d lands on the owning direction
    const long q = t - first[d];
    const long per = (long)w[d] * h[d];
    const int s = (int)(q / per);
    const long e = q - (long)s * per;
    const int r = (int)(e / w[d]), c = (int)(e - (long)r * w[d]);
    double* slab = a.base[s / a.kk] + a.slab * (s % a.kk);
    if (halo_is_fold(a, d)) {
      const int itype = a.itype[s / a.kk], grid = itype % 10;
      const int j = r + 1, col = c0[d] + c;            // Fortran j of the halo line, column of the sender
      if (PACK) {
        const int jo = (grid == 1 || grid == 3) ? a.jj - 1 - j : a.jj - j;
        const double v = slab[fidx(col, jo, a.nbdy, a.pitch)];
        b.buf[d][q] = (itype > 10 && v != 0.0) ? -v : v;   // sarc*a unless a == vland
      } else {
        const int sh = (grid == 2 || grid == 3) ? 1 : 0;
        const int ct = d == 3 ? col : d == 6 ? col + a.ii : col - a.ii;   // column relative to the twin
        const int i = a.ii + 1 + sh - ct;
        if (i >= 1 - a.mh && i <= a.ii + a.mh) slab[fidx(i, a.jj + j, a.nbdy, a.pitch)] = b.buf[d][q];
      }
      continue;
    }
    double* cell = slab + (long)(r0[d] + r) * a.pitch + (c0[d] + c);
    if (PACK) b.buf[d][q] = *cell;
    else *cell = b.buf[d] ? b.buf[d][q] : 0.0;  // vland
  }
}

__global__ void k_halo_outer_multi(HaloArrays a) {
  const int nb = a.nbdy;
  const int c_lo = nb - a.mh, c_hi = nb + a.ii + a.mh;
  const int r_lo = nb - a.nh, r_hi = nb + a.jj + a.nh;
  const int nfull = r_lo + (a.nrows - r_hi);
  const int wside = c_lo + (a.pitch - c_hi);
  const long per = (long)nfull * a.pitch + (long)(r_hi - r_lo) * wside;
  const long total = per * a.narr * a.kk;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long)gridDim.x * blockDim.x) {
    const int s = (int)(t / per);
    long q = t - (long)s * per;
    int r, c;
    if (q < (long)nfull * a.pitch) {
      r = (int)(q / a.pitch);
      c = (int)(q - (long)r * a.pitch);
      if (r >= r_lo) r += r_hi - r_lo;
    } else {
      q -= (long)nfull * a.pitch;
      r = r_lo + (int)(q / wside);
      c = (int)(q % wside);
      if (c >= c_lo) c += c_hi - c_lo;
    }
    a.base[s / a.kk][a.slab * (s % a.kk) + (long)r * a.pitch + c] = 0.0;
  }
}

static int halo_blocks(const HaloArrays& a) {
  const long cells = ((long)2 * a.mh * (a.jj + 2 * a.nh) + (long)2 * a.nh * a.ii) * a.narr * a.kk;
  long blocks = (cells + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int launch_halo_pack(const HaloArrays& a, const HaloBufs& b, cudaStream_t stream) {
  k_halo_xfer<true><<<halo_blocks(a), 256, 0, stream>>>(a, b);
  return (int)cudaGetLastError();
}
int launch_halo_unpack(const HaloArrays& a, const HaloBufs& b, cudaStream_t stream) {
  k_halo_xfer<false><<<halo_blocks(a), 256, 0, stream>>>(a, b);
  return (int)cudaGetLastError();
}
int launch_halo_outer_multi(const HaloArrays& a, cudaStream_t stream) {
  k_halo_outer_multi<<<148 * 4, 256, 0, stream>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace tsadvc
