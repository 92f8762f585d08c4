// Device-side data model of the tsadvc hot path (B200 / sm_100a).
//
// Layout in HBM (all fp64 unless noted):
//   a 2-D "slab" is nrows x pitch doubles, row-major with i (the Fortran first
//   index) contiguous; column c <-> Fortran i = c + 1 - nbdy, row r <-> Fortran
//   j = r + 1 - nbdy, exactly the reference's padded (1-nbdy:idm+nbdy,
//   1-nbdy:jdm+nbdy) array (mod_dimensions.F90:62-84) except that `pitch` is
//   rounded up to an even number of doubles so every row starts 16-byte aligned.
//   3-D/4-D fields are stacks of slabs: slab index = (k-1) + kdm*(t-1).
//   Masks are packed to one byte per cell (see MaskBits) at set_static time.
#pragma once
#include <cstdint>

namespace tsadvc {

enum MaskBits : unsigned {
  M_IP = 1u,    // ip(i,j)  != 0                    (bigrid.F90:205-214)
  M_IU = 2u,    // iu(i,j)  != 0                    (bigrid.F90:224-226)
  M_IV = 4u,    // iv(i,j)  != 0                    (bigrid.F90:227-229)
  M_PW = 8u,    // ip(i-1,j) != 0  <=> ipim1 = i-1  (bigrid.F90:322-326)
  M_PE = 16u,   // ip(i+1,j) != 0  <=> ipip1 = i+1
  M_PS = 32u,   // ip(i,j-1) != 0  <=> ipjm1 = j-1
  M_PN = 64u,   // ip(i,j+1) != 0  <=> ipjp1 = j+1
  M_OUT = 128u  // 1<=i<=ii, 1<=j<=jj and ip: the cell tsadvc writes
};

// one advected field: every layer k < nlay is one call of advem
// (mod_tsadvc.F90:1969-2034)
struct FieldDesc {
  const double* fld;   // (:,:,1,n) on entry: time level t-1
  const double* fldc;  // (:,:,1,m): time level t (unused by MPDATA/PCM)
  double* out;         // (:,:,1,n) of the ping-pong buffer: time level t+1
  double posdef;       // MPDATA offset (mod_tsadvc.F90:1762)
  int nlay;            // layers 1..nlay are advected (temp: nhybrd, :1855)
  int pad;
};
constexpr int kMaxFields = 2 + 16;

// a (field, layer) pair resolved to slabs
struct Job {
  const double* fld;
  const double* fldc;
  double* out;
  const double* u;     // uflx(:,:,k)
  const double* v;     // vflx(:,:,k)
  const double* dp;    // dp(:,:,k,n)
  double posdef;
};

struct Geo {
  int pitch;   // doubles per row (even)
  int ncols;   // idm + 2*nbdy  (columns that exist in the Fortran array)
  int nrows;   // jdm + 2*nbdy
  int nbdy;
  int ii, jj;  // tile extent
  const uint8_t* mask;  // nrows x pitch bytes, MaskBits
  const double* mask64; // the same byte in the low bits of a 64-bit word (TMA-staged path)
  const double* scp2;   // scal
  const double* scp2i;  // scali
  double delt1;         // dt2
  double onemm;
};

// strips/chunks of the marching decomposition: a warp stages 32*nc columns (nc adjacent
// cells per lane) and yields all but the apron of 3 on each side, the true dependency
// radius of FCT2/MPDATA (the reference's margins 4,3,3,2,1,0 are wider than needed)
constexpr int kApron = 3;
__host__ __device__ constexpr int strip_use(int nc) { return 32 * nc - 2 * kApron; }
// first column of strip s is s*strip_use - strip_lead: even, so that vector loads and the
// bulk copies of the TMA path are 16-byte aligned; the strip yields columns
// s*strip_use - 1 .. (s+1)*strip_use - 2
__host__ __device__ constexpr int strip_lead(int) { return 4; }
__host__ __device__ constexpr int strip_count(int pitch, int nc) {
  return (pitch + 1 + strip_use(nc) - 1) / strip_use(nc);
}

}  // namespace tsadvc
