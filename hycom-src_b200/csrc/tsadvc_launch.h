// host <-> kernel launch contract of the marching kernels
#pragma once
#include <cuda_runtime.h>

#include "eos.cuh"
#include "tsadvc_dev.h"

namespace tsadvc {

constexpr int kWarpsPerBlock = 4;

struct MarchRect {
  int strip0, nstrips;   // strips strip0 .. strip0+nstrips-1
  int row0, row1;        // rows [row0, row1) are stored
  int nchunks;           // ceil((row1-row0)/chunk_rows)
  int chunk_rows;        // rows one warp marches (plus 6 rows of pipeline fill)
  long unit0;            // first unit: unit = unit0 + job + njobs*(strip + nstrips*chunk)
};

// a run of rows of one strip (instead of the regular chunks of a rectangle)
struct MarchSeg { int strip, j0, j1, pad; };

struct MarchParams {
  FieldDesc fld[kMaxFields];
  int nfld;
  int kk;
  const double* u;   // uflx(:,:,1)
  const double* v;   // vflx(:,:,1)
  const double* dp;  // dp(:,:,1,n)
  // the mass fluxes the prolog (fco, mod_tsadvc.F90:1930-1938) is built from where they are not the
  // advecting ones: isopyc, tracers and q2, q2l of layer 1 (:2016-2048).  nullptr otherwise
  const double* u2;
  const double* v2;
  long slab;         // doubles per 2-D slab (pitch*nrows)
  int njobs;         // nfld*kk; job = field + nfld*(k-1): T and S of a layer adjacent
  Geo g;
  int nc;           // cells per lane (1 or 2)
  int minb;         // resident blocks per SM the variant is compiled for
  // the launch covers up to four rectangles of (strip, row) space: one for a whole slab or
  // for the interior of a tile, four for the frame that depends on halo cells
  // (interior/frame split: the frame waits for the halo exchange, the interior overlaps it)
  int nrect;
  MarchRect rect[4];
  long nunits;      // sum over rectangles of njobs*nstrips*nchunks
  // alternatively the launch covers a list of row segments: unit = job + njobs*segment
  const MarchSeg* seg;
  long nseg;
  int allsea;       // every staged cell of every segment is sea (mask byte 0xff): mask-free body
};

// scheme: 1 = MPDATA, 2 = FCT2 (advtyp of blkdat.input, mod_tsadvc.F90:87-90)

int launch_march_tma(int scheme, const MarchParams& P, cudaStream_t stream);

// aux kernels (halo.cu)
int launch_halo_local(double* base, long slab, int nslab, int pitch, int nbdy, int ii, int jj,
                      int mh, int nh, int periodic_i, int periodic_j, cudaStream_t stream);

int launch_halo_outer(double* base, long slab, int nslab, int pitch, int nrows, int nbdy, int ii,
                      int jj, int mh, int nh, cudaStream_t stream);

}  // namespace tsadvc

// ---- multi-tile halo exchange building blocks (halo.cu) ---------------------------------
namespace tsadvc {

constexpr int kMaxHaloArrays = 2 * kMaxFields + 4;

// the 3-D arrays one tsadvc(m,n) call exchanges (mod_tsadvc.F90:1829-1836): every entry is
// `kk` consecutive slabs; the message of one direction is [array][k][row][col] contiguous
struct HaloArrays {
  double* base[kMaxHaloArrays];
  int narr;
  int kk;
  long slab;
  int pitch, nrows, nbdy, ii, jj, mh, nh;
  // tile of the top row of a global grid across the arctic (nreg=2, mod_xc_mp.h:4114-4662): the
  // directions N, NW, NE carry the tripole fold instead of a plain strip.  itype is xctilr's grid and
  // field type of each array (1 p-grid scalar, 13 u-grid vector, 14 v-grid vector, mod_xc.F90:41-44)
  int fold;
  int itype[kMaxHaloArrays];
};
struct HaloBufs {
  double* buf[8];     // per direction (W,E,S,N,SW,SE,NW,NE); nullptr: skip (pack) / vland (unpack)
  long count[8];      // doubles per direction (0: no message)
};

// (width, height) of the strip exchanged with direction d and its first (column,row) in the
// slab: `recv` selects the halo cells that are filled, else the interior cells that are sent
__host__ __device__ inline void halo_region(const HaloArrays& a, int d, bool recv, int& w, int& h,
                                            int& c0, int& r0) {
  const int nb = a.nbdy;
  // x extent
  const int xs = (d == 0 || d == 4 || d == 6) ? -1 : (d == 1 || d == 5 || d == 7) ? 1 : 0;
  const int ys = (d == 2 || d == 4 || d == 5) ? -1 : (d == 3 || d == 6 || d == 7) ? 1 : 0;
  if (xs == 0) { w = a.ii; c0 = nb; }
  else if (xs < 0) { w = a.mh; c0 = recv ? nb - a.mh : nb; }
  else { w = a.mh; c0 = recv ? nb + a.ii : nb + a.ii - a.mh; }
  if (ys == 0) { h = a.jj; r0 = nb; }
  else if (ys < 0) { h = a.nh; r0 = recv ? nb - a.nh : nb; }
  else { h = a.nh; r0 = recv ? nb + a.jj : nb + a.jj - a.nh; }
}

// Fold messages (directions N=3, NW=6, NE=7 of a top-row arctic tile; all top-row tiles have the
// same ii).  The tile T' = twin of the receiver sends all ii columns (N), the tile east of T' its first
// mh+1 columns (NW) and the tile west of T' its last mh columns (NE); per array the rows are
// jj-1-j (p,u grid) or jj-j (q,v grid), j=1..nh, and the sender applies the sign rule of
// mod_xc_mp.h:4263-4372 (vector fields change sign unless the value is vland).  The receiver
// mirrors: column c of T' lands at i = ii+1+s-c (s = 1 on the u,q grids whose mirror is shifted by
// one column, mod_xc_mp.h:4283-4331) in row jj+j; cells that fall outside 1-mh..ii+mh are dropped.
__host__ __device__ inline bool halo_is_fold(const HaloArrays& a, int d) {
  return a.fold && (d == 3 || d == 6 || d == 7);
}
__host__ __device__ inline void halo_fold_region(const HaloArrays& a, int d, int& w, int& h, int& c1) {
  h = a.nh;
  if (d == 3) { w = a.ii; c1 = 1; }
  else if (d == 6) { w = a.mh + 1; c1 = 1; }
  else { w = a.mh; c1 = a.ii - a.mh + 1; }
}
// doubles of the message in direction d
__host__ __device__ inline long halo_count(const HaloArrays& a, int d) {
  int w, h, c0, r0;
  if (halo_is_fold(a, d)) halo_fold_region(a, d, w, h, c0);
  else halo_region(a, d, false, w, h, c0, r0);
  return (long)w * h * a.narr * a.kk;
}

int launch_halo_pack(const HaloArrays& a, const HaloBufs& b, cudaStream_t stream);
int launch_halo_unpack(const HaloArrays& a, const HaloBufs& b, cudaStream_t stream);
int launch_halo_outer_multi(const HaloArrays& a, cudaStream_t stream);

}  // namespace tsadvc

// ---- diffusion + equation of state after advection (tsdff.cu) ---------------------------
namespace tsadvc {

// one launch group of k_tsdff_march: the T/S/th3d triple with the equation-of-state epilogue (eos = 1:
// in[0] temp, in[1] saln, in[2] th3d; mod_tsadvc.F90:2166-2185 + :2199-2229) or one / two plain fields
// (tracer pairs, q2 & q2l; :2180-2198), all kk layers
struct DiffMarchParams {
  const double* in[3];
  double* out[3];
  int nf, eos;
  const double *dp, *oneta, *au, *av, *scp2, *mask64, *theta;
  long slab;
  int pitch, nrows, kk;
  int nstrips, chunk_rows, nchunks;
  long nunits;          // kk * nstrips * nchunks warps
  int nhybrd, isopyc;
  eos::Coef eosc;
  double temdfc, thbase, delt1;
};
int launch_tsdff_march(const DiffMarchParams& P, cudaStream_t stream);
int launch_diff_static(const double* aspux, const double* scuy, const double* aspvy, const double* scvx,
                       double temdf2, double* au, double* av, long n, cudaStream_t stream);
// uflux, vflux of mod_tsadvc.F90:1859-1897 from uflx(:,:,1), vflx(:,:,1)
int launch_isopyc_smooth(const double* u, const double* v, double* us, double* vs, const uint8_t* mask,
                         int pitch, int nrows, int nbdy, int ii, int jj, int margin, cudaStream_t stream);

}  // namespace tsadvc

// ---- advem_fct2c (btrmas), fct2c.cu ------------------------------------------------------
namespace tsadvc {

// one batch of nb layers starting at layer k0 (0-based), every advected field
struct Fct2cParams {
  const double* fld[kMaxFields];    // (:,:,1,n)  time level t-1
  const double* fldc[kMaxFields];   // (:,:,1,m)  time level t
  double* out[kMaxFields];          // ping-pong buffer of (:,:,1,n)
  int nlay[kMaxFields];             // layers 1..nlay of the field are advected
  int nf, k0, nb;
  int pitch, ncols, nrows, nbdy, ii, jj;
  long slab;
  const uint8_t* mask;
  const double *scp2, *scp2i, *oneta;   // oneta(:,:,n): onetamas(:,:,m) when btrmas (:1806)
  const double *u, *v, *dp;             // uflx(:,:,1), vflx(:,:,1), dp(:,:,1,n)
  double dt2;
  // scratch, per layer of the batch: [nb] slabs
  double *hloc, *dtloc, *ucum, *vcum, *uloc, *vloc;
  uint8_t* lcalc;
  // scratch, per field and layer: [nf][nb] slabs
  double *fldlo, *flx, *fly, *flxcum, *flycum;
};

// stage 0 init, 1 dtloc, 2 faces, 3 cells, 4 fax/fay, 5 rp/rm, 6 final update
int launch_fct2c(int stage, const Fct2cParams& P, cudaStream_t stream);

}  // namespace tsadvc

// ---- Robert-Asselin filter of the scalar fields (asselin.cu; mod_asselin.F90) ---------------
namespace tsadvc {

struct AsselinField { const double* o; double* fm; const double* fn; };   // t-1 (saved), t, t+1
struct AsselinCopy { double* o; const double* fn; };

struct AsselinParams {
  union { AsselinField f[kMaxFields + 1]; AsselinCopy cp[kMaxFields + 1]; };   // filter: saln, temp, th3d, tracers
  int nf, kcopy;                     // kcopy: layers of the save copies (kk, or kk+2 for oq2/oq2l)
  int pitch, nrows, nbdy, ii, jj, kk;
  long slab;
  const uint8_t* mask;
  const double *pbavg_n, *pbavg_m, *pbot;
  double *oneta_n, *oneta_m, *onetao_n, *onetao_m;
  const double *dpo_n, *dpo_m, *dp_n;
  double* dp_m;
  const double* theta;
  const double *q2_o, *q2l_o, *q2_n, *q2l_n;   // q2_o null unless mxlmy; all at slab 0 of (0:kk+1)
  double *q2_m, *q2l_m;
  int nhybrd, advflg, isopyc;
  eos::Coef eosc;
  double ra2fac, oneta0, thbase;
};
int launch_asselin(int stage, const AsselinParams& P, cudaStream_t stream);

}  // namespace tsadvc

// ---- cnuity(m,n) on the device mirrors (cnuity.cu; cnuity.F90) ---------------------------------
namespace tsadvc {

struct CnuityParams {
  int pitch, nrows, nbdy, ii, jj, kk;
  long slab;
  const uint8_t* mask;
  const double *scuy, *scvx, *scp2i, *depthu, *depthv, *pbot;
  double *dp_n, *dp_m, *dpo_n, *dpo_m;          // (:,:,1,n|m)
  const double *u_m, *v_m, *dpu_m, *dpv_m;      // (:,:,1,m)
  const double *ubavg_m, *vbavg_m;              // (:,:,m)
  double *dpmixl_n, *dpmold;
  double *uflx, *vflx, *p, *utotn, *vtotn;
  double *uflxav, *vflxav, *dpav;               // may be null: not accumulated
  double* dnew;                                 // scratch: dp after loop 76 (kk slabs)
  // interface-depth diffusion (:745-1124): coefficients at the u / v points, scp2, and five scratch slabs
  const double *thku, *thkv, *scp2;
  double *pold, *t1, *t2, *fu, *fv;
  int defer_av;                                 // the cumulative fluxes (:1326-1350) follow the diffusion
  double* dpkmin;                               // 2*kk (device)
  double delt1, ra2fac;
  int isopyc;
};
int launch_cnuity(int stage, const CnuityParams& P, cudaStream_t stream);
// the interface-depth diffusion behind its three exchanges; bih: thkdf4, else thkdf2; returns launches (< 0: error)
int launch_cnuity_thkdf(const CnuityParams& P, int bih, int nstep, cudaStream_t stream);
// hybrid .and. mxlkta (:1144-1324); mode 0: dpmixl is only advected, 1: + biharmonic, 2: + Laplacian diffusion
int launch_cnuity_mxlkta(const CnuityParams& P, int mode, cudaStream_t stream);

}  // namespace tsadvc
