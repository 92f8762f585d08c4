// host <-> kernel launch contract of the marching kernels
#pragma once
#include <cuda_runtime.h>

#include "tsadvc_dev.h"

namespace tsadvc {

constexpr int kWarpsPerBlock = 4;

struct MarchParams {
  FieldDesc fld[kMaxFields];
  int nfld;
  int kk;
  const double* u;   // uflx(:,:,1)
  const double* v;   // vflx(:,:,1)
  const double* dp;  // dp(:,:,1,n)
  long slab;         // doubles per 2-D slab (pitch*nrows)
  int njobs;         // nfld*kk; job = field + nfld*(k-1): T and S of a layer adjacent
  Geo g;
  int nc;           // cells per lane (1 or 2)
  int minb;         // resident blocks per SM the variant is compiled for
  int nstrips;      // strip_count(pitch, nc)
  int nchunks;      // ceil(nrows/chunk_rows)
  int chunk_rows;
  long nunits;      // njobs*nstrips*nchunks; unit = job + njobs*(strip + nstrips*chunk)
};

// scheme: 1 = MPDATA, 2 = FCT2 (advtyp of blkdat.input, mod_tsadvc.F90:87-90)
int launch_march(int scheme, const MarchParams& P, cudaStream_t stream);

// aux kernels (halo.cu)
int launch_halo_local(double* base, long slab, int nslab, int pitch, int nbdy, int ii, int jj,
                      int mh, int nh, int periodic_i, int periodic_j, cudaStream_t stream);

int launch_halo_outer(double* base, long slab, int nslab, int pitch, int nrows, int nbdy, int ii,
                      int jj, int mh, int nh, cudaStream_t stream);

}  // namespace tsadvc
