// private: the transport of a multi-tile run, the library's own mod_xc.
//
// mod_xc (mod_xc_mp.h:2317-3288) owns MPI_COMM_HYCOM and every xctilr call moves packed edge
// strips through it (mod_xc_mp.h:4664-4987).  Here the handle owns the communicator: NCCL
// send/recv between the GPUs of one box (one process per GPU; the library is loaded with
// dlopen, a single-tile host needs no NCCL), or an in-process transport between handles
// driven by host threads (one process, several handles; what the single-GPU tests use).
#pragma once
#include <cuda_runtime.h>

#include "tsadvc_handle.h"
#include "tsadvc_launch.h"

namespace tsadvc {

// xctilr of the arrays of `a`: pack -> transport -> unpack on stream `st`; `outer`: also set the
// never-refreshed halo lines beyond mh/nh to vland (k_halo_outer_multi)
int xc_exchange(hycom_tsadvc_handle* h, const HaloArrays& a, bool outer, cudaStream_t st);
// xcminr / xcmaxr of d_mm[0:kk] / d_mm[kk:2kk] (device, in place) over all tiles
int xc_minmax(hycom_tsadvc_handle* h, double* d_mm, int kk, cudaStream_t st);
// sum over all tiles of n 64-bit words (device, in place)
int xc_sum_u64(hycom_tsadvc_handle* h, unsigned long long* d, int n, cudaStream_t st);
// stream the overlapped exchange runs on (null without a communicator)
cudaStream_t xc_stream(hycom_tsadvc_handle* h);
void xc_detach(hycom_tsadvc_handle* h);

// 0-based tile index of the neighbour in direction d, -1 at a closed edge
int neighbour(const hycom_tsadvc_dims& d, int dir);
inline int arctic_fold(const hycom_tsadvc_dims& d) {
  return (d.nreg == 2 && d.ipr * d.jpr > 1 && d.nproc == d.jpr) ? 1 : 0;
}
// the direction in which the peer sent the message that arrives from direction d: the opposite one,
// except across the fold where N, NW, NE pair with themselves (both twins look north)
inline int opp_dir(const hycom_tsadvc_dims& g, int d) {
  static const int opp[8] = {1, 0, 3, 2, 7, 6, 5, 4};
  return (arctic_fold(g) && (d == 3 || d == 6 || d == 7)) ? d : opp[d];
}

}  // namespace tsadvc
