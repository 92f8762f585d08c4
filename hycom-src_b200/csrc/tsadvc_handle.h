// private: the handle behind hycom_tsadvc_handle* (shared by tsadvc_abi.cu and synth.cu)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/hycom_tsadvc_b200.h"

namespace tsadvc {
struct XcComm;   // communicator of a multi-tile run (xc_comm.cu)
struct Mirror {
  double* lev[2] = {nullptr, nullptr};  // time slots 1,2 (3-D fields: lev[0] only)
  double* spare = nullptr;              // ping-pong target of the next step
};
}  // namespace tsadvc

struct hycom_tsadvc_handle {
  hycom_tsadvc_dims d;
  int pitch, ncols, nrows;
  long slab;  // doubles per slab on the device
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // copy-in / copy-out streams and per-chunk events of the pipelined host-array call
  cudaStream_t up_stream = nullptr, down_stream = nullptr;
  std::vector<cudaEvent_t> ev_chunk;
  // optional stream of the PART_FRAME launch (caller-owned) and the event the main stream waits on
  cudaStream_t frame_stream = nullptr;
  cudaEvent_t ev_frame = nullptr;
  int64_t bytes = 0;
  int64_t launches = 0;
  bool have_static = false;
  uint8_t* mask = nullptr;
  std::vector<uint8_t> mask_host;   // the same bytes on the host (row segments of the FCT2 launch)
  // row segments of the marching launch per (part, cells per lane, chunk rows): device arrays of
  // MarchSeg, [0] the all-sea segments, [1] the rest (tsadvc_abi.cu, march_segments)
  struct SegLists { void* d[2] = {nullptr, nullptr}; long n[2] = {0, 0}; };
  std::map<long, SegLists> seg_cache;
  double *scp2 = nullptr, *scp2i = nullptr, *scuy = nullptr, *scvx = nullptr, *aspux = nullptr,
         *aspvy = nullptr;
  tsadvc::Mirror temp, saln, th3d, dp, uflx, vflx;
  tsadvc::Mirror oneta;  // (:,:,2): one slab per time slot
  tsadvc::Mirror theta;  // (:,:,kdm): lev[0] only
  tsadvc::Mirror q2, q2l;  // (:,:,0:kdm+1,2): kdm+2 slabs per slot (mxlmy)
  // mod_asselin.F90 operands: dpo (kdm per slot), onetao (1 per slot), pbavg (3 slabs), pbot (1),
  // otemp/osaln/oth3d/otracer (kdm), oq2/oq2l (kdm+2)
  tsadvc::Mirror dpo, onetao, pbavg, pbot, otemp, osaln, oth3d, oq2, oq2l;
  tsadvc::Mirror otracer[HYCOM_TSADVC_MXTRCR];
  // cnuity.F90 operands: u, v, dpu, dpv (kdm per slot), ubavg, vbavg (3 slabs), depthu, depthv (1), p (kdm+1),
  // dpmixl (1 per slot), uflxav, vflxav, dpav (kdm), utotn, vtotn, dpmold (1)
  tsadvc::Mirror u, v, dpu, dpv, ubavg, vbavg, depthu, depthv, p, dpmixl, uflxav, vflxav, dpav, utotn, vtotn, dpmold;
  tsadvc::Mirror thkdf4u, thkdf4v;     // coefficients of the interface-depth diffusion (1 slab)
  double* thk_scratch = nullptr;       // pold, util1, util2, uflux, vflux of cnuity.F90:745-1124 (5 slabs)
  double* cnuity_scratch = nullptr;   // kdm slabs: dp after loop 76
  double* d_dpkmin = nullptr;         // 2*kdm
  tsadvc::Mirror tracer[HYCOM_TSADVC_MXTRCR];
  // one allocation [dp(:,:,:,1) | uflx | vflx | dp(:,:,:,2)]
  double* flux_block = nullptr;
  // one allocation [scp2i | scp2 | mask byte in the low bits of a 64-bit word]
  double* static_block = nullptr;
  // every field buffer is allocated with one guard row in front and behind (the bulk copies
  // of the first/last strip start 4 columns before / end after their row); freed from here
  std::vector<void*> raw_allocs;
  // advem_fct2c scratch (btrmas): one block of (6 + 5*nf)*nb slabs + nb byte slabs
  double* fct2c_block = nullptr;
  uint8_t* fct2c_lcalc = nullptr;
  long fct2c_slabs = 0;
  int fct2c_nb = 0;
  double* diff_static = nullptr;   // au | av: temdf2*aspux*scuy, temdf2*aspvy*scvx (2 slabs), per call
  double* isopyc_flux = nullptr;   // smoothed uflux | vflux of layer 1 (isopyc), 2 slabs
  double* d_minmax = nullptr;  // 2*kdm
  uint8_t* d_sea = nullptr;    // synthetic generator: global sea mask
  // optional per-launch timing of the marching kernel (hycom_tsadvc_set_timing)
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending, ev_free;
  double march_ms = 0.0;
  int64_t march_n = 0;
  // multi-tile runs: the transport between tiles (NCCL or in-process), its stream and staging
  // buffers (xc_comm.cu); null: the host program moves the packed strips itself
  tsadvc::XcComm* xc = nullptr;
  bool overlap = true;          // exchange overlapped with the tile interior (hycom_tsadvc_set_overlap)
  // salinity range of the last diagnostic step: device -> pinned host, fetched by
  // hycom_tsadvc_saln_range (deferred mode: the step itself never waits for the device)
  bool deferred_range = false;
  double* range_host = nullptr;   // pinned, 2*kdm
  cudaEvent_t ev_range = nullptr;
  int32_t range_nstep = -1;
  unsigned long long* d_cksum = nullptr;
  cudaEvent_t ev_xc = nullptr;    // handle stream -> exchange stream
  char err[512];
};

// error reporting shared by the translation units of the library
namespace tsadvc {
int fail(hycom_tsadvc_handle* h, int code, const char* fmt, ...);
}
#define TSADVC_CU(h, call)                                                                 \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return tsadvc::fail(h, HYCOM_TSADVC_ECUDA, "%s failed: %s (%s:%d)", #call,           \
                          cudaGetErrorString(e_), __FILE__, __LINE__);                     \
  } while (0)
