// The transport of a multi-tile run inside the library (see xc_comm.h): what mod_xc does for the
// reference (mod_xc_mp.h:2317-3288 set-up, :4664-4987 xctilr, :6091-6389 xcminr/xcmaxr).
//
//   NCCL      one process per GPU; ncclSend/ncclRecv of the packed strips, all eight
//             neighbours in one group, on the stream the pack kernel ran on.  libnccl is
//             dlopen'ed (no link-time dependency: a single-tile host needs no NCCL); the host
//             program only has to broadcast the 128-byte unique id (MPI_Bcast in a Fortran/MPI
//             HYCOM, torch.distributed in this repository's bench).
//   in-process  several handles of one process, each driven by its own host thread (ctypes
//             releases the GIL); the strips move with cudaMemcpyAsync between the handles'
//             staging buffers, ordered by events.  This is what `pytest -m gpu` uses on its
//             single GPU: the same step code as under NCCL, only the byte moving differs.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "xc_comm.h"

using namespace tsadvc;

#define CU(h, call) TSADVC_CU(h, call)

// ---- NCCL through dlopen ---------------------------------------------------------------------
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclUint64 = 5, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2, kNcclMin = 3 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_load(hycom_tsadvc_handle* h) {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return 0;
  const char* env = getenv("HYCOM_TSADVC_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib)
    return fail(h, HYCOM_TSADVC_EUNSUPPORTED, "libnccl.so.2 not found (%s): set HYCOM_TSADVC_NCCL_LIB", dlerror());
  NcclApi a;
  a.lib = lib;
#define SYM(field, name)                                                               \
  *(void**)(&a.field) = dlsym(lib, name);                                              \
  if (!a.field) return fail(h, HYCOM_TSADVC_EUNSUPPORTED, "libnccl: symbol %s missing", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(GetErrorString, "ncclGetErrorString")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GetVersion, "ncclGetVersion")
#undef SYM
  g_nccl = a;
  return 0;
}

#define NC(h, call)                                                                        \
  do {                                                                                     \
    int r_ = (call);                                                                       \
    if (r_ != 0)                                                                           \
      return fail(h, HYCOM_TSADVC_ECUDA, "%s failed: %s (%s:%d)", #call,                   \
                  g_nccl.GetErrorString(r_), __FILE__, __LINE__);                          \
  } while (0)

}  // namespace

// ---- in-process group -------------------------------------------------------------------------
struct hycom_tsadvc_local_group {
  int n = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long gen = 0;
  bool broken = false;
  struct Slot {
    double* send[8] = {};
    cudaEvent_t packed = nullptr, copied = nullptr;
    bool have_copied = false;
    const void* host = nullptr;   // reductions
    bool attached = false;
  };
  std::vector<Slot> slot;
  // every rank waits until all have arrived; false if a rank gave up (error elsewhere) or after 120 s
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (broken) return false;
    const long g = gen;
    if (++arrived == n) {
      arrived = 0;
      ++gen;
      cv.notify_all();
      return true;
    }
    const bool ok = cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g || broken; });
    if (!ok || broken) { broken = true; cv.notify_all(); return false; }
    return true;
  }
};

namespace tsadvc {

struct XcComm {
  int kind = 0;   // 1 NCCL, 2 in-process
  int nranks = 0, rank = 0;
  int nbr[8];
  ncclComm_t nccl = nullptr;
  hycom_tsadvc_local_group* grp = nullptr;
  cudaStream_t stream = nullptr;      // the overlapped exchange runs here
  double* send[8] = {};
  double* recv[8] = {};
  size_t cap_send[8] = {}, cap_recv[8] = {};
  std::vector<double> host;           // in-process reductions
};

// 0-based tile index of the neighbour in direction d, -1 at a closed edge
// (mod_xc.F90:25-31: nreg 1,3 periodic in i; nreg 3,4 periodic in j).  Across the arctic (nreg=2)
// the tiles of the top row face their twins: idproc(m,jpr+1) = idproc(ipr+1-m,jpr)
// (mod_xc_mp.h:2830), so N is the twin of this tile, NW the twin of the western and NE the twin of
// the eastern neighbour.
int neighbour(const hycom_tsadvc_dims& d, int dir) {
  static const int dx[8] = {-1, 1, 0, 0, -1, 1, -1, 1};
  static const int dy[8] = {0, 0, -1, 1, -1, -1, 1, 1};
  const bool per_i = !(d.nreg == 0 || d.nreg == 4), per_j = d.nreg > 2;
  if (arctic_fold(d) && dy[dir] > 0) {
    const int mw = ((d.mproc - 1 + dx[dir]) % d.ipr + d.ipr) % d.ipr;
    return (d.ipr - 1 - mw) + d.ipr * (d.nproc - 1);
  }
  int mp = d.mproc - 1 + dx[dir], np = d.nproc - 1 + dy[dir];
  if (mp < 0 || mp >= d.ipr) { if (!per_i) return -1; mp = (mp + d.ipr) % d.ipr; }
  if (np < 0 || np >= d.jpr) { if (!per_j) return -1; np = (np + d.jpr) % d.jpr; }
  return mp + d.ipr * np;
}

cudaStream_t xc_stream(hycom_tsadvc_handle* h) { return h && h->xc ? h->xc->stream : nullptr; }

void xc_detach(hycom_tsadvc_handle* h) {
  if (!h || !h->xc) return;
  XcComm* x = h->xc;
  cudaSetDevice(h->d.device);
  cudaDeviceSynchronize();
  if (x->kind == 1 && x->nccl) g_nccl.CommDestroy(x->nccl);
  if (x->kind == 2 && x->grp) {
    std::lock_guard<std::mutex> lk(x->grp->mu);
    auto& s = x->grp->slot[x->rank];
    if (s.packed) cudaEventDestroy(s.packed);
    if (s.copied) cudaEventDestroy(s.copied);
    s = hycom_tsadvc_local_group::Slot();
  }
  for (int d = 0; d < 8; ++d) { cudaFree(x->send[d]); cudaFree(x->recv[d]); }
  if (x->stream) cudaStreamDestroy(x->stream);
  delete x;
  h->xc = nullptr;
}

static int xc_new(hycom_tsadvc_handle* h, int kind, XcComm** out) {
  if (!h) return fail(nullptr, HYCOM_TSADVC_EINVAL, "comm: null handle");
  if (h->xc) return fail(h, HYCOM_TSADVC_EINVAL, "comm: the handle already has a communicator");
  XcComm* x = new (std::nothrow) XcComm();
  if (!x) return fail(h, HYCOM_TSADVC_ENOMEM, "out of host memory");
  x->kind = kind;
  x->nranks = h->d.ipr * h->d.jpr;
  x->rank = h->d.mproc - 1 + h->d.ipr * (h->d.nproc - 1);
  for (int d = 0; d < 8; ++d) x->nbr[d] = neighbour(h->d, d);
  cudaError_t e = cudaSetDevice(h->d.device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete x;
    return fail(h, HYCOM_TSADVC_ECUDA, "comm: stream creation failed: %s", cudaGetErrorString(e));
  }
  *out = x;
  return 0;
}

// staging buffers, grown on demand (the first steps of a run)
static int xc_reserve(hycom_tsadvc_handle* h, XcComm* x, const long cnt[8], cudaStream_t st) {
  bool grow = false;
  for (int d = 0; d < 8; ++d) {
    if (x->nbr[d] < 0) continue;
    const bool self = x->nbr[d] == x->rank;
    if ((size_t)cnt[d] > x->cap_send[d] || (!self && (size_t)cnt[d] > x->cap_recv[d])) grow = true;
  }
  if (!grow) return 0;
  CU(h, cudaStreamSynchronize(st));
  CU(h, cudaStreamSynchronize(x->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  if (x->kind == 2) {   // peers may still be copying out of the old buffers (previous exchange)
    hycom_tsadvc_local_group* g = x->grp;
    for (int r = 0; r < g->n; ++r)
      if (r != x->rank && g->slot[r].have_copied) CU(h, cudaEventSynchronize(g->slot[r].copied));
  }
  for (int d = 0; d < 8; ++d) {
    if (x->nbr[d] < 0) continue;
    const bool self = x->nbr[d] == x->rank;
    const size_t need = (size_t)cnt[d] + (size_t)cnt[d] / 8;
    if ((size_t)cnt[d] > x->cap_send[d]) {
      if (x->send[d]) { cudaFree(x->send[d]); h->bytes -= (int64_t)(8 * x->cap_send[d]); }
      x->send[d] = nullptr; x->cap_send[d] = 0;
      cudaError_t e = cudaMalloc(&x->send[d], 8 * need);
      if (e != cudaSuccess) return fail(h, HYCOM_TSADVC_ENOMEM, "cudaMalloc(%zu) failed: %s", 8 * need, cudaGetErrorString(e));
      x->cap_send[d] = need; h->bytes += (int64_t)(8 * need);
    }
    if (!self && (size_t)cnt[d] > x->cap_recv[d]) {
      if (x->recv[d]) { cudaFree(x->recv[d]); h->bytes -= (int64_t)(8 * x->cap_recv[d]); }
      x->recv[d] = nullptr; x->cap_recv[d] = 0;
      cudaError_t e = cudaMalloc(&x->recv[d], 8 * need);
      if (e != cudaSuccess) return fail(h, HYCOM_TSADVC_ENOMEM, "cudaMalloc(%zu) failed: %s", 8 * need, cudaGetErrorString(e));
      x->cap_recv[d] = need; h->bytes += (int64_t)(8 * need);
    }
  }
  if (x->kind == 2) {
    auto& s = x->grp->slot[x->rank];
    std::lock_guard<std::mutex> lk(x->grp->mu);
    for (int d = 0; d < 8; ++d) s.send[d] = x->send[d];
  }
  return 0;
}

int xc_exchange(hycom_tsadvc_handle* h, const HaloArrays& a, bool outer, cudaStream_t st) {
  XcComm* x = h->xc;
  if (!x) return fail(h, HYCOM_TSADVC_EUNSUPPORTED, "no communicator attached to this multi-tile handle");
  long cnt[8];
  for (int d = 0; d < 8; ++d) cnt[d] = x->nbr[d] >= 0 ? halo_count(a, d) : 0;
  int rc;
  if ((rc = xc_reserve(h, x, cnt, st))) return rc;
  HaloBufs sb, rb;
  for (int d = 0; d < 8; ++d) {
    sb.buf[d] = x->nbr[d] >= 0 ? x->send[d] : nullptr; sb.count[d] = cnt[d];
    const int peer = x->nbr[d];
    // a periodic edge that wraps onto this tile: what leaves in the opposite direction arrives here
    rb.buf[d] = peer < 0 ? nullptr : (peer == x->rank ? x->send[opp_dir(h->d, d)] : x->recv[d]);
    rb.count[d] = cnt[d];
  }
  if (x->kind == 2) {   // peers must be done reading the staging buffers of the previous exchange
    hycom_tsadvc_local_group* g = x->grp;
    std::lock_guard<std::mutex> lk(g->mu);
    for (int r = 0; r < g->n; ++r)
      if (r != x->rank && g->slot[r].have_copied) CU(h, cudaStreamWaitEvent(st, g->slot[r].copied, 0));
  }
  rc = launch_halo_pack(a, sb, st);
  h->launches += 1;
  if (rc) return fail(h, HYCOM_TSADVC_ECUDA, "halo pack launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  if (x->kind == 1) {
    NC(h, g_nccl.GroupStart());
    for (int d = 0; d < 8; ++d) {
      const int peer = x->nbr[d];
      if (peer >= 0 && peer != x->rank && cnt[d] > 0)
        NC(h, g_nccl.Send(x->send[d], (size_t)cnt[d], kNcclFloat64, peer, x->nccl, st));
    }
    // receives in the order of the SENDER's direction: NCCL matches the k-th send of A to B with the
    // k-th receive B posts for A (a top-row arctic tile may hear from the same tile twice)
    int order[8];
    for (int e = 0; e < 8; ++e) order[e] = e;
    std::stable_sort(order, order + 8, [&](int e1, int e2) { return opp_dir(h->d, e1) < opp_dir(h->d, e2); });
    for (int q = 0; q < 8; ++q) {
      const int e = order[q], peer = x->nbr[e];
      if (peer >= 0 && peer != x->rank && cnt[e] > 0)
        NC(h, g_nccl.Recv(x->recv[e], (size_t)cnt[e], kNcclFloat64, peer, x->nccl, st));
    }
    NC(h, g_nccl.GroupEnd());
  } else {
    hycom_tsadvc_local_group* g = x->grp;
    auto& me = g->slot[x->rank];
    CU(h, cudaEventRecord(me.packed, st));
    if (!g->barrier()) return fail(h, HYCOM_TSADVC_ECUDA, "in-process exchange: a peer tile did not arrive");
    for (int e = 0; e < 8; ++e) {
      const int peer = x->nbr[e];
      if (peer < 0 || peer == x->rank || cnt[e] == 0) continue;
      const auto& ps = g->slot[peer];
      CU(h, cudaStreamWaitEvent(st, ps.packed, 0));
      CU(h, cudaMemcpyAsync(x->recv[e], ps.send[opp_dir(h->d, e)], 8 * (size_t)cnt[e], cudaMemcpyDefault, st));
    }
    CU(h, cudaEventRecord(me.copied, st));
    me.have_copied = true;
    if (!g->barrier()) return fail(h, HYCOM_TSADVC_ECUDA, "in-process exchange: a peer tile did not arrive");
  }
  rc = launch_halo_unpack(a, rb, st);
  h->launches += 1;
  if (!rc && outer) { rc = launch_halo_outer_multi(a, st); h->launches += 1; }
  if (rc) return fail(h, HYCOM_TSADVC_ECUDA, "halo unpack launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  return 0;
}

// host-side reduction of the in-process transport: every rank publishes its values, all combine
template <class T, class F>
static int local_reduce(hycom_tsadvc_handle* h, T* d, int n, cudaStream_t st, F op) {
  XcComm* x = h->xc;
  hycom_tsadvc_local_group* g = x->grp;
  x->host.resize((size_t)n * sizeof(T) / sizeof(double) + 1);
  T* mine = reinterpret_cast<T*>(x->host.data());
  CU(h, cudaMemcpyAsync(mine, d, sizeof(T) * n, cudaMemcpyDeviceToHost, st));
  CU(h, cudaStreamSynchronize(st));
  g->slot[x->rank].host = mine;
  if (!g->barrier()) return fail(h, HYCOM_TSADVC_ECUDA, "in-process reduction: a peer tile did not arrive");
  std::vector<T> res(mine, mine + n);
  for (int r = 0; r < g->n; ++r) {
    if (r == x->rank) continue;
    const T* o = static_cast<const T*>(g->slot[r].host);
    for (int i = 0; i < n; ++i) res[i] = op(res[i], o[i], i);
  }
  if (!g->barrier()) return fail(h, HYCOM_TSADVC_ECUDA, "in-process reduction: a peer tile did not arrive");
  memcpy(mine, res.data(), sizeof(T) * n);
  CU(h, cudaMemcpyAsync(d, mine, sizeof(T) * n, cudaMemcpyHostToDevice, st));
  CU(h, cudaStreamSynchronize(st));
  return 0;
}

int xc_minmax(hycom_tsadvc_handle* h, double* d_mm, int kk, cudaStream_t st) {
  XcComm* x = h->xc;
  if (!x) return fail(h, HYCOM_TSADVC_EUNSUPPORTED, "no communicator attached to this multi-tile handle");
  if (x->kind == 1) {   // xcminr, xcmaxr (mod_tsadvc.F90:2093-2094) as one group
    NC(h, g_nccl.GroupStart());
    NC(h, g_nccl.AllReduce(d_mm, d_mm, (size_t)kk, kNcclFloat64, kNcclMin, x->nccl, st));
    NC(h, g_nccl.AllReduce(d_mm + kk, d_mm + kk, (size_t)kk, kNcclFloat64, kNcclMax, x->nccl, st));
    NC(h, g_nccl.GroupEnd());
    return 0;
  }
  return local_reduce<double>(h, d_mm, 2 * kk, st, [kk](double a, double b, int i) {
    return i < kk ? (b < a ? b : a) : (b > a ? b : a);
  });
}

int xc_sum_u64(hycom_tsadvc_handle* h, unsigned long long* d, int n, cudaStream_t st) {
  XcComm* x = h->xc;
  if (!x) return fail(h, HYCOM_TSADVC_EUNSUPPORTED, "no communicator attached to this multi-tile handle");
  if (x->kind == 1) {
    NC(h, g_nccl.AllReduce(d, d, (size_t)n, kNcclUint64, kNcclSum, x->nccl, st));
    return 0;
  }
  return local_reduce<unsigned long long>(h, d, n, st,
                                          [](unsigned long long a, unsigned long long b, int) { return a + b; });
}

}  // namespace tsadvc

extern "C" {

int hycom_tsadvc_comm_unique_id(char id[HYCOM_TSADVC_COMM_ID_BYTES]) {
  if (!id) return fail(nullptr, HYCOM_TSADVC_EINVAL, "comm_unique_id: null argument");
  int rc = nccl_load(nullptr);
  if (rc) return rc;
  ncclUniqueId u;
  NC(nullptr, g_nccl.GetUniqueId(&u));
  static_assert(sizeof u == HYCOM_TSADVC_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  memcpy(id, &u, sizeof u);
  return 0;
}

int hycom_tsadvc_comm_init(hycom_tsadvc_handle* h, const char id[HYCOM_TSADVC_COMM_ID_BYTES]) {
  if (!h || !id) return fail(h, HYCOM_TSADVC_EINVAL, "comm_init: null argument");
  int rc = nccl_load(h);
  if (rc) return rc;
  XcComm* x = nullptr;
  if ((rc = xc_new(h, 1, &x))) return rc;
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  int r = g_nccl.CommInitRank(&x->nccl, x->nranks, u, x->rank);
  if (r != 0) {
    cudaStreamDestroy(x->stream);
    delete x;
    return fail(h, HYCOM_TSADVC_ECUDA, "ncclCommInitRank(%d of %d) failed: %s", x->rank, x->nranks,
                g_nccl.GetErrorString(r));
  }
  h->xc = x;
  return 0;
}

int hycom_tsadvc_comm_version(int32_t* version) {
  if (!version) return fail(nullptr, HYCOM_TSADVC_EINVAL, "comm_version: null argument");
  int rc = nccl_load(nullptr);
  if (rc) return rc;
  int v = 0;
  NC(nullptr, g_nccl.GetVersion(&v));
  *version = v;
  return 0;
}

int hycom_tsadvc_local_group_create(int32_t nranks, hycom_tsadvc_local_group** out) {
  if (nranks < 1 || !out) return fail(nullptr, HYCOM_TSADVC_EINVAL, "local_group_create: bad argument");
  hycom_tsadvc_local_group* g = new (std::nothrow) hycom_tsadvc_local_group();
  if (!g) return fail(nullptr, HYCOM_TSADVC_ENOMEM, "out of host memory");
  g->n = nranks;
  g->slot.resize(nranks);
  *out = g;
  return 0;
}

int hycom_tsadvc_local_group_destroy(hycom_tsadvc_local_group* g) {
  delete g;
  return 0;
}

int hycom_tsadvc_comm_attach_local(hycom_tsadvc_handle* h, hycom_tsadvc_local_group* g) {
  if (!h || !g) return fail(h, HYCOM_TSADVC_EINVAL, "comm_attach_local: null argument");
  if (g->n != h->d.ipr * h->d.jpr)
    return fail(h, HYCOM_TSADVC_EINVAL, "comm_attach_local: the group has %d ranks, the tiling %d x %d", g->n,
                h->d.ipr, h->d.jpr);
  XcComm* x = nullptr;
  int rc;
  if ((rc = xc_new(h, 2, &x))) return rc;
  x->grp = g;
  auto& s = g->slot[x->rank];
  if (s.attached) { cudaStreamDestroy(x->stream); delete x; return fail(h, HYCOM_TSADVC_EINVAL, "comm_attach_local: tile %d is already attached", x->rank); }
  CU(h, cudaEventCreateWithFlags(&s.packed, cudaEventDisableTiming));
  CU(h, cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
  s.attached = true;
  h->xc = x;
  return 0;
}

int hycom_tsadvc_comm_detach(hycom_tsadvc_handle* h) {
  if (!h) return fail(nullptr, HYCOM_TSADVC_EINVAL, "null handle");
  xc_detach(h);
  return 0;
}

int hycom_tsadvc_set_overlap(hycom_tsadvc_handle* h, int32_t enable) {
  if (!h) return fail(nullptr, HYCOM_TSADVC_EINVAL, "null handle");
  h->overlap = enable != 0;
  return 0;
}

}  // extern "C"
