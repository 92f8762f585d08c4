// probe 5: the TMA store (cp.async.bulk.tensor shared -> global) in the shapes the marching kernel uses.
//   variant 0: load box 64x1x1 at (4,5,1), store box 58x1x1 to (7,6,2) of a second tensor
//   variant 1: the same with a negative first store column (-1): clipped?
//   variant 2: store with a row coordinate far outside (-(1<<24)): dropped?
//   variant 3: load at column -4 (zero fill) and row -2
//   variant 4: tensor with one layer only (globalDim[2] = 1)
//   variant 5: descriptors in global memory, issued under elect.sync by a 32-thread warp, commit + wait_group.read
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

__global__ void k(const CUtensorMap* maps, int lc0, int lc1, int lc2, int sc0, int sc1, int sc2, double* dbg, int mode) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long barv;
  const uint32_t dst = s32(sm), bar = s32(&barv);
  const CUtensorMap* ml = maps;
  const CUtensorMap* ms = maps + 1;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(ml) : "memory");
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(ms) : "memory");
  }
  __syncwarp();
  if ((mode & 1) ? elect_one() : threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(512) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(ml), "r"(lc0), "r"(lc1), "r"(lc2), "r"(bar) : "memory");
  }
  uint32_t ok = 0;
  for (int spin = 0; !ok && spin < 2000000; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(0) : "memory");
  double* row = (double*)sm;
  dbg[threadIdx.x] = ok ? row[threadIdx.x] : -777.0;
  dbg[32 + threadIdx.x] = ok ? row[32 + threadIdx.x] : -777.0;
  // staging row behind a 128-byte pad: lane l writes its two cells at (2l-3)*8 relative to the aligned source
  unsigned char* stg = sm + 1024 + 128;
  double* mine = (double*)(stg + 8 * (2 * (int)threadIdx.x - 3));
  mine[0] = row[2 * threadIdx.x] + 0.5;
  mine[1] = row[2 * threadIdx.x + 1] + 0.5;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if ((mode & 2) && ((mode & 1) ? elect_one() : threadIdx.x == 0)) {
    if (mode & 8)
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(ms), "r"(s32(stg)), "r"(sc0), "r"(sc1), "r"(sc2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(ms), "r"(s32(stg)), "r"(sc0), "r"(sc1), "r"(sc2) : "memory");
    if (!(mode & 16)) {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (mode & 4) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      if (mode & 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
  __syncwarp();
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int mode = argc > 2 ? atoi(argv[2]) : 7;
  const int sbox = argc > 3 ? atoi(argv[3]) : 58;
  const int sdt = argc > 4 ? atoi(argv[4]) : 0;
  const int pitch = 128, nrows = 40, nslab = variant == 4 ? 1 : 3;
  std::vector<double> h((size_t)pitch * nrows * nslab), z(h.size(), -1.0);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
  double *d, *o, *dbg;
  cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, h.size() * 8); cudaMalloc(&dbg, 64 * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(o, z.data(), h.size() * 8, cudaMemcpyHostToDevice);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeFn enc = (EncodeFn)fn;
  CUtensorMap mp[2];
  cuuint64_t gd[3] = {(cuuint64_t)pitch, (cuuint64_t)nrows, (cuuint64_t)nslab};
  cuuint64_t gs[2] = {(cuuint64_t)pitch * 8, (cuuint64_t)pitch * nrows * 8};
  cuuint32_t es[3] = {1, 1, 1};
  cuuint32_t bl[3] = {64, 1, 1}, bs[3] = {(cuuint32_t)(sdt == 2 ? 2 * sbox : sbox), 1, 1};
  cuuint64_t gds[3] = {(cuuint64_t)(sdt == 2 ? 2 * pitch : pitch), (cuuint64_t)nrows, (cuuint64_t)nslab};
  const CUtensorMapDataType sdtype = sdt == 2 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : sdt == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  CUresult r1 = enc(&mp[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, gd, gs, bl, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&mp[1], sdtype, 3, o, gds, gs, bs, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d mode %d sbox %d sdt %d encode rc=%d %d\n", variant, mode, sbox, sdt, (int)r1, (int)r2);
  CUtensorMap* dm;
  cudaMalloc(&dm, sizeof mp);
  cudaMemcpy(dm, mp, sizeof mp, cudaMemcpyHostToDevice);
  int lc0 = 4, lc1 = 5, lc2 = nslab > 1 ? 1 : 0, sc0 = 7, sc1 = 6, sc2 = nslab > 1 ? 2 : 0;
  if (sdt == 2) sc0 *= 2;
  if (variant == 1) sc0 = -1;
  if (variant == 2) sc1 = -(1 << 24);
  if (variant == 3) { lc0 = -4; lc1 = -2; }
  k<<<1, 32, 4096>>>(dm, lc0, lc1, lc2, sc0, sc1, sc2, dbg, mode);
  cudaError_t le = cudaGetLastError(), se = cudaDeviceSynchronize();
  printf("launch: %s; sync: %s\n", cudaGetErrorString(le), cudaGetErrorString(se));
  if (se != cudaSuccess) return 1;
  double res[64];
  cudaMemcpy(res, dbg, sizeof res, cudaMemcpyDeviceToHost);
  cudaMemcpy(z.data(), o, h.size() * 8, cudaMemcpyDeviceToHost);
  const double base = (double)lc2 * pitch * nrows + (double)lc1 * pitch + lc0;
  printf("loaded[0..5]= %g %g %g %g %g %g (expect %g.. or zero fill)\n", res[0], res[1], res[2], res[3], res[4], res[5], base);
  long nw = 0, first = -1, last = -1;
  for (size_t i = 0; i < z.size(); ++i) if (z[i] != -1.0) { ++nw; if (first < 0) first = (long)i; last = (long)i; }
  const long want0 = (long)sc2 * pitch * nrows + (long)sc1 * pitch + sc0;
  printf("stored cells %ld first %ld last %ld (box starts at %ld); value at first %g (expect loaded[3 + clip]+0.5)\n", nw, first, last,
         want0, first >= 0 ? z[first] : 0.0);
  return 0;
}
