// TMA probe 2: official libcu++ wrappers vs raw PTX, rank-2 maps, box sizes
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);

__global__ void k_official(const __grid_constant__ CUtensorMap tm, double* out, int x, int y, int bytes) {
  __shared__ alignas(128) double smem[256];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem, &tm, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  out[threadIdx.x] = smem[threadIdx.x];
}

__global__ void k_raw2d(const __grid_constant__ CUtensorMap tm, double* out, int x, int y, int bytes) {
  __shared__ alignas(128) double smem[256];
  __shared__ alignas(8) unsigned long long barv;
  uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
  uint32_t bar = (uint32_t)__cvta_generic_to_shared(&barv);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(&tm), "r"(x), "r"(y), "r"(bar) : "memory");
  }
  uint32_t ok = 0; int spin = 0;
  while (!ok && spin < 1000000) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    ++spin;
  }
  out[threadIdx.x] = ok ? smem[threadIdx.x] : -777.0;
}

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int pitch = 64, nrows = 40;
  std::vector<double> h((size_t)pitch * nrows);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
  double *d, *out; cudaMalloc(&d, h.size() * 8); cudaMalloc(&out, 32 * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fn;
  CUtensorMap mp;
  int bw = (variant % 10 == 1) ? 16 : 32;
  cuuint64_t dims[2] = {pitch, nrows}, str[1] = {pitch * 8};
  cuuint32_t box[2] = {(cuuint32_t)bw, 1}, es[2] = {1, 1};
  CUresult r = enc(&mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode rc=%d query=%d\n", variant, (int)r, (int)q);
  if (variant >= 20) k_official<<<1, 32>>>(mp, out, 3, 5, bw * 8);
  else k_raw2d<<<1, 32>>>(mp, out, 3, 5, bw * 8);
  cudaError_t e = cudaDeviceSynchronize();
  double res[32]; cudaMemcpy(res, out, sizeof res, cudaMemcpyDeviceToHost);
  printf("  %s  out[0..3]= %g %g %g %g (expect %d..)\n", cudaGetErrorString(e), res[0], res[1], res[2], res[3], 5 * pitch + 3);
  return 0;
}
