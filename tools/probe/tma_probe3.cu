// probe 3: (a) 1-D bulk copy cp.async.bulk (no descriptor), (b) hexdump of an encoded tensor map
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);
__global__ void k_bulk(const double* src, double* out) {
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
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(256) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(256), "r"(bar) : "memory");
  }
  uint32_t ok = 0; int spin = 0;
  while (!ok && spin < 1000000) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    ++spin;
  }
  out[threadIdx.x] = ok ? smem[threadIdx.x] : -777.0;
}
int main() {
  const int pitch = 64, nrows = 40;
  std::vector<double> h((size_t)pitch * nrows);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
  double *d, *out; cudaMalloc(&d, h.size() * 8); cudaMalloc(&out, 32 * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
  printf("driver %d runtime %d  d=%p\n", drv, rt, (void*)d);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fn;
  CUtensorMap mp;
  cuuint64_t dims[2] = {pitch, nrows}, str[1] = {pitch * 8};
  cuuint32_t box[2] = {32, 1}, es[2] = {1, 1};
  CUresult r = enc(&mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  const uint64_t* w = (const uint64_t*)&mp;
  for (int i = 0; i < 16; ++i) printf("  [%2d] %016llx\n", i, (unsigned long long)w[i]);
  k_bulk<<<1, 32>>>(d + 5 * pitch + 4, out);
  cudaError_t e = cudaDeviceSynchronize();
  double res[32]; cudaMemcpy(res, out, sizeof res, cudaMemcpyDeviceToHost);
  printf("bulk 1-D: %s  out[0..3]= %g %g %g %g (expect %d..)\n", cudaGetErrorString(e), res[0], res[1], res[2], res[3], 5 * pitch + 4);
  return 0;
}
