#include "tma_maps.h"

#include <mutex>

namespace tsadvc {

namespace {
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
std::once_flag g_once;

void resolve() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_encode = (EncodeFn)fn;
}
}  // namespace

CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, void* base,
                      const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box) {
  std::call_once(g_once, resolve);
  if (!g_encode) return CUDA_ERROR_NOT_SUPPORTED;
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return g_encode(map, dt, rank, base, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int make_map_f64(CUtensorMap* map, const double* base, int pitch, int nrows, int nslab, int box_cols,
                 int box_slabs) {
  const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)nrows, (cuuint64_t)nslab};
  const cuuint64_t str[2] = {(cuuint64_t)pitch * 8, (cuuint64_t)pitch * nrows * 8};
  const cuuint32_t box[3] = {(cuuint32_t)box_cols, 1, (cuuint32_t)box_slabs};
  return (int)encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, str, box);
}

int make_map_f64_4d(CUtensorMap* map, const double* base, int pitch, int nrows, int kdm, int nplanes,
                    int box_cols, int box_planes) {
  const cuuint64_t dims[4] = {(cuuint64_t)pitch, (cuuint64_t)nrows, (cuuint64_t)kdm, (cuuint64_t)nplanes};
  const cuuint64_t str[3] = {(cuuint64_t)pitch * 8, (cuuint64_t)pitch * nrows * 8,
                             (cuuint64_t)pitch * nrows * 8 * (cuuint64_t)kdm};
  const cuuint32_t box[4] = {(cuuint32_t)box_cols, 1, 1, (cuuint32_t)box_planes};
  return (int)encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)base, dims, str, box);
}

}  // namespace tsadvc
