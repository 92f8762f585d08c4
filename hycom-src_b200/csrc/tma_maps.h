// Tensor maps (TMA descriptors) of the device mirrors: host-side construction.
// A mirror buffer of kdm slabs is a 3-D tensor (pitch, nrows, kdm) of fp64 (masks: uint8,
// depth 1); the marching kernel fetches one row of 32*NC columns per request with
// cp.async.bulk.tensor.3d.  Out-of-range coordinates (the apron of the first/last strip and
// chunk) are zero-filled by the hardware.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace tsadvc {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, void* base,
                      const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box);

// (pitch, nrows, nslab) tensor of doubles; box = (box_cols, 1, box_slabs)
int make_map_f64(CUtensorMap* map, const double* base, int pitch, int nrows, int nslab, int box_cols,
                 int box_slabs);
// (pitch, nrows, kdm, nplanes) tensor of doubles; box = (box_cols, 1, 1, box_planes)
int make_map_f64_4d(CUtensorMap* map, const double* base, int pitch, int nrows, int kdm, int nplanes,
                    int box_cols, int box_planes);

}  // namespace tsadvc
