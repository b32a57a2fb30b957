// Host-side access to cuTensorMapEncodeTiled (resolved through the runtime, no -lcuda link dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace fvla {

using TmaEncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                      const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmaEncodeTiledFn tma_encode_fn();  // null when the driver does not export it

}  // namespace fvla
