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

// 2-D bf16 tensor [rows, cols], row pitch ld (elements); box = box_rows x 64 columns, SWIZZLE_128B, zero OOB fill
int make_tmap_bf16(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int box_rows);

}  // namespace fvla
