// Tiled shared-memory depthwise 3x3x3 kernels (hot shapes).  Placeholder: not yet enabled.
#include "common.cuh"
namespace x3d {
int dwconv_fwd_tiled(const void*, const float*, void*, int64_t, int64_t, int64_t, int64_t, int64_t, int, const float*,
                     const float*, int, int, double*, x3d_dtype_t, cudaStream_t, bool* handled) {
  *handled = false;
  return 0;
}
}  // namespace x3d
