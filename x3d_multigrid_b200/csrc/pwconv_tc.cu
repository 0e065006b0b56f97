// tcgen05 / TMEM / TMA pointwise-conv GEMM (bf16).  Placeholder: not yet enabled.
#include "common.cuh"
namespace x3d {
int pwconv_fwd_tc(const void*, const void*, void*, int64_t, int64_t, int64_t, int64_t, double*, cudaStream_t,
                  bool* handled) {
  *handled = false;
  return 0;
}
}  // namespace x3d
