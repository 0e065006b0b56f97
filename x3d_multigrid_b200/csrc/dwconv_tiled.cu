// Dispatch of the tiled depthwise 3x3x3 kernels (forward and dgrad); the kernels themselves are in
// dwconv_tiled_impl.cuh, instantiated per storage type in dwconv_tiled_{bf16,f32}.cu.
#include <stdlib.h>

#include "common.cuh"
#include "dwconv_tiled.h"

using namespace x3d;

namespace {
int finish(const char* what, bool handled) {
  if (!handled) return 0;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  count_launch();
  return 0;
}
bool force_direct() {
  static const bool v = getenv("X3D_DW_DIRECT") != nullptr;   // A/B switch for tests and profiling
  return v;
}
}  // namespace

namespace x3d {
int dwconv_fwd_tiled(const void* x, const float* w_packed, void* y, int64_t N, int64_t T_, int64_t H, int64_t W,
                     int64_t Cp, int stride, const float* in_scale, const float* in_shift, int splits, int relu_in,
                     double* stats, x3d_dtype_t dt, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (force_direct()) return 0;
  if (in_scale != nullptr && !relu_in) return 0;   // the fused transform of the tiled kernel is BN + ReLU
  const int Ho = (int)((H + 2 - 3) / stride + 1), Wo = (int)((W + 2 - 3) / stride + 1);
  const DwTiledArgs a{w_packed, y, in_scale, in_shift, splits, nullptr, stats};
  const int mode = stride == 1 ? 0 : 1;
  if (dt == X3D_BF16)
    dw_tiled_run_bf16(mode, x, N, (int)T_, (int)H, (int)W, Ho, Wo, (int)Cp, a, stream, in_scale != nullptr, handled);
  else
    dw_tiled_run_f32(mode, x, N, (int)T_, (int)H, (int)W, Ho, Wo, (int)Cp, a, stream, in_scale != nullptr, handled);
  return finish("dwconv_fwd_tiled", *handled);
}

// dx[N][T][H][W][Cp] from dy[N][T][Ho][Wo][Cp]; optional mask/statistics epilogue with the saved conv1 output
int dwconv_dgrad_tiled(const void* dy, const float* w_packed, void* dx, int64_t N, int64_t T_, int64_t H, int64_t W,
                       int64_t Cp, int stride, const void* mask_src, const float* mask_scale,
                       const float* mask_shift, int splits, double* stats, x3d_dtype_t dt, cudaStream_t stream,
                       bool* handled) {
  *handled = false;
  if (force_direct()) return 0;
  const int Ho = (int)((H + 2 - 3) / stride + 1), Wo = (int)((W + 2 - 3) / stride + 1);
  const DwTiledArgs a{w_packed, dx, mask_scale, mask_shift, splits, mask_src, mask_src ? stats : nullptr};
  const int mode = stride == 1 ? 2 : 3;
  if (dt == X3D_BF16)
    dw_tiled_run_bf16(mode, dy, N, (int)T_, Ho, Wo, (int)H, (int)W, (int)Cp, a, stream, false, handled);
  else
    dw_tiled_run_f32(mode, dy, N, (int)T_, Ho, Wo, (int)H, (int)W, (int)Cp, a, stream, false, handled);
  return finish("dwconv_dgrad_tiled", *handled);
}

// dw[C][27] += wgrad(x (conv input, optional fused BN+ReLU), dy)
int dwconv_wgrad_tiled(const void* x, const void* dy, float* dw, int64_t N, int64_t T_, int64_t H, int64_t W,
                       int64_t C, int64_t Cp, int stride, const float* in_scale, const float* in_shift, int splits,
                       int relu_in, x3d_dtype_t dt, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (force_direct()) return 0;
  if (in_scale != nullptr && !relu_in) return 0;
  const int Ho = (int)((H + 2 - 3) / stride + 1), Wo = (int)((W + 2 - 3) / stride + 1);
  if (dt == X3D_BF16)
    dw_wgrad_tiled_bf16(stride, x, dy, dw, N, (int)T_, (int)H, (int)W, Ho, Wo, (int)C, (int)Cp, in_scale, in_shift,
                        splits, stream, handled);
  else
    dw_wgrad_tiled_f32(stride, x, dy, dw, N, (int)T_, (int)H, (int)W, Ho, Wo, (int)C, (int)Cp, in_scale, in_shift,
                       splits, stream, handled);
  return finish("dwconv_wgrad_tiled", *handled);
}
}  // namespace x3d
