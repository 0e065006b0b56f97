// shared declarations of the tiled depthwise kernels (see dwconv_tiled_impl.cuh)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace x3d {
struct DwTiledArgs {
  const float* w;
  void* y;
  const float* scale;
  const float* shift;
  int splits;
  const void* aux;
  double* stats;
};
// mode: 0 fwd s1, 1 fwd s2, 2 dgrad s1, 3 dgrad s2
int dw_tiled_run_bf16(int mode, const void* in, int64_t N, int T_, int Hin, int Win, int Ho, int Wo, int Cp,
                      const DwTiledArgs& a, cudaStream_t stream, bool nan_fill, bool* handled);
int dw_tiled_run_f32(int mode, const void* in, int64_t N, int T_, int Hin, int Win, int Ho, int Wo, int Cp,
                     const DwTiledArgs& a, cudaStream_t stream, bool nan_fill, bool* handled);
int dw_wgrad_tiled_bf16(int stride, const void* x, const void* dy, float* dw, int64_t N, int T_, int H, int W, int Ho,
                        int Wo, int C, int Cp, const float* scale, const float* shift, int splits,
                        cudaStream_t stream, bool* handled);
int dw_wgrad_tiled_f32(int stride, const void* x, const void* dy, float* dw, int64_t N, int T_, int H, int W, int Ho,
                       int Wo, int C, int Cp, const float* scale, const float* shift, int splits,
                       cudaStream_t stream, bool* handled);
}  // namespace x3d
