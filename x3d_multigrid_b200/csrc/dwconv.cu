// Channelwise (depthwise) Conv3d: forward, dgrad, wgrad.  NDHWC, 16-byte channel vectors.
// Replaces ATen's conv_depthwise3d_cuda_* behind conv3x3x3 (x3d.py:87-95) and conv1_t (x3d.py:202-208).
//
// "direct" kernels: one thread per output channel-vector, taps gathered through L1/L2.  They are the
// shape-generic path (any kernel extent, any stride); the tiled shared-memory kernels for the 3x3x3
// hot shapes live in dwconv_tiled.cu and are selected by x3d_dwconv_* when applicable.
#include "common.cuh"

using namespace x3d;

namespace x3d {
int dwconv_fwd_tiled(const void* x, const float* w_packed, void* y, int64_t N, int64_t T, int64_t H, int64_t W,
                     int64_t Cp, int stride, const float* in_scale, const float* in_shift, int splits, int relu_in,
                     double* stats, x3d_dtype_t dt, cudaStream_t stream, bool* handled);
int dwconv_dgrad_tiled(const void* dy, const float* w_packed, void* dx, int64_t N, int64_t T, int64_t H, int64_t W,
                       int64_t Cp, int stride, const void* mask_src, const float* mask_scale,
                       const float* mask_shift, int splits, double* stats, x3d_dtype_t dt, cudaStream_t stream,
                       bool* handled);
int dwconv_wgrad_tiled(const void* x, const void* dy, float* dw, int64_t N, int64_t T, int64_t H, int64_t W,
                       int64_t C, int64_t Cp, int stride, const float* in_scale, const float* in_shift, int splits,
                       int relu_in, x3d_dtype_t dt, cudaStream_t stream, bool* handled);
// dwconv_temporal.cu: kt x 1 x 1 streaming kernels (stem conv1_t)
int dwconv_temporal(const void* x, const float* w_packed, void* y, int64_t N, int64_t T_, int64_t P, int64_t Cp, int kt,
                    int flip, double* stats, x3d_dtype_t dt, cudaStream_t stream, bool* handled);
int dwconv_temporal_wgrad(const void* x, const void* dy, float* dw, int64_t N, int64_t T_, int64_t P, int64_t C,
                          int64_t Cp, int kt, x3d_dtype_t dt, cudaStream_t stream, bool* handled);
}

struct DwGeom {
  int T, H, W, Ho, Wo, Cp, stride;
};

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
template <typename T, int KT, int KH, int KW, bool XFORM, bool RELU, bool STATS>
__global__ void dw_fwd_direct_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                                     DwGeom g, const float* __restrict__ scale, const float* __restrict__ shift,
                                     int splits, double* __restrict__ stats, int64_t P, int cv, int rows,
                                     int64_t chunk) {
  x3d::pdl_prologue();
  extern __shared__ float s_acc[];
  const int Cp = g.Cp;
  constexpr int VEC = Vec<T>::N;
  const int n = blockIdx.y;
  const int cvec = threadIdx.x % cv;
  const int prow = threadIdx.x / cv;
  const int64_t p0 = (int64_t)blockIdx.x * chunk;
  const int64_t p1 = (p0 + chunk < P) ? p0 + chunk : P;
  const int c0 = cvec * VEC;
  const int b = n % splits;
  float sc[VEC], sh[VEC], a0[VEC], a1[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    sc[j] = XFORM ? scale[b * Cp + c0 + j] : 1.f;
    sh[j] = XFORM ? shift[b * Cp + c0 + j] : 0.f;
    a0[j] = a1[j] = 0.f;
  }
  const int64_t in_base = (int64_t)n * g.T * g.H * g.W;
  for (int64_t p = p0 + prow; p < p1; p += rows) {
    const int wo = (int)(p % g.Wo);
    const int ho = (int)((p / g.Wo) % g.Ho);
    const int t = (int)(p / ((int64_t)g.Wo * g.Ho));
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      const int tt = t + i - KT / 2;
      if (tt < 0 || tt >= g.T) continue;
#pragma unroll
      for (int jh = 0; jh < KH; ++jh) {
        const int hh = ho * g.stride + jh - KH / 2;
        if (hh < 0 || hh >= g.H) continue;
#pragma unroll
        for (int k = 0; k < KW; ++k) {
          const int ww = wo * g.stride + k - KW / 2;
          if (ww < 0 || ww >= g.W) continue;
          float xv[VEC], wv[VEC];
          load_vec<T>(x + (in_base + ((int64_t)tt * g.H + hh) * g.W + ww) * Cp + c0, xv);
          const float* wp = w + (int64_t)((i * KH + jh) * KW + k) * Cp + c0;
#pragma unroll
          for (int j = 0; j < VEC; j += 4) {
            float4 q = *reinterpret_cast<const float4*>(wp + j);
            wv[j] = q.x; wv[j + 1] = q.y; wv[j + 2] = q.z; wv[j + 3] = q.w;
          }
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            float v = xv[j];
            if (XFORM) {
              v = fmaf(v, sc[j], sh[j]);
              if (RELU) v = fmaxf(v, 0.f);
            }
            acc[j] = fmaf(wv[j], v, acc[j]);
          }
        }
      }
    }
    store_vec<T>(y + ((int64_t)n * P + p) * Cp + c0, acc);
    if (STATS) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float r = round_to<T>(acc[j]);
        a0[j] += r;
        a1[j] = fmaf(r, r, a1[j]);
      }
    }
  }
  if (STATS) block_stats_flush<VEC>(a0, a1, cvec, Cp, s_acc, stats + (int64_t)n * Cp * 2);
}

template <typename T, int KT, int KH, int KW>
static int launch_dw_fwd_direct(const void* x, const float* w, void* y, int64_t N, DwGeom g, const float* scale,
                                const float* shift, int splits, int relu_in, double* stats, cudaStream_t stream) {
  const int64_t P = (int64_t)g.T * g.Ho * g.Wo;
  RowGeom rg = make_row_geom<T>(N, P, g.Cp, 8 * kNumSMs);
  dim3 grid(rg.chunks, (unsigned)N);
  size_t smem = stats ? (size_t)rg.rows * g.Cp * 2 * sizeof(float) : 0;
#define L_(XF, RL, ST)                                                                                          \
  x3d::launch(dw_fwd_direct_kernel<T, KT, KH, KW, XF, RL, ST>, grid, rg.threads, smem, stream,                           \
      (const T*)x, w, (T*)y, g, scale, shift, splits, stats, P, rg.cv, rg.rows, rg.chunk)
  const bool xf = scale != nullptr;
  if (xf && relu_in && stats) L_(true, true, true);
  else if (xf && relu_in) L_(true, true, false);
  else if (xf && stats) L_(true, false, true);
  else if (xf) L_(true, false, false);
  else if (stats) L_(false, false, true);
  else L_(false, false, false);
#undef L_
  return 0;
}

extern "C" int x3d_dwconv_fwd(const void* x, const float* w_packed, void* y, int64_t N, int64_t T_, int64_t H,
                              int64_t W, int64_t Cp, int kt, int kh, int kw, int stride, const float* in_scale,
                              const float* in_shift, int splits, int relu_in, double* stats, x3d_dtype_t dt,
                              x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  X3D_CHECK_ARG((kt == 3 && kh == 3 && kw == 3) || (kt == 5 && kh == 1 && kw == 1), "kernel must be 3x3x3 or 5x1x1");
  X3D_CHECK_ARG(stride == 1 || stride == 2, "stride must be 1 or 2");
  X3D_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "scale/shift must come together");
  if (splits < 1) splits = 1;
  DwGeom g;
  g.T = (int)T_; g.H = (int)H; g.W = (int)W; g.Cp = (int)Cp; g.stride = stride;
  g.Ho = (int)((H + 2 * (kh / 2) - kh) / stride + 1);
  g.Wo = (int)((W + 2 * (kw / 2) - kw) / stride + 1);
  if (N * T_ * g.Ho * g.Wo == 0) return 0;
  if (kt == 5 && stride == 1 && in_scale == nullptr) {      // stem conv1_t: streaming temporal kernel
    bool handled = false;
    int rc = dwconv_temporal(x, w_packed, y, N, T_, H * W, Cp, kt, 0, stats, dt, as_stream(stream), &handled);
    if (handled) {
      count_path(X3D_PATH_DW_FWD_TEMPORAL);
      if (rc == 0) X3D_LAUNCH_CHECK();
      return rc;
    }
  }
  if (kt == 3) {
    bool handled = false;
    int rc = dwconv_fwd_tiled(x, w_packed, y, N, T_, H, W, Cp, stride, in_scale, in_shift, splits, relu_in, stats, dt,
                              as_stream(stream), &handled);
    if (handled) {
      count_path(X3D_PATH_DW_FWD_TILED);
      return rc;
    }
  }
  count_path(X3D_PATH_DW_FWD_DIRECT);
  X3D_DISPATCH_DTYPE(dt, {
    if (kt == 3) launch_dw_fwd_direct<T, 3, 3, 3>(x, w_packed, y, N, g, in_scale, in_shift, splits, relu_in, stats, as_stream(stream));
    else launch_dw_fwd_direct<T, 5, 1, 1>(x, w_packed, y, N, g, in_scale, in_shift, splits, relu_in, stats, as_stream(stream));
  });
  X3D_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------
// dgrad: dx[n,t,h,w,c] = sum_{i,j,k} W[c,i,j,k] * dy[n, t-i+pt, (h-j+ph)/s, (w-k+pw)/s, c]
// ---------------------------------------------------------------------------------------
template <typename T, int KT, int KH, int KW, bool MASK>
__global__ void dw_dgrad_direct_kernel(const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx,
                                       DwGeom g, const T* __restrict__ mask_src, const float* __restrict__ mscale,
                                       const float* __restrict__ mshift, int splits, double* __restrict__ stats,
                                       int64_t P /*input positions*/, int cv, int rows, int64_t chunk) {
  x3d::pdl_prologue();
  extern __shared__ float s_acc[];
  constexpr int VEC = Vec<T>::N;
  const int Cp = g.Cp;
  const int n = blockIdx.y;
  const int cvec = threadIdx.x % cv;
  const int prow = threadIdx.x / cv;
  const int64_t p0 = (int64_t)blockIdx.x * chunk;
  const int64_t p1 = (p0 + chunk < P) ? p0 + chunk : P;
  const int c0 = cvec * VEC;
  const int b = n % splits;
  float sc[VEC], sh[VEC], a0[VEC], a1[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    sc[j] = MASK ? mscale[b * Cp + c0 + j] : 1.f;
    sh[j] = MASK ? mshift[b * Cp + c0 + j] : 0.f;
    a0[j] = a1[j] = 0.f;
  }
  const int64_t out_base = (int64_t)n * g.T * g.Ho * g.Wo;
  for (int64_t p = p0 + prow; p < p1; p += rows) {
    const int wi = (int)(p % g.W);
    const int hi = (int)((p / g.W) % g.H);
    const int t = (int)(p / ((int64_t)g.W * g.H));
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      const int to = t - i + KT / 2;
      if (to < 0 || to >= g.T) continue;
#pragma unroll
      for (int jh = 0; jh < KH; ++jh) {
        const int hn = hi - jh + KH / 2;
        if (hn < 0 || (hn % g.stride) != 0) continue;
        const int ho = hn / g.stride;
        if (ho >= g.Ho) continue;
#pragma unroll
        for (int k = 0; k < KW; ++k) {
          const int wn = wi - k + KW / 2;
          if (wn < 0 || (wn % g.stride) != 0) continue;
          const int wo = wn / g.stride;
          if (wo >= g.Wo) continue;
          float dv[VEC], wv[VEC];
          load_vec<T>(dy + (out_base + ((int64_t)to * g.Ho + ho) * g.Wo + wo) * Cp + c0, dv);
          const float* wp = w + (int64_t)((i * KH + jh) * KW + k) * Cp + c0;
#pragma unroll
          for (int j = 0; j < VEC; j += 4) {
            float4 q = *reinterpret_cast<const float4*>(wp + j);
            wv[j] = q.x; wv[j + 1] = q.y; wv[j + 2] = q.z; wv[j + 3] = q.w;
          }
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] = fmaf(wv[j], dv[j], acc[j]);
        }
      }
    }
    const int64_t off = ((int64_t)n * P + p) * Cp + c0;
    if (MASK) {
      float m[VEC];
      load_vec<T>(mask_src + off, m);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float d = (fmaf(m[j], sc[j], sh[j]) > 0.f) ? round_to<T>(acc[j]) : 0.f;
        acc[j] = d;
        a0[j] += d;
        a1[j] = fmaf(d, m[j], a1[j]);
      }
    }
    store_vec<T>(dx + off, acc);
  }
  if (MASK && stats) block_stats_flush<VEC>(a0, a1, cvec, Cp, s_acc, stats + (int64_t)n * Cp * 2);
}

extern "C" int x3d_dwconv_dgrad(const void* dy, const float* w_packed, void* dx, int64_t N, int64_t T_, int64_t H,
                                int64_t W, int64_t Cp, int kt, int kh, int kw, int stride, const void* mask_src,
                                const float* mask_scale, const float* mask_shift, int splits, double* stats,
                                x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  X3D_CHECK_ARG((kt == 3 && kh == 3 && kw == 3) || (kt == 5 && kh == 1 && kw == 1), "kernel must be 3x3x3 or 5x1x1");
  X3D_CHECK_ARG(stride == 1 || stride == 2, "stride must be 1 or 2");
  X3D_CHECK_ARG(!mask_src || (mask_scale && mask_shift && stats), "mask epilogue needs scale, shift and stats");
  if (splits < 1) splits = 1;
  DwGeom g;
  g.T = (int)T_; g.H = (int)H; g.W = (int)W; g.Cp = (int)Cp; g.stride = stride;
  g.Ho = (int)((H + 2 * (kh / 2) - kh) / stride + 1);
  g.Wo = (int)((W + 2 * (kw / 2) - kw) / stride + 1);
  const int64_t P = T_ * H * W;
  if (N * P == 0) return 0;
  if (kt == 5 && stride == 1 && mask_src == nullptr) {
    bool handled = false;
    int rc = dwconv_temporal(dy, w_packed, dx, N, T_, H * W, Cp, kt, 1, nullptr, dt, as_stream(stream), &handled);
    if (handled) {
      count_path(X3D_PATH_DW_DGRAD_TEMPORAL);
      if (rc == 0) X3D_LAUNCH_CHECK();
      return rc;
    }
  }
  if (kt == 3) {
    bool handled = false;
    int rc = dwconv_dgrad_tiled(dy, w_packed, dx, N, T_, H, W, Cp, stride, mask_src, mask_scale, mask_shift, splits,
                                stats, dt, as_stream(stream), &handled);
    if (handled) {
      count_path(X3D_PATH_DW_DGRAD_TILED);
      return rc;
    }
  }
  count_path(X3D_PATH_DW_DGRAD_DIRECT);
#define L_(KT_, KH_, KW_, MK)                                                                                \
  x3d::launch(dw_dgrad_direct_kernel<T, KT_, KH_, KW_, MK>, grid, rg.threads, smem, as_stream(stream),                   \
      (const T*)dy, w_packed, (T*)dx, g, (const T*)mask_src, mask_scale, mask_shift, splits, stats, P, rg.cv,   \
      rg.rows, rg.chunk)
  X3D_DISPATCH_DTYPE(dt, {
    RowGeom rg = make_row_geom<T>(N, P, Cp, 8 * kNumSMs);
    dim3 grid(rg.chunks, (unsigned)N);
    size_t smem = mask_src ? (size_t)rg.rows * Cp * 2 * sizeof(float) : 0;
    if (kt == 3 && mask_src) L_(3, 3, 3, true);
    else if (kt == 3) L_(3, 3, 3, false);
    else if (mask_src) L_(5, 1, 1, true);
    else L_(5, 1, 1, false);
  });
#undef L_
  X3D_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------
// wgrad: dw[c][i][j][k] += sum dy[n,t,ho,wo,c] * xf[n, t+i-pt, s*ho+j-ph, s*wo+k-pw, c]
// grid.z = kt plane; each thread keeps KH*KW*VEC accumulators.
// ---------------------------------------------------------------------------------------
template <typename T, int KT, int KH, int KW, bool XFORM, bool RELU>
__global__ void dw_wgrad_direct_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                                       DwGeom g, int C, const float* __restrict__ scale,
                                       const float* __restrict__ shift, int splits, int64_t P /*out positions*/,
                                       int cv, int rows, int64_t chunk) {
  x3d::pdl_prologue();
  extern __shared__ float s_w[];  // [KH*KW][Cp]
  constexpr int VEC = Vec<T>::N;
  constexpr int NT = KH * KW;
  const int Cp = g.Cp;
  const int n = blockIdx.y;
  const int i = blockIdx.z;  // temporal tap
  const int cvec = threadIdx.x % cv;
  const int prow = threadIdx.x / cv;
  const int64_t p0 = (int64_t)blockIdx.x * chunk;
  const int64_t p1 = (p0 + chunk < P) ? p0 + chunk : P;
  const int c0 = cvec * VEC;
  const int b = n % splits;
  float sc[VEC], sh[VEC], acc[NT][VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    sc[j] = XFORM ? scale[b * Cp + c0 + j] : 1.f;
    sh[j] = XFORM ? shift[b * Cp + c0 + j] : 0.f;
#pragma unroll
    for (int q = 0; q < NT; ++q) acc[q][j] = 0.f;
  }
  const int64_t in_base = (int64_t)n * g.T * g.H * g.W;
  for (int64_t p = p0 + prow; p < p1; p += rows) {
    const int wo = (int)(p % g.Wo);
    const int ho = (int)((p / g.Wo) % g.Ho);
    const int t = (int)(p / ((int64_t)g.Wo * g.Ho));
    const int tt = t + i - KT / 2;
    if (tt < 0 || tt >= g.T) continue;
    float dv[VEC];
    load_vec<T>(dy + ((int64_t)n * P + p) * Cp + c0, dv);
#pragma unroll
    for (int jh = 0; jh < KH; ++jh) {
      const int hh = ho * g.stride + jh - KH / 2;
      if (hh < 0 || hh >= g.H) continue;
#pragma unroll
      for (int k = 0; k < KW; ++k) {
        const int ww = wo * g.stride + k - KW / 2;
        if (ww < 0 || ww >= g.W) continue;
        float xv[VEC];
        load_vec<T>(x + (in_base + ((int64_t)tt * g.H + hh) * g.W + ww) * Cp + c0, xv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float v = xv[j];
          if (XFORM) {
            v = fmaf(v, sc[j], sh[j]);
            if (RELU) v = fmaxf(v, 0.f);
          }
          acc[jh * KW + k][j] = fmaf(dv[j], v, acc[jh * KW + k][j]);
        }
      }
    }
  }
  for (int q = threadIdx.x; q < NT * Cp; q += blockDim.x) s_w[q] = 0.f;
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NT; ++q)
#pragma unroll
    for (int j = 0; j < VEC; ++j) atomicAdd(&s_w[q * Cp + c0 + j], acc[q][j]);
  __syncthreads();
  for (int q = threadIdx.x; q < NT * Cp; q += blockDim.x) {
    const int tap = q / Cp, c = q % Cp;
    const float v = s_w[q];
    if (c < C && v != 0.f) atomicAdd(&dw[(int64_t)c * (KT * NT) + i * NT + tap], v);
  }
}

extern "C" int x3d_dwconv_wgrad(const void* x, const void* dy, float* dw, int64_t N, int64_t T_, int64_t H,
                                int64_t W, int64_t C, int64_t Cp, int kt, int kh, int kw, int stride,
                                const float* in_scale, const float* in_shift, int splits, int relu_in,
                                x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  X3D_CHECK_ARG((kt == 3 && kh == 3 && kw == 3) || (kt == 5 && kh == 1 && kw == 1), "kernel must be 3x3x3 or 5x1x1");
  X3D_CHECK_ARG(stride == 1 || stride == 2, "stride must be 1 or 2");
  if (splits < 1) splits = 1;
  DwGeom g;
  g.T = (int)T_; g.H = (int)H; g.W = (int)W; g.Cp = (int)Cp; g.stride = stride;
  g.Ho = (int)((H + 2 * (kh / 2) - kh) / stride + 1);
  g.Wo = (int)((W + 2 * (kw / 2) - kw) / stride + 1);
  const int64_t P = T_ * g.Ho * g.Wo;
  if (N * P == 0) return 0;
  if (kt == 5 && stride == 1 && in_scale == nullptr) {
    bool handled = false;
    int rc = dwconv_temporal_wgrad(x, dy, dw, N, T_, H * W, C, Cp, kt, dt, as_stream(stream), &handled);
    if (handled) {
      count_path(X3D_PATH_DW_WGRAD_TEMPORAL);
      if (rc == 0) X3D_LAUNCH_CHECK();
      return rc;
    }
  }
  if (kt == 3) {
    bool handled = false;
    int rc = dwconv_wgrad_tiled(x, dy, dw, N, T_, H, W, C, Cp, stride, in_scale, in_shift, splits, relu_in, dt,
                                as_stream(stream), &handled);
    if (handled) {
      count_path(X3D_PATH_DW_WGRAD_TILED);
      return rc;
    }
  }
  count_path(X3D_PATH_DW_WGRAD_DIRECT);
#define L_(KT_, KH_, KW_, XF, RL)                                                                               \
  x3d::launch(dw_wgrad_direct_kernel<T, KT_, KH_, KW_, XF, RL>, grid, rg.threads, smem, as_stream(stream),               \
      (const T*)x, (const T*)dy, dw, g, (int)C, in_scale, in_shift, splits, P, rg.cv, rg.rows, rg.chunk)
  X3D_DISPATCH_DTYPE(dt, {
    RowGeom rg = make_row_geom<T>(N, P, Cp, 2 * kNumSMs);
    dim3 grid(rg.chunks, (unsigned)N, (unsigned)kt);
    size_t smem = (size_t)kh * kw * Cp * sizeof(float);
    const bool xf = in_scale != nullptr;
    if (kt == 3) {
      if (xf && relu_in) L_(3, 3, 3, true, true);
      else if (xf) L_(3, 3, 3, true, false);
      else L_(3, 3, 3, false, false);
    } else {
      if (xf && relu_in) L_(5, 1, 1, true, true);
      else if (xf) L_(5, 1, 1, true, false);
      else L_(5, 1, 1, false, false);
    }
  });
#undef L_
  X3D_LAUNCH_CHECK();
  return 0;
}
