// Pointwise (1x1x1) Conv3d as GEMM on CUDA cores: the fp32 parity path and the shape-generic
// fallback of the bf16 path (the tcgen05 tensor-core kernels live in pwconv_tc.cu).
// Replaces the cuDNN/cuBLAS kernels behind conv1x1x1 (x3d.py:98-103).
//   fwd  : Y[m][n]  = sum_k X[row(m)][k] * W[n][k]          (+ per-sample column statistics)
//   dgrad: dX[row(m)][k] (+)= sum_n dY[m][n] * Wt[k][n]     (same kernel, scatter on the output rows)
//   wgrad: dW[n][k] += sum_m dY[m][n] * X[row(m)][k]        (split-M, fp32 atomics)
#include "common.cuh"

using namespace x3d;

namespace x3d {
int pwconv_fwd_tc(const void* x, const void* w, void* y, int64_t M, int64_t Kp, int64_t Np, int64_t P_out,
                  const int* gather, const int* scatter, double* stats, cudaStream_t stream, bool* handled);
int pwconv_wgrad_tc(const void* x, const void* dy, float* dw, int64_t M, int64_t K, int64_t Kp, int64_t Nn,
                    int64_t Np, const int* gather, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                    bool* handled);
}

struct RowMap {
  int T, H, W, Ho, Wo, stride;  // maps a dense (n,t,ho,wo) row to the strided row of the big tensor
  __device__ __forceinline__ int64_t operator()(int64_t m) const {
    if (stride == 1) return m;
    const int wo = (int)(m % Wo);
    int64_t r = m / Wo;
    const int ho = (int)(r % Ho);
    r /= Ho;  // n*T + t
    return (r * H + (int64_t)ho * stride) * W + (int64_t)wo * stride;
  }
};

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int MAXS = 8;  // samples a tile may span on the fast statistics path

// A: [.][lda] gathered rows; B: [Nn][ldb] (row n holds the K coefficients); C: [.][ldc]
template <typename T, bool MAP_A, bool MAP_C, bool STATS, bool ACCUM>
__global__ void __launch_bounds__(256) pw_gemm_kernel(const T* __restrict__ A, int lda, const T* __restrict__ B,
                                                       int ldb, T* __restrict__ C, int ldc, int64_t M, int K, int Nn,
                                                       RowMap map, int64_t P_out, double* __restrict__ stats) {
  x3d::pdl_prologue();
  constexpr int VEC = Vec<T>::N;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  // fp64: sums of fp32 values are exact in double (order independent) -> run-to-run reproducible statistics
  __shared__ double s_stat[STATS ? MAXS * BN * 2 : 1];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // loader assignment: BK/VEC vectors per row, 64 rows
  constexpr int VPR = BK / VEC;          // vectors per row per k-tile (2 for bf16, 4 for fp32)
  constexpr int NLOAD = 64 * VPR;        // vector loads per tile (128 / 256)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // each thread loads (at most) one A vector and one B vector per k-tile
  const int lrow = tid / VPR, lvec = tid % VPR;
  const bool loader = tid < NLOAD;
  const T* a_ptr = nullptr;
  const T* b_ptr = nullptr;
  if (loader) {
    const int64_t m = m0 + lrow;
    if (m < M) a_ptr = A + (MAP_A ? map(m) : m) * (int64_t)lda;
    if (n0 + lrow < Nn) b_ptr = B + (int64_t)(n0 + lrow) * ldb;
  }

  for (int k0 = 0; k0 < K; k0 += BK) {
    if (loader) {
      const int k = k0 + lvec * VEC;
      float va[VEC], vb[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) va[j] = vb[j] = 0.f;
      if (a_ptr && k < K) load_vec<T>(a_ptr + k, va);
      if (b_ptr && k < K) load_vec<T>(b_ptr + k, vb);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        As[lvec * VEC + j][lrow] = va[j];
        Bs[lvec * VEC + j][lrow] = vb[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ---------------------------------------------------------------------
  const int nc = n0 + tx * 4;
  const bool col_ok = nc < Nn;  // Nn % 4 == 0
  int64_t s_first = 0;
  int s_count = 1;
  if (STATS) {
    s_first = m0 / P_out;
    const int64_t m_last = (m0 + BM - 1 < M - 1) ? m0 + BM - 1 : M - 1;
    s_count = (int)(m_last / P_out - s_first) + 1;
    if (s_count <= MAXS) {
      for (int i = tid; i < s_count * BN * 2; i += 256) s_stat[i] = 0.0;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M || !col_ok) continue;
    T* cp = C + (MAP_C ? map(m) : m) * (int64_t)ldc + nc;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j];
    if (ACCUM) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += to_float<T>(cp[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) cp[j] = from_float<T>(v[j]);
    if (STATS) {
      const int64_t s = m / P_out;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float r = round_to<T>(v[j]);
        if (s_count <= MAXS) {
          double* sp = &s_stat[((int)(s - s_first) * BN + tx * 4 + j) * 2];
          atomicAdd(sp, (double)r);
          atomicAdd(sp + 1, (double)(r * r));
        } else {
          double* gp = stats + (s * (int64_t)ldc + nc + j) * 2;
          atomicAdd(gp, (double)r);
          atomicAdd(gp + 1, (double)(r * r));
        }
      }
    }
  }
  if (STATS && s_count <= MAXS) {
    __syncthreads();
    for (int i = tid; i < s_count * BN * 2; i += 256) {
      const int s = i / (BN * 2), rem = i % (BN * 2);
      const int col = n0 + rem / 2;
      const float v = (float)s_stat[i];          // the CTA's partial as an fp32 value: the cross-CTA fp64 sum stays exact
      if (col < Nn && v != 0.f) atomicAdd(&stats[((s_first + s) * (int64_t)ldc + col) * 2 + (rem & 1)], (double)v);
    }
  }
}

static RowMap make_map(int64_t T_, int64_t H, int64_t W, int stride) {
  RowMap m;
  m.T = (int)T_; m.H = (int)H; m.W = (int)W; m.stride = stride;
  m.Ho = (int)((H - 1) / stride + 1);
  m.Wo = (int)((W - 1) / stride + 1);
  return m;
}

extern "C" int x3d_pwconv_fwd(const void* x, const void* w, void* y, int64_t N, int64_t T_, int64_t H, int64_t W,
                              int64_t Kp, int64_t Np, int stride, double* stats, x3d_dtype_t dt,
                              x3d_stream_t stream) {
  X3D_CHECK_ARG(Kp % 8 == 0 && Np % 8 == 0, "Kp, Np must be multiples of 8");
  X3D_CHECK_ARG(stride == 1 || stride == 2, "stride must be 1 or 2");
  RowMap map = make_map(T_, H, W, stride);
  const int64_t P_out = T_ * map.Ho * map.Wo;
  const int64_t M = N * P_out;
  if (M == 0) return 0;
  if (dt == X3D_BF16) {
    bool handled = false;
    const int gather[5] = {stride, (int)H, (int)W, map.Ho, map.Wo};
    int rc = pwconv_fwd_tc(x, w, y, M, Kp, Np, P_out, stride > 1 ? gather : nullptr, nullptr, stats, as_stream(stream), &handled);
    if (handled) {
      count_path(X3D_PATH_PW_FWD_TC);
      return rc;
    }
  }
  count_path(X3D_PATH_PW_FWD_SIMT);
  dim3 grid((unsigned)cdiv(M, BM), (unsigned)cdiv(Np, BN));
#define L_(MA, ST)                                                                                           \
  x3d::launch(pw_gemm_kernel<T, MA, false, ST, false>, grid, 256, 0, as_stream(stream),                               \
      (const T*)x, (int)Kp, (const T*)w, (int)Kp, (T*)y, (int)Np, M, (int)Kp, (int)Np, map, P_out, stats)
  X3D_DISPATCH_DTYPE(dt, {
    if (stride != 1 && stats) L_(true, true);
    else if (stride != 1) L_(true, false);
    else if (stats) L_(false, true);
    else L_(false, false);
  });
#undef L_
  X3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int x3d_pwconv_dgrad(const void* dy, const void* wT, void* dx, int64_t N, int64_t T_, int64_t H,
                                int64_t W, int64_t Kp, int64_t Np, int stride, int accumulate, x3d_dtype_t dt,
                                x3d_stream_t stream) {
  X3D_CHECK_ARG(Kp % 8 == 0 && Np % 8 == 0, "Kp, Np must be multiples of 8");
  X3D_CHECK_ARG(stride == 1 || stride == 2, "stride must be 1 or 2");
  RowMap map = make_map(T_, H, W, stride);
  const int64_t P_out = T_ * map.Ho * map.Wo;
  const int64_t M = N * P_out;
  if (M == 0) return 0;
  if (dt == X3D_BF16) {
    bool handled = false;
    // strided conv: the dense GEMM rows (nt,ho,wo) land on rows (nt, s*ho, s*wo) of dx (untouched rows keep their value);
    // accumulate: the epilogue adds to what dx already holds (residual branch gradient)
    const int scatter[6] = {stride, (int)H, (int)W, map.Ho, map.Wo, accumulate};
    int rc = pwconv_fwd_tc(dy, wT, dx, M, Np, Kp, P_out, nullptr, (stride > 1 || accumulate) ? scatter : nullptr, nullptr,
                           as_stream(stream), &handled);
    if (handled) {
      count_path(X3D_PATH_PW_DGRAD_TC);
      return rc;
    }
  }
  count_path(X3D_PATH_PW_DGRAD_SIMT);
  // GEMM roles: A = dy [M][Np], B = wT [Kp][Np] (row k holds the Np coefficients), C = dx [.][Kp]
  dim3 grid((unsigned)cdiv(M, BM), (unsigned)cdiv(Kp, BN));
#define L_(MC, AC)                                                                                           \
  x3d::launch(pw_gemm_kernel<T, false, MC, false, AC>, grid, 256, 0, as_stream(stream),                               \
      (const T*)dy, (int)Np, (const T*)wT, (int)Np, (T*)dx, (int)Kp, M, (int)Np, (int)Kp, map, P_out, nullptr)
  X3D_DISPATCH_DTYPE(dt, {
    if (stride != 1 && accumulate) L_(true, true);
    else if (stride != 1) L_(true, false);
    else if (accumulate) L_(false, true);
    else L_(false, false);
  });
#undef L_
  X3D_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------
constexpr int WM = 16;  // rows of m per smem step
template <typename T, bool MAP_X>
__global__ void __launch_bounds__(256) pw_wgrad_kernel(const T* __restrict__ X, int ldx, const T* __restrict__ DY,
                                                        int ldy, float* __restrict__ dW, int64_t M, int K, int Kp,
                                                        int Nn, int Np, RowMap map, int64_t m_per_block) {
  x3d::pdl_prologue();
  constexpr int VEC = Vec<T>::N;
  __shared__ __align__(16) float Ds[WM][64 + 4];
  __shared__ __align__(16) float Xs[WM][64 + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;   // ty -> n, tx -> k
  const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int64_t mb = (int64_t)blockIdx.z * m_per_block;
  const int64_t me = (mb + m_per_block < M) ? mb + m_per_block : M;
  constexpr int VPR = 64 / VEC;            // vectors per 64-wide row
  constexpr int NV = WM * VPR;             // vectors per tile (128 bf16 / 256 fp32)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t ms = mb; ms < me; ms += WM) {
    for (int q = tid; q < 2 * NV; q += 256) {
      const bool isx = q >= NV;
      const int qq = isx ? q - NV : q;
      const int r = qq / VPR, cvi = (qq % VPR) * VEC;
      const int64_t m = ms + r;
      float v[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] = 0.f;
      if (m < me) {
        if (isx) {
          if (k0 + cvi < Kp) load_vec<T>(X + (MAP_X ? map(m) : m) * (int64_t)ldx + k0 + cvi, v);
        } else {
          if (n0 + cvi < Np) load_vec<T>(DY + m * (int64_t)ldy + n0 + cvi, v);
        }
      }
      float* dst = isx ? &Xs[r][cvi] : &Ds[r][cvi];
#pragma unroll
      for (int j = 0; j < VEC; ++j) dst[j] = v[j];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < WM; ++r) {
      const float4 d4 = *reinterpret_cast<const float4*>(&Ds[r][ty * 4]);
      const float4 x4 = *reinterpret_cast<const float4*>(&Xs[r][tx * 4]);
      const float d[4] = {d4.x, d4.y, d4.z, d4.w};
      const float x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d[i], x[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= Nn) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K && acc[i][j] != 0.f) atomicAdd(&dW[(int64_t)n * K + k], acc[i][j]);
    }
  }
}

static int pwconv_wgrad_impl(const void* x, const void* dy, float* dw, int64_t N, int64_t T_, int64_t H,
                             int64_t W, int64_t K, int64_t Kp, int64_t Nn, int64_t Np, int stride, void* workspace,
                             size_t workspace_bytes, x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Kp % 8 == 0 && Np % 8 == 0, "Kp, Np must be multiples of 8");
  X3D_CHECK_ARG(stride == 1 || stride == 2, "stride must be 1 or 2");
  RowMap map = make_map(T_, H, W, stride);
  const int64_t M = N * T_ * map.Ho * map.Wo;
  if (M == 0) return 0;
  if (dt == X3D_BF16) {
    bool handled = false;
    const int gather[5] = {stride, (int)H, (int)W, map.Ho, map.Wo};
    int rc = pwconv_wgrad_tc(x, dy, dw, M, K, Kp, Nn, Np, stride > 1 ? gather : nullptr, workspace, workspace_bytes,
                             as_stream(stream), &handled);
    if (handled) {
      count_path(X3D_PATH_PW_WGRAD_TC);
      return rc;
    }
  }
  count_path(X3D_PATH_PW_WGRAD_SIMT);
  const int nt = (int)cdiv(Nn, 64), kt = (int)cdiv(K, 64);
  int64_t splits = cdiv(4 * kNumSMs, (int64_t)nt * kt);
  int64_t max_splits = cdiv(M, 4 * WM);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t mpb = cdiv(cdiv(M, splits), WM) * WM;
  splits = cdiv(M, mpb);
  dim3 grid(nt, kt, (unsigned)splits);
  X3D_DISPATCH_DTYPE(dt, {
    if (stride != 1)
      x3d::launch(pw_wgrad_kernel<T, true>, grid, 256, 0, as_stream(stream), (const T*)x, (int)Kp, (const T*)dy, (int)Np, dw, M,
                                                                   (int)K, (int)Kp, (int)Nn, (int)Np, map, mpb);
    else
      x3d::launch(pw_wgrad_kernel<T, false>, grid, 256, 0, as_stream(stream), (const T*)x, (int)Kp, (const T*)dy, (int)Np, dw, M,
                                                                    (int)K, (int)Kp, (int)Nn, (int)Np, map, mpb);
  });
  X3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int x3d_pwconv_wgrad(const void* x, const void* dy, float* dw, int64_t N, int64_t T_, int64_t H,
                                int64_t W, int64_t K, int64_t Kp, int64_t Nn, int64_t Np, int stride,
                                x3d_dtype_t dt, x3d_stream_t stream) {
  return pwconv_wgrad_impl(x, dy, dw, N, T_, H, W, K, Kp, Nn, Np, stride, nullptr, 0, dt, stream);
}
extern "C" int x3d_pwconv_wgrad_ws(const void* x, const void* dy, float* dw, int64_t N, int64_t T_, int64_t H,
                                   int64_t W, int64_t K, int64_t Kp, int64_t Nn, int64_t Np, int stride,
                                   void* workspace, size_t workspace_bytes, x3d_dtype_t dt, x3d_stream_t stream) {
  return pwconv_wgrad_impl(x, dy, dw, N, T_, H, W, K, Kp, Nn, Np, stride, workspace, workspace_bytes, dt, stream);
}
extern "C" size_t x3d_pwconv_wgrad_workspace_bytes(void) { return (size_t)32 << 20; }
