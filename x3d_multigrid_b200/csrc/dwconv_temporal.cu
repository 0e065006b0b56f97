// Temporal depthwise conv kt x 1 x 1 (the stem's conv1_t, x3d.py:202-208,318: kt = 5, pad kt/2, stride 1), NDHWC.
//
//   y[n,t,p,c] = sum_k w[k][c] * x[n, t + k - kt/2, p, c]            p = flattened (h, w)
//
// A position is independent of its neighbours, so this is a pure streaming kernel: thread = (position, 16-byte
// channel vector) marching over t with the last KT input vectors held in registers -- every element of x is read
// from HBM exactly once and every element of y written once (algorithmic bytes = traffic).  dgrad is the same
// kernel with the taps reversed.  wgrad keeps the KT x VEC partial weight gradients of the thread in registers over
// its whole march, reduces them over the block in shared memory and issues one fp32 red per (channel, tap) per CTA.
// Forward optionally emits the per-(sample, channel) sum / sum of squares of the stored output (bn1 statistics).
#include "common.cuh"

using namespace x3d;

namespace {

constexpr int KT = 5;

// block = cv channel vectors x rows positions; grid = (position chunks, N)
template <typename T, bool FLIP, bool STATS>
__global__ void __launch_bounds__(256)
dw_temporal_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y, int T_, int64_t P, int Cp,
                   int cv, int rows, double* __restrict__ stats) {
  x3d::pdl_prologue();
  constexpr int VEC = Vec<T>::N;
  extern __shared__ float s_acc[];
  const int cvec = threadIdx.x % cv, prow = threadIdx.x / cv;
  const int n = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * rows + prow;
  const int c0 = cvec * VEC;
  const bool live = prow < rows && p < P;
  float wk[KT][VEC];
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int j = 0; j < VEC; ++j) wk[k][j] = w[(FLIP ? KT - 1 - k : k) * Cp + c0 + j];
  float s1[VEC], s2[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) s1[j] = s2[j] = 0.f;
  if (live) {
    const int64_t plane = P * Cp;
    const T* xp = x + ((int64_t)n * T_ * P + p) * Cp + c0;
    T* yp = y + ((int64_t)n * T_ * P + p) * Cp + c0;
    // window[k] = x[t + k - 2]; planes outside [0, T) are zero
    float win[KT][VEC];
#pragma unroll
    for (int k = 0; k < KT; ++k)
#pragma unroll
      for (int j = 0; j < VEC; ++j) win[k][j] = 0.f;
#pragma unroll
    for (int k = KT / 2; k < KT; ++k)
      if (k - KT / 2 < T_) load_vec<T>(xp + (int64_t)(k - KT / 2) * plane, win[k]);
    for (int t = 0; t < T_; ++t) {
      float nxt[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) nxt[j] = 0.f;
      if (t + KT / 2 + 1 < T_) load_vec<T>(xp + (int64_t)(t + KT / 2 + 1) * plane, nxt);   // needed by the NEXT output
      float o[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) a = fmaf(wk[k][j], win[k][j], a);
        o[j] = a;
      }
      store_vec<T>(yp + (int64_t)t * plane, o);
      if (STATS) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float r = round_to<T>(o[j]);
          s1[j] += r;
          s2[j] = fmaf(r, r, s2[j]);
        }
      }
#pragma unroll
      for (int k = 0; k + 1 < KT; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) win[k][j] = win[k + 1][j];
#pragma unroll
      for (int j = 0; j < VEC; ++j) win[KT - 1][j] = nxt[j];
    }
  }
  if (STATS) block_stats_flush<VEC>(s1, s2, cvec, Cp, s_acc, stats + (int64_t)n * Cp * 2);
}

// dw[c][k] += sum_{n,t,p} dy[n,t,p,c] * x[n, t + k - 2, p, c];  persistent blocks over (n, position chunk) units
template <typename T>
__global__ void __launch_bounds__(256)
dw_temporal_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, int T_, int64_t P,
                         int C, int Cp, int cv, int rows, int chunks, int64_t nunits) {
  x3d::pdl_prologue();
  constexpr int VEC = Vec<T>::N;
  extern __shared__ float s_red[];                 // [rows][KT][Cp]
  const int cvec = threadIdx.x % cv, prow = threadIdx.x / cv;
  const int c0 = cvec * VEC;
  float g[KT][VEC];
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int j = 0; j < VEC; ++j) g[k][j] = 0.f;
  const int64_t plane = P * Cp;
  for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int n = (int)(u / chunks);
    const int64_t p = (u - (int64_t)n * chunks) * rows + prow;
    if (prow >= rows || p >= P) continue;
    const T* xp = x + ((int64_t)n * T_ * P + p) * Cp + c0;
    const T* dp = dy + ((int64_t)n * T_ * P + p) * Cp + c0;
    float win[KT][VEC];
#pragma unroll
    for (int k = 0; k < KT; ++k)
#pragma unroll
      for (int j = 0; j < VEC; ++j) win[k][j] = 0.f;
#pragma unroll
    for (int k = KT / 2; k < KT; ++k)
      if (k - KT / 2 < T_) load_vec<T>(xp + (int64_t)(k - KT / 2) * plane, win[k]);
    for (int t = 0; t < T_; ++t) {
      float nxt[VEC], d[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) nxt[j] = 0.f;
      if (t + KT / 2 + 1 < T_) load_vec<T>(xp + (int64_t)(t + KT / 2 + 1) * plane, nxt);
      load_vec<T>(dp + (int64_t)t * plane, d);
#pragma unroll
      for (int k = 0; k < KT; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) g[k][j] = fmaf(d[j], win[k][j], g[k][j]);
#pragma unroll
      for (int k = 0; k + 1 < KT; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) win[k][j] = win[k + 1][j];
#pragma unroll
      for (int j = 0; j < VEC; ++j) win[KT - 1][j] = nxt[j];
    }
  }
  // block reduction without atomics: every thread parks its KT x VEC partials in its own row of shared memory,
  // then one thread per (tap, channel) sums the rows and issues ONE fp32 red
  float* mine = s_red + (size_t)prow * KT * Cp;
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int j = 0; j < VEC; ++j) mine[k * Cp + c0 + j] = g[k][j];
  __syncthreads();
  for (int i = threadIdx.x; i < KT * Cp; i += blockDim.x) {
    const int k = i / Cp, ch = i - k * Cp;
    float v = 0.f;
    for (int r = 0; r < rows; ++r) v += s_red[(size_t)r * KT * Cp + i];
    if (ch < C && v != 0.f) atomicAdd(&dw[(int64_t)ch * KT + k], v);
  }
}

struct TGeom {
  int cv, rows, threads, chunks;
};
template <typename T>
TGeom temporal_geom(int64_t P, int64_t Cp) {
  TGeom g;
  g.cv = (int)(Cp / Vec<T>::N);
  g.rows = 256 / g.cv;
  if (g.rows < 1) g.rows = 1;
  if ((int64_t)g.rows > P) g.rows = (int)P;
  g.threads = g.cv * g.rows;
  g.chunks = (int)cdiv(P, g.rows);
  return g;
}

}  // namespace

namespace x3d {

// forward (flip = 0) / dgrad (flip = 1) of the kt x 1 x 1 depthwise conv; returns handled = false for shapes it does
// not cover (kt != 5, more than 256 channel vectors)
int dwconv_temporal(const void* x, const float* w_packed, void* y, int64_t N, int64_t T_, int64_t P, int64_t Cp, int kt,
                    int flip, double* stats, x3d_dtype_t dt, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (kt != KT || N > 65535 || (flip && stats)) return 0;
  X3D_DISPATCH_DTYPE(dt, {
    if (Cp / Vec<T>::N > 256) return 0;
    TGeom g = temporal_geom<T>(P, Cp);
    dim3 grid((unsigned)g.chunks, (unsigned)N);
    const size_t smem = stats ? (size_t)g.rows * Cp * 2 * sizeof(float) : 0;
    if (smem > 48 * 1024) return 0;
    if (flip)
      x3d::launch(dw_temporal_kernel<T, true, false>, grid, g.threads, 0, stream, (const T*)x, w_packed, (T*)y, (int)T_, P,
                  (int)Cp, g.cv, g.rows, (double*)nullptr);
    else if (stats)
      x3d::launch(dw_temporal_kernel<T, false, true>, grid, g.threads, smem, stream, (const T*)x, w_packed, (T*)y, (int)T_,
                  P, (int)Cp, g.cv, g.rows, stats);
    else
      x3d::launch(dw_temporal_kernel<T, false, false>, grid, g.threads, 0, stream, (const T*)x, w_packed, (T*)y, (int)T_,
                  P, (int)Cp, g.cv, g.rows, (double*)nullptr);
  });
  *handled = true;
  return 0;
}

int dwconv_temporal_wgrad(const void* x, const void* dy, float* dw, int64_t N, int64_t T_, int64_t P, int64_t C,
                          int64_t Cp, int kt, x3d_dtype_t dt, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (kt != KT) return 0;
  X3D_DISPATCH_DTYPE(dt, {
    if (Cp / Vec<T>::N > 256) return 0;
    TGeom g = temporal_geom<T>(P, Cp);
    const int64_t nunits = (int64_t)g.chunks * N;
    int64_t blocks = nunits < 4 * kNumSMs ? nunits : 4 * kNumSMs;
    const size_t smem = (size_t)g.rows * KT * Cp * sizeof(float);
    if (smem > 48 * 1024) return 0;
    x3d::launch(dw_temporal_wgrad_kernel<T>, (unsigned)blocks, g.threads, smem, stream, (const T*)x, (const T*)dy, dw,
                (int)T_, P, (int)C, (int)Cp, g.cv, g.rows, g.chunks, nunits);
  });
  *handled = true;
  return 0;
}

}  // namespace x3d
