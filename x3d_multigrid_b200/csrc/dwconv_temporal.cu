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

// Memory-level parallelism is what bounds these kernels: with one 16-byte load in flight per thread and ~120 registers
// (2 CTAs/SM) an SM had ~12 KB outstanding -- 2.6 TB/s.  A thread now keeps the RAW 16-byte vectors of the KT window
// planes plus LOOK planes of look-ahead in a register ring (plane p lives in slot (p + 2) % RING; the march is unrolled
// over RING steps so every slot index is static and a loaded register is not touched before its plane is needed).
constexpr int LOOK = 4;              // weight gradient: two rings (x and dy) next to the 5 x VEC accumulators
constexpr int RING = KT + LOOK;
constexpr int FLOOK = 6;             // forward / dgrad: one ring, deeper look-ahead (8 spills under the 128-register cap)
constexpr int FRING = KT + FLOOK;

template <typename T>
__device__ __forceinline__ void unpack16(const uint4& q, float (&v)[Vec<T>::N]);
template <>
__device__ __forceinline__ void unpack16<float>(const uint4& q, float (&v)[4]) {
  v[0] = __uint_as_float(q.x); v[1] = __uint_as_float(q.y); v[2] = __uint_as_float(q.z); v[3] = __uint_as_float(q.w);
}
template <>
__device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4& q, float (&v)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <typename T>
__device__ __forceinline__ uint4 ldg16(const T* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// block = cv channel vectors x rows positions; grid = (position chunks, N)
template <typename T, bool FLIP, bool STATS>
__global__ void __launch_bounds__(256, 2)
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
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);     // planes outside [0, T) are zero
    uint4 ring[FRING];
#pragma unroll
    for (int i = 0; i < FRING; ++i) {
      const int pl = i - KT / 2;
      ring[i] = (pl >= 0 && pl < T_) ? ldg16(xp + (int64_t)pl * plane) : zero;
    }
    for (int t0 = 0; t0 < T_; t0 += FRING) {
#pragma unroll
      for (int s = 0; s < FRING; ++s) {
        const int t = t0 + s;
        if (t < T_) {
          float o[VEC];
#pragma unroll
          for (int j = 0; j < VEC; ++j) o[j] = 0.f;
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            float v[VEC];
            unpack16<T>(ring[(s + k) % FRING], v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) o[j] = fmaf(wk[k][j], v[j], o[j]);
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
          const int pl = t - KT / 2 + FRING;            // plane t-2 retires, its slot takes plane t-2+FRING
          ring[s] = pl < T_ ? ldg16(xp + (int64_t)pl * plane) : zero;
        }
      }
    }
  }
  if (STATS) block_stats_flush<VEC>(s1, s2, cvec, Cp, s_acc, stats + (int64_t)n * Cp * 2);
}

// dw[c][k] += sum_{n,t,p} dy[n,t,p,c] * x[n, t + k - 2, p, c];  persistent blocks over (n, position chunk) units
template <typename T>
__global__ void __launch_bounds__(256, 2)
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
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int n = (int)(u / chunks);
    const int64_t p = (u - (int64_t)n * chunks) * rows + prow;
    if (prow >= rows || p >= P) continue;
    const T* xp = x + ((int64_t)n * T_ * P + p) * Cp + c0;
    const T* dp = dy + ((int64_t)n * T_ * P + p) * Cp + c0;
    uint4 ring[RING], dring[LOOK];                 // x planes t-2 .. t+2+LOOK; dy planes t .. t+LOOK-1 (slot t % LOOK)
#pragma unroll
    for (int i = 0; i < RING; ++i) {
      const int pl = i - KT / 2;
      ring[i] = (pl >= 0 && pl < T_) ? ldg16(xp + (int64_t)pl * plane) : zero;
    }
#pragma unroll
    for (int i = 0; i < LOOK; ++i) dring[i] = i < T_ ? ldg16(dp + (int64_t)i * plane) : zero;
    // RING * LOOK is the common period of the two rings; the march is unrolled over it so all slot indices are static
    for (int t0 = 0; t0 < T_; t0 += RING * LOOK) {
#pragma unroll
      for (int s = 0; s < RING * LOOK; ++s) {
        const int t = t0 + s;
        if (t < T_) {
          float d[VEC];
          unpack16<T>(dring[s % LOOK], d);
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            float v[VEC];
            unpack16<T>(ring[(s + k) % RING], v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) g[k][j] = fmaf(d[j], v[j], g[k][j]);
          }
          const int pl = t - KT / 2 + RING;
          ring[s % RING] = pl < T_ ? ldg16(xp + (int64_t)pl * plane) : zero;
          dring[s % LOOK] = t + LOOK < T_ ? ldg16(dp + (int64_t)(t + LOOK) * plane) : zero;
        }
      }
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
