// Shared device/host helpers for libx3d_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/x3d_b200.h"

#ifndef __CUDA_ARCH_LIST__
#endif

namespace x3d {

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void count_path(int path);      // x3d_path_t
// Per-DEVICE one-time set-up (cudaFuncSetAttribute is a per-device attribute): returns true the first time the
// calling site is reached with the current device; `mask` is one static bitmask per call site.
bool first_use_on_device(unsigned long long* mask);

#define X3D_CHECK_ARG(cond, msg)                    \
  do {                                              \
    if (!(cond)) {                                  \
      x3d::set_error("%s: %s", __func__, msg);      \
      return -1;                                    \
    }                                               \
  } while (0)

#define X3D_LAUNCH_CHECK()                                                   \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      x3d::set_error("%s: launch failed: %s", __func__, cudaGetErrorString(e__)); \
      return (int)e__;                                                       \
    }                                                                        \
    x3d::count_launch();                                                     \
  } while (0)

static inline cudaStream_t as_stream(x3d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch.  A training step is ~700 dependent kernels, most of them 3-30 us long: every
// kernel lets its successor become resident right away (pdl_trigger) and orders itself behind its predecessor
// with pdl_wait() before it touches global memory, so launch latency and the block-scheduling ramp overlap the
// predecessor's tail.  (griddepcontrol.wait returns once all prerequisite grids have completed and flushed; both
// instructions are no-ops for a kernel launched without the attribute.)  X3D_NO_PDL=1 launches the classic way.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_trigger();
  pdl_wait();
}
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through cudaGetLastError
}
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;

// ---------------------------------------------------------------------------------------
// 16-byte vector access: VEC channels per thread (8 x bf16 or 4 x fp32)
// ---------------------------------------------------------------------------------------
template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int N = 4;
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
};

template <typename T>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&v)[Vec<T>::N]);
template <>
__device__ __forceinline__ void load_vec<float>(const float* __restrict__ p, float (&v)[4]) {
  float4 q = *reinterpret_cast<const float4*>(p);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16>(const __nv_bfloat16* __restrict__ p, float (&v)[8]) {
  uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <typename T>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&v)[Vec<T>::N]);
template <>
__device__ __forceinline__ void store_vec<float>(float* __restrict__ p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16>(__nv_bfloat16* __restrict__ p, const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = q;
}

// VEC (4 or 8) consecutive per-channel constants (scale / shift / coefficients), 16-byte aligned: vector loads --
// the element-wise kernels read up to six such sets per thread before their loop (48 scalar loads used to be a large
// part of their run time on the small late-stage tensors)
template <int VEC>
__device__ __forceinline__ void load_consts(const float* __restrict__ p, float (&v)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; i += 4) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p + i));
    v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
  }
}

// value as it will be re-read from storage (so statistics/masks agree with what is stored)
template <typename T>
__device__ __forceinline__ float round_to(float x);
template <>
__device__ __forceinline__ float round_to<float>(float x) { return x; }
template <>
__device__ __forceinline__ float round_to<__nv_bfloat16>(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

template <typename T>
__device__ __forceinline__ float to_float(T x);
template <>
__device__ __forceinline__ float to_float<float>(float x) { return x; }
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T>
__device__ __forceinline__ T from_float(float x);
template <>
__device__ __forceinline__ float from_float<float>(float x) { return x; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// dispatch on the activation dtype
#define X3D_DISPATCH_DTYPE(dt, ...)                         \
  do {                                                      \
    if ((dt) == X3D_F32) {                                  \
      using T = float;                                      \
      __VA_ARGS__;                                          \
    } else if ((dt) == X3D_BF16) {                          \
      using T = __nv_bfloat16;                              \
      __VA_ARGS__;                                          \
    } else {                                                \
      x3d::set_error("%s: bad dtype %d", __func__, (int)(dt)); \
      return -1;                                            \
    }                                                       \
  } while (0)

// ---------------------------------------------------------------------------------------
// Geometry of "per-sample channel-vector" kernels: block = CV x R threads where
// CV = Cp/VEC channel vectors and R position rows; grid = (chunks, N).
// ---------------------------------------------------------------------------------------
struct RowGeom {
  int cv;        // channel vectors per position
  int rows;      // position rows per block iteration
  int threads;   // cv * rows
  int chunks;    // blocks per sample
  int64_t chunk; // positions per block
};

template <typename T>
static inline RowGeom make_row_geom(int64_t N, int64_t P, int64_t Cp, int target_blocks = 4 * kNumSMs) {
  RowGeom g;
  g.cv = (int)(Cp / Vec<T>::N);
  g.rows = g.cv >= 256 ? 1 : 256 / g.cv;
  if ((int64_t)g.rows > P) g.rows = (int)P;
  g.threads = g.cv * g.rows;
  int64_t want = cdiv(target_blocks, N);               // chunks per sample we would like
  int64_t min_chunk = (int64_t)g.rows * 4;             // at least 4 iterations per block
  int64_t chunk = cdiv(P, want);
  if (chunk < min_chunk) chunk = min_chunk;
  chunk = cdiv(chunk, g.rows) * g.rows;
  g.chunk = chunk;
  g.chunks = (int)cdiv(P, chunk);
  return g;
}

// Block-level reduction of per-thread channel accumulators into double stats[n][Cp][2].
// acc0/acc1: VEC partial sums owned by this thread for channels [cvec*VEC, +VEC).  Block = cv x rows threads
// (thread = cvec + cv*prow).  Every thread parks its partials in its own shared slot (no atomics), then one
// thread per (channel, quantity) sums the `rows` slots and issues ONE fp64 atomic.  s_acc: rows*Cp*2 floats.
template <int VEC>
__device__ __forceinline__ void block_stats_flush(float (&acc0)[VEC], float (&acc1)[VEC], int cvec, int Cp,
                                                  float* s_acc, double* stats_n /* [Cp][2] */) {
  const int cv = Cp / VEC;
  const int prow = threadIdx.x / cv, rows = blockDim.x / cv;
  float* mine = s_acc + (size_t)prow * Cp * 2;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    mine[(cvec * VEC + j) * 2 + 0] = acc0[j];
    mine[(cvec * VEC + j) * 2 + 1] = acc1[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cp * 2; i += blockDim.x) {
    float v = 0.f;
    for (int r = 0; r < rows; ++r) v += s_acc[(size_t)r * Cp * 2 + i];
    if (v != 0.f) atomicAdd(&stats_n[i], (double)v);
  }
}

}  // namespace x3d
