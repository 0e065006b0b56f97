// Tiled depthwise 3x3x3 forward for the hot shapes (stride (1,1,1) and (1,2,2), pad 1), NDHWC.
//
// Design (see DESIGN.md "dwconv"):
//  * CTA = (sample n, TH x TW output tile, chunk of CC channels), marching over ALL T planes.
//    Each input plane tile (with its 1-pixel halo) is brought ONCE into shared memory with 16-byte
//    cp.async (zero-filled outside the image), double-buffered against the FMA phase.
//  * thread = (channel PAIR, 2 x 4 output patch).  The 27 taps of the pair live in registers as float2;
//    every shared-memory word (2 channels) feeds up to 27 packed FFMA2 (fma.rn.f32x2): one input plane
//    contributes to three output planes held in rolling register accumulators, so an input element is
//    read from shared memory once per thread-window, never from global more than once per CTA.
//  * the preceding SubBatchNorm3d+ReLU (scale/shift per (split,channel)) is applied on the fly to the
//    window values; padding positions are forced to exact zero AFTER the transform (x3d.py:147-150).
//  * epilogue: bf16/fp32 store + per-(sample,channel) sum / sum-of-squares of the stored values
//    (bn2 statistics and the SE global pool) -> shared atomics -> one fp64 atomic per channel per CTA.
// Arithmetic intensity at stride 1 in bf16 is 27 FMA / 4 B = 6.75 FMA/B, above the B200 balance of
// ~5.7 FMA/B (37 TFMA/s fp32 vs 6.5 TB/s): the stride-1 layers are bound by the fp32 FMA pipe, which is
// why the inner loop is FFMA2 and everything else is kept off that pipe.
#include "common.cuh"

using namespace x3d;

namespace {

constexpr int PH = 2, PW = 4;     // output patch per thread

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <typename T>
__device__ __forceinline__ float2 lds_pair(const T* p);
template <>
__device__ __forceinline__ float2 lds_pair<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <>
__device__ __forceinline__ float2 lds_pair<float>(const float* p) {
  return *reinterpret_cast<const float2*>(p);
}
// store a channel pair; returns the values as stored (for the statistics)
template <typename T>
__device__ __forceinline__ float2 st_pair(T* p, float2 v);
template <>
__device__ __forceinline__ float2 st_pair<__nv_bfloat16>(__nv_bfloat16* p, float2 v) {
  const uint32_t w = pack_bf16x2(v.x, v.y);
  *reinterpret_cast<uint32_t*>(p) = w;
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <>
__device__ __forceinline__ float2 st_pair<float>(float* p, float2 v) {
  *reinterpret_cast<float2*>(p) = v;
  return v;
}

struct TileGeom {
  int T, H, W, Ho, Wo, Cp;
  int TH, TW, CC;        // output tile, channels per CTA
  int tiles_w;
};

template <typename T, int S, bool XFORM, bool RELU, bool STATS>
__global__ void __launch_bounds__(384, 1)
dw3_fwd_tiled_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y, TileGeom g,
                     const float* __restrict__ scale, const float* __restrict__ shift, int splits,
                     double* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int VEC = Vec<T>::N;
  constexpr int WR = (PH - 1) * S + 3, WC = (PW - 1) * S + 3;   // per-thread input window
  const int IH = (g.TH - 1) * S + 3, IW = (g.TW - 1) * S + 3;    // CTA input tile (with halo)
  const int CC = g.CC, Cp = g.Cp;
  const int plane_elems = IH * IW * CC;
  T* buf[2] = {reinterpret_cast<T*>(smem_raw), reinterpret_cast<T*>(smem_raw) + plane_elems};
  float* s_stat = reinterpret_cast<float*>(reinterpret_cast<T*>(smem_raw) + 2 * plane_elems);

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int ho0 = (blockIdx.x / g.tiles_w) * g.TH, wo0 = (blockIdx.x % g.tiles_w) * g.TW;
  const int cbase = blockIdx.y * CC;
  const int n = blockIdx.z;
  const int pairs = CC / 2;
  const int pair = tid % pairs, patch = tid / pairs;
  const int ppr = g.TW / PW;                         // patches per tile row
  const int py = patch / ppr, px = patch % ppr;
  const int c = cbase + 2 * pair;
  const bool ch_ok = c < Cp;

  // ---- loader plan (shared table, built once per CTA): global element offset of every 16-byte
  //      vector of the input tile inside its plane, or -1 for zero fill (outside the image)
  const int vpp = CC / VEC;                           // vectors per position
  const int nvec = IH * IW * vpp;
  int* s_goff = reinterpret_cast<int*>(s_stat + 2 * CC);
  const int hi0 = ho0 * S - 1, wi0 = wo0 * S - 1;
  for (int v = tid; v < nvec; v += nthr) {
    const int pos = v / vpp, cv = v % vpp;
    const int r = pos / IW, cc = pos % IW;
    const int gh = hi0 + r, gw = wi0 + cc;
    const bool ok = gh >= 0 && gh < g.H && gw >= 0 && gw < g.W && (cbase + cv * VEC) < Cp;
    s_goff[v] = ok ? (gh * g.W + gw) * Cp + cbase + cv * VEC : -1;
  }
  __syncthreads();
  const int64_t plane_stride = (int64_t)g.H * g.W * Cp;
  const T* xn = x + (int64_t)n * g.T * plane_stride;
  auto issue = [&](int t, T* dst) {
    const T* xp = xn + (int64_t)t * plane_stride;
    for (int v = tid; v < nvec; v += nthr) {
      const int go = s_goff[v];
      cp_async16(dst + (int64_t)v * VEC, go >= 0 ? (const void*)(xp + go) : (const void*)xn, go >= 0 ? 16 : 0);
    }
    cp_async_commit();
  };

  // ---- per-thread constants ----------------------------------------------------------------
  float2 wreg[27];
#pragma unroll
  for (int tap = 0; tap < 27; ++tap)
    wreg[tap] = ch_ok ? *reinterpret_cast<const float2*>(w + (int64_t)tap * Cp + c) : make_float2(0.f, 0.f);
  float2 sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f);
  if (XFORM && ch_ok) {
    const int b = n % splits;
    sc = *reinterpret_cast<const float2*>(scale + (int64_t)b * Cp + c);
    sh = *reinterpret_cast<const float2*>(shift + (int64_t)b * Cp + c);
  }
  // validity of the window rows / columns (zero padding is applied after the BN+ReLU transform)
  uint32_t rmask = 0, cmask = 0;
#pragma unroll
  for (int r = 0; r < WR; ++r) {
    const int gh = hi0 + py * PH * S + r;
    rmask |= (gh >= 0 && gh < g.H) ? (1u << r) : 0u;
  }
#pragma unroll
  for (int cc = 0; cc < WC; ++cc) {
    const int gw = wi0 + px * PW * S + cc;
    cmask |= (gw >= 0 && gw < g.W) ? (1u << cc) : 0u;
  }
  const bool interior = rmask == ((1u << WR) - 1) && cmask == ((1u << WC) - 1);
  const int win_base = ((py * PH * S) * IW + px * PW * S) * CC + 2 * pair;
  const int row_stride = IW * CC;

  float2 acc[3][PH * PW];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int o = 0; o < PH * PW; ++o) acc[k][o] = make_float2(0.f, 0.f);
  float2 ssum = make_float2(0.f, 0.f), ssq = make_float2(0.f, 0.f);

  const int ho_t = ho0 + py * PH, wo_t = wo0 + px * PW;
  auto store_plane = [&](int tout) {
    T* yp = y + ((((int64_t)n * g.T + tout) * g.Ho + ho_t) * g.Wo + wo_t) * Cp + c;
#pragma unroll
    for (int oy = 0; oy < PH; ++oy) {
#pragma unroll
      for (int ox = 0; ox < PW; ++ox) {
        if (ch_ok && ho_t + oy < g.Ho && wo_t + ox < g.Wo) {
          const float2 r = st_pair<T>(yp + ((int64_t)oy * g.Wo + ox) * Cp, acc[2][oy * PW + ox]);
          if (STATS) {
            ssum.x += r.x; ssum.y += r.y;
            ssq = __ffma2_rn(r, r, ssq);
          }
        }
      }
    }
  };

  issue(0, buf[0]);
  for (int tin = 0; tin < g.T; ++tin) {
    cp_async_wait_all();
    __syncthreads();                       // plane tin landed; everybody finished reading the other buffer
    if (tin + 1 < g.T) issue(tin + 1, buf[(tin + 1) & 1]);
    const T* bp = buf[tin & 1] + win_base;
#pragma unroll
    for (int r = 0; r < WR; ++r) {
#pragma unroll
      for (int cc = 0; cc < WC; ++cc) {
        float2 xv = lds_pair<T>(bp + r * row_stride + cc * CC);
        if (XFORM) {
          xv = __ffma2_rn(xv, sc, sh);
          if (RELU) { xv.x = fmaxf(xv.x, 0.f); xv.y = fmaxf(xv.y, 0.f); }
          if (!interior && !(((rmask >> r) & (cmask >> cc)) & 1u)) xv = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int oy = 0; oy < PH; ++oy) {
          const int kh = r - oy * S;
          if (kh < 0 || kh > 2) continue;
#pragma unroll
          for (int ox = 0; ox < PW; ++ox) {
            const int kw = cc - ox * S;
            if (kw < 0 || kw > 2) continue;
#pragma unroll
            for (int kt = 0; kt < 3; ++kt)
              acc[kt][oy * PW + ox] = __ffma2_rn(wreg[(kt * 3 + kh) * 3 + kw], xv, acc[kt][oy * PW + ox]);
          }
        }
      }
    }
    // acc[kt] holds output plane tin + 1 - kt;  plane tin-1 (kt = 2) is now complete
    if (tin >= 1) store_plane(tin - 1);
#pragma unroll
    for (int o = 0; o < PH * PW; ++o) {
      acc[2][o] = acc[1][o];
      acc[1][o] = acc[0][o];
      acc[0][o] = make_float2(0.f, 0.f);
    }
  }
  store_plane(g.T - 1);

  if (STATS) {
    __syncthreads();
    for (int i = tid; i < CC * 2; i += nthr) s_stat[i] = 0.f;
    __syncthreads();
    atomicAdd(&s_stat[(2 * pair) * 2 + 0], ssum.x);
    atomicAdd(&s_stat[(2 * pair) * 2 + 1], ssq.x);
    atomicAdd(&s_stat[(2 * pair + 1) * 2 + 0], ssum.y);
    atomicAdd(&s_stat[(2 * pair + 1) * 2 + 1], ssq.y);
    __syncthreads();
    for (int i = tid; i < CC * 2; i += nthr) {
      const int ch = cbase + i / 2;
      const float v = s_stat[i];
      if (ch < Cp && v != 0.f) atomicAdd(&stats[((int64_t)n * Cp + ch) * 2 + (i & 1)], (double)v);
    }
  }
}

struct TilePlan {
  TileGeom g;
  dim3 grid;
  int threads;
  size_t smem;
  bool ok;
};

template <typename T>
TilePlan plan_tiles(int64_t N, int T_, int H, int W, int Cp, int S) {
  TilePlan p;
  p.ok = false;
  TileGeom& g = p.g;
  g.T = T_; g.H = H; g.W = W; g.Cp = Cp;
  g.Ho = (H + 2 - 3) / S + 1;
  g.Wo = (W + 2 - 3) / S + 1;
  const int esz = (int)sizeof(T);
  // tile: width up to 16 (stride 1) / 8 (stride 2) outputs, height up to 8
  int TW = (g.Wo + PW - 1) / PW * PW;
  const int maxTW = (S == 1) ? 16 : 8;
  if (TW > maxTW) {
    // split the row into equal tiles of at most maxTW (multiple of PW) to limit waste
    int nt = (g.Wo + maxTW - 1) / maxTW;
    TW = ((g.Wo + nt - 1) / nt + PW - 1) / PW * PW;
  }
  int TH = (g.Ho + PH - 1) / PH * PH;
  if (TH > 8) {
    int nt = (g.Ho + 7) / 8;
    TH = ((g.Ho + nt - 1) / nt + PH - 1) / PH * PH;
  }
  const int patches = (TH / PH) * (TW / PW);
  // channel chunk: as many pairs as fit in <= 384 threads (168 registers per thread), multiple of 8 channels
  int pairs = 384 / patches;
  if (pairs > Cp / 2) pairs = Cp / 2;
  int CC = pairs * 2 / 8 * 8;
  if (CC < 8) return p;
  // even out the chunks (e.g. Cp=216 -> 3 x 72 instead of 208 + 8)
  const int nchunk = (Cp + CC - 1) / CC;
  CC = ((Cp + nchunk - 1) / nchunk + 7) / 8 * 8;
  g.TH = TH; g.TW = TW; g.CC = CC;
  g.tiles_w = (g.Wo + TW - 1) / TW;
  const int tiles_h = (g.Ho + TH - 1) / TH;
  const int IH = (TH - 1) * S + 3, IW = (TW - 1) * S + 3;
  p.threads = patches * (CC / 2);
  const int nvec = IH * IW * (CC / Vec<T>::N);
  if (p.threads > 384 || p.threads < 32) return p;
  p.smem = (size_t)2 * IH * IW * CC * esz + (size_t)CC * 2 * sizeof(float) + (size_t)nvec * sizeof(int);
  if (p.smem > 200 * 1024) return p;
  p.grid = dim3((unsigned)(g.tiles_w * tiles_h), (unsigned)((Cp + CC - 1) / CC), (unsigned)N);
  if (N > 65535) return p;
  p.ok = true;
  return p;
}

template <typename T, int S>
int launch_tiled(const TilePlan& p, const void* x, const float* w, void* y, const float* scale, const float* shift,
                 int splits, int relu_in, double* stats, cudaStream_t stream) {
#define L_(XF, RL, ST)                                                                                         \
  do {                                                                                                         \
    auto kfn = dw3_fwd_tiled_kernel<T, S, XF, RL, ST>;                                                         \
    if (p.smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem); \
    kfn<<<p.grid, p.threads, p.smem, stream>>>((const T*)x, w, (T*)y, p.g, scale, shift, splits, stats);       \
  } while (0)
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

}  // namespace

namespace x3d {
int dwconv_fwd_tiled(const void* x, const float* w_packed, void* y, int64_t N, int64_t T_, int64_t H, int64_t W,
                     int64_t Cp, int stride, const float* in_scale, const float* in_shift, int splits, int relu_in,
                     double* stats, x3d_dtype_t dt, cudaStream_t stream, bool* handled) {
  *handled = false;
  static const bool force_direct = getenv("X3D_DW_DIRECT") != nullptr;
  if (force_direct) return 0;
  if (dt == X3D_BF16) {
    using T = __nv_bfloat16;
    TilePlan p = plan_tiles<T>(N, (int)T_, (int)H, (int)W, (int)Cp, stride);
    if (!p.ok) return 0;
    if (stride == 1) launch_tiled<T, 1>(p, x, w_packed, y, in_scale, in_shift, splits, relu_in, stats, stream);
    else launch_tiled<T, 2>(p, x, w_packed, y, in_scale, in_shift, splits, relu_in, stats, stream);
  } else {
    using T = float;
    TilePlan p = plan_tiles<T>(N, (int)T_, (int)H, (int)W, (int)Cp, stride);
    if (!p.ok) return 0;
    if (stride == 1) launch_tiled<T, 1>(p, x, w_packed, y, in_scale, in_shift, splits, relu_in, stats, stream);
    else launch_tiled<T, 2>(p, x, w_packed, y, in_scale, in_shift, splits, relu_in, stats, stream);
  }
  *handled = true;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("dwconv_fwd_tiled: launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  count_launch();
  return 0;
}
}  // namespace x3d
