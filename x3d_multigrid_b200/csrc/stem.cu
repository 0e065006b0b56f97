// Stem conv1_s: dense (1,3,3) conv, stride (1,2,2), pad (0,1,1), Ci (<=3) -> Co channels.
// Reads the user-facing NCDHW fp32 clip directly, writes NDHWC.  Replaces the cuDNN kernel behind
// nn.Conv3d at x3d.py:196-201 / :317.   TAPS = Ci*9 <= 27.
#include <stdlib.h>

#include "common.cuh"

using namespace x3d;

constexpr int STEM_MAX_TAPS = 27;

// Where the stem reads the clip from.
//  * SrcF32: the user-facing fp32 NCDHW clip (x3d.py:316).
//  * SrcU8:  decoded uint8 frames [B][T][Hs][Ws][3] (NTHWC) + per-clip crop window / flip: the crop, the horizontal
//            flip, ToTensor(255) and Normalize(mean, std) of the reference's input pipeline
//            (transforms/spatial_transforms.py:35-119,331-349,472-501; train_x3d_kinetics_multigrid.py:70-73) are
//            applied on the fly with the same fp32 operations ((v / 255 - mean) / std, each rounded), so the stem sees
//            bit-identical values -- no fp32 clip is ever materialised (4x less H2D and 2 x 154 MB less HBM traffic).
struct SrcF32 {
  const float* x;
  int Ci, T, H, W;
  __device__ __forceinline__ float operator()(int64_t n, int ci, int t, int hh, int ww) const {
    return __ldg(&x[((((int64_t)n * Ci + ci) * T + t) * H + hh) * W + ww]);
  }
};
struct SrcU8 {
  const unsigned char* src;
  const x3d_crop_t* crops;
  int T, Hs, Ws, S;
  float norm, mean[3], stdv[3];
  __device__ __forceinline__ float operator()(int64_t n, int ci, int t, int hh, int ww) const {
    const x3d_crop_t c = crops[n];
    const int xs = c.flip ? c.x1 + S - 1 - ww : c.x1 + ww;
    const unsigned char v = __ldg(&src[((((int64_t)n * T + t) * Hs + (c.y1 + hh)) * Ws + xs) * 3 + ci]);
    return __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, norm), mean[ci]), stdv[ci]);
  }
};

// Thread = PQ = 4 consecutive output ROWS at one output column x all output channels; lanes run along the output
// columns (stride-2 coalesced input reads).  The weights are read from shared memory (warp-uniform 16-byte loads); one
// output per thread made that the bound: a broadcast LDS.128 still delivers 512 B to the register file (4 cycles of the
// SM's one load pipe) for only 8 FMAs of the warp -- 175 us for 308 MB.  With four positions per thread each weight
// vector feeds 32 FMAs, and the 9 x 3 x 3 input window is shared by the four rows (81 loads instead of 108).
constexpr int STEM_PQ = 4;
template <typename T, typename SRC>
__global__ void __launch_bounds__(128) stem_fwd_kernel(const SRC x, const float* __restrict__ w,
                                                        T* __restrict__ y, int Ci, int T_, int H, int W, int Ho,
                                                        int Wo, int Co, int Cop, int HQ, int64_t total) {
  x3d::pdl_prologue();
  extern __shared__ float s_w[];  // [taps][Cop]
  const int taps = Ci * 9;
  for (int i = threadIdx.x; i < taps * Cop; i += blockDim.x) {
    const int tap = i / Cop, c = i % Cop;
    s_w[i] = (c < Co) ? w[(int64_t)c * taps + tap] : 0.f;  // w[co][ci][0][j][k] -> tap = ci*9 + j*3 + k
  }
  __syncthreads();
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= total) return;
  const int wo = (int)(u % Wo);
  int64_t r = u / Wo;
  const int hq = (int)(r % HQ);
  r /= HQ;
  const int t = (int)(r % T_);
  const int n = (int)(r / T_);
  const int ho0 = hq * STEM_PQ;
  constexpr int XR = 2 * STEM_PQ + 1;          // input rows of the thread
  float xin[3][XR][3];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
    for (int j = 0; j < XR; ++j) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int hh = 2 * ho0 + j - 1, ww = 2 * wo + k - 1;
        float v = 0.f;
        if (ci < Ci && hh >= 0 && hh < H && ww >= 0 && ww < W) v = x(n, ci, t, hh, ww);
        xin[ci][j][k] = v;
      }
    }
  }
  const int64_t orow = (int64_t)Wo * Cop;
  T* yp = y + ((((int64_t)n * T_ + t) * Ho + ho0) * Wo + wo) * Cop;
  for (int c8 = 0; c8 < Cop; c8 += 8) {
    float acc[STEM_PQ][8];
#pragma unroll
    for (int q = 0; q < STEM_PQ; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      if (ci < Ci) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int tap = ci * 9 + j * 3 + k;
            const float4 w0 = *reinterpret_cast<const float4*>(&s_w[tap * Cop + c8]);
            const float4 w1 = *reinterpret_cast<const float4*>(&s_w[tap * Cop + c8 + 4]);
#pragma unroll
            for (int q = 0; q < STEM_PQ; ++q) {
              const float xv = xin[ci][2 * q + j][k];
              acc[q][0] = fmaf(xv, w0.x, acc[q][0]); acc[q][1] = fmaf(xv, w0.y, acc[q][1]);
              acc[q][2] = fmaf(xv, w0.z, acc[q][2]); acc[q][3] = fmaf(xv, w0.w, acc[q][3]);
              acc[q][4] = fmaf(xv, w1.x, acc[q][4]); acc[q][5] = fmaf(xv, w1.y, acc[q][5]);
              acc[q][6] = fmaf(xv, w1.z, acc[q][6]); acc[q][7] = fmaf(xv, w1.w, acc[q][7]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < STEM_PQ; ++q) {
      if (ho0 + q >= Ho) continue;
      if (sizeof(T) == 2) {
        uint4 o;
        o.x = pack_bf16x2(acc[q][0], acc[q][1]); o.y = pack_bf16x2(acc[q][2], acc[q][3]);
        o.z = pack_bf16x2(acc[q][4], acc[q][5]); o.w = pack_bf16x2(acc[q][6], acc[q][7]);
        *reinterpret_cast<uint4*>(yp + q * orow + c8) = o;
      } else {
        float* yf = reinterpret_cast<float*>(yp + q * orow) + c8;
        *reinterpret_cast<float4*>(yf) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
        *reinterpret_cast<float4*>(yf + 4) = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
      }
    }
  }
}

extern "C" int x3d_stem_conv_s_fwd(const float* x, const float* w, void* y, int64_t N, int64_t Ci, int64_t T_,
                                   int64_t H, int64_t W, int64_t Co, int64_t Cop, x3d_dtype_t dt,
                                   x3d_stream_t stream) {
  X3D_CHECK_ARG(Ci >= 1 && Ci <= 3, "n_input_channels must be <= 3");
  X3D_CHECK_ARG(Cop % 8 == 0 && Cop >= Co, "Cop");
  const int Ho = (int)((H + 2 - 3) / 2 + 1), Wo = (int)((W + 2 - 3) / 2 + 1);
  const int HQ = (Ho + STEM_PQ - 1) / STEM_PQ;
  const int64_t total = N * T_ * HQ * Wo;
  if (total == 0) return 0;
  size_t smem = (size_t)Ci * 9 * Cop * sizeof(float);
  const SrcF32 srcx{x, (int)Ci, (int)T_, (int)H, (int)W};
  X3D_DISPATCH_DTYPE(dt, (x3d::launch(stem_fwd_kernel<T, SrcF32>, (unsigned)cdiv(total, 128), 128, smem, as_stream(stream),
                             srcx, w, (T*)y, (int)Ci, (int)T_, (int)H, (int)W, Ho, Wo, (int)Co, (int)Cop, HQ, total)));
  X3D_LAUNCH_CHECK();
  return 0;
}

static SrcU8 make_src_u8(const unsigned char* src, const x3d_crop_t* crops, int64_t T_, int64_t Hs, int64_t Ws, int64_t S,
                         const float* mean_std, float norm_value) {
  SrcU8 u;
  u.src = src; u.crops = crops; u.T = (int)T_; u.Hs = (int)Hs; u.Ws = (int)Ws; u.S = (int)S; u.norm = norm_value;
  for (int i = 0; i < 3; ++i) { u.mean[i] = mean_std[i]; u.stdv[i] = mean_std[3 + i]; }
  return u;
}

extern "C" int x3d_stem_conv_s_fwd_u8(const uint8_t* src, const x3d_crop_t* crops_dev, const float* w, void* y, int64_t N,
                                      int64_t T_, int64_t Hs, int64_t Ws, int64_t S, const float* mean_std,
                                      float norm_value, int64_t Co, int64_t Cop, x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cop % 8 == 0 && Cop >= Co, "Cop");
  X3D_CHECK_ARG(S >= 1 && S <= Hs && S <= Ws && mean_std != nullptr && crops_dev != nullptr, "crop larger than the frames");
  const int Ho = (int)((S + 2 - 3) / 2 + 1);
  const int HQ = (Ho + STEM_PQ - 1) / STEM_PQ;
  const int64_t total = N * T_ * HQ * Ho;
  if (total == 0) return 0;
  size_t smem = (size_t)27 * Cop * sizeof(float);
  const SrcU8 srcx = make_src_u8(src, crops_dev, T_, Hs, Ws, S, mean_std, norm_value);
  X3D_DISPATCH_DTYPE(dt, (x3d::launch(stem_fwd_kernel<T, SrcU8>, (unsigned)cdiv(total, 128), 128, smem, as_stream(stream),
                             srcx, w, (T*)y, 3, (int)T_, (int)S, (int)S, Ho, Ho, (int)Co, (int)Cop, HQ, total)));
  X3D_LAUNCH_CHECK();
  return 0;
}

// uint8 frames -> the fp32 NCDHW clip the reference's transforms produce (crop, flip, ToTensor(255), Normalize).
// The 3 x 256 possible results are tabulated per CTA with the reference's exact fp32 operations; a thread then turns 4
// consecutive pixels (12 source bytes) into one float4 per channel.
__global__ void __launch_bounds__(256) clip_u8_to_f32_kernel(const SrcU8 x, float* __restrict__ dst, int T_, int S,
                                                             int64_t quads) {
  x3d::pdl_prologue();
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int ci = i >> 8, v = i & 255;
    lut[ci][v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, x.norm), x.mean[ci]), x.stdv[ci]);
  }
  __syncthreads();
  const int qpr = (S + 3) >> 2;                                  // quads per output row
  const int64_t plane = (int64_t)T_ * S * S;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
    const int wq = (int)(q % qpr);
    int64_t r = q / qpr;
    const int hh = (int)(r % S);
    r /= S;
    const int t = (int)(r % T_);
    const int64_t n = r / T_;
    const x3d_crop_t c = x.crops[n];
    const int ww0 = wq * 4;
    const unsigned char* row = x.src + (((int64_t)n * x.T + t) * x.Hs + (c.y1 + hh)) * x.Ws * 3;
    float o[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ww = ww0 + e < S ? ww0 + e : S - 1;
      const int xs = c.flip ? c.x1 + S - 1 - ww : c.x1 + ww;
      const unsigned char* px = row + xs * 3;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) o[ci][e] = lut[ci][__ldg(px + ci)];
    }
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      float* d = dst + (n * 3 + ci) * plane + ((int64_t)t * S + hh) * S + ww0;
      if (ww0 + 3 < S && (S & 3) == 0) {
        *reinterpret_cast<float4*>(d) = make_float4(o[ci][0], o[ci][1], o[ci][2], o[ci][3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (ww0 + e < S) d[e] = o[ci][e];
      }
    }
  }
}
extern "C" int x3d_clip_u8_to_f32(const uint8_t* src, const x3d_crop_t* crops_dev, float* dst, int64_t N, int64_t T_,
                                  int64_t Hs, int64_t Ws, int64_t S, const float* mean_std, float norm_value,
                                  x3d_stream_t stream) {
  X3D_CHECK_ARG(S >= 1 && S <= Hs && S <= Ws && mean_std != nullptr && crops_dev != nullptr, "crop larger than the frames");
  const int64_t quads = N * T_ * S * ((S + 3) / 4);
  if (quads == 0) return 0;
  const SrcU8 srcx = make_src_u8(src, crops_dev, T_, Hs, Ws, S, mean_std, norm_value);
  int64_t blocks = cdiv(quads, 256);
  if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
  x3d::launch(clip_u8_to_f32_kernel, (unsigned)blocks, 256, 0, as_stream(stream), srcx, dst, (int)T_, (int)S, quads);
  X3D_LAUNCH_CHECK();
  return 0;
}

// wgrad: dw[co][tap] += sum_p dy[p][co] * xcol[p][tap].  Block: 256 threads = 32 taps x 8 channel groups.
constexpr int SW_POS = 128;
constexpr long long kNoPos = (long long)0x8000000000000000ull;   // sentinel: position beyond the tensor
template <typename T>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy,
                                                          float* __restrict__ dw, int Ci, int T_, int H, int W,
                                                          int Ho, int Wo, int Co, int Cop, int64_t total,
                                                          int64_t pos_per_block) {
  x3d::pdl_prologue();
  extern __shared__ float sm[];
  float* Xs = sm;                   // [SW_POS][32]
  float* Ds = sm + SW_POS * 32;     // [SW_POS][Cop]
  long long* s_base = reinterpret_cast<long long*>(Ds + SW_POS * Cop);   // [SW_POS]
  int* s_flag = reinterpret_cast<int*>(s_base + SW_POS);                   // [SW_POS]
  const int taps = Ci * 9;
  // register tile: a thread owns 3 taps x 8 channels and every `nslice`-th position of the tile
  const int ncg = Cop / 8;                      // channel groups
  const int roles = 9 * ncg;                    // (tap group, channel group) pairs
  const int nslice = 256 / roles;
  const int role = threadIdx.x % roles, slice = threadIdx.x / roles;
  const int tg = role % 9, cg = role / 9;
  const bool active = slice < nslice && tg * 3 < taps;
  float acc[3][8];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int64_t pb = (int64_t)blockIdx.x * pos_per_block;
  const int64_t pe = (pb + pos_per_block < total) ? pb + pos_per_block : total;
  for (int64_t ps = pb; ps < pe; ps += SW_POS) {
    // per-position decode once (n, t, ho, wo -> offset of tap (ci=0, j=0, k=0) and border flags)
    if (threadIdx.x < SW_POS) {
      const int64_t p = ps + threadIdx.x;
      long long base = kNoPos;
      int flags = 0;
      if (p < pe) {
        const int wo = (int)(p % Wo);
        int64_t r = p / Wo;
        const int ho = (int)(r % Ho);
        r /= Ho;
        const int t = (int)(r % T_);
        const int n = (int)(r / T_);
        base = ((((int64_t)n * Ci) * T_ + t) * H + (2 * ho - 1)) * (int64_t)W + (2 * wo - 1);
        // bit j: row 2ho+j-1 inside the image; bit 3+k: column 2wo+k-1 inside the image
        for (int j = 0; j < 3; ++j) flags |= (2 * ho + j - 1 >= 0 && 2 * ho + j - 1 < H) ? (1 << j) : 0;
        for (int k = 0; k < 3; ++k) flags |= (2 * wo + k - 1 >= 0 && 2 * wo + k - 1 < W) ? (8 << k) : 0;
      }
      s_base[threadIdx.x] = base;
      s_flag[threadIdx.x] = flags;
    }
    __syncthreads();
    // im2col tile: item = (position, tap)
    const int64_t chan_stride = (int64_t)T_ * H * W;
    for (int i = threadIdx.x; i < SW_POS * 32; i += 256) {
      const int pp = i / 32, tp = i % 32;
      float v = 0.f;
      const long long base = s_base[pp];
      if (base != kNoPos && tp < taps) {
        const int ci = tp / 9, j = (tp % 9) / 3, k = tp % 3;
        const int fl = s_flag[pp];
        if ((fl >> j) & (fl >> (3 + k)) & 1) v = __ldg(&x[base + ci * chan_stride + (int64_t)j * W + k]);
      }
      Xs[i] = v;
    }
    {
      // dy tile: 16-byte vector loads (Cop % 8 == 0), converted to fp32
      constexpr int VEC = Vec<T>::N;
      const int vpr = Cop / VEC;
      for (int i = threadIdx.x; i < SW_POS * vpr; i += 256) {
        const int pp = i / vpr, cvq = (i % vpr) * VEC;
        const int64_t p = ps + pp;
        float v[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) v[q] = 0.f;
        if (p < pe) load_vec<T>(dy + p * Cop + cvq, v);
#pragma unroll
        for (int q = 0; q < VEC; ++q) Ds[pp * Cop + cvq + q] = v[q];
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll 2
      for (int pp = slice; pp < SW_POS; pp += nslice) {
        const float* xp = &Xs[pp * 32 + tg * 3];
        const float x0 = xp[0], x1 = xp[1], x2 = xp[2];
        const float4 d0 = *reinterpret_cast<const float4*>(&Ds[pp * Cop + cg * 8]);
        const float4 d1 = *reinterpret_cast<const float4*>(&Ds[pp * Cop + cg * 8 + 4]);
        const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] = fmaf(x0, d[j], acc[0][j]);
          acc[1][j] = fmaf(x1, d[j], acc[1][j]);
          acc[2][j] = fmaf(x2, d[j], acc[2][j]);
        }
      }
    }
    __syncthreads();
  }
  // slices -> shared (reuse Xs as [taps][Cop] accumulators) -> one atomic per (channel, tap) per block
  float* red = Xs;
  for (int i = threadIdx.x; i < 32 * Cop; i += 256) red[i] = 0.f;
  __syncthreads();
  if (active) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[(tg * 3 + i) * Cop + cg * 8 + j], acc[i][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < taps * Cop; i += 256) {
    const int tap = i / Cop, c = i % Cop;
    const float v = red[i];
    if (c < Co && v != 0.f) atomicAdd(&dw[(int64_t)c * taps + tap], v);
  }
}

// ---- wgrad, register-tiled (the one that runs for the stock networks) ------------------------------------
// dw[co][ci][j][k] += sum_p dy[p][co] * x[n, ci, t, 2ho+j-1, 2wo+k-1].  M = N*T*Ho*Wo is 3.2 M positions at the benchmark
// shape, the result is a 24 x 27 matrix: the kernel is a skinny reduction that should run at memory speed (x once as
// fp32 NCDHW, dy once).  Warp = role (input channel ci, group of 8 output channels); its 32 lanes take 32 consecutive
// output positions of one (n, t) plane, load their 3 x 3 input patch of channel ci straight from the clip (stride-2
// lanes: every 128-byte line is used by the three k offsets) and one 16-byte vector of dy, and keep the 9 x 8 partial
// sums in registers over the CTA's whole share of chunks -- no shared-memory staging, 72 FMAs per 10 loads.  All roles
// of a CTA walk the same chunks, so the patch / dy lines are fetched from L2 once and re-read from L1.  One butterfly
// reduction per warp at the very end, then one fp32 red per (co, tap) per warp.
template <typename T, typename SRC, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) stem_wgrad_reg_kernel(const SRC x, const T* __restrict__ dy,
                                                              float* __restrict__ dw, int Ci, int T_, int H, int W,
                                                              int Ho, int Wo, int Co, int Cop, int64_t planes,
                                                              int chunks_per_plane, int64_t nchunks) {
  x3d::pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci = warp % Ci, cg = warp / Ci;
  const int HoWo = Ho * Wo;
  float2 acc[9][4];
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int64_t plane = ch / chunks_per_plane;            // = n * T + t   (warp-uniform)
    const int pp = (int)(ch - plane * chunks_per_plane) * 32 + lane;
    const bool live = pp < HoWo;
    const int ho = live ? pp / Wo : 0, wo = live ? pp - (pp / Wo) * Wo : 0;
    const int64_t n = plane / T_;
    const int t = (int)(plane - n * T_);
    float xv[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int hh = 2 * ho + j - 1;
      const bool rok = live && hh >= 0 && hh < H;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int ww = 2 * wo + k - 1;
        const bool ok = rok && ww >= 0 && ww < W;
        xv[j * 3 + k] = ok ? x(n, ci, t, hh, ww) : 0.f;
      }
    }
    float d[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) d[q] = 0.f;
    if (live) {
      const T* dp = dy + (plane * HoWo + pp) * Cop + cg * 8;
      if (sizeof(T) == 2) {
        float v[8];
        load_vec<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(dp), v);
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = v[q];
      } else {
        const float4 a = *reinterpret_cast<const float4*>(dp), b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dp) + 4);
        d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float2 xx = make_float2(xv[i], xv[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(xx, make_float2(d[2 * j], d[2 * j + 1]), acc[i][j]);
    }
  }
  // butterfly reduction over the 32 lanes; afterwards lane l publishes elements l, l+32, l+64 of the 72
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = acc[i][j].x, b = acc[i][j].y;
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      acc[i][j] = make_float2(a, b);
    }
  const int taps = Ci * 9;
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = i * 8 + 2 * j;                            // element index (tap i, channel 2j / 2j+1)
      if (lane == (e & 31) || lane == ((e + 1) & 31)) {
        const bool second = lane != (e & 31);
        const int c = cg * 8 + 2 * j + (second ? 1 : 0);
        const float v = second ? acc[i][j].y : acc[i][j].x;
        if (c < Co && v != 0.f) atomicAdd(&dw[(int64_t)c * taps + ci * 9 + i], v);
      }
    }
}

// ---- wgrad, cp.async-pipelined (fp32 NCDHW clips whose rows are 16-byte aligned: the benchmark shapes) ----------
// Same role decomposition as stem_wgrad_reg_kernel (warp = input channel x 8 output channels, 72 register accumulators),
// but the operands of a chunk -- 32 consecutive output positions of one output row: a 3 x 3 x 72-float input patch and
// 32 dy rows -- are staged through an NST-deep shared-memory ring with 16-byte cp.async (zero fill at the image
// borders).  In the register version the 9 warps of a CTA share one chunk, so only ~8 KB of distinct bytes were in
// flight per SM and the kernel ran at a quarter of the HBM rate (0.44 ms for 308 MB); here every CTA keeps NST-1
// chunks (4 KB each) in flight without holding registers for them.
constexpr int SP_NST = 6;        // ring depth
constexpr int SP_XSEG = 72;      // floats per (ci, row) segment: input columns 2*wo0-4 .. 2*wo0+67
template <typename T>
__global__ void __launch_bounds__(288, 2) stem_wgrad_pipe_kernel(const float* __restrict__ x, const T* __restrict__ dy,
                                                                 float* __restrict__ dw, int T_, int H, int W, int Ho,
                                                                 int Wo, int Co, int Cop, int wblocks, int64_t nchunks,
                                                                 int64_t chunks_per_cta, int aligned16) {
  x3d::pdl_prologue();
  extern __shared__ __align__(16) unsigned char sp_smem[];
  const int dy_bytes = 32 * Cop * (int)sizeof(T);
  const int stage_bytes = 9 * SP_XSEG * 4 + dy_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ci = warp % 3, cg = warp / 3;
  const int64_t c_begin = (int64_t)blockIdx.x * chunks_per_cta;
  int64_t c_end = c_begin + chunks_per_cta;
  if (c_end > nchunks) c_end = nchunks;
  const int n_mine = c_end > c_begin ? (int)(c_end - c_begin) : 0;
  const int64_t HW = (int64_t)H * W;
  const int n_x16 = 9 * (SP_XSEG / 4), n_d16 = dy_bytes / 16;        // 16-byte copies per chunk (162 + 96 | 192)

  // per-thread copy slots (chunk invariant): slot 0 = copy number tid, slot 1 = tid + 288 (fp32 dy only)
  int k_seg[2], k_piece[2], k_byte0[2], k_pos[2];
  int64_t k_xoff[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int q = tid + u * 288;
    k_seg[u] = -2;                                           // -2: unused, -1: dy copy, >= 0: x segment
    k_piece[u] = k_byte0[u] = k_pos[u] = 0;
    k_xoff[u] = 0;
    if (q < n_x16) {
      k_seg[u] = q / (SP_XSEG / 4);
      k_piece[u] = q - k_seg[u] * (SP_XSEG / 4);
      k_xoff[u] = (int64_t)(k_seg[u] / 3) * T_ * HW;         // channel plane offset
    } else if (q < n_x16 + n_d16) {
      k_seg[u] = -1;
      k_byte0[u] = (q - n_x16) * 16;
      k_pos[u] = k_byte0[u] / (Cop * (int)sizeof(T));
    }
  }
  // coordinates of the next chunk to be issued, advanced incrementally (no divisions in the loop)
  int is_wb, is_ho, is_t;
  int64_t is_n;
  {
    const int64_t c = c_begin;
    is_wb = (int)(c % wblocks);
    int64_t r = c / wblocks;
    is_ho = (int)(r % Ho);
    const int64_t plane = r / Ho;
    is_n = plane / T_;
    is_t = (int)(plane - is_n * T_);
  }
  auto issue = [&](int i) {                  // stage the next chunk into ring slot i % SP_NST
    const int wo0 = is_wb * 32, ho = is_ho;
    const int64_t plane = is_n * T_ + is_t;
    unsigned char* st = sp_smem + (size_t)(i % SP_NST) * stage_bytes;
    const float* xplane = x + (is_n * 3 * T_ + is_t) * HW;
    const unsigned char* dyrow = reinterpret_cast<const unsigned char*>(dy) +
                                 ((plane * Ho + ho) * (int64_t)Wo + wo0) * Cop * sizeof(T);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (k_seg[u] >= 0) {
        const int j = k_seg[u] % 3;
        const int hh = 2 * ho + j - 1;
        const int col0 = 2 * wo0 - 4 + 4 * k_piece[u];
        int valid = 0;
        if (hh >= 0 && hh < H && col0 >= 0) {
          valid = W - col0;
          valid = valid > 4 ? 4 : (valid < 0 ? 0 : valid);
        }
        const float* src = xplane + k_xoff[u] + (valid ? (int64_t)hh * W + col0 : 0);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(st + (k_seg[u] * SP_XSEG + k_piece[u] * 4) * 4);
        if (aligned16) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(valid * 4) : "memory");
        } else {
          // image rows that are not 16-byte aligned (W = 111, 158 of the multigrid shapes): four 4-byte copies
#pragma unroll
          for (int e = 0; e < 4; ++e)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst + 4 * e), "l"(src + (e < valid ? e : 0)),
                         "r"(e < valid ? 4 : 0)
                         : "memory");
        }
      } else if (k_seg[u] == -1) {
        const bool ok = wo0 + k_pos[u] < Wo;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(st + 9 * SP_XSEG * 4 + k_byte0[u]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(dyrow + (ok ? k_byte0[u] : 0)),
                     "r"(ok ? 16 : 0)
                     : "memory");
      }
    }
    if (++is_wb == wblocks) {
      is_wb = 0;
      if (++is_ho == Ho) {
        is_ho = 0;
        if (++is_t == T_) { is_t = 0; ++is_n; }
      }
    }
  };

  float2 acc[9][4];
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int i = 0; i < SP_NST - 1; ++i) {
    if (i < n_mine) issue(i);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
#pragma unroll 1
  for (int i = 0; i < n_mine; ++i) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(SP_NST - 2) : "memory");    // this thread's copies of chunk i landed
    __syncthreads();                                                           // ... everybody's; slot of chunk i-1 is free
    if (i + SP_NST - 1 < n_mine) issue(i + SP_NST - 1);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    const unsigned char* st = sp_smem + (size_t)(i % SP_NST) * stage_bytes;
    const float* xs = reinterpret_cast<const float*>(st) + (ci * 3) * SP_XSEG + 3 + 2 * lane;      // col 2*wo-1
    float xv[9];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) xv[j * 3 + k] = xs[j * SP_XSEG + k];
    float d[8];
    const T* dp = reinterpret_cast<const T*>(st + 9 * SP_XSEG * 4) + lane * Cop + cg * 8;
    if (sizeof(T) == 2) {
      load_vec<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(dp), d);
    } else {
      const float4 a = *reinterpret_cast<const float4*>(dp), b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dp) + 4);
      d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
    }
#pragma unroll
    for (int ii = 0; ii < 9; ++ii) {
      const float2 xx = make_float2(xv[ii], xv[ii]);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[ii][j] = __ffma2_rn(xx, make_float2(d[2 * j], d[2 * j + 1]), acc[ii][j]);
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = acc[i][j].x, b = acc[i][j].y;
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      acc[i][j] = make_float2(a, b);
    }
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = i * 8 + 2 * j;
      if (lane == (e & 31) || lane == ((e + 1) & 31)) {
        const bool second = lane != (e & 31);
        const int c = cg * 8 + 2 * j + (second ? 1 : 0);
        const float v = second ? acc[i][j].y : acc[i][j].x;
        if (c < Co && v != 0.f) atomicAdd(&dw[(int64_t)c * 27 + ci * 9 + i], v);
      }
    }
}

extern "C" int x3d_stem_conv_s_wgrad(const float* x, const void* dy, float* dw, int64_t N, int64_t Ci, int64_t T_,
                                     int64_t H, int64_t W, int64_t Co, int64_t Cop, x3d_dtype_t dt,
                                     x3d_stream_t stream) {
  X3D_CHECK_ARG(Ci >= 1 && Ci <= 3, "n_input_channels must be <= 3");
  X3D_CHECK_ARG(Cop % 8 == 0 && Cop >= Co && Cop <= 64, "Cop must be a multiple of 8, <= 64");
  const int Ho = (int)((H + 2 - 3) / 2 + 1), Wo = (int)((W + 2 - 3) / 2 + 1);
  const int64_t total = N * T_ * Ho * Wo;
  if (total == 0) return 0;
  const int roles = (int)(Ci * (Cop / 8));
  static const bool old_kernel = getenv("X3D_STEM_WGRAD_SMEM") != nullptr;      // A/B switch
  static const bool no_pipe = getenv("X3D_STEM_WGRAD_REG") != nullptr;          // A/B switch
  if (Ci == 3 && Cop == 24 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && !old_kernel && !no_pipe) {
    const int aligned16 = (W % 4 == 0) ? 1 : 0;
    const int wblocks = (int)cdiv(Wo, 32);
    const int64_t nchunks = N * T_ * Ho * wblocks;
    int64_t blocks = 2 * kNumSMs;
    if (blocks > nchunks) blocks = nchunks;
    const int64_t cpc = cdiv(nchunks, blocks);
    blocks = cdiv(nchunks, cpc);
    X3D_DISPATCH_DTYPE(dt, {
      const size_t smem = (size_t)SP_NST * (9 * SP_XSEG * 4 + 32 * Cop * sizeof(T));
      x3d::launch(stem_wgrad_pipe_kernel<T>, (unsigned)blocks, 288, smem, as_stream(stream), x, (const T*)dy, dw, (int)T_,
                  (int)H, (int)W, Ho, Wo, (int)Co, (int)Cop, wblocks, nchunks, cpc, aligned16);
    });
    X3D_LAUNCH_CHECK();
    return 0;
  }
  if (roles <= 12 && (int64_t)Ho * Wo < (1ll << 30) && !old_kernel) {
    const int cpp = (int)cdiv((int64_t)Ho * Wo, 32);
    const int64_t planes = N * T_, nchunks = planes * cpp;
    int64_t blocks = 2 * kNumSMs;
    if (blocks > nchunks) blocks = nchunks;
    // <= 9 roles (the stock 3 -> 24 stem): 288 threads at <= 112 registers, two CTAs per SM
    const SrcF32 srcx{x, (int)Ci, (int)T_, (int)H, (int)W};
#define SW_(MAXT, MINB)                                                                                          \
  x3d::launch(stem_wgrad_reg_kernel<T, SrcF32, MAXT, MINB>, (unsigned)blocks, roles * 32, 0, as_stream(stream), srcx,     \
              (const T*)dy, dw, (int)Ci, (int)T_, (int)H, (int)W, Ho, Wo, (int)Co, (int)Cop, planes, cpp, nchunks)
    X3D_DISPATCH_DTYPE(dt, {
      if (roles <= 9) SW_(288, 2);
      else SW_(384, 1);
    });
#undef SW_
    X3D_LAUNCH_CHECK();
    return 0;
  }
  int64_t blocks = 4 * kNumSMs;
  int64_t ppb = cdiv(cdiv(total, blocks), SW_POS) * SW_POS;
  blocks = cdiv(total, ppb);
  size_t smem = (size_t)(SW_POS * 32 + SW_POS * Cop) * sizeof(float) + SW_POS * (sizeof(long long) + sizeof(int));
  X3D_DISPATCH_DTYPE(dt, (x3d::launch(stem_wgrad_kernel<T>, (unsigned)blocks, 256, smem, as_stream(stream), 
                             x, (const T*)dy, dw, (int)Ci, (int)T_, (int)H, (int)W, Ho, Wo, (int)Co, (int)Cop, total, ppb)));
  X3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int x3d_stem_conv_s_wgrad_u8(const uint8_t* src, const x3d_crop_t* crops_dev, const void* dy, float* dw,
                                        int64_t N, int64_t T_, int64_t Hs, int64_t Ws, int64_t S, const float* mean_std,
                                        float norm_value, int64_t Co, int64_t Cop, x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cop % 8 == 0 && Cop >= Co && Cop <= 32, "Cop must be a multiple of 8, <= 32");
  X3D_CHECK_ARG(S >= 1 && S <= Hs && S <= Ws && mean_std != nullptr && crops_dev != nullptr, "crop larger than the frames");
  const int Ho = (int)((S + 2 - 3) / 2 + 1);
  if (N * T_ * Ho == 0) return 0;
  const int roles = (int)(3 * (Cop / 8));
  const int cpp = (int)cdiv((int64_t)Ho * Ho, 32);
  const int64_t planes = N * T_, nchunks = planes * cpp;
  int64_t blocks = 2 * kNumSMs;
  if (blocks > nchunks) blocks = nchunks;
  const SrcU8 srcx = make_src_u8(src, crops_dev, T_, Hs, Ws, S, mean_std, norm_value);
#define SW_(MAXT, MINB)                                                                                          \
  x3d::launch(stem_wgrad_reg_kernel<T, SrcU8, MAXT, MINB>, (unsigned)blocks, roles * 32, 0, as_stream(stream), srcx,      \
              (const T*)dy, dw, 3, (int)T_, (int)S, (int)S, Ho, Ho, (int)Co, (int)Cop, planes, cpp, nchunks)
  X3D_DISPATCH_DTYPE(dt, {
    if (roles <= 9) SW_(288, 2);
    else SW_(384, 1);
  });
#undef SW_
  X3D_LAUNCH_CHECK();
  return 0;
}
