// Tiled depthwise 3x3x3 forward for the hot shapes (stride (1,1,1) and (1,2,2), pad 1), NDHWC.
//
// Design (see DESIGN.md "dwconv"):
//  * CTA = (sample n, TH x TW output tile, chunk of CC channels), marching over ALL T planes.
//    Each input plane tile (with its 1-pixel halo) is brought ONCE into shared memory by TMA
//    (cp.async.bulk.tensor.5d, one instruction issued by one thread per plane, completion on an
//    mbarrier) into a 3-deep ring: two planes are in flight while one is consumed.
//  * thread = (channel PAIR, 2 x PW output patch).  The 27 taps of the pair live in registers as float2;
//    every shared-memory word (2 channels) feeds up to 27 packed FFMA2 (fma.rn.f32x2): one input plane
//    contributes to three output planes held in register accumulators whose roles rotate by a 3x
//    unrolled plane loop (no register moves).  CC and TW are template parameters so that every window
//    LDS uses an immediate offset (no address arithmetic in the inner loop).
//  * the preceding SubBatchNorm3d+ReLU (scale/shift per (split,channel)) is applied on the fly to the
//    window values.  Zero padding must be applied AFTER that transform (x3d.py:147-150): the tensor map
//    fills out-of-image halo elements with NaN (CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA) and
//    relu is fmaxf(v, 0), which returns the non-NaN operand -- so padded taps become exact zeros without
//    any select in the inner loop.  Without the fused transform the OOB fill is plain zero.
//  * epilogue: bf16/fp32 store + per-(sample,channel) sum / sum-of-squares of the stored values
//    (bn2 statistics and the SE global pool) -> shared atomics -> one fp64 atomic per channel per CTA.
// Arithmetic intensity at stride 1 in bf16 is 27 FMA / 4 B = 6.75 FMA/B, above the B200 balance of
// ~5.7 FMA/B (37 TFMA/s fp32 vs 6.5 TB/s): the stride-1 layers are bound by the fp32 FMA pipe, which is
// why the inner loop is FFMA2 and everything else is kept off that pipe.
#include <cuda.h>

#include "common.cuh"

using namespace x3d;

namespace {

constexpr int PH = 2;             // output patch per thread: PH x PW
constexpr int NSTAGE = 3;         // input-plane ring
template <int PW> struct Cfg { static constexpr int MAXT = PW == 4 ? 384 : 256, MINB = PW == 4 ? 1 : 2; };

// ---- mbarrier / TMA primitives -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

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
// predicated store of a channel pair; returns the values as stored (for the statistics)
template <typename T>
__device__ __forceinline__ float2 st_pair_if(T* p, float2 v, uint32_t pred);
template <>
__device__ __forceinline__ float2 st_pair_if<__nv_bfloat16>(__nv_bfloat16* p, float2 v, uint32_t pred) {
  const uint32_t w = pack_bf16x2(v.x, v.y);
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p st.global.b32 [%0], %1;\n}\n" ::"l"(p), "r"(w), "r"(pred) : "memory");
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <>
__device__ __forceinline__ float2 st_pair_if<float>(float* p, float2 v, uint32_t pred) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %3, 0;\n@p st.global.v2.f32 [%0], {%1, %2};\n}\n" ::"l"(p), "f"(v.x), "f"(v.y),
               "r"(pred)
               : "memory");
  return v;
}

struct TileGeom {
  int T, Ho, Wo, Cp;
  int TH;                // output tile height (TW and CC are template parameters)
  int tiles_w;
  int stage_elems;       // ring-slot stride in elements (128-byte aligned)
};

template <typename T, int S, int CC, int TW, int PW, bool XF>
__global__ void __launch_bounds__(Cfg<PW>::MAXT, Cfg<PW>::MINB)
dw3_fwd_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ w, T* __restrict__ y,
                     const TileGeom g, const float* __restrict__ scale, const float* __restrict__ shift, int splits,
                     double* __restrict__ stats) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int WR = (PH - 1) * S + 3, WC = (PW - 1) * S + 3;   // per-thread input window
  constexpr int IW = (TW - 1) * S + 3;                           // CTA input tile width (with halo)
  constexpr int ROW = IW * CC;                                   // smem row stride (elements)
  constexpr int PAIRS = CC / 2;
  constexpr int PPR = TW / PW;                                   // patches per tile row
  constexpr int NO = PH * PW;
  const int IH = (g.TH - 1) * S + 3;
  const int Cp = g.Cp;
  T* const sbuf = reinterpret_cast<T*>(smem_raw);               // NSTAGE plane buffers (ring)
  float* s_stat = reinterpret_cast<float*>(sbuf + NSTAGE * g.stage_elems);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_stat + 2 * CC);

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int ho0 = (blockIdx.x / g.tiles_w) * g.TH, wo0 = (blockIdx.x % g.tiles_w) * TW;
  const int cbase = blockIdx.y * CC;
  const int n = blockIdx.z;
  const int pair = tid % PAIRS, patch = tid / PAIRS;
  const int py = patch / PPR, px = patch % PPR;
  const int c = cbase + 2 * pair;
  const bool ch_ok = c < Cp;
  const int hi0 = ho0 * S - 1, wi0 = wo0 * S - 1;
  const int nT = g.T;
  const uint32_t plane_bytes = (uint32_t)(IH * ROW * sizeof(T));

  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < NSTAGE; ++k) mbar_init(&full_bar[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int t, int slot) {      // one thread: arm the barrier, fire one TMA box load
    mbar_expect_tx(&full_bar[slot], plane_bytes);
    tma_load_5d(sbuf + slot * g.stage_elems, &tmap, &full_bar[slot], cbase, wi0, hi0, t, n);
  };
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < NSTAGE - 1; ++k)
      if (k < nT) issue(k, k);
  }

  // ---- per-thread constants ----------------------------------------------------------------
  float2 wreg[27];
#pragma unroll
  for (int tap = 0; tap < 27; ++tap)
    wreg[tap] = ch_ok ? *reinterpret_cast<const float2*>(w + (int64_t)tap * Cp + c) : make_float2(0.f, 0.f);
  float2 sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f);
  if (XF && ch_ok) {
    const int b = n % splits;
    sc = *reinterpret_cast<const float2*>(scale + (int64_t)b * Cp + c);
    sh = *reinterpret_cast<const float2*>(shift + (int64_t)b * Cp + c);
  }
  const int win_base = ((py * PH * S) * IW + px * PW * S) * CC + 2 * pair;

  float2 accA[NO], accB[NO], accC[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) accA[o] = accB[o] = accC[o] = make_float2(0.f, 0.f);
  float2 ssum = make_float2(0.f, 0.f), ssq = make_float2(0.f, 0.f);

  const int ho_t = ho0 + py * PH, wo_t = wo0 + px * PW;
  uint32_t omask = 0;                      // validity of the patch outputs (bit oy*PW+ox)
#pragma unroll
  for (int oy = 0; oy < PH; ++oy)
#pragma unroll
    for (int ox = 0; ox < PW; ++ox)
      if (ch_ok && ho_t + oy < g.Ho && wo_t + ox < g.Wo) omask |= 1u << (oy * PW + ox);
  const int out_plane = g.Ho * g.Wo * Cp;  // < 2^31 elements (checked on the host)
  const int orow = g.Wo * Cp;
  T* yp = y + ((int64_t)n * nT * g.Ho + ho_t) * (int64_t)orow + (int64_t)wo_t * Cp + c;   // output plane 0

  // store a finished accumulator as the next output plane (planes are finished in order 0,1,2,...)
  auto store_plane = [&](float2 (&a)[NO]) {
#pragma unroll
    for (int oy = 0; oy < PH; ++oy) {
#pragma unroll
      for (int ox = 0; ox < PW; ++ox) {
        const uint32_t ok = omask & (1u << (oy * PW + ox));
        float2 r = st_pair_if<T>(yp + oy * orow + ox * Cp, a[oy * PW + ox], ok);
        if (!ok) r = make_float2(0.f, 0.f);
        ssum.x += r.x; ssum.y += r.y;
        ssq = __ffma2_rn(r, r, ssq);
        a[oy * PW + ox] = make_float2(0.f, 0.f);
      }
    }
    yp += out_plane;
  };

  // One plane: input plane `tin` feeds output planes tin+1 (kt=0, accumulator `nw`), tin (kt=1, `md`)
  // and tin-1 (kt=2, `od`), which is complete afterwards and is stored.
  int slot = 0;
  uint32_t parity = 0;
  auto plane = [&](int tin, float2 (&nw)[NO], float2 (&md)[NO], float2 (&od)[NO]) {
    mbar_wait(&full_bar[slot], parity);    // TMA bytes of plane tin have landed
    __syncthreads();                       // everybody is done with plane tin-1: its ring slot is free
    if (tid == 0) {
      const int tn = tin + NSTAGE - 1;
      int sn = slot + NSTAGE - 1;
      if (sn >= NSTAGE) sn -= NSTAGE;
      if (tn < nT) issue(tn, sn);
    }
    const T* bp = sbuf + slot * g.stage_elems + win_base;
    if (++slot == NSTAGE) { slot = 0; parity ^= 1u; }
#pragma unroll
    for (int r = 0; r < WR; ++r) {
#pragma unroll
      for (int cc = 0; cc < WC; ++cc) {
        float2 xv = lds_pair<T>(bp + r * ROW + cc * CC);
        if (XF) {
          xv = __ffma2_rn(xv, sc, sh);
          xv.x = fmaxf(xv.x, 0.f);          // also maps the NaN halo to exact zero
          xv.y = fmaxf(xv.y, 0.f);
        }
#pragma unroll
        for (int oy = 0; oy < PH; ++oy) {
          const int kh = r - oy * S;
          if (kh < 0 || kh > 2) continue;
#pragma unroll
          for (int ox = 0; ox < PW; ++ox) {
            const int kw = cc - ox * S;
            if (kw < 0 || kw > 2) continue;
            const int o = oy * PW + ox, tap = kh * 3 + kw;
            nw[o] = __ffma2_rn(wreg[tap], xv, nw[o]);
            md[o] = __ffma2_rn(wreg[9 + tap], xv, md[o]);
            od[o] = __ffma2_rn(wreg[18 + tap], xv, od[o]);
          }
        }
      }
    }
    if (tin >= 1) store_plane(od);
    else {
#pragma unroll
      for (int o = 0; o < NO; ++o) od[o] = make_float2(0.f, 0.f);
    }
  };

  int tin = 0;
  for (; tin + 3 <= nT; tin += 3) {
    plane(tin, accA, accB, accC);
    plane(tin + 1, accC, accA, accB);
    plane(tin + 2, accB, accC, accA);
  }
  // remainder; the last executed plane's `md` accumulator holds output plane T-1
  const int rem = nT - tin;
  if (rem == 0) {
    store_plane(accC);                     // last call was plane(.., accB, accC, accA): md = accC
  } else if (rem == 1) {
    plane(tin, accA, accB, accC);
    store_plane(accB);
  } else {
    plane(tin, accA, accB, accC);
    plane(tin + 1, accC, accA, accB);
    store_plane(accA);
  }

  if (stats != nullptr) {
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
  int threads, CC, TW, PW, IH, IW;
  size_t smem;
  bool ok;
};

// candidates compiled below
constexpr int kCC[3] = {48, 56, 72};

template <typename T>
TilePlan plan_tiles(int64_t N, int T_, int H, int W, int Cp, int S, int PW) {
  TilePlan p;
  p.ok = false;
  p.PW = PW;
  const int MAX_THREADS = PW == 4 ? 384 : 256;
  TileGeom& g = p.g;
  g.T = T_; g.Cp = Cp;
  g.Ho = (H + 2 - 3) / S + 1;
  g.Wo = (W + 2 - 3) / S + 1;
  const int esz = (int)sizeof(T);
  // Pick (TH, TW, CC) maximising the fraction of useful lanes (ragged tiles, partial channel chunks)
  // with a mild penalty for halo re-reads and for small CTAs.
  double best = -1.0;
  for (int ci = 0; ci < 3; ++ci) {
    const int CC = kCC[ci];
    const int nchunk = (Cp + CC - 1) / CC;
    for (int twi = 0; twi < 2; ++twi) {
      const int TW = (S == 1) ? (twi ? 16 : 8) : (twi ? 8 : 4);
      if (TW % PW) continue;
      for (int TH = PH; TH <= 8; TH += PH) {
        const int patches = (TH / PH) * (TW / PW);
        const int threads = patches * (CC / 2);
        if (threads < 96 || threads > MAX_THREADS) continue;
        const int th = (g.Ho + TH - 1) / TH, tw = (g.Wo + TW - 1) / TW;
        const int IH = (TH - 1) * S + 3, IW = (TW - 1) * S + 3;
        if (IH > 256 || IW > 256) continue;
        const size_t smem = (size_t)NSTAGE * IH * IW * CC * esz;
        if (smem > 160 * 1024) continue;
        const double useful = (double)g.Ho * g.Wo * Cp / ((double)th * TH * tw * TW * nchunk * CC);
        const double halo = (double)IH * IW / ((double)TH * S * TW * S);
        const double score = useful / (1.0 + 0.3 * (halo - 1.0)) * (threads >= 192 ? 1.0 : 0.9);
        if (score > best) {
          best = score;
          g.TH = TH; p.TW = TW; p.CC = CC; p.threads = threads;
        }
      }
    }
  }
  if (best < 0.45) return p;    // too wasteful (exotic channel counts): let the direct kernel do it
  g.tiles_w = (g.Wo + p.TW - 1) / p.TW;
  const int tiles_h = (g.Ho + g.TH - 1) / g.TH;
  p.IH = (g.TH - 1) * S + 3;
  p.IW = (p.TW - 1) * S + 3;
  const size_t plane_bytes = (size_t)p.IH * p.IW * p.CC * esz;
  const size_t stage_bytes = (plane_bytes + 127) / 128 * 128;
  g.stage_elems = (int)(stage_bytes / esz);
  p.smem = NSTAGE * stage_bytes + (size_t)p.CC * 2 * sizeof(float) + NSTAGE * sizeof(uint64_t) + 16;
  p.grid = dim3((unsigned)(g.tiles_w * tiles_h), (unsigned)((Cp + p.CC - 1) / p.CC), (unsigned)N);
  if (N > 65535 || (int64_t)g.Ho * g.Wo * Cp >= (1ll << 31)) return p;
  p.ok = true;
  return p;
}

// ---- tensor map (driver entry point fetched through the runtime: no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <typename T>
bool make_input_map(CUtensorMap* map, const void* x, int64_t N, int T_, int H, int W, int Cp, int CC, int IW, int IH,
                    bool nan_fill) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  const cuuint64_t esz = sizeof(T);
  cuuint64_t dims[5] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T_, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)Cp * esz, (cuuint64_t)W * Cp * esz, (cuuint64_t)H * W * Cp * esz,
                           (cuuint64_t)T_ * H * W * Cp * esz};
  cuuint32_t box[5] = {(cuuint32_t)CC, (cuuint32_t)IW, (cuuint32_t)IH, 1u, 1u};
  cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  const CUtensorMapDataType dt = sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(map, dt, 5, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <typename T, int S, int CC, int TW, int PW, bool XF>
void launch_one(const TilePlan& p, const CUtensorMap& map, const float* w, void* y, const float* scale,
                const float* shift, int splits, double* stats, cudaStream_t stream) {
  auto kfn = dw3_fwd_tiled_kernel<T, S, CC, TW, PW, XF>;
  static bool attr_done = false;   // per instantiation
  if (!attr_done) {
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  kfn<<<p.grid, p.threads, p.smem, stream>>>(map, w, (T*)y, p.g, scale, shift, splits, stats);
}

template <typename T, int S, int PW>
int launch_tiled(const TilePlan& p, const CUtensorMap& map, const float* w, void* y, const float* scale,
                 const float* shift, int splits, double* stats, cudaStream_t stream) {
  constexpr int TWa = (S == 1) ? 8 : 4, TWb = (S == 1) ? 16 : 8;
  const bool xf = scale != nullptr;
#define L2_(CCv, TWv)                                                                              \
  do {                                                                                             \
    if (xf) launch_one<T, S, CCv, TWv, PW, true>(p, map, w, y, scale, shift, splits, stats, stream); \
    else launch_one<T, S, CCv, TWv, PW, false>(p, map, w, y, scale, shift, splits, stats, stream);  \
  } while (0)
#define L_(CCv)                                    \
  if (p.CC == CCv) {                               \
    if (p.TW == TWa) {                             \
      if (TWa % PW == 0) L2_(CCv, (TWa % PW == 0 ? TWa : TWb)); \
    } else L2_(CCv, TWb);                          \
    return 0;                                      \
  }
  L_(48) L_(56) L_(72)
#undef L_
#undef L2_
  return -1;
}

template <typename T>
int run_tiled(const void* x, const float* w, void* y, int64_t N, int T_, int H, int W, int Cp, int stride,
              const float* scale, const float* shift, int splits, double* stats, cudaStream_t stream, int PW,
              bool* handled) {
  TilePlan p = plan_tiles<T>(N, T_, H, W, Cp, stride, PW);
  if (!p.ok) return 0;
  CUtensorMap map;
  if (!make_input_map<T>(&map, x, N, T_, H, W, Cp, p.CC, p.IW, p.IH, scale != nullptr)) return 0;
  int rc;
  if (PW == 4)
    rc = stride == 1 ? launch_tiled<T, 1, 4>(p, map, w, y, scale, shift, splits, stats, stream)
                     : launch_tiled<T, 2, 4>(p, map, w, y, scale, shift, splits, stats, stream);
  else
    rc = stride == 1 ? launch_tiled<T, 1, 2>(p, map, w, y, scale, shift, splits, stats, stream)
                     : launch_tiled<T, 2, 2>(p, map, w, y, scale, shift, splits, stats, stream);
  if (rc != 0) return 0;
  *handled = true;
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
  static const int PWsel = getenv("X3D_DW_PW") ? atoi(getenv("X3D_DW_PW")) : 2;
  if (in_scale != nullptr && !relu_in) return 0;   // the fused transform of this kernel is BN + ReLU
  const int PW = PWsel == 4 ? 4 : 2;
  if (dt == X3D_BF16)
    run_tiled<__nv_bfloat16>(x, w_packed, y, N, (int)T_, (int)H, (int)W, (int)Cp, stride, in_scale, in_shift, splits,
                             stats, stream, PW, handled);
  else
    run_tiled<float>(x, w_packed, y, N, (int)T_, (int)H, (int)W, (int)Cp, stride, in_scale, in_shift, splits, stats,
                     stream, PW, handled);
  if (!*handled) return 0;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("dwconv_fwd_tiled: launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  count_launch();
  return 0;
}
}  // namespace x3d
