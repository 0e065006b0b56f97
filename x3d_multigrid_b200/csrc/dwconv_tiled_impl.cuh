// Tiled depthwise 3x3x3 forward AND dgrad for the hot shapes (stride (1,1,1) / (1,2,2), pad 1), NDHWC.
//
// Design (see DESIGN.md "dwconv"):
//  * CTA = (sample n, TH x TW output tile, chunk of CC channels), marching over ALL T planes.
//    Each input plane tile (with its halo) is brought ONCE into shared memory by TMA
//    (cp.async.bulk.tensor.5d, one instruction issued by one thread per plane, completion on an
//    mbarrier) into a 3-deep ring: two planes are in flight while one is consumed.
//  * thread = (channel PAIR, 2 x PW output patch).  The 27 taps of the pair live in registers as float2;
//    every shared-memory word (2 channels) feeds up to 27 packed FFMA2 (fma.rn.f32x2): one input plane
//    contributes to three output planes held in register accumulators whose roles rotate by a 3x
//    unrolled plane loop (no register moves).  CC and TW are template parameters so that every window
//    LDS uses an immediate offset (no address arithmetic in the inner loop).
//  * four tap mappings share the kernel (template MODE): forward stride 1 / 2, and dgrad stride 1 / 2
//    (dgrad s1 = correlation with the flipped taps; dgrad s2 gathers, per output parity, only the taps
//    that hit a sampled position: 6.75 FMA per dx element instead of 27).
//  * forward: the preceding SubBatchNorm3d+ReLU (scale/shift per (split,channel)) is applied on the fly
//    to the window values.  Zero padding must be applied AFTER that transform (x3d.py:147-150): the
//    tensor map fills out-of-image halo elements with NaN (CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_
//    ZERO_FMA) and relu is fmaxf(v, 0), which returns the non-NaN operand -- padded taps become exact
//    zeros without any select in the inner loop.  Without the fused transform the OOB fill is zero.
//  * forward epilogue: store + per-(sample,channel) sum / sum-of-squares of the stored values (bn2
//    statistics and the SE global pool).  dgrad epilogue: ReLU mask of the preceding BN+ReLU recomputed
//    from the saved conv1 output (aux), store, and the two BN-backward sums (sum d, sum d*aux).
//    Both: ordered block reduction through a scratch array carved from the idle ring (no shared atomics), then one
//    fp64 atomic per channel per CTA.
// Arithmetic intensity of the stride-1 forward in bf16 is 27 FMA / 4 B = 6.75 FMA/B, above the B200
// balance of ~5.7 FMA/B (37 TFMA/s fp32 vs 6.5 TB/s): those layers are bound by the fp32 FMA pipe, which
// is why the inner loop is FFMA2 and everything else is kept off that pipe.
#include <cuda.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "dwconv_tiled.h"

using namespace x3d;

namespace {

constexpr int PH = 2;             // output patch per thread: PH x PW
constexpr int NSTAGE = 3;         // input-plane ring
template <int PW> struct Cfg { static constexpr int MAXT = PW == 4 ? 384 : 256, MINB = PW == 4 ? 1 : 2; };

enum { M_FWD1 = 0, M_FWD2 = 1, M_DG1 = 2, M_DG2 = 3 };
// geometry of the four tap mappings; "out" = tensor being produced, "in" = tensor staged through smem
template <int MODE>
struct Map {
  static constexpr bool DG = MODE >= 2;
  // input extent needed for an output extent of n
  __host__ __device__ static constexpr int in_ext(int n) {
    return MODE == M_FWD1 ? n + 2 : MODE == M_FWD2 ? (n - 1) * 2 + 3 : MODE == M_DG1 ? n + 2 : n / 2 + 1;
  }
  // input coordinate of the first staged row/col for an output tile starting at o0
  __host__ __device__ static constexpr int in_org(int o0) {
    return MODE == M_FWD1 ? o0 - 1 : MODE == M_FWD2 ? o0 * 2 - 1 : MODE == M_DG1 ? o0 - 1 : o0 / 2;
  }
  // offset (in input positions) of a patch window that starts p output positions into the tile
  __host__ __device__ static constexpr int win_off(int p) {
    return MODE == M_FWD1 ? p : MODE == M_FWD2 ? p * 2 : MODE == M_DG1 ? p : p / 2;
  }
  // spatial tap index used by window element r for patch output o (valid iff in [0,2])
  __host__ __device__ static constexpr int tap(int r, int o) {
    return MODE == M_FWD1 ? r - o : MODE == M_FWD2 ? r - 2 * o : MODE == M_DG1 ? o - r + 2 : o + 1 - 2 * r;
  }
  // temporal tap used for the output planes tin+1 / tin / tin-1
  static constexpr int KT_NEW = DG ? 2 : 0, KT_MID = 1, KT_OLD = DG ? 0 : 2;
};

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
__device__ __forceinline__ float2 unpack_pair(uint32_t w0, uint32_t w1);
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
// predicated global load of a channel pair (0 when off)
template <typename T>
__device__ __forceinline__ float2 ldg_pair_if(const T* p, uint32_t pred);
template <>
__device__ __forceinline__ float2 ldg_pair_if<__nv_bfloat16>(const __nv_bfloat16* p, uint32_t pred) {
  uint32_t w = 0;
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p ld.global.nc.b32 %0, [%1];\n}\n" : "+r"(w) : "l"(p), "r"(pred));
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <>
__device__ __forceinline__ float2 ldg_pair_if<float>(const float* p, uint32_t pred) {
  float2 v = make_float2(0.f, 0.f);
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %3, 0;\n@p ld.global.nc.v2.f32 {%0, %1}, [%2];\n}\n"
               : "+f"(v.x), "+f"(v.y)
               : "l"(p), "r"(pred));
  return v;
}
// raw variants: the loaded register(s) are NOT touched at the load site, so the scoreboard wait happens where the
// value is finally unpacked (the store phase), not right behind the load
template <typename T> struct RawPair;
template <> struct RawPair<__nv_bfloat16> { uint32_t w; };
template <> struct RawPair<float> { float2 v; };
__device__ __forceinline__ RawPair<__nv_bfloat16> ldg_raw_if(const __nv_bfloat16* p, uint32_t pred) {
  RawPair<__nv_bfloat16> r;
  r.w = 0;
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p ld.global.nc.b32 %0, [%1];\n}\n" : "+r"(r.w) : "l"(p), "r"(pred));
  return r;
}
__device__ __forceinline__ RawPair<float> ldg_raw_if(const float* p, uint32_t pred) {
  RawPair<float> r;
  r.v = ldg_pair_if<float>(p, pred);
  return r;
}
__device__ __forceinline__ float2 unpack_raw(RawPair<__nv_bfloat16> r) {
  return make_float2(__uint_as_float(r.w << 16), __uint_as_float(r.w & 0xffff0000u));
}
__device__ __forceinline__ float2 unpack_raw(RawPair<float> r) { return r.v; }
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
  int T, Ho, Wo, Cp;     // extents of the tensor being PRODUCED (y for forward, dx for dgrad)
  int TH;                // output tile height (TW and CC are template parameters)
  int tiles_w;
  int tiles_n;           // wgrad: number of samples (units = samples x spatial tiles)
  int stage_elems;       // ring-slot stride in elements (128-byte aligned)
  int aux_stage_elems;   // stride-2 dgrad: aux-tile ring-slot stride in elements
};

// XF: fused relu(x*scale+shift) on the staged input (forward modes)
// AUX: dgrad epilogue with the ReLU mask / BN sums from the saved conv1 output
template <typename T, int MODE, int CC, int TW, int PW, bool XF, bool AUX>
__global__ void __launch_bounds__(Cfg<PW>::MAXT, Cfg<PW>::MINB)
dw3_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap amap,
                 const float* __restrict__ w, T* __restrict__ y,
                 const TileGeom g, const float* __restrict__ scale, const float* __restrict__ shift, int splits,
                 const T* __restrict__ aux, double* __restrict__ stats) {
  x3d::pdl_trigger();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using M = Map<MODE>;
  // Stride-2 dgrad does only 6.75 FMA per output: its plane time is about one global-load latency, and the per-thread
  // aux loads (issued at the start of a plane, needed at its end) were 55 % of its stall samples (ncu).  For that
  // mapping the aux tile of output plane t rides the TMA ring together with input plane t (ring of 4, look-ahead 2)
  // and is read from shared memory one call later.
  constexpr bool AUXT = AUX && (MODE == M_DG2 || MODE == M_DG1);
  constexpr int NSTAGE = AUXT ? 4 : 3;
  constexpr int LOOK = AUXT ? 2 : NSTAGE - 1;                    // planes of TMA look-ahead
  constexpr int WR = M::in_ext(PH), WC = M::in_ext(PW);         // per-thread input window
  constexpr int IW = M::in_ext(TW);                              // CTA input tile width (with halo)
  constexpr int ROW = IW * CC;                                   // smem row stride (elements)
  constexpr int PAIRS = CC / 2;
  constexpr int PPR = TW / PW;                                   // patches per tile row
  constexpr int NO = PH * PW;
  const int IH = M::in_ext(g.TH);
  const int Cp = g.Cp;
  T* const sbuf = reinterpret_cast<T*>(smem_raw);               // NSTAGE plane buffers (ring)
  T* const abuf = sbuf + NSTAGE * g.stage_elems;                // AUXT: NSTAGE aux tiles [TH][TW][CC]
  float* s_stat = reinterpret_cast<float*>(abuf + (AUXT ? NSTAGE * g.aux_stage_elems : 0));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_stat + 2 * CC);

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int ho0 = (blockIdx.x / g.tiles_w) * g.TH, wo0 = (blockIdx.x % g.tiles_w) * TW;
  const int cbase = blockIdx.y * CC;
  const int n = blockIdx.z;
  const int pair = tid % PAIRS, patch = tid / PAIRS;
  const int py = patch / PPR, px = patch % PPR;
  const int c = cbase + 2 * pair;
  const bool ch_ok = c < Cp;
  const int hi0 = M::in_org(ho0), wi0 = M::in_org(wo0);
  const int nT = g.T;
  const uint32_t plane_bytes = (uint32_t)(IH * ROW * sizeof(T));

  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < NSTAGE; ++k) mbar_init(&full_bar[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  x3d::pdl_wait();                         // on-chip setup done; global memory is touched only from here on
  const uint32_t aux_bytes = AUXT ? (uint32_t)(g.TH * TW * CC * sizeof(T)) : 0u;
  auto issue = [&](int t, int slot) {      // one thread: arm the barrier, fire the TMA box load(s) of plane t
    mbar_expect_tx(&full_bar[slot], plane_bytes + aux_bytes);
    tma_load_5d(sbuf + slot * g.stage_elems, &tmap, &full_bar[slot], cbase, wi0, hi0, t, n);
    if (AUXT) tma_load_5d(abuf + slot * g.aux_stage_elems, &amap, &full_bar[slot], cbase, wo0, ho0, t, n);
  };
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < LOOK; ++k)
      if (k < nT) issue(k, k);
  }

  // ---- per-thread constants ----------------------------------------------------------------
  float2 wreg[27];
#pragma unroll
  for (int tap = 0; tap < 27; ++tap)
    wreg[tap] = ch_ok ? *reinterpret_cast<const float2*>(w + (int64_t)tap * Cp + c) : make_float2(0.f, 0.f);
  float2 sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f);
  if ((XF || AUX) && ch_ok) {
    const int b = n % splits;
    sc = *reinterpret_cast<const float2*>(scale + (int64_t)b * Cp + c);
    sh = *reinterpret_cast<const float2*>(shift + (int64_t)b * Cp + c);
  }
  const int win_base = (M::win_off(py * PH) * IW + M::win_off(px * PW)) * CC + 2 * pair;

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
  const int64_t out0 = ((int64_t)n * nT * g.Ho + ho_t) * (int64_t)orow + (int64_t)wo_t * Cp + c;   // plane 0
  T* yp = y + out0;
  const T* ap = AUX ? aux + out0 : nullptr;
  RawPair<T> av[NO];                       // aux values (raw) of the plane stored next (dgrad epilogue, register path)
  const T* aslot = abuf;                   // AUXT: aux tile of the plane stored next
  const int aux_base = ((py * PH) * TW + px * PW) * CC + 2 * pair;

  // store a finished accumulator as the next output plane (planes are finished in order 0,1,2,...)
  auto store_plane = [&](float2 (&a)[NO]) {
#pragma unroll
    for (int oy = 0; oy < PH; ++oy) {
#pragma unroll
      for (int ox = 0; ox < PW; ++ox) {
        const int o = oy * PW + ox;
        const uint32_t ok = omask & (1u << o);
        float2 v = a[o];
        float2 ax = make_float2(0.f, 0.f);
        if (AUX) {                          // d = dgrad * [relu(bn(aux)) > 0]
          ax = AUXT ? lds_pair<T>(aslot + aux_base + (oy * TW + ox) * CC) : unpack_raw(av[o]);
          const float2 t = __ffma2_rn(ax, sc, sh);
          v.x = t.x > 0.f ? v.x : 0.f;
          v.y = t.y > 0.f ? v.y : 0.f;
        }
        float2 r = st_pair_if<T>(yp + oy * orow + ox * Cp, v, ok);
        if (!ok) r = make_float2(0.f, 0.f);
        ssum.x += r.x; ssum.y += r.y;
        ssq = __ffma2_rn(r, AUX ? ax : r, ssq);
        a[o] = make_float2(0.f, 0.f);
      }
    }
    yp += out_plane;
  };
  auto load_aux = [&]() {
    if (AUX && !AUXT) {
#pragma unroll
      for (int oy = 0; oy < PH; ++oy)
#pragma unroll
        for (int ox = 0; ox < PW; ++ox)
          av[oy * PW + ox] = ldg_raw_if(ap + oy * orow + ox * Cp, omask & (1u << (oy * PW + ox)));
      ap += out_plane;
    }
  };

  // One plane: input plane `tin` feeds output planes tin+1 (accumulator `nw`), tin (`md`) and tin-1
  // (`od`), which is complete afterwards and is stored.
  int slot = 0;
  uint32_t parity = 0;
  auto plane = [&](int tin, float2 (&nw)[NO], float2 (&md)[NO], float2 (&od)[NO]) {
    mbar_wait(&full_bar[slot], parity);    // TMA bytes of plane tin have landed
    __syncthreads();                       // everybody is done with plane tin-1: its ring slot is free
    if (tid == 0) {
      const int tn = tin + LOOK;
      int sn = slot + LOOK;
      if (sn >= NSTAGE) sn -= NSTAGE;
      if (tn < nT) issue(tn, sn);
    }
    if (tin >= 1) load_aux();              // aux of output plane tin-1, consumed after the FMA phase
    const T* bp = sbuf + slot * g.stage_elems + win_base;
    if (AUXT) {                            // aux of output plane tin-1 sits in the previous slot (still intact)
      int ps = slot - 1;
      if (ps < 0) ps += NSTAGE;
      aslot = abuf + ps * g.aux_stage_elems;
    }
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
          const int kh = M::tap(r, oy);
          if (kh < 0 || kh > 2) continue;
#pragma unroll
          for (int ox = 0; ox < PW; ++ox) {
            const int kw = M::tap(cc, ox);
            if (kw < 0 || kw > 2) continue;
            const int o = oy * PW + ox, tap = kh * 3 + kw;
            nw[o] = __ffma2_rn(wreg[M::KT_NEW * 9 + tap], xv, nw[o]);
            md[o] = __ffma2_rn(wreg[M::KT_MID * 9 + tap], xv, md[o]);
            od[o] = __ffma2_rn(wreg[M::KT_OLD * 9 + tap], xv, od[o]);
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
  auto final_aux = [&]() {                 // AUXT: aux[T-1] arrived with input plane T-1 = the slot before `slot`
    if (AUXT) {
      int ps = slot - 1;
      if (ps < 0) ps += NSTAGE;
      aslot = abuf + ps * g.aux_stage_elems;
    }
  };
  const int rem = nT - tin;
  if (rem == 0) {
    load_aux();
    final_aux();
    store_plane(accC);                     // last call was plane(.., accB, accC, accA): md = accC
  } else if (rem == 1) {
    plane(tin, accA, accB, accC);
    load_aux();
    final_aux();
    store_plane(accB);
  } else {
    plane(tin, accA, accB, accC);
    plane(tin + 1, accC, accA, accB);
    load_aux();
    final_aux();
    store_plane(accA);
  }

  if (stats != nullptr) {
    // block reduction in a fixed order within the CTA (no shared atomics): every thread parks its four partial sums
    // in its patch's row of a scratch array carved from the (now idle) plane ring, then one thread per (channel,
    // quantity) adds the rows and issues one fp64 atomic
    __syncthreads();                       // every plane of the ring has been consumed
    float* scr = reinterpret_cast<float*>(smem_raw);      // [patches][CC][2] <= 8 * 72 * 2 floats, fits one ring slot
    float* mine = scr + patch * (CC * 2) + (2 * pair) * 2;
    *reinterpret_cast<float4*>(mine) = make_float4(ssum.x, ssq.x, ssum.y, ssq.y);
    __syncthreads();
    const int patches = nthr / PAIRS;
    for (int i = tid; i < CC * 2; i += nthr) {
      const int ch = cbase + i / 2;
      float v = 0.f;
      for (int q = 0; q < patches; ++q) v += scr[q * (CC * 2) + i];
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

template <int MODE>
constexpr int tw_small() { return (MODE == M_FWD2) ? 4 : 8; }
template <int MODE>
constexpr int tw_tiny() { return 4; }   // narrow images (W = 4, 5, 10, 20 of the 112/158-pixel multigrid shapes)
template <int MODE>
constexpr int tw_big() { return (MODE == M_FWD2) ? 8 : 16; }

// (Ho, Wo): extents of the produced tensor
template <typename T, int MODE>
TilePlan plan_tiles(int64_t N, int T_, int Ho, int Wo, int Cp, int PW, bool persistent = false) {
  using M = Map<MODE>;
  TilePlan p;
  p.ok = false;
  p.PW = PW;
  const int MAX_THREADS = PW == 4 ? 384 : 256;
  TileGeom& g = p.g;
  g.T = T_; g.Cp = Cp; g.Ho = Ho; g.Wo = Wo;
  const int esz = (int)sizeof(T);
  // Pick (TH, TW, CC).  The stride-2 forward / dgrad mappings: maximise the fraction of useful lanes
  // (ragged tiles, partial channel chunks) with a mild penalty for halo re-reads and for small CTAs.  Stride-1 forward /
  // dgrad (one CTA per tile): minimise the PADDED work  waves x resident CTAs per SM x tile size  -- ragged tiles, partial
  // channel chunks and, above all, wave quantisation (448 CTAs of 224 threads on 296 slots run as 2 waves at 76 %).
  // Fitted to a sweep over all tiles on the X3D-M shapes (tools/dw_tile_sweep.py, profiles/r02_dw_tile_sweep.json): it
  // picks the measured best tile or one within 2 % of it on every layer; 28^2 x 108: 62 -> 50 us, 14^2 x 216: 37 -> 33 us.
  double best = -1.0;
  int fTH = 0, fTW = 0, fCC = 0;            // tuning knob: X3D_DW_FORCE="TH,TW,CC" pins the tile (tools/dw_tile_sweep.py)
  if (const char* f = getenv("X3D_DW_FORCE")) sscanf(f, "%d,%d,%d", &fTH, &fTW, &fCC);
  const bool wave_model = !persistent && (MODE == M_FWD1 || MODE == M_DG1);
  double best_cost = 1e300, best_useful = 0.0;
  for (int ci = 0; ci < 3; ++ci) {
    const int CC = kCC[ci];
    const int nchunk = (Cp + CC - 1) / CC;
    for (int twi = 0; twi < 3; ++twi) {
      const int TW = twi == 2 ? tw_tiny<MODE>() : twi ? tw_big<MODE>() : tw_small<MODE>();
      if (twi == 2 && TW == tw_small<MODE>()) continue;
      if (TW % PW) continue;
      for (int TH = PH; TH <= 8; TH += PH) {
        const int patches = (TH / PH) * (TW / PW);
        const int threads = patches * (CC / 2);
        if (fTH && (TH != fTH || TW != fTW || CC != fCC)) continue;
        if ((threads < 96 && !fTH) || threads > MAX_THREADS) continue;
        const int th = (Ho + TH - 1) / TH, tw = (Wo + TW - 1) / TW;
        const int IH = M::in_ext(TH), IW = M::in_ext(TW);
        if (IH > 256 || IW > 256) continue;
        const size_t smem = (size_t)NSTAGE * IH * IW * CC * esz;
        if (smem > 160 * 1024) continue;
        const double useful = (double)Ho * Wo * Cp / ((double)th * TH * tw * TW * nchunk * CC);
        if (wave_model) {
          const size_t smem_cta = (size_t)(NSTAGE + 1) * IH * IW * CC * esz + 4 * (size_t)TH * TW * CC * esz + 1024;
          int per_sm = 2048 / threads;
          const int by_regs = 65536 / (threads * 128), by_smem = (int)((220 * 1024) / smem_cta);
          if (per_sm > by_regs) per_sm = by_regs;
          if (per_sm > by_smem) per_sm = by_smem;
          if (per_sm < 1) per_sm = 1;
          const double ctas = (double)th * tw * nchunk * (double)N;
          const double waves = ceil(ctas / (double)(kNumSMs * per_sm));
          const double cost = waves * per_sm * (double)(TH * TW * CC);
          if (cost < best_cost || (cost == best_cost && useful > best_useful)) {
            best_cost = cost; best_useful = useful; best = useful;
            g.TH = TH; p.TW = TW; p.CC = CC; p.threads = threads;
          }
          continue;
        }
        if (persistent) {
          // weight gradient: CTAs walk (sample, tile) units in rounds; padded work = rounds x (tile + per-unit pipeline
          // restart, ~2000 output-channel elements), same sweep (28^2 x 108 stride 2: 107 -> 72 us)
          const double units = (double)th * tw * (double)N;
          double gx = floor((2.0 * kNumSMs) / nchunk);
          if (gx < 1.0) gx = 1.0;
          if (gx > units) gx = units;
          const double cost = ceil(units / gx) * ((double)(TH * TW * CC) + 2000.0);
          if (cost < best_cost || (cost == best_cost && useful > best_useful)) {
            best_cost = cost; best_useful = useful; best = useful;
            g.TH = TH; p.TW = TW; p.CC = CC; p.threads = threads;
          }
          continue;
        }
        const double in_per_out = MODE == M_FWD2 ? 4.0 : MODE == M_DG2 ? 0.25 : 1.0;
        const double halo = (double)IH * IW / ((double)TH * TW * in_per_out);
        const double score = useful / (1.0 + 0.3 * (halo - 1.0)) * (threads >= 192 ? 1.0 : 0.9);
        if (score > best) {
          best = score;
          g.TH = TH; p.TW = TW; p.CC = CC; p.threads = threads;
        }
      }
    }
  }
  if (best < 0.2) return p;     // too wasteful (exotic channel counts): let the direct kernel do it
  g.tiles_w = (Wo + p.TW - 1) / p.TW;
  const int tiles_h = (Ho + g.TH - 1) / g.TH;
  p.IH = M::in_ext(g.TH);
  p.IW = M::in_ext(p.TW);
  const size_t plane_bytes = (size_t)p.IH * p.IW * p.CC * esz;
  const size_t stage_bytes = (plane_bytes + 127) / 128 * 128;
  g.stage_elems = (int)(stage_bytes / esz);
  // stride-2 dgrad with the mask epilogue: 4 ring slots, each with an aux tile [TH][TW][CC] (sized for that case always)
  const size_t aux_stage_bytes = MODE >= M_DG1 ? ((size_t)g.TH * p.TW * p.CC * esz + 127) / 128 * 128 : 0;
  g.aux_stage_elems = (int)(aux_stage_bytes / esz);
  p.smem = (NSTAGE + 1) * stage_bytes + 4 * aux_stage_bytes + (size_t)p.CC * 2 * sizeof(float) +
           (NSTAGE + 1) * sizeof(uint64_t) + 16;
  p.grid = dim3((unsigned)(g.tiles_w * tiles_h), (unsigned)((Cp + p.CC - 1) / p.CC), (unsigned)N);
  if (N > 65535 || (int64_t)Ho * Wo * Cp >= (1ll << 31)) return p;
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

// NDHWC tensor [N][T_][H][W][Cp] staged in boxes of [IH][IW][CC]
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

using Args = x3d::DwTiledArgs;

template <typename T, int MODE, int CC, int TW, int PW, bool XF, bool AUX>
void launch_one(const TilePlan& p, const CUtensorMap& map, const CUtensorMap& amap, const Args& a, cudaStream_t stream) {
  auto kfn = dw3_tiled_kernel<T, MODE, CC, TW, PW, XF, AUX>;
  static unsigned long long attr_mask = 0;   // per instantiation, one bit per device
  if (first_use_on_device(&attr_mask)) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  x3d::launch(kfn, p.grid, p.threads, p.smem, stream, map, amap, a.w, (T*)a.y, p.g, a.scale, a.shift, a.splits, (const T*)a.aux,
                                             a.stats);
}

template <typename T, int MODE, int PW>
int launch_tiled(const TilePlan& p, const CUtensorMap& map, const CUtensorMap& amap, const Args& a, cudaStream_t stream) {
  constexpr bool DG = MODE >= 2;
  constexpr int TWa = tw_small<MODE>(), TWb = tw_big<MODE>(), TWc = tw_tiny<MODE>();
  const bool flag = DG ? (a.aux != nullptr) : (a.scale != nullptr);
#define L2_(CCv, TWv)                                                               \
  do {                                                                              \
    if (flag) launch_one<T, MODE, CCv, TWv, PW, !DG, DG>(p, map, amap, a, stream);  \
    else launch_one<T, MODE, CCv, TWv, PW, false, false>(p, map, amap, a, stream);  \
  } while (0)
#define L_(CCv)                                    \
  if (p.CC == CCv) {                               \
    if (p.TW == TWa) L2_(CCv, TWa);                \
    else if (p.TW == TWb) L2_(CCv, TWb);           \
    else L2_(CCv, TWc);                            \
    return 0;                                      \
  }
  L_(48) L_(56) L_(72)
#undef L_
#undef L2_
  return -1;
}

// in: staged tensor [N][T_][Hin][Win][Cp]; produced tensor has extents (Ho, Wo)
template <typename T, int MODE>
int run_tiled(const void* in, int64_t N, int T_, int Hin, int Win, int Ho, int Wo, int Cp, const Args& a,
              cudaStream_t stream, int PW, bool nan_fill, bool* handled) {
  TilePlan p = plan_tiles<T, MODE>(N, T_, Ho, Wo, Cp, PW);
  if (!p.ok) return 0;
  CUtensorMap map;
  if (!make_input_map<T>(&map, in, N, T_, Hin, Win, Cp, p.CC, p.IW, p.IH, nan_fill)) return 0;
  CUtensorMap amap = map;                  // only dereferenced by the stride-2 dgrad with the mask epilogue
  if (MODE >= M_DG1 && a.aux != nullptr &&
      !make_input_map<T>(&amap, a.aux, N, T_, Ho, Wo, Cp, p.CC, p.TW, p.g.TH, false))
    return 0;
  const int rc = launch_tiled<T, MODE, 2>(p, map, amap, a, stream);   // PW = 4 (1 CTA/SM) measured slower
  if (rc != 0) return 0;
  *handled = true;
  return 0;
}

// =================================================================================================
// wgrad:  dW[c][kt][kh][kw] += sum_{n,t,ho,wo} dy[n,t,ho,wo,c] * xf[n, t+kt-1, s*ho+kh-1, s*wo+kw-1, c]
// Same CTA / thread decomposition as the forward kernel.  Per T step the ring delivers the input-plane
// tile x[tin] (with halo, BN+ReLU fused on load, NaN -> 0 padding) and the dy tile of plane tin+1; a
// thread keeps its dy patch of planes tin+1 / tin / tin-1 in registers (they rotate like the forward
// accumulators) and its 27 x 2 weight-gradient accumulators: every window word feeds up to 27 FFMA2.
// Reduction: ordered sum over the patches of the CTA through shared scratch, then one fp32 red per (channel, tap) per CTA
// (the CTAs are persistent over (sample, tile) units, so this happens once per CTA).
// =================================================================================================
template <typename T, int MODE, int CC, int TW, int PW, bool XF>
__global__ void __launch_bounds__(Cfg<PW>::MAXT, Cfg<PW>::MINB)
dw3_wgrad_tiled_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap dymap,
                       float* __restrict__ dw, const TileGeom g, const float* __restrict__ scale,
                       const float* __restrict__ shift, int splits, int C, int dy_stage_elems, int nunits) {
  x3d::pdl_trigger();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using M = Map<MODE>;                                           // forward geometry (MODE = M_FWD1 / M_FWD2)
  constexpr int WR = M::in_ext(PH), WC = M::in_ext(PW);
  constexpr int IW = M::in_ext(TW);
  constexpr int ROW = IW * CC;
  constexpr int PAIRS = CC / 2;
  constexpr int PPR = TW / PW;
  constexpr int NO = PH * PW;
  const int IH = M::in_ext(g.TH);
  const int Cp = g.Cp;
  T* const xbuf = reinterpret_cast<T*>(smem_raw);                            // NSTAGE x tiles
  T* const dbuf = xbuf + NSTAGE * g.stage_elems;                              // NSTAGE dy tiles
  float* s_red = reinterpret_cast<float*>(dbuf + NSTAGE * dy_stage_elems);    // [27][CC]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_red + 27 * CC);

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int cbase = blockIdx.y * CC;
  const int pair = tid % PAIRS, patch = tid / PAIRS;
  const int py = patch / PPR, px = patch % PPR;
  const int c = cbase + 2 * pair;
  const bool ch_ok = c < Cp;
  const int nT = g.T;
  // PERSISTENT over work units (sample n, spatial tile): the 27 x 2 weight-gradient accumulators live in registers
  // across units, so the block reduction and the global fp32 reds (all CTAs of a channel chunk hit the same
  // C x 27 addresses, and same-address reds serialise in L2) happen once per CTA instead of once per tile.
  const int tiles_per_n = nunits / g.tiles_n;                    // spatial tiles per sample
  int n = 0, ho0 = 0, wo0 = 0, hi0 = 0, wi0 = 0;
  const uint32_t x_bytes = (uint32_t)(IH * ROW * sizeof(T));
  const uint32_t dy_bytes = (uint32_t)(g.TH * TW * CC * sizeof(T));

  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < NSTAGE; ++k) mbar_init(&full_bar[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  x3d::pdl_wait();
  // step s (s = -1 .. T-1) needs x[s] (if s >= 0) and dy[s+1] (if s+1 < T)
  auto issue = [&](int s, int slot) {
    const bool hx = s >= 0, hd = s + 1 < nT;
    mbar_expect_tx(&full_bar[slot], (hx ? x_bytes : 0u) + (hd ? dy_bytes : 0u));
    if (hx) tma_load_5d(xbuf + slot * g.stage_elems, &xmap, &full_bar[slot], cbase, wi0, hi0, s, n);
    if (hd) tma_load_5d(dbuf + slot * dy_stage_elems, &dymap, &full_bar[slot], cbase, wo0, ho0, s + 1, n);
  };
  float2 sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f);
  const int win_base = (M::win_off(py * PH) * IW + M::win_off(px * PW)) * CC + 2 * pair;
  const int dy_base = ((py * PH) * TW + px * PW) * CC + 2 * pair;

  float2 gacc[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) gacc[k] = make_float2(0.f, 0.f);
  float2 dyA[NO], dyB[NO], dyC[NO];

  int slot = 0;
  uint32_t parity = 0;
  // step s: `nw` receives dy plane s+1; with x plane s:  kt=0 pairs with nw (plane s+1), kt=1 with md
  // (plane s), kt=2 with od (plane s-1)
  auto step = [&](int s, float2 (&nw)[NO], const float2 (&md)[NO], const float2 (&od)[NO]) {
    mbar_wait(&full_bar[slot], parity);
    __syncthreads();
    if (tid == 0) {
      const int sn_step = s + NSTAGE - 1;
      int sn = slot + NSTAGE - 1;
      if (sn >= NSTAGE) sn -= NSTAGE;
      if (sn_step < nT) issue(sn_step, sn);
    }
    const T* bp = xbuf + slot * g.stage_elems + win_base;
    const T* dp = dbuf + slot * dy_stage_elems + dy_base;
    if (++slot == NSTAGE) { slot = 0; parity ^= 1u; }
    if (s + 1 < nT) {
#pragma unroll
      for (int oy = 0; oy < PH; ++oy)
#pragma unroll
        for (int ox = 0; ox < PW; ++ox) nw[oy * PW + ox] = lds_pair<T>(dp + (oy * TW + ox) * CC);
    } else {
#pragma unroll
      for (int o = 0; o < NO; ++o) nw[o] = make_float2(0.f, 0.f);
    }
    if (s < 0) return;
#pragma unroll
    for (int r = 0; r < WR; ++r) {
#pragma unroll
      for (int cc = 0; cc < WC; ++cc) {
        float2 xv = lds_pair<T>(bp + r * ROW + cc * CC);
        if (XF) {
          xv = __ffma2_rn(xv, sc, sh);
          xv.x = fmaxf(xv.x, 0.f);
          xv.y = fmaxf(xv.y, 0.f);
        }
#pragma unroll
        for (int oy = 0; oy < PH; ++oy) {
          const int kh = M::tap(r, oy);
          if (kh < 0 || kh > 2) continue;
#pragma unroll
          for (int ox = 0; ox < PW; ++ox) {
            const int kw = M::tap(cc, ox);
            if (kw < 0 || kw > 2) continue;
            const int o = oy * PW + ox, tap = kh * 3 + kw;
            gacc[tap] = __ffma2_rn(nw[o], xv, gacc[tap]);
            gacc[9 + tap] = __ffma2_rn(md[o], xv, gacc[9 + tap]);
            gacc[18 + tap] = __ffma2_rn(od[o], xv, gacc[18 + tap]);
          }
        }
      }
    }
  };

  for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    n = unit / tiles_per_n;
    const int tile = unit - n * tiles_per_n;
    ho0 = (tile / g.tiles_w) * g.TH;
    wo0 = (tile % g.tiles_w) * TW;
    hi0 = M::in_org(ho0);
    wi0 = M::in_org(wo0);
    // the ring slots the first loads of this unit go to were last read two / three steps ago (all threads passed
    // the barrier of the previous unit's final step since), so the pipeline restarts without another barrier
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < NSTAGE - 1; ++k) {
        int sl = slot + k;
        if (sl >= NSTAGE) sl -= NSTAGE;
        if (k - 1 < nT) issue(k - 1, sl);
      }
    }
    if (XF && ch_ok) {
      const int b = n % splits;
      sc = *reinterpret_cast<const float2*>(scale + (int64_t)b * Cp + c);
      sh = *reinterpret_cast<const float2*>(shift + (int64_t)b * Cp + c);
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) dyA[o] = dyB[o] = dyC[o] = make_float2(0.f, 0.f);
    // steps -1, 0, ..., T-1 with the dy register sets rotating (A,B,C) -> (C,A,B) -> (B,C,A)
    int s = -1;
    for (; s + 3 <= nT; s += 3) {
      step(s, dyA, dyB, dyC);
      step(s + 1, dyC, dyA, dyB);
      step(s + 2, dyB, dyC, dyA);
    }
    if (s < nT) {
      step(s, dyA, dyB, dyC);
      if (s + 1 < nT) step(s + 1, dyC, dyA, dyB);
    }
  }

  // ---- reduction: patches of the CTA -> shared, CTA -> global ------------------------------------
  // No shared atomics: in three rounds of 9 taps every thread parks its partial sums in its patch's row of a scratch
  // array carved from the (now idle) ring -- patches x 9 x CC floats = 72 bytes per thread, always smaller than the
  // ring -- and one thread per (tap, channel) adds the rows in a fixed order and issues ONE fp32 red.
  float* scr = reinterpret_cast<float*>(smem_raw);
  const int patches = nthr / PAIRS;
#pragma unroll
  for (int rnd = 0; rnd < 3; ++rnd) {
    __syncthreads();                       // ring consumed (rnd 0) / previous round summed
#pragma unroll
    for (int k = 0; k < 9; ++k)
      *reinterpret_cast<float2*>(scr + (patch * 9 + k) * CC + 2 * pair) = gacc[rnd * 9 + k];
    __syncthreads();
    for (int i = tid; i < 9 * CC; i += nthr) {
      const int k = i / CC, cl = i - k * CC, ch = cbase + cl;
      float v = 0.f;
      for (int q = 0; q < patches; ++q) v += scr[(q * 9 + k) * CC + cl];
      if (ch < C && v != 0.f) atomicAdd(&dw[(int64_t)ch * 27 + rnd * 9 + k], v);
    }
  }
}

template <typename T, int MODE, int CC, int TW, bool XF>
void launch_wgrad_one(const TilePlan& p, const CUtensorMap& xmap, const CUtensorMap& dymap, float* dw,
                      const float* scale, const float* shift, int splits, int C, int dy_stage_elems, size_t smem,
                      cudaStream_t stream) {
  auto kfn = dw3_wgrad_tiled_kernel<T, MODE, CC, TW, 2, XF>;
  static unsigned long long attr_mask = 0;   // per instantiation, one bit per device
  if (first_use_on_device(&attr_mask)) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  // persistent grid: as many CTAs per channel chunk as fit the 2 x 148 resident slots, each walking
  // ceil(units / CTAs) (sample, tile) units
  const int nunits = (int)(p.grid.x * p.grid.z);
  const int chunks = (int)p.grid.y;
  int gx = (2 * kNumSMs) / chunks;
  if (gx < 1) gx = 1;
  if (gx > nunits) gx = nunits;
  const int per = (nunits + gx - 1) / gx;
  gx = (nunits + per - 1) / per;                       // same number of rounds, no idle CTAs
  TileGeom g = p.g;
  g.tiles_n = (int)p.grid.z;
  x3d::launch(kfn, dim3((unsigned)gx, (unsigned)chunks, 1u), p.threads, smem, stream, xmap, dymap, dw, g, scale, shift, splits,
              C, dy_stage_elems, nunits);
}

// x: [N][T_][H][W][Cp] (conv input), dy: [N][T_][Ho][Wo][Cp]
template <typename T, int MODE>
int run_wgrad_tiled(const void* x, const void* dy, float* dw, int64_t N, int T_, int H, int W, int Ho, int Wo, int C,
                    int Cp, const float* scale, const float* shift, int splits, cudaStream_t stream, bool* handled) {
  TilePlan p = plan_tiles<T, MODE>(N, T_, Ho, Wo, Cp, 2, true);
  if (!p.ok) return 0;
  const size_t esz = sizeof(T);
  const size_t dy_stage_bytes = ((size_t)p.g.TH * p.TW * p.CC * esz + 127) / 128 * 128;
  const size_t smem = NSTAGE * ((size_t)p.g.stage_elems * esz + dy_stage_bytes) + (size_t)27 * p.CC * sizeof(float) +
                      NSTAGE * sizeof(uint64_t) + 16;
  if (smem > 200 * 1024) return 0;
  CUtensorMap xmap, dymap;
  if (!make_input_map<T>(&xmap, x, N, T_, H, W, Cp, p.CC, p.IW, p.IH, scale != nullptr)) return 0;
  if (!make_input_map<T>(&dymap, dy, N, T_, Ho, Wo, Cp, p.CC, p.TW, p.g.TH, false)) return 0;
  constexpr int TWa = tw_small<MODE>(), TWb = tw_big<MODE>(), TWc = tw_tiny<MODE>();
  const bool xf = scale != nullptr;
  const int dse = (int)(dy_stage_bytes / esz);
#define W2_(CCv, TWv)                                                                                             \
  do {                                                                                                            \
    if (xf) launch_wgrad_one<T, MODE, CCv, TWv, true>(p, xmap, dymap, dw, scale, shift, splits, C, dse, smem, stream);   \
    else launch_wgrad_one<T, MODE, CCv, TWv, false>(p, xmap, dymap, dw, scale, shift, splits, C, dse, smem, stream);    \
  } while (0)
#define W_(CCv)                                    \
  if (p.CC == CCv) {                               \
    if (p.TW == TWa) W2_(CCv, TWa);                \
    else if (p.TW == TWb) W2_(CCv, TWb);           \
    else W2_(CCv, TWc);                            \
    *handled = true;                               \
    return 0;                                      \
  }
  W_(48) W_(56) W_(72)
#undef W_
#undef W2_
  return 0;
}

}  // namespace

namespace x3d {
int DW_WGRAD_ENTRY(int stride, const void* x, const void* dy, float* dw, int64_t N, int T_, int H, int W, int Ho,
                   int Wo, int C, int Cp, const float* scale, const float* shift, int splits, cudaStream_t stream,
                   bool* handled) {
  if (stride == 1)
    return run_wgrad_tiled<DW_T, M_FWD1>(x, dy, dw, N, T_, H, W, Ho, Wo, C, Cp, scale, shift, splits, stream, handled);
  return run_wgrad_tiled<DW_T, M_FWD2>(x, dy, dw, N, T_, H, W, Ho, Wo, C, Cp, scale, shift, splits, stream, handled);
}

// one translation unit per storage type (DW_T): halves the build time of the 48 kernel instantiations
int DW_TILED_ENTRY(int mode, const void* in, int64_t N, int T_, int Hin, int Win, int Ho, int Wo, int Cp,
                   const DwTiledArgs& a, cudaStream_t stream, bool nan_fill, bool* handled) {
  const int PW = 2;
  switch (mode) {
    case M_FWD1: return run_tiled<DW_T, M_FWD1>(in, N, T_, Hin, Win, Ho, Wo, Cp, a, stream, PW, nan_fill, handled);
    case M_FWD2: return run_tiled<DW_T, M_FWD2>(in, N, T_, Hin, Win, Ho, Wo, Cp, a, stream, PW, nan_fill, handled);
    case M_DG1: return run_tiled<DW_T, M_DG1>(in, N, T_, Hin, Win, Ho, Wo, Cp, a, stream, PW, nan_fill, handled);
    default: return run_tiled<DW_T, M_DG2>(in, N, T_, Hin, Win, Ho, Wo, Cp, a, stream, PW, nan_fill, handled);
  }
}
}  // namespace x3d
