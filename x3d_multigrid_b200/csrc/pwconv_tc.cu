// Pointwise (1x1x1) Conv3d forward / dgrad as a tcgen05 GEMM (bf16 in, fp32 accumulate in TMEM).
//
//   C[m][n] = sum_k A[m][k] * B[n][k]      A: activations [M][Kp] (NDHWC rows), B: weights [Np][Kp]
//
// Persistent CTAs (2-3 per SM), each walking a contiguous range of 128-row tiles of one BN <= 128 wide column part;
// 6 warps:
//   warp 0  : producer -- TMA boxes of A (128 x 64) and B (BN x 64) per 64-wide K chunk into a 2-3 stage shared
//             ring (SWIZZLE_128B, "full" mbarriers); when the A rows are narrower than 128 bytes (K < 64) the
//             whole warp copies the tile with coalesced 16-byte cp.async into the same swizzled layout instead
//             (TMA retires only ~1 box row per 20 cycles) and B stays resident; strided (downsample) convs gather
//             their input rows here;
//   warp 1  : allocates TMEM, then one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//             (M=128, N=BN, K=16) per 16-wide K step into one of two TMEM accumulators, tcgen05.commit frees the
//             ring slot ("empty") and finally signals the accumulator ("accum_full") to the epilogue;
//   warps 2-5: epilogue -- tcgen05.ld of their 32 TMEM lanes (= 32 output rows), bf16 pack into a padded shared
//             staging tile, coalesced 16-byte global stores (optionally scattered / accumulated: dgrad of a
//             strided conv), and (forward) the per-sample column sums / sums of squares that feed the following
//             SubBatchNorm3d, kept in registers across the CTA's tiles and flushed once per sample.
// K and N of this network are small (24..432) and M is huge: every layer is HBM-bound,
// so the design goal is simply to stream A and C at full bandwidth with several CTAs in flight per SM.
#include <cuda.h>

#include "common.cuh"

using namespace x3d;

namespace {

constexpr int BM = 128, BK = 64;
constexpr int NTHREADS = 192;

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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);          // start address
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                            // layout: SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format  = F32
  d |= 1u << 7;                       // a_format  = BF16
  d |= 1u << 10;                      // b_format  = BF16
  d |= (uint32_t)(n >> 3) << 17;      // n_dim
  d |= (uint32_t)(BM >> 4) << 24;     // m_dim
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// two 16-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr, uint32_t (&a)[16], uint32_t (&b)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]),
        "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
      : "r"(taddr));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]),
        "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
      : "r"(taddr + 16));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

struct TcParams {
  int M, Kp, Np, BN, nk, tmem_cols, ldc;
  int tiles_m;       // number of 128-row tiles (CTAs stride over them)
  int stages;        // operand ring depth
  int cpitch;        // staging row pitch in elements (BN + 8: conflict-free 16-byte st.shared)
  int cpa;           // 1: A tile copied by cp.async (narrow rows), 0: by TMA
  int gs, gH, gW, gHo, gWo;   // row gather of a strided 1x1x1 conv: output row (nt,ho,wo) reads input row (nt, gs*ho, gs*wo)
  int ss;            // row SCATTER (dgrad of a strided conv): GEMM row (nt,ho,wo) is written to row (nt, ss*ho, ss*wo)
                     // of C (same gH,gW,gHo,gWo fields); ss = 0/1: dense
  int accum;         // C += instead of C = (scatter mode: dx already holds the main-branch gradient)
  int dbg;           // profiling experiments (X3D_TC_DBG): 1 = skip copy-out, 2 = skip statistics, 4 = skip staging
  long long P_out;
};


// Persistent kernel: grid = (min(tiles_m, resident CTAs), column parts).  Each CTA walks its 128-row tiles;
// the three roles run as independent pipelines connected by mbarriers:
//   full/empty[stage]      TMA producer  <-> MMA issuer      (operand ring)
//   accum_full/empty[buf]  MMA issuer    <-> epilogue warps  (two TMEM accumulators: tile i+1 is being
//                                                              multiplied while tile i is drained)
template <bool STATS>
__global__ void __launch_bounds__(NTHREADS)
pw_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __nv_bfloat16* __restrict__ A, __nv_bfloat16* __restrict__ C, const TcParams p,
             double* __restrict__ stats) {
  x3d::pdl_trigger();
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte aligned operand ring (SWIZZLE_128B atoms)
  const uint32_t base_u32 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base = smem_dyn + (base_u32 - smem_u32(smem_dyn));
  const int BN = p.BN, S = p.stages;
  const uint32_t a_bytes = BM * BK * 2, b_bytes = (uint32_t)BN * BK * 2;
  unsigned char* a_s = base;
  unsigned char* b_s = base + S * a_bytes;
  // (cp.async path: the weights are resident in ONE B slot, the ring holds A tiles only)
  __nv_bfloat16* c_s = reinterpret_cast<__nv_bfloat16*>(b_s + (p.cpa ? 1 : S) * b_bytes);   // [128][cpitch] staging
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(c_s + BM * p.cpitch);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* accum_full = empty_bar + S;
  uint64_t* accum_empty = accum_full + 2;
  uint64_t* b_bar = accum_empty + 2;                            // resident-weights barrier (cp.async path)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_bar + 1);
  float* s_stat = reinterpret_cast<float*>(tmem_slot + 2);      // [BN][2]
  float* s_part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_stat + 2 * BN) + 15) & ~(uintptr_t)15);
  int* s_rowoff = reinterpret_cast<int*>(s_part + 1024);        // [128] gathered input rows of the current tile
  int* s_rowout = s_rowoff + BM;                                 // [128] scattered output rows of the current tile                              // [row groups][ncols][2] partial sums (<= 4 KB)

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n0 = blockIdx.y * BN;
  const int ncols = (p.Np - n0 < BN) ? p.Np - n0 : BN;          // valid columns of this column part
  const int acc_off = p.tmem_cols / 2;                          // TMEM column offset of accumulator 1

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mapB) : "memory");
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], p.cpa ? 32 : 1);    // cp.async path: one arrival per producer lane
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&accum_full[b], 1);
      mbar_init(&accum_empty[b], 4);          // one arrival per epilogue warp
    }
    mbar_init(b_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  x3d::pdl_wait();      // everything above was on-chip setup; global memory is touched only from here on
  // contiguous range of 128-row tiles of this CTA (keeps consecutive tiles in the same sample: statistics
  // are flushed once per sample, not once per tile)
  const int tpc = (p.tiles_m + gridDim.x - 1) / gridDim.x;
  const int tile_begin = blockIdx.x * tpc;
  const int tile_end = tile_begin + tpc < p.tiles_m ? tile_begin + tpc : p.tiles_m;

  if (warp == 0) {
    if (p.cpa) {
      // ===== producer, narrow A (row pitch < 128 B): TMA issues one request per box row and crawls on
      // 48..112-byte rows, whereas the 128 x Kp tile is ONE contiguous span of global memory.  The warp
      // copies it with coalesced 16-byte cp.async, scattering every chunk to its SWIZZLE_128B position
      // (chunk c of row r -> r*128 + ((c ^ (r & 7)) << 4)), i.e. the same canonical K-major layout the
      // TMA path produces.  B (weights, L2 resident) still comes by TMA.
      const int cpr = p.Kp >> 3;                                // 16-byte chunks per row (3..7)
      for (int s = 0; s < S; ++s)                               // chunk columns never written: zero once
        for (int i = lane; i < BM * (8 - cpr); i += 32) {
          const int r = i / (8 - cpr), c = cpr + i % (8 - cpr);
          *reinterpret_cast<uint4*>(a_s + s * a_bytes + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      // the weights (one K chunk) are loaded ONCE and stay resident: TMA needs ~20 cycles per box row of
      // this width, which would dominate if repeated per tile
      if (lane == 0) {
        mbar_expect_tx(accum_empty + 2, b_bytes);
        tma_load_2d(b_s, &mapB, accum_empty + 2, 0, n0);
      }
      const unsigned char* Ab = reinterpret_cast<const unsigned char*>(A);
      // D = S-1 tiles of look-ahead: tile `it` is published (full barrier) once D younger copy groups are in flight
      const int D = S - 1;
      int it = 0;
      for (int tm = tile_begin; tm < tile_end; ++tm, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int m0 = tm * BM;
        const int rows = (p.M - m0 < BM) ? p.M - m0 : BM;
        const unsigned char* src0 = Ab + (int64_t)m0 * p.Kp * 2;
        const uint32_t dst0 = smem_u32(a_s + s * a_bytes);
        if (p.gs > 1) {
          // strided conv (downsample branch): the tile's rows are gathered; input row index per tile row first
          __syncwarp();
          for (int r = lane; r < BM; r += 32) {
            const int m = m0 + r;
            const int hw = p.gHo * p.gWo;
            const int nt = m / hw, rem = m - nt * hw;
            const int ho = rem / p.gWo, wo = rem - ho * p.gWo;
            s_rowoff[r] = (nt * p.gH + ho * p.gs) * p.gW + wo * p.gs;     // < 2^31 rows (checked on the host)
          }
          __syncwarp();
          const int dr = 32 / cpr, dc = 32 - dr * cpr;
          int r = lane / cpr, c = lane - r * cpr;
          for (; r < BM;) {
            const bool ok = r < rows;
            const unsigned char* src = Ab + (ok ? ((int64_t)s_rowoff[r] * p.Kp + c * 8) * 2 : 0);
            const uint32_t dst = dst0 + r * 128 + ((c ^ (r & 7)) << 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
            r += dr;
            c += dc;
            if (c >= cpr) { c -= cpr; ++r; }
          }
        } else {
          for (int ch = lane; ch < BM * cpr; ch += 32) {
            const int r = ch / cpr, c = ch - r * cpr;
            const bool ok = r < rows;
            const unsigned char* src = src0 + (ok ? (int64_t)ch * 16 : 0);
            const uint32_t dst = dst0 + r * 128 + ((c ^ (r & 7)) << 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        if (it >= D) {                                           // publish tile it - D (its copies are done)
          switch (D) {
            case 1: asm volatile("cp.async.wait_group 1;\n" ::: "memory"); break;
            case 2: asm volatile("cp.async.wait_group 2;\n" ::: "memory"); break;
            case 3: asm volatile("cp.async.wait_group 3;\n" ::: "memory"); break;
            case 4: asm volatile("cp.async.wait_group 4;\n" ::: "memory"); break;
            default: asm volatile("cp.async.wait_group 5;\n" ::: "memory"); break;
          }
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&full_bar[(it - D) % S])) : "memory");
        }
      }
      if (it > 0) {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        for (int j = (it > D ? it - D : 0); j < it; ++j)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&full_bar[j % S])) : "memory");
      }
    } else if (lane == 0) {
      // ===== TMA producer =====
      int it = 0;
      for (int tm = tile_begin; tm < tile_end; ++tm) {
        for (int kc = 0; kc < p.nk; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], a_bytes + b_bytes);
          tma_load_2d(a_s + s * a_bytes, &mapA, &full_bar[s], kc * BK, tm * BM);
          tma_load_2d(b_s + s * b_bytes, &mapB, &full_bar[s], kc * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BN);
      int it = 0, tcount = 0;
      if (p.cpa && tile_begin < tile_end) mbar_wait(b_bar, 0);
      for (int tm = tile_begin; tm < tile_end; ++tm, ++tcount) {
        const int buf = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        mbar_wait(&accum_empty[buf], aph ^ 1u);          // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * acc_off;
        for (int kc = 0; kc < p.nk; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const int krem = p.Kp - kc * BK;
          const int ksteps = krem >= BK ? BK / 16 : (krem + 15) / 16;
          const uint32_t a_addr = smem_u32(a_s + s * a_bytes), b_addr = smem_u32(b_s + (p.cpa ? 0 : s) * b_bytes);
          for (int k = 0; k < ksteps; ++k)
            umma_f16(tmem_d, make_smem_desc(a_addr + k * 32), make_smem_desc(b_addr + k * 32), idesc,
                     (kc | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);          // ring slot reusable once these MMAs have read it
        }
        umma_commit(&accum_full[buf]);         // accumulator complete
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> smem staging -> coalesced global stores (+ column statistics) =====
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    const int et = threadIdx.x - 64;                       // 0..127 among the epilogue threads
    const int rloc = q * 32 + lane;                        // row of the tile owned by this thread
    // statistics: every thread keeps register partial sums (sum, sum of squares of 4 columns over its row group)
    // of the sample `cur_samp`; they are combined through a small scratch array and flushed to global (fp64
    // atomics) only when the CTA's contiguous tile range moves on to another sample.
    // thread = (4-column group cq, row group grp): one 8-byte shared load per row feeds packed f32x2 adds / FMAs
    const int nquad = ncols >> 2;
    int groups = 128 / nquad;
    if (groups > 16) groups = 16;
    const int cq = et % nquad, grp = et / nquad;
    const bool worker = et < nquad * groups;
    const int pstride = ncols * 2;                                  // scratch row: [col][sum, sumsq]
    float2 s1a = make_float2(0.f, 0.f), s1b = s1a, s2a = s1a, s2b = s1a;
    long long cur_samp = -1;
    long long samp_lim = 0;                                         // first global row of the next sample
    auto flush_stats = [&]() {                                      // called by all 128 epilogue threads
      if (cur_samp >= 0) {
        if (worker) {
          float4* dst = reinterpret_cast<float4*>(s_part + grp * pstride + cq * 8);
          dst[0] = make_float4(s1a.x, s2a.x, s1a.y, s2a.y);
          dst[1] = make_float4(s1b.x, s2b.x, s1b.y, s2b.y);
        }
        asm volatile("bar.sync 2, 128;\n" ::: "memory");
        for (int i = et; i < pstride; i += 128) {
          float v = 0.f;
          for (int gq = 0; gq < groups; ++gq) v += s_part[gq * pstride + i];
          if (v != 0.f) atomicAdd(&stats[(cur_samp * p.ldc + n0 + (i >> 1)) * 2 + (i & 1)], (double)v);
        }
        asm volatile("bar.sync 2, 128;\n" ::: "memory");             // scratch reusable
        s1a = s1b = s2a = s2b = make_float2(0.f, 0.f);
      }
    };
    // copy-out: consecutive threads write consecutive 16-byte chunks of the output rows; chunk -> (row, chunk in
    // row) is advanced incrementally (no division in the loop)
    const int cpr = ncols >> 3;                              // 16-byte chunks per row
    const int co_r0 = et / cpr, co_c0 = et - co_r0 * cpr;
    const int co_dr = 128 / cpr, co_dc = 128 - co_dr * cpr;
    int tcount = 0;
    for (int tm = tile_begin; tm < tile_end; ++tm, ++tcount) {
      const int buf = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      const int m0 = tm * BM;
      // staging buffer is free (previous tile fully copied out / summed)
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      if (p.ss > 1) {                                       // scatter mode: output row of tile row `et`
        const int m = m0 + et;
        const int hw = p.gHo * p.gWo;
        const int nt = m / hw, rem = m - nt * hw;
        const int ho = rem / p.gWo, wo = rem - ho * p.gWo;
        s_rowout[et] = (nt * p.gH + ho * p.ss) * p.gW + wo * p.ss;   // published by the bar.sync after staging
      }
      mbar_wait(&accum_full[buf], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * acc_off + ((uint32_t)(q * 32) << 16);
      __nv_bfloat16* srow = c_s + rloc * p.cpitch;
      auto stage16 = [&](int c0, const uint32_t (&r)[16]) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        *reinterpret_cast<uint4*>(srow + c0) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(srow + c0 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
      };
      // only the chunks that hold real columns are read (BN is the MMA width, ncols <= BN); two loads per wait
      const int cend = (p.dbg & 4) ? 0 : ncols;
      int c0 = 0;
      for (; c0 + 16 < cend; c0 += 32) {
        uint32_t ra[16], rb[16];
        tmem_ld16x2(taddr + c0, ra, rb);
        stage16(c0, ra);
        stage16(c0 + 16, rb);
      }
      if (c0 < cend) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        stage16(c0, r);
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&accum_empty[buf])) : "memory");
      }
      asm volatile("bar.sync 1, 128;\n" ::: "memory");     // tile staged in shared memory
      const int rows_valid = (p.M - m0 < BM) ? p.M - m0 : BM;
      if (!(p.dbg & 1)) {
        int r = co_r0, cc = co_c0;
        while (r < rows_valid) {
          uint4 v = *reinterpret_cast<const uint4*>(c_s + r * p.cpitch + cc * 8);
          const int64_t orow = p.ss > 1 ? (int64_t)s_rowout[r] : (int64_t)(m0 + r);
          uint4* gp = reinterpret_cast<uint4*>(C + orow * p.ldc + n0 + cc * 8);
          if (p.accum) {
            const uint4 o = *gp;
            const uint32_t a4[4] = {v.x, v.y, v.z, v.w}, o4[4] = {o.x, o.y, o.z, o.w};
            uint32_t r4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              r4[j] = pack_bf16x2(__uint_as_float(a4[j] << 16) + __uint_as_float(o4[j] << 16),
                                  __uint_as_float(a4[j] & 0xffff0000u) + __uint_as_float(o4[j] & 0xffff0000u));
            v = make_uint4(r4[0], r4[1], r4[2], r4[3]);
          }
          *gp = v;
          r += co_dr;
          cc += co_dc;
          if (cc >= cpr) { cc -= cpr; ++r; }
        }
      }
      if (STATS && !(p.dbg & 2)) {
        // column sums of the staged (bf16-rounded) tile, one row segment per sample touched by the tile
        int r0 = 0;
        while (r0 < rows_valid) {
          const long long g0 = (long long)m0 + r0;
          if (cur_samp < 0 || g0 >= samp_lim) {                     // uniform across the 128 threads
            flush_stats();
            cur_samp = g0 / p.P_out;
            samp_lim = (cur_samp + 1) * p.P_out;
          }
          const long long lim = samp_lim - m0;
          const int r1 = lim < rows_valid ? (int)lim : rows_valid;
          if (worker) {
            const __nv_bfloat16* cp = c_s + 4 * cq;
#pragma unroll 4
            for (int r = r0 + grp; r < r1; r += groups) {
              const uint2 wv = *reinterpret_cast<const uint2*>(cp + r * p.cpitch);
              const float2 a = make_float2(__uint_as_float(wv.x << 16), __uint_as_float(wv.x & 0xffff0000u));
              const float2 b2 = make_float2(__uint_as_float(wv.y << 16), __uint_as_float(wv.y & 0xffff0000u));
              s1a = __fadd2_rn(s1a, a);
              s1b = __fadd2_rn(s1b, b2);
              s2a = __ffma2_rn(a, a, s2a);
              s2b = __ffma2_rn(b2, b2, s2b);
            }
          }
          r0 = r1;
        }
      }
    }
    if (STATS) flush_stats();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
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
// row-major bf16 matrix [rows][cols]; boxes of [box_rows][64] with 128-byte swizzle
bool make_map_2d(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// =================================================================================================
// wgrad:  dW[n][k] += sum_m dY[m][n] * X[m][k]
// Both operands have the reduction index m as their SLOW memory index, i.e. they are "MN-major" UMMA
// operands: a TMA box of 64 rows (m) x 64 columns gives exactly the canonical SWIZZLE_128B MN-major
// atom layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: SBO = 1024 B between 8-row groups,
// LBO = 8192 B between 64-column blocks.  A = dY^T (M' = 128 columns n per CTA), B = X^T (N' <= 256
// columns k), D[n][k] accumulates in TMEM over the CTA's chunk of rows, then fp32 red.global.add into dW.
// =================================================================================================
constexpr int WG_ROWS = 64;                       // m rows per ring stage (4 MMAs of K = 16)
constexpr int WG_MAX_STAGES = 8;
constexpr int WG_BOX_BYTES = WG_ROWS * 128;       // one 64 x 64 bf16 box

__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)(WG_BOX_BYTES >> 4) << 16;          // LBO: next 64-column block
  d |= (uint64_t)(1024 >> 4) << 32;                  // SBO: next group of 8 rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {
  return make_idesc(n) | (1u << 15) | (1u << 16);    // A and B are MN-major
}

struct WgParams {
  int M, K, Kp, Nn, Np;
  int NB;            // MMA N (columns k handled by this CTA's part), multiple of 16
  int a_boxes, b_boxes_full;   // 64-wide boxes per stage for A (1 or 2 valid) and for a full B part
  int rows_per_cta;
  int tmem_cols;
  int cpa;           // allow the cp.async loader for operands narrower than 64 columns
  int stages;        // operand ring depth (3..8)
  int ldp;           // two-stage mode: row pitch (floats) of the partial tiles, = NB * number of k parts
  int gs, gH, gW, gHo, gWo;   // strided (downsample) conv: row m = (nt,ho,wo) of dY pairs with row (nt, gs*ho, gs*wo) of X
                              // (gathered by the cp.async loader; gs = 1: dense)
};

__global__ void __launch_bounds__(NTHREADS)
pw_wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapX,
                   const __nv_bfloat16* __restrict__ DY, const __nv_bfloat16* __restrict__ X, float* __restrict__ dW,
                   const WgParams p, float* __restrict__ partial) {
  x3d::pdl_trigger();
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base = smem_dyn + (base_u32 - smem_u32(smem_dyn));
  const int A_BYTES = p.a_boxes * WG_BOX_BYTES;               // 64 or 128 columns n
  const int b_stage_bytes = p.b_boxes_full * WG_BOX_BYTES;    // up to 256 columns k
  const int stage_bytes = A_BYTES + b_stage_bytes;
  const int S = p.stages;                                     // ring depth (<= WG_MAX_STAGES)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + WG_MAX_STAGES * 0 + S * stage_bytes);
  uint64_t* empty_bar = full_bar + WG_MAX_STAGES;
  uint64_t* accum_bar = empty_bar + WG_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n0 = blockIdx.y * 128;
  const int k0 = blockIdx.z * p.NB;
  const int64_t row0 = (int64_t)blockIdx.x * p.rows_per_cta;
  int64_t row1 = row0 + p.rows_per_cta;
  if (row1 > p.M) row1 = p.M;
  const int nsteps = (int)((row1 - row0 + WG_ROWS - 1) / WG_ROWS);
  // boxes actually worth loading (the rest of the MMA operand is don't-care: those D rows/cols are never stored)
  const int a_boxes = (p.Np - n0 > 64) ? 2 : 1;
  int b_boxes = (p.Kp - k0 + 63) / 64;
  const int b_need = (p.NB + 63) / 64;
  if (b_boxes > b_need) b_boxes = b_need;
  const bool cpa_a = p.cpa && p.Np < 64, cpa_b = p.cpa && p.Kp < 64;
  const int full_count = (cpa_a || cpa_b) ? 160 + ((cpa_a && cpa_b) ? 0 : 1) : 1;   // 5 producer warps (+ expect_tx)

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mapDY) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mapX) : "memory");
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], full_count);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  x3d::pdl_wait();      // everything above was on-chip setup; global memory is touched only from here on

  if ((cpa_a || cpa_b) && warp != 1) {
    // narrow operand(s) (< 64 columns, i.e. < 128-byte rows): coalesced cp.async of the contiguous 64-row span
    // into the SWIZZLE_128B atom layout (see pw_tc_kernel); wide operands still use TMA boxes.  All five
    // non-MMA warps copy (the epilogue warps are idle until the accumulator is complete); the chunk -> (row,
    // column) decomposition is the same for every step and is computed once.
    constexpr int NPROD = 160, MAXQ = 8;
    const int pl = warp == 0 ? lane : 32 * (warp - 1) + lane;       // 0..159
    const unsigned char* dyb = reinterpret_cast<const unsigned char*>(DY);
    const unsigned char* xb = reinterpret_cast<const unsigned char*>(X);
    const int cpr_a = p.Np >> 3, cpr_b = p.Kp >> 3;
    const int nA = cpa_a ? WG_ROWS * cpr_a : 0, nB = cpa_b ? WG_ROWS * cpr_b : 0;
    int q_row[MAXQ], q_src[MAXQ], q_dst[MAXQ];                       // row in the step, byte offsets (src<0: unused)
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) {
      const int q = pl + i * NPROD;
      q_row[i] = 0; q_src[i] = -1; q_dst[i] = 0;
      if (q < nA) {
        const int r = q / cpr_a, c = q - r * cpr_a;
        q_row[i] = r; q_src[i] = q * 16; q_dst[i] = r * 128 + ((c ^ (r & 7)) << 4);
      } else if (q < nA + nB) {
        const int qq = q - nA;
        const int r = qq / cpr_b, c = qq - r * cpr_b;
        q_row[i] = r; q_src[i] = qq * 16 + (1 << 30); q_dst[i] = A_BYTES + r * 128 + ((c ^ (r & 7)) << 4);
      }
    }
    const int D = S - 2 > 6 ? 6 : S - 2;
    for (int it = 0; it < nsteps; ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      unsigned char* st = base + s * stage_bytes;
      const int64_t m = row0 + (int64_t)it * WG_ROWS;
      const int rows = (int)((row1 - m < WG_ROWS) ? row1 - m : WG_ROWS);
      if (pl == 0 && !(cpa_a && cpa_b)) {
        mbar_expect_tx(&full_bar[s], (uint32_t)(((cpa_a ? 0 : a_boxes) + (cpa_b ? 0 : b_boxes)) * WG_BOX_BYTES));
        if (!cpa_a)
          for (int j = 0; j < a_boxes; ++j) tma_load_2d(st + j * WG_BOX_BYTES, &mapDY, &full_bar[s], n0 + 64 * j, (int)m);
        if (!cpa_b)
          for (int j = 0; j < b_boxes; ++j)
            tma_load_2d(st + A_BYTES + j * WG_BOX_BYTES, &mapX, &full_bar[s], k0 + 64 * j, (int)m);
      }
      const unsigned char* srcA = dyb + m * p.Np * 2;
      const unsigned char* srcB = xb + m * p.Kp * 2;
      const uint32_t dst0 = smem_u32(st);
#pragma unroll
      for (int i = 0; i < MAXQ; ++i) {
        if (q_src[i] >= 0) {
          const bool isb = q_src[i] >= (1 << 30);
          const bool ok = q_row[i] < rows;
          const unsigned char* src = (isb ? srcB + (q_src[i] - (1 << 30)) : srcA + q_src[i]);
          if (isb && p.gs > 1 && ok) {
            // gathered X row of GEMM row mm (downsample branch, x3d.py:272: stride (1,s,s) on the input positions)
            const int mm = (int)m + q_row[i];
            const int hw = p.gHo * p.gWo;
            const int nt = mm / hw, rem = mm - nt * hw;
            const int ho = rem / p.gWo, wo = rem - ho * p.gWo;
            const int64_t srow = ((int64_t)nt * p.gH + ho * p.gs) * p.gW + wo * p.gs;
            const int cc = ((q_src[i] - (1 << 30)) >> 4) - q_row[i] * cpr_b;       // 16-byte chunk within the row
            src = xb + (srow * p.Kp + cc * 8) * 2;
          }
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst0 + q_dst[i]),
                       "l"(ok ? src : srcA), "r"(ok ? 16 : 0)
                       : "memory");
        }
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
      // keep D = S-2 younger groups in flight: publish the step whose copies are now guaranteed complete
      if (it >= D) {
        switch (D) {
          case 1: asm volatile("cp.async.wait_group 1;\n" ::: "memory"); break;
          case 2: asm volatile("cp.async.wait_group 2;\n" ::: "memory"); break;
          case 3: asm volatile("cp.async.wait_group 3;\n" ::: "memory"); break;
          case 4: asm volatile("cp.async.wait_group 4;\n" ::: "memory"); break;
          case 5: asm volatile("cp.async.wait_group 5;\n" ::: "memory"); break;
          default: asm volatile("cp.async.wait_group 6;\n" ::: "memory"); break;
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&full_bar[(it - D) % S])) : "memory");
      }
    }
    if (nsteps > 0) {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      for (int j = (nsteps > D ? nsteps - D : 0); j < nsteps; ++j)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&full_bar[j % S])) : "memory");
    }
  }
  if (warp == 0) {
    if (cpa_a || cpa_b) {
      // (copied above together with the epilogue warps)
    } else if (lane == 0) {
      for (int it = 0; it < nsteps; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], (uint32_t)((a_boxes + b_boxes) * WG_BOX_BYTES));
        unsigned char* st = base + s * stage_bytes;
        const int m = (int)(row0 + (int64_t)it * WG_ROWS);
        for (int j = 0; j < a_boxes; ++j) tma_load_2d(st + j * WG_BOX_BYTES, &mapDY, &full_bar[s], n0 + 64 * j, m);
        for (int j = 0; j < b_boxes; ++j)
          tma_load_2d(st + A_BYTES + j * WG_BOX_BYTES, &mapX, &full_bar[s], k0 + 64 * j, m);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_mn(p.NB);
      for (int it = 0; it < nsteps; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(base + s * stage_bytes), b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int j = 0; j < WG_ROWS / 16; ++j) {
          // 16 rows of m = two 8-row groups = 2048 bytes further into every 64-column block
          umma_f16(tmem_base, make_smem_desc_mn(a_addr + j * 2048), make_smem_desc_mn(b_addr + j * 2048), idesc,
                   (it | j) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(accum_bar);
    }
  } else {
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int n = n0 + q * 32 + lane;
    const bool n_ok = n < p.Nn;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (partial != nullptr) {
      // two-stage reduction: this CTA's tile goes to its own slice of the workspace with plain 16-byte stores
      // ([M split][Np][ldp] fp32); pw_wgrad_reduce_kernel adds the slices in a fixed order afterwards.  (One fp32 red
      // per element per CTA -- up to 2 M same-address reds per launch -- was 60-70 % of the run time of the stage-3/4
      // layers, and made the result depend on the arrival order.)
      float* prow = partial + ((size_t)blockIdx.x * p.Np + n) * p.ldp + k0;
      const bool st_ok = n < p.Np;
      for (int c0 = 0; c0 < p.NB; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        if (st_ok) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<uint4*>(prow + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
      }
    } else {
      float* wrow = dW + (int64_t)n * p.K;
      for (int c0 = 0; c0 < p.NB; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        if (n_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = k0 + c0 + j;
            const float v = __uint_as_float(r[j]);
            if (k < p.K && v != 0.f) atomicAdd(wrow + k, v);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// second stage: dW[n][k] += sum over the M splits of partial[s][n][k], in a fixed order (deterministic).
// block = 32 consecutive outputs x 8 slices of the split index; slices are combined through shared memory.
__global__ void __launch_bounds__(256) pw_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW,
                                                              int nsplit, int Np, int ldp, int Nn, int K) {
  x3d::pdl_prologue();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + tx;                 // output index n*K + k
  const bool ok = o < Nn * K;
  const int n = ok ? o / K : 0, k = ok ? o - n * K : 0;
  const float* src = partial + (size_t)n * ldp + k;
  const size_t sstride = (size_t)Np * ldp;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (ok) {
    int s = ty;
    for (; s + 24 < nsplit; s += 32) {
      a0 += __ldcg(src + (size_t)s * sstride);
      a1 += __ldcg(src + (size_t)(s + 8) * sstride);
      a2 += __ldcg(src + (size_t)(s + 16) * sstride);
      a3 += __ldcg(src + (size_t)(s + 24) * sstride);
    }
    for (; s < nsplit; s += 8) a0 += __ldcg(src + (size_t)s * sstride);
  }
  red[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && ok) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += red[q][tx];
    dW[o] += v;
  }
}

}  // namespace

namespace x3d {
// dw[Nn][K] (fp32, += ) from x[M][Kp], dy[M][Np] (bf16, dense rows).  *handled = false -> SIMT path.
// workspace (optional, fp32, >= 1 MB): two-stage deterministic reduction instead of fp32 reds.
int pwconv_wgrad_tc(const void* x, const void* dy, float* dw, int64_t M, int64_t K, int64_t Kp, int64_t Nn,
                    int64_t Np, const int* gather /* nullptr | {stride, H, W, Ho, Wo}: X rows gathered */,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream, bool* handled) {
  *handled = false;
  static const bool off = getenv("X3D_PW_SIMT") != nullptr;
  if (off) return 0;
  if (M < 1 || Kp % 8 || Np % 8 || M >= (1ll << 31)) return 0;
  WgParams p;
  p.M = (int)M; p.K = (int)K; p.Kp = (int)Kp; p.Nn = (int)Nn; p.Np = (int)Np;
  p.gs = 1; p.gH = p.gW = p.gHo = p.gWo = 0;
  int64_t x_rows = M;
  if (gather != nullptr && gather[0] > 1) {
    static const bool no_cpa_g = getenv("X3D_TC_NOCPA") != nullptr;
    if (Kp >= 64 || no_cpa_g) return 0;   // the row gather lives in the cp.async loader of narrow operands
    p.gs = gather[0]; p.gH = gather[1]; p.gW = gather[2]; p.gHo = gather[3]; p.gWo = gather[4];
    x_rows = (M / ((int64_t)p.gHo * p.gWo)) * p.gH * p.gW;
    if (x_rows >= (1ll << 31)) return 0;
  }
  const int kparts = (int)((Kp + 255) / 256);
  p.NB = (int)(((Kp + kparts - 1) / kparts + 15) / 16 * 16);
  if (p.NB > 256) return 0;
  p.tmem_cols = p.NB <= 32 ? 32 : p.NB <= 64 ? 64 : p.NB <= 128 ? 128 : 256;
  p.a_boxes = Np > 64 ? 2 : 1; p.b_boxes_full = (p.NB + 63) / 64;
  const int ntiles = (int)((Np + 127) / 128);
  const int kz = (int)((Kp + p.NB - 1) / p.NB);
  // Every CTA ends with an fp32 red per element of its dW tile and same-address reds serialise in L2 (~40 ns
  // each): few M-splits for small M (1 CTA per SM, deep ring), up to 4 CTAs per SM only when M is huge and the
  // per-step producer overhead has to be hidden by co-resident CTAs.
  const int stage_kb = 8 * (p.a_boxes + p.b_boxes_full);
  static const int ctas_env = getenv("X3D_WG_CTAS") ? atoi(getenv("X3D_WG_CTAS")) : 0;   // tuning knob
  int64_t target = ctas_env > 0 ? ctas_env : M / 2700;
  if (target < kNumSMs) target = kNumSMs;
  if (target > 4 * kNumSMs) target = 4 * kNumSMs;
  int per_sm = (int)((target + kNumSMs - 1) / kNumSMs);
  static const int smem_kb = getenv("X3D_WG_SMEM_KB") ? atoi(getenv("X3D_WG_SMEM_KB")) : 200;   // tuning knob
  while (per_sm > 1 && per_sm * 3 * stage_kb > smem_kb) --per_sm;
  if (target > (int64_t)per_sm * kNumSMs) target = (int64_t)per_sm * kNumSMs;
  p.stages = (smem_kb / per_sm) / stage_kb;
  if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
  if (p.stages < 3) p.stages = 3;
  int64_t msplit = (target + ntiles * kz - 1) / (ntiles * kz);
  const int64_t max_split = (M + 511) / 512;
  if (msplit > max_split) msplit = max_split;
  p.ldp = p.NB * kz;
  const size_t slice_bytes = (size_t)Np * p.ldp * sizeof(float);
  static const bool no_ws = getenv("X3D_WG_ATOMIC") != nullptr;             // A/B switch
  const bool two_stage = workspace != nullptr && !no_ws && workspace_bytes >= slice_bytes &&
                         (reinterpret_cast<uintptr_t>(workspace) & 15) == 0;
  if (two_stage) {
    // one slice of the workspace per M split
    const int64_t ws_cap = (int64_t)(workspace_bytes / slice_bytes);
    if (msplit > ws_cap) msplit = ws_cap;
  } else {
    // every CTA ends with one fp32 red per element of its dW tile: keep the total under ~2M reds
    const int64_t red_cap = 2000000 / (Np * Kp) > 1 ? 2000000 / (Np * Kp) : 1;
    if (msplit > red_cap) msplit = red_cap;
  }
  if (msplit < 1) msplit = 1;
  int64_t rpc = ((M + msplit - 1) / msplit + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  msplit = (M + rpc - 1) / rpc;
  p.rows_per_cta = (int)rpc;
  CUtensorMap mapDY, mapX;
  if (!make_map_2d(&mapDY, dy, M, Np, WG_ROWS)) return 0;
  if (!make_map_2d(&mapX, x, x_rows, Kp, WG_ROWS)) return 0;
  const size_t smem = 1024 + (size_t)p.stages * (p.a_boxes + p.b_boxes_full) * WG_BOX_BYTES + (2 * WG_MAX_STAGES + 1) * 8 + 16;
  static unsigned long long attr_mask = 0;        // per device (the attribute is a per-device property)
  if (first_use_on_device(&attr_mask))
    cudaFuncSetAttribute(pw_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  dim3 grid((unsigned)msplit, (unsigned)ntiles, (unsigned)kz);
  static const bool no_cpa = getenv("X3D_TC_NOCPA") != nullptr;
  p.cpa = no_cpa ? 0 : 1;
  x3d::launch(pw_wgrad_tc_kernel, grid, NTHREADS, smem, stream, mapDY, mapX, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, dw, p,
              two_stage ? (float*)workspace : nullptr);
  *handled = true;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("pwconv_wgrad_tc: launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  count_launch();
  if (two_stage) {
    const int64_t outs = Nn * K;
    x3d::launch(pw_wgrad_reduce_kernel, dim3((unsigned)((outs + 31) / 32)), 256, 0, stream, (const float*)workspace, dw,
                (int)msplit, (int)Np, p.ldp, (int)Nn, (int)K);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("pwconv_wgrad_tc: reduce launch failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    count_launch();
  }
  return 0;
}

// y[M][Np] = x[M][Kp] * w[Np][Kp]^T  (all bf16, dense rows).  *handled = false -> caller uses the SIMT path.
int pwconv_fwd_tc(const void* x, const void* w, void* y, int64_t M, int64_t Kp, int64_t Np, int64_t P_out,
                  const int* gather /* nullptr | {stride, H, W, Ho, Wo}: A rows gathered */,
                  const int* scatter /* nullptr | {stride, H, W, Ho, Wo, accumulate}: C rows scattered */,
                  double* stats, cudaStream_t stream, bool* handled) {
  *handled = false;
  static const bool off = getenv("X3D_PW_SIMT") != nullptr;      // A/B switch for tests and profiling
  if (off) return 0;
  if (M < 1 || Kp % 8 || Np % 8 || M >= (1ll << 31)) return 0;
  // column tiles of width BN (multiple of 16, <= 256): the whole N when it fits one MMA, else N split in
  // near-equal parts; tile j covers columns [j*BN, min((j+1)*BN, Np)), weight rows beyond Np are TMA zero fill
  // <= 128 columns per CTA: accumulators take <= 256 TMEM columns, so >= 2 CTAs (8 epilogue warps) per SM
  static const int max_bn = getenv("X3D_TC_MAXBN") ? atoi(getenv("X3D_TC_MAXBN")) : 128;
  const int parts = (int)((Np + max_bn - 1) / max_bn);
  const int BN = (int)(((Np + parts - 1) / parts + 15) / 16 * 16);
  if (BN > 256) return 0;
  TcParams p;
  p.M = (int)M; p.Kp = (int)Kp; p.Np = (int)Np; p.BN = BN; p.ldc = (int)Np;
  p.nk = (int)((Kp + BK - 1) / BK);
  // two accumulators of BN fp32 columns each
  p.tmem_cols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  p.P_out = P_out;
  p.tiles_m = (int)((M + BM - 1) / BM);
  p.cpitch = BN + 8;
  static const int dbg = getenv("X3D_TC_DBG") ? atoi(getenv("X3D_TC_DBG")) : 0;
  p.dbg = dbg;
  static const bool no_cpa = getenv("X3D_TC_NOCPA") != nullptr;
  p.cpa = (Kp < BK && !no_cpa) ? 1 : 0;
  p.gs = 1; p.gH = p.gW = p.gHo = p.gWo = 0;
  if (gather != nullptr && gather[0] > 1) {
    if (!p.cpa) return 0;                 // the row gather lives in the cp.async loader (Kp < 64)
    p.gs = gather[0]; p.gH = gather[1]; p.gW = gather[2]; p.gHo = gather[3]; p.gWo = gather[4];
    if ((M / ((int64_t)p.gHo * p.gWo)) * p.gH * p.gW >= (1ll << 31)) return 0;
  }
  p.ss = 0; p.accum = 0;
  if (scatter != nullptr && scatter[0] > 1) {
    if (p.gs > 1 || stats != nullptr) return 0;
    p.ss = scatter[0]; p.gH = scatter[1]; p.gW = scatter[2]; p.gHo = scatter[3]; p.gWo = scatter[4];
    p.accum = scatter[5];
    if ((M / ((int64_t)p.gHo * p.gWo)) * p.gH * p.gW >= (1ll << 31)) return 0;
  } else if (scatter != nullptr) {
    if (stats != nullptr) return 0;
    p.accum = scatter[5];                 // dense rows, C += (identity-residual gradient already in dx)
  }
  p.stages = BN > 128 ? 3 : 2;      // small-N layers: 2 stages so that 2-3 CTAs fit per SM (nk is 1-2 there)
  static const int st_env = getenv("X3D_TC_STAGES") ? atoi(getenv("X3D_TC_STAGES")) : 0;      // tuning knob
  if (st_env >= 2 && st_env <= 6) p.stages = st_env;
  CUtensorMap mapA, mapB;
  if (!make_map_2d(&mapA, x, M, Kp, BM)) return 0;
  if (!make_map_2d(&mapB, w, Np, Kp, BN)) return 0;
  const size_t smem = 1024 + (size_t)p.stages * (BM * BK * 2) + (size_t)(p.cpa ? 1 : p.stages) * ((size_t)BN * BK * 2) +
                      (size_t)BM * p.cpitch * 2 + (2 * p.stages + 6) * 8 + 16 + (size_t)2 * BN * 2 * sizeof(float) + 4096 +
                      16 + 1024;
  // resident CTAs: TMEM (512 columns per SM) and shared memory (227 KB per SM) bound the co-residency
  int per_sm = 512 / p.tmem_cols;
  const int by_smem = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > by_smem) per_sm = by_smem;
  if (per_sm < 1) per_sm = 1;
  static const int per_sm_cap = getenv("X3D_TC_PERSM") ? atoi(getenv("X3D_TC_PERSM")) : 4;
  if (per_sm > per_sm_cap) per_sm = per_sm_cap;
  const int parts_n = (int)((Np + BN - 1) / BN);
  int gx = (kNumSMs * per_sm) / parts_n;
  if (gx < 1) gx = 1;
  if (gx > p.tiles_m) gx = p.tiles_m;
  gx = (p.tiles_m + (p.tiles_m + gx - 1) / gx - 1) / ((p.tiles_m + gx - 1) / gx);   // no empty CTAs
  dim3 grid((unsigned)gx, (unsigned)parts_n);
  static unsigned long long attr_mask[2] = {0, 0};   // per device
  if (stats) {
    if (first_use_on_device(&attr_mask[1])) cudaFuncSetAttribute(pw_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    x3d::launch(pw_tc_kernel<true>, grid, NTHREADS, smem, stream, mapA, mapB, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, p, stats);
  } else {
    if (first_use_on_device(&attr_mask[0])) cudaFuncSetAttribute(pw_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    x3d::launch(pw_tc_kernel<false>, grid, NTHREADS, smem, stream, mapA, mapB, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, p, nullptr);
  }
  *handled = true;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("pwconv_fwd_tc: launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  count_launch();
  return 0;
}
}  // namespace x3d
