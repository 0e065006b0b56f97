// Bandwidth-bound elementwise / reduction kernels of the X3D hot path:
// layout converters, parameter repack, split-BN finalize / apply / backward, Swish + SE gate,
// head pooling.  All activation tensors are [N][P][Cp] (NDHWC, Cp % 8 == 0), accessed with
// 16-byte vectors; block = (channel vectors) x (position rows), grid = (chunks, N).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace x3d {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static const bool on = getenv("X3D_NO_PDL") == nullptr;
  return on;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static std::atomic<int64_t> g_paths[X3D_PATH_COUNT];
void count_path(int path) {
  if (path >= 0 && path < X3D_PATH_COUNT) g_paths[path].fetch_add(1, std::memory_order_relaxed);
}
bool first_use_on_device(unsigned long long* mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  auto* m = reinterpret_cast<std::atomic<unsigned long long>*>(mask);
  if (m->load(std::memory_order_acquire) & bit) return false;
  m->fetch_or(bit, std::memory_order_acq_rel);     // a racing thread repeats the (idempotent) set-up: harmless
  return true;
}
}  // namespace x3d

using namespace x3d;

extern "C" const char* x3d_last_error(void) { return x3d::g_err; }
extern "C" int x3d_abi_version(void) { return 2; }
extern "C" int64_t x3d_launch_count(void) { return x3d::g_launches.load(); }
extern "C" int64_t x3d_path_count(int path) {
  return (path >= 0 && path < X3D_PATH_COUNT) ? x3d::g_paths[path].load() : -1;
}
extern "C" const char* x3d_path_name(int path) {
  static const char* names[X3D_PATH_COUNT] = {
      "dw_fwd_tiled",   "dw_fwd_temporal",   "dw_fwd_direct",   "dw_dgrad_tiled", "dw_dgrad_temporal",
      "dw_dgrad_direct", "dw_wgrad_tiled",   "dw_wgrad_temporal", "dw_wgrad_direct", "pw_fwd_tc",
      "pw_fwd_simt",    "pw_dgrad_tc",       "pw_dgrad_simt",   "pw_wgrad_tc",    "pw_wgrad_simt"};
  return (path >= 0 && path < X3D_PATH_COUNT) ? names[path] : "?";
}

#define ROW_PROLOGUE()                                    \
  constexpr int VEC = Vec<T>::N;                          \
  const int n = blockIdx.y;                               \
  const int cvec = threadIdx.x % cv;                      \
  const int prow = threadIdx.x / cv;                      \
  const int64_t p0 = (int64_t)blockIdx.x * chunk;         \
  const int64_t p1 = (p0 + chunk < P) ? p0 + chunk : P;   \
  const int c0 = cvec * VEC;                              \
  (void)c0

// =======================================================================================
// layout converters
// =======================================================================================
// src NCDHW fp32 -> dst NDHWC T with Cp lanes.  32x32 smem transpose over (c, spatial).
template <typename T>
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int Cp,
                                      int64_t S /*T*H*W*/) {
  x3d::pdl_prologue();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32;
  const int cb = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = cb + i;
    int64_t s = s0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && s < S) ? src[((int64_t)n * C + c) * S + s] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t s = s0 + i;
    int c = cb + threadIdx.x;
    if (s < S && c < Cp) dst[((int64_t)n * S + s) * Cp + c] = from_float<T>(tile[threadIdx.x][i]);
  }
}
template <typename T>
__global__ void ndhwc_to_ncdhw_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int Cp,
                                      int64_t S) {
  x3d::pdl_prologue();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32;
  const int cb = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t s = s0 + i;
    int c = cb + threadIdx.x;
    tile[i][threadIdx.x] = (s < S && c < C) ? to_float<T>(src[((int64_t)n * S + s) * Cp + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = cb + i;
    int64_t s = s0 + threadIdx.x;
    if (c < C && s < S) dst[((int64_t)n * C + c) * S + s] = tile[threadIdx.x][i];
  }
}

extern "C" int x3d_ncdhw_to_ndhwc(const float* src, void* dst, int64_t N, int64_t C, int64_t Cp, int64_t T_,
                                  int64_t H, int64_t W, x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0 && Cp >= C, "Cp must be a multiple of 8 and >= C");
  int64_t S = T_ * H * W;
  if (N * S == 0) return 0;
  dim3 grid((unsigned)cdiv(S, 32), (unsigned)cdiv(Cp, 32), (unsigned)N), block(32, 8);
  X3D_DISPATCH_DTYPE(dt, (x3d::launch(ncdhw_to_ndhwc_kernel<T>, grid, block, 0, as_stream(stream), src, (T*)dst, (int)C, (int)Cp, S)));
  X3D_LAUNCH_CHECK();
  return 0;
}
extern "C" int x3d_ndhwc_to_ncdhw(const void* src, float* dst, int64_t N, int64_t C, int64_t Cp, int64_t T_,
                                  int64_t H, int64_t W, x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0 && Cp >= C, "Cp must be a multiple of 8 and >= C");
  int64_t S = T_ * H * W;
  if (N * S == 0) return 0;
  dim3 grid((unsigned)cdiv(S, 32), (unsigned)cdiv(C, 32), (unsigned)N), block(32, 8);
  X3D_DISPATCH_DTYPE(dt, (x3d::launch(ndhwc_to_ncdhw_kernel<T>, grid, block, 0, as_stream(stream), (const T*)src, dst, (int)C, (int)Cp, S)));
  X3D_LAUNCH_CHECK();
  return 0;
}

// =======================================================================================
// parameter repack (one launch for the whole network)
// =======================================================================================
__global__ void pack_params_kernel(const x3d_pack_desc_t* __restrict__ descs) {
  x3d::pdl_prologue();
  const x3d_pack_desc_t d = descs[blockIdx.y];
  const int64_t total = (int64_t)d.dst_rows * d.dst_cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int dr = (int)(i / d.dst_cols), dc = (int)(i % d.dst_cols);
    int r = d.transpose ? dc : dr;
    int c = d.transpose ? dr : dc;
    float v = (r < d.rows && c < d.cols) ? d.src[(int64_t)r * d.cols + c] : 0.f;
    if (d.dtype == X3D_BF16)
      reinterpret_cast<__nv_bfloat16*>(d.dst)[i] = __float2bfloat16_rn(v);
    else
      reinterpret_cast<float*>(d.dst)[i] = v;
  }
}
extern "C" int x3d_pack_params(const x3d_pack_desc_t* descs_dev, int n_desc, int64_t max_dst_elems,
                               x3d_stream_t stream) {
  if (n_desc == 0) return 0;
  int64_t bx = cdiv(max_dst_elems, 256);
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  x3d::launch(pack_params_kernel, dim3((unsigned)bx, (unsigned)n_desc), 256, 0, as_stream(stream), descs_dev);
  X3D_LAUNCH_CHECK();
  return 0;
}

// =======================================================================================
// split-BN finalize: per-sample sums -> per-(split,channel) scale/shift + running stats
// =======================================================================================
// sum of the per-sample {s1, s2} pairs of split b, channel c: 16-byte loads, 4 independent accumulators so that
// the (cold) global loads overlap instead of forming one dependent chain
__device__ __forceinline__ void sum_split(const double* __restrict__ stats, int N, int splits, int b, int c, int Cp,
                                          double& s1, double& s2) {
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
  const double2* p = reinterpret_cast<const double2*>(stats) + ((int64_t)b * Cp + c);
  const int64_t step = (int64_t)splits * Cp;
  const int cnt = (N - b + splits - 1) / splits;
  int i = 0;
  for (; i + 4 <= cnt; i += 4) {
    const double2 v0 = __ldg(p + (int64_t)(i + 0) * step), v1 = __ldg(p + (int64_t)(i + 1) * step);
    const double2 v2 = __ldg(p + (int64_t)(i + 2) * step), v3 = __ldg(p + (int64_t)(i + 3) * step);
    a0 += v0.x; a1 += v0.y; b0 += v1.x; b1 += v1.y; c0 += v2.x; c1 += v2.y; d0 += v3.x; d1 += v3.y;
  }
  for (; i < cnt; ++i) {
    const double2 v = __ldg(p + (int64_t)i * step);
    a0 += v.x; a1 += v.y;
  }
  s1 = (a0 + b0) + (c0 + d0);
  s2 = (a1 + b1) + (c1 + d1);
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, int N, int splits, double inv_m, double unbias, int C,
                                   int Cp,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ run_mean, float* __restrict__ run_var,
                                   int64_t* __restrict__ nbt, float momentum, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_o, float* __restrict__ rstd_o) {
  x3d::pdl_prologue();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0 && nbt) *nbt += 1;
  if (idx >= splits * Cp) return;
  const int b = idx / Cp, c = idx % Cp;
  if (c >= C) {
    scale[idx] = 0.f; shift[idx] = 0.f; mean_o[idx] = 0.f; rstd_o[idx] = 0.f;
    return;
  }
  double s1, s2;
  sum_split(stats, N, splits, b, c, Cp, s1, s2);
  // No fp64 division / square root here (software sequences of a few hundred cycles each on this part, and this
  // kernel sits between every conv and its consumer): 1/m and m/(m-1) come from the host, and rstd -- stored as fp32
  // anyway -- is rsqrtf refined by one Newton step (<= 1 ulp of fp32).
  const double mu = s1 * inv_m;
  double var = s2 * inv_m - mu * mu;
  if (var < 0.0) var = 0.0;
  const float vf = (float)(var + (double)eps);
  float rf = rsqrtf(vf);
  rf = rf * fmaf(-0.5f * vf * rf, rf, 1.5f);
  const double r = (double)rf;
  const float sc = (float)((double)gamma[c] * r);
  scale[idx] = sc;
  shift[idx] = (float)((double)beta[c] - mu * (double)gamma[c] * r);
  mean_o[idx] = (float)mu;
  rstd_o[idx] = (float)r;
  if (run_mean) {
    const int ri = b * C + c;  // x3d.py:50: channel index of the (n//s, c*s) view is b*C + c
    const double unb = var * unbias;
    run_mean[ri] = (float)((1.0 - momentum) * (double)run_mean[ri] + momentum * mu);
    run_var[ri] = (float)((1.0 - momentum) * (double)run_var[ri] + momentum * unb);
  }
}
extern "C" int x3d_bn_finalize(const double* stats, int64_t N, int splits, int64_t P, int64_t C, int64_t Cp,
                               const float* gamma, const float* beta, float* run_mean, float* run_var,
                               int64_t* nbt, float momentum, float eps, float* scale, float* shift,
                               float* mean, float* rstd, x3d_stream_t stream) {
  X3D_CHECK_ARG(splits >= 1 && N % splits == 0, "batch must be divisible by num_splits (x3d.py:50)");
  int total = (int)(splits * Cp);
  const double m = (double)(N / splits) * (double)P;
  x3d::launch(bn_finalize_kernel, (unsigned)cdiv(total, 64), 64, 0, as_stream(stream),
      stats, (int)N, splits, 1.0 / m, m > 1.0 ? m / (m - 1.0) : 1.0, (int)C, (int)Cp, gamma, beta, run_mean, run_var, nbt,
      momentum, eps, scale, shift, mean, rstd);
  X3D_LAUNCH_CHECK();
  return 0;
}

__global__ void bn_eval_params_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, int C, int Cp,
                                      float eps, float* scale, float* shift, float* mean, float* rstd) {
  x3d::pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  float sc = 0.f, sh = 0.f, mu = 0.f, r = 0.f;
  if (c < C) {
    r = (float)(1.0 / sqrt((double)rv[c] + (double)eps));
    mu = rm[c];
    sc = gamma[c] * r;
    sh = beta[c] - mu * sc;
  }
  scale[c] = sc; shift[c] = sh;
  if (mean) mean[c] = mu;
  if (rstd) rstd[c] = r;
}
extern "C" int x3d_bn_eval_params(const float* gamma, const float* beta, const float* run_mean,
                                  const float* run_var, int64_t C, int64_t Cp, float eps, float* scale,
                                  float* shift, float* mean, float* rstd, x3d_stream_t stream) {
  x3d::launch(bn_eval_params_kernel, (unsigned)cdiv(Cp, 128), 128, 0, as_stream(stream), gamma, beta, run_mean, run_var, (int)C,
                                                                            (int)Cp, eps, scale, shift, mean, rstd);
  X3D_LAUNCH_CHECK();
  return 0;
}

// =======================================================================================
// out = [relu](a*scale+shift [+ res | + res*rs+rh])
// =======================================================================================
template <typename T, int RES /*0 none, 1 raw, 2 affine*/, bool RELU>
__global__ void bn_act_fwd_kernel(const T* __restrict__ a, const float* __restrict__ scale,
                                  const float* __restrict__ shift, int splits, const T* __restrict__ res,
                                  const float* __restrict__ rscale, const float* __restrict__ rshift,
                                  T* __restrict__ out, int64_t P, int Cp, int cv, int rows, int64_t chunk) {
  x3d::pdl_prologue();
  ROW_PROLOGUE();
  const int b = n % splits;
  float sc[VEC], sh[VEC], rs[VEC], rh[VEC];
  load_consts<VEC>(scale + b * Cp + c0, sc);
  load_consts<VEC>(shift + b * Cp + c0, sh);
  if (RES == 2) {
    load_consts<VEC>(rscale + b * Cp + c0, rs);
    load_consts<VEC>(rshift + b * Cp + c0, rh);
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) { rs[j] = 1.f; rh[j] = 0.f; }
  }
  auto body = [&](float (&v)[VEC], const float (&r)[VEC]) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float y = fmaf(v[j], sc[j], sh[j]);
      if (RES == 1) y += r[j];
      if (RES == 2) y += fmaf(r[j], rs[j], rh[j]);
      if (RELU) y = fmaxf(y, 0.f);
      v[j] = y;
    }
  };
  int64_t p = p0 + prow;
  for (; p + rows < p1; p += 2 * (int64_t)rows) {               // two rows per iteration: all loads in flight first
    const int64_t off0 = ((int64_t)n * P + p) * Cp + c0, off1 = off0 + (int64_t)rows * Cp;
    float v0[VEC], r0[VEC], v1[VEC], r1[VEC];
    load_vec<T>(a + off0, v0);
    if (RES) load_vec<T>(res + off0, r0);
    load_vec<T>(a + off1, v1);
    if (RES) load_vec<T>(res + off1, r1);
    body(v0, r0);
    body(v1, r1);
    store_vec<T>(out + off0, v0);
    store_vec<T>(out + off1, v1);
  }
  for (; p < p1; p += rows) {
    const int64_t off = ((int64_t)n * P + p) * Cp + c0;
    float v[VEC], r[VEC];
    load_vec<T>(a + off, v);
    if (RES) load_vec<T>(res + off, r);
    body(v, r);
    store_vec<T>(out + off, v);
  }
}

extern "C" int x3d_bn_act_fwd(const void* a, const float* scale, const float* shift, int splits,
                              const void* res, const float* res_scale, const float* res_shift, int relu,
                              void* out, int64_t N, int64_t P, int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  if (N * P == 0) return 0;
  const int mode = res ? (res_scale ? 2 : 1) : 0;
#define LAUNCH_(M, R)                                                                                          \
  x3d::launch(bn_act_fwd_kernel<T, M, R>, grid, g.threads, 0, as_stream(stream), (const T*)a, scale, shift, splits,      \
                                                                         (const T*)res, res_scale, res_shift,  \
                                                                         (T*)out, P, (int)Cp, g.cv, g.rows, g.chunk)
  X3D_DISPATCH_DTYPE(dt, {
    RowGeom g = make_row_geom<T>(N, P, Cp, 8 * kNumSMs);
    dim3 grid(g.chunks, (unsigned)N);
    if (mode == 0 && relu) LAUNCH_(0, true);
    else if (mode == 0) LAUNCH_(0, false);
    else if (mode == 1 && relu) LAUNCH_(1, true);
    else if (mode == 1) LAUNCH_(1, false);
    else if (relu) LAUNCH_(2, true);
    else LAUNCH_(2, false);
  });
#undef LAUNCH_
  X3D_LAUNCH_CHECK();
  return 0;
}

// =======================================================================================
// BN backward
// =======================================================================================
template <typename T, bool MASK, bool STORE>
__global__ void bn_bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ mask_out,
                                     const T* __restrict__ a, double* __restrict__ stats, T* __restrict__ dpre_out,
                                     int64_t P, int Cp, int cv, int rows, int64_t chunk) {
  x3d::pdl_prologue();
  extern __shared__ float s_acc[];
  ROW_PROLOGUE();
  float a0[VEC], a1[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) a0[j] = a1[j] = 0.f;
  auto body = [&](float (&d)[VEC], const float (&x)[VEC], const float (&m)[VEC]) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float dp = (!MASK || m[j] > 0.f) ? d[j] : 0.f;
      d[j] = dp;
      a0[j] += dp;
      a1[j] = fmaf(dp, x[j], a1[j]);
    }
  };
  int64_t p = p0 + prow;
  for (; p + rows < p1; p += 2 * (int64_t)rows) {               // two rows per iteration: all loads in flight first
    const int64_t off0 = ((int64_t)n * P + p) * Cp + c0, off1 = off0 + (int64_t)rows * Cp;
    float d0[VEC], x0[VEC], m0[VEC], d1[VEC], x1[VEC], m1[VEC];
    load_vec<T>(dout + off0, d0);
    load_vec<T>(a + off0, x0);
    if (MASK) load_vec<T>(mask_out + off0, m0);
    load_vec<T>(dout + off1, d1);
    load_vec<T>(a + off1, x1);
    if (MASK) load_vec<T>(mask_out + off1, m1);
    body(d0, x0, m0);
    body(d1, x1, m1);
    if (STORE) {
      store_vec<T>(dpre_out + off0, d0);
      store_vec<T>(dpre_out + off1, d1);
    }
  }
  for (; p < p1; p += rows) {
    const int64_t off = ((int64_t)n * P + p) * Cp + c0;
    float d[VEC], x[VEC], m[VEC];
    load_vec<T>(dout + off, d);
    load_vec<T>(a + off, x);
    if (MASK) load_vec<T>(mask_out + off, m);
    body(d, x, m);
    if (STORE) store_vec<T>(dpre_out + off, d);
  }
  block_stats_flush<VEC>(a0, a1, cvec, Cp, s_acc, stats + (int64_t)n * Cp * 2);
}
static int bn_bwd_reduce_impl(const void* dout, const void* mask_out, const void* a, double* stats, void* dpre_out,
                              int64_t N, int64_t P, int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  if (N * P == 0) return 0;
#define RL_(MK, STO)                                                                                         \
  x3d::launch(bn_bwd_reduce_kernel<T, MK, STO>, grid, g.threads, smem, as_stream(stream), (const T*)dout,       \
              (const T*)mask_out, (const T*)a, stats, (T*)dpre_out, P, (int)Cp, g.cv, g.rows, g.chunk)
  X3D_DISPATCH_DTYPE(dt, {
    RowGeom g = make_row_geom<T>(N, P, Cp);
    dim3 grid(g.chunks, (unsigned)N);
    size_t smem = (size_t)g.rows * Cp * 2 * sizeof(float);
    if (mask_out && dpre_out) RL_(true, true);
    else if (mask_out) RL_(true, false);
    else if (dpre_out) RL_(false, true);
    else RL_(false, false);
  });
#undef RL_
  X3D_LAUNCH_CHECK();
  return 0;
}
extern "C" int x3d_bn_bwd_reduce(const void* dout, const void* mask_out, const void* a, double* stats,
                                 int64_t N, int64_t P, int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream) {
  return bn_bwd_reduce_impl(dout, mask_out, a, stats, nullptr, N, P, Cp, dt, stream);
}
extern "C" int x3d_bn_bwd_reduce_store(const void* dout, const void* mask_out, const void* a, double* stats,
                                       void* dpre_out, int64_t N, int64_t P, int64_t Cp, x3d_dtype_t dt,
                                       x3d_stream_t stream) {
  X3D_CHECK_ARG(dpre_out != nullptr, "dpre_out");
  return bn_bwd_reduce_impl(dout, mask_out, a, stats, dpre_out, N, P, Cp, dt, stream);
}

// coefficients: da = A*dpre + B*a + Cc.  One thread per channel, loops splits and samples.
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ stats, int N, int splits, double inv_m, int C, int Cp,
                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ rstd, int train, float* __restrict__ coef,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  x3d::pdl_prologue();
  // one thread per (split, channel)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int sC = splits * Cp;
  if (idx >= sC) return;
  const int b = idx / Cp, c = idx % Cp;
  if (c >= C) {
    coef[0 * sC + idx] = 0.f; coef[1 * sC + idx] = 0.f; coef[2 * sC + idx] = 0.f;
    return;
  }
  const double g = gamma[c];
  double s1, s2;
  sum_split(stats, N, splits, b, c, Cp, s1, s2);
  const double mu = mean[idx], r = rstd[idx];
  const double sxh = (s2 - mu * s1) * r;  // sum dpre * xhat
  double A = g * r, B = 0.0, Cc = 0.0;
  if (train) {
    const double m1 = s1 * inv_m, m2 = sxh * inv_m;   // 1/m from the host: no fp64 division
    B = -g * r * r * m2;
    Cc = g * r * (-m1 + mu * r * m2);
  }
  coef[0 * sC + idx] = (float)A;
  coef[1 * sC + idx] = (float)B;
  coef[2 * sC + idx] = (float)Cc;
  if (dgamma) atomicAdd(&dgamma[c], (float)sxh);
  if (dbeta) atomicAdd(&dbeta[c], (float)s1);
}
extern "C" int x3d_bn_bwd_finalize(const double* stats, int64_t N, int splits, int64_t P, int64_t C, int64_t Cp,
                                   const float* gamma, const float* mean, const float* rstd, int train,
                                   float* coef, float* dgamma, float* dbeta, x3d_stream_t stream) {
  x3d::launch(bn_bwd_finalize_kernel, (unsigned)cdiv(splits * Cp, 64), 64, 0, as_stream(stream), stats, (int)N, splits,
              1.0 / ((double)(N / splits) * (double)P), (int)C,
                                                                            (int)Cp, gamma, mean, rstd, train, coef,
                                                                            dgamma, dbeta);
  X3D_LAUNCH_CHECK();
  return 0;
}

template <typename T, bool MASK>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dout, const T* __restrict__ mask_out,
                                    const T* __restrict__ a, const float* __restrict__ coef, int splits,
                                    T* __restrict__ da, int64_t P, int Cp, int cv, int rows, int64_t chunk) {
  x3d::pdl_prologue();
  ROW_PROLOGUE();
  const int b = n % splits;
  const int sC = splits * Cp;
  float A[VEC], B[VEC], Cc[VEC];
  load_consts<VEC>(coef + 0 * sC + b * Cp + c0, A);
  load_consts<VEC>(coef + 1 * sC + b * Cp + c0, B);
  load_consts<VEC>(coef + 2 * sC + b * Cp + c0, Cc);
  auto body = [&](float (&d)[VEC], const float (&x)[VEC], const float (&m)[VEC]) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float dp = (!MASK || m[j] > 0.f) ? d[j] : 0.f;
      d[j] = fmaf(A[j], dp, fmaf(B[j], x[j], Cc[j]));
    }
  };
  int64_t p = p0 + prow;
  for (; p + rows < p1; p += 2 * (int64_t)rows) {               // two rows per iteration: all loads in flight first
    const int64_t off0 = ((int64_t)n * P + p) * Cp + c0, off1 = off0 + (int64_t)rows * Cp;
    float d0[VEC], x0[VEC], m0[VEC], d1[VEC], x1[VEC], m1[VEC];
    load_vec<T>(dout + off0, d0);
    load_vec<T>(a + off0, x0);
    if (MASK) load_vec<T>(mask_out + off0, m0);
    load_vec<T>(dout + off1, d1);
    load_vec<T>(a + off1, x1);
    if (MASK) load_vec<T>(mask_out + off1, m1);
    body(d0, x0, m0);
    body(d1, x1, m1);
    store_vec<T>(da + off0, d0);
    store_vec<T>(da + off1, d1);
  }
  for (; p < p1; p += rows) {
    const int64_t off = ((int64_t)n * P + p) * Cp + c0;
    float d[VEC], x[VEC], m[VEC];
    load_vec<T>(dout + off, d);
    load_vec<T>(a + off, x);
    if (MASK) load_vec<T>(mask_out + off, m);
    body(d, x, m);
    store_vec<T>(da + off, d);
  }
}
extern "C" int x3d_bn_bwd_apply(const void* dout, const void* mask_out, const void* a, const float* coef,
                                int splits, void* da, int64_t N, int64_t P, int64_t Cp, x3d_dtype_t dt,
                                x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  if (N * P == 0) return 0;
  X3D_DISPATCH_DTYPE(dt, {
    RowGeom g = make_row_geom<T>(N, P, Cp, 8 * kNumSMs);
    dim3 grid(g.chunks, (unsigned)N);
    if (mask_out)
      x3d::launch(bn_bwd_apply_kernel<T, true>, grid, g.threads, 0, as_stream(stream), 
          (const T*)dout, (const T*)mask_out, (const T*)a, coef, splits, (T*)da, P, (int)Cp, g.cv, g.rows, g.chunk);
    else
      x3d::launch(bn_bwd_apply_kernel<T, false>, grid, g.threads, 0, as_stream(stream), 
          (const T*)dout, nullptr, (const T*)a, coef, splits, (T*)da, P, (int)Cp, g.cv, g.rows, g.chunk);
  });
  X3D_LAUNCH_CHECK();
  return 0;
}

template <typename T>
__global__ void relu_bwd_add_kernel(const T* __restrict__ dout, const T* __restrict__ out, T* __restrict__ dx,
                                    int64_t nvec) {
  x3d::pdl_prologue();
  constexpr int VEC = Vec<T>::N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float d[VEC], o[VEC], x[VEC];
    load_vec<T>(dout + i * VEC, d);
    load_vec<T>(out + i * VEC, o);
    load_vec<T>(dx + i * VEC, x);
#pragma unroll
    for (int j = 0; j < VEC; ++j) x[j] += (o[j] > 0.f) ? d[j] : 0.f;
    store_vec<T>(dx + i * VEC, x);
  }
}
extern "C" int x3d_relu_bwd_add(const void* dout, const void* out, void* dx, int64_t numel, x3d_dtype_t dt,
                                x3d_stream_t stream) {
  X3D_CHECK_ARG(numel % 8 == 0, "numel % 8");
  if (numel == 0) return 0;
  X3D_DISPATCH_DTYPE(dt, {
    int64_t nvec = numel / Vec<T>::N;
    int64_t blocks = cdiv(nvec, 256);
    if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
    x3d::launch(relu_bwd_add_kernel<T>, (unsigned)blocks, 256, 0, as_stream(stream), (const T*)dout, (const T*)out, (T*)dx, nvec);
  });
  X3D_LAUNCH_CHECK();
  return 0;
}

// =======================================================================================
// SE forward (tiny): one block per sample
// =======================================================================================
__global__ void se_fwd_kernel(const double* __restrict__ stats, const float* __restrict__ scale,
                              const float* __restrict__ shift, int splits, double invP, int C, int Cp, int sw,
                              const float* __restrict__ W1, const float* __restrict__ b1,
                              const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ pooled,
                              float* __restrict__ hidden, float* __restrict__ gate) {
  x3d::pdl_prologue();
  extern __shared__ float sm[];
  float* s_p = sm;        // [C]
  float* s_h = sm + C;    // [sw]
  const int n = blockIdx.x, b = n % splits;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mu = (float)(stats[((int64_t)n * Cp + c) * 2] * invP);
    float p = fmaf(scale[b * Cp + c], mu, shift[b * Cp + c]);
    s_p[c] = p;
    pooled[(int64_t)n * C + c] = p;
  }
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
  for (int j = warp; j < sw; j += nw) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(W1[(int64_t)j * C + c], s_p[c], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float h = fmaxf(acc + b1[j], 0.f);
      s_h[j] = h;
      hidden[(int64_t)n * sw + j] = h;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    float g = 0.f;
    if (c < C) {
      float acc = b2[c];
      for (int j = 0; j < sw; ++j) acc = fmaf(W2[(int64_t)c * sw + j], s_h[j], acc);
      g = 1.f / (1.f + expf(-acc));
    }
    gate[(int64_t)n * Cp + c] = g;
  }
}
extern "C" int x3d_se_fwd(const double* stats, const float* scale, const float* shift, int splits, int64_t N,
                          int64_t P, int64_t C, int64_t Cp, int sw, const float* W1, const float* b1,
                          const float* W2, const float* b2, float* pooled, float* hidden, float* gate,
                          x3d_stream_t stream) {
  if (N == 0) return 0;
  size_t smem = (C + sw) * sizeof(float);
  x3d::launch(se_fwd_kernel, (unsigned)N, 256, smem, as_stream(stream), stats, scale, shift, splits, 1.0 / (double)P, (int)C,
                                                              (int)Cp, sw, W1, b1, W2, b2, pooled, hidden, gate);
  X3D_LAUNCH_CHECK();
  return 0;
}

// =======================================================================================
// Swish (+ SE gate) forward / backward
// =======================================================================================
// sigmoid for bf16-storage tensors: one MUFU (tanh.approx, rel. error 2^-11, four times finer than the bf16 the result
// is stored in) instead of ex2 + rcp; fp32 storage keeps the exact-ish exp / divide path (parity mode)
template <typename T>
__device__ __forceinline__ float sigmoid_t(float x) { return sigmoidf_(x); }
template <>
__device__ __forceinline__ float sigmoid_t<__nv_bfloat16>(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
template <typename T, bool GATE>
__global__ void swish_gate_fwd_kernel(const T* __restrict__ a2, const float* __restrict__ scale,
                                      const float* __restrict__ shift, int splits, const float* __restrict__ gate,
                                      T* __restrict__ v, int64_t P, int Cp, int cv, int rows, int64_t chunk) {
  x3d::pdl_prologue();
  ROW_PROLOGUE();
  const int b = n % splits;
  float sc[VEC], sh[VEC];
  load_consts<VEC>(scale + b * Cp + c0, sc);
  load_consts<VEC>(shift + b * Cp + c0, sh);
  if (GATE) {
    float g[VEC];
    load_consts<VEC>(gate + (int64_t)n * Cp + c0, g);
#pragma unroll
    for (int j = 0; j < VEC; ++j) { sc[j] *= g[j]; sh[j] *= g[j]; }   // z = g*(sc*a+sh)
  }
  int64_t p = p0 + prow;
  for (; p + rows < p1; p += 2 * (int64_t)rows) {               // two rows per iteration: both loads in flight first
    const int64_t off0 = ((int64_t)n * P + p) * Cp + c0, off1 = off0 + (int64_t)rows * Cp;
    float x0[VEC], x1[VEC];
    load_vec<T>(a2 + off0, x0);
    load_vec<T>(a2 + off1, x1);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float z0 = fmaf(x0[j], sc[j], sh[j]), z1 = fmaf(x1[j], sc[j], sh[j]);
      x0[j] = z0 * sigmoid_t<T>(z0);
      x1[j] = z1 * sigmoid_t<T>(z1);
    }
    store_vec<T>(v + off0, x0);
    store_vec<T>(v + off1, x1);
  }
  for (; p < p1; p += rows) {
    const int64_t off = ((int64_t)n * P + p) * Cp + c0;
    float x[VEC];
    load_vec<T>(a2 + off, x);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float z = fmaf(x[j], sc[j], sh[j]);
      x[j] = z * sigmoid_t<T>(z);
    }
    store_vec<T>(v + off, x);
  }
}
extern "C" int x3d_swish_gate_fwd(const void* a2, const float* scale, const float* shift, int splits,
                                  const float* gate, void* v, int64_t N, int64_t P, int64_t Cp, x3d_dtype_t dt,
                                  x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  if (N * P == 0) return 0;
  X3D_DISPATCH_DTYPE(dt, {
    RowGeom g = make_row_geom<T>(N, P, Cp, 8 * kNumSMs);
    dim3 grid(g.chunks, (unsigned)N);
    if (gate)
      x3d::launch(swish_gate_fwd_kernel<T, true>, grid, g.threads, 0, as_stream(stream), (const T*)a2, scale, shift, splits, gate,
                                                                               (T*)v, P, (int)Cp, g.cv, g.rows, g.chunk);
    else
      x3d::launch(swish_gate_fwd_kernel<T, false>, grid, g.threads, 0, as_stream(stream), (const T*)a2, scale, shift, splits,
                                                                                nullptr, (T*)v, P, (int)Cp, g.cv, g.rows,
                                                                                g.chunk);
  });
  X3D_LAUNCH_CHECK();
  return 0;
}

// d swish(z)/dz = s*(1 + z*(1-s)), s = sigmoid(z)   (x3d.py:81-84)
template <typename T>
__device__ __forceinline__ float swish_grad(float z) {
  float s = sigmoid_t<T>(z);
  return s * fmaf(z, 1.f - s, 1.f);
}

template <typename T, bool GATE>
__global__ void swish_gate_bwd_reduce_kernel(const T* __restrict__ dv, const T* __restrict__ a2,
                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                             int splits, const float* __restrict__ gate, double* __restrict__ stats,
                                             int64_t P, int Cp, int cv, int rows, int64_t chunk) {
  x3d::pdl_prologue();
  extern __shared__ float s_acc[];
  ROW_PROLOGUE();
  const int b = n % splits;
  float sc[VEC], sh[VEC], a0[VEC], a1[VEC];
  load_consts<VEC>(scale + b * Cp + c0, sc);
  load_consts<VEC>(shift + b * Cp + c0, sh);
  if (GATE) {
    float g[VEC];
    load_consts<VEC>(gate + (int64_t)n * Cp + c0, g);
#pragma unroll
    for (int j = 0; j < VEC; ++j) { sc[j] *= g[j]; sh[j] *= g[j]; }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) a0[j] = a1[j] = 0.f;
  int64_t p = p0 + prow;
  for (; p + rows < p1; p += 2 * (int64_t)rows) {               // two rows per iteration: four loads in flight first
    const int64_t off0 = ((int64_t)n * P + p) * Cp + c0, off1 = off0 + (int64_t)rows * Cp;
    float d0[VEC], x0[VEC], d1[VEC], x1[VEC];
    load_vec<T>(dv + off0, d0);
    load_vec<T>(a2 + off0, x0);
    load_vec<T>(dv + off1, d1);
    load_vec<T>(a2 + off1, x1);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float dz0 = d0[j] * swish_grad<T>(fmaf(x0[j], sc[j], sh[j]));
      const float dz1 = d1[j] * swish_grad<T>(fmaf(x1[j], sc[j], sh[j]));
      a0[j] += dz0;
      a1[j] = fmaf(dz0, x0[j], a1[j]);
      a0[j] += dz1;
      a1[j] = fmaf(dz1, x1[j], a1[j]);
    }
  }
  for (; p < p1; p += rows) {
    const int64_t off = ((int64_t)n * P + p) * Cp + c0;
    float d[VEC], x[VEC];
    load_vec<T>(dv + off, d);
    load_vec<T>(a2 + off, x);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float dz = d[j] * swish_grad<T>(fmaf(x[j], sc[j], sh[j]));
      a0[j] += dz;
      a1[j] = fmaf(dz, x[j], a1[j]);
    }
  }
  block_stats_flush<VEC>(a0, a1, cvec, Cp, s_acc, stats + (int64_t)n * Cp * 2);
}
extern "C" int x3d_swish_gate_bwd_reduce(const void* dv, const void* a2, const float* scale, const float* shift,
                                         int splits, const float* gate, double* stats, int64_t N, int64_t P,
                                         int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  if (N * P == 0) return 0;
  X3D_DISPATCH_DTYPE(dt, {
    RowGeom g = make_row_geom<T>(N, P, Cp);
    dim3 grid(g.chunks, (unsigned)N);
    size_t smem = (size_t)g.rows * Cp * 2 * sizeof(float);
    if (gate)
      x3d::launch(swish_gate_bwd_reduce_kernel<T, true>, grid, g.threads, smem, as_stream(stream), 
          (const T*)dv, (const T*)a2, scale, shift, splits, gate, stats, P, (int)Cp, g.cv, g.rows, g.chunk);
    else
      x3d::launch(swish_gate_bwd_reduce_kernel<T, false>, grid, g.threads, smem, as_stream(stream), 
          (const T*)dv, (const T*)a2, scale, shift, splits, nullptr, stats, P, (int)Cp, g.cv, g.rows, g.chunk);
  });
  X3D_LAUNCH_CHECK();
  return 0;
}

// SE backward for one sample (block per sample): produces work[n][c] = dp[n][c]/P plus dz2[n][c] and dz1[n][j], the
// per-sample factors of the SE parameter gradients (summed over the batch, in a fixed order, by se_param_grad_kernel:
// no atomics -- with the 64-256 clips per GPU of the multigrid shapes, N blocks hammering the same C x w addresses
// took 38 us per SE block).  u = scale*a2+shift (bn2 output), z = gate*u, dz = dv*swish'(z).
__global__ void se_bwd_sample_kernel(const double* __restrict__ fwd_stats, const double* __restrict__ bwd_stats,
                                     int splits, double inv_P, int C, int Cp, int sw, const float* __restrict__ scale,
                                     const float* __restrict__ shift, const float* __restrict__ W1,
                                     const float* __restrict__ W2, const float* __restrict__ hidden,
                                     const float* __restrict__ gate, float* __restrict__ work,
                                     float* __restrict__ dz2_out /*[N][Cp]*/, float* __restrict__ dz1_out /*[N][sw]*/) {
  x3d::pdl_prologue();
  extern __shared__ float sm[];
  float* s_dz2 = sm;           // [C]
  float* s_dz1 = sm + C;       // [sw]
  float* s_h = sm + C + sw;    // [sw]
  const int n = blockIdx.x, b = n % splits;
  for (int j = threadIdx.x; j < sw; j += blockDim.x) s_h[j] = hidden[(int64_t)n * sw + j];
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    float dz2 = 0.f;
    if (c < C) {
      const double R1 = bwd_stats[((int64_t)n * Cp + c) * 2 + 0];
      const double R2 = bwd_stats[((int64_t)n * Cp + c) * 2 + 1];
      const float g = gate[(int64_t)n * Cp + c];
      // dgate = sum_thw dz*u = scale*sum(dz*a2) + shift*sum(dz)
      const float dgate = (float)((double)scale[b * Cp + c] * R2 + (double)shift[b * Cp + c] * R1);
      dz2 = dgate * g * (1.f - g);
      s_dz2[c] = dz2;
    }
    dz2_out[(int64_t)n * Cp + c] = dz2;
  }
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
  for (int j = warp; j < sw; j += nw) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(W2[(int64_t)c * sw + j], s_dz2[c], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float dz1 = s_h[j] > 0.f ? acc : 0.f;
      s_dz1[j] = dz1;
      dz1_out[(int64_t)n * sw + j] = dz1;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    float dpP = 0.f;
    if (c < C) {
      float dp = 0.f;
      for (int j = 0; j < sw; ++j) dp = fmaf(W1[(int64_t)j * C + c], s_dz1[j], dp);
      dpP = (float)((double)dp * inv_P);
    }
    work[(int64_t)n * Cp + c] = dpP;
  }
}

// SE parameter gradients: dW2[c][j] += sum_n dz2[n][c] h[n][j]; dW1[j][c] += sum_n dz1[n][j] p[n][c]; db2, db1.
// One thread per output element, samples added in order.
__global__ void se_param_grad_kernel(const float* __restrict__ dz2, const float* __restrict__ dz1,
                                     const float* __restrict__ pooled, const float* __restrict__ hidden, int N, int C,
                                     int Cp, int sw, int G, float* __restrict__ dW1, float* __restrict__ db1,
                                     float* __restrict__ dW2, float* __restrict__ db2) {
  x3d::pdl_prologue();
  // G (a power of two <= 32) consecutive lanes share one output: lane g takes samples g, g+G, ... (8 loads in flight per
  // operand), then a fixed shuffle tree adds the G partial sums -- deterministic, and a multigrid batch of 256 clips is
  // one round of loads instead of 32 dependent ones (22 us per SE block at 256 x 4 x 111^2).
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int o = t / G, g = t & (G - 1);
  const int nW = C * sw;
  auto dot = [&](const float* __restrict__ a, int64_t sa, const float* __restrict__ b, int64_t sb) {
    float acc = 0.f;
    int n = g;
    for (; n + 7 * G < N; n += 8 * G) {
      float av[8], bv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        av[u] = __ldg(a + (int64_t)(n + u * G) * sa);
        bv[u] = b ? __ldg(b + (int64_t)(n + u * G) * sb) : 1.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc = fmaf(av[u], bv[u], acc);
    }
    for (; n < N; n += G) acc = fmaf(__ldg(a + (int64_t)n * sa), b ? __ldg(b + (int64_t)n * sb) : 1.f, acc);
    return acc;
  };
  float v = 0.f;
  float* dst = nullptr;
  if (o < nW) {                                  // dW2[c][j]
    const int c = o / sw, j = o - c * sw;
    v = dot(dz2 + c, Cp, hidden + j, sw);
    dst = dW2 + o;
  } else if (o < 2 * nW) {                       // dW1[j][c]
    const int q = o - nW;
    const int j = q / C, c = q - j * C;
    v = dot(dz1 + j, sw, pooled + c, C);
    dst = dW1 + q;
  } else if (o < 2 * nW + C) {                   // db2[c]
    const int c = o - 2 * nW;
    v = dot(dz2 + c, Cp, nullptr, 0);
    dst = db2 + c;
  } else if (o < 2 * nW + C + sw) {              // db1[j]
    const int j = o - 2 * nW - C;
    v = dot(dz1 + j, sw, nullptr, 0);
    dst = db1 + j;
  }
  for (int off = G >> 1; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (dst != nullptr && g == 0) *dst += v;
}

// BN2 backward coefficients per (n, c): da2 = E1*dz + E2*a2 + E3, with du = dz*g + dpP.
__global__ void se_bn_bwd_coef_kernel(const double* __restrict__ fwd_stats, const double* __restrict__ bwd_stats,
                                      int N, int splits, double P, double inv_m, int C, int Cp,
                                      const float* __restrict__ gamma, const float* __restrict__ mean,
                                      const float* __restrict__ rstd, int train, const float* __restrict__ gate /*nullable*/, const float* __restrict__ work,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta,
                                      float* __restrict__ coef) {
  x3d::pdl_prologue();
  // one thread per (split, channel)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= splits * Cp) return;
  const int b = idx / Cp, c = idx % Cp;
  const int64_t NC = (int64_t)N * Cp;             // coef is planar: [3][N][Cp]
  if (c >= C) {
    for (int n = b; n < N; n += splits) {
      const int64_t o = (int64_t)n * Cp + c;
      coef[o] = 0.f; coef[NC + o] = 0.f; coef[2 * NC + o] = 0.f;
    }
    return;
  }
  const double gm = gamma[c];
  double s1 = 0.0, s2 = 0.0;   // sum du, sum du*a2 over the split (du = dz*gate + dp/P)
  if (gate) {
#pragma unroll 4
    for (int n = b; n < N; n += splits) {
      const int64_t o = (int64_t)n * Cp + c;
      const double g = gate[o], dpP = work[o];
      const double2 bs = __ldg(reinterpret_cast<const double2*>(bwd_stats) + o);
      s1 += g * bs.x + dpP * P;
      s2 += g * bs.y + dpP * __ldg(fwd_stats + o * 2);
    }
  } else {
    sum_split(bwd_stats, N, splits, b, c, Cp, s1, s2);
  }
  const double mu = mean[idx], r = rstd[idx];
  const double sxh = (s2 - mu * s1) * r;
  double A = gm * r, B = 0.0, Cc = 0.0;
  if (train) {
    const double m1 = s1 * inv_m, m2 = sxh * inv_m;   // 1/m from the host: no fp64 division
    B = -gm * r * r * m2;
    Cc = gm * r * (-m1 + mu * r * m2);
  }
  for (int n = b; n < N; n += splits) {
    const int64_t o = (int64_t)n * Cp + c;
    const double g = gate ? (double)gate[o] : 1.0;
    const double dpP = gate ? (double)work[o] : 0.0;
    coef[o] = (float)(A * g);
    coef[NC + o] = (float)B;
    coef[2 * NC + o] = (float)(Cc + A * dpP);
  }
  if (dgamma) atomicAdd(&dgamma[c], (float)sxh);
  if (dbeta) atomicAdd(&dbeta[c], (float)s1);
}

extern "C" int x3d_se_bn_bwd(const double* fwd_stats, const double* bwd_stats, int64_t N, int splits, int64_t P,
                             int64_t C, int64_t Cp, int sw, const float* gamma, const float* mean,
                             const float* rstd, const float* scale, const float* shift, int train,
                             const float* W1, const float* W2, const float* pooled, const float* hidden,
                             const float* gate, float* dW1, float* db1, float* dW2, float* db2, float* dgamma,
                             float* dbeta, float* work, float* coef, x3d_stream_t stream) {
  if (N == 0) return 0;
  if (gate) {
    size_t smem = (C + 2 * sw) * sizeof(float);
    float* dz2 = work + N * Cp;                 // work: [N][Cp] dp/P | [N][Cp] dz2 | [N][sw] dz1
    float* dz1 = dz2 + N * Cp;
    x3d::launch(se_bwd_sample_kernel, (unsigned)N, 256, smem, as_stream(stream), fwd_stats, bwd_stats, splits, 1.0 / (double)P, (int)C,
                (int)Cp, sw, scale, shift, W1, W2, hidden, gate, work, dz2, dz1);
    X3D_LAUNCH_CHECK();
    const int64_t outs = 2 * C * sw + C + sw;
    int G = 1;                                   // lanes per output: ~8 samples per lane
    while (G < 32 && G * 8 < N) G *= 2;
    x3d::launch(se_param_grad_kernel, (unsigned)cdiv(outs * G, 128), 128, 0, as_stream(stream), (const float*)dz2,
                (const float*)dz1, pooled, hidden, (int)N, (int)C, (int)Cp, sw, G, dW1, db1, dW2, db2);
    X3D_LAUNCH_CHECK();
  }
  x3d::launch(se_bn_bwd_coef_kernel, (unsigned)cdiv(splits * Cp, 64), 64, 0, as_stream(stream), fwd_stats, bwd_stats, (int)N, splits,
              (double)P, 1.0 / ((double)(N / splits) * (double)P), (int)C, (int)Cp, gamma, mean, rstd, train, gate, work,
              dgamma, dbeta, coef);
  X3D_LAUNCH_CHECK();
  return 0;
}

template <typename T, bool GATE>
__global__ void swish_gate_bwd_apply_kernel(const T* __restrict__ dv, const T* __restrict__ a2,
                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                            int splits, const float* __restrict__ gate, const float* __restrict__ coef,
                                            T* __restrict__ da2, int64_t P, int Cp, int cv, int rows, int64_t chunk) {
  x3d::pdl_prologue();
  ROW_PROLOGUE();
  const int b = n % splits;
  float sc[VEC], sh[VEC], E1[VEC], E2[VEC], E3[VEC];
  load_consts<VEC>(scale + b * Cp + c0, sc);
  load_consts<VEC>(shift + b * Cp + c0, sh);
  if (GATE) {
    float g[VEC];
    load_consts<VEC>(gate + (int64_t)n * Cp + c0, g);
#pragma unroll
    for (int j = 0; j < VEC; ++j) { sc[j] *= g[j]; sh[j] *= g[j]; }
  }
  const int64_t NC = (int64_t)gridDim.y * Cp;                 // coef is planar: [3][N][Cp]
  load_consts<VEC>(coef + 0 * NC + (int64_t)n * Cp + c0, E1);
  load_consts<VEC>(coef + 1 * NC + (int64_t)n * Cp + c0, E2);
  load_consts<VEC>(coef + 2 * NC + (int64_t)n * Cp + c0, E3);
  int64_t p = p0 + prow;
  for (; p + rows < p1; p += 2 * (int64_t)rows) {               // two rows per iteration: four loads in flight first
    const int64_t off0 = ((int64_t)n * P + p) * Cp + c0, off1 = off0 + (int64_t)rows * Cp;
    float d0[VEC], x0[VEC], d1[VEC], x1[VEC];
    load_vec<T>(dv + off0, d0);
    load_vec<T>(a2 + off0, x0);
    load_vec<T>(dv + off1, d1);
    load_vec<T>(a2 + off1, x1);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float dz0 = d0[j] * swish_grad<T>(fmaf(x0[j], sc[j], sh[j]));
      const float dz1 = d1[j] * swish_grad<T>(fmaf(x1[j], sc[j], sh[j]));
      d0[j] = fmaf(E1[j], dz0, fmaf(E2[j], x0[j], E3[j]));
      d1[j] = fmaf(E1[j], dz1, fmaf(E2[j], x1[j], E3[j]));
    }
    store_vec<T>(da2 + off0, d0);
    store_vec<T>(da2 + off1, d1);
  }
  for (; p < p1; p += rows) {
    const int64_t off = ((int64_t)n * P + p) * Cp + c0;
    float d[VEC], x[VEC];
    load_vec<T>(dv + off, d);
    load_vec<T>(a2 + off, x);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float dz = d[j] * swish_grad<T>(fmaf(x[j], sc[j], sh[j]));
      d[j] = fmaf(E1[j], dz, fmaf(E2[j], x[j], E3[j]));
    }
    store_vec<T>(da2 + off, d);
  }
}
extern "C" int x3d_swish_gate_bwd_apply(const void* dv, const void* a2, const float* scale, const float* shift,
                                        int splits, const float* gate, const float* coef, void* da2, int64_t N,
                                        int64_t P, int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  if (N * P == 0) return 0;
  X3D_DISPATCH_DTYPE(dt, {
    RowGeom g = make_row_geom<T>(N, P, Cp, 8 * kNumSMs);
    dim3 grid(g.chunks, (unsigned)N);
    if (gate)
      x3d::launch(swish_gate_bwd_apply_kernel<T, true>, grid, g.threads, 0, as_stream(stream), 
          (const T*)dv, (const T*)a2, scale, shift, splits, gate, coef, (T*)da2, P, (int)Cp, g.cv, g.rows, g.chunk);
    else
      x3d::launch(swish_gate_bwd_apply_kernel<T, false>, grid, g.threads, 0, as_stream(stream), 
          (const T*)dv, (const T*)a2, scale, shift, splits, nullptr, coef, (T*)da2, P, (int)Cp, g.cv, g.rows, g.chunk);
  });
  X3D_LAUNCH_CHECK();
  return 0;
}

// =======================================================================================
// head: relu(bn5) -> average pool ; backward into BN
// rows r = n (pool_t) or n*T+t; Pp positions per row.
// =======================================================================================
template <typename T>
__global__ void bn_relu_pool_fwd_kernel(const T* __restrict__ a, const float* __restrict__ scale,
                                        const float* __restrict__ shift, int splits, float* __restrict__ pooled,
                                        int rows_per_sample, int64_t P /*positions per row*/, int C, int Cp, int cv,
                                        int rows, int64_t chunk) {
  x3d::pdl_prologue();
  extern __shared__ float s_acc[];
  ROW_PROLOGUE();   // here "n" is the pooled row r
  const int b = (n / rows_per_sample) % splits;
  float sc[VEC], sh[VEC], a0[VEC], a1[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    sc[j] = scale[b * Cp + c0 + j];
    sh[j] = shift[b * Cp + c0 + j];
    a0[j] = 0.f; a1[j] = 0.f;
  }
  int64_t p = p0 + prow;
  for (; p + 3 * (int64_t)rows < p1; p += 4 * (int64_t)rows) {        // 4 independent 16-byte loads in flight
    float x[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) load_vec<T>(a + ((int64_t)n * P + p + (int64_t)u * rows) * Cp + c0, x[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < VEC; ++j) a0[j] += fmaxf(fmaf(x[u][j], sc[j], sh[j]), 0.f);
  }
  for (; p < p1; p += rows) {
    const int64_t off = ((int64_t)n * P + p) * Cp + c0;
    float x[VEC];
    load_vec<T>(a + off, x);
#pragma unroll
    for (int j = 0; j < VEC; ++j) a0[j] += fmaxf(fmaf(x[j], sc[j], sh[j]), 0.f);
  }
  // ordered block reduction (no atomics: the pooled features feed fc1/fc2 and, through the loss, every gradient of
  // the network -- last-bit noise here is amplified ~1e4x by the train-mode BN backward chain):
  // every thread parks its partial sums in its own slot, then one thread per channel adds the `rows` slots in order.
  // One CTA per pooled row, so the result is written, not accumulated.
#pragma unroll
  for (int j = 0; j < VEC; ++j) s_acc[(size_t)prow * Cp + c0 + j] = a0[j];
  __syncthreads();
  const float inv = 1.f / (float)P;
  const int nrows = blockDim.x / cv;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float v = 0.f;
    for (int r = 0; r < nrows; ++r) v += s_acc[(size_t)r * Cp + i];
    pooled[(int64_t)n * C + i] = v * inv;
  }
  (void)a1;
}
extern "C" int x3d_bn_relu_pool_fwd(const void* a5, const float* scale, const float* shift, int splits,
                                    float* pooled, int64_t N, int64_t T_, int64_t HW, int pool_t, int64_t C,
                                    int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  const int64_t R = pool_t ? N : N * T_;
  const int64_t Pp = pool_t ? T_ * HW : HW;
  if (R * Pp == 0) return 0;
  X3D_DISPATCH_DTYPE(dt, {
    RowGeom g = make_row_geom<T>(R, Pp, Cp);
    g.chunk = Pp;                      // ONE CTA per pooled row: deterministic, no cross-CTA accumulation
    g.chunks = 1;
    g.rows = 1024 / g.cv;              // as many position rows as a CTA holds (the grid is only R CTAs)
    if ((int64_t)g.rows > Pp) g.rows = (int)Pp;
    if (g.rows < 1) g.rows = 1;
    g.threads = g.cv * g.rows;
    dim3 grid(1, (unsigned)R);
    x3d::launch(bn_relu_pool_fwd_kernel<T>, grid, g.threads, (size_t)g.rows * Cp * sizeof(float), as_stream(stream),
        (const T*)a5, scale, shift, splits, pooled, pool_t ? 1 : (int)T_, Pp, (int)C, (int)Cp, g.cv, g.rows, g.chunk);
  });
  X3D_LAUNCH_CHECK();
  return 0;
}

template <typename T, bool APPLY>
__global__ void bn_relu_pool_bwd_kernel(const T* __restrict__ a, const float* __restrict__ scale,
                                        const float* __restrict__ shift, int splits,
                                        const float* __restrict__ dpooled, double* __restrict__ stats,
                                        const float* __restrict__ coef, T* __restrict__ da, int rows_per_sample,
                                        int64_t P, int C, int Cp, int cv, int rows, int64_t chunk) {
  x3d::pdl_prologue();
  extern __shared__ float s_acc[];
  ROW_PROLOGUE();   // n = pooled row r
  const int ns = n / rows_per_sample;   // sample index
  const int b = ns % splits;
  const int sC = splits * Cp;
  float sc[VEC], sh[VEC], dpl[VEC], a0[VEC], a1[VEC], A[VEC], B[VEC], Cc[VEC];
  const float inv = 1.f / (float)P;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    sc[j] = scale[b * Cp + c0 + j];
    sh[j] = shift[b * Cp + c0 + j];
    dpl[j] = (c0 + j < C) ? dpooled[(int64_t)n * C + c0 + j] * inv : 0.f;
    a0[j] = a1[j] = 0.f;
    if (APPLY) {
      A[j] = coef[0 * sC + b * Cp + c0 + j];
      B[j] = coef[1 * sC + b * Cp + c0 + j];
      Cc[j] = coef[2 * sC + b * Cp + c0 + j];
    }
  }
  for (int64_t p = p0 + prow; p < p1; p += rows) {
    const int64_t off = ((int64_t)n * P + p) * Cp + c0;
    float x[VEC];
    load_vec<T>(a + off, x);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float dp = (fmaf(x[j], sc[j], sh[j]) > 0.f) ? dpl[j] : 0.f;
      if (APPLY) {
        x[j] = fmaf(A[j], dp, fmaf(B[j], x[j], Cc[j]));
      } else {
        a0[j] += dp;
        a1[j] = fmaf(dp, x[j], a1[j]);
      }
    }
    if (APPLY) store_vec<T>(da + off, x);
  }
  if (!APPLY) block_stats_flush<VEC>(a0, a1, cvec, Cp, s_acc, stats + (int64_t)ns * Cp * 2);
}
extern "C" int x3d_bn_relu_pool_bwd_reduce(const void* a5, const float* scale, const float* shift, int splits,
                                           const float* dpooled, double* stats, int64_t N, int64_t T_, int64_t HW,
                                           int pool_t, int64_t C, int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  const int64_t R = pool_t ? N : N * T_;
  const int64_t Pp = pool_t ? T_ * HW : HW;
  if (R * Pp == 0) return 0;
  X3D_DISPATCH_DTYPE(dt, {
    RowGeom g = make_row_geom<T>(R, Pp, Cp);
    dim3 grid(g.chunks, (unsigned)R);
    x3d::launch(bn_relu_pool_bwd_kernel<T, false>, grid, g.threads, (size_t)g.rows * Cp * 2 * sizeof(float), as_stream(stream), 
        (const T*)a5, scale, shift, splits, dpooled, stats, nullptr, nullptr, pool_t ? 1 : (int)T_, Pp, (int)C, (int)Cp,
        g.cv, g.rows, g.chunk);
  });
  X3D_LAUNCH_CHECK();
  return 0;
}
extern "C" int x3d_bn_relu_pool_bwd_apply(const void* a5, const float* scale, const float* shift, int splits,
                                          const float* dpooled, const float* coef, void* da5, int64_t N, int64_t T_,
                                          int64_t HW, int pool_t, int64_t C, int64_t Cp, x3d_dtype_t dt,
                                          x3d_stream_t stream) {
  X3D_CHECK_ARG(Cp % 8 == 0, "Cp % 8");
  const int64_t R = pool_t ? N : N * T_;
  const int64_t Pp = pool_t ? T_ * HW : HW;
  if (R * Pp == 0) return 0;
  X3D_DISPATCH_DTYPE(dt, {
    RowGeom g = make_row_geom<T>(R, Pp, Cp, 8 * kNumSMs);
    dim3 grid(g.chunks, (unsigned)R);
    x3d::launch(bn_relu_pool_bwd_kernel<T, true>, grid, g.threads, Cp * 2 * sizeof(float), as_stream(stream), 
        (const T*)a5, scale, shift, splits, dpooled, nullptr, coef, (T*)da5, pool_t ? 1 : (int)T_, Pp, (int)C, (int)Cp,
        g.cv, g.rows, g.chunk);
  });
  X3D_LAUNCH_CHECK();
  return 0;
}

// =======================================================================================
// small dense fp32 GEMM (head fc1/fc2 and their gradients): 32x32x32 smem tiles
// =======================================================================================
// 64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread; either operand may be k- or row-contiguous (the four head
// GEMM flavours NT / NN / TN of fc1 / fc2 and their gradients) -- the tile loaders pick the coalesced orientation.
// No split-K: deterministic.  (The multigrid shapes put 64-256 clips on a GPU, so M = batch is no longer "skinny".)
__global__ void __launch_bounds__(256) small_gemm_kernel(const float* __restrict__ A, int64_t sai, int64_t sak,
                                                         const float* __restrict__ B, int64_t sbk, int64_t sbj,
                                                         float* __restrict__ C, int64_t ldc, int M, int Nn, int K,
                                                         const float* __restrict__ bias, int relu,
                                                         const float* __restrict__ mul, int accumulate) {
  x3d::pdl_prologue();
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ __align__(16) float As[TK][TM + 4];  // [k][i]
  __shared__ __align__(16) float Bs[TK][TN + 4];  // [k][j]
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += TK) {
    // 64 x 16 elements per operand = 4 per thread; the fastest thread index follows the unit-stride dimension
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int q = tid + e * 256;
      int i, k;
      if (sak == 1) { k = q % TK; i = q / TK; } else { i = q % TM; k = q / TM; }
      As[k][i] = (i0 + i < M && k0 + k < K) ? __ldg(&A[(int64_t)(i0 + i) * sai + (int64_t)(k0 + k) * sak]) : 0.f;
      int j, kb;
      if (sbk == 1) { kb = q % TK; j = q / TK; } else { j = q % TN; kb = q / TN; }
      Bs[kb][j] = (j0 + j < Nn && k0 + kb < K) ? __ldg(&B[(int64_t)(k0 + kb) * sbk + (int64_t)(j0 + j) * sbj]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ii = i0 + ty * 4 + i;
    if (ii >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jj = j0 + tx * 4 + j;
      if (jj >= Nn) continue;
      float v = acc[i][j] + (bias ? bias[jj] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      if (mul) v *= mul[(int64_t)ii * Nn + jj];
      float* dst = &C[(int64_t)ii * ldc + jj];
      *dst = accumulate ? (*dst + v) : v;
    }
  }
}
// Split-K flavour of the tile kernel (the head GEMMs have 7 - 32 column tiles and K up to 2048: without a K split a
// handful of CTAs walk 128 dependent k-tiles -- fc1 dgrad at 64 rows took 228 us for 0.1 GFLOP).  grid.z CTAs share a
// tile, each reduces one K slice (operands of the next k-tile are fetched into registers while the current one is
// multiplied) and parks its 64 x 64 partial in the caller's workspace; the CTA that arrives last at the tile's ticket adds
// the partials IN SLICE ORDER (deterministic) and runs the epilogue.  Tickets are left at zero for the next call.
__global__ void __launch_bounds__(256)
small_gemm_splitk_kernel(const float* __restrict__ A, int64_t sai, int64_t sak, const float* __restrict__ B, int64_t sbk,
                         int64_t sbj, float* __restrict__ C, int64_t ldc, int M, int Nn, int K, int kchunk,
                         const float* __restrict__ bias, int relu, const float* __restrict__ mul, int accumulate,
                         float* __restrict__ part, unsigned* __restrict__ tickets) {
  x3d::pdl_prologue();
  // A K slice is consumed in stages of 64: ALL 32 loads of a thread for a stage are issued back to back (the kernel is
  // latency-bound: 3.5 MB of weights, a few microseconds), the next stage's loads fly while this one is multiplied.
  constexpr int TM = 64, TN = 64, TK = 64, PER = TM * TK / 256;
  __shared__ __align__(16) float As[TK][TM + 4];  // [k][i]
  __shared__ __align__(16) float Bs[TK][TN + 4];  // [k][j]
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  const int S = gridDim.z, sl = blockIdx.z;
  const int kbeg = sl * kchunk, kend = min(K, kbeg + kchunk);
  // element e of a thread: q = tid + 256 e; the fastest thread index follows the operand's unit-stride dimension
  const bool a_kfast = sak == 1, b_kfast = sbk == 1;
  float ra[PER], rb[PER];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int q = tid + e * 256;
      const int ai = a_kfast ? q / TK : q % TM, ak = a_kfast ? q % TK : q / TM;
      const int bj = b_kfast ? q / TK : q % TN, bk = b_kfast ? q % TK : q / TN;
      ra[e] = (i0 + ai < M && k0 + ak < kend) ? __ldg(&A[(int64_t)(i0 + ai) * sai + (int64_t)(k0 + ak) * sak]) : 0.f;
      rb[e] = (j0 + bj < Nn && k0 + bk < kend) ? __ldg(&B[(int64_t)(k0 + bk) * sbk + (int64_t)(j0 + bj) * sbj]) : 0.f;
    }
  };
  auto park = [&]() {
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int q = tid + e * 256;
      const int ai = a_kfast ? q / TK : q % TM, ak = a_kfast ? q % TK : q / TM;
      const int bj = b_kfast ? q / TK : q % TN, bk = b_kfast ? q % TK : q / TN;
      As[ak][ai] = ra[e];
      Bs[bk][bj] = rb[e];
    }
  };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
    park();
    __syncthreads();
    if (k0 + TK < kend) fetch(k0 + TK);
    const int kn = min(TK, kend - k0);
#pragma unroll 8
    for (int k = 0; k < kn; ++k) {
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
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  if (S > 1) {
    float4* mine = reinterpret_cast<float4*>(part + ((size_t)sl * gridDim.x * gridDim.y + tile) * (TM * TN)) + tid * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) __stcg(mine + i, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&tickets[tile], 1u) == (unsigned)(S - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int s2 = 0; s2 < S; ++s2) {       // fixed order: bit-identical from run to run
      const float4* src = reinterpret_cast<const float4*>(part + ((size_t)s2 * gridDim.x * gridDim.y + tile) * (TM * TN)) + tid * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldcg(src + i);
        acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
      }
    }
    if (tid == 0) tickets[tile] = 0u;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ii = i0 + ty * 4 + i;
    if (ii >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jj = j0 + tx * 4 + j;
      if (jj >= Nn) continue;
      float v = acc[i][j] + (bias ? bias[jj] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      if (mul) v *= mul[(int64_t)ii * Nn + jj];
      float* dst = &C[(int64_t)ii * ldc + jj];
      *dst = accumulate ? (*dst + v) : v;
    }
  }
}
constexpr int64_t kGemmTicketBytes = 4096;     // 1024 tile tickets at the head of the workspace

// skinny variants for M <= 32 rows (the head: M = batch).  The weight matrix (N x K, 3.5 MB for fc1 / fc2 of X3D-M) is the
// only real traffic; what matters is that enough CTAs pull on it at once and that no thread walks a long serial K loop
// (fc1 dgrad with 14 CTAs x 128 dependent iterations took 115 us for 3.5 MB).  Both variants are deterministic.
// (a) B k-contiguous (sbk == 1): a warp owns JW output columns and one of KW slices of K (lanes stride through the
//     slice, shuffle reduction); the KW slice results of a column group are added in a fixed order through shared memory
template <int MR>
__global__ void __launch_bounds__(256)
skinny_gemm_kcontig_kernel(const float* __restrict__ A, int64_t sai, int64_t sak, const float* __restrict__ B, int64_t sbj,
                           float* __restrict__ C, int64_t ldc, int M, int Nn, int K, int KW, const float* __restrict__ bias,
                           int relu, const float* __restrict__ mul, int accumulate) {
  x3d::pdl_prologue();
  constexpr int JW = MR <= 16 ? 4 : 2;
  __shared__ float red[8][JW * MR];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gpb = 8 / KW;                                       // column groups per block
  const int grp = wib / KW, ks = wib - grp * KW;
  const int j0 = (blockIdx.x * gpb + grp) * JW;
  const int kchunk = ((K + KW - 1) / KW + 31) / 32 * 32;
  const int kbeg = ks * kchunk, kend = min(K, kbeg + kchunk);
  float acc[JW][MR];
#pragma unroll
  for (int q = 0; q < JW; ++q)
#pragma unroll
    for (int i = 0; i < MR; ++i) acc[q][i] = 0.f;
  if (j0 < Nn) {
#pragma unroll 2
    for (int k = kbeg + lane; k < kend; k += 32) {
      float a[MR], b[JW];
#pragma unroll
      for (int q = 0; q < JW; ++q) b[q] = (j0 + q < Nn) ? __ldg(&B[(int64_t)(j0 + q) * sbj + k]) : 0.f;
#pragma unroll
      for (int i = 0; i < MR; ++i) a[i] = (i < M) ? __ldg(&A[(int64_t)i * sai + (int64_t)k * sak]) : 0.f;
#pragma unroll
      for (int q = 0; q < JW; ++q)
#pragma unroll
        for (int i = 0; i < MR; ++i) acc[q][i] = fmaf(a[i], b[q], acc[q][i]);
    }
  }
#pragma unroll
  for (int q = 0; q < JW; ++q)
#pragma unroll
    for (int i = 0; i < MR; ++i) {
#pragma unroll
      for (int o = 16; o; o >>= 1) acc[q][i] += __shfl_xor_sync(0xffffffffu, acc[q][i], o);
    }
  // every lane now holds all JW*MR sums of its warp; lane e keeps element e (and e+32, ...)
#pragma unroll
  for (int q = 0; q < JW; ++q)
#pragma unroll
    for (int i = 0; i < MR; ++i)
      if (((q * MR + i) & 31) == lane) red[wib][q * MR + i] = acc[q][i];
  __syncthreads();
  for (int e = threadIdx.x; e < gpb * JW * MR; e += 256) {
    const int g = e / (JW * MR), r = e - g * (JW * MR);
    const int q = r / MR, i = r - q * MR;
    const int j = (blockIdx.x * gpb + g) * JW + q;
    if (j >= Nn || i >= M) continue;
    float v = 0.f;
    for (int s2 = 0; s2 < KW; ++s2) v += red[g * KW + s2][r];
    v += bias ? bias[j] : 0.f;
    if (relu) v = fmaxf(v, 0.f);
    if (mul) v *= mul[(int64_t)i * Nn + j];
    float* dst = &C[(int64_t)i * ldc + j];
    *dst = accumulate ? (*dst + v) : v;
  }
}
// (b) B j-contiguous (sbj == 1): block = CW output columns x 512 / CW k-slices; every thread walks its slice of K for its
//     column, the slices are added in a fixed order through shared memory -- no atomics: dh / dpooled of the head
//     backward feed the whole backward pass (see bn_relu_pool_fwd_kernel).  CW = 8 when 32-column blocks would leave
//     most SMs idle (a 32-byte row segment per k is still a full DRAM sector).
template <int MR, int CW>
__global__ void __launch_bounds__(512)
skinny_gemm_jcontig_kernel(const float* __restrict__ A, int64_t sai, int64_t sak, const float* __restrict__ B, int64_t sbk,
                           float* __restrict__ C, int64_t ldc, int M, int Nn, int K, int accumulate) {
  x3d::pdl_prologue();
  constexpr int SL = (MR <= 16 ? 512 : 256) / CW;              // k slices (shared scratch <= 37 KB)
  __shared__ float red[SL][MR][CW + 1];
  const int tx = threadIdx.x % CW, ty = threadIdx.x / CW;      // CW x SL
  const int j = blockIdx.x * CW + tx;
  const int kchunk = (K + SL - 1) / SL;
  const int k0 = ty * kchunk;
  const int k1 = k0 + kchunk < K ? k0 + kchunk : K;
  float acc[MR];
#pragma unroll
  for (int i = 0; i < MR; ++i) acc[i] = 0.f;
  if (j < Nn && ty < SL) {
#pragma unroll 2
    for (int k = k0; k < k1; ++k) {
      const float b = __ldg(&B[(int64_t)k * sbk + j]);
#pragma unroll
      for (int i = 0; i < MR; ++i)
        if (i < M) acc[i] = fmaf(__ldg(&A[(int64_t)i * sai + (int64_t)k * sak]), b, acc[i]);
    }
  }
  if (ty < SL) {
#pragma unroll
    for (int i = 0; i < MR; ++i) red[ty][i][tx] = acc[i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < MR * CW; e += blockDim.x) {
    const int i = e / CW, c = e % CW, jj = blockIdx.x * CW + c;
    if (i < M && jj < Nn) {
      float v = 0.f;
#pragma unroll 8
      for (int q = 0; q < SL; ++q) v += red[q][i][c];
      float* dst = &C[(int64_t)i * ldc + jj];
      *dst = accumulate ? (*dst + v) : v;
    }
  }
}

// slices of K for the split-K kernel: enough CTAs for two per SM, at least 64 k per slice, capped by the workspace
static int splitk_plan(int64_t M, int64_t Nn, int64_t K, int64_t ws_bytes, int* kchunk) {
  const int64_t tiles = cdiv(M, 64) * cdiv(Nn, 64);
  int64_t S = cdiv(2 * kNumSMs, tiles);
  if (S > K / 64) S = K / 64;
  if (S > 16) S = 16;             // the last CTA of a tile adds the slices alone: 16 x 16 KB is ~2 us of L2 reads
  const int64_t fit = (ws_bytes - kGemmTicketBytes) / (tiles * 64 * 64 * (int64_t)sizeof(float));
  if (S > fit) S = fit;
  if (S < 1 || tiles > kGemmTicketBytes / 4) S = 1;
  int64_t kc = cdiv(cdiv(K, S), 16) * 16;
  if (kc < 16) kc = 16;
  *kchunk = (int)kc;
  return (int)cdiv(K, kc);
}
extern "C" int64_t x3d_small_gemm_workspace_bytes(int64_t M, int64_t Nn, int64_t K) {
  const int64_t tiles = cdiv(M, 64) * cdiv(Nn, 64);
  int64_t S = cdiv(2 * kNumSMs, tiles);
  if (S > K / 64) S = K / 64;
  if (S > 16) S = 16;
  if (S < 1) S = 1;
  return kGemmTicketBytes + S * tiles * 64 * 64 * (int64_t)sizeof(float);
}
// Dispatch (micro-benchmarked per flavour, tools/head_gemm_microbench.py): batch <= 32 rows -> the skinny kernels (except a
// j-contiguous weight with K > 1024, fc1 dgrad: 49 us skinny vs 20 us split-K); enough tiles to fill the machine or a
// K of one or two k-tiles (the weight gradients, K = batch) -> the plain tile kernel; otherwise split-K when a workspace
// was given.
static int small_gemm_impl(const float* A, int64_t sai, int64_t sak, const float* B, int64_t sbk, int64_t sbj, float* C,
                           int64_t ldc, int64_t M, int64_t Nn, int64_t K, const float* bias, int relu, const float* mul,
                           int accumulate, void* ws, int64_t ws_bytes, x3d_stream_t stream) {
  if (M == 0 || Nn == 0) return 0;
  const int64_t tiles = cdiv(M, 64) * cdiv(Nn, 64);
  const bool can_split = ws != nullptr && tiles < kNumSMs && K > 64;
  if (M <= 32 && sbk == 1) {
    const int JW = M <= 16 ? 4 : 2;
    const int64_t groups = cdiv(Nn, JW);
    int KW = 1;                                                  // slices of K per column group: aim at >= 8 warps per SM
    while (KW < 8 && groups * KW < 8 * kNumSMs && K / (2 * KW) >= 64) KW *= 2;
    const unsigned grid = (unsigned)cdiv(groups, 8 / KW);
#define SK_(MR) x3d::launch(skinny_gemm_kcontig_kernel<MR>, grid, 256, 0, as_stream(stream),  \
      A, sai, sak, B, sbj, C, ldc, (int)M, (int)Nn, (int)K, KW, bias, relu, mul, accumulate)
    if (M <= 8) SK_(8); else if (M <= 16) SK_(16); else SK_(32);
#undef SK_
  } else if (M <= 32 && sbj == 1 && !bias && !relu && !mul && !(can_split && K > 1024)) {
    const bool narrow = cdiv(Nn, 32) < kNumSMs;
#define SJ_(MR, CWv) x3d::launch(skinny_gemm_jcontig_kernel<MR, CWv>, (unsigned)cdiv(Nn, CWv), (MR <= 16 ? 512 : 256), 0,  \
      as_stream(stream), A, sai, sak, B, sbk, C, ldc, (int)M, (int)Nn, (int)K, accumulate)
#define SJ2_(MR) do { if (narrow) SJ_(MR, 8); else SJ_(MR, 32); } while (0)
    if (M <= 8) SJ2_(8); else if (M <= 16) SJ2_(16); else SJ2_(32);
#undef SJ2_
#undef SJ_
  } else if (can_split) {
    int kchunk = 16;
    const int S = splitk_plan(M, Nn, K, ws_bytes, &kchunk);
    dim3 grid((unsigned)cdiv(Nn, 64), (unsigned)cdiv(M, 64), (unsigned)S);
    x3d::launch(small_gemm_splitk_kernel, grid, 256, 0, as_stream(stream), A, sai, sak, B, sbk, sbj, C, ldc, (int)M, (int)Nn,
                (int)K, kchunk, bias, relu, mul, accumulate,
                reinterpret_cast<float*>(static_cast<char*>(ws) + kGemmTicketBytes), static_cast<unsigned*>(ws));
  } else {
    dim3 grid((unsigned)cdiv(Nn, 64), (unsigned)cdiv(M, 64));
    x3d::launch(small_gemm_kernel, grid, 256, 0, as_stream(stream), A, sai, sak, B, sbk, sbj, C, ldc, (int)M, (int)Nn, (int)K,
                                                            bias, relu, mul, accumulate);
  }
  X3D_LAUNCH_CHECK();
  return 0;
}
extern "C" int x3d_small_gemm_ws(const float* A, int64_t sai, int64_t sak, const float* B, int64_t sbk, int64_t sbj,
                                 float* C, int64_t ldc, int64_t M, int64_t Nn, int64_t K, const float* bias, int relu,
                                 const float* mul, int accumulate, void* ws, int64_t ws_bytes, x3d_stream_t stream) {
  X3D_CHECK_ARG(ws != nullptr && ws_bytes >= kGemmTicketBytes && (reinterpret_cast<uintptr_t>(ws) & 15) == 0,
                "workspace must be 16-byte aligned and hold at least the 4 KB ticket area (zero-filled once by the caller)");
  return small_gemm_impl(A, sai, sak, B, sbk, sbj, C, ldc, M, Nn, K, bias, relu, mul, accumulate, ws, ws_bytes, stream);
}
extern "C" int x3d_small_gemm(const float* A, int64_t sai, int64_t sak, const float* B, int64_t sbk, int64_t sbj,
                              float* C, int64_t ldc, int64_t M, int64_t Nn, int64_t K, const float* bias, int relu,
                              const float* mul, int accumulate, x3d_stream_t stream) {
  return small_gemm_impl(A, sai, sak, B, sbk, sbj, C, ldc, M, Nn, K, bias, relu, mul, accumulate, nullptr, 0, stream);
}

__global__ void colsum_kernel(const float* __restrict__ src, int M, int Nn, float* __restrict__ dst) {
  x3d::pdl_prologue();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Nn) return;
  float acc = 0.f;
  for (int i = 0; i < M; ++i) acc += src[(int64_t)i * Nn + j];
  dst[j] += acc;
}
extern "C" int x3d_colsum(const float* src, int64_t M, int64_t Nn, float* dst, x3d_stream_t stream) {
  if (Nn == 0) return 0;
  x3d::launch(colsum_kernel, (unsigned)cdiv(Nn, 128), 128, 0, as_stream(stream), src, (int)M, (int)Nn, dst);
  X3D_LAUNCH_CHECK();
  return 0;
}

__global__ void relu_mask_mul_kernel(const float* __restrict__ src, const float* __restrict__ ref,
                                     const float* __restrict__ mul, float* __restrict__ dst, int64_t n) {
  x3d::pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = ref[i] > 0.f ? src[i] : 0.f;
    if (mul) v *= mul[i];
    dst[i] = v;
  }
}
extern "C" int x3d_relu_mask_mul(const float* src, const float* ref, const float* mul, float* dst, int64_t numel,
                                 x3d_stream_t stream) {
  if (numel == 0) return 0;
  int64_t blocks = cdiv(numel, 256);
  if (blocks > 1024) blocks = 1024;
  x3d::launch(relu_mask_mul_kernel, (unsigned)blocks, 256, 0, as_stream(stream), src, ref, mul, dst, numel);
  X3D_LAUNCH_CHECK();
  return 0;
}

// =======================================================================================
// fused SGD (momentum, weight decay; torch.optim.SGD semantics, dampening 0, no nesterov)
// =======================================================================================
// Work is cut into 4096-element blocks over ALL tensors (prefix sum over the descriptor table, in shared memory) and a
// machine-sized grid walks them: the old (64, n_desc) grid launched ~20 k CTAs of which most found nothing to do, and
// the two fc weights (0.9 M elements each) were 54 dependent scalar iterations per thread -- 58 us for 76 MB.
constexpr int kSgdBlock = 4096;
constexpr int kSgdMaxDesc = 2048;
__global__ void __launch_bounds__(256)
sgd_kernel(const x3d_sgd_desc_t* __restrict__ descs, int n_desc, float lr, float momentum, float wd, float grad_scale,
           int first_step, const float* __restrict__ hyper) {
  x3d::pdl_prologue();
  __shared__ int s_cum[kSgdMaxDesc + 1];           // s_cum[d] = blocks before descriptor d
  __shared__ int s_warp[8];
  __shared__ int s_run;
  if (hyper) { lr = hyper[0]; momentum = hyper[1]; wd = hyper[2]; grad_scale = hyper[3]; }
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_run = 0;
  __syncthreads();
  for (int base = 0; base < n_desc; base += 256) {
    const int d = base + tid;
    const int nb = d < n_desc ? (int)((descs[d].numel + kSgdBlock - 1) / kSgdBlock) : 0;
    int inc = nb;                                   // inclusive scan over the block
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    int before = s_run;
    for (int w = 0; w < wid; ++w) before += s_warp[w];
    if (d < n_desc) s_cum[d] = before + inc - nb;
    __syncthreads();
    if (tid == 255) s_run = before + inc;
    __syncthreads();
  }
  const int total = s_run;
  if (tid == 0) s_cum[n_desc] = total;
  __syncthreads();
  for (int vb = blockIdx.x; vb < total; vb += gridDim.x) {
    int lo = 0, hi = n_desc;                        // last descriptor with s_cum[d] <= vb (empty tensors share a value)
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_cum[mid] <= vb) lo = mid; else hi = mid;
    }
    const x3d_sgd_desc_t d = descs[lo];
    const int64_t e0 = (int64_t)(vb - s_cum[lo]) * kSgdBlock;
    const int64_t e1 = e0 + kSgdBlock < d.numel ? e0 + kSgdBlock : d.numel;
    auto upd = [&](float p, float g, float m, float& mo) {
      g = fmaf(wd, p, g * grad_scale);
      if (momentum != 0.f) {
        mo = first_step ? g : fmaf(momentum, m, g);
        g = mo;
      }
      return fmaf(-lr, g, p);
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(d.param) | reinterpret_cast<uintptr_t>(d.grad) |
                       reinterpret_cast<uintptr_t>(d.momentum_buf)) & 15) == 0;
    int64_t i = e0;
    if (vec) {
      for (int64_t q = e0 + (int64_t)tid * 4; q + 4 <= e1; q += 256 * 4) {
        float4 p = *reinterpret_cast<const float4*>(d.param + q);
        const float4 g = *reinterpret_cast<const float4*>(d.grad + q);
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
        if (momentum != 0.f && !first_step) m = *reinterpret_cast<const float4*>(d.momentum_buf + q);
        p.x = upd(p.x, g.x, m.x, m.x); p.y = upd(p.y, g.y, m.y, m.y);
        p.z = upd(p.z, g.z, m.z, m.z); p.w = upd(p.w, g.w, m.w, m.w);
        if (momentum != 0.f) *reinterpret_cast<float4*>(d.momentum_buf + q) = m;
        *reinterpret_cast<float4*>(d.param + q) = p;
      }
      i = e0 + ((e1 - e0) & ~(int64_t)3);            // scalar tail
    }
    for (int64_t q = i + tid; q < e1; q += 256) {
      const float p = d.param[q];
      float m = (momentum != 0.f && !first_step) ? d.momentum_buf[q] : 0.f;
      const float pn = upd(p, d.grad[q], m, m);
      if (momentum != 0.f) d.momentum_buf[q] = m;
      d.param[q] = pn;
    }
  }
}
static int sgd_launch(const x3d_sgd_desc_t* descs_dev, int n_desc, int64_t max_numel, float lr, float momentum, float wd,
                      float grad_scale, int first_step, const float* hyper, x3d_stream_t stream) {
  X3D_CHECK_ARG(n_desc <= kSgdMaxDesc, "at most 2048 tensors per x3d_sgd_step call");
  int64_t grid = cdiv(max_numel, kSgdBlock) * (int64_t)n_desc;   // upper bound on the blocks of work
  if (grid > 4 * kNumSMs) grid = 4 * kNumSMs;
  if (grid < 1) grid = 1;
  x3d::launch(sgd_kernel, (unsigned)grid, 256, 0, as_stream(stream), descs_dev, n_desc, lr, momentum, wd, grad_scale, first_step,
              hyper);
  X3D_LAUNCH_CHECK();
  return 0;
}
extern "C" int x3d_sgd_step(const x3d_sgd_desc_t* descs_dev, int n_desc, int64_t max_numel, float lr,
                            float momentum, float weight_decay, float grad_scale, int first_step,
                            x3d_stream_t stream) {
  if (n_desc == 0) return 0;
  return sgd_launch(descs_dev, n_desc, max_numel, lr, momentum, weight_decay, grad_scale, first_step, nullptr, stream);
}
extern "C" int x3d_sgd_step_dev(const x3d_sgd_desc_t* descs_dev, int n_desc, int64_t max_numel, const float* hyper_dev,
                                int first_step, x3d_stream_t stream) {
  if (n_desc == 0) return 0;
  X3D_CHECK_ARG(hyper_dev != nullptr, "hyper_dev");
  return sgd_launch(descs_dev, n_desc, max_numel, 0.f, 0.f, 0.f, 1.f, first_step, hyper_dev, stream);
}
