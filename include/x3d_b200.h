/*
 * x3d_b200.h -- C ABI of libx3d_b200.so: hand-written sm_100a kernels for the X3D
 * training hot path (reference: KiyoshiKAWASAKI/X3D-Multigrid, x3d.py).
 *
 * The reference has no FFI of its own: its hot path is the chain of torch.nn calls
 * in x3d.py, executed by ATen/cuDNN.  Each entry point below replaces the ATen
 * kernel(s) behind one group of those calls (cited as x3d.py:line).  A binding
 * needs only ctypes/cffi: plain pointers, integers, a CUDA stream handle.
 *
 * Conventions
 *  - All activation tensors are NDHWC ("channels last"), dense, channel count
 *    padded to Cp = round_up(C, 8); pad lanes hold zeros.  `x3d_dtype_t` selects
 *    the storage type of activations (fp32 or bf16); all accumulation, all BN
 *    statistics and all parameter gradients are fp32/fp64.
 *  - "P" is positions per sample (T*H*W); split-BN split of sample n is n % splits
 *    (x3d.py:50: view(n//s, c*s, ...)).
 *  - Per-sample statistic buffers are double[N][Cp][2]; they are ACCUMULATED into
 *    (callers zero them).  Parameter-gradient outputs are fp32 in the PyTorch
 *    parameter layout and are ACCUMULATED into as well.
 *  - Every function is asynchronous on `stream` (a cudaStream_t passed as void*),
 *    never allocates, is re-entrant, and returns 0 or a cudaError_t / negative
 *    own code; x3d_last_error() gives a thread-local message.
 */
#ifndef X3D_B200_H
#define X3D_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { X3D_F32 = 0, X3D_BF16 = 1 } x3d_dtype_t;
typedef void* x3d_stream_t;

const char* x3d_last_error(void);
int x3d_abi_version(void);
/* number of kernel launches issued through this library since load (all threads) */
int64_t x3d_launch_count(void);
/* Which kernel family served a call: every conv entry point below counts the path it took, so that callers
 * (and the tests) can assert that a hot shape ran the TMA-tiled / tcgen05 kernel and did not silently drop to
 * the shape-generic fallback.  Counters are process-wide and monotonic. */
typedef enum {
  X3D_PATH_DW_FWD_TILED = 0,
  X3D_PATH_DW_FWD_TEMPORAL = 1,
  X3D_PATH_DW_FWD_DIRECT = 2,
  X3D_PATH_DW_DGRAD_TILED = 3,
  X3D_PATH_DW_DGRAD_TEMPORAL = 4,
  X3D_PATH_DW_DGRAD_DIRECT = 5,
  X3D_PATH_DW_WGRAD_TILED = 6,
  X3D_PATH_DW_WGRAD_TEMPORAL = 7,
  X3D_PATH_DW_WGRAD_DIRECT = 8,
  X3D_PATH_PW_FWD_TC = 9,
  X3D_PATH_PW_FWD_SIMT = 10,
  X3D_PATH_PW_DGRAD_TC = 11,
  X3D_PATH_PW_DGRAD_SIMT = 12,
  X3D_PATH_PW_WGRAD_TC = 13,
  X3D_PATH_PW_WGRAD_SIMT = 14,
  X3D_PATH_COUNT = 15
} x3d_path_t;
int64_t x3d_path_count(int path);
const char* x3d_path_name(int path);

/* ---- layout converters (user-facing NCDHW fp32 <-> internal NDHWC) -------------------- */
int x3d_ncdhw_to_ndhwc(const float* src, void* dst, int64_t N, int64_t C, int64_t Cp, int64_t T,
                       int64_t H, int64_t W, x3d_dtype_t dt, x3d_stream_t stream);
int x3d_ndhwc_to_ncdhw(const void* src, float* dst, int64_t N, int64_t C, int64_t Cp, int64_t T,
                       int64_t H, int64_t W, x3d_dtype_t dt, x3d_stream_t stream);

/* ---- parameter repack: fp32 master [rows][cols] -> padded operand buffers ------------- */
typedef struct {
  const float* src; /* row-major [rows][cols] */
  void* dst;        /* !transpose: [dst_rows][dst_cols]; transpose: dst[c][r] = src[r][c] */
  int32_t rows, cols;
  int32_t dst_rows, dst_cols; /* padded extents of dst (zero filled) */
  int32_t transpose;
  int32_t dtype; /* x3d_dtype_t of dst */
} x3d_pack_desc_t;
int x3d_pack_params(const x3d_pack_desc_t* descs_dev, int n_desc, int64_t max_dst_elems,
                    x3d_stream_t stream);

/* ---- stem conv1_s: dense 1x3x3, stride (1,2,2), pad (0,1,1)  (x3d.py:196-201,317) ------ */
int x3d_stem_conv_s_fwd(const float* x_ncdhw, const float* w /*[Co][Ci][1][3][3]*/, void* y,
                        int64_t N, int64_t Ci, int64_t T, int64_t H, int64_t W, int64_t Co,
                        int64_t Cop, x3d_dtype_t dt, x3d_stream_t stream);
int x3d_stem_conv_s_wgrad(const float* x_ncdhw, const void* dy, float* dw, int64_t N, int64_t Ci,
                          int64_t T, int64_t H, int64_t W, int64_t Co, int64_t Cop, x3d_dtype_t dt,
                          x3d_stream_t stream);

/* ---- GPU-side input pipeline ---------------------------------------------
 * Source: decoded uint8 frames [B][T][Hs][Ws][3] (NTHWC, RGB).  Per clip a crop window of S x S pixels at (x1, y1) and
 * an optional horizontal flip -- the integer-window part of MultiScaleRandomCropMultigrid and RandomHorizontalFlip
 * (transforms/spatial_transforms.py:472-501, 331-349) -- then ToTensor(norm_value) and Normalize(mean, std)
 * (:35-119; train_x3d_kinetics_multigrid.py:70-73), computed as ((v / norm_value) - mean[c]) / std[c] with the same
 * fp32 roundings as the reference, i.e. bit-identical values.  mean_std = {mean R,G,B, std R,G,B} (HOST pointer).
 * x3d_clip_u8_to_f32 materialises the fp32 NCDHW clip the reference's DataLoader would deliver (drop-in input of
 * ResNet.forward); the *_u8 stem entry points read the frames directly, so no fp32 clip exists at all. */
typedef struct {
  int32_t x1, y1; /* top-left corner of the crop window in the source frame */
  int32_t flip;   /* != 0: horizontal flip */
  int32_t reserved;
} x3d_crop_t;
int x3d_clip_u8_to_f32(const uint8_t* src, const x3d_crop_t* crops_dev, float* dst, int64_t N, int64_t T,
                       int64_t Hs, int64_t Ws, int64_t S, const float* mean_std, float norm_value,
                       x3d_stream_t stream);
int x3d_stem_conv_s_fwd_u8(const uint8_t* src, const x3d_crop_t* crops_dev, const float* w, void* y, int64_t N,
                           int64_t T, int64_t Hs, int64_t Ws, int64_t S, const float* mean_std, float norm_value,
                           int64_t Co, int64_t Cop, x3d_dtype_t dt, x3d_stream_t stream);
int x3d_stem_conv_s_wgrad_u8(const uint8_t* src, const x3d_crop_t* crops_dev, const void* dy, float* dw, int64_t N,
                             int64_t T, int64_t Hs, int64_t Ws, int64_t S, const float* mean_std, float norm_value,
                             int64_t Co, int64_t Cop, x3d_dtype_t dt, x3d_stream_t stream);

/* ---- depthwise conv: conv3x3x3 (x3d.py:87-95,150) and conv1_t 5x1x1 (x3d.py:202-208,318) -
 * kernel (kt,kh,kw) in {(3,3,3),(5,1,1)}, pad k/2, stride (1,s,s).  w_packed: fp32 [taps][Cp].
 * Optional fused input transform relu(x*scale[n%splits][c]+shift[..]) (the preceding
 * SubBatchNorm3d + ReLU, x3d.py:147-148), zero padding applied AFTER the transform.
 * Optional epilogue: per-sample sum / sum-of-squares of the output (for bn2 and the SE pool). */
int x3d_dwconv_fwd(const void* x, const float* w_packed, void* y, int64_t N, int64_t T, int64_t H,
                   int64_t W, int64_t Cp, int kt, int kh, int kw, int stride, const float* in_scale,
                   const float* in_shift, int splits, int relu_in, double* stats, x3d_dtype_t dt,
                   x3d_stream_t stream);
/* dx = dgrad(dy); optional epilogue dx *= (mask_src*mask_scale+mask_shift > 0) plus per-sample
 * sums of dx and dx*mask_src (BN backward of the preceding SubBatchNorm3d). N,T,H,W: input dims. */
int x3d_dwconv_dgrad(const void* dy, const float* w_packed, void* dx, int64_t N, int64_t T,
                     int64_t H, int64_t W, int64_t Cp, int kt, int kh, int kw, int stride,
                     const void* mask_src, const float* mask_scale, const float* mask_shift,
                     int splits, double* stats, x3d_dtype_t dt, x3d_stream_t stream);
/* dw[C][taps] += sum dy * transform(x) */
int x3d_dwconv_wgrad(const void* x, const void* dy, float* dw, int64_t N, int64_t T, int64_t H,
                     int64_t W, int64_t C, int64_t Cp, int kt, int kh, int kw, int stride,
                     const float* in_scale, const float* in_shift, int splits, int relu_in,
                     x3d_dtype_t dt, x3d_stream_t stream);

/* ---- pointwise conv: conv1x1x1 (x3d.py:98-103; :146,162,272,327) ------------------------
 * y[m][n] = sum_k x[row(m)][k] * w[n][k];  rows gathered with stride (1,s,s).
 * w: [Np][Kp] in the activation dtype (from x3d_pack_params).  N,T,H,W: INPUT dims. */
int x3d_pwconv_fwd(const void* x, const void* w, void* y, int64_t N, int64_t T, int64_t H, int64_t W,
                   int64_t Kp, int64_t Np, int stride, double* stats, x3d_dtype_t dt,
                   x3d_stream_t stream);
/* dx[row(m)][k] (+)= sum_n dy[m][n] * wT[k][n];  wT: [Kp][Np].  stride 2: only the sampled
 * rows are touched (accumulate=0 then requires dx pre-zeroed by the caller). */
int x3d_pwconv_dgrad(const void* dy, const void* wT, void* dx, int64_t N, int64_t T, int64_t H,
                     int64_t W, int64_t Kp, int64_t Np, int stride, int accumulate, x3d_dtype_t dt,
                     x3d_stream_t stream);
/* dw[Nn][K] += sum_m dy[m][n] * x[row(m)][k]   (fp32, PyTorch [out][in] layout) */
int x3d_pwconv_wgrad(const void* x, const void* dy, float* dw, int64_t N, int64_t T, int64_t H,
                     int64_t W, int64_t K, int64_t Kp, int64_t Nn, int64_t Np, int stride,
                     x3d_dtype_t dt, x3d_stream_t stream);

/* same with a caller-provided scratch buffer (device memory, 16-byte aligned, not shared with a concurrently running
 * call): the tensor-core kernel then writes one partial dW tile per M split into it and a second small kernel adds
 * the partials in a fixed order -- deterministic, and free of the per-element fp32 reds of the plain entry point.
 * Any size >= 1 MB works (the M split adapts); x3d_pwconv_wgrad_workspace_bytes() is the size that never limits it. */
int x3d_pwconv_wgrad_ws(const void* x, const void* dy, float* dw, int64_t N, int64_t T, int64_t H, int64_t W,
                        int64_t K, int64_t Kp, int64_t Nn, int64_t Np, int stride, void* workspace,
                        size_t workspace_bytes, x3d_dtype_t dt, x3d_stream_t stream);
size_t x3d_pwconv_wgrad_workspace_bytes(void);

/* ---- SubBatchNorm3d (x3d.py:9-58) -------------------------------------------------------- */
/* train: per-(split,channel) mean/var from per-sample sums -> scale/shift (+saved mean, rstd),
 * running-stat update of split_bn (momentum, unbiased var) and num_batches_tracked += 1. */
int x3d_bn_finalize(const double* stats, int64_t N, int splits, int64_t P, int64_t C, int64_t Cp,
                    const float* gamma, const float* beta, float* run_mean, float* run_var,
                    int64_t* num_batches_tracked, float momentum, float eps, float* scale,
                    float* shift, float* mean, float* rstd, x3d_stream_t stream);
/* eval: scale/shift from bn.running_* (x3d.py:54) ; mean/rstd outputs optional */
int x3d_bn_eval_params(const float* gamma, const float* beta, const float* run_mean,
                       const float* run_var, int64_t C, int64_t Cp, float eps, float* scale,
                       float* shift, float* mean, float* rstd, x3d_stream_t stream);
/* out = [relu]( a*scale+shift  [+ res | + res*res_scale+res_shift] )   (x3d.py:163-169, 319-320) */
int x3d_bn_act_fwd(const void* a, const float* scale, const float* shift, int splits,
                   const void* res, const float* res_scale, const float* res_shift, int relu,
                   void* out, int64_t N, int64_t P, int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream);
/* BN backward, pass 1: dpre = dout * (mask_out > 0 if mask_out) ; stats += {sum dpre, sum dpre*a} */
int x3d_bn_bwd_reduce(const void* dout, const void* mask_out, const void* a, double* stats,
                      int64_t N, int64_t P, int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream);
/* same, and dpre is also written out: the identity-residual block keeps it as the start value of its input
 * gradient (dx = dpre + conv1 dgrad, x3d.py:168-169), so neither the apply pass below nor a separate
 * "dx += dout*(out>0)" pass reads dout / the mask source again */
int x3d_bn_bwd_reduce_store(const void* dout, const void* mask_out, const void* a, double* stats,
                            void* dpre_out, int64_t N, int64_t P, int64_t Cp, x3d_dtype_t dt,
                            x3d_stream_t stream);
/* pass 2 (tiny): coefficients of da = A*dpre + B*a + C per (split,channel); dgamma/dbeta += .
 * train=0: BN used running stats (eval) -> A=scale, B=C=0. */
int x3d_bn_bwd_finalize(const double* stats, int64_t N, int splits, int64_t P, int64_t C, int64_t Cp,
                        const float* gamma, const float* mean, const float* rstd, int train,
                        float* coef /*[3][splits][Cp]*/, float* dgamma, float* dbeta,
                        x3d_stream_t stream);
/* pass 3: da = A*dpre + B*a + C */
int x3d_bn_bwd_apply(const void* dout, const void* mask_out, const void* a, const float* coef,
                     int splits, void* da, int64_t N, int64_t P, int64_t Cp, x3d_dtype_t dt,
                     x3d_stream_t stream);
/* dx += dout * (out > 0)   (identity-residual branch of x3d.py:168-169) */
int x3d_relu_bwd_add(const void* dout, const void* out, void* dx, int64_t numel, x3d_dtype_t dt,
                     x3d_stream_t stream);

/* ---- squeeze-excitation + Swish (x3d.py:61-84,120-125,153-160) --------------------------- */
/* pooled[n][c] = scale*mean_thw(a2)+shift from the dw-conv's per-sample sums; h = relu(W1 p + b1);
 * gate = sigmoid(W2 h + b2).  W1: [w][C] fp32 master, W2: [C][w]. gate is [N][Cp] fp32. */
int x3d_se_fwd(const double* stats, const float* scale, const float* shift, int splits, int64_t N,
               int64_t P, int64_t C, int64_t Cp, int sw, const float* W1, const float* b1,
               const float* W2, const float* b2, float* pooled, float* hidden, float* gate,
               x3d_stream_t stream);
/* v = swish(gate[n][c] * (a2*scale+shift))   (gate may be NULL: blocks without SE) */
int x3d_swish_gate_fwd(const void* a2, const float* scale, const float* shift, int splits,
                       const float* gate, void* v, int64_t N, int64_t P, int64_t Cp, x3d_dtype_t dt,
                       x3d_stream_t stream);
/* dz = dv * swish'(gate*(a2*scale+shift));  stats += {sum dz, sum dz*a2} per sample */
int x3d_swish_gate_bwd_reduce(const void* dv, const void* a2, const float* scale, const float* shift,
                              int splits, const float* gate, double* stats, int64_t N, int64_t P,
                              int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream);
/* tiny: SE backward + BN2 backward coefficients.
 * fwd_stats: the dw-conv's per-sample {sum a2, sum a2^2}; bwd_stats: from the reduce above.
 * Produces the planar coef[3][N][Cp] = (E1,E2,E3) with da2 = E1*dz + E2*a2 + E3, accumulates (in sample order, no atomics) the SE
 * parameter gradients (NULL pointers: block has no SE) and bn2's dgamma/dbeta. */
int x3d_se_bn_bwd(const double* fwd_stats, const double* bwd_stats, int64_t N, int splits, int64_t P,
                  int64_t C, int64_t Cp, int sw, const float* gamma, const float* mean,
                  const float* rstd, const float* scale, const float* shift, int train,
                  const float* W1, const float* W2, const float* pooled, const float* hidden,
                  const float* gate, float* dW1, float* db1, float* dW2, float* db2, float* dgamma,
                  float* dbeta, float* work /*fp32 [N][2*Cp + sw]*/, float* coef /*[3][N][Cp]*/,
                  x3d_stream_t stream);
int x3d_swish_gate_bwd_apply(const void* dv, const void* a2, const float* scale, const float* shift,
                             int splits, const float* gate, const float* coef, void* da2, int64_t N,
                             int64_t P, int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream);

/* ---- head (x3d.py:327-345) --------------------------------------------------------------- */
/* pooled[r][c] = mean over the pooled positions of relu(a5*scale+shift); rows r = n (pool_t=1,
 * task 'class') or n*T+t (pool_t=0, task 'loc').  pooled is fp32 [R][C]. */
int x3d_bn_relu_pool_fwd(const void* a5, const float* scale, const float* shift, int splits,
                         float* pooled, int64_t N, int64_t T, int64_t HW, int pool_t, int64_t C,
                         int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream);
/* backward of the above through ReLU into BN: dpre = dpooled[r][c]/count * (a5*scale+shift > 0) */
int x3d_bn_relu_pool_bwd_reduce(const void* a5, const float* scale, const float* shift, int splits,
                                const float* dpooled, double* stats, int64_t N, int64_t T, int64_t HW,
                                int pool_t, int64_t C, int64_t Cp, x3d_dtype_t dt, x3d_stream_t stream);
int x3d_bn_relu_pool_bwd_apply(const void* a5, const float* scale, const float* shift, int splits,
                               const float* dpooled, const float* coef, void* da5, int64_t N,
                               int64_t T, int64_t HW, int pool_t, int64_t C, int64_t Cp,
                               x3d_dtype_t dt, x3d_stream_t stream);
/* small dense fp32 GEMM for fc1 / fc2 (x3d.py:242-243,333-343) and their gradients:
 * C[i][j] (+)= sum_k A(i,k) * B(k,j) [+ bias[j]] ; A(i,k)=A[i*sai+k*sak], B(k,j)=B[k*sbk+j*sbj];
 * epilogue: relu, then optional elementwise multiply by mul[i][j] (dropout mask). */
int x3d_small_gemm(const float* A, int64_t sai, int64_t sak, const float* B, int64_t sbk, int64_t sbj,
                   float* C, int64_t ldc, int64_t M, int64_t Nn, int64_t K, const float* bias, int relu,
                   const float* mul, int accumulate, x3d_stream_t stream);
/* Same GEMM with K split over CTAs (the head has 7 - 32 column tiles and K up to 2048: without the split a handful of CTAs
 * walk the whole K).  Slice partials go through `ws`; the last CTA of a tile adds them in slice order (deterministic).
 * ws: 16-byte aligned, >= x3d_small_gemm_workspace_bytes(M, Nn, K) for the full split (less = fewer slices), its first
 * 4 KB (tile tickets) ZERO before the first call -- every call leaves them zero again.  One workspace per stream. */
int x3d_small_gemm_ws(const float* A, int64_t sai, int64_t sak, const float* B, int64_t sbk, int64_t sbj,
                      float* C, int64_t ldc, int64_t M, int64_t Nn, int64_t K, const float* bias, int relu,
                      const float* mul, int accumulate, void* ws, int64_t ws_bytes, x3d_stream_t stream);
int64_t x3d_small_gemm_workspace_bytes(int64_t M, int64_t Nn, int64_t K);
/* dst[j] += sum_i src[i][j] (bias gradients) */
int x3d_colsum(const float* src, int64_t M, int64_t Nn, float* dst, x3d_stream_t stream);
/* dst = src * (ref > 0) * mul  (ReLU + dropout backward on the [R][2048] head activations) */
int x3d_relu_mask_mul(const float* src, const float* ref, const float* mul, float* dst, int64_t numel,
                      x3d_stream_t stream);

/* ---- fused SGD over a flat parameter table (train_x3d_kinetics_multigrid.py:183,277) ------ */
typedef struct {
  float* param;
  const float* grad;
  float* momentum_buf;
  int64_t numel;
} x3d_sgd_desc_t;
/* one launch over all tensors (n_desc <= 2048): the work is cut into 4096-element blocks across the table, 16-byte
 * accesses where the three pointers of a tensor are 16-byte aligned.  max_numel: largest numel of the table. */
int x3d_sgd_step(const x3d_sgd_desc_t* descs_dev, int n_desc, int64_t max_numel, float lr,
                 float momentum, float weight_decay, float grad_scale, int first_step,
                 x3d_stream_t stream);
/* same, hyper-parameters read from device memory hyper_dev = {lr, momentum, weight_decay, grad_scale}: lets a
 * captured CUDA graph follow LR schedules (the long-cycle LR law, train_x3d_kinetics_multigrid.py:229-233) */
int x3d_sgd_step_dev(const x3d_sgd_desc_t* descs_dev, int n_desc, int64_t max_numel, const float* hyper_dev,
                     int first_step, x3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* X3D_B200_H */
