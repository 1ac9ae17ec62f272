/*
 * ogv.h -- C-ABI of libogvit.so: hand-written sm_100a kernels for the OutGridBlock hot path
 * (Outlooker -> MBConv -> GridAttn -> MLP, forward and backward).
 *
 * Every entry point replaces a stretch of PyTorch-eager code in the reference
 * (pablo-reyes8/outlook-grid-vision-transformer); the reference file:line each one stands for
 * is cited beside it.  The reference has no FFI of its own (it is pure Python), so this header
 * IS the boundary a maintainer would bind (ctypes stub in INTEGRATION.md).
 *
 * Conventions
 *  - all pointers are DEVICE pointers owned by the caller; no hidden allocation, no host sync;
 *  - activations are "rows x channels" tensors: row m = (b*H + h)*W + w (NHWC flattened),
 *    channels contiguous, dtype = OGV_F32 or OGV_BF16; statistics / parameters / gradients of
 *    parameters are always fp32;
 *  - `stream` is a cudaStream_t passed as void*;
 *  - return value: 0 on success, negative error code otherwise (ogv_last_error() has the text).
 */
#ifndef OGV_H_
#define OGV_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OGV_OK 0
#define OGV_ERR_ARG (-1)
#define OGV_ERR_CUDA (-2)
#define OGV_ERR_UNSUPPORTED (-3)

#define OGV_F32 0
#define OGV_BF16 1

#define OGV_ACT_NONE 0
#define OGV_ACT_GELU 1
#define OGV_ACT_SILU 2
#define OGV_ACT_SIGMOID 3
#define OGV_ACT_RELU 4
#define OGV_ACT_MUL 5 /* `dact` only: multiply by dact_src itself (a derivative saved by the forward pass) */

#define OGV_ENGINE_AUTO 0
#define OGV_ENGINE_SIMT 1 /* fp32 FFMA tiles (exact-fp32 parity mode, debugging) */
#define OGV_ENGINE_TC 2   /* tcgen05 + TMEM + TMA (bf16 operands, fp32 accumulate) */

int ogv_version(void);
const char* ogv_last_error(void);
int ogv_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Pointwise GEMM with fused epilogue.   D[m,n] = epi( sum_k A(m,k) * B(n,k) )
 * Stands for every nn.Conv2d 1x1 / nn.Linear on the path, forward, dgrad and wgrad:
 *   outlook_attention.py:41-47,83-88,100,110,122 ; mbc_conv.py:17-19,67-85 ;
 *   grid_attention.py:57-59,70,86-87 ; Out_Grid_Block.py:18-22,27-31 (and their autograd).
 * A(m,k) = A[m*a_rs + k*a_cs], B(n,k) = B[n*b_rs + k*b_cs]  (element strides).
 *   forward : A = activations [M,K] (a_cs=1), B = weight [N,K] (b_cs=1)
 *   dgrad   : A = dY [M,N'],  B = W^T [K',N'] (pre-transposed copy, b_cs=1)
 *   wgrad   : A(m,k) = dY[k, m] (a_rs=1), B(n,k) = X[k, n] (b_rs=1), reduction over rows
 * epilogue order: v = acc (+bias[n]); pre_out[m,n] = v (or act'(v) if pre_out_grad); v = act(v); v *= act'(dact_src[m,n]);
 *                 v *= row_scale[m / rows_per_scale]; v += residual[m,n]; D[m,n] = v (or += v).
 * ------------------------------------------------------------------------------------------- */
typedef struct ogv_gemm_args {
  const void* A;
  long long a_rs, a_cs;
  const void* B;
  long long b_rs, b_cs;
  void* D;
  long long ldd;
  int M, N, K;
  int in_dtype;  /* dtype of A and B */
  int out_dtype; /* dtype of D, pre_out, residual, dact_src */
  const float* bias; /* [N] or NULL */
  void* pre_out;     /* optional [M,N] (ld_pre) */
  long long ld_pre;
  int act; /* OGV_ACT_* applied after bias */
  const void* dact_src; /* optional [M,N]: multiply by act'(dact_src), act code in `dact` */
  long long ld_dact;
  int dact;
  const float* row_scale; /* optional per-sample scale (stochastic depth) */
  int rows_per_scale;
  const void* residual; /* optional [M,N] */
  long long ld_res;
  int accumulate; /* 1: D is fp32 and is atomically accumulated into (split-K partial sums) */
  int split_k;    /* >=1; >1 requires accumulate */
  float* col_sum;   /* optional [N]: += sum_m of stored value (BatchNorm batch statistics) */
  float* col_sumsq; /* optional [N]: += sum_m of stored value^2 */
  int pre_out_grad; /* 1: pre_out receives act'(v) instead of v, so that the backward pass multiplies by it
                       (dact = OGV_ACT_MUL) without re-evaluating a transcendental */
  float* row_sum;   /* optional [M], accumulate mode only: += sum_k A(m,k).  With the wgrad operands (A(m,k) = dY[k,m])
                       this is the BIAS gradient colsum(dY), produced by one extra N=16 MMA per K step against a tile
                       of ones instead of a separate pass over dY */
} ogv_gemm_args;

int ogv_gemm(const ogv_gemm_args* args, int engine, void* stream);

/* Layout boundary: the reference hands the block NCHW tensors (Out_Grid_Block.py:88-107). */
int ogv_nchw_to_nhwc(const void* src, void* dst, int B, int C, int HW, int dtype, void* stream);
int ogv_nhwc_to_nchw(const void* src, void* dst, int B, int C, int HW, int dtype, void* stream);

/* fp32 master weight [rows, cols] -> compute-dtype copy dst[r*ld_dst + c] and/or transposed copy
 * dst_t[c*ld_dst_t + r] (what torch.autocast's weight cast does per step; autocast.py:60-66). */
int ogv_cast_transpose(const float* src, void* dst, long long ld_dst, void* dst_t, long long ld_dst_t,
                       int rows, int cols, int dtype, void* stream);
/* The same cast for MANY weights in one launch (the per-step refresh of every compute-dtype weight copy of a model:
 * ~120 tiny launches otherwise).  `items` is a DEVICE array of n_items descriptors, caller-owned; item i covers the
 * 32x32 tiles [tile0_i, tile0_{i+1}) of the batch (tile0 = exclusive prefix sum of ceil(rows/32)*ceil(cols/32)),
 * total_tiles = the sum.  dst_dtype is per item (OGV_F32 items are plain copies, e.g. concatenated biases). */
typedef struct ogv_cast_item {
  const float* src;   /* [rows, cols] fp32, contiguous */
  void* dst;          /* dst[r*ld_dst + c], or NULL */
  void* dst_t;        /* dst_t[c*ld_dst_t + r], or NULL */
  long long ld_dst, ld_dst_t;
  int rows, cols;
  int tile0;
  int dst_dtype;
} ogv_cast_item;
int ogv_cast_batch(const ogv_cast_item* items, int n_items, int total_tiles, void* stream);
/* fp32 [rows, cols] -> three bf16 planes (hi, hi, lo) [pattern 0] or (hi, lo, hi) [pattern 1], `plane_stride` dst
 * elements apart: the operands of the "bf16 x 3" tensor-core emulation of an fp32 GEMM (a*b ~= a_hi*b_hi +
 * a_hi*b_lo + a_lo*b_hi over a 3x longer reduction axis; relative error ~2^-17).  Used for the fp32 parity mode. */
int ogv_split3(const float* src, long long ld_src, void* dst, long long ld_dst, long long plane_stride, long long rows,
               int cols, int pattern, void* stream);
/* y = x * scale[row / rows_per_scale]  (DropPath, Outlook_Block.py:15-22) */
int ogv_rowscale(const void* x, const float* scale, void* y, long long rows, int cols, int rows_per_scale,
                 int dtype, void* stream);
/* y = x * scale[row / rows_per_scale] and out[n] += sum_m y[m,n] in one pass (DropPath backward + bias gradient) */
int ogv_rowscale_colsum(const void* x, const float* scale, void* y, float* out, long long rows, int cols,
                        int rows_per_scale, int dtype, void* stream);
/* out[n] += sum_m x[m,n]  (bias gradients) */
int ogv_colsum(const void* x, long long ld, float* out, long long M, int N, int dtype, void* stream);
/* out = a * act'(pre) elementwise, n elements (SE gate backward) */
int ogv_mul_dact(const void* a, const void* pre, void* out, long long n, int act, int dtype, void* stream);
/* y = a + b elementwise */
int ogv_add(const void* a, const void* b, void* y, long long n, int dtype, void* stream);

/* LayerNorm over channels (outlook_attention.py:24-31 eps 1e-6; Out_Grid_Block.py:69,84 eps 1e-5). */
int ogv_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      long long M, int C, float eps, int dtype, void* stream);
/* dx = dres + LN'(dy); dgamma/dbeta are accumulated (+=) */
int ogv_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                      const void* dres, void* dx, float* dgamma, float* dbeta, long long M, int C, int dtype,
                      void* stream);

/* Outlook core (outlook_attention.py:104-120): `va` holds v in columns [0,C) and the k*k logits of
 * head h in columns [C + h*9, C + h*9 + 9); row stride ld_va.  y[p,c] = sum_t softmax(logits)[t] * v[p+d_t,c]. */
int ogv_outlook_core_fwd(const void* va, long long ld_va, void* y, int B, int H, int W, int C, int heads,
                         int dtype, void* stream);
/* dva gets dv in [0,C), dlogits in [C, C+9*heads), zeros up to ld_va. */
int ogv_outlook_core_bwd(const void* va, long long ld_va, const void* dy, void* dva, int B, int H, int W, int C,
                         int heads, int dtype, void* stream);

/* BatchNorm2d (mbc_conv.py:61; training = batch statistics).  stats: sum/sumsq -> scale/shift,
 * mean/rstd, and the running-stat update (momentum, unbiased variance). */
int ogv_colstats(const void* x, long long ld, float* sum, float* sumsq, long long M, int N, int dtype, void* stream);
int ogv_bn_finalize(const float* sum, const float* sumsq, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* rstd,
                    long long n, int C, float eps, float momentum, int training, void* stream);
/* out = res + scale*x + shift  (res optional) */
int ogv_bn_apply(const void* x, const float* scale, const float* shift, const void* res, void* out, long long M,
                 int C, int dtype, void* stream);
/* dbeta += sum dy ; dgamma += sum dy * (x-mean)*rstd */
int ogv_bn_bwd_reduce(const void* dy, const void* x, const float* mean, const float* rstd, float* dgamma,
                      float* dbeta, long long M, int C, int dtype, void* stream);
/* dx = gamma*rstd*(dy - dbeta/n - xhat*dgamma/n) */
int ogv_bn_bwd_apply(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                     const float* dgamma, const float* dbeta, void* dx, long long M, int C, int dtype,
                     void* stream);

/* Squeeze-excite MLP (mbc_conv.py:17-19,24-26) on [B, Cm] rows, both 1x1 convs in one kernel per direction (bf16
 * compute mode; ogv_se_mlp_supported tells which shapes).  w1t = W1^T [Cm, Cs], w2t = W2^T [Cs, Cm] (forward),
 * w2 = W2 [Cm, Cs], w1 = W1 [Cs, Cm] (backward): the K-major bf16 copies.  pool_c / s1_pre / s1a / dgate_c / ds1_pre are
 * bf16 (operands of the weight-gradient GEMMs), gate_pre / gate / dpool fp32.
 *   forward : s1_pre = pool W1^T + b1, s1a = act(s1_pre), gate_pre = s1a W2^T + b2, gate = sigmoid(gate_pre)
 *   backward: dgate_c = dgate * sigmoid'(gate_pre), ds1_pre = (dgate_c W2) * act'(s1_pre), dpool = ds1_pre W1 */
int ogv_se_mlp_supported(int Cm, int Cs, int dtype);
int ogv_se_mlp_fwd(const float* pool, const void* w1t, const float* b1, const void* w2t, const float* b2, void* pool_c,
                   void* s1_pre, void* s1a, float* gate_pre, float* gate, int B, int Cm, int Cs, int act, void* stream);
int ogv_se_mlp_bwd(const float* dgate, const float* gate_pre, const void* s1_pre, const void* w2, const void* w1,
                   void* dgate_c, void* ds1_pre, float* dpool, int B, int Cm, int Cs, int act, void* stream);

/* 3x3 stride-1 pad-1 patches of a channels_last image x[B,H,W,Cin] as GEMM rows cols[B*H*W, Kpad], column
 * (ky*3+kx)*Cin + ci, zeros outside the image and in columns >= 9*Cin: the stem convolution (stem_head.py:23-32) becomes
 * ogv_gemm forward and weight gradient. */
int ogv_im2col3x3(const void* x, void* cols, int B, int H, int W, int Cin, int Kpad, int dtype, void* stream);

/* The Downsample convolution (downsampling.py:41-47: nn.Conv2d(C, 2C, 3, stride 2, padding 1)) and any other 3x3 / pad 1 /
 * stride 1|2 convolution over a channels_last image with Cin % 8 == 0, as GEMM rows:
 *   ogv_im2col3x3_vec: cols[(b, oy, ox), (ky*3+kx)*Cin + c] = x[b, oy*s-1+ky, ox*s-1+kx, c], zero outside the image;
 *                      Ho = (H-1)/s + 1, Wo = (W-1)/s + 1; forward = ogv_gemm(cols, W2), weight gradient = ogv_gemm
 *                      (dY^T cols) with W2[co, (ky*3+kx)*Cin + c] = weight[co, c, ky, kx];
 *   ogv_col2im3x3_vec: the input gradient from dcols = dY x W2 (ogv_gemm): dx[b, iy, ix, c] = sum of the patch entries
 *                      that were copies of that pixel (replaces autograd's convolution_backward input branch). */
int ogv_im2col3x3_vec(const void* x, void* cols, int B, int H, int W, int Cin, int stride, int dtype, void* stream);
int ogv_col2im3x3_vec(const void* dcols, void* dx, int B, int H, int W, int Cin, int stride, int dtype, void* stream);

/* The same convolution as an IMPLICIT GEMM on the tcgen05 engine (bf16, Cin % 64 == 0, output rows that tile 64 / 128
 * pixels): the patch matrix is never written -- the GEMM's TMA producer fetches each (tap, 64-channel) slice of a tile of
 * output pixels as one 5-D box of x viewed as {2*Cin, W/2, 2, H/2, B} (stride 2) or {Cin, W, 1, H, B} (stride 1), the zero
 * padding being the descriptor's out-of-bounds fill.
 *   ogv_conv3x3_fwd  : y[(b,oy,ox), co] = sum_k cols[.., k] * w2[co, k]          (downsampling.py:41-47 forward)
 *                      col_sum / col_sumsq (fp32 [Co], nullable): += per-channel sum / sum of squares of the stored
 *                      outputs -- the BatchNorm batch statistics of the unit, from the GEMM's epilogue
 *   ogv_conv3x3_wgrad: dw2[co, k] += sum_pixels dy[pixel, co] * cols[pixel, k]   (fp32, split over pixels, pre-zeroed)
 * with w2[co, (ky*3+kx)*Cin + c] = weight[co, c, ky, kx].  ogv_conv3x3_supported() -> 1 when the geometry is served. */
int ogv_conv3x3_supported(int B, int H, int W, int Cin, int Co, int stride);
int ogv_conv3x3_fwd(const void* x, const void* w2, void* y, float* col_sum, float* col_sumsq, int B, int H, int W, int Cin,
                    int Co, int stride, void* stream);
int ogv_conv3x3_wgrad(const void* x, const void* dy, float* dw2, int B, int H, int W, int Cin, int Co, int stride,
                      void* stream);

/* conv -> BatchNorm -> act units around the blocks (stem_head.py:23-32, downsampling.py:28-65), the BN + act part:
 * out = act(scale*x + shift); backward with g = dy * act'(scale*x + shift) recomputed in both passes:
 * dbeta += sum g, dgamma += sum g*xhat; dx = gamma*rstd*(g - dbeta/n - xhat*dgamma/n). */
int ogv_bn_act_apply(const void* x, const float* scale, const float* shift, void* out, long long M, int C, int act,
                     int dtype, void* stream);
int ogv_bn_act_bwd_reduce(const void* dy, const void* x, const float* scale, const float* shift, const float* mean,
                          const float* rstd, float* dgamma, float* dbeta, long long M, int C, int act, int dtype,
                          void* stream);
int ogv_bn_act_bwd_apply(const void* dy, const void* x, const float* scale, const float* shift, const float* mean,
                         const float* rstd, const float* gamma, const float* dgamma, const float* dbeta, void* dx,
                         long long M, int C, int act, int dtype, void* stream);

/* Depthwise 3x3 (mbc_conv.py:73-78) with BN1-affine + SiLU applied to the input on load and the
 * batch statistics of the output accumulated on store.  w is [Cm, 9] fp32. */
int ogv_dwconv_fwd(const void* e_pre, const float* scale1, const float* shift1, const float* w, void* d_pre,
                   float* sum2, float* sumsq2, int B, int H, int W, int Cm, int act, int dtype, void* stream);
/* SE squeeze (mbc_conv.py:22-23): pool[b,c] = mean_hw act(scale2*d_pre+shift2) */
int ogv_se_pool(const void* d_pre, const float* scale2, const float* shift2, float* pool, int B, int HW, int Cm,
                int act, int dtype, void* stream);
/* d_act = act(scale2*d_pre+shift2) * gate[b,c]   (mbc_conv.py:27) */
int ogv_bn_act_gate(const void* d_pre, const float* scale2, const float* shift2, const float* gate, void* d_act,
                    int B, int HW, int Cm, int act, int dtype, void* stream);
/* dgate[b,c] = sum_hw dd_act * act(scale2*d_pre+shift2) */
int ogv_se_bwd_reduce(const void* dd_act, const void* d_pre, const float* scale2, const float* shift2,
                      float* dgate, int B, int HW, int Cm, int act, int dtype, void* stream);
/* ONE pass over (dd_act, d_pre) giving every per-(image, channel) sum the SE backward and the BN2
 * backward reductions need (mbc_conv.py:22-27,61 backward).  With u = scale2*d_pre+shift2, a = act(u),
 * a' = act'(u), xh = (d_pre-mean2)*rstd2, g = dd_act:   stats[q][b][c], q = 0..4 =
 *   sum_hw g*a | sum_hw g*a' | sum_hw g*a'*xh | sum_hw a' | sum_hw a'*xh      (stats: [5,B,Cm] fp32)
 * stats[0] IS dgate (same values as ogv_se_bwd_reduce). */
int ogv_mbconv_bwd_stats(const void* dd_act, const void* d_pre, const float* scale2, const float* shift2,
                         const float* mean2, const float* rstd2, float* stats, int B, int HW, int Cm, int act,
                         int dtype, void* stream);
/* du = (dd_act*gate + dpool/HW)*a' :  dbeta2[c] += sum_b gate*stats1 + dpool/HW*stats3 ;
 *                                      dgamma2[c] += sum_b gate*stats2 + dpool/HW*stats4 */
int ogv_mbconv_bn2_finalize(const float* stats, const float* gate, const float* dpool, float* dgamma2,
                            float* dbeta2, int B, int HW, int Cm, void* stream);
/* dd_pre = gamma2*rstd2*(du - dbeta2/n - xhat2*dgamma2/n), du as above, n = B*HW */
int ogv_dw_bn2_bwd_apply(const void* dd_act, const void* d_pre, const float* gate, const float* dpool,
                         const float* scale2, const float* shift2, const float* mean2, const float* rstd2,
                         const float* gamma2, const float* dgamma2, const float* dbeta2, void* dd_pre, int B,
                         int HW, int Cm, int act, int dtype, void* stream);
/* Depthwise backward: du1 = corr(dd_pre, w) * act'(u1); dw[c,t] += sum dd_pre[p]*act(u1[p+d_t]);
 * dbeta1 += sum du1 ; dgamma1 += sum du1*xhat1. */
int ogv_dwconv_bwd(const void* dd_pre, const void* e_pre, const float* scale1, const float* shift1,
                   const float* mean1, const float* rstd1, const float* w, void* du1, float* dw, float* dgamma1,
                   float* dbeta1, int B, int H, int W, int Cm, int act, int dtype, void* stream);

/* Grid attention core (grid_partition.py:13-15 + grid_attention.py:70-86) on qkv [M,3C] in
 * NHWC row order; the grid partition/unpartition is index arithmetic inside the kernel. */
int ogv_grid_attn_fwd(const void* qkv, void* out, int B, int H, int W, int C, int heads, int g, int dtype,
                      void* stream);
int ogv_grid_attn_bwd(const void* qkv, const void* dout, void* dqkv, int B, int H, int W, int C, int heads, int g,
                      int dtype, void* stream);
/* attention probabilities for analysis hooks (grid_attention.py:77-78): attn [B*g*g, heads, N, N] fp32 */
int ogv_grid_attn_probs(const void* qkv, float* attn, int B, int H, int W, int C, int heads, int g, int dtype,
                        void* stream);

/* Fused multi-tensor AdamW over a flat fp32 arena (train_full_model.py:56-57). */
int ogv_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
              float eps, float weight_decay, float bias_c1, float bias_c2, float grad_scale, void* stream);


/* ---------------------------------------------------------------------------------------------
 * Sync-free tail of the train step over FLAT fp32 arenas (parameters, gradients, Adam moments laid out
 * back to back, every parameter starting on an 8-float granule) -- what lets one captured CUDA graph follow the
 * reference's per-step host logic (one_epoch_train.py:98-166) without a host sync:
 *   ogv_sumsq         : out[0] += sum g^2  -- the global gradient norm of clip_grad_norm_ (one_epoch_train.py:141).
 *   ogv_adamw_flat    : clip + AdamW (train_full_model.py:56-57) in one pass.  hyper (DEVICE fp32[5]) =
 *                       { lr, 1-beta1^t, 1-beta2^t, grad_scale (1/world), max_norm (<=0: no clip) } is rewritten by the
 *                       host before every replay, so WarmupCosineLR (warmup.py:29-59) works under graph replay;
 *                       decay_bits: bit i = granule i takes weight decay (the two param groups of warmup.py:4-26);
 *                       the update is SKIPPED when *loss or *gnorm_sq is not finite (one_epoch_train.py:99-109) and
 *                       *skipped is incremented; gnorm_sq / loss / skipped may be null.
 *   ogv_train_metrics : acc[5] += { loss*B, top-1, top-3, top-5 hits, B } from fp32 logits [B, K] (row stride ld) and
 *                       int64 labels (metrics.py:7-24, one_epoch_train.py:155-166) -- read back once per epoch. */
int ogv_sumsq(const float* g, long long n, float* out, void* stream);
int ogv_adamw_flat(float* p, const float* g, float* m, float* v, const unsigned* decay_bits, long long n,
                   const float* hyper, const float* gnorm_sq, const float* loss, float beta1, float beta2, float eps,
                   float weight_decay, float* skipped, void* stream);
int ogv_train_metrics(const float* logits, long long ld, const long long* labels, int B, int K, const float* loss,
                      float* acc, void* stream);

/* The training criterion nn.CrossEntropyLoss(label_smoothing=eps), mean reduction (train_full_model.py:52,
 * one_epoch_train.py:95-97) as one kernel per direction:
 *   ogv_xent_fwd: lse[i] = logsumexp(logits[i, :]); *loss += mean_i (1-eps)*(lse_i - x_i[y_i]) + eps*(lse_i - mean_k x_i[k])
 *                 (loss pre-zeroed by the caller; labels int64 in [0, K), no ignore_index)
 *   ogv_xent_bwd: dlogits[i, k] = *gout / B * (exp(x_ik - lse_i) - (1-eps)*[k == y_i] - eps/K)   (gout NULL: 1) */
int ogv_xent_fwd(const void* logits, long long ld, const long long* labels, int B, int K, float smoothing, int dtype,
                 float* lse, float* loss, void* stream);
int ogv_xent_bwd(const void* logits, long long ld, const long long* labels, const float* lse, const float* gout, int B,
                 int K, float smoothing, int dtype, void* dlogits, long long ldd, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused MLP (north_star kernel 4):  y = residual + row_scale[m / rows_per_scale] * ( act(x W1^T + b1) W2^T + b2 )
 * in ONE tcgen05 kernel -- MLP2d (outlook_attention.py:43-49) and MLP (Out_Grid_Block.py:24-32) with the residual /
 * DropPath of Outlook_Block.py:63 and Out_Grid_Block.py:102-104.  The [M, hidden] activation stays on chip: fc1 chunk ->
 * TMEM -> activation in registers -> bf16 tile in swizzled shared memory -> A operand of fc2.  bf16 tensors, fp32
 * accumulation / bias / scale.  x: [M, C] (row stride ldx; the LayerNorm-ed rows), w1: [hidden, C], w2: [C, hidden]
 * (both row-major, dense), residual: [M, C] or null, y: [M, C].  ogv_mlp_fused_supported(C, hidden) says whether the
 * shape is served (C in {64, 128}, hidden a multiple of the chunk width); other shapes take the ogv_gemm route. */
int ogv_mlp_fused_supported(int C, int hidden);
int ogv_mlp_fwd(const void* x, long long ldx, const void* w1, const float* b1, const void* w2, const float* b2,
                const void* residual, long long ldr, const float* row_scale, int rows_per_scale, void* y, long long ldy,
                long long M, int C, int hidden, int act, void* stream);
/* Backward of the same (autograd of the reference lines above) with the hidden activation RECOMPUTED on chip:
 *   z = xn W1^T + b1 ;  dz = (dy W2) * act'(z) * s ;  hs = act(z) * s ;  dxn = dz W1      (s = row_scale[m / rows_per_scale] or 1)
 * dz, hs: [M, hidden] bf16 outputs for the weight-gradient GEMMs (dW1 = dz^T xn, db1 = colsum dz, dW2 = dy^T hs);
 * w2t = W2^T [hidden, C] and w1t = W1^T [C, hidden] are the pre-transposed bf16 copies the dgrad GEMMs use. */
int ogv_mlp_bwd(const void* xn, long long ldx, const void* dy, long long ldg, const void* w1, const void* w2t,
                const void* w1t, const float* b1, const float* row_scale, int rows_per_scale, void* dz, void* hs,
                void* dxn, long long lddx, long long M, int C, int hidden, int act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OGV_H_ */
