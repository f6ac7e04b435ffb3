/*
 * deeplab_b200.h -- C ABI of libdeeplab_b200.so: the sm_100a DeepLabV3+ hot path.
 *
 * This is the drop-in boundary for the reference's model/CRF hot path.  Every entry
 * point replaces the work a Keras layer call (or a pydensecrf call) did in the reference; the reference
 * site each one stands in for is cited as  <file>:<line>  into Golbstein/Keras-segmentation-deeplab-v3.1.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers + explicit sizes, no torch / C++ types.  The caller owns every buffer.
 *   - activations are NHWC ("channels last", the TF data format the reference uses, deeplabv3p.py:216-217),
 *     a 1x1 convolution therefore sees a row-major matrix [M = B*H*W, C].
 *   - `dtype` is the storage type of activation tensors (DLB_F16 / DLB_BF16 / DLB_F32); parameters,
 *     statistics and gradients of parameters are always fp32 (BN sums are fp64 accumulators).
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), performs no hidden
 *     synchronisation and no device allocation, and is CUDA-graph capturable.
 *   - kernels are launched with programmatic stream serialization and execute griddepcontrol.wait before their
 *     first global access: consecutive calls on one stream overlap launch latency, never data.  Kernels the CALLER
 *     enqueues in between see ordinary stream order.  DLB_PDL=0 in the environment disables the attribute.
 *   - return value: 0 (DLB_OK) or a negative dlb_status; dlb_last_error() gives the thread-local message.
 */
#ifndef DEEPLAB_B200_H_
#define DEEPLAB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLB_ABI_VERSION 2

typedef enum { DLB_OK = 0, DLB_ERR_INVALID = -1, DLB_ERR_UNSUPPORTED = -2, DLB_ERR_CUDA = -3 } dlb_status;
typedef enum { DLB_F16 = 0, DLB_BF16 = 1, DLB_F32 = 2 } dlb_dtype;
typedef enum { DLB_ACT_NONE = 0, DLB_ACT_RELU = 1, DLB_ACT_RELU6 = 2 } dlb_act;

int dlb_version(void);
const char* dlb_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
int64_t dlb_launch_count(void);
/* 1 if a CUDA device of compute capability 10.x is current, 0 otherwise (no compute is done) */
int dlb_device_ok(void);

/* ---------------------------------------------------------------------------------------------------
 * Training-mode BatchNorm finalisation done by the CONSUMER of the statistics instead of a kernel of its own
 * (dlb_bn_finalize below was 54 single-wave launches per training step in round 1).  A kernel that is handed a
 * dlb_bn_fin computes scale / shift for the channels it needs from the fp64 sums in its prologue -- the arithmetic of
 * dlb_bn_finalize -- and one of its CTAs also writes scale / shift / mean / rstd (the backward pass reads them) and
 * updates the moving statistics.  sum / sqs are NOT cleared (every CTA reads them): zero the accumulators once per
 * step before the forward pass.  moving_* / mean / rstd may be NULL.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const double* sum; const double* sqs;     /* [C] fp64 sums of x and x*x over `count` elements */
  const float* gamma; const float* beta;
  float eps, momentum;
  double count;
  float* moving_mean; float* moving_var;
  float* scale; float* shift; float* mean; float* rstd;
} dlb_bn_fin;

/* ---------------------------------------------------------------------------------------------------
 * Pointwise (1x1) convolution = GEMM on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA).
 * Replaces every Conv2D(.., (1,1)) of the graph: deeplabv3p.py:78-82, :175-177, :194-196, :385, :406,
 * :420, :438; utils.py:189; subpixel.py:90-91 (the conv half of Subpixel).
 *
 *   C[M, N] = epilogue( A[M, K] * Bt[N, K]^T )
 *
 * A  : activations, row pitch lda elements (16-byte multiple), dtype f16/bf16 -> tcgen05 kind::f16;
 *      dtype f32 -> exact-fp32 SIMT path (parity mode).
 * Bt : weights stored [N, K] (K contiguous), same dtype as A, row pitch ldb.
 * epilogue (all optional, applied in this order, in fp32):
 *      v = acc * col_scale[n] + col_shift[n] + row_bias[(m / rows_per_img) * ld_row_bias + n]
 *      stats: sum[n] += v, sqs[n] += v*v     (v rounded to out dtype first; training-mode BatchNorm, K13)
 *      v = act(v) ; v += R[m, n]
 *      store C[m, n] (out_dtype), n < n_store.  If shuffle_r > 0 the store is the Subpixel phase shift:
 *      column j = (jj*r + i)*Cs + k of row (b, a, bb) goes to out[b, a*r+jj, bb*r+i, k]  (subpixel.py:77-88
 *      with the weight columns pre-permuted from k*r*r + i*r + jj by the host).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int M, N, K;
  int dtype;            /* of A and Bt */
  int out_dtype;        /* of C and R */
  const void* A;  int lda;
  const void* Bt; int ldb;
  void* C;        int ldc;
  int n_store;          /* columns of C written, <= ldc; multiple of 8 (16-bit out) or 4 (f32 out) */
  const float* col_scale;   /* [N] or NULL */
  const float* col_shift;   /* [N] or NULL */
  const float* row_bias;    /* [M / rows_per_img, ld_row_bias] or NULL */
  int rows_per_img, ld_row_bias;
  int act;
  const void* R;  int ldr;  /* residual or NULL */
  double* stat_sum; double* stat_sqs;   /* [N] fp64 accumulators (atomically added) or NULL */
  int shuffle_r, shuffle_h, shuffle_w;  /* Subpixel store: r, low-res H, W ; 0 = plain store */
  /* A-operand transform (training forward of the project conv, deeplabv3p.py:189-196): the GEMM reads the RAW output of
   * the previous layer and applies A := act(A * a_scale[k] + a_shift[k]) (BatchNorm affine + ReLU6) to every landed
   * tile in shared memory, so the normalised activation never exists in HBM.  NULL = A is used as is.  16-bit path:
   * needs a plain 16-bit output (no output affine / bias / residual / shuffle). */
  const float* a_scale;     /* [K] or NULL */
  const float* a_shift;     /* [K] (required with a_scale) */
  int a_act;
  /* fp32 operands run as 3xTF32 on the tensor cores (x = hi + lo, hi = x with the 13 low mantissa bits cleared; three
   * tcgen05 kind::tf32 MMAs per k-step).  The kernel splits the landed tiles itself; for constant weights (inference)
   * the caller may pass them pre-split: Bt = hi parts, Bt_lo = remainders (same shape / pitch), which removes two
   * thirds of the in-kernel split work.  NULL = split in the kernel. */
  const void* Bt_lo;
  /* A-operand transform whose scale / shift the kernel derives from the producing layer's batch statistics (instead
   * of a_scale / a_shift; a_act still applies).  16-bit tensor-core path only, else DLB_ERR_INVALID. */
  const dlb_bn_fin* a_fin;
} dlb_pw_gemm_params;
int dlb_pw_gemm(const dlb_pw_gemm_params* p, void* stream);
/* Tiling plan dlb_pw_gemm would use for a 16-bit GEMM of this shape (host arithmetic only, no device needed):
 * plan[19] = {epilogue warp sets, chunk_n, n_chunks, chunks_per_group, n_groups, accumulator columns per stage,
 *             accumulator stages, alt_tiles, smem pipeline stages, dynamic shared memory bytes, grid size,
 *             CTAs serving column group 0..7}.  A CTA serves ONE column group for the whole kernel (its epilogue
 *             warps keep that group's BatchNorm statistics in registers); groups start on multiples of 64 columns so
 *             the staged 32 x 64 output tiles leave through bulk tensor stores.  Bit 16 of shuffle_r selects the plan
 *             of the A-operand-transform instance (one warp set rewrites A tiles), bit 17 the plan of the fp32-operand
 *             (3xTF32) instance.  0 or DLB_ERR_INVALID. */
int dlb_pw_gemm_plan(int M, int N, int K, int out_dtype, int shuffle_r, int* plan);

/* Weight gradient of a 1x1 convolution: dW[K, N] (+)= A[M, K]^T * dY[M, N]  (fp32 out, ld = N).
 * tcgen05 with MN-major operands for 16-bit dtypes; SIMT for f32.  beta = 0 overwrites, 1 accumulates (only 0 / 1).
 * Split-M partial tiles are added into dW with fp32 reductions in L2, so the last bits depend on the arrival order. */
typedef struct {
  int M, N, K;
  int dtype;
  const void* A;  int lda;
  const void* dY; int ldy;
  float* dW;      int ldw;
  float* dbias;             /* [N] column sums of dY, or NULL */
  float beta;
  void* workspace;          /* unused since the partials are reduced in L2 (dlb_pw_wgrad_workspace_bytes() = 0); may be NULL */
  int64_t workspace_bytes;
  /* A-operand transform, as in dlb_pw_gemm: A := act(A * a_scale[k] + a_shift[k]) on the landed tiles (or NULL) */
  const float* a_scale;
  const float* a_shift;
  int a_act;
} dlb_pw_wgrad_params;
int dlb_pw_wgrad(const dlb_pw_wgrad_params* p, void* stream);
int64_t dlb_pw_wgrad_workspace_bytes(int M, int N, int K);

/* ---------------------------------------------------------------------------------------------------
 * Depthwise 3x3 convolution, stride 1/2, any dilation, TF "SAME" or explicit padding.
 * Replaces DepthwiseConv2D at deeplabv3p.py:73-74 and :186-188 (+ ZeroPadding2D :61-69).
 *   x  : [B, H, W, C] raw producer output; the consumer-side prologue a = act(x*in_scale[c] + in_shift[c])
 *        (BatchNorm + ReLU6 of the *previous* layer, deeplabv3p.py:178-181) is applied on load when
 *        in_scale != NULL; padding is applied to a (zeros), exactly as the reference pads the activated map.
 *   y  : [B, Ho, Wo, C] raw conv output, optional fp64 per-channel sum / sum-of-squares for its own BN.
 *   pad_top / pad_left: rows/cols of zeros before the first pixel (TF SAME: pad_total // 2).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int B, H, W, C, Ho, Wo;
  int stride, dilation, pad_top, pad_left;
  int dtype;
  const void* x; void* y;
  const float* w;                 /* [3, 3, C] fp32 (Keras depthwise_kernel (3,3,C,1)) */
  const float* in_scale; const float* in_shift; int in_act;
  /* optional inference epilogue y = act(y*out_scale + out_shift) (folded BN) */
  const float* out_scale; const float* out_shift; int out_act;
  double* stat_sum; double* stat_sqs;
  const dlb_bn_fin* in_fin;       /* optional: in_scale / in_shift from the batch statistics (dlb_bn_fin above) */
} dlb_dw_conv_params;
int dlb_dw_conv_fwd(const dlb_dw_conv_params* p, void* stream);

/* backward: dx_act[B,H,W,C] = d loss / d a (gradient w.r.t. the activated input) and dw[3,3,C] += ... */
typedef struct {
  int B, H, W, C, Ho, Wo;
  int stride, dilation, pad_top, pad_left;
  int dtype;
  const void* x;                  /* raw input as in forward (prologue re-applied) */
  const void* dy;                 /* [B, Ho, Wo, C] */
  void* dx;                       /* [B, H, W, C] or NULL */
  const float* w;
  float* dw;                      /* [3,3,C] fp32, atomically accumulated (caller zeroes) or NULL */
  const float* in_scale; const float* in_shift; int in_act;
} dlb_dw_conv_bwd_params;
int dlb_dw_conv_bwd(const dlb_dw_conv_bwd_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Stem: Lambda(x/127.5 - 1) + Conv2D(32, 3, strides 2, 'same', no bias)  (deeplabv3p.py:270, :317-321;
 * Xception entry_flow_conv1_1 :283-284).  Input fp32 NHWC [B,H,W,3] in 0..255, TF-SAME padding (0,1).
 * 16-bit dtypes run on warp-level tensor-core MMAs with the exact operand x - 127.5 (integer inputs assumed, as the
 * reference's uint8 images are) and hi+lo split weights; DLB_F32 is the exact SIMT parity path.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int B, H, W, Cout, Ho, Wo;
  int dtype;                      /* of y */
  const float* x; void* y;
  const float* w;                 /* [3,3,3,Cout] HWIO fp32 */
  const float* out_scale; const float* out_shift; int out_act;   /* inference folded BN, or NULL */
  double* stat_sum; double* stat_sqs;
} dlb_stem_conv_params;
int dlb_stem_conv_fwd(const dlb_stem_conv_params* p, void* stream);
/* dw[3,3,3,Cout] += sum x_pre * dy  (dy: [B,Ho,Wo,Cout] dtype) */
int dlb_stem_conv_wgrad(int B, int H, int W, int Cout, int dtype, const float* x, const void* dy, float* dw,
                        void* stream);

/* Xception-only helpers (backbone='xception', deeplabv3p.py:272-313, decoder :414-429):
 *   dense 3x3 stride-1 SAME conv (entry_flow_conv1_2, :287), [3,3,Cin,Cout] HWIO fp32 weights, folded BN + act epilogue;
 *   pixel subsampling = the input gather of a 1x1 stride-2 'valid' conv (_conv2d_same with kernel_size=1, :106-116);
 *   legacy TF1 bilinear resize of a feature map (:418), output written with channel pitch ldo (concat slice). */
int dlb_conv3x3_fwd(int B, int H, int W, int Cin, int Cout, int dtype, const void* x, const float* w, void* y,
                    const float* out_scale, const float* out_shift, int out_act, void* stream);
int dlb_subsample(int B, int H, int W, int C, int step, int dtype, const void* x, void* y, void* stream);
int dlb_resize_bilinear(int B, int h, int w, int C, int H, int W, int ldo, int dtype, const void* x, void* y,
                        void* stream);
/* Fused ASPP atrous depthwise stage (deeplabv3p.py:392-399): the depthwise 3x3 + BN + ReLU halves of aspp1..3
 * (three dilation rates) in one pass over x [B,H,W,C]: reads x once, writes y[0..2] [B,H,W,C].
 * w[i]: [3,3,C] fp32, scale/shift[i]: folded depthwise BN.  H*W*32 bytes must fit in shared memory (<= 80x80). */
int dlb_aspp_dw3_fwd(int B, int H, int W, int C, int dtype, const void* x, const float* const* w, const int* rates,
                     const float* const* scale, const float* const* shift, void* const* y, void* stream);

/* Fused separable-conv branches (tcgen05): for every branch i
 *     out[i][m, 0:N] = pw_act( (dw_act( dw3x3_{rate[i]}(x) * dw_scale + dw_shift ) @ w_pw[i]^T) * pw_scale[i] + pw_shift[i] )
 * in ONE kernel -- the depthwise result is produced in shared memory as the tensor-core A operand and never written.
 * Replaces the spatial ASPP branches aspp0 (rate 0 = plain 1x1, no depthwise stage) and aspp1..3 = SepConv_BN(x, 256,
 * rate=atrous_rates[i], depth_activation=True, epsilon=1e-5) (deeplabv3p.py:385-399, SepConv_BN :47-84; stride 1,
 * TF 'same' zero padding) and, with one branch, decoder_conv0/1 (:426-429), BatchNorms folded (inference).
 * x: [B,H,W,C] NHWC f16/bf16, W <= 128, C % 8 == 0.  w_pw[i]: [N, C] (C contiguous), N a multiple of 8, 8..256.
 * out[i]: rows of pitch ldc (a channel slice of the concat buffer).  dw_pack: dlb_sepconv_pack_dw() output for the
 * branches with rate > 0, in branch order.
 * With res[i] != NULL the kernel is also the inference form of the second half of _inverted_res_block
 * (deeplabv3p.py:186-206): out = res + BN(project(relu6(BN(depthwise(x))))) with dw_act = RELU6, pw_act = NONE -- the
 * 6C-wide expanded activation is then read once and its depthwise result never exists in memory. */
typedef struct {
  int B, H, W, C, N;
  int dtype;
  int n_branches;             /* 1..4 */
  int rates[4];               /* 0 = pointwise-only branch */
  const void* x;
  const void* w_pw[4];
  const void* dw_pack;
  const float* pw_scale[4];   /* [N] folded pointwise BN (NULL = 1) */
  const float* pw_shift[4];   /* [N] (NULL = 0) */
  void* out[4];
  int ldc;
  int dw_act, pw_act;         /* dlb_act after the depthwise BN / pointwise BN */
  const void* res[4];         /* optional residual [B,H,W,N] rows of pitch ldr, added after pw_act (NULL = none) */
  int ldr;
} dlb_sepconv_fused_params;
int dlb_sepconv_fused_fwd(const dlb_sepconv_fused_params* p, void* stream);
/* Packs depthwise kernels w_dw[i] [3,3,C] fp32 + folded BN scale/shift[i] [C] into the per-64-channel-chunk blocks the
 * fused kernel bulk-copies to shared memory (16-bit taps, fp32 affine); pack holds dlb_sepconv_pack_bytes() bytes. */
int64_t dlb_sepconv_pack_bytes(int C, int n_branches);
int dlb_sepconv_pack_dw(int C, int dtype, int n_branches, const float* const* w_dw, const float* const* scale,
                        const float* const* shift, void* pack, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * BatchNormalization (deeplabv3p.py:76,80,178,189,197,322,379,386,408), training mode = batch statistics.
 * ------------------------------------------------------------------------------------------------- */
/* From fp64 sums over `count` elements: mean, biased var -> scale = gamma*rstd, shift = beta - mean*scale;
 * moving stats m = m*momentum + (1-momentum)*batch (variance fed unbiased, Keras 2.2.4 TF backend).
 * Zeroes sum/sqs afterwards when `reset` != 0.  moving_* may be NULL (frozen layer). */
int dlb_bn_finalize(int C, double count, double* sum, double* sqs, const float* gamma, const float* beta,
                    float eps, float momentum, float* moving_mean, float* moving_var, float* scale, float* shift,
                    float* mean, float* rstd, int reset, void* stream);
/* Inference fold: scale = gamma / sqrt(moving_var + eps), shift = beta - moving_mean * scale. */
int dlb_bn_fold(int C, const float* gamma, const float* beta, const float* moving_mean, const float* moving_var,
                float eps, float* scale, float* shift, void* stream);
/* y = act(x*scale[c] + shift[c]) (+ res) (* dropout keep-mask / keep_prob); x,[res],y: [M, C] dtype */
typedef struct {
  int64_t M; int C; int dtype;
  const void* x; void* y; const void* res;
  const float* scale; const float* shift; int act;
  float drop_rate; uint64_t drop_seed;      /* Dropout(0.1), deeplabv3p.py:410; 0 = off */
  const int64_t* drop_seed_dev;             /* optional device counter added to drop_seed (graph replay) */
  const dlb_bn_fin* fin;                    /* optional: scale / shift from the batch statistics (dlb_bn_fin above) */
} dlb_bn_apply_params;
int dlb_bn_act_apply(const dlb_bn_apply_params* p, void* stream);
/* Backward through  a = dropout(act(z)), z = x*scale + shift  where scale/shift come from batch statistics.
 *   pass 1 (reduce): dbeta[c] = sum dz, dgamma[c] = sum dz * xhat     (fp64 atomics into red[2*C])
 *   pass 2 (apply) : dx = scale * (dz - dbeta/M - xhat * dgamma/M)     (batch-stat mode)
 *                    dx = scale * dz                                   (frozen_stats mode)
 * with dz = da * act'(z) * dropmask. */
typedef struct {
  int64_t M; int C; int dtype;
  const void* x;          /* raw conv output */
  const void* da;         /* gradient w.r.t. activated output */
  void* dx;               /* gradient w.r.t. raw conv output (pass 2) */
  const float* scale; const float* shift; const float* mean; const float* rstd; int act;
  double* red;            /* [2*C]: dbeta then dgamma */
  float* dgamma; float* dbeta;   /* fp32 results written by pass 2 (may be NULL) */
  float drop_rate; uint64_t drop_seed;
  int frozen_stats;
  const int64_t* drop_seed_dev;
} dlb_bn_bwd_params;
int dlb_bn_bwd_reduce(const dlb_bn_bwd_params* p, void* stream);
int dlb_bn_bwd_apply(const dlb_bn_bwd_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * ASPP image pooling (deeplabv3p.py:375-382): global average over H*W of act(x*scale+shift) -> [B, C] fp32,
 * and its backward (broadcast of dpool / (H*W) added into dx).
 * ------------------------------------------------------------------------------------------------- */
int dlb_global_avgpool_fwd(int B, int HW, int C, int dtype, const void* x, const float* in_scale,
                           const float* in_shift, int in_act, float* out, void* stream);
int dlb_global_avgpool_bwd(int B, int HW, int C, int dtype, const float* dout, void* dx, int accumulate,
                           void* stream);
/* small dense fp32 GEMM for the [B, C] pooled branch: C[M,N] = act(A[M,K] W[K,N] * scale + shift) etc.
 * generic tiny SIMT matmul: C = alpha * op(A) * op(B) + beta * C  (row-major, fp32) */
int dlb_small_gemm(int M, int N, int K, const float* A, int lda, int transA, const float* B, int ldb, int transB,
                   float* C, int ldc, float alpha, float beta, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Head: legacy TF1 bilinear resize (align_corners=False, no half-pixel centres; deeplabv3p.py:439,
 * utils.py:190) + softmax (deeplabv3p.py:440-444, utils.py:191-192) + the void-ignoring sparse
 * cross-entropy (utils.py:127-130, Keras clip 1e-7) and its gradient, fused.
 *   logits : [B, h, w, ldl] (C valid channels) dtype fp32;  scale = H/h = W/w integer (1 for Subpixel head)
 * ------------------------------------------------------------------------------------------------- */
/* probs[B, H*W, C] fp32 (model.predict output) and/or argmax[B, H*W] uint8 */
int dlb_resize_softmax_fwd(int B, int h, int w, int C, int ldl, int H, int W, const float* logits, float* probs,
                           uint8_t* argmax, void* stream);
/* training: per-pixel loss_i = -log clip(p[y_i]) (0 for void label == C), weighted:
 *   loss_sum  += sum_i sw_i * loss_i ; wcount += #(sw_i != 0)        (fp64 accumulators)
 *   dlogits[B,h,w,ldl] += resize^T( sw_i * grad_scale * (p - onehot) )   (caller zeroes dlogits)
 * grad_scale is read from device memory (*grad_scale_dev) so that 1/(N * mean(sw != 0)) can be produced by a
 * previous launch without a host round trip; labels are the reference's float [B, H*W, 1] tensor. */
typedef struct {
  int B, h, w, C, ldl, H, W;
  const float* logits; const float* labels; const float* sample_w;   /* sample_w may be NULL (= 1) */
  const float* grad_scale_dev;
  float* dlogits; double* loss_sum; double* wcount; uint8_t* argmax;
} dlb_softmax_ce_params;
int dlb_resize_softmax_ce(const dlb_softmax_ce_params* p, void* stream);
/* counts sample weights != 0 -> *grad_scale = loss_scale / (n_pix_total * mean(sw != 0)) = loss_scale / #(sw != 0);
 * loss_scale = the static factor x loss_scale_state[0] (the dynamic fp16 scale, see dlb_adam_step; may be NULL) */
int dlb_ce_grad_scale(int64_t n, const float* sample_w, float* grad_scale_dev, double* wcount, float loss_scale,
                      const float* loss_scale_state, void* stream);

/* Backward of the phase-shift store fused into dlb_pw_gemm (whose columns are ordered (jj, i, k)):
 *   dst[n, a, b, (jj*r + i)*Cs + k] = dlogits[n, a*r + jj, b*r + i, k], fp32 -> dst_dtype.  dlogits [B, h*r, w*r, Cs]. */
int dlb_subpixel_grad_gather(int B, int h, int w, int Cs, int r, const float* dlogits, int dst_dtype, void* dst,
                             void* stream);
/* Standalone Subpixel phase shift (subpixel.py:77-88): out[n, a*r+j, b*r+i, k] = in[n, a, b, k*r*r + i*r + j] */
int dlb_phase_shift(int B, int h, int w, int Cs, int r, int dtype, const void* in, void* out, int inverse,
                    void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Keras Adam on a flat fp32 parameter buffer (ipynb:107; Keras 2.2.4 optimizers.Adam):
 *   lr_t = lr / (1 + decay*iter) * sqrt(1 - b2^t) / (1 - b1^t) ; p -= lr_t * m / (sqrt(v) + eps)
 * `step_dev` is a device int64 iteration counter (incremented by the kernel) so the step is graph-replayable.
 * grads are multiplied by grad_mult first (1/world_size after the NCCL sum).
 * train_mask (optional, n floats): elements with mask 0 belong to layers with trainable=False -- Keras leaves them
 *   out of the update entirely, so neither p nor m / v are touched (utils.py / ipynb:147-155 freeze regime).
 * loss_scale_state (optional, device float[4] = {scale, good_steps, found_inf, growth_interval}): dynamic fp16 loss
 *   scaling.  Gradients are divided by scale; if found_inf != 0 (set by dlb_grad_finite_check) the whole update and
 *   the iteration counter are skipped and scale is halved, otherwise scale doubles every growth_interval good steps.
 * ------------------------------------------------------------------------------------------------- */
int dlb_adam_step(int64_t n, float* param, const float* grad, float* m, float* v, int64_t* step_dev, float lr,
                  float beta1, float beta2, float eps, float decay, float grad_mult, const float* train_mask,
                  float* loss_scale_state, void* stream);
/* sets loss_scale_state[2] = 1 if any of the n gradients (n % 4 == 0, 16-byte aligned) is inf / NaN */
int dlb_grad_finite_check(int64_t n, const float* grad, float* loss_scale_state, void* stream);
/* fp32 [K, N] master weight -> 16-bit W[K,N] and Wt[N,K] copies used by the GEMMs (either may be NULL) */
int dlb_cast_weight(int K, int N, const float* w, int dtype, void* w_kn, void* w_nk, void* stream);
/* The same for every 1x1 layer of the model in ONE launch (after each optimizer step).  `table` is a DEVICE array of
 * n_entries x 8 int64: { w (const float*), w_kn (void* or 0), w_nk (void* or 0), K, N, ld_kn (row pitch of w_kn in
 * elements, >= N), dtype (DLB_F16 / DLB_BF16 / DLB_F32) of the copies, first flat element index of the entry };
 * total = sum of K*N. */
int dlb_cast_weights_batched(int n_entries, const int64_t* table, int64_t total, void* stream);
int dlb_cast(int64_t n, int src_dtype, const void* src, int dst_dtype, void* dst, void* stream);
int dlb_fill_zero(void* p, int64_t bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Metrics on the argmax map (utils.py:132-157): confusion counts per image, conf[B, C+1, C] int64
 * (row = true label incl. void, col = predicted).  Jaccard / accuracy are finished on the host from it.
 * ------------------------------------------------------------------------------------------------- */
int dlb_confusion(int B, int64_t npix, int C, const float* labels, const uint8_t* argmax, unsigned long long* conf,
                  void* stream);

/* ---------------------------------------------------------------------------------------------------
 * SegmentationGenerator label contract (utils.py:360-399), the step right before the hot path: void remap
 * (labels outside 0..n_classes-1 -> n_classes) into y[B, npix] and per-image 'balanced' class weights
 * sw[B, npix] = n_valid / (n_present * count[label]) (0 for void), float64 arithmetic stored as float32 --
 * bit-exact with sklearn.utils.class_weight.compute_class_weight('balanced').  label_type: 0 = uint8, 1 = int32,
 * 2 = float32.  counts: caller-owned workspace [B, n_classes + 1] uint64 (zeroed here).  y or sw may be NULL.
 * ------------------------------------------------------------------------------------------------- */
int dlb_label_weights(int B, int64_t npix, int n_classes, int label_type, const void* labels,
                      unsigned long long* counts, float* y, float* sw, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * SegmentationGenerator.__getitem__'s augmentations on the device (utils.py:319-365), bit-exact with OpenCV's uint8
 * paths: GaussianBlur((k,k), 0) for k in {3,5,7} -> flips -> gamma LUT -> warpAffine (INTER_LINEAR for image AND label,
 * constant border 0) -> "labels the decoded image did not contain become void" -> float32 image + label map (feed
 * label_out to dlb_label_weights for Y / SW).  The random decisions stay with the caller (the reference draws them from
 * Python's `random`); they arrive as one dlb_aug_params per image in DEVICE memory.
 *   img [B,H,W,3] uint8 (cv2 BGR), label [B,H,W] uint8, luts [B,256] uint8 or NULL (identity),
 *   tmp_img [B,H,W,3] / tmp_label [B,H,W] uint8 scratch, present [B,8] uint32 scratch,
 *   X [B,H,W,3] float32 out, label_out [B,H,W] uint8 out (values 0..n_classes).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int hflip, vflip;          /* cv2.flip(.., 1) / cv2.flip(.., 0) */
  int blur_ksize;            /* 0 = none, else 3 / 5 / 7 */
  int warp;                  /* 0 = none */
  double minv[6];            /* INVERSE of the 2x3 affine matrix passed to cv2.warpAffine (cv2.invertAffineTransform) */
} dlb_aug_params;
int dlb_augment_batch(int B, int H, int W, int n_classes, const uint8_t* img, const uint8_t* label,
                      const dlb_aug_params* params_dev, const uint8_t* luts, uint8_t* tmp_img, uint8_t* tmp_label,
                      unsigned int* present, float* X, uint8_t* label_out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Dense CRF (utils.py:74-91 -> pydensecrf DenseCRF2D.inference): permutohedral lattice mean field.
 *   unary   : [M, N] fp32 energies (N = H*W pixels, label-major as pydensecrf)
 *   image   : [H, W, 3] uint8
 *   Q       : [M, N] fp32 marginals out
 * Workspace sizes are queried first; the lattice (hash table, blur neighbours, barycentric weights) is built
 * on the device for the Gaussian (d=2) and bilateral (d=5) kernels and reused by all iterations.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int H, W, M, iters;
  float sxy_gauss, compat_gauss;             /* addPairwiseGaussian(sxy=3, compat=3)   utils.py:82 */
  float sxy_bilat, srgb_bilat, compat_bilat; /* addPairwiseBilateral(80, 13, compat=10) utils.py:85 */
  int unary_layout;   /* 0: `unary` = energies, label-major [M, N] (pydensecrf setUnaryEnergy);
                         1: `unary` = class probabilities, pixel-major [N, M] -- the network's softmax output as it
                            lies in device memory; -U = log p (soft unaries, no argmax -> unary_from_labels round trip) */
} dlb_crf_config;
int64_t dlb_crf_workspace_bytes(const dlb_crf_config* cfg);
int dlb_crf_inference(const dlb_crf_config* cfg, const float* unary, const uint8_t* image, float* Q,
                      uint8_t* map_out /* [N] argmax or NULL */, void* workspace, int64_t workspace_bytes,
                      void* stream);
/* The same for a batch of independent images in ONE set of launches (batch = grid.y of every kernel):
 *   unary [B, M, N], image [B, H, W, 3], Q [B, M, N], map_out [B, N] or NULL.
 * The Gaussian lattice depends on (H, W, sxy) only and is built once for the whole batch; bilateral lattices are per
 * image.  Per iteration: 2 splats (CSR gathers; lattice vertices are numbered in raster order of their first pixel so
 * the gathers are cache-coherent), 3 + 6 blur passes, and ONE slice kernel that gathers both lattices, adds the Potts
 * messages to -U and normalises (softmax) -- 12 launches per iteration for the whole batch. */
int64_t dlb_crf_workspace_bytes_batched(const dlb_crf_config* cfg, int batch);
int dlb_crf_inference_batched(const dlb_crf_config* cfg, int batch, const float* unary, const uint8_t* image, float* Q,
                              uint8_t* map_out, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPLAB_B200_H_ */
