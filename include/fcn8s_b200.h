/*
 * fcn8s_b200.h -- C ABI of libfcn8s_sm100.so: the B200 (sm_100a) kernels behind the FCN-8s forward+backward path.
 *
 * The reference (pierluigiferrari/fcn8s_tensorflow) has no FFI: its hot path is the set of TensorFlow-1.x ops that
 * `tf.Session.run` executes for the graph built in fcn8s_tensorflow.py.  Each entry point below replaces one group
 * of those ops; the comment above each names the reference call site (file:line under the reference root) whose
 * arithmetic it implements.  A maintainer binds them with ctypes; this repo's own binding is
 * fcn8s_tensorflow_b200/_capi.py.
 *
 * Conventions
 *   - every function returns 0 on success or a negative FCN8_ERR_* code; fcn8_last_error() gives the message
 *     (thread-local).
 *   - the library never allocates, frees or synchronises: all pointers are device pointers owned by the caller
 *     (torch.Tensor.data_ptr()), 16-byte aligned, NHWC contiguous; `stream` is a cudaStream_t passed as void*.
 *   - `dtype` selects the activation storage / tensor-core operand type: FCN8_BF16 (bf16 operands, kind::f16) or
 *     FCN8_F32 (fp32 storage, kind::tf32; with `nseg == 3` the error-compensated 3xTF32 product hi*hi+hi*lo+lo*hi).
 *     Accumulation is always fp32 (TMEM).  Parameters, gradients, Adam state and the decoder are fp32.
 *   - workspaces are caller-provided; the *_workspace_bytes query never launches anything.
 */
#ifndef FCN8S_B200_H_
#define FCN8S_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCN8_VERSION 100

enum {
  FCN8_OK = 0,
  FCN8_ERR_BAD_SHAPE = -1,
  FCN8_ERR_BAD_ALIGN = -2,
  FCN8_ERR_WORKSPACE = -3,
  FCN8_ERR_CUDA = -4,
  FCN8_ERR_UNSUPPORTED = -5
};

/* activation storage formats.  FCN8_BF16X2 ("fp32-equivalent"): a tensor [N,H,W,2C] of bf16 whose first C channels
 * of every pixel are hi = bf16(v) and the next C channels lo = bf16(v - hi); v' = hi + lo carries 16-17 mantissa
 * bits and the GEMMs form hi*hi + hi*lo + lo*hi on the bf16 tensor cores (nseg = 3). */
enum { FCN8_BF16 = 0, FCN8_F32 = 1, FCN8_BF16X2 = 2 };

/* epilogue flags of fcn8_conv_gemm (bit-or) */
enum {
  FCN8_EPI_BIAS = 1,
  FCN8_EPI_RELU = 2,
  FCN8_EPI_DROPOUT = 4,
  FCN8_EPI_MASK = 8,
  FCN8_EPI_RESIDUAL = 16,
  FCN8_EPI_ROUND_TF32 = 64, /* FCN8_F32 only: round outputs to the nearest tf32 (the MMA truncates its operands) */
  FCN8_EPI_COLSUM = 128     /* colsum[co] += sum over pixels of the stored values: the BiasAddGrad of the layer whose
                               dY this dgrad call produces, fused into its epilogue (fp32 atomics, caller zeroes) */
};

int32_t fcn8_version(void);
const char* fcn8_last_error(void);
/* 0 iff device `dev` is compute capability 10.x (B200); the kernels are sm_100a-only. */
int32_t fcn8_device_check(int32_t dev);
/* number of CUDA kernels this library has launched in this process so far (what bench.py reports as gpu_launches). */
uint64_t fcn8_launch_count(void);
/* Persistent GEMM kernels launch min(tiles, #SMs) CTAs with static round-robin tile assignment; while a collective
 * (the overlapped gradient all-reduce) occupies some SMs, n > 0 caps the grids at n CTAs so that no CTA has to wait for
 * an SM and run a whole second wave.  n = 0 restores the device's SM count.  Affects launches issued afterwards. */
int32_t fcn8_set_sm_limit(int32_t n);
/* Host-side CRC-32C (Castagnoli) of `n` bytes, continuing from `crc` (0 to start): the checksum of TensorFlow's
 * tensor-bundle checkpoint files (fcn8s_tensorflow.py:74,134,922-934 read / write them through tf.saved_model and
 * tf.train.Saver); used by fcn8s_tensorflow_b200/tf_bundle.py.  No device work. */
uint32_t fcn8_crc32c(const void* data, size_t n, uint32_t crc);
/* Bring-up / measurement knobs -- tests and profiling only (also settable as FCN8_DEBUG="key=value,..." at load time):
 *   0 = 1: no round-toward-zero compensation of the GEMM accumulators (the library multiplies every tcgen05
 *          accumulator by 1 + n_mma * 2.1e-8, the expected relative loss of TMEM's truncating accumulation over n_mma
 *          instructions; scripts/bringup.py::rz_accumulation_probe measures the constant)
 *   1 = 1: no wgrad_halo_kernel          2 = 1: no resident-weights variant of conv_halo_kernel<64>
 *   3    : bit 0: conv epilogues skip their output stores, bit 1: ... and their mask / residual loads (timing only)
 *   5 = 1: single-CTA kernels instead of the CTA-pair (cta_group::2) variants of conv_gemm / wgrad_gemm
 *   6 = 1: launch every kernel with the programmatic-dependent-launch attribute (kernels.h) */
int32_t fcn8_debug_set(int32_t key, int32_t value);
/* Measurement only: `buf` = device buffer of slots*148*8 int64; every following fcn8_conv_gemm / fcn8_wgrad_gemm launch
 * takes the next slot and its CTAs write their MMA-warp wait-cycle counters there (csrc/conv_gemm.cuh,
 * ConvGemmArgs::dbg).  NULL switches it off. */
int32_t fcn8_debug_buffer(void* buf, int32_t slots);

/* ---- feed: fcn8s_tensorflow.py:558,686,765 (image_input) + the encoder graph's RGB->BGR / mean subtraction [EXT].
 * uint8 RGB [N,H,W,3] -> im2col of the mean-subtracted BGR image for conv1_1: out[N,H,W,KP], column tap*3+c
 * (tap = kh*3+kw, c in B,G,R order) = value at (y+kh-1, x+kw-1) or 0 outside (SAME padding), columns 27..KP-1 = 0.
 * KP = 64 (bf16) / 32 (f32).  conv1_1 then runs as a 1x1 fcn8_conv_gemm over it. */
typedef struct {
  const uint8_t* images;
  void* out;
  int32_t N, H, W;
  int32_t dtype;
} Fcn8PreprocessParams;
int32_t fcn8_preprocess_im2col(const Fcn8PreprocessParams* p, void* stream);

/* ---- encoder convolutions (external VGG-16 graph loaded at fcn8s_tensorflow.py:127-152; conv kxk stride 1 SAME):
 * out[N,H,W,Cout] = epilogue( sum_{kh,kw,ci} x[N, y+kh-pad, x+kw-pad, ci] * wp[co][(kh*k+kw)*Cin + ci] ).
 * fprop: wp from fcn8_pack_weights(mode 0), flags BIAS|RELU(|DROPOUT).  dgrad (autodiff of the same op, :257):
 * x = dY, wp from fcn8_pack_weights(mode 1), flags MASK (ReLU/dropout backward against `mask_src`) and/or RESIDUAL.
 * nseg == 3: x_lo / wp_lo are the low halves of the operands (tf32 split: fcn8_split_tf32; bf16 pairs: the lo planes);
 * nseg == 2 forms only x*w + x*w_lo (x rounded to its high half, weights exact), nseg == 1 only x*w -- the reduced
 * products of the measured "fp32 forward / 2- or 1-term backward" modes (Engine(backward_terms=...)). */
typedef struct {
  const void* x;
  const void* x_lo;
  const void* wp;
  const void* wp_lo;
  void* out;
  const float* bias;
  const void* mask_src;
  const void* residual;
  int32_t N, H, W, Cin, Cout, ksize;
  int32_t dtype, nseg, flags;
  float mask_scale;
  float keep_prob;
  uint32_t seed;
  int32_t force_splits; /* 0 = heuristic */
  int32_t force_bn;     /* 0 = heuristic, else 64/128/256 */
  /* --- optional (zero = the defaults above) --- */
  int32_t x_ld;         /* pixel stride of x / x_lo in elements (0 = Cin; 2*Cin for a FCN8_BF16X2 tensor) */
  int32_t out_ld;       /* pixel stride of out / out_lo / mask_src / residual (0 = Cout) */
  void* out_lo;         /* FCN8_BF16 only: also store lo = bf16(v - bf16(v)) (hi/lo pair output) */
  const void* residual_lo;
  /* w_mode 0: wp is the packed operand of fcn8_pack_weights.  FCN8_BF16 only: w_mode 1 (fprop) / 2 (dgrad): wp (and
   * wp_lo) is a bf16 copy of the weight tensor in its TF layout [k,k,Cin_w,Cout_w] read in place -- fprop: Cin_w = Cin,
   * Cout_w = Cout; dgrad: Cin_w = Cout, Cout_w = Cin (rotation and transposition happen in the TMA coordinates). */
  int32_t w_mode;
  float* colsum;        /* FCN8_EPI_COLSUM target, [Cout] fp32 */
  const uint32_t* seed_ptr; /* optional device scalar: dropout seed = *seed_ptr * 2 + seed (see fcn8_set_step_scalars) */
  int32_t algo;         /* 0 = heuristic; 1 = per-tap implicit GEMM; 2 = halo-tile kernel (3x3, bf16, w_mode 1/2: the
                           activation patch of a tile is loaded once with its halo and the nine taps are shifted UMMA
                           descriptors into it) */
} Fcn8ConvParams;
size_t fcn8_conv_gemm_workspace_bytes(const Fcn8ConvParams* p);
int32_t fcn8_conv_gemm(const Fcn8ConvParams* p, void* workspace, size_t workspace_bytes, void* stream);

/* ---- filter gradient (autodiff of the conv, fcn8s_tensorflow.py:257):
 * dw[(kh*k+kw)*Cin + ci][co] (fp32, TF HWIO order) = sum_{n,y,x} x[n,y+kh-pad,x+kw-pad,ci] * dy[n,y,x,co].
 * rows_valid < ksize*ksize*Cin keeps only the first rows (27 for the im2col'ed conv1_1).
 * accumulate != 0 adds into dw instead of overwriting. */
typedef struct {
  const void* x;
  const void* x_lo;
  const void* dy;
  const void* dy_lo;
  float* dw;
  int32_t N, H, W, Cin, Cout, ksize;
  int32_t rows_valid;
  int32_t dtype, nseg;
  int32_t force_splits;
  int32_t force_bn;
  int32_t x_ld;  /* pixel strides in elements, 0 = Cin / Cout */
  int32_t dy_ld;
} Fcn8WgradParams;
size_t fcn8_wgrad_gemm_workspace_bytes(const Fcn8WgradParams* p);
int32_t fcn8_wgrad_gemm(const Fcn8WgradParams* p, void* workspace, size_t workspace_bytes, void* stream);

/* ---- weight packing: fp32 master weights in TF layout [k,k,Cin,Cout] (HWIO) -> tensor-core operand layout.
 * mode 0 (fprop): out[co][tap*CinPad + ci] = w[tap][ci][co]           (CinPad >= Cin, zero filled)
 * mode 1 (dgrad): out[ci][tap'*Cout + co]  = w[taps-1-tap'][ci][co]    (180-degree rotation, in/out swapped)
 * dtype BF16: out bf16.  dtype F32: out_lo == NULL -> out = round_tf32(w); out_lo != NULL -> out = w (the MMA
 * truncates it to its tf32 high part) and out_lo = round_tf32(w - trunc_tf32(w)). */
typedef struct {
  const float* w;
  void* out;
  void* out_lo;
  int32_t ksize, Cin, Cout, CinPad;
  int32_t mode, dtype;
} Fcn8PackParams;
int32_t fcn8_pack_weights(const Fcn8PackParams* p, void* stream);

/* tf32 operand preparation (the tf32 MMA truncates fp32 operands to their top 19 bits):
 *   hi != NULL, lo != NULL: hi = trunc_tf32(x), lo = round_tf32(x - hi)
 *   hi == NULL:             lo only (x is its own high operand: the MMA truncates it)
 *   lo == NULL:             hi = round_tf32(x)  (single-pass tf32: makes the MMA's truncation exact) */
int32_t fcn8_split_tf32(const float* x, float* hi, float* lo, size_t n, void* stream);

/* ---- 2x2 / stride 2 SAME max pooling of the encoder graph [EXT] and its gradient.
 * bwd: dx = (first arg-max of the window && x > 0) ? dy : 0  -- MaxPoolGrad fused with the ReLUGrad of the producer. */
typedef struct {
  const void* x;  /* [N,H,W,C]   */
  void* y;        /* fwd: out [N,ceil(H/2),ceil(W/2),C];  bwd: dy (same shape) */
  void* dx;       /* bwd only */
  int32_t N, H, W, C;
  int32_t dtype;
  float* db;      /* bwd only, optional: db[c] += sum over pixels of dx (bias gradient of the producing conv) */
} Fcn8PoolParams;
int32_t fcn8_maxpool_fwd(const Fcn8PoolParams* p, void* stream);
int32_t fcn8_maxpool_bwd(const Fcn8PoolParams* p, void* stream);

/* ---- bias gradient: db[c] = sum_p dy[p][c]  (BiasAddGrad).  dy [P][C] in `dtype`, db fp32. */
typedef struct {
  const void* dy;
  float* db;
  int64_t P;
  int32_t C;
  int32_t dtype;
} Fcn8BiasGradParams;
size_t fcn8_bias_grad_workspace_bytes(const Fcn8BiasGradParams* p);
int32_t fcn8_bias_grad(const Fcn8BiasGradParams* p, void* workspace, size_t workspace_bytes, void* stream);

/* ---- decoder 1x1 score heads: fcn8s_tensorflow.py:171-200 (tf.multiply scale + tf.layers.conv2d 1x1 + bias).
 * fwd: s[p][c] = scale * sum_ci x[p][ci] * K[ci][c] + b[c]            (x in `dtype`, everything else fp32)
 * bwd: dK[ci][c] = scale * sum_p x[p][ci]*ds[p][c]; db[c] = sum_p ds[p][c];
 *      dx[p][ci] = scale * sum_c ds[p][c]*K[ci][c]  (* (x>0)*mask_scale if mask != 0: dropout+ReLU backward of fc7) */
typedef struct {
  const void* x;
  const float* K;
  const float* b;
  float* s;        /* fwd out / bwd: ds in */
  float* dK;
  float* db;
  void* dx;        /* `dtype` */
  int64_t P;
  int32_t Cin, C;
  float scale;
  int32_t dtype;
  int32_t mask;
  float mask_scale;
} Fcn8HeadParams;
size_t fcn8_score_head_fwd_workspace_bytes(const Fcn8HeadParams* p);
int32_t fcn8_score_head_fwd(const Fcn8HeadParams* p, void* workspace, size_t workspace_bytes, void* stream);
size_t fcn8_score_head_bwd_workspace_bytes(const Fcn8HeadParams* p);
int32_t fcn8_score_head_bwd(const Fcn8HeadParams* p, void* workspace, size_t workspace_bytes, void* stream);

/* ---- transposed convolutions: fcn8s_tensorflow.py:204-233 (tf.layers.conv2d_transpose k=2s, stride s, 'same')
 * plus the skip adds :213,:224.   T[a][b][co][ci] (TF layout), all fp32.
 * fwd: y[n, s*i+a-p, s*j+b-p, co] = sum x[n,i,j,ci]*T[a,b,co,ci] + bias[co] (+ skip[...]),  p = s/2.
 * bwd: dx = strided conv of dy with T;  dT[a,b,co,ci] = sum x*dy;  dbias = sum dy. */
typedef struct {
  const float* x;  /* [N,h,w,C] */
  const float* T;  /* [2s,2s,C,C] */
  const float* bias;
  const float* skip; /* [N,h*s,w*s,C] or NULL */
  float* y;        /* fwd out / bwd: dy in */
  float* dx;
  float* dT;
  float* dbias;
  int32_t N, h, w, C, stride;
} Fcn8UpscoreParams;
int32_t fcn8_upscore_fwd(const Fcn8UpscoreParams* p, void* stream);
size_t fcn8_upscore_bwd_workspace_bytes(const Fcn8UpscoreParams* p);
int32_t fcn8_upscore_bwd(const Fcn8UpscoreParams* p, void* workspace, size_t workspace_bytes, void* stream);

/* ---- loss / predictor: fcn8s_tensorflow.py:253 (softmax_cross_entropy_with_logits + reduce_mean), :268-269.
 * logits (and dlogits, same layout): pixel (n,y,x) at ((n*(H+2*pad) + y+pad)*(W+2*pad) + x+pad)*CP, CP >= C -- pad = 0,
 * CP = C is the dense [N,H,W,C] tensor; pad = 4, CP = fcn8_upscore_tc_cp(C, 8) is the padded output of
 * fcn8_upscore_tc_fwd.  labels: uint8 dense [N,H,W,C] one-hot as yielded by the generators (numpy bool), used as fp32
 * weights y_c like TF does.
 * loss_sum (fp32 scalar, accumulated; caller zeroes) += sum_p (sum_c y_c)*lse(z) - sum_c y_c z_c;
 * dlogits[p][c] = (softmax_c * sum_c y_c - y_c) * grad_scale   (grad_scale = 1/(N*H*W)), channels C..CP-1 = 0;
 * dbias[c] (accumulated; caller zeroes) += sum_p dlogits[p][c]  (bias gradient of the last transposed conv);
 * softmax fp32 dense [N,H,W,C] / argmax int64 dense [N,H,W] (first max) for predict().  Any output may be NULL;
 * softmax and dlogits are mutually exclusive. */
typedef struct {
  const float* logits;
  const uint8_t* labels;
  float* loss_sum;
  float* dlogits;
  float* dbias;
  float* softmax;
  int64_t* argmax;
  int32_t N, H, W, C, CP, pad;
  float grad_scale;
} Fcn8SoftmaxParams;
int32_t fcn8_softmax_xent(const Fcn8SoftmaxParams* p, void* stream);

/* ---- transposed convolutions on the tensor cores (tf.layers.conv2d_transpose k=2s, stride s, 'same',
 * fcn8s_tensorflow.py:204-233, and its autodiff :257) as phase GEMMs on tcgen05 (kind::tf32; nseg = 3: 3xTF32):
 * an s x s block of outputs depends on a 2x2 input neighbourhood, so with rows = blocks (J,I), J in [0,h], I in [0,w],
 *   Zblock[(dy,dx,co)] = sum_{(ty,tx,ci)} x[n, J-1+ty, I-1+tx, ci] * T[dy+s(1-ty), dx+s(1-tx), co, ci] + bias[co].
 * The blocks tile a PADDED tensor zp[N, s*(h+1), s*(w+1), CP] whose interior [s/2 : s/2+s*h, s/2 : s/2+s*w] is the
 * transposed convolution's output (pixel (oy,ox) at zp[n, oy+s/2, ox+s/2, :]); CP = fcn8_upscore_tc_cp(C, s).
 * The border of zp holds don't-care values after fwd and MUST be zero in the dz passed to dx / dw.
 * x / dx: [N,h,w,ldx] fp32, ldx a multiple of 4 >= C, channels >= C zero.
 * Operands come from fcn8_upscore_tc_pack (w_fwd [s*s*CP][128], w_dx [64][4*s*s*CP], bias_big [s*s*CP]; the *_lo
 * arrays are the low halves of the tf32 split, NULL for single-pass tf32). */
typedef struct {
  const float* T;     /* [2s,2s,C,C] (kh,kw,out,in) */
  const float* bias;  /* [C] */
  float* w_fwd;
  float* w_fwd_lo;
  float* w_dx;
  float* w_dx_lo;
  float* bias_big;
  int32_t C, stride;
} Fcn8UpscorePackParams;
typedef struct {
  const float* x;
  const float* x_lo;
  const float* w;     /* fwd: w_fwd, dx: w_dx */
  const float* w_lo;
  const float* bias_big;
  float* zp;          /* fwd: output; dx / dw: dz input */
  const float* zp_lo;
  float* dx;
  float* dT;          /* dw: [2s,2s,C,C] TF layout */
  int32_t N, h, wd, C, stride, ldx, nseg; /* x is [N,h,wd,ldx] */
} Fcn8UpscoreTcParams;
int32_t fcn8_upscore_tc_cp(int32_t C, int32_t stride);
/* Skip connections around the padded tensors (tf.add, fcn8s_tensorflow.py:213,224).  (h, w) are the INPUT dims of the
 * transposed convolution, its output is [N, stride*h, stride*w, .].
 * gather:  f[n,y,x,c] = zp[n, y+stride/2, x+stride/2, c] + skip[n,y,x,c] (skip may be NULL), c < C; channels C..ldf-1 = 0.
 * scatter: dzp interior = g (border / pad channels untouched: keep them zero); dbias[c] += sum over pixels of g. */
int32_t fcn8_upscore_tc_gather(const float* zp, const float* skip, float* f, int32_t N, int32_t h, int32_t w,
                               int32_t C, int32_t stride, int32_t ldf, int32_t ld_skip, void* stream);
int32_t fcn8_upscore_tc_scatter(const float* g, float* dzp, float* dbias, int32_t N, int32_t h, int32_t w, int32_t C,
                                int32_t stride, int32_t ldg, void* stream);
int32_t fcn8_upscore_tc_pack(const Fcn8UpscorePackParams* p, void* stream);
int32_t fcn8_upscore_tc_fwd(const Fcn8UpscoreTcParams* p, void* stream);
int32_t fcn8_upscore_tc_dx(const Fcn8UpscoreTcParams* p, void* stream);
size_t fcn8_upscore_tc_dw_workspace_bytes(const Fcn8UpscoreTcParams* p);
int32_t fcn8_upscore_tc_dw(const Fcn8UpscoreTcParams* p, void* workspace, size_t workspace_bytes, void* stream);

/* ---- streaming metrics: fcn8s_tensorflow.py:280-301 (labels_argmax, tf.metrics.mean_iou / accuracy); the device
 * analogue of cityscapesscripts/evaluation/addToConfusionMatrix_impl.c:3-16: conf[gt*C + pred] += 1 (uint64). */
int32_t fcn8_confusion_matrix(const int64_t* pred, const uint8_t* labels_onehot, unsigned long long* conf, int64_t P,
                              int32_t C, void* stream);

/* ---- optimiser: tf.train.AdamOptimizer.minimize, fcn8s_tensorflow.py:256-257 (TF form: eps outside sqrt(v)):
 * g' = g*grad_scale;  m = b1*m + (1-b1)*g';  v = b2*v + (1-b2)*g'^2;  p -= lr_t * m / (sqrt(v) + eps),
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed by the caller. */
int32_t fcn8_adam(float* p, const float* g, float* m, float* v, size_t n, float lr_t, float beta1, float beta2,
                  float eps, float grad_scale, void* w_hi, void* w_lo, const float* lr_ptr, const void* g_bf16,
                  void* stream);
/* Data-parallel option (new functionality, the reference has no collective): the flat gradient goes over NVLink as
 * bf16 (half the bytes of the one all-reduce per step).  fcn8_cast_bf16 writes out = bf16(x); fcn8_adam then reads the
 * reduced gradient from g_bf16 (bf16 [n]) instead of g (which may be NULL). */
int32_t fcn8_cast_bf16(const float* x, void* out, size_t n, void* stream);
/* Per-step scalars in device memory so that a captured CUDA graph of the whole training step can be replayed with a
 * new learning rate and dropout seed (the reference feeds both every step, fcn8s_tensorflow.py:558-562):
 * scalars[0] = lr_t (read by fcn8_adam through lr_ptr), scalars[1] = seed bits (read through Fcn8ConvParams.seed_ptr). */
int32_t fcn8_set_step_scalars(float* scalars, float lr_t, uint32_t seed, void* stream);
/* w_hi / w_lo (optional, bf16 [n]): the tensor-core shadow of the parameters, w_hi = bf16(p), w_lo = bf16(p - w_hi),
 * refreshed by fcn8_adam in the same pass; the GEMMs read it in TF layout (Fcn8ConvParams.w_mode 1 / 2), so no
 * per-step re-packing exists.  fcn8_shadow_weights builds it after a weight load. */
int32_t fcn8_shadow_weights(const float* p, void* w_hi, void* w_lo, size_t n, void* stream);
/* L2 regulariser of the six decoder kernels (:179..232, :250-251): g += rate*w;  loss_sum += 0.5*rate*sum w^2. */
int32_t fcn8_l2_reg(const float* w, float* g, float* loss_sum, size_t n, float rate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FCN8S_B200_H_ */
