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
  FCN8_EPI_COLSUM = 128,    /* colsum[co] += sum over pixels of the stored values: the BiasAddGrad of the layer whose
                               dY this dgrad call produces, fused into its epilogue (fp32 atomics, caller zeroes) */
  FCN8_EPI_POOL = 256       /* FCN8_BF16 only: also emit the 2x2 / stride-2 SAME max-pool of the output to pool_out
                               (the encoder's MaxPool fused into the epilogue of conv{1_2,2_2,3_3,4_3,5_3}); H and W
                               must be even; out may be NULL (inference: keep only the pooled tensor) */
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
 *   7 = c: experiment only -- multiply every tcgen05 accumulator by 1 + n_mma * c * 1e-10 (a statistical correction of
 *          TMEM's truncating accumulation; OFF by default: the loss is bounded structurally instead, key 9); 0 = 1: off
 *   9    : promoted accumulation of the error-compensated (3-product) convolutions: the hi*hi segment is accumulated
 *          in chunks of P k-blocks (4 MMAs each) that the epilogue warps add up in fp32 registers with round-to-
 *          nearest adds (csrc/conv_gemm.cuh, ConvGemmArgs::promo_kb).  0 = library default (12), P > 0 = chunk length,
 *          < 0 = off (one TMEM accumulation per tile; scripts/promo_sweep.py measures what that costs in accuracy)
 *   1 = 1: no wgrad_halo_kernel          2 = 1: no resident-weights variant of conv_halo_kernel<64>
 *   3    : bit 0: conv epilogues skip their output stores, bit 1: ... and their mask / residual loads (timing only)
 *   5 = 1: single-CTA kernels instead of the CTA-pair (cta_group::2) variants of conv_gemm / wgrad_gemm
 *   6 = 1: launch every kernel with the programmatic-dependent-launch attribute (kernels.h)
 *  11 = 1: single-CTA conv_halo_kernel instead of the CTA-pair variant for the 64- / 128-column halo tiles
 *  10 = 1: dynamic tile assignment in the persistent GEMM kernels (cluster launch control: one CTA / CTA pair per tile
 *          in the grid, running CTAs cancel pending ones and take their tiles) instead of the static round-robin over
 *          min(tiles, #SMs) CTAs; the engine switches it on when a gradient all-reduce runs under the backward pass */
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

/* ---- conv1_1 straight from the uint8 image: the feed (:558,686,765), the encoder graph's RGB->BGR / mean subtraction
 * and its first 3x3 convolution 3 -> 64 + bias + ReLU [EXT], and that convolution's filter gradient (:257), with the
 * 27-column im2col operand built in shared memory (csrc/conv1.cu).  w / w_lo: fcn8_pack_weights(mode 0, ksize 1,
 * Cin 27, CinPad 64) of conv1_1/filter seen as [27][64] (bf16, K-major [64 co][64 k]).  pair != 0: hi / lo operands and
 * a hi / lo output (out, out_lo with pixel stride out_ld); dy / dy_lo likewise.  N*H*W must be a multiple of 128. */
typedef struct {
  const uint8_t* images;  /* [N,H,W,3] RGB */
  int32_t N, H, W;
  const void* w;
  const void* w_lo;
  const float* bias;      /* [64] */
  void* out;              /* fwd: [N,H,W,out_ld] bf16 */
  void* out_lo;
  int32_t out_ld;         /* 0 = 64 */
  const void* dy;         /* wgrad: [N,H,W,dy_ld] bf16 */
  const void* dy_lo;
  int32_t dy_ld;          /* 0 = 64 */
  float* dw;              /* wgrad: [27][64] fp32 = conv1_1/filter in TF layout */
  int32_t pair;
} Fcn8Conv1Params;
int32_t fcn8_conv1_fwd(const Fcn8Conv1Params* p, void* stream);
size_t fcn8_conv1_wgrad_workspace_bytes(const Fcn8Conv1Params* p);
int32_t fcn8_conv1_wgrad(const Fcn8Conv1Params* p, void* workspace, size_t workspace_bytes, void* stream);

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
  float out_scale;      /* 0 = 1: the accumulator is multiplied by it before bias / residual (the 1e-4 / 1e-2 skip
                           scales of the score heads, fcn8s_tensorflow.py:171,182, and of their input gradients) */
  int32_t colsum_n;     /* 0 = Cout: only the first colsum_n columns are added to colsum */
  void* pool_out;       /* FCN8_EPI_POOL: [N, H/2, W/2, pool_ld] in the storage format of out (pool_out_lo: lo half) */
  void* pool_out_lo;
  int32_t pool_ld;      /* 0 = out_ld */
  /* strided views (elements; 0 = dense NHWC): row / image strides of x (and x_lo) and of out (and out_lo, mask_src,
   * residual) -- e.g. the interior of a zero-bordered padded tensor of the transposed-convolution stages */
  int64_t x_sH, x_sN, out_sH, out_sN;
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
  int32_t out_cols;  /* 0 = Cout: only the first out_cols columns of every row are written, dw is [rows][out_cols]
                        (score heads: Cout is the class count padded to 64) */
  float out_scale;   /* 0 = 1 */
  int64_t x_sH, x_sN, dy_sH, dy_sN;   /* strided views as in Fcn8ConvParams (0 = dense) */
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
  float* db;      /* bwd only, optional: db[c] += sum over pixels of dx (bias gradient of the producing conv); needs
                     C / 8 (C / 4 for FCN8_F32) to divide 256 -- every VGG width does; FCN8_ERR_UNSUPPORTED otherwise */
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

/* ---- decoder on the tensor cores: fcn8s_tensorflow.py:164-235 and its autodiff (:257), loss :253, predictor :268-269,
 * metrics :280-301.  All decoder activations are bf16 hi / lo PLANES: two bf16 tensors of the same geometry with
 * v ~ hi + lo; a pixel's channel vector is one 128-byte operand row (64 channels, classes zero-padded).
 *
 * 1x1 score heads (:171-200, tf.multiply scale + tf.layers.conv2d 1x1 + bias): fcn8_head_pack turns the kernel
 * K [1,1,Cin,C] into a TF-layout weight tensor with the class dimension zero-padded to 64, w[ci][c64] (bf16 hi / lo),
 * and bias64; then
 *   s   = fcn8_conv_gemm (ksize 1, Cout 64, w_mode 1, BIAS, out_scale = the skip scale, out / out_lo = the planes)
 *   dX  = fcn8_conv_gemm (x = ds planes, Cin 64, Cout = Cin_head, w_mode 2, out_scale = scale, MASK for fc7)
 *   dK  = fcn8_wgrad_gemm(x, dy = ds planes, Cout 64, out_cols = C, out_scale = scale)
 *   db  = the column sums the kernel that produced ds emitted (FCN8_EPI_COLSUM / Fcn8DeconvParams.colsum). */
int32_t fcn8_head_pack(const float* K, const float* bias, int32_t Cin, int32_t C, void* w_hi, void* w_lo, float* bias64,
                       void* stream);

/* Transposed convolutions (tf.layers.conv2d_transpose k = 2s, stride s, 'same'; kernel T [2s,2s,C_out,C_in] in TF layout)
 * as phase GEMMs on tcgen05: every s x s output block (J, I), J in [0,h], I in [0,w], is one GEMM row,
 *   Zblock[(dy,dx,co)] = sum_{(ty,tx,ci)} x[J-1+ty, I-1+tx, ci] * T[dy+s(1-ty), dx+s(1-tx), co, ci],
 * K = 4 taps x 64 channels, N = s*s*CP columns (CP = fcn8_deconv_cp(s): 64 for s = 2, 32 for s = 8).
 *   fcn8_deconv_pack  T, bias -> w_fwd [s*s*CP][256], w_dx [64][4*s*s*CP] (bf16 hi / lo), bias_big [s*s*CP]
 *   fcn8_deconv_fwd   (s = 2: upscore2 :204-211, upscore_pool4 :215-222) dense output planes [N, s*h, s*w, .]; the skip
 *                     tensor (tf.add :213 / :224) is added in the epilogue
 *   fcn8_deconv_loss  (s = 8: upscore8 :226-235) with the loss / predictor in the epilogue -- any subset of: loss_sum +=
 *                     sum over pixels of softmax-CE with `labels` (one-hot uint8 [N,8h,8w,C]); the loss gradient
 *                     dz = (softmax - labels) * grad_scale as bf16 planes in the padded blocked layout
 *                     [N, 8(h+1), 8(w+1), 32] (interior only: zero the planes once) and dbias[C] += its class sums;
 *                     logits / softmax fp32 [N,8h,8w,C]; argmax int64 (and / or uint8) [N,8h,8w];
 *                     conf[label*C + prediction] += 1
 *   fcn8_deconv_dx    input gradient from dz planes in the padded blocked layout [N, s(h+1), s(w+1), CP] (zero border,
 *                     zero beyond C) into out planes [N,h,w,64] (+ colsum[c] += column sums: the bias gradient of the
 *                     layer below)
 *   fcn8_deconv_dw    dT [2s,2s,C,C] fp32 from x planes and dz planes
 * nseg = 3: hi*hi + hi*lo + lo*hi; 1: hi planes only (the bf16 precision mode). */
typedef struct {
  const void* x;       /* input planes [N,h,w,.]: 64 readable channels per pixel, zero beyond C */
  const void* x_lo;
  int32_t x_ld;        /* pixel stride of x in elements */
  int64_t x_sH, x_sN;  /* row / image strides in elements (0 = dense) */
  const void* w;       /* packed operand: w_fwd (fwd, loss) or w_dx (dx) */
  const void* w_lo;
  const float* bias_big;
  void* out;           /* fwd: [N, s*h, s*w, .] planes; dx: [N,h,w,.] planes */
  void* out_lo;
  int32_t out_ld;
  int64_t out_sH, out_sN;
  const void* skip;    /* fwd: same geometry as out */
  const void* skip_lo;
  const void* dz;      /* dx, dw: padded blocked planes */
  const void* dz_lo;
  float* dT;           /* dw */
  float* colsum;       /* dx, optional */
  int32_t colsum_n;    /* 0 = C */
  int32_t N, h, wd, C, stride, nseg;   /* x is [N,h,wd,.] */
  /* fcn8_deconv_loss outputs / inputs (NULL = not requested) */
  const uint8_t* labels;
  float* loss_sum;
  float* dbias;
  void* dz_hi_out;
  void* dz_lo_out;
  float* logits;
  float* softmax;
  int64_t* argmax;
  uint64_t* conf;
  float grad_scale;
  uint8_t* argmax_u8;  /* the class map of `argmax` as one byte per pixel [N,8h,8w]: what FCN8s.predict() brings back over
                          PCIe (1/8 of the bytes of the int64 map tf.argmax returns, widened on the host) */
} Fcn8DeconvParams;
int32_t fcn8_deconv_cp(int32_t stride);
int32_t fcn8_deconv_pack(const float* T, const float* bias, int32_t C, int32_t stride, void* w_fwd, void* w_fwd_lo,
                         void* w_dx, void* w_dx_lo, float* bias_big, void* stream);
int32_t fcn8_deconv_fwd(const Fcn8DeconvParams* p, void* stream);
int32_t fcn8_deconv_loss(const Fcn8DeconvParams* p, void* stream);
int32_t fcn8_deconv_dx(const Fcn8DeconvParams* p, void* stream);
size_t fcn8_deconv_dw_workspace_bytes(const Fcn8DeconvParams* p);
int32_t fcn8_deconv_dw(const Fcn8DeconvParams* p, void* workspace, size_t workspace_bytes, void* stream);


/* ---- label feed: the `labels` placeholder (fcn8s_tensorflow.py:110, fed at :559,687) receives the generators' bool
 * one-hot batches [n,H,W,C] (helpers/ground_truth_conversion_utils.py:84-88, batch_generator_KITTI.py:82-84), which
 * TensorFlow widens to int32 on the host (4C bytes per pixel over PCIe).  Here the feed thread packs a batch to one
 * class id per pixel on the host and the device restores the one-hot tensor the loss kernel reads (1 byte per pixel).
 * fcn8_pack_labels (HOST code, `threads` worker threads, no CUDA call): returns 0 when every pixel's C bytes are
 * exactly one-hot (a single byte equal to 1) and ids[pixels] is filled; 1 when some pixel is not -- the caller then
 * ships the one-hot batch unchanged, so soft / multi-hot / all-zero label rows keep the reference's semantics;
 * < 0 on bad arguments.  fcn8_expand_labels: onehot[p*C + c] = (ids[p] == c), device pointers. */
int32_t fcn8_pack_labels(const uint8_t* onehot, int64_t pixels, int32_t C, uint8_t* ids, int32_t threads);
int32_t fcn8_expand_labels(const uint8_t* ids, uint8_t* onehot, int64_t pixels, int32_t C, void* stream);

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
