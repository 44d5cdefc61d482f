// Implicit-GEMM convolution kernels on tcgen05 (sm_100a) -- device side.
//
//   conv_gemm_kernel : stride-1 SAME k x k convolution of an NHWC tensor as a GEMM
//        D[pixels, Cout] = sum_{tap, cblock} A_tap[pixels, CH] * B[(tap, cblock), Cout]
//     used for fprop and for dgrad.  B is either a packed K-major operand [Cout][K] (conv1_1, decoder GEMMs) or the
//     bf16 shadow of the TF-layout (HWIO) weight tensor read in place (b_mode 1 / 2: rotation and in/out swap of the
//     dgrad live in the TMA coordinates).  Also carries the phase GEMMs of the transposed convolutions (blocked
//     output / 5-D blocked A operand).
//     Replaces TF's Conv2D / Conv2DBackpropInput behind the external VGG-16 graph the reference loads at
//     fcn8s_tensorflow.py:127-152 and differentiates at :256-257, and conv2d_transpose at :204-233.
//   conv_halo_kernel : the same 3x3 convolution with one activation patch per tile and nine shifted UMMA descriptors.
//   wgrad_gemm_kernel : D[(tap, ci), co] = sum_{pixels} X[pixel + tap, ci] * dY[pixel, co]
//     (TF's Conv2DBackpropFilter, implied by fcn8s_tensorflow.py:257), both operands MN-major.
//
//   PAIR variants of conv_gemm_kernel and wgrad_gemm_kernel (bf16 operands, 256-column tiles): clusters of two CTAs on
//     one SM pair, tcgen05 cta_group::2 -- each CTA owns a 128-row M tile and loads its own A tile plus HALF of the B
//     tile, the leader CTA issues M = 256 MMAs over both CTAs' shared memory (protocol: ptx.cuh, "CTA pairs";
//     measurements that led there: profiles/r01_probes.md).
//
// Shared design: one CTA = 10 warps: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (both loops run by
// the whole warp, one elected lane issues), warps 2..9 = epilogue (TMEM -> registers -> global; two warps per TMEM
// lane quarter).  Persistent over tiles, smem ring of kStages operand stages, 2 accumulator stages in TMEM so the
// epilogue of tile i overlaps the main loop of tile i+1.
// Everything is expressed in bytes: an operand row is always 128 B (64 bf16 or 32 tf32/fp32 values) with the
// 128-byte swizzle, so one code path serves kind::f16 (bf16) and kind::tf32.
#pragma once
#include <cuda_bf16.h>

#include <type_traits>

#include "kernels.h"
#include "ptx.cuh"

namespace fcn8 {

enum EpilogueFlags : int {
  EPI_BIAS = 1,      // + bias[col]
  EPI_RELU = 2,      // max(.,0)
  EPI_DROPOUT = 4,   // * keep(seed, index) / keep_prob   (forward dropout, fcn8s_tensorflow.py:561 feeds keep_prob)
  EPI_MASK = 8,      // * (mask_src[index] > 0) * mask_scale   (ReLU / dropout backward)
  EPI_RESIDUAL = 16, // + residual[index]                      (gradient fan-in at pool3 / pool4)
  EPI_PARTIAL = 32,  // raw fp32 partial sums to the split-K workspace, no epilogue math
  EPI_ROUND_TF32 = 64,  // fp32 outputs rounded to the nearest tf32 (single-pass tf32 mode: the MMA truncates operands)
  EPI_DBG_NOSTORE = 1 << 20,  // measurement only (fcn8_debug_set(3, 1)): the epilogue skips its output stores
  EPI_DBG_NOLOAD = 1 << 21,   // measurement only (fcn8_debug_set(3, 2)): ... and its mask / residual loads
  EPI_COLSUM = 128,     // colsum[col] += sum over rows of the final values (the bias gradient of the layer whose dY
                        // this dgrad produces): per-CTA shared-memory accumulation, one global atomic per column
  EPI_POOL = 256,       // also emit the 2x2 / stride-2 max-pool of the output (fused max-pool of conv{1_2,2_2,3_3,4_3,
                        // 5_3}): needs the 8 x 16 pixel tile (row = y*8 + x), so that a window is 4 lanes of one warp
};
constexpr int kColsumMax = 4096;  // widest dgrad output (fc6 activations)

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// Counter-based uniform in [0,1): one 32-bit mix of (seed, element index). Shared by the forward dropout epilogue,
// the split-K reduce epilogue and (host side, numpy) the parity tests that inject the same mask into the oracle.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t seed, uint64_t idx) {
  uint32_t x = static_cast<uint32_t>(idx) ^ (static_cast<uint32_t>(idx >> 32) * 0x9E3779B9u) ^ seed;
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  x += seed * 0x9E3779B9u;
  x ^= x >> 15;
  x *= 0x2c1b3c6du;
  x ^= x >> 12;
  return x;
}
__host__ __device__ __forceinline__ bool dropout_keep(uint32_t seed, uint64_t idx, uint32_t keep_threshold) {
  // keep iff the top 24 bits are below keep_prob * 2^24
  return (mix32(seed, idx) >> 8) < keep_threshold;
}

struct ConvGemmArgs {
  void* out;             // [N,H,W,ldc] OutT
  const float* bias;     // [ldc]
  const void* mask_src;  // same shape/type as out
  const void* residual;  // same shape/type as out
  float* partial;        // [splits][N*H*W*ldc] fp32 (EPI_PARTIAL)
  int N, H, W, ldc;
  int taps, taps_w, pad;
  int cblocks;  // Cin / CH
  int nseg;     // 1, or 3 for the error-compensated tf32 product (hi*hi + hi*lo + lo*hi)
  int lbw, lbh, lbn;
  int tiles_x, tiles_y, tiles_b, tiles_n;
  int splits, kb_per_split;
  int flags;
  float mask_scale;
  float inv_keep;
  uint32_t keep_threshold;
  uint32_t seed;
  // Output addressing: element offset of row (n, y, x) = n*osN + y*osH + x*osW; column c lands at +c (out_mode 0) or,
  // for the blocked output of a transposed convolution (out_mode 1: a row is one s x s output block whose columns are
  // (dy, dx, co)), at + (c / blk_row) * os_dy + c % blk_row.  store_cols > 0 keeps only the first store_cols columns.
  // out_mode 2: DENSE output of a transposed convolution (stride blk_s, blk_cp columns per pixel): row = block (y, x)
  // of the (h+1) x (w+1) block grid, column c = (q, ch), q = c / blk_cp = dy * blk_s + dx, lands at channel ch of pixel
  // (blk_s*y - blk_s/2 + dy, blk_s*x - blk_s/2 + dx) of the [N, out_H, out_W, .] tensor (strides osN / osH / osW);
  // pixels outside are dropped.  Together with EPI_RESIDUAL this is the skip-fusion add of fcn8s_tensorflow.py:213,224
  // in the epilogue of the transposed convolution that precedes it.
  long long osN, osH, osW, os_dy;
  int out_mode, blk_row, store_cols;
  int blk_s, blk_cp, out_H, out_W;
  int colsum_n;          // > 0: only the first colsum_n columns go to `colsum` (class counts < the padded width)
  // EPI_POOL: pooled output [N, H/2, W/2, .] (same storage format as `out`), element strides per image / row / pixel
  void* pool_out;
  void* pool_out_lo;
  long long psN, psH, psW;
  // A operand: a_mode 0 = NHWC map (C, W, H, N), tap (kh, kw) shifts the box by (kw - pad, kh - pad);
  // a_mode 1 = blocked 5-D map (s*CP, Wb, s, Hb, N) of a padded transposed-conv output: k-block (tap, cb) reads row
  // dy = cb / blk_chunks, chunk cb % blk_chunks of block (y + pad - kh, x + pad - kw).
  int a_mode, blk_chunks;
  // B operand: b_mode 0 = packed K-major 2-D map [Cout][K]; b_mode 1 / 2 read the TF-layout (HWIO) weight tensor
  // itself through a 3-D map (co, ci, tap): 1 = dgrad (K = (tap', co) K-major rows ci, tap = taps-1-tap' is the
  // 180-degree rotation), 2 = fprop (K = (tap, ci), N = co contiguous: MN-major chunks of [64 ci][64 co]).
  int b_mode;
  // bf16 hi/lo pair output (fp32-equivalent storage): out_lo != nullptr stores hi = bf16(v) to out, lo = bf16(v - hi)
  // to out_lo at the same element index; residual_lo is the low half of the residual.
  void* out_lo;
  const void* residual_lo;
  float* colsum;  // EPI_COLSUM target [tiles_n * BN] fp32, accumulated with atomics (caller zeroes)
  // Per-step dropout seed in device memory (CUDA-graph replays cannot change kernel arguments):
  // effective seed = *seed_ptr * 2 + seed when seed_ptr != nullptr.
  const uint32_t* seed_ptr;
  // Round-toward-zero compensation: the TMEM accumulator truncates on every tcgen05.mma, an expected relative loss of
  // kRzBiasPerMma per accumulation step given the final value (measured, scripts/bringup.py::rz_accumulation_probe).
  // acc_scale = 1 + (MMAs accumulated into one accumulator) * kRzBiasPerMma multiplies the accumulator in the epilogue.
  float acc_scale;   // base output scale (1, or the 1e-4 / 1e-2 skip scales of the score heads, fcn8s_tensorflow.py:171,182)
  // The low-order segments of an error-compensated product (x*w_lo, x_lo*w) are accumulated FIRST, while the
  // accumulator is still 2^-8 of its final size, and the hi*hi segment last: only its MMAs truncate at full magnitude.
  // rz_c = expected relative loss per k-block (4 MMAs) of the hi*hi segment; the epilogue scales a tile's accumulator
  // by acc_scale * (1 + rz_c * [hi*hi k-blocks accumulated into it]).
  float rz_c;
  // Promoted accumulation (PROMO kernels): the hi*hi segment of a tile is accumulated in CHUNKS of promo_kb k-blocks
  // (4 MMAs each), every chunk into a fresh TMEM accumulator stage; the epilogue warps drain each finished chunk into
  // an fp32 running sum in registers (round-to-nearest adds) while the next chunk accumulates into the other stage,
  // and write the total back to tensor memory before the tile's epilogue.  The truncating adds of the tensor core then
  // only ever see accumulators of <= 4 * promo_kb MMAs: the round-toward-zero loss is bounded by the chunk length
  // instead of growing with K (and with it the dependence of the bias on the sign structure of the data).
  int promo_kb;
  // Dynamic tile scheduling (cluster launch control): the grid holds one CTA (pair) per tile instead of one per SM; a
  // running CTA (pair) takes the tiles of clusters that have not been launched yet (TileSched below).  While another
  // kernel -- the overlapped NCCL all-reduce of the gradients -- holds some SMs, the resident CTAs absorb all the work
  // instead of leaving a statically assigned share to CTAs that cannot start (which doubled the GEMMs' time under the
  // collective).  0 = static round-robin over min(tiles, #SMs) CTAs.
  int dyn;
  // measurement only (fcn8_debug_buffer): when set, CTA b writes dbg[8b + 0..3] = cycles its MMA warp spent in the
  // tile loop / waiting for operand stages (full barriers) / waiting for a free accumulator, and k-blocks issued;
  // dbg[8b + 4] = cycles the TMA producer waited for free stages; [5] = (halo kernels) waiting for weight stages;
  // [6] / [7] = cycles the first epilogue warp waited for a finished accumulator / spent in its tile loop
  long long* dbg;
  // ---- EPI = 1 (loss / predictor epilogue of the upscore8 phase GEMM, fcn8s_tensorflow.py:226-235 + :253 / :268-269):
  // a 256-column tile is one row (dy = tile index) of 8 output pixels x 32 padded classes of each block, so every
  // epilogue thread holds whole pixels' logits.  Any subset of the outputs below may be requested (nullptr = skip).
  const uint8_t* labels;        // one-hot [N, out_H, out_W, C] (bool / uint8)
  float* loss_sum;              // += sum over pixels of softmax-CE
  float* dbias;                 // [C] += sum over pixels of dz (gradient of the transposed convolution's bias)
  __nv_bfloat16* dz_hi;         // dlogits in the padded blocked layout [N, 8*(h+1), 8*(w+1), 32], bf16 hi / lo planes
  __nv_bfloat16* dz_lo;         //   (nullptr: hi only), border and channels >= C are zero
  float* logits;                // dense [N, out_H, out_W, C]
  float* softmax;               // dense [N, out_H, out_W, C]
  long long* argmax;            // dense [N, out_H, out_W]
  unsigned char* argmax_u8;     // the same class map as one byte per pixel
  unsigned long long* conf;     // [C, C] confusion matrix, conf[label * C + prediction] += 1
  int num_classes;
  float gscale;                 // dz = (softmax - y) * gscale
};
constexpr float kRzBiasPerMma = 0.f;   // no statistical correction by default (promoted accumulation bounds the loss)

// Chunk structure of a promoted tile over k-blocks [kb0, kb1): a new chunk starts at every kb = kb_hi0 + j * P that
// lies strictly inside (max(kb0, kb_hi0), kb1) -- the low-order segments and the first P hi*hi k-blocks share chunk 0.
__device__ __forceinline__ bool promo_boundary(int kb, int kb0, int kb_hi0, int P) {
  return kb > kb0 && kb > kb_hi0 && (kb - kb_hi0) % P == 0;
}
__device__ __forceinline__ int promo_chunks(int kb0, int kb1, int kb_hi0, int P) {
  const int lo = max(kb0, kb_hi0);
  if (kb1 - 1 <= lo) return 1;
  return 1 + (kb1 - 1 - kb_hi0) / P - (lo - kb_hi0) / P;
}

struct TensorMaps3 {
  CUtensorMap a[3];
  CUtensorMap b[3];
};

// PAIR: the kernel runs as clusters of two CTAs on one SM pair (tcgen05 cta_group::2).  Each CTA owns one 128-row M
// tile and loads its own A tile but only HALF of the B tile; the leader issues M = 256 MMAs over both CTAs' shared
// memory.  Per k-block a CTA then streams 32 KB instead of 48 KB through TMA and shared memory, and six stages fit
// where four did -- the single-CTA kernel's main loop is bound by exactly that (wait-cycle profile: profiles/).
constexpr int kEpiWarpsC = 8;           // (= kEpiWarps, declared below)
constexpr int kStoreWarpBytes = 2048;   // 32 rows x 64 B: one epilogue warp's fragment of one bf16 plane
template <int BN, bool PAIR = false>
struct GemmCfg {
  static constexpr int kStages = PAIR ? 6 : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int kABytes = 128 * 128;  // 128 rows x 128 B
  static constexpr int kBRows = PAIR ? BN / 2 : BN;
  static constexpr int kBBytes = kBRows * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages (power of two >= 32 for BN in {64,128,256})
  static constexpr int kBarBytes = 1024;
  static constexpr int kColsumBytes = kColsumMax * 4;
  static constexpr int kStoreBytes = kEpiWarpsC * kStoreWarpBytes;   // per-warp staging of the coalesced epilogue stores
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kColsumBytes + kStoreBytes + 1024;  // +1024: manual alignment
};

constexpr int kGemmThreads = 320;  // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue
constexpr int kEpiWarps = 8;        // two epilogue warps per TMEM lane quarter, each takes half of the tile's columns
// PROMO kernels: three warpgroups -- warps 0..3 = TMA producer, MMA issuer and two idle warps (72 registers per thread
// after setmaxnreg.dec), warps 4..11 = epilogue (216 registers after setmaxnreg.inc: BN / 2 running sums per thread)
constexpr int kPromoThreads = 384;
constexpr int kPromoKbDefault = 12;   // 48 MMAs per chunk: worst-case (same-sign) truncation loss ~3e-6 per layer
constexpr int kPromoCtlRegs = 72, kPromoEpiRegs = 216;

// ------------------------------------------------------------------------------------------------------------------
// Tile scheduler shared by the three warp roles of a persistent kernel.  Static mode: tile += stride.  Dynamic mode
// (ConvGemmArgs::dyn): the TMA producer warp of the (leader) CTA issues, when it starts tile number `it`, the cluster
// launch control query for the tile after it; the answer is multicast into slot it % kSchedStages of every CTA of
// the pair; every role reads it (producer: after it has issued the tile's loads, the others: before they start the
// tile) and releases the slot on the leader's barrier.  One query is in flight ahead of the tile being loaded and none
// is issued after a failed one.
constexpr int kSchedStages = 4;
struct TileSched {
  uint32_t resp;    // shared::cta address of the answer slots (16 B each)
  uint64_t* full;   // [kSchedStages] 16 transaction bytes each
  uint64_t* empty;  // [kSchedStages] one arrival per consumer warp of the cluster (on the leader CTA)
  int dyn, stride, total;
};
template <bool PAIR>
__device__ __forceinline__ void sched_issue(const TileSched& s, int it, int lane) {
  const int slot = it % kSchedStages;
  mbar_wait(&s.empty[slot], ((static_cast<uint32_t>(it) / kSchedStages) & 1u) ^ 1u);
  if (lane < (PAIR ? 2 : 1)) {
    if constexpr (PAIR)
      mbar_expect_tx_cluster(map_to_cta(smem_u32(&s.full[slot]), lane), 16);
    else
      mbar_expect_tx(&s.full[slot], 16);
  }
  __syncwarp();
  if (lane == 0) clc_try_cancel<PAIR>(s.resp + slot * 16, smem_u32(&s.full[slot]));
  __syncwarp();
}
// the tile after tile number `it` of this CTA (pair): >= total when there is none
template <bool PAIR>
__device__ __forceinline__ int sched_next(const TileSched& s, int it, int t, int lane) {
  if (!s.dyn) return t + s.stride;
  const int slot = it % kSchedStages;
  mbar_wait(&s.full[slot], (static_cast<uint32_t>(it) / kSchedStages) & 1u);
  const int x = clc_decode(s.resp + slot * 16);
  fence_proxy_async_smem();   // this (generic-proxy) read is ordered before the next asynchronous write of the slot
  __syncwarp();
  if (lane == 0) {
    if constexpr (PAIR)
      mbar_arrive_cluster(map_to_cta(smem_u32(&s.empty[slot]), 0));
    else
      mbar_arrive(&s.empty[slot]);
  }
  return x < 0 ? s.total : (PAIR ? x >> 1 : x);
}

template <bool TF32>
__device__ __forceinline__ void umma_issue(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (TF32)
    umma_tf32(d, a, b, idesc, acc);
  else
    umma_f16(d, a, b, idesc, acc);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------------------------------
// Epilogue math on 32 consecutive columns of one output row. `idx` = element index of column c0 of this row.
// Sum of f[j] over the 32 lanes of the warp, delivered to lane j (transpose-reduce: 31 shuffles).
__device__ __forceinline__ float warp_colsum32(float (&f)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = up ? f[j] : f[j + s];
      const float keep = up ? f[j + s] : f[j];
      f[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return f[0];
}

// Coalesced store of a warp's bf16 fragment: lane l holds 64 bytes (32 values) of ITS row; written directly, one
// STG.128 of the warp touches 32 different 128-byte lines with 16 bytes each (ncu, round 2: the LSU queue backs up and
// the epilogue warps of the narrow-tile kernels are >95 % busy with it).  Instead the fragment goes through a per-warp
// 2 KB staging tile in shared memory (16-byte chunks XOR-swizzled by (row >> 1) & 3: both passes conflict-free) and
// leaves as four STG.128 in which four consecutive lanes cover one row's 64 contiguous bytes: 8 lines per instruction,
// whole 32-byte sectors only.  Row addresses travel by shuffle (dst = this lane's row, `ok` = row inside the tensor);
// every lane of the warp must call.
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void warp_store_frag64(uint8_t* stage, const uint32_t (&w)[16], const void* dst, bool ok,
                                                  int lane) {
  const uint32_t sbase = smem_u32(stage);
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    st_shared_v4(sbase + lane * 64 + ((j ^ sw) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
  __syncwarp();
  const unsigned long long mine = ok ? reinterpret_cast<unsigned long long>(dst) : 0ull;
  const int c = lane & 3;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = (lane >> 2) + 8 * it;
    const uint4 q = ld_shared_v4(sbase + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
    const unsigned long long rp = __shfl_sync(0xffffffffu, mine, r);
    if (rp) *reinterpret_cast<uint4*>(rp + c * 16) = q;
  }
  __syncwarp();
}

template <bool TF32>
__device__ __forceinline__ void epilogue_row32(const ConvGemmArgs& g, uint32_t (&v)[32], size_t idx, int c0,
                                               int ncols, size_t dense_idx, bool valid, float* colsum_s, int lane,
                                               uint32_t seed, float acc_scale, uint8_t* stage, size_t pool_idx = 0,
                                               bool pool_writer = false) {
  using OutT = typename std::conditional<TF32, float, __nv_bfloat16>::type;
  float f[32];
  if (acc_scale != 1.f) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * acc_scale;
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  }
  if (!valid) {
    // rows outside the tensor come along as zeros: the column sums, the pooling and the staged stores are
    // warp-collective
    if (TF32 && !(g.flags & EPI_COLSUM)) return;
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = 0.f;
  }
  if (valid && (g.flags & EPI_BIAS)) {
    const float4* b4 = reinterpret_cast<const float4*>(g.bias + c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = __ldg(b4 + i);
      f[4 * i + 0] += b.x;
      f[4 * i + 1] += b.y;
      f[4 * i + 2] += b.z;
      f[4 * i + 3] += b.w;
    }
  }
  if (valid && (g.flags & EPI_RESIDUAL) && !(g.flags & EPI_DBG_NOLOAD)) {
    const OutT* r = reinterpret_cast<const OutT*>(g.residual) + idx;
    if constexpr (TF32) {
      const float4* r4 = reinterpret_cast<const float4*>(r);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 b = __ldg(r4 + i);
        f[4 * i + 0] += b.x;
        f[4 * i + 1] += b.y;
        f[4 * i + 2] += b.z;
        f[4 * i + 3] += b.w;
      }
    } else {
      const uint4* r4 = reinterpret_cast<const uint4*>(r);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 q = __ldg(r4 + i);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 t = __bfloat1622float2(h[j]);
          f[8 * i + 2 * j] += t.x;
          f[8 * i + 2 * j + 1] += t.y;
        }
      }
      if (g.residual_lo) {
        const uint4* l4 = reinterpret_cast<const uint4*>(reinterpret_cast<const OutT*>(g.residual_lo) + idx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 q = __ldg(l4 + i);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 t = __bfloat1622float2(h[j]);
            f[8 * i + 2 * j] += t.x;
            f[8 * i + 2 * j + 1] += t.y;
          }
        }
      }
    }
  }
  if (g.flags & EPI_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
  }
  if (g.flags & EPI_DROPOUT) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      f[i] = dropout_keep(seed, static_cast<uint64_t>(dense_idx) + i, g.keep_threshold) ? f[i] * g.inv_keep : 0.f;
  }
  if (valid && (g.flags & EPI_MASK) && !(g.flags & EPI_DBG_NOLOAD)) {
    const OutT* m = reinterpret_cast<const OutT*>(g.mask_src) + idx;
    if constexpr (TF32) {
      const float4* m4 = reinterpret_cast<const float4*>(m);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 b = __ldg(m4 + i);
        f[4 * i + 0] = b.x > 0.f ? f[4 * i + 0] * g.mask_scale : 0.f;
        f[4 * i + 1] = b.y > 0.f ? f[4 * i + 1] * g.mask_scale : 0.f;
        f[4 * i + 2] = b.z > 0.f ? f[4 * i + 2] * g.mask_scale : 0.f;
        f[4 * i + 3] = b.w > 0.f ? f[4 * i + 3] * g.mask_scale : 0.f;
      }
    } else {
      const uint4* m4 = reinterpret_cast<const uint4*>(m);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 q = __ldg(m4 + i);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 t = __bfloat1622float2(h[j]);
          f[8 * i + 2 * j] = t.x > 0.f ? f[8 * i + 2 * j] * g.mask_scale : 0.f;
          f[8 * i + 2 * j + 1] = t.y > 0.f ? f[8 * i + 2 * j + 1] * g.mask_scale : 0.f;
        }
      }
    }
  }
  OutT* o = reinterpret_cast<OutT*>(g.out) + idx;
  if (g.flags & EPI_COLSUM) {
    float t[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) t[i] = valid ? f[i] : 0.f;
    const float cs = warp_colsum32(t, lane);
    atomicAdd(&colsum_s[c0 + lane], cs);
  }
  if (g.flags & EPI_DBG_NOSTORE) {
    float keep = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) keep += f[i];
    if (valid && keep == 123.456f) *reinterpret_cast<float*>(o) = keep;   // keeps the math alive
    return;
  }
  if constexpr (TF32) {
    if (!valid) return;
    if (g.flags & EPI_ROUND_TF32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = round_tf32(f[i]);
    }
    float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (4 * i < ncols) o4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
  } else {
    uint32_t hi[16], lo[16];
    // hi/lo pair storage format: also when only the pooled tensor is kept (inference: out == out_lo == nullptr)
    const bool pair_fmt = g.out_lo != nullptr || g.pool_out_lo != nullptr;
#pragma unroll
    for (int i = 0; i < 16; ++i) hi[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
    if (pair_fmt) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi[i]));
        lo[i] = pack_bf16x2(f[2 * i] - h.x, f[2 * i + 1] - h.y);
      }
    }
    if (g.out) {   // (out == nullptr: inference with a fused pool keeps only the pooled tensor)
      warp_store_frag64(stage, hi, o, valid, lane);
      if (g.out_lo) warp_store_frag64(stage, lo, reinterpret_cast<OutT*>(g.out_lo) + idx, valid, lane);
    }
    if ((g.flags & EPI_POOL) && !pair_fmt) {
      // bf16 storage: the 2x2 max of the stored values on the packed pairs (16 registers instead of 32 floats)
      uint32_t ph[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&hi[i]);
        uint32_t o = __shfl_xor_sync(0xffffffffu, hi[i], 1);
        a = __hmax2(a, *reinterpret_cast<const __nv_bfloat162*>(&o));
        o = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<const uint32_t*>(&a), 8);
        a = __hmax2(a, *reinterpret_cast<const __nv_bfloat162*>(&o));
        ph[i] = *reinterpret_cast<const uint32_t*>(&a);
      }
      if (pool_writer) {
        uint4* p4 = reinterpret_cast<uint4*>(reinterpret_cast<OutT*>(g.pool_out) + pool_idx);
#pragma unroll
        for (int i = 0; i < 4; ++i) p4[i] = make_uint4(ph[4 * i], ph[4 * i + 1], ph[4 * i + 2], ph[4 * i + 3]);
      }
    } else if (g.flags & EPI_POOL) {
      // 2x2 max over the STORED values (hi, or hi + lo for pairs: exactly what a stand-alone pool of the stored tensor
      // would see); the window's other three pixels are lanes ^1 (x) and ^8 (y) of this warp (tile row = y*8 + x)
      float m[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi[i]));
        if (pair_fmt) {
          const float2 l = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lo[i]));
          h.x += l.x;
          h.y += l.y;
        }
        m[2 * i] = h.x;
        m[2 * i + 1] = h.y;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], 1));
        m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], 8));
      }
      if (pool_writer) {
        uint32_t ph[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) ph[i] = pack_bf16x2(m[2 * i], m[2 * i + 1]);
        uint4* p4 = reinterpret_cast<uint4*>(reinterpret_cast<OutT*>(g.pool_out) + pool_idx);
#pragma unroll
        for (int i = 0; i < 4; ++i) p4[i] = make_uint4(ph[4 * i], ph[4 * i + 1], ph[4 * i + 2], ph[4 * i + 3]);
        if (g.pool_out_lo) {
          uint32_t pl[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ph[i]));
            pl[i] = pack_bf16x2(m[2 * i] - h.x, m[2 * i + 1] - h.y);
          }
          uint4* q4 = reinterpret_cast<uint4*>(reinterpret_cast<OutT*>(g.pool_out_lo) + pool_idx);
#pragma unroll
          for (int i = 0; i < 4; ++i) q4[i] = make_uint4(pl[4 * i], pl[4 * i + 1], pl[4 * i + 2], pl[4 * i + 3]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Loss / predictor epilogue of the upscore8 phase GEMM (EPI = 1): the thread's TMEM row is one 8 x 8 output block, the
// tile's 256 columns are row dy = `dy` of that block (8 pixels x 32 padded classes), this warp handles pixels
// dx = 4*chalf .. 4*chalf+3.  Per pixel: logits = acc * scale + bias  ->  softmax-CE (fcn8s_tensorflow.py:253), its
// gradient, softmax / argmax (:268-269), confusion-matrix update (:280-301) -- whichever outputs are requested.
// The loss gradient follows TensorFlow's fused op: backprop = softmax - labels (NOT softmax * sum(labels) - labels; the
// two agree for one-hot labels, which is all the reference's generators produce).
struct LossAcc {
  float loss;
  float db[32];
};
__device__ __forceinline__ void loss_epilogue_pixel(const ConvGemmArgs& g, const uint32_t (&v)[32], float acc_scale,
                                                    const float* bias, int n, int oy, int ox, size_t pad_idx,
                                                    LossAcc& acc, unsigned int* hist) {
  const int C = g.num_classes;
  float z[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) z[c] = (c < C) ? __uint_as_float(v[c]) * acc_scale + bias[c] : -INFINITY;
  const size_t p = (static_cast<size_t>(n) * g.out_H + oy) * g.out_W + ox;
  const bool vec4 = (C & 3) == 0;   // a pixel's C floats are then 16-byte aligned: 128-bit stores
  if (g.logits) {
    float* dst = g.logits + p * C;
    if (vec4) {
#pragma unroll
      for (int m = 0; m < 8; ++m)
        if (4 * m < C) reinterpret_cast<float4*>(dst)[m] = make_float4(z[4 * m], z[4 * m + 1], z[4 * m + 2], z[4 * m + 3]);
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < C) dst[c] = z[c];
    }
  }
  float mx = z[0];
  int am = 0;
#pragma unroll
  for (int c = 1; c < 32; ++c)
    if (c < C && z[c] > mx) {
      mx = z[c];
      am = c;
    }
  if (g.argmax) g.argmax[p] = am;
  if (g.argmax_u8) g.argmax_u8[p] = static_cast<unsigned char>(am);
  if (!(g.labels || g.softmax)) return;
  float e[32];
  float se = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    e[c] = (c < C) ? __expf(z[c] - mx) : 0.f;
    se += e[c];
  }
  const float inv = 1.f / se;
  if (g.softmax) {
    float* dst = g.softmax + p * C;
    if (vec4) {
#pragma unroll
      for (int m = 0; m < 8; ++m)
        if (4 * m < C)
          reinterpret_cast<float4*>(dst)[m] =
              make_float4(e[4 * m] * inv, e[4 * m + 1] * inv, e[4 * m + 2] * inv, e[4 * m + 3] * inv);
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < C) dst[c] = e[c] * inv;
    }
  }
  if (!g.labels) return;
  const uint8_t* lp = g.labels + p * C;
  float yv[32];
  if ((C & 3) == 0) {   // the pixel's C label bytes are 4-byte aligned
    const uint32_t* l4 = reinterpret_cast<const uint32_t*>(lp);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const uint32_t w = (4 * m < C) ? __ldg(l4 + m) : 0u;
      yv[4 * m] = static_cast<float>(w & 0xffu);
      yv[4 * m + 1] = static_cast<float>((w >> 8) & 0xffu);
      yv[4 * m + 2] = static_cast<float>((w >> 16) & 0xffu);
      yv[4 * m + 3] = static_cast<float>(w >> 24);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 32; ++c) yv[c] = (c < C) ? static_cast<float>(__ldg(lp + c)) : 0.f;
  }
  if (g.loss_sum) {
    const float lse = mx + __logf(se);
    float ysum = 0.f, yz = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (c < C) {
        ysum += yv[c];
        yz += yv[c] * z[c];
      }
    acc.loss += ysum * lse - yz;
  }
  if (g.conf) {   // labels_argmax (first maximum), fcn8s_tensorflow.py:280
    int gt = 0;
    float best = yv[0];
#pragma unroll
    for (int c = 1; c < 32; ++c)
      if (c < C && yv[c] > best) {
        best = yv[c];
        gt = c;
      }
    atomicAdd(&hist[gt * C + am], 1u);
  }
  if (g.dz_hi) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      const float d0 = (c < C) ? (e[c] * inv - yv[c]) * g.gscale : 0.f;
      const float d1 = (c + 1 < C) ? (e[c + 1] * inv - yv[c + 1]) * g.gscale : 0.f;
      acc.db[c] += d0;
      acc.db[c + 1] += d1;
      hi[c >> 1] = pack_bf16x2(d0, d1);
      const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi[c >> 1]));
      lo[c >> 1] = pack_bf16x2(d0 - h.x, d1 - h.y);
    }
    uint4* h4 = reinterpret_cast<uint4*>(g.dz_hi + pad_idx);
#pragma unroll
    for (int i = 0; i < 4; ++i) h4[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
    if (g.dz_lo) {
      uint4* l4 = reinterpret_cast<uint4*>(g.dz_lo + pad_idx);
#pragma unroll
      for (int i = 0; i < 4; ++i) l4[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Epilogue warps (8 warps: two per TMEM lane quarter, each takes half of the tile's columns) of the implicit-GEMM
// convolution kernels: persistent loop over the CTA's tiles, TMEM -> registers -> epilogue math -> global.
// PAIR (CTA pairs, see GemmCfg): `total_tiles` / `m_tiles` count PAIRS of M tiles, CTA `rank` of the pair owns M tile
// 2 * pair + rank (a partner past the last tile computes on zero-filled operands and stores nothing), and the
// accumulator is handed back on the leader CTA's barrier.
// EPI = 1: the loss / predictor epilogue above instead of the generic one (BN = 256 only); `colsum_s` then holds
// [0] the CTA's loss sum, [32..64) its class sums of dz, [64..64 + C*C) its confusion-matrix histogram.
// ALT (narrow tiles, BN <= 128): the two warps of a lane quarter take ALTERNATE tiles (all BN columns each) instead
// of half of every tile's columns.  A tile's epilogue is one dependent chain per warp (tcgen05.ld -> math -> staged
// stores, 2000+ cycles for 32 columns) that cannot overlap with itself; with short main loops (64-column tiles: 36
// MMAs) that latency, not the issue rate, bounded the kernel (accumulator waits of 40-57 % in the bf16 mode).  Two
// groups working on consecutive tiles hide it; accumulator stage = tile parity = group.
template <int BN, bool TF32, bool PAIR = false, int EPI = 0, bool PROMO = false, bool ALT = false>
__device__ __forceinline__ void conv_epilogue_loop(const ConvGemmArgs& g, uint32_t tmem_base, uint64_t* acc_full,
                                                   uint64_t* acc_empty, float* colsum_s, uint8_t* store_s,
                                                   int total_tiles, int m_tiles, int warp, int lane,
                                                   uint32_t rank = 0, const TileSched* sched = nullptr) {
    static_assert(!ALT || (!PROMO && (EPI == 1 || BN <= 128)), "alternating epilogue groups: narrow tiles, loss epilogue");
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int chalf = (warp - (PROMO ? 4 : 2)) >> 2;  // which half of the tile's columns (ALT: which tiles) this warp handles
    constexpr int kColsPerWarp = ALT ? BN : BN / 2;
    const int col_begin = ALT ? 0 : chalf * (BN / 2);
    uint8_t* stage = store_s + (warp - (PROMO ? 4 : 2)) * kStoreWarpBytes;   // this warp's store staging tile
    const int row = quarter * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    const size_t out_elems = static_cast<size_t>(g.N) * g.H * g.W * g.ldc;
    const uint32_t seed = g.seed_ptr ? (__ldg(g.seed_ptr) * 2u + g.seed) : g.seed;
    const uint32_t lead_acc_empty = PAIR ? map_to_cta(smem_u32(acc_empty), 0) : 0;
    const int t_step = PAIR ? gridDim.x >> 1 : gridDim.x;
    const int kb_per_seg = g.taps * g.cblocks;
    const int total_kb = g.nseg * kb_per_seg;
    const int kb_hi0 = total_kb - kb_per_seg;   // the hi*hi segment is the last one
    long long dbg_wait = 0;   // measurement only: cycles this warp waited for a finished accumulator
    const long long dbg_e0 = g.dbg ? clock64() : 0;
    LossAcc lacc;
    if constexpr (EPI == 1) {
      lacc.loss = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) lacc.db[c] = 0.f;
    }
    int t_next = 0;
    for (int t = PAIR ? blockIdx.x >> 1 : blockIdx.x, it = 0; t < total_tiles; t = t_next, ++it) {
      t_next = sched ? sched_next<PAIR>(*sched, it, t, lane) : t + t_step;
      if constexpr (ALT) {
        if ((it & 1) != chalf) continue;   // the other group's tile
        as = chalf;
        aphase = (static_cast<uint32_t>(it) >> 1) & 1u;
      }
      const int nb = t % g.tiles_n;
      const int mt = PAIR ? 2 * ((t / g.tiles_n) % m_tiles) + static_cast<int>(rank) : (t / g.tiles_n) % m_tiles;
      const int sp = t / (g.tiles_n * m_tiles);
      const int tx = mt % g.tiles_x;
      const int ty = (mt / g.tiles_x) % g.tiles_y;
      const int tb = mt / (g.tiles_x * g.tiles_y);
      const int x = (tx << g.lbw) + (row & ((1 << g.lbw) - 1));
      const int y = (ty << g.lbh) + ((row >> g.lbw) & ((1 << g.lbh) - 1));
      const int n = (tb << g.lbn) + (row >> (g.lbw + g.lbh));
      const bool valid = (x < g.W) && (y < g.H) && (n < g.N);
      const size_t pix = (static_cast<size_t>(n) * g.H + y) * g.W + x;
      const size_t row_off = static_cast<size_t>(n) * g.osN + static_cast<size_t>(y) * g.osH +
                             static_cast<size_t>(x) * g.osW;
      const int kb0 = sp * g.kb_per_split;
      const int kb1 = min(total_kb, kb0 + g.kb_per_split);
      const float acc_scale = g.acc_scale * (1.f + g.rz_c * static_cast<float>(max(0, kb1 - max(kb0, kb_hi0))));
      if constexpr (PROMO) {
        // every chunk but the last: drain the finished accumulator stage into the running sum, hand the stage back
        const int nch = promo_chunks(kb0, kb1, kb_hi0, g.promo_kb);
        if (nch > 1) {
          float sums[BN / 2];
#pragma unroll
          for (int i = 0; i < BN / 2; ++i) sums[i] = 0.f;
#pragma unroll 1
          for (int ch = 0; ch + 1 < nch; ++ch) {
            mbar_wait(&acc_full[as], aphase);
            tc_fence_after();
            const uint32_t ta = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16) + chalf * (BN / 2);
#pragma unroll
            for (int q = 0; q < BN / 64; ++q) {
              uint32_t v[32];
              tmem_ld32(ta + q * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) sums[q * 32 + i] += __uint_as_float(v[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR)
                mbar_arrive_cluster(lead_acc_empty + as * 8);
              else
                mbar_arrive_relaxed(&acc_empty[as]);
            }
            if (++as == 2) {
              as = 0;
              aphase ^= 1;
            }
          }
          // last chunk: total = running sum + this accumulator, written back in place (this warp's own lanes / columns)
          mbar_wait(&acc_full[as], aphase);
          tc_fence_after();
          const uint32_t ta = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16) + chalf * (BN / 2);
#pragma unroll
          for (int q = 0; q < BN / 64; ++q) {
            uint32_t v[32];
            tmem_ld32(ta + q * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(sums[q * 32 + i] + __uint_as_float(v[i]));
            tmem_st32(ta + q * 32, v);
          }
          tmem_st_wait();
        }
      }
      const long long dbg_tw = g.dbg ? clock64() : 0;
      mbar_wait(&acc_full[as], aphase);
      if (g.dbg) dbg_wait += clock64() - dbg_tw;
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c = col_begin; c < col_begin + kColsPerWarp; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        if (c + 32 >= col_begin + kColsPerWarp) {
          // the warp's last chunk is in registers: hand the accumulator stage back BEFORE the math and the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR)
              mbar_arrive_cluster(lead_acc_empty + as * 8);
            else
              mbar_arrive_relaxed(&acc_empty[as]);
          }
        }
        const int c0 = nb * BN + c;
        if constexpr (EPI == 1) {
          // block (y, x) of image n, block row dy = nb, pixel dx = c / 32 of the row
          const int dx = c >> 5;
          const int oy = 8 * y - 4 + nb, ox = 8 * x - 4 + dx;
          if (valid && oy >= 0 && oy < g.out_H && ox >= 0 && ox < g.out_W) {
            const size_t pad_idx = ((static_cast<size_t>(n) * (8 * g.H) + 8 * y + nb) * (8 * g.W) + 8 * x + dx) * 32;
            loss_epilogue_pixel(g, v, acc_scale, g.bias, n, oy, ox, pad_idx, lacc,
                                reinterpret_cast<unsigned int*>(colsum_s) + 64);
          }
        } else if (g.flags & EPI_PARTIAL) {
          if (valid) {
            const size_t idx = pix * g.ldc + c0;
            float4* o4 = reinterpret_cast<float4*>(g.partial + static_cast<size_t>(sp) * out_elems + idx);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              o4[i] = make_float4(__uint_as_float(v[4 * i]) * acc_scale, __uint_as_float(v[4 * i + 1]) * acc_scale,
                                  __uint_as_float(v[4 * i + 2]) * acc_scale,
                                  __uint_as_float(v[4 * i + 3]) * acc_scale);
          }
        } else {
          size_t idx = row_off + c0;
          bool vrow = valid;
          if (g.out_mode == 1) {
            const int bdy = c0 / g.blk_row;
            idx = row_off + static_cast<size_t>(bdy) * g.os_dy + (c0 - bdy * g.blk_row);
          } else if (g.out_mode == 2) {
            const int q = c0 / g.blk_cp;
            const int bdy = q / g.blk_s, bdx = q - bdy * g.blk_s;
            const int oy = g.blk_s * y - (g.blk_s >> 1) + bdy, ox = g.blk_s * x - (g.blk_s >> 1) + bdx;
            vrow = valid && oy >= 0 && oy < g.out_H && ox >= 0 && ox < g.out_W;
            idx = static_cast<size_t>(n) * g.osN + static_cast<size_t>(oy) * g.osH + static_cast<size_t>(ox) * g.osW +
                  (c0 - q * g.blk_cp);
          }
          const int ncols = g.store_cols > 0 ? min(32, g.store_cols - c0) : 32;
          size_t pool_idx = 0;
          bool pool_writer = false;
          if (g.flags & EPI_POOL) {
            pool_writer = valid && !(lane & 9);   // even x, even y of the 8 x 16 tile: the window's top-left pixel
            pool_idx = static_cast<size_t>(n) * g.psN + static_cast<size_t>(y >> 1) * g.psH +
                       static_cast<size_t>(x >> 1) * g.psW + c0;
          }
          if (ncols > 0)
            epilogue_row32<TF32>(g, v, idx, c0, ncols, pix * g.ldc + c0, vrow, colsum_s, lane, seed, acc_scale, stage,
                                 pool_idx, pool_writer);
        }
      }
      if constexpr (!ALT) {
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
    if (g.dbg && lane == 0 && warp == (PROMO ? 4 : 2) && blockIdx.x < 148) {   // first epilogue warp: [6] = idle cycles, [7] = loop cycles
      g.dbg[8 * blockIdx.x + 6] = dbg_wait;
      g.dbg[8 * blockIdx.x + 7] = clock64() - dbg_e0;
    }
    if constexpr (EPI == 1) {
      // CTA-level reduction of the loss and the class sums of dz (flushed to global memory by the kernel's tail)
      if (g.loss_sum) {
        float l = lacc.loss;
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        if (lane == 0) atomicAdd(&colsum_s[0], l);
      }
      if (g.dz_hi && g.dbias) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float a = lacc.db[c];
          for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
          if (lane == 0 && c < g.num_classes) atomicAdd(&colsum_s[32 + c], a);
        }
      }
    }
  }

// ------------------------------------------------------------------------------------------------------------------
template <int BN, bool TF32, bool PAIR = false, int EPI = 0, bool PROMO = false>
__global__ void __launch_bounds__(PROMO ? kPromoThreads : kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ TensorMaps3 maps, const ConvGemmArgs g) {
  pdl_launch_dependents();
  static_assert(!PAIR || (!TF32 && BN == 256), "CTA pairs: bf16 operands, 256-column tiles");
  static_assert(EPI == 0 || (!TF32 && BN == 256), "loss epilogue: bf16 operands, one block row per 256-column tile");
  static_assert(!PROMO || (EPI == 0 && !TF32), "promoted accumulation: bf16 operands, generic epilogue");
  constexpr int kThreads = PROMO ? kPromoThreads : kGemmThreads;
  constexpr int kEpiWarp0 = PROMO ? 4 : 2;
  // narrow single-CTA tiles (score heads, decoder GEMMs, small images): alternating epilogue groups (conv_epilogue_loop)
  // ... and the loss / predictor epilogue of upscore8 (~300 dependent instructions per pixel, 12-k-block main loops)
  constexpr bool kAlt = (BN <= 128 && !TF32 && !PAIR && !PROMO && EPI == 0) || EPI == 1;
  using Cfg = GemmCfg<BN, PAIR>;
  constexpr int CH = TF32 ? 32 : 64;  // elements per 128-byte operand row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* acc_full = empty_bar + Cfg::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* colsum_s = reinterpret_cast<float*>(bar_base + Cfg::kBarBytes);
  uint8_t* store_s = bar_base + Cfg::kBarBytes + Cfg::kColsumBytes;
  // tile scheduler (second half of the barrier block): answer slots, then their barriers
  TileSched sched;
  sched.resp = smem_u32(bar_base + 512);
  sched.full = reinterpret_cast<uint64_t*>(bar_base + 512 + 16 * kSchedStages);
  sched.empty = sched.full + kSchedStages;
  sched.dyn = g.dyn;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (EPI == 1) {
    for (int c = threadIdx.x; c < 64 + 32 * 32; c += kThreads) colsum_s[c] = 0.f;   // loss, dz class sums, histogram
  } else if (g.flags & EPI_COLSUM) {
    for (int c = threadIdx.x; c < g.tiles_n * BN; c += kThreads) colsum_s[c] = 0.f;
  }

  // scheduling units: M tiles, or (PAIR) pairs of consecutive M tiles, one per CTA of the cluster
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  const int real_m_tiles = g.tiles_x * g.tiles_y * g.tiles_b;
  const int m_tiles = PAIR ? (real_m_tiles + 1) / 2 : real_m_tiles;
  const int total_tiles = m_tiles * g.tiles_n * g.splits;
  const int t_first = PAIR ? blockIdx.x >> 1 : blockIdx.x;
  const int t_step = PAIR ? gridDim.x >> 1 : gridDim.x;
  const int kb_per_seg = g.taps * g.cblocks;
  const int total_kb = g.nseg * kb_per_seg;
  sched.stride = t_step;
  sched.total = total_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nseg; ++s) {
      tma_prefetch_desc(&maps.a[s]);
      tma_prefetch_desc(&maps.b[s]);
    }
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], PAIR ? 2 : 1);   // pair: the leader's expect_tx arrive + the partner's plain arrive
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], (PAIR ? 2 : 1) * (kAlt ? kEpiWarps / 2 : kEpiWarps));
    }
    for (int s = 0; s < kSchedStages; ++s) {
      mbar_init(&sched.full[s], 1);
      // consumers: producer + MMA + epilogue warps of the leader, producer + epilogue warps of the partner
      mbar_init(&sched.empty[s], PAIR ? 2 * kEpiWarps + 3 : kEpiWarps + 2);
    }
    fence_mbar_init();
  }
  if (PAIR) cluster_sync_all();   // both CTAs' barriers exist before anything is signalled across
  if (warp == 1) {
    if (PAIR)
      tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
    else
      tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched from here on

  // (PROMO: the register re-partitioning sits INSIDE the two role branches, so that the compiler allocates the control
  // warps' code against 72 registers and the epilogue's against 216)
  if (warp < kEpiWarp0) {
  if constexpr (PROMO) setmaxnreg_dec<kPromoCtlRegs>();
  if (warp == 0) {
    // ============================== TMA producer (whole warp, one elected lane issues) ==============================
    {
      int stage = 0;
      uint32_t phase = 0;
      long long dbg_wait_empty = 0;
      for (int t = t_first, it = 0; t < total_tiles; ++it) {
        if (g.dyn && rank == 0) sched_issue<PAIR>(sched, it, lane);   // ask for the tile after this one
        const int nb = t % g.tiles_n;
        const int mt = PAIR ? 2 * ((t / g.tiles_n) % m_tiles) + static_cast<int>(rank) : (t / g.tiles_n) % m_tiles;
        const int sp = t / (g.tiles_n * m_tiles);
        const int tx = mt % g.tiles_x;
        const int ty = (mt / g.tiles_x) % g.tiles_y;
        const int tb = mt / (g.tiles_x * g.tiles_y);
        const int x0 = tx << g.lbw, y0 = ty << g.lbh, n0 = tb << g.lbn;
        const int kb0 = sp * g.kb_per_split;
        const int kb1 = min(total_kb, kb0 + g.kb_per_split);
        int seg = kb0 / kb_per_seg;
        int rem = kb0 - seg * kb_per_seg;
        int tap = rem / g.cblocks;
        int cb = rem - tap * g.cblocks;
        int kh = tap / g.taps_w;
        int kw = tap - kh * g.taps_w;
        for (int kb = kb0; kb < kb1; ++kb) {
          const long long tw = g.dbg ? clock64() : 0;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (g.dbg) dbg_wait_empty += clock64() - tw;
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (PAIR) {
            // operands land in this CTA's shared memory, their bytes are counted on the leader's barrier; an M tile
            // past the end (odd tile count) reads rows outside the tensor: zero-filled by TMA
            if (elect_one()) {
              const uint32_t lead_bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
              if (rank == 0)
                mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
              else
                mbar_arrive_cluster(lead_bar);
              tma_load_4d_pair(&maps.a[seg], lead_bar, sa, cb * CH, x0 + kw - g.pad, y0 + kh - g.pad, n0);
              const int nrow = nb * BN + static_cast<int>(rank) * (BN / 2);
              if (g.b_mode == 0) {
                tma_load_2d_pair(&maps.b[seg], lead_bar, sb, (tap * g.cblocks + cb) * CH, nrow);
              } else if (g.b_mode == 1) {
                tma_load_3d_pair(&maps.b[seg], lead_bar, sb, cb * CH, nrow, g.taps - 1 - tap);
              } else {
#pragma unroll
                for (int j = 0; j < BN / 128; ++j)
                  tma_load_3d_pair(&maps.b[seg], lead_bar, sb + j * 8192, nrow + j * 64, cb * CH, tap);
              }
            }
          } else if (elect_one()) {
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            if (g.a_mode == 0) {
              tma_load_4d(&maps.a[seg], &full_bar[stage], sa, cb * CH, x0 + kw - g.pad, y0 + kh - g.pad, n0);
            } else {
              const int bdy = cb / g.blk_chunks;
              tma_load_5d(&maps.a[seg], &full_bar[stage], sa, (cb - bdy * g.blk_chunks) * CH, x0 + g.pad - kw, bdy,
                          y0 + g.pad - kh, n0);
            }
            if (g.b_mode == 0) {
              tma_load_2d(&maps.b[seg], &full_bar[stage], sb, (tap * g.cblocks + cb) * CH, nb * BN);
            } else if (g.b_mode == 1) {
              tma_load_3d(&maps.b[seg], &full_bar[stage], sb, cb * CH, nb * BN, g.taps - 1 - tap);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_3d(&maps.b[seg], &full_bar[stage], sb + j * 8192, nb * BN + j * 64, cb * CH, tap);
            }
          }
          __syncwarp();
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
          if (++cb == g.cblocks) {
            cb = 0;
            ++tap;
            if (++kw == g.taps_w) {
              kw = 0;
              ++kh;
            }
            if (tap == g.taps) {
              tap = 0;
              kh = 0;
              kw = 0;
              ++seg;
            }
          }
        }
        t = sched_next<PAIR>(sched, it, t, lane);
      }
      if (g.dbg && lane == 0 && blockIdx.x < 148) g.dbg[8 * blockIdx.x + 4] = dbg_wait_empty;
    }
  } else if (warp == 1 && rank == 0) {
    // ============================== MMA issuer (the leader CTA of a pair issues for both) ==============================
    // The WHOLE warp runs this loop: its control flow and every address / descriptor are then warp-uniform and live in
    // uniform registers, and one elected lane issues the tcgen05 instructions.  (With the loop inside `if (lane == 0)`
    // the compiler wraps every UTCHMMA operand in ELECT + R2UR.BROADCAST: ~600 cycles of issue overhead per k-block,
    // measured with ncu's source view, which capped the tensor pipe at ~65 % for N = 256 and ~20 % for N = 64 tiles.)
    const bool b_mn = g.b_mode == 2;  // B = [64 k][64 n] MN-major chunks (bf16 only)
    const uint32_t idesc = make_idesc(TF32 ? 2u : 1u, 0u, b_mn ? 1u : 0u, PAIR ? 256u : 128u, BN);
    // K-major B: 128-byte rows, +32 B per MMA; MN-major B: LBO = 8 KB between 64-wide N chunks, SBO = 1 KB between
    // 8-row K groups, +16 K-rows (2 KB) per MMA
    const uint64_t adesc0 = make_smem_desc_sw128(0, 16, 1024);
    const uint64_t bdesc0 = b_mn ? make_smem_desc_sw128(0, 8192, 1024) : make_smem_desc_sw128(0, 16, 1024);
    const uint32_t badv = b_mn ? 128u : 2u;
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    long long dbg_full = 0, dbg_acc = 0, dbg_kb = 0;
    const long long dbg_t0 = g.dbg ? clock64() : 0;
    int t_next = 0;
    for (int t = t_first, it = 0; t < total_tiles; t = t_next, ++it) {
      t_next = sched_next<PAIR>(sched, it, t, lane);
      const int sp = t / (g.tiles_n * m_tiles);
      const int kb0 = sp * g.kb_per_split;
      const int kb1 = min(total_kb, kb0 + g.kb_per_split);
      long long tw = g.dbg ? clock64() : 0;
      mbar_wait(&acc_empty[as], aphase ^ 1);
      if (g.dbg) dbg_acc += clock64() - tw;
      tc_fence_after();
      uint32_t d_tmem = tmem_base + as * BN;
      int kb_first = kb0;   // first k-block of the current accumulation chunk (PROMO: chunks, else the whole tile)
      for (int kb = kb0; kb < kb1; ++kb) {
        if constexpr (PROMO) {
          if (promo_boundary(kb, kb0, total_kb - kb_per_seg, g.promo_kb)) {
            // chunk complete: publish this accumulator stage, continue in the other one from zero
            if (elect_one()) {
              if constexpr (PAIR)
                umma_commit_pair(&acc_full[as], 3);
              else
                umma_commit(&acc_full[as]);
            }
            __syncwarp();
            if (++as == 2) {
              as = 0;
              aphase ^= 1;
            }
            if (g.dbg) tw = clock64();
            mbar_wait(&acc_empty[as], aphase ^ 1);
            if (g.dbg) dbg_acc += clock64() - tw;
            tc_fence_after();
            d_tmem = tmem_base + as * BN;
            kb_first = kb;
          }
        }
        if (g.dbg) tw = clock64();
        mbar_wait(&full_bar[stage], phase);
        if (g.dbg) {
          dbg_full += clock64() - tw;
          ++dbg_kb;
        }
        tc_fence_after();
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint64_t adesc = adesc0 | static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
        const uint64_t bdesc = bdesc0 | static_cast<uint64_t>(((sa + Cfg::kABytes) & 0x3FFFF) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // A: advance 32 bytes along K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
            if constexpr (PAIR)
              umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + badv * k, idesc, (kb > kb_first || k > 0) ? 1u : 0u);
            else
              umma_issue<TF32>(d_tmem, adesc + 2 * k, bdesc + badv * k, idesc, (kb > kb_first || k > 0) ? 1u : 0u);
          }
          if constexpr (PAIR)
            umma_commit_pair(&empty_bar[stage], 3);   // frees the stage in both CTAs
          else
            umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) {
        if constexpr (PAIR)
          umma_commit_pair(&acc_full[as], 3);   // the accumulator halves of both CTAs are complete
        else
          umma_commit(&acc_full[as]);
      }
      __syncwarp();
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (g.dbg && lane == 0 && blockIdx.x < 148) {
      g.dbg[8 * blockIdx.x + 0] = clock64() - dbg_t0;
      g.dbg[8 * blockIdx.x + 1] = dbg_full;
      g.dbg[8 * blockIdx.x + 2] = dbg_acc;
      g.dbg[8 * blockIdx.x + 3] = dbg_kb;
    }
  }
  } else {
    // ============================== epilogue ==============================
    if constexpr (PROMO) setmaxnreg_inc<kPromoEpiRegs>();
    conv_epilogue_loop<BN, TF32, PAIR, EPI, PROMO, kAlt>(g, tmem_base, acc_full, acc_empty, colsum_s, store_s,
                                                         total_tiles, m_tiles, warp, lane, rank, &sched);
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // nobody leaves while the partner may still read its operands or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    else
      tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
  if (EPI == 1) {
    if (threadIdx.x == 0 && g.loss_sum) atomicAdd(g.loss_sum, colsum_s[0]);
    if (threadIdx.x < g.num_classes && g.dz_hi && g.dbias) atomicAdd(g.dbias + threadIdx.x, colsum_s[32 + threadIdx.x]);
    if (g.conf) {
      const unsigned int* hist = reinterpret_cast<const unsigned int*>(colsum_s) + 64;
      for (int c = threadIdx.x; c < g.num_classes * g.num_classes; c += kThreads)
        if (hist[c]) atomicAdd(g.conf + c, static_cast<unsigned long long>(hist[c]));
    }
  } else if ((g.flags & EPI_COLSUM) && !(g.flags & EPI_PARTIAL)) {
    const int ncs = g.colsum_n > 0 ? min(g.colsum_n, g.tiles_n * BN) : g.tiles_n * BN;
    for (int c = threadIdx.x; c < ncs; c += kThreads) {
      const float t = colsum_s[c];
      if (t != 0.f) atomicAdd(g.colsum + c, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// conv_halo_kernel: the same 3x3 stride-1 SAME convolution for the layers whose implicit GEMM is bound by L2 -> shared
// memory traffic (few channels: conv1_x, conv2_x): the nine taps of a tile read nine shifted windows of the SAME
// activation patch, so the patch is loaded ONCE with its one-pixel halo -- a TMA box of 18 rows x 16 pixels x 128 B
// (64 bf16 channels) for a tile of 16 rows x 8 pixels -- and every tap's A operand is a UMMA descriptor that starts
// (kh*16 + kw) rows into that box: the 8-row core-matrix groups of the 128 output pixels are the patch rows, 2048 B
// (= 16 pixels) apart (SBO), so groups and the 128-byte swizzle phase of every row stay address-consistent.
// A traffic per tile drops from 9 x 16 KB to 36 KB; B (weights, TF layout as in conv_gemm_kernel) is streamed through
// its own ring.  bf16 operands only (kind::f16); epilogue shared with conv_gemm_kernel.
struct HaloCfg {
  static constexpr int kABytes = 18 * 16 * 128;  // 36,864
};
// The activation patches go through a ring of 3-4 stages: a 36 KB box of 288 rows has a TMA latency of ~3000 cycles,
// longer than the 36 MMAs that consume it, so with two stages the kernel was latency-bound (tensor pipe 43-57 %).
// RB ("resident B", N = 64 only): when the CTA's whole weight slice is 9 taps x 8 KB (conv1_2 forward / dgrad in bf16)
// it is loaded into shared memory ONCE and every tile only streams its activation patch.
template <int BN, bool RB = false>
struct HaloSmem {
  // filter taps per B pipeline stage: the three kw taps of one kh row for the narrow N = 64 tiles (12 MMAs per barrier
  // round trip instead of 4: the MMA time of a single 128x64x64 stage is shorter than the issue + wait overhead)
  static constexpr int kTaps = (BN == 64) ? 3 : 1;
  static constexpr int kBBytes = kTaps * BN * 128;
  static constexpr int kAStages = 3;
  static constexpr int kBStages = RB ? 3 : (BN == 128 ? 6 : (BN == 64 ? 4 : 3));   // RB: the 3 resident tap-row groups
  static constexpr int kBarBytes = 1024;
  static constexpr int kColsumBytes = 1024;   // the halo layers have at most 256 output channels
  static constexpr int kStoreBytes = kEpiWarpsC * kStoreWarpBytes;   // coalesced epilogue stores (warp_store_frag64)
  static constexpr int kBytes = kAStages * HaloCfg::kABytes + kBStages * kBBytes + kBarBytes + kColsumBytes +
                                kStoreBytes + 1024;
  static_assert(kBytes <= 232448, "shared memory of the halo kernel");
};

template <int BN, bool RB = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_halo_kernel(const __grid_constant__ TensorMaps3 maps, const ConvGemmArgs g) {
  pdl_launch_dependents();
  static_assert(!RB || BN == 64, "resident weights only for the N = 64 tiles");
  using HS = HaloSmem<BN, RB>;
  constexpr int kAS = HS::kAStages, kBS = HS::kBStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + kAS * HaloCfg::kABytes;
  uint8_t* bar_base = smem_b + kBS * HS::kBBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* a_empty = a_full + kAS;
  uint64_t* b_full = a_empty + kAS;
  uint64_t* b_empty = b_full + kBS;
  uint64_t* acc_full = b_empty + kBS;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* colsum_s = reinterpret_cast<float*>(bar_base + HS::kBarBytes);
  uint8_t* store_s = bar_base + HS::kBarBytes + HS::kColsumBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (g.flags & EPI_COLSUM)
    for (int c = threadIdx.x; c < g.tiles_n * BN; c += kGemmThreads) colsum_s[c] = 0.f;

  const int m_tiles = g.tiles_x * g.tiles_y * g.tiles_b;
  const int total_tiles = m_tiles * g.tiles_n;
  TileSched sched;   // as in conv_gemm_kernel
  sched.resp = smem_u32(bar_base + 512);
  sched.full = reinterpret_cast<uint64_t*>(bar_base + 512 + 16 * kSchedStages);
  sched.empty = sched.full + kSchedStages;
  sched.dyn = g.dyn;
  sched.stride = gridDim.x;
  sched.total = total_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nseg; ++s) {
      tma_prefetch_desc(&maps.a[s]);
      tma_prefetch_desc(&maps.b[s]);
    }
    for (int s = 0; s < kAS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kBS; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kEpiWarps);
    }
    for (int s = 0; s < kSchedStages; ++s) {
      mbar_init(&sched.full[s], 1);
      mbar_init(&sched.empty[s], kEpiWarps + 2);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched from here on

  if (warp == 0) {
    // ============================== TMA producer (whole warp, one elected lane issues) ==============================
    {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      if constexpr (RB) {
        // resident weights: groups (set, cb, kh) of three taps; set 0 = hi (maps.b[0]), set 1 = lo (maps.b[1])
        const int nsets = g.nseg >= 2 ? 2 : 1;
        if (elect_one()) {
          mbar_expect_tx(&b_full[0], static_cast<uint32_t>(nsets * g.cblocks * 3) * HS::kBBytes);
          for (int set = 0; set < nsets; ++set)
            for (int cb = 0; cb < g.cblocks; ++cb)
              for (int kh = 0; kh < 3; ++kh) {
                uint8_t* sb = smem_b + ((set * g.cblocks + cb) * 3 + kh) * HS::kBBytes;
                if (g.b_mode == 1)
                  tma_load_3d(&maps.b[set], &b_full[0], sb, cb * 64, 0, 8 - 3 * kh - 2);
                else
                  tma_load_3d(&maps.b[set], &b_full[0], sb, 0, cb * 64, 3 * kh);
              }
        }
        __syncwarp();
      }
      for (int t = blockIdx.x, it = 0; t < total_tiles; ++it) {
        if (g.dyn) sched_issue<false>(sched, it, lane);
        const int nb = t % g.tiles_n;
        const int mt = t / g.tiles_n;
        const int tx = mt % g.tiles_x;
        const int ty = (mt / g.tiles_x) % g.tiles_y;
        const int n0 = mt / (g.tiles_x * g.tiles_y);
        const int x0 = tx << 3, y0 = ty << 4;
        for (int seg = 0; seg < g.nseg; ++seg) {
          for (int cb = 0; cb < g.cblocks; ++cb) {
            mbar_wait(&a_empty[as], aph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&a_full[as], HaloCfg::kABytes);
              tma_load_4d(&maps.a[seg], &a_full[as], smem + as * HaloCfg::kABytes, cb * 64, x0 - 1, y0 - 1, n0);
            }
            __syncwarp();
            if (++as == kAS) {
              as = 0;
              aph ^= 1;
            }
            for (int tap = 0; tap < (RB ? 0 : 9); tap += HS::kTaps) {
              mbar_wait(&b_empty[bs], bph ^ 1);
              uint8_t* sb = smem_b + bs * HS::kBBytes;
              if (elect_one()) {
                mbar_expect_tx(&b_full[bs], HS::kBBytes);
                if (g.b_mode == 1) {
                  // rotated taps 8-tap ... 8-tap-(kTaps-1): one box starting at the lowest of them (slots are then in
                  // ascending filter-tap order, i.e. descending kw -- see the MMA loop)
                  tma_load_3d(&maps.b[seg], &b_full[bs], sb, cb * 64, nb * BN, 8 - tap - (HS::kTaps - 1));
                } else {
#pragma unroll
                  for (int j = 0; j < BN / 64; ++j)
                    tma_load_3d(&maps.b[seg], &b_full[bs], sb + j * (HS::kTaps * 8192), nb * BN + j * 64, cb * 64, tap);
                }
              }
              __syncwarp();
              if (++bs == kBS) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
        }
        t = sched_next<false>(sched, it, t, lane);
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (whole warp, one elected lane issues) ==============================
    const bool b_mn = g.b_mode == 2;
    const uint32_t idesc = make_idesc(1u, 0u, b_mn ? 1u : 0u, 128u, BN);
    const uint64_t adesc0 = make_smem_desc_sw128(0, 16, 2048);
    // MN-major B (fprop): the 64-wide N chunks of a tap are kTaps * 8 KB apart (LBO) when a stage holds several taps
    const uint64_t bdesc0 = b_mn ? make_smem_desc_sw128(0, HS::kTaps * 8192, 1024) : make_smem_desc_sw128(0, 16, 1024);
    const uint32_t badv = b_mn ? 128u : 2u;
    const uint32_t sa_base = smem_u32(smem), sb_base = smem_u32(smem_b);
    int as = 0, bs = 0, acs = 0;
    uint32_t aph = 0, bph = 0, acph = 0;
    if constexpr (RB) {
      mbar_wait(&b_full[0], 0);   // the resident weight slice has landed
      tc_fence_after();
    }
    long long dbg_a = 0, dbg_b = 0, dbg_acc = 0, dbg_kb = 0, tw = 0;
    const long long dbg_t0 = g.dbg ? clock64() : 0;
    int t_next = 0;
    for (int t = blockIdx.x, it = 0; t < total_tiles; t = t_next, ++it) {
      t_next = sched_next<false>(sched, it, t, lane);
      if (g.dbg) tw = clock64();
      mbar_wait(&acc_empty[acs], acph ^ 1);
      if (g.dbg) dbg_acc += clock64() - tw;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acs * BN;
      uint32_t first = 0;
      for (int kbk = 0; kbk < g.nseg * g.cblocks; ++kbk) {
        if (g.dbg) tw = clock64();
        mbar_wait(&a_full[as], aph);
        if (g.dbg) {
          dbg_a += clock64() - tw;
          ++dbg_kb;
        }
        tc_fence_after();
        const uint32_t sa = sa_base + as * HaloCfg::kABytes;
        const int rb_seg = kbk / g.cblocks;
        const int rb_group0 = ((rb_seg == 1 ? 1 : 0) * g.cblocks + (kbk - rb_seg * g.cblocks)) * 3;
#pragma unroll 1
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
          for (int kw0 = 0; kw0 < 3; kw0 += HS::kTaps) {
            if constexpr (!RB) {
              if (g.dbg) tw = clock64();
              mbar_wait(&b_full[bs], bph);
              if (g.dbg) dbg_b += clock64() - tw;
              tc_fence_after();
            }
            const uint32_t sb = sb_base + (RB ? (rb_group0 + kh) : bs) * HS::kBBytes;
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < HS::kTaps; ++j) {
                const int kw = kw0 + j;
                // tap window: starts (kh*16 + kw) 128-byte rows into the halo box; 8-row groups 2048 B apart.  (The
                // 128-byte swizzle XOR is taken from the absolute shared-memory address bits [7,10) -- measured in
                // scripts/bringup.py::halo_conv: a start that is not 1024-byte aligned needs NO base-offset field.)
                const uint32_t a_addr = sa + static_cast<uint32_t>(kh * 16 + kw) * 128u;
                const uint64_t adesc = adesc0 | static_cast<uint64_t>((a_addr & 0x3FFFF) >> 4);
                // slot of this tap inside the stage: fprop boxes hold taps in order; dgrad boxes hold the rotated taps
                // in ascending filter order, i.e. kw descending
                const int slot = b_mn ? j : (HS::kTaps - 1 - j);
                const uint32_t b_addr = sb + static_cast<uint32_t>(slot) * (b_mn ? 8192u : static_cast<uint32_t>(BN) * 128u);
                const uint64_t bdesc = bdesc0 | static_cast<uint64_t>((b_addr & 0x3FFFF) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_f16(d_tmem, adesc + 2 * k, bdesc + badv * k, idesc, first | static_cast<uint32_t>(j + k > 0));
                }
                first = 1;
              }
              if constexpr (!RB) umma_commit(&b_empty[bs]);
            }
            __syncwarp();
            first = 1;
            if constexpr (!RB) {
              if (++bs == kBS) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
        }
        if (elect_one()) umma_commit(&a_empty[as]);
        __syncwarp();
        if (++as == kAS) {
          as = 0;
          aph ^= 1;
        }
      }
      if (elect_one()) umma_commit(&acc_full[acs]);
      __syncwarp();
      if (++acs == 2) {
        acs = 0;
        acph ^= 1;
      }
    }
    if (g.dbg && lane == 0 && blockIdx.x < 148) {   // as in conv_gemm_kernel; [1] = activation patches, [5] = weight stages, [3] = patches
      g.dbg[8 * blockIdx.x + 0] = clock64() - dbg_t0;
      g.dbg[8 * blockIdx.x + 1] = dbg_a;
      g.dbg[8 * blockIdx.x + 2] = dbg_acc;
      g.dbg[8 * blockIdx.x + 3] = dbg_kb;
      g.dbg[8 * blockIdx.x + 5] = dbg_b;
    }
  } else {
    // ============================== epilogue ==============================
    conv_epilogue_loop<BN, false>(g, tmem_base, acc_full, acc_empty, colsum_s, store_s, total_tiles, m_tiles, warp,
                                  lane, 0, &sched);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
  if (g.flags & EPI_COLSUM) {
    const int ncs = g.colsum_n > 0 ? min(g.colsum_n, g.tiles_n * BN) : g.tiles_n * BN;
    for (int c = threadIdx.x; c < ncs; c += kGemmThreads) {
      const float t = colsum_s[c];
      if (t != 0.f) atomicAdd(g.colsum + c, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// conv_halo_pair_kernel: conv_halo_kernel on CTA pairs (tcgen05 cta_group::2) for the narrow tiles (N = 64 / 128).
// A single CTA's N = 64 MMA needs 4 KB of A + 2 KB of B from shared memory per 32-cycle instruction, N = 128 needs
// 4 + 4 KB, and the TMA fills of the weight stages go through the same 128 B/clk port: the single-CTA kernel runs at
// 85 cycles per N = 64 MMA (36 % tensor pipe) and 50 % for N = 128 (profiles/r02_launches_fp32.md).  In a pair each
// CTA owns one 16 x 8-pixel tile and its halo patch but only HALF of the weight rows (N/2): the M = 256 MMA reads
// 4 KB + 1 (2) KB per CTA, and the weight traffic per CTA halves -- for N = 64 with <= 2 (set, channel-block)
// combinations the CTA's share of ALL nine taps (hi and lo) is 72 KB and stays resident for the whole kernel (RB).
// K-major weights (dgrad, b_mode 1; packed forward operand, b_mode 0) for every width; MN-major weights (fprop from the
// TF layout, b_mode 2) need N/2 >= 64 columns per CTA, i.e. BN >= 128 -- conv1_2's forward therefore reads a packed copy.
// Roles, barriers and the scheduler as in conv_gemm_kernel<.., PAIR>; epilogue shared.
template <int BN, bool RB = false>
struct HaloPairSmem {
  static constexpr int kTaps = 3;                            // filter taps per weight stage: one kh row (see HaloSmem)
  static constexpr int kBRows = BN / 2;                      // weight rows this CTA holds
  static constexpr int kBBytes = kTaps * kBRows * 128;       // one stage / one resident (set, cb, kh) group
  static constexpr int kAStages = 3;
  static constexpr int kBStages = RB ? 6 : (BN == 64 ? 8 : 4);   // 96 KB of weight stages in flight
  static constexpr int kBarBytes = 1024;
  static constexpr int kColsumBytes = 1024;
  static constexpr int kStoreBytes = kEpiWarpsC * kStoreWarpBytes;
  static constexpr int kBytes = kAStages * HaloCfg::kABytes + kBStages * kBBytes + kBarBytes + kColsumBytes +
                                kStoreBytes + 1024;
  static_assert(kBytes <= 232448, "shared memory of the pair halo kernel");
};

template <int BN, bool RB = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_halo_pair_kernel(const __grid_constant__ TensorMaps3 maps, const ConvGemmArgs g) {
  pdl_launch_dependents();
  static_assert(BN == 64 || BN == 128, "pair halo tiles: N = 64 or 128");
  static_assert(!RB || BN == 64, "resident weights only for the N = 64 tiles");
  using HS = HaloPairSmem<BN, RB>;
  constexpr int kAS = HS::kAStages, kBS = HS::kBStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + kAS * HaloCfg::kABytes;
  uint8_t* bar_base = smem_b + kBS * HS::kBBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* a_empty = a_full + kAS;
  uint64_t* b_full = a_empty + kAS;
  uint64_t* b_empty = b_full + kBS;
  uint64_t* acc_full = b_empty + kBS;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* colsum_s = reinterpret_cast<float*>(bar_base + HS::kBarBytes);
  uint8_t* store_s = bar_base + HS::kBarBytes + HS::kColsumBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (g.flags & EPI_COLSUM)
    for (int c = threadIdx.x; c < g.tiles_n * BN; c += kGemmThreads) colsum_s[c] = 0.f;

  const int real_m_tiles = g.tiles_x * g.tiles_y * g.tiles_b;
  const int m_pairs = (real_m_tiles + 1) / 2;
  const int total_units = m_pairs * g.tiles_n;
  const int t_first = blockIdx.x >> 1, t_step = gridDim.x >> 1;
  TileSched sched;
  sched.resp = smem_u32(bar_base + 512);
  sched.full = reinterpret_cast<uint64_t*>(bar_base + 512 + 16 * kSchedStages);
  sched.empty = sched.full + kSchedStages;
  sched.dyn = g.dyn;
  sched.stride = t_step;
  sched.total = total_units;
  const int nsets = g.nseg >= 2 ? 2 : 1;   // distinct weight operands: segment s reads set min(s, 1)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nseg; ++s) {
      tma_prefetch_desc(&maps.a[s]);
      tma_prefetch_desc(&maps.b[s]);
    }
    for (int s = 0; s < kAS; ++s) {
      mbar_init(&a_full[s], 2);    // the leader's expect_tx arrive + the partner's plain arrive
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kBS; ++s) {
      mbar_init(&b_full[s], 2);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kEpiWarps);   // alternating epilogue groups: 4 warps per tile in each CTA
    }
    for (int s = 0; s < kSchedStages; ++s) {
      mbar_init(&sched.full[s], 1);
      mbar_init(&sched.empty[s], 2 * kEpiWarps + 3);
    }
    fence_mbar_init();
  }
  cluster_sync_all();
  if (warp == 1) tmem_alloc_pair<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ============================== TMA producer (both CTAs; bytes are counted on the leader's barriers) =============
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    if constexpr (RB) {
      // this CTA's half of every weight row group (set, cb, kh), once
      const int ngroups = nsets * g.cblocks * 3;
      if (elect_one()) {
        const uint32_t lead = map_to_cta(smem_u32(&b_full[0]), 0);
        if (rank == 0)
          mbar_expect_tx(&b_full[0], static_cast<uint32_t>(2 * ngroups) * HS::kBBytes);
        else
          mbar_arrive_cluster(lead);
        for (int set = 0; set < nsets; ++set)
          for (int cb = 0; cb < g.cblocks; ++cb)
            for (int kh = 0; kh < 3; ++kh) {
              uint8_t* sb = smem_b + ((set * g.cblocks + cb) * 3 + kh) * HS::kBBytes;
              if (g.b_mode == 1) {
                tma_load_3d_pair(&maps.b[set], lead, sb, cb * 64, static_cast<int>(rank) * HS::kBRows, 8 - 3 * kh - 2);
              } else {   // b_mode 0: packed K-major [Cout][(tap, ci)], one box per tap, taps in filter order
#pragma unroll
                for (int j = 0; j < 3; ++j)
                  tma_load_2d_pair(&maps.b[set], lead, sb + j * (HS::kBRows * 128), ((3 * kh + j) * g.cblocks + cb) * 64,
                                   static_cast<int>(rank) * HS::kBRows);
              }
            }
      }
      __syncwarp();
    }
    for (int t = t_first, it = 0; t < total_units; ++it) {
      if (g.dyn && rank == 0) sched_issue<true>(sched, it, lane);
      const int nb = t % g.tiles_n;
      const int mt = 2 * (t / g.tiles_n) + static_cast<int>(rank);
      const int tx = mt % g.tiles_x;
      const int ty = (mt / g.tiles_x) % g.tiles_y;
      const int n0 = mt / (g.tiles_x * g.tiles_y);   // (a partner past the last tile reads rows outside: zero fill)
      const int x0 = tx << 3, y0 = ty << 4;
      const int nrow = nb * BN + static_cast<int>(rank) * HS::kBRows;
      for (int seg = 0; seg < g.nseg; ++seg) {
        for (int cb = 0; cb < g.cblocks; ++cb) {
          mbar_wait(&a_empty[as], aph ^ 1);
          if (elect_one()) {
            const uint32_t lead = map_to_cta(smem_u32(&a_full[as]), 0);
            if (rank == 0)
              mbar_expect_tx(&a_full[as], 2 * HaloCfg::kABytes);
            else
              mbar_arrive_cluster(lead);
            tma_load_4d_pair(&maps.a[seg], lead, smem + as * HaloCfg::kABytes, cb * 64, x0 - 1, y0 - 1, n0);
          }
          __syncwarp();
          if (++as == kAS) {
            as = 0;
            aph ^= 1;
          }
          for (int tap = 0; tap < (RB ? 0 : 9); tap += HS::kTaps) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            uint8_t* sb = smem_b + bs * HS::kBBytes;
            if (elect_one()) {
              const uint32_t lead = map_to_cta(smem_u32(&b_full[bs]), 0);
              if (rank == 0)
                mbar_expect_tx(&b_full[bs], 2 * HS::kBBytes);
              else
                mbar_arrive_cluster(lead);
              if (g.b_mode == 1) {
                tma_load_3d_pair(&maps.b[seg], lead, sb, cb * 64, nrow, 8 - tap - (HS::kTaps - 1));
              } else if (g.b_mode == 0) {
#pragma unroll
                for (int j = 0; j < HS::kTaps; ++j)
                  tma_load_2d_pair(&maps.b[seg], lead, sb + j * (HS::kBRows * 128), ((tap + j) * g.cblocks + cb) * 64, nrow);
              } else {
#pragma unroll
                for (int j = 0; j < HS::kBRows / 64; ++j)
                  tma_load_3d_pair(&maps.b[seg], lead, sb + j * (HS::kTaps * 8192), nrow + j * 64, cb * 64, tap);
              }
            }
            __syncwarp();
            if (++bs == kBS) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
      }
      t = sched_next<true>(sched, it, t, lane);
    }
  } else if (warp == 1 && rank == 0) {
    // ============================== MMA issuer (leader CTA, M = 256 over both CTAs' shared memory) ==================
    const bool b_mn = g.b_mode == 2;
    const uint32_t idesc = make_idesc(1u, 0u, b_mn ? 1u : 0u, 256u, BN);
    const uint64_t adesc0 = make_smem_desc_sw128(0, 16, 2048);
    const uint64_t bdesc0 = b_mn ? make_smem_desc_sw128(0, HS::kTaps * 8192, 1024) : make_smem_desc_sw128(0, 16, 1024);
    const uint32_t badv = b_mn ? 128u : 2u;
    const uint32_t sa_base = smem_u32(smem), sb_base = smem_u32(smem_b);
    int as = 0, bs = 0, acs = 0;
    uint32_t aph = 0, bph = 0, acph = 0;
    if constexpr (RB) {
      mbar_wait(&b_full[0], 0);
      tc_fence_after();
    }
    int t_next = 0;
    long long dbg_a = 0, dbg_b = 0, dbg_acc = 0, dbg_kb = 0, tw = 0;
    const long long dbg_t0 = g.dbg ? clock64() : 0;
    for (int t = t_first, it = 0; t < total_units; t = t_next, ++it) {
      t_next = sched_next<true>(sched, it, t, lane);
      if (g.dbg) tw = clock64();
      mbar_wait(&acc_empty[acs], acph ^ 1);
      if (g.dbg) dbg_acc += clock64() - tw;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acs * BN;
      uint32_t first = 0;
      for (int kbk = 0; kbk < g.nseg * g.cblocks; ++kbk) {
        if (g.dbg) tw = clock64();
        mbar_wait(&a_full[as], aph);
        if (g.dbg) {
          dbg_a += clock64() - tw;
          ++dbg_kb;
        }
        tc_fence_after();
        const uint32_t sa = sa_base + as * HaloCfg::kABytes;
        const int seg = kbk / g.cblocks;
        const int rb_group0 = ((seg >= 1 ? nsets - 1 : 0) * g.cblocks + (kbk - seg * g.cblocks)) * 3;
#pragma unroll 1
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
          for (int kw0 = 0; kw0 < 3; kw0 += HS::kTaps) {
            if constexpr (!RB) {
              if (g.dbg) tw = clock64();
              mbar_wait(&b_full[bs], bph);
              if (g.dbg) dbg_b += clock64() - tw;
              tc_fence_after();
            }
            const uint32_t sb = sb_base + (RB ? (rb_group0 + kh) : bs) * HS::kBBytes;
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < HS::kTaps; ++j) {
                const int kw = kw0 + j;
                const uint32_t a_addr = sa + static_cast<uint32_t>(kh * 16 + kw) * 128u;
                const uint64_t adesc = adesc0 | static_cast<uint64_t>((a_addr & 0x3FFFF) >> 4);
                // slot of this tap in the stage: forward operands hold the taps in filter order, the dgrad box holds
                // the rotated taps in ascending filter order, i.e. kw descending
                const int slot = (g.b_mode != 1) ? j : (HS::kTaps - 1 - j);
                const uint32_t b_addr = sb + static_cast<uint32_t>(slot) * (b_mn ? 8192u : static_cast<uint32_t>(HS::kBRows) * 128u);
                const uint64_t bdesc = bdesc0 | static_cast<uint64_t>((b_addr & 0x3FFFF) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + badv * k, idesc, first | static_cast<uint32_t>(j + k > 0));
                first = 1;
              }
              if constexpr (!RB) umma_commit_pair(&b_empty[bs], 3);
            }
            __syncwarp();
            first = 1;
            if constexpr (!RB) {
              if (++bs == kBS) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
        }
        if (elect_one()) umma_commit_pair(&a_empty[as], 3);
        __syncwarp();
        if (++as == kAS) {
          as = 0;
          aph ^= 1;
        }
      }
      if (elect_one()) umma_commit_pair(&acc_full[acs], 3);
      __syncwarp();
      if (++acs == 2) {
        acs = 0;
        acph ^= 1;
      }
    }
    if (g.dbg && lane == 0 && blockIdx.x < 148) {   // as in conv_halo_kernel
      g.dbg[8 * blockIdx.x + 0] = clock64() - dbg_t0;
      g.dbg[8 * blockIdx.x + 1] = dbg_a;
      g.dbg[8 * blockIdx.x + 2] = dbg_acc;
      g.dbg[8 * blockIdx.x + 3] = dbg_kb;
      g.dbg[8 * blockIdx.x + 5] = dbg_b;
    }
  } else if (warp >= 2) {
    conv_epilogue_loop<BN, false, true, 0, false, true>(g, tmem_base, acc_full, acc_empty, colsum_s, store_s, total_units,
                                                        m_pairs, warp, lane, rank, &sched);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<2 * BN>(tmem_base);
  }
  if (g.flags & EPI_COLSUM) {
    const int ncs = g.colsum_n > 0 ? min(g.colsum_n, g.tiles_n * BN) : g.tiles_n * BN;
    for (int c = threadIdx.x; c < ncs; c += kGemmThreads) {
      const float t = colsum_s[c];
      if (t != 0.f) atomicAdd(g.colsum + c, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// wgrad: D[(tap,ci) rows, co cols] = sum over pixels.  A = X (shifted per tap), B = dY, both MN-major:
// a smem "chunk" is [64 pixels][128 B of channels], 128B-swizzled, exactly what a TMA box {CH, pbw, pbh, pbn} gives.
struct WgradArgs {
  float* out;      // [rows_valid, ldc] fp32 (HWIO flattened: row = tap*Cin + ci)
  float* partial;  // [splits][rows_pad][ldc]
  int N, H, W;
  int Cin, ldc;  // ldc = Cout
  int taps, taps_w, pad;
  int nseg;
  int rows_valid;    // taps*Cin, or 27 for the im2col'ed conv1_1
  int total_chunks;  // taps*Cin/CH
  int m_tiles, tiles_n;
  int lbw, lbh, lbn;           // 64-pixel patch
  int pb_x, pb_y, pb_b;        // pixel blocks per dim
  int splits, pb_per_split;
  int flags;  // EPI_PARTIAL or 0
  // b_mode 1: dY is the padded, blocked output gradient of a transposed convolution (5-D map (s*CP, Wb, s, Hb, N));
  // output column c = (dy, dx, co) reads row dy = c / blk_row, element c % blk_row of the blocks.
  int b_mode, blk_row;
  float acc_scale;  // round-toward-zero compensation of the TMEM accumulation (see ConvGemmArgs::acc_scale)
  long long* dbg;   // measurement only: per-CTA wait-cycle counters (see ConvGemmArgs::dbg)
  int dyn;          // dynamic tile scheduling (see ConvGemmArgs::dyn)
};

template <int BN, bool TF32>
struct WgradPix {
  static constexpr int value = !TF32 ? 128 : 64;
};

// PAIR: CTA pairs as in conv_gemm_kernel -- the two CTAs own consecutive M tiles (rows of the flattened filter), share
// the pixel range and the N tile, and each loads half of the dY tile: 64 KB per 128-pixel stage, three stages (six
// 64-pixel stages of 32 KB measured 8 % slower: twice the TMA operations per MMA).
template <int BN, bool TF32, bool PAIR = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
wgrad_gemm_kernel(const __grid_constant__ TensorMaps3 maps, const WgradArgs g) {
  pdl_launch_dependents();
  static_assert(!PAIR || (!TF32 && BN == 256), "CTA pairs: bf16 operands, 256-column tiles");
  using Cfg = GemmCfg<BN>;
  constexpr int CH = TF32 ? 32 : 64;
  // pixels (K) per pipeline stage: 128 for the narrow bf16 tiles, whose per-stage MMA time would otherwise be shorter
  // than the issue + barrier round trip of a stage; 64 where shared memory is tight (BN = 256) and for tf32
  constexpr int PIX = PAIR ? 128 : WgradPix<BN, TF32>::value;
  constexpr int kChunkBytes = PIX * 128;       // PIX pixels x 128 B
  constexpr int MCH = 128 / CH;                // A chunks per M tile
  constexpr int NCH = (PAIR ? BN / 2 : BN) / CH;   // B chunks per N tile that this CTA loads
  constexpr int kAStage = MCH * kChunkBytes;   // 16 KB bf16 / 32 KB tf32
  constexpr int kBStage = NCH * kChunkBytes;
  constexpr int kStage = kAStage + kBStage;
  constexpr int kStages = (200 * 1024) / kStage;
  constexpr int KPM = TF32 ? 8 : 16;           // K (pixels) per MMA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + kStages * kStage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  const int m_sched = PAIR ? (g.m_tiles + 1) / 2 : g.m_tiles;   // scheduling units along M: tiles or pairs of tiles
  const int total_tiles = m_sched * g.tiles_n * g.splits;
  const int t_first = PAIR ? blockIdx.x >> 1 : blockIdx.x;
  const int t_step = PAIR ? gridDim.x >> 1 : gridDim.x;
  const int pb_per_seg = g.pb_x * g.pb_y * g.pb_b;
  const int total_pb = g.nseg * pb_per_seg;
  const int cin_chunks = g.Cin / CH;
  TileSched sched;   // as in conv_gemm_kernel
  sched.resp = smem_u32(bar_base + 512);
  sched.full = reinterpret_cast<uint64_t*>(bar_base + 512 + 16 * kSchedStages);
  sched.empty = sched.full + kSchedStages;
  sched.dyn = g.dyn;
  sched.stride = t_step;
  sched.total = total_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nseg; ++s) {
      tma_prefetch_desc(&maps.a[s]);
      tma_prefetch_desc(&maps.b[s]);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], PAIR ? 2 : 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], PAIR ? 2 * kEpiWarps : kEpiWarps);
    }
    for (int s = 0; s < kSchedStages; ++s) {
      mbar_init(&sched.full[s], 1);
      mbar_init(&sched.empty[s], PAIR ? 2 * kEpiWarps + 3 : kEpiWarps + 2);
    }
    fence_mbar_init();
  }
  if (PAIR) cluster_sync_all();
  if (warp == 1) {
    if (PAIR)
      tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
    else
      tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched from here on

  if (warp == 0) {
    {   // TMA producer: whole warp runs the uniform loop, one elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_first, it = 0; t < total_tiles; ++it) {
        if (g.dyn && rank == 0) sched_issue<PAIR>(sched, it, lane);
        const int nb = t % g.tiles_n;
        const int mt = PAIR ? 2 * ((t / g.tiles_n) % m_sched) + static_cast<int>(rank) : (t / g.tiles_n) % m_sched;
        const int sp = t / (g.tiles_n * m_sched);
        // per-chunk tap offsets for this M tile
        int ccb[MCH], cdx[MCH], cdy[MCH];
        int nvalid = 0;
#pragma unroll
        for (int c = 0; c < MCH; ++c) {
          const int gch = mt * MCH + c;
          if (gch < g.total_chunks) {
            const int tap = gch / cin_chunks;
            ccb[c] = (gch - tap * cin_chunks) * CH;
            cdy[c] = tap / g.taps_w - g.pad;
            cdx[c] = tap % g.taps_w - g.pad;
            nvalid = c + 1;
          } else {
            ccb[c] = 0;
            cdx[c] = 0;
            cdy[c] = 0;
          }
        }
        const uint32_t tx_bytes = nvalid * kChunkBytes + kBStage;
        const int pb0 = sp * g.pb_per_split;
        const int pb1 = min(total_pb, pb0 + g.pb_per_split);
        for (int pb = pb0; pb < pb1; ++pb) {
          const int seg = pb / pb_per_seg;
          const int r = pb - seg * pb_per_seg;
          const int bx = r % g.pb_x;
          const int by = (r / g.pb_x) % g.pb_y;
          const int bb = r / (g.pb_x * g.pb_y);
          const int x0 = bx << g.lbw, y0 = by << g.lbh, n0 = bb << g.lbn;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStage;
          uint8_t* sb = sa + kAStage;
          if (PAIR) {
            // every chunk is loaded (rows past the filter read channel 0 / tap 0 and are never stored), so both CTAs
            // always deliver kStage bytes to the leader's barrier
            if (elect_one()) {
              const uint32_t lead_bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
              if (rank == 0)
                mbar_expect_tx(&full_bar[stage], 2 * kStage);
              else
                mbar_arrive_cluster(lead_bar);
#pragma unroll
              for (int c = 0; c < MCH; ++c)
                tma_load_4d_pair(&maps.a[seg], lead_bar, sa + c * kChunkBytes, ccb[c], x0 + cdx[c], y0 + cdy[c], n0);
#pragma unroll
              for (int j = 0; j < NCH; ++j)
                tma_load_4d_pair(&maps.b[seg], lead_bar, sb + j * kChunkBytes,
                                 nb * BN + (static_cast<int>(rank) * NCH + j) * CH, x0, y0, n0);
            }
          } else if (elect_one()) {
            mbar_expect_tx(&full_bar[stage], tx_bytes);
#pragma unroll
            for (int c = 0; c < MCH; ++c)
              if (c < nvalid)
                tma_load_4d(&maps.a[seg], &full_bar[stage], sa + c * kChunkBytes, ccb[c], x0 + cdx[c], y0 + cdy[c], n0);
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
              const int col = nb * BN + j * CH;
              if (g.b_mode == 0) {
                tma_load_4d(&maps.b[seg], &full_bar[stage], sb + j * kChunkBytes, col, x0, y0, n0);
              } else {
                const int bdy = col / g.blk_row;
                tma_load_5d(&maps.b[seg], &full_bar[stage], sb + j * kChunkBytes, col - bdy * g.blk_row, x0, bdy, y0,
                            n0);
              }
            }
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        t = sched_next<PAIR>(sched, it, t, lane);
      }
    }
  } else if (warp == 1 && rank == 0) {
    // MMA issuer: whole warp runs the uniform loop, one elected lane issues (see conv_gemm_kernel)
    const uint32_t idesc = make_idesc(TF32 ? 2u : 1u, 1u, 1u, PAIR ? 256u : 128u, BN);
    // MN-major, 128B swizzle: LBO = stride between 128-byte channel chunks, SBO = stride between 8-pixel groups
    // (tf32: 32-byte-granule swizzle, 4-pixel / 512-byte atoms -- the only MN-major layout the tf32 MMA accepts)
    const uint64_t desc0 = make_smem_desc_sw128(0, kChunkBytes, TF32 ? 512 : 1024, TF32 ? 1u : 2u);
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    long long dbg_full = 0, dbg_acc = 0, dbg_kb = 0;
    const long long dbg_t0 = g.dbg ? clock64() : 0;
    int t_next = 0;
    for (int t = t_first, it = 0; t < total_tiles; t = t_next, ++it) {
      t_next = sched_next<PAIR>(sched, it, t, lane);
      const int sp = t / (g.tiles_n * m_sched);
      const int pb0 = sp * g.pb_per_split;
      const int pb1 = min(total_pb, pb0 + g.pb_per_split);
      long long tw = g.dbg ? clock64() : 0;
      mbar_wait(&acc_empty[as], aphase ^ 1);
      if (g.dbg) dbg_acc += clock64() - tw;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int pb = pb0; pb < pb1; ++pb) {
        if (g.dbg) tw = clock64();
        mbar_wait(&full_bar[stage], phase);
        if (g.dbg) {
          dbg_full += clock64() - tw;
          ++dbg_kb;
        }
        tc_fence_after();
        const uint32_t sa = smem_base + stage * kStage;
        const uint64_t adesc = desc0 | static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
        const uint64_t bdesc = desc0 | static_cast<uint64_t>(((sa + kAStage) & 0x3FFFF) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < PIX / KPM; ++k) {
            const uint32_t adv = (k * KPM * 128) >> 4;
            if constexpr (PAIR)
              umma_f16_pair(d_tmem, adesc + adv, bdesc + adv, idesc, (pb > pb0 || k > 0) ? 1u : 0u);
            else
              umma_issue<TF32>(d_tmem, adesc + adv, bdesc + adv, idesc, (pb > pb0 || k > 0) ? 1u : 0u);
          }
          if constexpr (PAIR)
            umma_commit_pair(&empty_bar[stage], 3);
          else
            umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) {
        if constexpr (PAIR)
          umma_commit_pair(&acc_full[as], 3);
        else
          umma_commit(&acc_full[as]);
      }
      __syncwarp();
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (g.dbg && lane == 0 && blockIdx.x < 148) {
      g.dbg[8 * blockIdx.x + 0] = clock64() - dbg_t0;
      g.dbg[8 * blockIdx.x + 1] = dbg_full;
      g.dbg[8 * blockIdx.x + 2] = dbg_acc;
      g.dbg[8 * blockIdx.x + 3] = dbg_kb;
    }
  } else if (warp >= 2) {
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    const size_t rows_pad = static_cast<size_t>(g.m_tiles) * 128;
    const uint32_t lead_acc_empty = PAIR ? map_to_cta(smem_u32(acc_empty), 0) : 0;
    int t_next = 0;
    for (int t = t_first, it = 0; t < total_tiles; t = t_next, ++it) {
      t_next = sched_next<PAIR>(sched, it, t, lane);
      const int nb = t % g.tiles_n;
      const int mt = PAIR ? 2 * ((t / g.tiles_n) % m_sched) + static_cast<int>(rank) : (t / g.tiles_n) % m_sched;
      const int sp = t / (g.tiles_n * m_sched);
      const int grow = mt * 128 + row;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      float* dst = (g.flags & EPI_PARTIAL) ? g.partial + (static_cast<size_t>(sp) * rows_pad + grow) * g.ldc
                                           : g.out + static_cast<size_t>(grow) * g.ldc;
      const bool valid = (mt < g.m_tiles) && ((g.flags & EPI_PARTIAL) ? true : (grow < g.rows_valid));
#pragma unroll 1
      for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        if (valid) {
          float4* o4 = reinterpret_cast<float4*>(dst + nb * BN + c);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o4[i] = make_float4(__uint_as_float(v[4 * i]) * g.acc_scale, __uint_as_float(v[4 * i + 1]) * g.acc_scale,
                                __uint_as_float(v[4 * i + 2]) * g.acc_scale,
                                __uint_as_float(v[4 * i + 3]) * g.acc_scale);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR)
          mbar_arrive_cluster(lead_acc_empty + as * 8);
        else
          mbar_arrive_relaxed(&acc_empty[as]);
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    else
      tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// wgrad_halo_kernel: filter gradient of the 3x3 layers with 64 INPUT channels (conv1_2, conv2_1), where the generic
// kernel is bound by L2 -> shared-memory traffic (N = 64 / 128 tiles, one tap-shifted copy of X per m-tile).
// One CTA owns ALL 576 = 9 taps x 64 channels output rows of a 64-column slice and a range of 128-pixel patches:
// per patch it loads the activation patch ONCE with its halo (the same 18 x 16-pixel box as conv_halo_kernel) plus
// the 16 x 8-pixel dY tile, and issues 5 x 8 MMAs: accumulator j holds taps (2j, 2j+1) -- the A operand is MN-major
// with M = 128 = two 64-channel chunks that are the two taps' windows into the patch (LBO = their byte distance), K =
// pixels in groups of 8 (one patch row, 2048 B apart); the dY tile is the MN-major B operand.  52 KB of operands per
// 40 MMAs instead of 5 x 48 KB.  Five accumulators = 320 TMEM columns, no double buffering (one output tile per CTA);
// fp32 partials [split][640][Cout] go through wgrad_splitk_reduce_kernel.
struct WgradHaloArgs {
  float* partial;    // [splits][640][ldc]
  int N, H, W, ldc;  // ldc = Cout
  int nseg;
  int tiles_x, tiles_y;          // 8 x 16 pixel patches per image
  int total_patches, patches_per_split, splits, tiles_n;
  float acc_scale;
};
struct WgradHaloCfg {
  static constexpr int kXBytes = 18 * 16 * 128;  // halo box of X
  static constexpr int kDyBytes = 128 * 128;     // 16 x 8 pixels x 64 co
  static constexpr int kStage = kXBytes + kDyBytes;
  static constexpr int kStages = 4;
  static constexpr int kSmemBytes = kStages * kStage + 1024 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
wgrad_halo_kernel(const __grid_constant__ TensorMaps3 maps, const WgradHaloArgs g) {
  pdl_launch_dependents();
  static_assert(BN == 64, "one 64-column slice per CTA");
  using Cfg = WgradHaloCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + Cfg::kStages * Cfg::kStage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* acc_full = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nb = blockIdx.x % g.tiles_n;
  const int sp = blockIdx.x / g.tiles_n;
  const int p0 = sp * g.patches_per_split;
  const int p1 = min(g.total_patches, p0 + g.patches_per_split);
  const int per_img = g.tiles_x * g.tiles_y;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nseg; ++s) {
      tma_prefetch_desc(&maps.a[s]);
      tma_prefetch_desc(&maps.b[s]);
    }
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched from here on

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int seg = 0; seg < g.nseg; ++seg) {
      for (int pt = p0; pt < p1; ++pt) {
        const int n = pt / per_img;
        const int r = pt - n * per_img;
        const int x0 = (r % g.tiles_x) << 3, y0 = (r / g.tiles_x) << 4;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sx = smem + stage * Cfg::kStage;
          mbar_expect_tx(&full_bar[stage], Cfg::kStage);
          tma_load_4d(&maps.a[seg], &full_bar[stage], sx, 0, x0 - 1, y0 - 1, n);
          tma_load_4d(&maps.b[seg], &full_bar[stage], sx + Cfg::kXBytes, nb * 64, x0, y0, n);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(1u, 1u, 1u, 128u, 64);
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t first = 0;
    for (int it = 0; it < g.nseg * (p1 - p0); ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sx = smem_base + stage * Cfg::kStage;
      const uint32_t sdy = sx + Cfg::kXBytes;
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int t0 = 2 * j, t1 = (2 * j + 1 < 9) ? 2 * j + 1 : 2 * j;
          const uint32_t off0 = static_cast<uint32_t>((t0 / 3) * 16 + (t0 % 3)) * 128u;
          const uint32_t off1 = static_cast<uint32_t>((t1 / 3) * 16 + (t1 % 3)) * 128u;
          // A: MN-major, M = two 64-channel chunks (the two taps' windows, LBO apart), K groups of 8 pixels 2048 B apart
          const uint64_t adesc = make_smem_desc_sw128(sx + off0, off1 - off0, 2048, 2u);
          const uint64_t bdesc = make_smem_desc_sw128(sdy, 8192, 1024, 2u);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // 16 pixels per MMA = two patch rows of X (2 x 2048 B) / 16 rows of dY (2048 B)
            umma_f16(tmem_base + j * 64, adesc + ((k * 4096u) >> 4), bdesc + ((k * 2048u) >> 4), idesc,
                     first | static_cast<uint32_t>(k > 0));
          }
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      first = 1;
      if (++stage == Cfg::kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* dst = g.partial + static_cast<size_t>(sp) * 640 * g.ldc + nb * 64 + chalf * 32;
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {
      uint32_t v[32];
      tmem_ld32(tmem_base + j * 64 + chalf * 32 + (static_cast<uint32_t>(quarter * 32) << 16), v);
      tmem_ld_wait();
      float4* o4 = reinterpret_cast<float4*>(dst + static_cast<size_t>(j * 128 + row) * g.ldc);
      if (p1 > p0) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o4[i] = make_float4(__uint_as_float(v[4 * i]) * g.acc_scale, __uint_as_float(v[4 * i + 1]) * g.acc_scale,
                              __uint_as_float(v[4 * i + 2]) * g.acc_scale, __uint_as_float(v[4 * i + 3]) * g.acc_scale);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace fcn8
