// C ABI of libfcn8s_sm100.so (declared in include/fcn8s_b200.h): argument validation, TMA tensor-map encoding,
// tile / split-K heuristics and kernel launches.  No allocation, no synchronisation, no torch types.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/fcn8s_b200.h"
#include "conv1.h"
#include "conv_gemm.cuh"
#include "kernels.h"

using namespace fcn8;

namespace fcn8 {
unsigned long long g_launch_count = 0;
int g_pdl_off = 1;
}

namespace {

thread_local char g_err[512] = "";
int g_debug[16] = {0};
// measurement only: device buffer of [slots][148][8] int64 wait-cycle counters, one slot per conv / wgrad launch
long long* g_dbg_buf = nullptr;
int g_dbg_slots = 0, g_dbg_next = 0;
long long* next_dbg_slot() {
  if (!g_dbg_buf || g_dbg_next >= g_dbg_slots) return nullptr;
  return g_dbg_buf + (size_t)(g_dbg_next++) * 148 * 8;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  return fail(FCN8_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// NHWC activation map: dims (C, W, H, N), box (CH, bw, bh, bn), 128B swizzle, OOB -> 0 (= SAME zero padding).
int encode_act_map(CUtensorMap* m, const void* ptr, int dtype, int N, int H, int W, int C, int bw, int bh, int bn,
                   bool atom32 = false, int ld = 0, long long sH = 0, long long sN = 0) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const int es = dtype == FCN8_BF16 ? 2 : 4;
  if (ld <= 0) ld = C;
  if (sH <= 0) sH = (long long)W * ld;     // row / image strides in elements (strided views: see Fcn8ConvParams)
  if (sN <= 0) sN = (long long)H * sH;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)ld * es, (cuuint64_t)sH * es, (cuuint64_t)sN * es};
  const cuuint32_t box[4] = {(cuuint32_t)(128 / es), (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, dtype == FCN8_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled(act N%d H%d W%d C%d box %d,%d,%d) failed: %d", N, H, W, C, bw,
                bh, bn, (int)r);
  return 0;
}
// Packed weight map: rows = Cout, cols = Ktot (K-major), box (CH, BN).
int encode_w_map(CUtensorMap* m, const void* ptr, int dtype, int rows, int ktot, int bn) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const int es = dtype == FCN8_BF16 ? 2 : 4;
  const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ktot * es};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)bn};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dtype == FCN8_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled(w rows %d k %d bn %d) failed: %d", rows, ktot, bn, (int)r);
  return 0;
}

// bf16 weight tensor in TF layout [taps][Cin_w][Cout_w] as dims (co, ci, tap); box (64 co, box_ci, 1).
int encode_hwio_map(CUtensorMap* m, const void* ptr, int taps, int cin_w, int cout_w, int box_ci, int box_taps = 1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const cuuint64_t dims[3] = {(cuuint64_t)cout_w, (cuuint64_t)cin_w, (cuuint64_t)taps};
  const cuuint64_t strides[2] = {(cuuint64_t)cout_w * 2, (cuuint64_t)cin_w * cout_w * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)box_ci, (cuuint32_t)box_taps};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled(hwio taps %d ci %d co %d box %d) failed: %d", taps, cin_w, cout_w,
                box_ci, (int)r);
  return 0;
}

int ilog2_ceil(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
// Choose a power-of-two patch (bw, bh, bn) with bw*bh*bn = 2^total_log covering [N,H,W] with the fewest tiles.
void choose_patch(int N, int H, int W, int total_log, int* lbw, int* lbh, int* lbn) {
  long long best = -1;
  for (int lw = total_log; lw >= 0; --lw)
    for (int lh = total_log - lw; lh >= 0; --lh) {
      const int ln = total_log - lw - lh;
      const long long tiles = (long long)((W + (1 << lw) - 1) >> lw) * ((H + (1 << lh) - 1) >> lh) *
                              ((N + (1 << ln) - 1) >> ln);
      if (best < 0 || tiles < best) {
        best = tiles;
        *lbw = lw;
        *lbh = lh;
        *lbn = ln;
      }
    }
}

// cudaFuncSetAttribute is per device: the "already set" flags of the launchers are indexed by the current device
int cur_dev() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & 63;
}
// expected relative loss per tcgen05.mma of a round-toward-zero TMEM accumulation at full magnitude (see
// ConvGemmArgs::rz_c); fcn8_debug_set(7, v) overrides it with v * 1e-10 for calibration runs, (0, 1) switches it off
float rz_per_mma();
int g_sm_limit = 0;  // fcn8_set_sm_limit: persistent GEMM grids leave SMs free for a co-running collective
int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

float rz_per_mma() {
  if (g_debug[0]) return 0.f;
  return g_debug[7] > 0 ? (float)g_debug[7] * 1e-10f : kRzBiasPerMma;
}

struct ConvPlan {
  int BN, splits, kb_per_split, total_kb;
  int lbw, lbh, lbn, tiles_x, tiles_y, tiles_b, m_tiles, tiles_n;
  size_t out_elems;
};

int plan_conv(const Fcn8ConvParams* p, ConvPlan* pl) {
  const int CH = p->dtype == FCN8_BF16 ? 64 : 32;
  if (p->N <= 0 || p->H <= 0 || p->W <= 0) return fail(FCN8_ERR_BAD_SHAPE, "conv: empty tensor");
  if (p->Cin % CH) return fail(FCN8_ERR_BAD_SHAPE, "conv: Cin=%d must be a multiple of %d", p->Cin, CH);
  if (p->Cout % 64) return fail(FCN8_ERR_BAD_SHAPE, "conv: Cout=%d must be a multiple of 64", p->Cout);
  if (!(p->ksize & 1)) return fail(FCN8_ERR_BAD_SHAPE, "conv: ksize must be odd");
  if (p->nseg < 1 || p->nseg > 3) return fail(FCN8_ERR_UNSUPPORTED, "conv: nseg must be 1, 2 or 3");
  if (p->dtype != FCN8_BF16 && (p->w_mode || p->out_lo || p->residual_lo))
    return fail(FCN8_ERR_UNSUPPORTED, "conv: w_mode / out_lo / residual_lo need FCN8_BF16 operands");
  if (p->w_mode < 0 || p->w_mode > 2) return fail(FCN8_ERR_BAD_SHAPE, "conv: w_mode must be 0, 1 or 2");
  if (p->w_mode == 2 && p->Cout % 64) return fail(FCN8_ERR_BAD_SHAPE, "conv: dgrad w_mode needs Cout %% 64 == 0");
  choose_patch(p->N, p->H, p->W, 7, &pl->lbw, &pl->lbh, &pl->lbn);
  if (p->flags & FCN8_EPI_POOL) {   // 8 x 16 pixel tiles: a 2x2 pooling window is four lanes of one epilogue warp
    pl->lbw = 3;
    pl->lbh = 4;
    pl->lbn = 0;
  }
  pl->tiles_x = (p->W + (1 << pl->lbw) - 1) >> pl->lbw;
  pl->tiles_y = (p->H + (1 << pl->lbh) - 1) >> pl->lbh;
  pl->tiles_b = (p->N + (1 << pl->lbn) - 1) >> pl->lbn;
  pl->m_tiles = pl->tiles_x * pl->tiles_y * pl->tiles_b;
  pl->total_kb = p->nseg * p->ksize * p->ksize * (p->Cin / CH);
  const int sms = num_sms();
  int bn = 0;
  if (p->force_bn) {
    bn = p->force_bn;
    if ((bn != 64 && bn != 128 && bn != 256) || p->Cout % bn)
      return fail(FCN8_ERR_BAD_SHAPE, "conv: force_bn=%d invalid for Cout=%d", bn, p->Cout);
  } else {
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
      if (p->Cout % cands[i]) continue;
      if (!bn) bn = cands[i];  // largest valid
      if ((long long)pl->m_tiles * (p->Cout / cands[i]) >= (sms * 4) / 5) {
        bn = cands[i];
        break;
      }
      if (cands[i] == 128) break;  // never go below 128 just to get more tiles; split-K handles the rest
    }
    if (p->Cout % 128) bn = 64;
  }
  pl->BN = bn;
  pl->tiles_n = p->Cout / bn;
  const long long tiles = (long long)pl->m_tiles * pl->tiles_n;
  int splits = 1;
  if (p->flags & (FCN8_EPI_COLSUM | FCN8_EPI_POOL)) {
    splits = 1;  // the fused column sums / pooling live in the GEMM kernel's epilogue, not in the split-K reduce
  } else if (p->force_splits > 0) {
    splits = p->force_splits;
  } else if (tiles * 2 <= sms) {
    splits = (int)(sms / tiles);
    const int max_by_k = pl->total_kb / 16 > 0 ? pl->total_kb / 16 : 1;  // keep >= 16 k-blocks per split
    if (splits > max_by_k) splits = max_by_k;
    if (splits > 16) splits = 16;
  } else if (pl->total_kb >= 256 && (size_t)p->N * p->H * p->W * p->Cout * sizeof(float) * 4 <= (160u << 20)) {
    // deep reductions with a small output (fc6: K = 25 088, 256 tiles = 1.73 waves): a 2..4-way split-K evens out
    // the last wave (86 % -> 99 % of the SMs busy) for ~0.3 GB of partial traffic, and shortens the round-toward-zero
    // accumulation chains of the TMEM accumulators fourfold
    double best_eff = (double)tiles / (double)(((tiles + sms - 1) / sms) * sms);
    for (int s2 = 2; s2 <= 4; ++s2) {
      const long long items = tiles * s2;
      const double eff = (double)items / (double)(((items + sms - 1) / sms) * sms);
      if (eff > best_eff + 0.08) {
        best_eff = eff;
        splits = s2;
      }
    }
  }
  if (splits > pl->total_kb) splits = pl->total_kb;
  pl->kb_per_split = (pl->total_kb + splits - 1) / splits;
  pl->splits = (pl->total_kb + pl->kb_per_split - 1) / pl->kb_per_split;
  pl->out_elems = (size_t)p->N * p->H * p->W * p->Cout;
  return 0;
}

// PROMO: promoted accumulation (ConvGemmArgs::promo_kb), bf16 operands only
template <int BN, bool TF32, bool PROMO = false>
cudaError_t launch_conv_t(const TensorMaps3& maps, const ConvGemmArgs& a, int grid, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[cur_dev()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN, TF32, false, 0, PROMO>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  (void)launch_k(conv_gemm_kernel<BN, TF32, false, 0, PROMO>, dim3(grid), dim3(PROMO ? kPromoThreads : kGemmThreads),
                 Cfg::kSmemBytes, st, maps, a);
  return cudaGetLastError();
}

// CTA-pair variant (clusters of two CTAs, tcgen05 cta_group::2): bf16 operands, 256-column tiles; `grid` is even.
template <bool PROMO = false>
cudaError_t launch_conv_pair(const TensorMaps3& maps, const ConvGemmArgs& a, int grid, cudaStream_t st) {
  using Cfg = GemmCfg<256, true>;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[cur_dev()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<256, false, true, 0, PROMO>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(PROMO ? kPromoThreads : kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_off ? 1 : 2;
  count_launch();
  return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<256, false, true, 0, PROMO>, maps, a);
}

template <int BN, bool RB = false>
cudaError_t launch_halo_t(const TensorMaps3& maps, const ConvGemmArgs& a, int grid, cudaStream_t st) {
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[cur_dev()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BN, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         HaloSmem<BN, RB>::kBytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  (void)launch_k(conv_halo_kernel<BN, RB>, dim3(grid), dim3(kGemmThreads), HaloSmem<BN, RB>::kBytes, st, maps, a);
  return cudaGetLastError();
}

// CTA-pair halo kernel (clusters of two, cta_group::2); `grid` is even
template <int BN, bool RB = false>
cudaError_t launch_halo_pair_t(const TensorMaps3& maps, const ConvGemmArgs& a, int grid, cudaStream_t st) {
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[cur_dev()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_pair_kernel<BN, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         HaloPairSmem<BN, RB>::kBytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = HaloPairSmem<BN, RB>::kBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_off ? 1 : 2;
  count_launch();
  return cudaLaunchKernelEx(&cfg, conv_halo_pair_kernel<BN, RB>, maps, a);
}

template <int BN, bool TF32>
constexpr int wgrad_smem_bytes() {
  constexpr int CH = TF32 ? 32 : 64;
  constexpr int chunk = WgradPix<BN, TF32>::value * 128;
  constexpr int stage = (128 / CH) * chunk + (BN / CH) * chunk;
  constexpr int stages = (200 * 1024) / stage;
  return stages * stage + 1024 + 1024;
}
template <int BN, bool TF32>
cudaError_t launch_wgrad_t(const TensorMaps3& maps, const WgradArgs& a, int grid, cudaStream_t st) {
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[cur_dev()];
  constexpr int smem = wgrad_smem_bytes<BN, TF32>();
  if (!attr_done) {
    cudaError_t e =
        cudaFuncSetAttribute(wgrad_gemm_kernel<BN, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  (void)launch_k(wgrad_gemm_kernel<BN, TF32>, dim3(grid), dim3(kGemmThreads), smem, st, maps, a);
  return cudaGetLastError();
}

cudaError_t launch_wgrad_pair(const TensorMaps3& maps, const WgradArgs& a, int grid, cudaStream_t st) {
  constexpr int smem = 3 * (4 * 128 * 128) + 1024 + 1024;   // three stages of (2 + 2) 128-pixel chunks
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[cur_dev()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_kernel<256, false, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_off ? 1 : 2;
  count_launch();
  return cudaLaunchKernelEx(&cfg, wgrad_gemm_kernel<256, false, true>, maps, a);
}

cudaError_t dispatch_conv(int BN, bool tf32, const TensorMaps3& maps, const ConvGemmArgs& a, int grid,
                          cudaStream_t st) {
  if (BN == 256) return tf32 ? launch_conv_t<256, true>(maps, a, grid, st) : launch_conv_t<256, false>(maps, a, grid, st);
  if (BN == 128) return tf32 ? launch_conv_t<128, true>(maps, a, grid, st) : launch_conv_t<128, false>(maps, a, grid, st);
  return tf32 ? launch_conv_t<64, true>(maps, a, grid, st) : launch_conv_t<64, false>(maps, a, grid, st);
}
cudaError_t dispatch_wgrad(int BN, bool tf32, const TensorMaps3& maps, const WgradArgs& a, int grid, cudaStream_t st) {
  if (BN == 256)
    return tf32 ? launch_wgrad_t<256, true>(maps, a, grid, st) : launch_wgrad_t<256, false>(maps, a, grid, st);
  if (BN == 128)
    return tf32 ? launch_wgrad_t<128, true>(maps, a, grid, st) : launch_wgrad_t<128, false>(maps, a, grid, st);
  return tf32 ? launch_wgrad_t<64, true>(maps, a, grid, st) : launch_wgrad_t<64, false>(maps, a, grid, st);
}

// Blocked view of the padded output (or output gradient) of a stride-s transposed convolution:
// zp[N][s*Hb][s*Wb][CP] fp32 seen as dims (s*CP, Wb, s, Hb, N); box (32, bw, 1, bh, bn) = bw*bh*bn blocks x 128 B.
int encode_blocked_map(CUtensorMap* m, const void* ptr, int N, int Hb, int Wb, int s, int CP, int bw, int bh, int bn,
                       bool atom32) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const cuuint64_t row = (cuuint64_t)s * CP * 4;       // bytes of one block row
  const cuuint64_t line = (cuuint64_t)Wb * row;        // bytes of one padded image row
  const cuuint64_t dims[5] = {(cuuint64_t)s * CP, (cuuint64_t)Wb, (cuuint64_t)s, (cuuint64_t)Hb, (cuuint64_t)N};
  const cuuint64_t strides[4] = {row, line, (cuuint64_t)s * line, (cuuint64_t)s * Hb * line};
  const cuuint32_t box[5] = {32, (cuuint32_t)bw, 1, (cuuint32_t)bh, (cuuint32_t)bn};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled(blocked N%d Hb%d Wb%d s%d CP%d) failed: %d", N, Hb, Wb, s, CP,
                (int)r);
  return 0;
}

struct WgradPlan {
  bool pair;   // CTA-pair kernel (bf16, 256-column tiles): two M tiles per cluster
  int BN, splits, pb_per_split, total_pb;
  int lbw, lbh, lbn, pb_x, pb_y, pb_b;
  int m_tiles, tiles_n, total_chunks;
  size_t rows_pad;
};
int plan_wgrad(const Fcn8WgradParams* p, WgradPlan* pl) {
  const int CH = p->dtype == FCN8_BF16 ? 64 : 32;
  if (p->N <= 0 || p->H <= 0 || p->W <= 0) return fail(FCN8_ERR_BAD_SHAPE, "wgrad: empty tensor");
  if (p->Cin % CH) return fail(FCN8_ERR_BAD_SHAPE, "wgrad: Cin=%d must be a multiple of %d", p->Cin, CH);
  if (p->Cout % 64) return fail(FCN8_ERR_BAD_SHAPE, "wgrad: Cout=%d must be a multiple of 64", p->Cout);
  if (p->nseg < 1 || p->nseg > 3) return fail(FCN8_ERR_UNSUPPORTED, "wgrad: nseg must be 1, 2 or 3");
  int bn = p->force_bn ? p->force_bn : (p->Cout % 256 == 0 ? 256 : (p->Cout % 128 == 0 ? 128 : 64));
  if (p->dtype == FCN8_F32 && bn == 256 && !p->force_bn) bn = 128;  // keep >= 3 smem stages with 4-byte operands
  if ((bn != 64 && bn != 128 && bn != 256) || p->Cout % bn)
    return fail(FCN8_ERR_BAD_SHAPE, "wgrad: BN=%d invalid for Cout=%d", bn, p->Cout);
  pl->BN = bn;
  pl->tiles_n = p->Cout / bn;
  pl->total_chunks = p->ksize * p->ksize * (p->Cin / CH);
  const int mch = 128 / CH;
  pl->m_tiles = (pl->total_chunks + mch - 1) / mch;
  pl->rows_pad = (size_t)pl->m_tiles * 128;
  pl->pair = bn == 256 && p->dtype == FCN8_BF16 && !g_debug[5];
  const int pix_log = pl->pair ? 7 : (p->dtype == FCN8_BF16 ? 7 : 6);   // = log2(pixels per pipeline stage)
  choose_patch(p->N, p->H, p->W, pix_log, &pl->lbw, &pl->lbh, &pl->lbn);
  pl->pb_x = (p->W + (1 << pl->lbw) - 1) >> pl->lbw;
  pl->pb_y = (p->H + (1 << pl->lbh) - 1) >> pl->lbh;
  pl->pb_b = (p->N + (1 << pl->lbn) - 1) >> pl->lbn;
  pl->total_pb = p->nseg * pl->pb_x * pl->pb_y * pl->pb_b;
  // CTAs that work at the same time: one per SM; a pair kernel schedules pairs of M tiles on SM pairs
  const int sms = pl->pair ? num_sms() / 2 : num_sms();
  const long long tiles = (long long)(pl->pair ? (pl->m_tiles + 1) / 2 : pl->m_tiles) * pl->tiles_n;
  int splits = 1;
  if (p->force_splits > 0) {
    splits = p->force_splits;
  } else if (tiles < sms) {
    splits = (int)(sms / tiles);  // floor: tiles * splits <= #SMs, one full wave (ceil would leave a 2nd, ~empty wave)
    const int max_by_k = pl->total_pb / 8 > 0 ? pl->total_pb / 8 : 1;
    if (splits > max_by_k) splits = max_by_k;
    if (splits > sms) splits = sms;
  }
  if (splits > pl->total_pb) splits = pl->total_pb;
  pl->pb_per_split = (pl->total_pb + splits - 1) / splits;
  pl->splits = (pl->total_pb + pl->pb_per_split - 1) / pl->pb_per_split;
  return 0;
}

}  // namespace

extern "C" {

int32_t fcn8_version(void) { return FCN8_VERSION; }
const char* fcn8_last_error(void) { return g_err; }

int32_t fcn8_device_check(int32_t dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
  if (prop.major != 10) return fail(FCN8_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is sm_100a-only", dev,
                                    prop.major, prop.minor);
  return 0;
}
uint64_t fcn8_launch_count(void) { return g_launch_count; }
int32_t fcn8_set_sm_limit(int32_t n) {
  g_sm_limit = n > 0 ? n : 0;
  return 0;
}
// CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), slicing-by-8, host code: the checksum TensorFlow's
// tensor-bundle checkpoints carry per table block and per tensor (tf_bundle.py).  `crc` = running value (0 to start).
uint32_t fcn8_crc32c(const void* data, size_t n, uint32_t crc) {
  static uint32_t table[8][256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 0xffu];
    ready = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7u)) {
    c = (c >> 8) ^ table[0][(c ^ *p++) & 0xffu];
    --n;
  }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;   // little-endian host (x86-64 / aarch64)
    c = table[7][w & 0xff] ^ table[6][(w >> 8) & 0xff] ^ table[5][(w >> 16) & 0xff] ^ table[4][(w >> 24) & 0xff] ^
        table[3][(w >> 32) & 0xff] ^ table[2][(w >> 40) & 0xff] ^ table[1][(w >> 48) & 0xff] ^ table[0][w >> 56];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ table[0][(c ^ *p++) & 0xffu];
  return ~c;
}
int32_t fcn8_debug_buffer(void* buf, int32_t slots) {
  g_dbg_buf = static_cast<long long*>(buf);
  g_dbg_slots = buf ? slots : 0;
  g_dbg_next = 0;
  return 0;
}
// Label feed (helpers/ground_truth_conversion_utils.py:84-88, batch_generator_KITTI.py:82-84 -> `labels` placeholder,
// fcn8s_tensorflow.py:110,559): the generators yield bool one-hot [n,H,W,C]; host code packs a batch to one class id
// per pixel (C bytes -> 1 byte over PCIe), fcn8_expand_labels restores the one-hot tensor on the device.
// Returns 0 when every pixel's C bytes are exactly one-hot (one byte == 1, the others 0) and `ids` is filled,
// 1 when some pixel is not (the caller then ships the one-hot batch as it is), < 0 on bad arguments.
//
// The batch is scanned as ONE flat byte string: it is one-hot iff no byte exceeds 1 and the k-th set byte (in address
// order) lies inside row k, for every k -- then its offset in the row is the class id.  Set bytes are found 64 bytes at
// a time (SSE2 compare + movemask on x86-64; a multiply-gather of the bytes' low bits elsewhere), so the cost is one
// pass of loads plus ~10 instructions per pixel: about what the memcpy into pinned memory that it replaces costs.
static inline uint64_t set_byte_mask64(const uint8_t* p, uint64_t* any) {
#if defined(__SSE2__)
  const __m128i one = _mm_set1_epi8(1);
  const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p));
  const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p + 16));
  const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p + 32));
  const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p + 48));
  const __m128i o = _mm_or_si128(_mm_or_si128(a, b), _mm_or_si128(c, d));
  *any |= static_cast<uint64_t>(_mm_cvtsi128_si64(o)) | static_cast<uint64_t>(_mm_cvtsi128_si64(_mm_srli_si128(o, 8)));
  return static_cast<uint64_t>(static_cast<uint32_t>(_mm_movemask_epi8(_mm_cmpeq_epi8(a, one)))) |
         (static_cast<uint64_t>(static_cast<uint32_t>(_mm_movemask_epi8(_mm_cmpeq_epi8(b, one)))) << 16) |
         (static_cast<uint64_t>(static_cast<uint32_t>(_mm_movemask_epi8(_mm_cmpeq_epi8(c, one)))) << 32) |
         (static_cast<uint64_t>(static_cast<uint32_t>(_mm_movemask_epi8(_mm_cmpeq_epi8(d, one)))) << 48);
#else
  uint64_t m = 0;
  for (int w = 0; w < 8; ++w) {
    uint64_t v;
    memcpy(&v, p + 8 * w, 8);
    *any |= v;
    // bit 0 of each of the 8 bytes -> 8 adjacent bits (bytes are 0 / 1 in a valid batch; others are caught by `any`)
    m |= (((v & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56) << (8 * w);
  }
  return m;
#endif
}
// Rows [p0, p1): returns false as soon as the range cannot be one-hot.
static bool pack_label_range(const uint8_t* onehot, int64_t p0, int64_t p1, int C, uint8_t* ids) {
  const uint64_t uc = static_cast<uint64_t>(C);
  const uint64_t b1 = static_cast<uint64_t>(p1) * uc;
  uint64_t b = static_cast<uint64_t>(p0) * uc;     // flat byte position
  int64_t k = p0;                                   // next row to receive its set byte
  uint64_t row0 = b;                                // first byte of row k
  uint64_t any = 0;
  for (; b + 64 <= b1; b += 64) {
    uint64_t m = set_byte_mask64(onehot + b, &any);
    while (m) {
      const uint64_t off = b + static_cast<uint64_t>(__builtin_ctzll(m)) - row0;   // wraps when the byte is before row k
      if (off >= uc || k >= p1) return false;
      ids[k++] = static_cast<uint8_t>(off);
      row0 += uc;
      m &= m - 1;
    }
  }
  for (; b < b1; ++b) {
    const uint8_t v = onehot[b];
    any |= v;
    if (v) {
      const uint64_t off = b - row0;
      if (off >= uc || k >= p1) return false;
      ids[k++] = static_cast<uint8_t>(off);
      row0 += uc;
    }
  }
  return k == p1 && !(any & ~0x0101010101010101ull);
}
int32_t fcn8_pack_labels(const uint8_t* onehot, int64_t pixels, int32_t C, uint8_t* ids, int32_t threads) {
  if (!onehot || !ids) return fail(FCN8_ERR_BAD_SHAPE, "pack_labels: null pointer");
  if (pixels < 0 || C < 1 || C > 255) return fail(FCN8_ERR_BAD_SHAPE, "pack_labels: pixels=%lld, C=%d", (long long)pixels, C);
  if (threads < 1) threads = 1;
  if (threads > 16) threads = 16;
  if (pixels < (1 << 16)) threads = 1;
  if (threads == 1) return pack_label_range(onehot, 0, pixels, C, ids) ? 0 : 1;
  std::atomic<int> bad(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      if (!pack_label_range(onehot, pixels * t / threads, pixels * (t + 1) / threads, C, ids)) bad.store(1);
    });
  for (auto& th : pool) th.join();
  return bad.load() ? 1 : 0;
}
int32_t fcn8_expand_labels(const uint8_t* ids, uint8_t* onehot, int64_t pixels, int32_t C, void* stream) {
  if (!ids || !onehot) return fail(FCN8_ERR_BAD_SHAPE, "expand_labels: null pointer");
  if (pixels < 0 || C < 1 || C > 255) return fail(FCN8_ERR_BAD_SHAPE, "expand_labels: pixels=%lld, C=%d", (long long)pixels, C);
  if (reinterpret_cast<uintptr_t>(onehot) & 3u) return fail(FCN8_ERR_BAD_ALIGN, "expand_labels: onehot not 4-byte aligned");
  if (!pixels) return 0;
  cudaError_t e = launch_expand_labels(ids, onehot, pixels, C, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "expand_labels launch");
}

int32_t fcn8_debug_set(int32_t key, int32_t value) {
  if (key < 0 || key >= 16) return fail(FCN8_ERR_BAD_SHAPE, "debug key out of range");
  g_debug[key] = value;
  if (key == 6) g_pdl_off = !value;
  return 0;
}

int32_t fcn8_preprocess_im2col(const Fcn8PreprocessParams* p, void* stream) {
  if (!p || !p->images || !p->out) return fail(FCN8_ERR_BAD_SHAPE, "preprocess: null pointer");
  if (!aligned16(p->out)) return fail(FCN8_ERR_BAD_ALIGN, "preprocess: out not 16-byte aligned");
  if (p->N <= 0 || p->H <= 0 || p->W <= 0) return fail(FCN8_ERR_BAD_SHAPE, "preprocess: empty tensor");
  cudaError_t e = launch_preprocess(p->images, p->out, p->N, p->H, p->W, p->dtype, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "preprocess launch");
}

size_t fcn8_conv_gemm_workspace_bytes(const Fcn8ConvParams* p) {
  ConvPlan pl;
  if (plan_conv(p, &pl)) return 0;
  return pl.splits > 1 ? (size_t)pl.splits * pl.out_elems * sizeof(float) : 0;
}

int32_t fcn8_conv_gemm(const Fcn8ConvParams* p, void* workspace, size_t workspace_bytes, void* stream) {
  if (!p || !p->x || !p->wp || (!p->out && !((p->flags & FCN8_EPI_POOL) && p->pool_out)))
    return fail(FCN8_ERR_BAD_SHAPE, "conv: null pointer");
  if (p->flags & FCN8_EPI_POOL) {
    if (p->dtype != FCN8_BF16 || !p->pool_out || (p->H & 1) || (p->W & 1) || !aligned16(p->pool_out) ||
        (p->pool_out_lo && !aligned16(p->pool_out_lo)))
      return fail(FCN8_ERR_BAD_SHAPE, "conv: POOL needs bf16 operands, pool_out (16-byte aligned) and even H, W");
  }
  if (!aligned16(p->x) || !aligned16(p->wp) || !aligned16(p->out) || (p->bias && !aligned16(p->bias)) ||
      (p->mask_src && !aligned16(p->mask_src)) || (p->residual && !aligned16(p->residual)))
    return fail(FCN8_ERR_BAD_ALIGN, "conv: pointers must be 16-byte aligned");
  if ((p->flags & FCN8_EPI_BIAS) && !p->bias) return fail(FCN8_ERR_BAD_SHAPE, "conv: BIAS flag without bias");
  if ((p->flags & FCN8_EPI_MASK) && !p->mask_src) return fail(FCN8_ERR_BAD_SHAPE, "conv: MASK flag without mask_src");
  if ((p->flags & FCN8_EPI_RESIDUAL) && !p->residual)
    return fail(FCN8_ERR_BAD_SHAPE, "conv: RESIDUAL flag without residual");
  if ((p->flags & FCN8_EPI_COLSUM) && (!p->colsum || p->Cout > kColsumMax))
    return fail(FCN8_ERR_BAD_SHAPE, "conv: COLSUM flag needs colsum and Cout <= %d", kColsumMax);
  if ((p->nseg == 3 && !p->x_lo) || (p->nseg >= 2 && !p->wp_lo))
    return fail(FCN8_ERR_BAD_SHAPE, "conv: nseg=3 needs x_lo and wp_lo (nseg=2: wp_lo)");
  ConvPlan pl;
  int rc = plan_conv(p, &pl);
  if (rc) return rc;
  const size_t need = pl.splits > 1 ? (size_t)pl.splits * pl.out_elems * sizeof(float) : 0;
  if (need > workspace_bytes || (need && !workspace))
    return fail(FCN8_ERR_WORKSPACE, "conv: workspace %zu < required %zu", workspace_bytes, need);

  // Halo-tile kernel for the L2-traffic-bound 3x3 layers (few input channels: the activation patch is loaded once per
  // tile instead of once per tap).  algo: 0 = heuristic, 1 = per-tap kernel, 2 = halo kernel.
  // (packed forward weights, w_mode 0, only through the CTA-pair halo kernel and its 64-column tiles)
  const bool pair_halo_on = !g_debug[11] && !g_debug[5];
  const bool halo_ok = p->dtype == FCN8_BF16 && p->ksize == 3 && p->Cin % 64 == 0 &&
                       (p->w_mode != 0 || (pair_halo_on && p->Cout == 64 && !(p->flags & FCN8_EPI_COLSUM)));
  // (heuristic: up to 128 input channels; the pair kernel's 128-column dgrad tiles also pay off at 256: conv3_1's dgrad)
  const bool use_halo = halo_ok && p->Cout <= 256 &&
                        (p->algo >= 2 || (p->algo == 0 && (p->Cin <= 128 || (pair_halo_on && p->w_mode == 2 &&
                                                                              p->Cout == 128 && p->Cin <= 256))));
  if (p->algo >= 2 && !halo_ok) return fail(FCN8_ERR_UNSUPPORTED, "conv: halo kernel needs bf16, 3x3, w_mode 1/2");
  if (use_halo) {
    pl.lbw = 3;
    pl.lbh = 4;
    pl.lbn = 0;
    pl.tiles_x = (p->W + 7) >> 3;
    pl.tiles_y = (p->H + 15) >> 4;
    pl.tiles_b = p->N;
    pl.m_tiles = pl.tiles_x * pl.tiles_y * pl.tiles_b;
    pl.splits = 1;
  }
  // CTA pairs for the 256-column bf16 tiles of the per-tap kernel (debug key 5 = 1 keeps the single-CTA kernel)
  const bool use_pair = !use_halo && pl.BN == 256 && p->dtype == FCN8_BF16 && !g_debug[5];
  // CTA pairs for the narrow halo tiles too (conv_halo_pair_kernel): K-major weights (dgrad) at N = 64 / 128, MN-major
  // weights (fprop) need N/2 >= 64 columns per CTA; debug key 11 = 1 keeps the single-CTA halo kernel
  const bool halo_pair = use_halo && pair_halo_on && pl.BN <= 128 && (p->w_mode != 1 || pl.BN == 128);
  const int b_rows = (use_pair || halo_pair) ? pl.BN / 2 : pl.BN;   // rows of the weight tile one CTA loads
  TensorMaps3 maps;
  memset(&maps, 0, sizeof(maps));
  const int ktot = p->ksize * p->ksize * p->Cin;
  // segment order: the low-order products first, hi*hi last (ConvGemmArgs::rz_c)
  const void* xs[3] = {p->x, p->x_lo, p->x};
  const void* ws[3] = {p->wp_lo, p->wp, p->wp};
  if (p->nseg == 2) {
    xs[1] = p->x;
  } else if (p->nseg == 1) {
    ws[0] = p->wp;
  }
  const int taps = p->ksize * p->ksize;
  for (int s = 0; s < p->nseg; ++s) {
    if (use_halo)
      rc = encode_act_map(&maps.a[s], xs[s], p->dtype, p->N, p->H, p->W, p->Cin, 16, 18, 1, false, p->x_ld, p->x_sH,
                          p->x_sN);
    else
      rc = encode_act_map(&maps.a[s], xs[s], p->dtype, p->N, p->H, p->W, p->Cin, 1 << pl.lbw, 1 << pl.lbh,
                          1 << pl.lbn, false, p->x_ld, p->x_sH, p->x_sN);
    if (rc) return rc;
    if (p->w_mode == 0)
      rc = encode_w_map(&maps.b[s], ws[s], p->dtype, p->Cout, ktot, b_rows);
    else if (p->w_mode == 1)
      rc = encode_hwio_map(&maps.b[s], ws[s], taps, p->Cin, p->Cout, 64, use_halo && (pl.BN == 64 || halo_pair) ? 3 : 1);
    else
      rc = encode_hwio_map(&maps.b[s], ws[s], taps, p->Cout, p->Cin, b_rows, use_halo && (pl.BN == 64 || halo_pair) ? 3 : 1);
    if (rc) return rc;
  }
  ConvGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.dbg = next_dbg_slot();
  a.out = p->out;
  a.bias = p->bias;
  a.mask_src = p->mask_src;
  a.residual = p->residual;
  a.partial = static_cast<float*>(workspace);
  a.N = p->N;
  a.H = p->H;
  a.W = p->W;
  a.ldc = p->Cout;
  a.taps = p->ksize * p->ksize;
  a.taps_w = p->ksize;
  a.pad = p->ksize / 2;
  a.cblocks = p->Cin / (p->dtype == FCN8_BF16 ? 64 : 32);
  a.nseg = p->nseg;
  a.lbw = pl.lbw;
  a.lbh = pl.lbh;
  a.lbn = pl.lbn;
  a.tiles_x = pl.tiles_x;
  a.tiles_y = pl.tiles_y;
  a.tiles_b = pl.tiles_b;
  a.tiles_n = pl.tiles_n;
  a.splits = pl.splits;
  a.kb_per_split = pl.kb_per_split;
  a.flags = p->flags & (31 | 64 | 128 | 256);
  if (g_debug[3] & 1) a.flags |= EPI_DBG_NOSTORE;
  if (g_debug[3] & 2) a.flags |= EPI_DBG_NOLOAD;
  a.colsum = p->colsum;
  a.colsum_n = p->colsum_n;
  a.seed_ptr = p->seed_ptr;
  a.acc_scale = p->out_scale != 0.f ? p->out_scale : 1.f;
  a.rz_c = 4.f * rz_per_mma();   // per k-block of 128 operand bytes = 4 MMAs
  if (use_halo) a.kb_per_split = pl.total_kb;   // a halo tile accumulates every segment
  const int out_ld = p->out_ld > 0 ? p->out_ld : p->Cout;
  a.osW = out_ld;
  a.osH = p->out_sH > 0 ? p->out_sH : (long long)p->W * out_ld;
  a.osN = p->out_sN > 0 ? p->out_sN : (long long)p->H * a.osH;
  if (p->flags & FCN8_EPI_POOL) {
    const int pld = p->pool_ld > 0 ? p->pool_ld : out_ld;
    a.pool_out = p->pool_out;
    a.pool_out_lo = p->pool_out_lo;
    a.psW = pld;
    a.psH = (long long)(p->W / 2) * pld;
    a.psN = (long long)(p->H / 2) * a.psH;
  }
  a.b_mode = p->w_mode == 1 ? 2 : (p->w_mode == 2 ? 1 : 0);
  a.out_lo = p->out_lo;
  a.residual_lo = p->residual_lo;
  a.mask_scale = p->mask_scale;
  a.seed = p->seed;
  if (p->flags & FCN8_EPI_DROPOUT) {
    if (!(p->keep_prob > 0.f && p->keep_prob <= 1.f)) return fail(FCN8_ERR_BAD_SHAPE, "conv: keep_prob out of (0,1]");
    a.inv_keep = 1.f / p->keep_prob;
    a.keep_threshold = (uint32_t)((double)p->keep_prob * 16777216.0);
  }
  // promoted accumulation for the error-compensated products whose hi*hi segment is longer than one chunk
  // (debug key 9: > 0 = chunk length in k-blocks, < 0 = off)
  const int promo_kb = g_debug[9] < 0 ? 0 : (g_debug[9] > 0 ? g_debug[9] : kPromoKbDefault);
  // (forward only: the 1e-4 bar is on logits and loss; the dgrad sums are mixed-sign and their bar is 1e-2)
  const bool promo = !use_halo && p->dtype == FCN8_BF16 && p->nseg == 3 && promo_kb > 0 && p->w_mode != 2 &&
                     (pl.total_kb / p->nseg > promo_kb) && (pl.kb_per_split > promo_kb);
  a.promo_kb = promo ? promo_kb : 0;
  ConvGemmArgs kernel_args = a;
  if (pl.splits > 1) kernel_args.flags = EPI_PARTIAL;
  if (use_halo) {
    const long long tiles = (long long)pl.m_tiles * pl.tiles_n;
    a.dyn = g_debug[10] ? 1 : 0;   // dynamic tile scheduling, see below
    const int hgrid = (int)((a.dyn || tiles < num_sms()) ? tiles : num_sms());
    if ((long long)pl.tiles_n * pl.BN > 256) return fail(FCN8_ERR_UNSUPPORTED, "conv: halo kernel needs Cout <= 256");
    if (halo_pair) {
      const long long units = (long long)((pl.m_tiles + 1) / 2) * pl.tiles_n;
      const int pairs = (int)((a.dyn || units < num_sms() / 2) ? units : num_sms() / 2);
      // the CTA's half of all nine taps of every (set, channel block) stays resident when it fits: 6 groups of 12 KB
      const bool res = pl.BN == 64 && pl.tiles_n == 1 && p->w_mode != 1 && (p->nseg >= 2 ? 2 : 1) * (p->Cin / 64) <= 2 &&
                       !g_debug[2];
      cudaError_t pe = pl.BN == 128 ? launch_halo_pair_t<128>(maps, a, 2 * pairs, (cudaStream_t)stream)
                       : res        ? launch_halo_pair_t<64, true>(maps, a, 2 * pairs, (cudaStream_t)stream)
                                    : launch_halo_pair_t<64>(maps, a, 2 * pairs, (cudaStream_t)stream);
      return pe == cudaSuccess ? 0 : cuda_fail(pe, "conv_halo_pair launch");
    }
    // weights resident in shared memory when the CTA's whole slice is one group of 9 taps (conv1_2 fwd / dgrad, bf16)
    const bool resident = pl.BN == 64 && pl.tiles_n == 1 && (p->nseg >= 2 ? 2 : 1) * (p->Cin / 64) <= 1 && !g_debug[2];
    cudaError_t he = pl.BN == 256 ? launch_halo_t<256>(maps, a, hgrid, (cudaStream_t)stream)
                   : pl.BN == 128 ? launch_halo_t<128>(maps, a, hgrid, (cudaStream_t)stream)
                   : resident     ? launch_halo_t<64, true>(maps, a, hgrid, (cudaStream_t)stream)
                                  : launch_halo_t<64>(maps, a, hgrid, (cudaStream_t)stream);
    return he == cudaSuccess ? 0 : cuda_fail(he, "conv_halo launch");
  }
  // dynamic tile scheduling (ConvGemmArgs::dyn): one CTA (pair) per tile in the grid, the resident ones take over the
  // tiles of those not yet launched; switched on with debug key 10 = 1 (the engine does when it runs data parallel); default: static round-robin over
  // min(tiles, #SMs) CTAs, which is 2-3 % faster for the short tiles of the halo kernels when nothing else holds SMs
  const bool dyn = g_debug[10] != 0;
  kernel_args.dyn = dyn ? 1 : 0;
  const long long total_tiles = (long long)pl.m_tiles * pl.tiles_n * pl.splits;
  const int grid = (int)((dyn || total_tiles < num_sms()) ? total_tiles : num_sms());
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  const bool tf32 = p->dtype == FCN8_F32;
  if (use_pair) {
    const long long units = (long long)((pl.m_tiles + 1) / 2) * pl.tiles_n * pl.splits;
    const int pairs = (int)((dyn || units < num_sms() / 2) ? units : num_sms() / 2);
    e = promo ? launch_conv_pair<true>(maps, kernel_args, 2 * pairs, st)
              : launch_conv_pair<false>(maps, kernel_args, 2 * pairs, st);
  } else
#define FCN8_DISPATCH(BNV)                                                    \
  e = tf32    ? launch_conv_t<BNV, true>(maps, kernel_args, grid, st)         \
      : promo ? launch_conv_t<BNV, false, true>(maps, kernel_args, grid, st)  \
              : launch_conv_t<BNV, false>(maps, kernel_args, grid, st)
  if (pl.BN == 256) {
    FCN8_DISPATCH(256);
  } else if (pl.BN == 128) {
    FCN8_DISPATCH(128);
  } else {
    FCN8_DISPATCH(64);
  }
#undef FCN8_DISPATCH
  if (e != cudaSuccess) return cuda_fail(e, "conv_gemm launch");
  if (pl.splits > 1) {
    e = launch_conv_splitk_reduce(static_cast<const float*>(workspace), pl.splits, pl.out_elems, p->Cout, a, p->dtype,
                                  st);
    if (e != cudaSuccess) return cuda_fail(e, "conv split-K reduce launch");
  }
  return 0;
}

namespace {
// wgrad_halo_kernel: 3x3, 64 input channels, bf16 operands (conv1_2, conv2_1)
bool wgrad_use_halo(const Fcn8WgradParams* p) {
  return p->dtype == FCN8_BF16 && p->ksize == 3 && p->Cin == 64 && p->Cout % 64 == 0 && p->rows_valid <= 0 &&
         p->force_splits <= 0 && p->force_bn <= 0 && !g_debug[1] && p->out_cols <= 0;
}
// score heads: the class dimension is padded to 64 columns, only the first out_cols go to dw [rows][out_cols]
bool wgrad_clip(const Fcn8WgradParams* p) {
  return (p->out_cols > 0 && p->out_cols < p->Cout) || (p->out_scale != 0.f && p->out_scale != 1.f);
}
struct WgradHaloPlan {
  int tiles_x, tiles_y, total_patches, tiles_n, splits, patches_per_split;
};
void plan_wgrad_halo(const Fcn8WgradParams* p, WgradHaloPlan* pl) {
  pl->tiles_x = (p->W + 7) >> 3;
  pl->tiles_y = (p->H + 15) >> 4;
  pl->total_patches = pl->tiles_x * pl->tiles_y * p->N;
  pl->tiles_n = p->Cout / 64;
  int splits = num_sms() / pl->tiles_n;
  if (splits > pl->total_patches) splits = pl->total_patches;
  if (splits < 1) splits = 1;
  pl->patches_per_split = (pl->total_patches + splits - 1) / splits;
  pl->splits = (pl->total_patches + pl->patches_per_split - 1) / pl->patches_per_split;
}
}  // namespace

size_t fcn8_wgrad_gemm_workspace_bytes(const Fcn8WgradParams* p) {
  if (wgrad_use_halo(p)) {
    WgradHaloPlan hp;
    plan_wgrad_halo(p, &hp);
    return (size_t)hp.splits * 640 * p->Cout * sizeof(float);
  }
  WgradPlan pl;
  if (plan_wgrad(p, &pl)) return 0;
  return (pl.splits > 1 || wgrad_clip(p)) ? (size_t)pl.splits * pl.rows_pad * p->Cout * sizeof(float) : 0;
}

int32_t fcn8_wgrad_gemm(const Fcn8WgradParams* p, void* workspace, size_t workspace_bytes, void* stream) {
  if (!p || !p->x || !p->dy || !p->dw) return fail(FCN8_ERR_BAD_SHAPE, "wgrad: null pointer");
  if (!aligned16(p->x) || !aligned16(p->dy) || !aligned16(p->dw))
    return fail(FCN8_ERR_BAD_ALIGN, "wgrad: pointers must be 16-byte aligned");
  if ((p->nseg == 3 && !p->x_lo) || (p->nseg >= 2 && !p->dy_lo))
    return fail(FCN8_ERR_BAD_SHAPE, "wgrad: nseg=3 needs x_lo and dy_lo (nseg=2: dy_lo)");
  if (p->nseg < 1 || p->nseg > 3) return fail(FCN8_ERR_UNSUPPORTED, "wgrad: nseg must be 1, 2 or 3");
  if (wgrad_use_halo(p)) {
    if (p->N <= 0 || p->H <= 0 || p->W <= 0) return fail(FCN8_ERR_BAD_SHAPE, "wgrad: empty tensor");
    WgradHaloPlan hp;
    plan_wgrad_halo(p, &hp);
    const size_t hneed = (size_t)hp.splits * 640 * p->Cout * sizeof(float);
    if (hneed > workspace_bytes || !workspace)
      return fail(FCN8_ERR_WORKSPACE, "wgrad: workspace %zu < required %zu", workspace_bytes, hneed);
    TensorMaps3 hm;
    memset(&hm, 0, sizeof(hm));
    const void* hxs[3] = {p->x, p->x_lo, p->x};      // low-order products first (ConvGemmArgs::rz_c)
    const void* hds[3] = {p->dy_lo, p->dy, p->dy};
    if (p->nseg == 2) hxs[1] = p->x;
    if (p->nseg == 1) hds[0] = p->dy;
    for (int s = 0; s < p->nseg; ++s) {
      int hrc = encode_act_map(&hm.a[s], hxs[s], FCN8_BF16, p->N, p->H, p->W, 64, 16, 18, 1, false, p->x_ld, p->x_sH,
                               p->x_sN);
      if (hrc) return hrc;
      hrc = encode_act_map(&hm.b[s], hds[s], FCN8_BF16, p->N, p->H, p->W, p->Cout, 8, 16, 1, false, p->dy_ld, p->dy_sH,
                           p->dy_sN);
      if (hrc) return hrc;
    }
    WgradHaloArgs ha;
    memset(&ha, 0, sizeof(ha));
    ha.partial = static_cast<float*>(workspace);
    ha.N = p->N;
    ha.H = p->H;
    ha.W = p->W;
    ha.ldc = p->Cout;
    ha.nseg = p->nseg;
    ha.tiles_x = hp.tiles_x;
    ha.tiles_y = hp.tiles_y;
    ha.total_patches = hp.total_patches;
    ha.patches_per_split = hp.patches_per_split;
    ha.splits = hp.splits;
    ha.tiles_n = hp.tiles_n;
    ha.acc_scale = 1.f + 8.f * (float)hp.patches_per_split * rz_per_mma();   // the hi*hi segment's MMAs
    static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[cur_dev()];
    if (!attr_done) {
      cudaError_t ae = cudaFuncSetAttribute(wgrad_halo_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            WgradHaloCfg::kSmemBytes);
      if (ae != cudaSuccess) return cuda_fail(ae, "wgrad_halo attribute");
      attr_done = true;
    }
    cudaStream_t hst = (cudaStream_t)stream;
    (void)launch_k(wgrad_halo_kernel<64>, dim3(hp.tiles_n * hp.splits), dim3(kGemmThreads), WgradHaloCfg::kSmemBytes, hst, hm, ha);
    cudaError_t he = cudaGetLastError();
    if (he != cudaSuccess) return cuda_fail(he, "wgrad_halo launch");
    he = launch_wgrad_splitk_reduce(static_cast<const float*>(workspace), p->dw, hp.splits, 640, 576, p->Cout, p->Cout,
                                    1.f, hst);
    return he == cudaSuccess ? 0 : cuda_fail(he, "wgrad_halo reduce launch");
  }
  WgradPlan pl;
  int rc = plan_wgrad(p, &pl);
  if (rc) return rc;
  const bool via_partial = pl.splits > 1 || wgrad_clip(p);   // clipped / scaled outputs always go through the reduce
  const size_t need = via_partial ? (size_t)pl.splits * pl.rows_pad * p->Cout * sizeof(float) : 0;
  if (need > workspace_bytes || (need && !workspace))
    return fail(FCN8_ERR_WORKSPACE, "wgrad: workspace %zu < required %zu", workspace_bytes, need);
  const int rows_all = p->ksize * p->ksize * p->Cin;
  const int rows_valid = (p->rows_valid > 0 && p->rows_valid < rows_all) ? p->rows_valid : rows_all;

  TensorMaps3 maps;
  memset(&maps, 0, sizeof(maps));
  const void* xs[3] = {p->x, p->x_lo, p->x};        // low-order products first (ConvGemmArgs::rz_c)
  const void* ds[3] = {p->dy_lo, p->dy, p->dy};
  if (p->nseg == 2) xs[1] = p->x;
  if (p->nseg == 1) ds[0] = p->dy;
  for (int s = 0; s < p->nseg; ++s) {
    const bool atom32 = p->dtype == FCN8_F32;  // MN-major tf32 operands need the 32-byte-granule swizzle
    rc = encode_act_map(&maps.a[s], xs[s], p->dtype, p->N, p->H, p->W, p->Cin, 1 << pl.lbw, 1 << pl.lbh, 1 << pl.lbn,
                        atom32, p->x_ld, p->x_sH, p->x_sN);
    if (rc) return rc;
    rc = encode_act_map(&maps.b[s], ds[s], p->dtype, p->N, p->H, p->W, p->Cout, 1 << pl.lbw, 1 << pl.lbh,
                        1 << pl.lbn, atom32, p->dy_ld, p->dy_sH, p->dy_sN);
    if (rc) return rc;
  }
  WgradArgs a;
  memset(&a, 0, sizeof(a));
  a.dbg = next_dbg_slot();
  a.out = p->dw;
  a.partial = static_cast<float*>(workspace);
  a.N = p->N;
  a.H = p->H;
  a.W = p->W;
  a.Cin = p->Cin;
  a.ldc = p->Cout;
  a.taps = p->ksize * p->ksize;
  a.taps_w = p->ksize;
  a.pad = p->ksize / 2;
  a.nseg = p->nseg;
  a.rows_valid = rows_valid;
  a.total_chunks = pl.total_chunks;
  a.m_tiles = pl.m_tiles;
  a.tiles_n = pl.tiles_n;
  a.lbw = pl.lbw;
  a.lbh = pl.lbh;
  a.lbn = pl.lbn;
  a.pb_x = pl.pb_x;
  a.pb_y = pl.pb_y;
  a.pb_b = pl.pb_b;
  a.splits = pl.splits;
  a.pb_per_split = pl.pb_per_split;
  a.flags = via_partial ? EPI_PARTIAL : 0;
  {
    // the hi*hi segment's share of a split's MMAs truncates at full magnitude (on average 1 / nseg of them)
    const int pix = 1 << (pl.lbw + pl.lbh + pl.lbn);
    const int mma_per_pb = pix / (p->dtype == FCN8_BF16 ? 16 : 8);
    a.acc_scale = 1.f + (float)mma_per_pb * (float)pl.pb_per_split / (float)p->nseg * rz_per_mma();
  }
  const long long total_tiles = (long long)pl.m_tiles * pl.tiles_n * pl.splits;
  a.dyn = g_debug[10] ? 1 : 0;   // dynamic tile scheduling (ConvGemmArgs::dyn)
  const int grid = (int)((a.dyn || total_tiles < num_sms()) ? total_tiles : num_sms());
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  const bool tf32 = p->dtype == FCN8_F32;
#define FCN8_DISPATCH(BNV) \
  e = tf32 ? launch_wgrad_t<BNV, true>(maps, a, grid, st) : launch_wgrad_t<BNV, false>(maps, a, grid, st)
  if (pl.pair) {
    const long long units = (long long)((pl.m_tiles + 1) / 2) * pl.tiles_n * pl.splits;
    const int pairs = (int)((a.dyn || units < num_sms() / 2) ? units : num_sms() / 2);
    e = launch_wgrad_pair(maps, a, 2 * pairs, st);
  } else if (pl.BN == 256) {
    FCN8_DISPATCH(256);
  } else if (pl.BN == 128) {
    FCN8_DISPATCH(128);
  } else {
    FCN8_DISPATCH(64);
  }
#undef FCN8_DISPATCH
  if (e != cudaSuccess) return cuda_fail(e, "wgrad_gemm launch");
  if (via_partial) {
    e = launch_wgrad_splitk_reduce(static_cast<const float*>(workspace), p->dw, pl.splits, pl.rows_pad, rows_valid,
                                   p->Cout, p->out_cols > 0 ? p->out_cols : p->Cout,
                                   p->out_scale != 0.f ? p->out_scale : 1.f, st);
    if (e != cudaSuccess) return cuda_fail(e, "wgrad split-K reduce launch");
  }
  return 0;
}

int32_t fcn8_pack_weights(const Fcn8PackParams* p, void* stream) {
  if (!p || !p->w || !p->out) return fail(FCN8_ERR_BAD_SHAPE, "pack: null pointer");
  if (p->mode != 0 && p->mode != 1) return fail(FCN8_ERR_BAD_SHAPE, "pack: mode must be 0 or 1");
  if (p->mode == 0 && p->CinPad < p->Cin) return fail(FCN8_ERR_BAD_SHAPE, "pack: CinPad < Cin");
  cudaError_t e = launch_pack(p->w, p->out, p->out_lo, p->ksize, p->Cin, p->Cout, p->CinPad, p->mode, p->dtype,
                              (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "pack launch");
}

int32_t fcn8_split_tf32(const float* x, float* hi, float* lo, size_t n, void* stream) {
  if (!x || (!hi && !lo)) return fail(FCN8_ERR_BAD_SHAPE, "split_tf32: null pointer");
  if (n % 4) return fail(FCN8_ERR_BAD_SHAPE, "split_tf32: n must be a multiple of 4");
  if (!aligned16(x) || (hi && !aligned16(hi)) || (lo && !aligned16(lo)))
    return fail(FCN8_ERR_BAD_ALIGN, "split_tf32: alignment");
  cudaError_t e = launch_split_tf32(x, hi, lo, n, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "split_tf32 launch");
}

static int check_pool(const Fcn8PoolParams* p) {
  if (!p || !p->x || !p->y) return fail(FCN8_ERR_BAD_SHAPE, "pool: null pointer");
  const int vec = p->dtype == FCN8_F32 ? 4 : 8;
  if (p->C % vec) return fail(FCN8_ERR_BAD_SHAPE, "pool: C=%d must be a multiple of %d", p->C, vec);
  if (p->N <= 0 || p->H <= 0 || p->W <= 0) return fail(FCN8_ERR_BAD_SHAPE, "pool: empty tensor");
  if (!aligned16(p->x) || !aligned16(p->y)) return fail(FCN8_ERR_BAD_ALIGN, "pool: alignment");
  return 0;
}
int32_t fcn8_maxpool_fwd(const Fcn8PoolParams* p, void* stream) {
  int rc = check_pool(p);
  if (rc) return rc;
  cudaError_t e = launch_maxpool_fwd(p->x, p->y, p->N, p->H, p->W, p->C, p->dtype, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "maxpool_fwd launch");
}
int32_t fcn8_maxpool_bwd(const Fcn8PoolParams* p, void* stream) {
  int rc = check_pool(p);
  if (rc) return rc;
  if (!p->dx || !aligned16(p->dx)) return fail(FCN8_ERR_BAD_SHAPE, "pool bwd: dx missing / unaligned");
  if (p->db) {   // a thread keeps the column sums of ONE channel vector: the grid stride must be a multiple of C / vec
    const int CV = p->C / (p->dtype == FCN8_F32 ? 4 : 8);
    if (CV <= 0 || 256 % CV) return fail(FCN8_ERR_UNSUPPORTED, "pool bwd: db needs C/vec (= %d) to divide 256", CV);
  }
  cudaError_t e = launch_maxpool_bwd(p->x, p->y, p->dx, p->db, p->N, p->H, p->W, p->C, p->dtype, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "maxpool_bwd launch");
}

size_t fcn8_bias_grad_workspace_bytes(const Fcn8BiasGradParams* p) {
  return (size_t)bias_grad_blocks(p->P, p->C) * p->C * sizeof(float);
}
int32_t fcn8_bias_grad(const Fcn8BiasGradParams* p, void* workspace, size_t workspace_bytes, void* stream) {
  if (!p || !p->dy || !p->db) return fail(FCN8_ERR_BAD_SHAPE, "bias_grad: null pointer");
  const int vec = p->dtype == FCN8_F32 ? 4 : 8;
  const int CV = p->C / vec;
  if (p->C % vec || (CV & (CV - 1))) return fail(FCN8_ERR_BAD_SHAPE, "bias_grad: C/%d must be a power of two", vec);
  if (workspace_bytes < fcn8_bias_grad_workspace_bytes(p) || !workspace)
    return fail(FCN8_ERR_WORKSPACE, "bias_grad: workspace too small");
  cudaError_t e = launch_bias_grad(p->dy, p->db, p->P, p->C, p->dtype, static_cast<float*>(workspace),
                                   (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "bias_grad launch");
}

int32_t fcn8_confusion_matrix(const int64_t* pred, const uint8_t* labels_onehot, unsigned long long* conf, int64_t P,
                              int32_t C, void* stream) {
  if (!pred || !labels_onehot || !conf) return fail(FCN8_ERR_BAD_SHAPE, "confusion: null pointer");
  if (C < 1 || C > 32) return fail(FCN8_ERR_UNSUPPORTED, "confusion: num_classes=%d not in [1,32]", C);
  cudaError_t e = launch_confusion(reinterpret_cast<const long long*>(pred), labels_onehot, conf, P, C,
                                   (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "confusion launch");
}

int32_t fcn8_set_step_scalars(float* scalars, float lr_t, uint32_t seed, void* stream) {
  if (!scalars) return fail(FCN8_ERR_BAD_SHAPE, "set_step_scalars: null pointer");
  cudaError_t e = launch_set_step_scalars(scalars, lr_t, seed, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "set_step_scalars launch");
}
int32_t fcn8_cast_bf16(const float* x, void* out, size_t n, void* stream) {
  if (!x || !out) return fail(FCN8_ERR_BAD_SHAPE, "cast_bf16: null pointer");
  if (!aligned16(x) || !aligned16(out)) return fail(FCN8_ERR_BAD_ALIGN, "cast_bf16: pointers must be 16-byte aligned");
  cudaError_t e = launch_cast_bf16(x, out, n, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "cast_bf16 launch");
}
int32_t fcn8_adam(float* p, const float* g, float* m, float* v, size_t n, float lr_t, float beta1, float beta2,
                  float eps, float grad_scale, void* w_hi, void* w_lo, const float* lr_ptr, const void* g_bf16,
                  void* stream) {
  if (!p || (!g && !g_bf16) || !m || !v) return fail(FCN8_ERR_BAD_SHAPE, "adam: null pointer");
  if (g_bf16 && (reinterpret_cast<uintptr_t>(g_bf16) & 7u)) return fail(FCN8_ERR_BAD_ALIGN, "adam: g_bf16 alignment");
  if (!g) g = p;   // never read when g_bf16 is set; keeps the alignment check below meaningful
  if (!aligned16(p) || !aligned16(g) || !aligned16(m) || !aligned16(v) || (w_hi && !aligned16(w_hi)) ||
      (w_lo && !aligned16(w_lo)))
    return fail(FCN8_ERR_BAD_ALIGN, "adam: pointers must be 16-byte aligned");
  if (w_lo && !w_hi) return fail(FCN8_ERR_BAD_SHAPE, "adam: w_lo without w_hi");
  cudaError_t e =
      launch_adam(p, g, m, v, n, lr_t, beta1, beta2, eps, grad_scale, w_hi, w_lo, lr_ptr, g_bf16, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "adam launch");
}
int32_t fcn8_shadow_weights(const float* p, void* w_hi, void* w_lo, size_t n, void* stream) {
  if (!p || !w_hi) return fail(FCN8_ERR_BAD_SHAPE, "shadow: null pointer");
  if (!aligned16(p) || !aligned16(w_hi) || (w_lo && !aligned16(w_lo)))
    return fail(FCN8_ERR_BAD_ALIGN, "shadow: pointers must be 16-byte aligned");
  cudaError_t e = launch_shadow(p, w_hi, w_lo, n, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "shadow launch");
}
int32_t fcn8_l2_reg(const float* w, float* g, float* loss_sum, size_t n, float rate, void* stream) {
  if (!w) return fail(FCN8_ERR_BAD_SHAPE, "l2_reg: null pointer");
  cudaError_t e = launch_l2_reg(w, g, loss_sum, n, rate, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "l2_reg launch");
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Decoder on the tensor cores (fcn8s_tensorflow.py:164-235 and its autodiff).  The three transposed convolutions are
// phase GEMMs over bf16 hi / lo planes: for stride s, kernel 2s, pad s/2 every s x s output block (J, I),
// J in [0,h], I in [0,w], depends on the 2x2 inputs (J-1+ty, I-1+tx):
//   Zblock[(dy,dx,co)] = sum_{(ty,tx,ci)} x[J-1+ty, I-1+tx, ci] * T[dy+s(1-ty), dx+s(1-tx), co, ci]
// rows = blocks, K = 4 taps x 64 channels (classes zero-padded to one 128-byte operand row), N = s*s*CP columns.
//   fwd  (s = 2): the block columns are scattered to the dense output in the epilogue, + skip tensor (tf.add :213,224)
//   loss (s = 8): CP = 32, a 256-column tile is one block row of 8 pixels: softmax-CE / gradient / softmax / argmax /
//                 confusion matrix in the epilogue, logits are not written unless asked for
//   dx:  A = dz planes in the padded blocked layout [N, s(h+1), s(w+1), CP] through a 5-D TMA map, K = 4 * s*s*CP
//   dw:  dWbig[(ty,tx,ci), (dy,dx,co)] = sum_blocks Xnbr * dZblock (wgrad kernel, dz as the 5-D B operand), un-permuted
// The 1x1 score heads run through fcn8_conv_gemm / fcn8_wgrad_gemm with the class dimension padded to 64
// (fcn8_head_pack); see INTEGRATION.md.
namespace {
int deconv_cp(int s) { return s == 8 ? 32 : 64; }
int check_deconv(const Fcn8DeconvParams* p, const char* what) {
  if (!p) return fail(FCN8_ERR_BAD_SHAPE, "%s: null params", what);
  if (p->C < 1 || p->C > 32) return fail(FCN8_ERR_UNSUPPORTED, "%s: num_classes=%d not in [1,32]", what, p->C);
  if (p->stride != 2 && p->stride != 8) return fail(FCN8_ERR_UNSUPPORTED, "%s: stride must be 2 or 8", what);
  if (p->N <= 0 || p->h <= 0 || p->wd <= 0) return fail(FCN8_ERR_BAD_SHAPE, "%s: empty tensor", what);
  if (p->nseg < 1 || p->nseg > 3) return fail(FCN8_ERR_BAD_SHAPE, "%s: nseg must be 1, 2 or 3", what);
  return 0;
}
// zp planes [N][s*Hb][s*Wb][CP] bf16 seen as dims (s*CP, Wb, s, Hb, N); box (64, bw, 1, bh, bn) = bw*bh*bn blocks x 128 B
int encode_blocked_map(CUtensorMap* m, const void* ptr, int N, int Hb, int Wb, int s, int CP, int bw, int bh, int bn) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const cuuint64_t row = (cuuint64_t)s * CP * 2;       // bytes of one block row
  const cuuint64_t line = (cuuint64_t)Wb * row;        // bytes of one padded image row
  const cuuint64_t dims[5] = {(cuuint64_t)s * CP, (cuuint64_t)Wb, (cuuint64_t)s, (cuuint64_t)Hb, (cuuint64_t)N};
  const cuuint64_t strides[4] = {row, line, (cuuint64_t)s * line, (cuuint64_t)s * Hb * line};
  const cuuint32_t box[5] = {64, (cuuint32_t)bw, 1, (cuuint32_t)bh, (cuuint32_t)bn};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled(blocked N%d Hb%d Wb%d s%d CP%d) failed: %d", N, Hb, Wb, s, CP,
                (int)r);
  return 0;
}
// operand pointers of segment i in the order low-order products first (ConvGemmArgs::rz_c): (a, b_lo), (a_lo, b), (a, b)
void seg_ptrs(int nseg, const void* a, const void* a_lo, const void* b, const void* b_lo, const void* as[3],
              const void* bs[3]) {
  as[0] = a; as[1] = a_lo; as[2] = a;
  bs[0] = b_lo; bs[1] = b; bs[2] = b;
  if (nseg == 2) as[1] = a;
  if (nseg == 1) bs[0] = b;
}

cudaError_t launch_conv_loss(const TensorMaps3& maps, const ConvGemmArgs& a, int grid, bool pair, cudaStream_t st) {
  static bool attr_done_dev[64][2] = {};
  bool& attr_done = attr_done_dev[cur_dev()][pair ? 1 : 0];
  if (pair) {
    using Cfg = GemmCfg<256, true>;
    if (!attr_done) {
      cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<256, false, true, 1>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
      if (e != cudaSuccess) return e;
      attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl_off ? 1 : 2;
    count_launch();
    return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<256, false, true, 1>, maps, a);
  }
  using Cfg = GemmCfg<256>;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<256, false, false, 1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  (void)launch_k(conv_gemm_kernel<256, false, false, 1>, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, st, maps, a);
  return cudaGetLastError();
}

// common part of fwd / loss: A = x planes (2x2 taps over the (h+1) x (w+1) block grid), B = packed w_fwd
int setup_deconv_fwd(const Fcn8DeconvParams* p, int BN, TensorMaps3* maps, ConvGemmArgs* a) {
  const int s = p->stride, CP = deconv_cp(s);
  const int Hb = p->h + 1, Wb = p->wd + 1;
  const int ncols = s * s * CP;
  int lbw, lbh, lbn;
  choose_patch(p->N, Hb, Wb, 7, &lbw, &lbh, &lbn);
  memset(maps, 0, sizeof(*maps));
  const void* xs[3];
  const void* ws[3];
  seg_ptrs(p->nseg, p->x, p->x_lo, p->w, p->w_lo, xs, ws);
  const int b_rows = BN;
  for (int i = 0; i < p->nseg; ++i) {
    int rc = encode_act_map(&maps->a[i], xs[i], FCN8_BF16, p->N, p->h, p->wd, 64, 1 << lbw, 1 << lbh, 1 << lbn, false,
                            p->x_ld, p->x_sH, p->x_sN);
    if (rc) return rc;
    rc = encode_w_map(&maps->b[i], ws[i], FCN8_BF16, ncols, 256, b_rows);
    if (rc) return rc;
  }
  memset(a, 0, sizeof(*a));
  a->bias = p->bias_big;
  a->N = p->N;
  a->H = Hb;
  a->W = Wb;
  a->ldc = ncols;
  a->taps = 4;
  a->taps_w = 2;
  a->pad = 1;
  a->cblocks = 1;
  a->nseg = p->nseg;
  a->lbw = lbw;
  a->lbh = lbh;
  a->lbn = lbn;
  a->tiles_x = (Wb + (1 << lbw) - 1) >> lbw;
  a->tiles_y = (Hb + (1 << lbh) - 1) >> lbh;
  a->tiles_b = (p->N + (1 << lbn) - 1) >> lbn;
  a->tiles_n = ncols / BN;
  a->splits = 1;
  a->kb_per_split = p->nseg * 4;
  a->acc_scale = 1.f;
  a->rz_c = 4.f * rz_per_mma();
  a->flags = EPI_BIAS;
  a->blk_s = s;
  a->blk_cp = CP;
  a->out_H = s * p->h;
  a->out_W = s * p->wd;
  return 0;
}
}  // namespace

extern "C" {

int32_t fcn8_deconv_cp(int32_t stride) { return deconv_cp(stride); }

int32_t fcn8_deconv_pack(const float* T, const float* bias, int32_t C, int32_t stride, void* w_fwd, void* w_fwd_lo,
                         void* w_dx, void* w_dx_lo, float* bias_big, void* stream) {
  if (!T || !bias || !w_fwd || !w_dx || !bias_big) return fail(FCN8_ERR_BAD_SHAPE, "deconv pack: null pointer");
  if (C < 1 || C > 32) return fail(FCN8_ERR_UNSUPPORTED, "deconv pack: num_classes=%d not in [1,32]", C);
  if (stride != 2 && stride != 8) return fail(FCN8_ERR_UNSUPPORTED, "deconv pack: stride must be 2 or 8");
  cudaError_t e = launch_deconv_pack(T, bias, w_fwd, w_fwd_lo, w_dx, w_dx_lo, bias_big, C, deconv_cp(stride), stride,
                                     (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "deconv pack launch");
}

int32_t fcn8_head_pack(const float* K, const float* bias, int32_t Cin, int32_t C, void* w_hi, void* w_lo, float* bias64,
                       void* stream) {
  if (!K || !w_hi) return fail(FCN8_ERR_BAD_SHAPE, "head pack: null pointer");
  if (C < 1 || C > 32) return fail(FCN8_ERR_UNSUPPORTED, "head pack: num_classes=%d not in [1,32]", C);
  cudaError_t e = launch_head_pack(K, bias, Cin, C, w_hi, w_lo, bias64, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "head pack launch");
}

// stride-2 stages (upscore2, upscore_pool4): dense output planes + skip tensor in the epilogue
int32_t fcn8_deconv_fwd(const Fcn8DeconvParams* p, void* stream) {
  int rc = check_deconv(p, "deconv_fwd");
  if (rc) return rc;
  if (!p->x || !p->w || !p->bias_big || !p->out) return fail(FCN8_ERR_BAD_SHAPE, "deconv_fwd: null pointer");
  if ((p->nseg == 3 && !p->x_lo) || (p->nseg >= 2 && !p->w_lo)) return fail(FCN8_ERR_BAD_SHAPE, "deconv_fwd: lo operands");
  const int s = p->stride, CP = deconv_cp(s);
  const int ncols = s * s * CP;
  const int BN = ncols % 256 == 0 ? 256 : 128;
  TensorMaps3 maps;
  ConvGemmArgs a;
  rc = setup_deconv_fwd(p, BN, &maps, &a);
  if (rc) return rc;
  a.out = p->out;
  a.out_lo = p->out_lo;
  a.out_mode = 2;
  const int out_ld = p->out_ld > 0 ? p->out_ld : CP;
  a.osW = out_ld;
  a.osH = p->out_sH > 0 ? p->out_sH : (long long)a.out_W * out_ld;
  a.osN = p->out_sN > 0 ? p->out_sN : (long long)a.out_H * a.osH;
  if (p->skip) {
    a.flags |= EPI_RESIDUAL;
    a.residual = p->skip;
    a.residual_lo = p->skip_lo;
  }
  const long long tiles = (long long)a.tiles_x * a.tiles_y * a.tiles_b * a.tiles_n;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  cudaError_t e = dispatch_conv(BN, false, maps, a, grid, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "deconv_fwd launch");
}

// stride-8 stage (upscore8) with the loss / predictor fused into its epilogue
int32_t fcn8_deconv_loss(const Fcn8DeconvParams* p, void* stream) {
  int rc = check_deconv(p, "deconv_loss");
  if (rc) return rc;
  if (p->stride != 8) return fail(FCN8_ERR_UNSUPPORTED, "deconv_loss: stride must be 8");
  if (!p->x || !p->w || !p->bias_big) return fail(FCN8_ERR_BAD_SHAPE, "deconv_loss: null pointer");
  if ((p->nseg == 3 && !p->x_lo) || (p->nseg >= 2 && !p->w_lo)) return fail(FCN8_ERR_BAD_SHAPE, "deconv_loss: lo operands");
  if ((p->loss_sum || p->dz_hi_out || p->conf) && !p->labels) return fail(FCN8_ERR_BAD_SHAPE, "deconv_loss: needs labels");
  if (!p->loss_sum && !p->dz_hi_out && !p->conf && !p->logits && !p->softmax && !p->argmax && !p->argmax_u8)
    return fail(FCN8_ERR_BAD_SHAPE, "deconv_loss: no output requested");
  if ((p->dz_hi_out && !aligned16(p->dz_hi_out)) || (p->dz_lo_out && !aligned16(p->dz_lo_out)))
    return fail(FCN8_ERR_BAD_ALIGN, "deconv_loss: dz planes must be 16-byte aligned");
  if ((p->C & 3) == 0 && ((p->logits && !aligned16(p->logits)) || (p->softmax && !aligned16(p->softmax))))
    return fail(FCN8_ERR_BAD_ALIGN, "deconv_loss: logits / softmax must be 16-byte aligned (128-bit stores)");
  TensorMaps3 maps;
  ConvGemmArgs a;
  rc = setup_deconv_fwd(p, 256, &maps, &a);
  if (rc) return rc;
  a.flags = 0;     // the loss epilogue adds the bias itself
  a.labels = p->labels;
  a.loss_sum = p->loss_sum;
  a.dbias = p->dbias;
  a.dz_hi = static_cast<__nv_bfloat16*>(p->dz_hi_out);
  a.dz_lo = static_cast<__nv_bfloat16*>(p->dz_lo_out);
  a.logits = p->logits;
  a.softmax = p->softmax;
  a.argmax = reinterpret_cast<long long*>(p->argmax);
  a.argmax_u8 = p->argmax_u8;
  a.conf = reinterpret_cast<unsigned long long*>(p->conf);
  a.num_classes = p->C;
  a.gscale = p->grad_scale;
  const bool pair = !g_debug[5];
  const int m_tiles = a.tiles_x * a.tiles_y * a.tiles_b;
  cudaError_t e;
  if (pair) {
    const long long units = (long long)((m_tiles + 1) / 2) * a.tiles_n;
    const int pairs = (int)(units < num_sms() / 2 ? units : num_sms() / 2);
    // each CTA of a pair loads half of the 256-row weight tile
    for (int i = 0; i < p->nseg; ++i) {
      const void* xs[3];
      const void* ws[3];
      seg_ptrs(p->nseg, p->x, p->x_lo, p->w, p->w_lo, xs, ws);
      rc = encode_w_map(&maps.b[i], ws[i], FCN8_BF16, 8 * 8 * 32, 256, 128);
      if (rc) return rc;
    }
    e = launch_conv_loss(maps, a, 2 * pairs, true, (cudaStream_t)stream);
  } else {
    const long long tiles = (long long)m_tiles * a.tiles_n;
    e = launch_conv_loss(maps, a, (int)(tiles < num_sms() ? tiles : num_sms()), false, (cudaStream_t)stream);
  }
  return e == cudaSuccess ? 0 : cuda_fail(e, "deconv_loss launch");
}

// dx[n,i,j,ci] = sum_{ty,tx} sum_{dy,dx,co} dz[n, s(i+1-ty)+dy, s(j+1-tx)+dx, co] * T[dy+s(1-ty), dx+s(1-tx), co, ci]
int32_t fcn8_deconv_dx(const Fcn8DeconvParams* p, void* stream) {
  int rc = check_deconv(p, "deconv_dx");
  if (rc) return rc;
  if (!p->dz || !p->w || !p->out) return fail(FCN8_ERR_BAD_SHAPE, "deconv_dx: null pointer");
  if ((p->nseg == 3 && !p->dz_lo) || (p->nseg >= 2 && !p->w_lo)) return fail(FCN8_ERR_BAD_SHAPE, "deconv_dx: lo operands");
  const int s = p->stride, CP = deconv_cp(s);
  const int Hb = p->h + 1, Wb = p->wd + 1;
  const int chunks_row = s * CP / 64;
  const int cblocks = s * chunks_row;  // k-blocks per tap
  int lbw, lbh, lbn;
  choose_patch(p->N, p->h, p->wd, 7, &lbw, &lbh, &lbn);
  TensorMaps3 maps;
  memset(&maps, 0, sizeof(maps));
  const void* zs[3];
  const void* ws[3];
  seg_ptrs(p->nseg, p->dz, p->dz_lo, p->w, p->w_lo, zs, ws);
  for (int i = 0; i < p->nseg; ++i) {
    rc = encode_blocked_map(&maps.a[i], zs[i], p->N, Hb, Wb, s, CP, 1 << lbw, 1 << lbh, 1 << lbn);
    if (rc) return rc;
    rc = encode_w_map(&maps.b[i], ws[i], FCN8_BF16, 64, 4 * s * s * CP, 64);
    if (rc) return rc;
  }
  ConvGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.out = p->out;
  a.out_lo = p->out_lo;
  a.N = p->N;
  a.H = p->h;
  a.W = p->wd;
  a.ldc = 64;
  a.taps = 4;
  a.taps_w = 2;
  a.pad = 1;
  a.cblocks = cblocks;
  a.nseg = p->nseg;
  a.lbw = lbw;
  a.lbh = lbh;
  a.lbn = lbn;
  a.tiles_x = (p->wd + (1 << lbw) - 1) >> lbw;
  a.tiles_y = (p->h + (1 << lbh) - 1) >> lbh;
  a.tiles_b = (p->N + (1 << lbn) - 1) >> lbn;
  a.tiles_n = 1;
  a.splits = 1;
  a.kb_per_split = p->nseg * 4 * cblocks;
  a.acc_scale = 1.f;
  a.rz_c = 4.f * rz_per_mma();
  a.flags = 0;
  a.a_mode = 1;
  a.blk_chunks = chunks_row;
  const int out_ld = p->out_ld > 0 ? p->out_ld : 64;
  a.osW = out_ld;
  a.osH = p->out_sH > 0 ? p->out_sH : (long long)p->wd * out_ld;
  a.osN = p->out_sN > 0 ? p->out_sN : (long long)p->h * a.osH;
  if (p->colsum) {
    a.flags |= EPI_COLSUM;
    a.colsum = p->colsum;
    a.colsum_n = p->colsum_n > 0 ? p->colsum_n : p->C;
  }
  const long long tiles = (long long)a.tiles_x * a.tiles_y * a.tiles_b;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  cudaError_t e = dispatch_conv(64, false, maps, a, grid, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "deconv_dx launch");
}

namespace {
struct DeconvDwPlan {
  int BN, splits, pb_per_split, total_pb, lbw, lbh, lbn, pb_x, pb_y, pb_b, tiles_n, ncols;
};
void plan_deconv_dw(const Fcn8DeconvParams* p, DeconvDwPlan* pl) {
  const int s = p->stride, CP = deconv_cp(s);
  pl->ncols = s * s * CP;
  pl->BN = 128;
  pl->tiles_n = pl->ncols / pl->BN;
  choose_patch(p->N, p->h + 1, p->wd + 1, 7, &pl->lbw, &pl->lbh, &pl->lbn);
  pl->pb_x = (p->wd + 1 + (1 << pl->lbw) - 1) >> pl->lbw;
  pl->pb_y = (p->h + 1 + (1 << pl->lbh) - 1) >> pl->lbh;
  pl->pb_b = (p->N + (1 << pl->lbn) - 1) >> pl->lbn;
  pl->total_pb = p->nseg * pl->pb_x * pl->pb_y * pl->pb_b;
  int splits = num_sms() / (2 * pl->tiles_n);   // two 128-row M tiles (4 taps x 64 channels) per N tile
  const int max_by_k = pl->total_pb / 4 > 0 ? pl->total_pb / 4 : 1;
  if (splits > max_by_k) splits = max_by_k;
  if (splits > 64) splits = 64;
  if (splits < 1) splits = 1;
  pl->pb_per_split = (pl->total_pb + splits - 1) / splits;
  pl->splits = (pl->total_pb + pl->pb_per_split - 1) / pl->pb_per_split;
}
}  // namespace

size_t fcn8_deconv_dw_workspace_bytes(const Fcn8DeconvParams* p) {
  if (check_deconv(p, "deconv_dw")) return 0;
  DeconvDwPlan pl;
  plan_deconv_dw(p, &pl);
  return (size_t)pl.splits * 256 * pl.ncols * sizeof(float);
}

// dT[a,b,co,ci] = sum_{n,i,j} x[n,i,j,ci] * dz[n, s*i+a-p, s*j+b-p, co] as the phase GEMM
// dWbig[(ty,tx,ci), (dy,dx,co)] = sum_blocks Xnbr * dZblock, un-permuted into the TF layout by the reduce kernel.
int32_t fcn8_deconv_dw(const Fcn8DeconvParams* p, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_deconv(p, "deconv_dw");
  if (rc) return rc;
  if (!p->x || !p->dz || !p->dT) return fail(FCN8_ERR_BAD_SHAPE, "deconv_dw: null pointer");
  if ((p->nseg == 3 && !p->x_lo) || (p->nseg >= 2 && !p->dz_lo)) return fail(FCN8_ERR_BAD_SHAPE, "deconv_dw: lo operands");
  const size_t need = fcn8_deconv_dw_workspace_bytes(p);
  if (workspace_bytes < need || !workspace) return fail(FCN8_ERR_WORKSPACE, "deconv_dw: workspace too small");
  const int s = p->stride, CP = deconv_cp(s);
  DeconvDwPlan pl;
  plan_deconv_dw(p, &pl);
  TensorMaps3 maps;
  memset(&maps, 0, sizeof(maps));
  const void* xs[3];
  const void* zs[3];
  seg_ptrs(p->nseg, p->x, p->x_lo, p->dz, p->dz_lo, xs, zs);
  for (int i = 0; i < p->nseg; ++i) {
    rc = encode_act_map(&maps.a[i], xs[i], FCN8_BF16, p->N, p->h, p->wd, 64, 1 << pl.lbw, 1 << pl.lbh, 1 << pl.lbn, false,
                        p->x_ld, p->x_sH, p->x_sN);
    if (rc) return rc;
    rc = encode_blocked_map(&maps.b[i], zs[i], p->N, p->h + 1, p->wd + 1, s, CP, 1 << pl.lbw, 1 << pl.lbh, 1 << pl.lbn);
    if (rc) return rc;
  }
  float* partial = static_cast<float*>(workspace);               // [splits][256][ncols]
  WgradArgs a;
  memset(&a, 0, sizeof(a));
  a.out = partial;
  a.partial = partial;
  a.N = p->N;
  a.H = p->h + 1;
  a.W = p->wd + 1;
  a.Cin = 64;
  a.ldc = pl.ncols;
  a.taps = 4;
  a.taps_w = 2;
  a.pad = 1;
  a.nseg = p->nseg;
  a.rows_valid = 256;
  a.total_chunks = 4;
  a.m_tiles = 2;
  a.tiles_n = pl.tiles_n;
  a.lbw = pl.lbw;
  a.lbh = pl.lbh;
  a.lbn = pl.lbn;
  a.pb_x = pl.pb_x;
  a.pb_y = pl.pb_y;
  a.pb_b = pl.pb_b;
  a.splits = pl.splits;
  a.pb_per_split = pl.pb_per_split;
  a.flags = EPI_PARTIAL;
  a.b_mode = 1;
  a.blk_row = s * CP;
  a.acc_scale = 1.f + 8.f * (float)pl.pb_per_split / (float)p->nseg * rz_per_mma();   // 128 blocks / 16 per bf16 MMA
  const long long tiles = (long long)a.m_tiles * pl.tiles_n * pl.splits;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = dispatch_wgrad(pl.BN, false, maps, a, grid, st);
  if (e != cudaSuccess) return cuda_fail(e, "deconv_dw launch");
  e = launch_deconv_unpack_dw(partial, pl.splits, p->dT, p->C, CP, s, st);
  return e == cudaSuccess ? 0 : cuda_fail(e, "deconv_dw unpack launch");
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// conv1_1 straight from the uint8 image (conv1.cu): the im2col operand exists only in shared memory.
namespace {
// rows [P][ld] bf16 seen as a 2-D map (64 columns, P rows), box (64, 128)
int encode_rows_map(CUtensorMap* m, const void* ptr, long long P, int ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const cuuint64_t dims[2] = {64, (cuuint64_t)P};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FCN8_ERR_CUDA, "cuTensorMapEncodeTiled(rows P %lld ld %d) failed: %d", P, ld, (int)r);
  return 0;
}
int check_conv1(const Fcn8Conv1Params* p, const char* what) {
  if (!p || !p->images) return fail(FCN8_ERR_BAD_SHAPE, "%s: null pointer", what);
  if (p->N <= 0 || p->H <= 0 || p->W <= 0) return fail(FCN8_ERR_BAD_SHAPE, "%s: empty tensor", what);
  if (((long long)p->N * p->H * p->W) % 128) return fail(FCN8_ERR_BAD_SHAPE, "%s: N*H*W must be a multiple of 128", what);
  return 0;
}
int conv1_grid(const Fcn8Conv1Params* p) {
  const long long tiles = (long long)p->N * p->H * p->W / 128;
  return (int)(tiles < num_sms() ? tiles : num_sms());
}
}  // namespace

extern "C" {

int32_t fcn8_conv1_fwd(const Fcn8Conv1Params* p, void* stream) {
  int rc = check_conv1(p, "conv1_fwd");
  if (rc) return rc;
  if (!p->w || !p->bias || !p->out || (p->pair && (!p->w_lo || !p->out_lo)))
    return fail(FCN8_ERR_BAD_SHAPE, "conv1_fwd: null pointer");
  if (!aligned16(p->out) || (p->out_lo && !aligned16(p->out_lo)) || !aligned16(p->w) || !aligned16(p->bias))
    return fail(FCN8_ERR_BAD_ALIGN, "conv1_fwd: pointers must be 16-byte aligned");
  CUtensorMap mh, ml;
  memset(&ml, 0, sizeof(ml));
  rc = encode_w_map(&mh, p->w, FCN8_BF16, 64, 64, 64);
  if (rc) return rc;
  if (p->pair) {
    rc = encode_w_map(&ml, p->w_lo, FCN8_BF16, 64, 64, 64);
    if (rc) return rc;
  } else {
    ml = mh;
  }
  Conv1Args a;
  memset(&a, 0, sizeof(a));
  a.img = p->images;
  a.N = p->N;
  a.H = p->H;
  a.W = p->W;
  a.tiles = (int)((long long)p->N * p->H * p->W / 128);
  a.pair = p->pair ? 1 : 0;
  a.out = static_cast<__nv_bfloat16*>(p->out);
  a.out_lo = static_cast<__nv_bfloat16*>(p->out_lo);
  a.bias = p->bias;
  a.out_ld = p->out_ld > 0 ? p->out_ld : 64;
  a.rz_c = 4.f * rz_per_mma();
  cudaError_t e = launch_conv1_fwd(mh, ml, a, conv1_grid(p), cur_dev(), (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "conv1_fwd launch");
}

size_t fcn8_conv1_wgrad_workspace_bytes(const Fcn8Conv1Params* p) {
  if (check_conv1(p, "conv1_wgrad")) return 0;
  return (size_t)conv1_grid(p) * (p->pair ? 2 : 1) * 4096 * sizeof(float);
}

int32_t fcn8_conv1_wgrad(const Fcn8Conv1Params* p, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_conv1(p, "conv1_wgrad");
  if (rc) return rc;
  if (!p->dy || !p->dw || (p->pair && !p->dy_lo)) return fail(FCN8_ERR_BAD_SHAPE, "conv1_wgrad: null pointer");
  const size_t need = fcn8_conv1_wgrad_workspace_bytes(p);
  if (workspace_bytes < need || !workspace) return fail(FCN8_ERR_WORKSPACE, "conv1_wgrad: workspace too small");
  const long long P = (long long)p->N * p->H * p->W;
  const int ld = p->dy_ld > 0 ? p->dy_ld : 64;
  CUtensorMap mh, ml;
  rc = encode_rows_map(&mh, p->dy, P, ld);
  if (rc) return rc;
  if (p->pair) {
    rc = encode_rows_map(&ml, p->dy_lo, P, ld);
    if (rc) return rc;
  } else {
    ml = mh;
  }
  Conv1Args a;
  memset(&a, 0, sizeof(a));
  a.img = p->images;
  a.N = p->N;
  a.H = p->H;
  a.W = p->W;
  a.tiles = (int)(P / 128);
  a.pair = p->pair ? 1 : 0;
  a.partial = static_cast<float*>(workspace);
  const int grid = conv1_grid(p);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = launch_conv1_wgrad(mh, ml, a, grid, cur_dev(), st);
  if (e != cudaSuccess) return cuda_fail(e, "conv1_wgrad launch");
  // rows 0..26 of every [64][64] partial are the filter taps (kh, kw, c) in TF order
  e = launch_wgrad_splitk_reduce(static_cast<const float*>(workspace), p->dw, grid * (p->pair ? 2 : 1), 64, 27, 64, 64,
                                 1.f, st);
  return e == cudaSuccess ? 0 : cuda_fail(e, "conv1_wgrad reduce launch");
}

}  // extern "C"
