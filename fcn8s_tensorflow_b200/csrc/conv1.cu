// conv1_1 (3 -> 64 channels, 3x3 SAME) straight from the uint8 image: forward and filter gradient on tcgen05 with the
// 27-column im2col operand built in SHARED MEMORY.
//
// Reference sites: the feed of image_input (fcn8s_tensorflow.py:558,686,765), the encoder graph's RGB -> BGR / mean
// subtraction and first convolution [EXT, SURVEY.md A.2], and the filter gradient implied by :257.
//
// Why a dedicated kernel: K = 27.  As an implicit GEMM over a global im2col tensor the layer moved 3.2 GB per c2 step
// (a 64-column bf16 hi/lo im2col written by the feed kernel, read by the forward GEMM and again by the filter
// gradient) for 1.1 GB of compulsory traffic.  Here builder warps read the 3x3x3 uint8 neighbourhood of each pixel,
// subtract the means (zero padding applies AFTER the subtraction: out-of-image taps are zeros, not -mean), convert to
// bf16 (hi / lo) and write the pixel's operand row into shared memory in the 128-byte-swizzled K-major layout the UMMA
// descriptors expect; nothing but the 6 MB image is read for the A operand.
//   tiles   = 128 consecutive pixels of the flattened [N*H*W] pixel list (H, W multiples of 32: no ragged tile)
//   forward : D[128 px, 64 co] = A[128, 64 k] * W[64 co, 64 k]^T, bias + ReLU, bf16 / hi-lo pair store
//   wgrad   : dW[(hi | lo) 64 k, 64 co] += A'[128 px, (hi | lo) 64 k]^T * dY[128 px, 64 co]: the SAME shared-memory tile
//             read MN-major, hi and lo halves side by side as the two 64-row chunks of one M = 128 operand; one TMEM
//             accumulator per CTA over all its tiles, partials reduced by wgrad_splitk_reduce.
// Warp roles (18 warps): 0 = TMA (weights / dY), 1 = TMEM allocator + MMA issuer, 2..9 = epilogue (two warps per TMEM
// lane quarter, 32 columns each; the filter gradient only uses 2..5), 10..17 = builders in two groups of 128 threads
// that take alternate tiles: building a tile is one global-load latency plus ~250 instructions per thread, and the
// layer is bound by its 0.5 GB of output (forward) / dY (gradient) only if two tiles are always under construction.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv1.h"
#include "conv_gemm.cuh"
#include "kernels.h"

namespace fcn8 {

// One pixel's operand row: 27 values (kh, kw, c = B,G,R) as bf16 hi (and lo), chunks 0..3 of the swizzled 128-byte row.
__device__ __forceinline__ void build_row(const Conv1Args& g, long long p, int r, uint8_t* a_hi, uint8_t* a_lo) {
  const int x = static_cast<int>(p % g.W);
  const int y = static_cast<int>((p / g.W) % g.H);
  const long long n = p / (static_cast<long long>(g.W) * g.H);
  float v[32];
#pragma unroll
  for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int yy = y + kh - 1;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int xx = x + kw - 1;
      const bool inb = yy >= 0 && yy < g.H && xx >= 0 && xx < g.W;
      const uint8_t* px = g.img + ((n * g.H + yy) * g.W + xx) * 3;
      // c = 0,1,2 -> B,G,R = RGB byte 2,1,0 minus the ImageNet means
      v[(kh * 3 + kw) * 3 + 0] = inb ? static_cast<float>(__ldg(px + 2)) - 103.939f : 0.f;
      v[(kh * 3 + kw) * 3 + 1] = inb ? static_cast<float>(__ldg(px + 1)) - 116.779f : 0.f;
      v[(kh * 3 + kw) * 3 + 2] = inb ? static_cast<float>(__ldg(px + 0)) - 123.68f : 0.f;
    }
  }
  const int sw = r & 7;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = v[8 * j + 2 * i], b = v[8 * j + 2 * i + 1];
      hi[i] = pack_bf16x2(a, b);
      const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi[i]));
      lo[i] = pack_bf16x2(a - h.x, b - h.y);
    }
    const int off = r * 128 + ((j ^ sw) << 4);
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (a_lo) *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
template <bool PAIR>
__global__ void __launch_bounds__(kC1Threads, 1)
conv1_fwd_kernel(const __grid_constant__ CUtensorMap wmap_hi, const __grid_constant__ CUtensorMap wmap_lo,
                 const Conv1Args g) {
  pdl_launch_dependents();
  constexpr int kStages = 4;
  constexpr int kStageBytes = PAIR ? 2 * kC1Tile : kC1Tile;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_hi = smem;                      // [64 co][128 B]
  uint8_t* w_lo = smem + 8192;
  uint8_t* a_base = smem + 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_base + kStages * kStageBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kStages;
  uint64_t* acc_full = a_empty + kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  uint8_t* store_s = reinterpret_cast<uint8_t*>(bars) + 1024;   // 8 x 2 KB: coalesced epilogue stores (warp_store_frag64)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // columns 27..63 of every operand row are zero for the whole kernel: the builders only ever write chunks 0..3
  for (int i = threadIdx.x; i < kStages * kStageBytes / 16; i += kC1Threads)
    reinterpret_cast<uint4*>(a_base)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&wmap_hi);
    if (PAIR) tma_prefetch_desc(&wmap_lo);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&a_full[s], 128);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  fence_proxy_async_smem();     // the zero fill above is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_full, PAIR ? 16384 : 8192);
      tma_load_2d(&wmap_hi, w_full, w_hi, 0, 0);
      if (PAIR) tma_load_2d(&wmap_lo, w_full, w_lo, 0, 0);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    const uint32_t idesc = make_idesc(1u, 0u, 0u, 128u, 64u);
    const uint64_t desc0 = make_smem_desc_sw128(0, 16, 1024);
    const uint32_t sw_hi = smem_u32(w_hi), sw_lo = smem_u32(w_lo), sa0 = smem_u32(a_base);
    mbar_wait(w_full, 0);
    tc_fence_after();
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int t = blockIdx.x; t < g.tiles; t += gridDim.x) {
      mbar_wait(&acc_empty[as], aphase ^ 1);
      mbar_wait(&a_full[stage], phase);
      tc_fence_after();
      const uint32_t sa = sa0 + stage * kStageBytes;
      const uint32_t d_tmem = tmem_base + as * 64;
      if (elect_one()) {
        uint32_t acc = 0;
        // low-order products first (ConvGemmArgs::rz_c): A_hi * W_lo, A_lo * W_hi, then A_hi * W_hi
        const int nseg = PAIR ? 3 : 1;
#pragma unroll
        for (int seg = 0; seg < nseg; ++seg) {
          const uint32_t a_addr = (PAIR && seg == 1) ? sa + kC1Tile : sa;
          const uint32_t b_addr = (PAIR && seg == 0) ? sw_lo : sw_hi;
          const uint64_t adesc = desc0 | static_cast<uint64_t>((a_addr & 0x3FFFF) >> 4);
          const uint64_t bdesc = desc0 | static_cast<uint64_t>((b_addr & 0x3FFFF) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
            acc = 1;
          }
        }
        umma_commit(&a_empty[stage]);
        umma_commit(&acc_full[as]);
      }
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else if (warp < 10) {
    // ============================== epilogue: bias + ReLU, bf16 / hi-lo pair store ==============================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int c = ((warp - 2) >> 2) * 32;     // this warp's half of the 64 output channels
    const float acc_scale = 1.f + g.rz_c;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < g.tiles; t += gridDim.x) {
      const size_t p = static_cast<size_t>(t) * 128 + row;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * 64 + (static_cast<uint32_t>(quarter * 32) << 16);
      {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(&acc_empty[as]);
        uint32_t hi[16], lo[16];
        const float4* b4 = reinterpret_cast<const float4*>(g.bias + c);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = __ldg(b4 + i);
          const float f0 = fmaxf(__uint_as_float(v[4 * i]) * acc_scale + b.x, 0.f);
          const float f1 = fmaxf(__uint_as_float(v[4 * i + 1]) * acc_scale + b.y, 0.f);
          const float f2 = fmaxf(__uint_as_float(v[4 * i + 2]) * acc_scale + b.z, 0.f);
          const float f3 = fmaxf(__uint_as_float(v[4 * i + 3]) * acc_scale + b.w, 0.f);
          hi[2 * i] = pack_bf16x2(f0, f1);
          hi[2 * i + 1] = pack_bf16x2(f2, f3);
          const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi[2 * i]));
          const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi[2 * i + 1]));
          lo[2 * i] = pack_bf16x2(f0 - h0.x, f1 - h0.y);
          lo[2 * i + 1] = pack_bf16x2(f2 - h1.x, f3 - h1.y);
        }
        uint8_t* stage = store_s + (warp - 2) * kStoreWarpBytes;
        warp_store_frag64(stage, hi, g.out + p * g.out_ld + c, true, lane);
        if (PAIR) warp_store_frag64(stage, lo, g.out_lo + p * g.out_ld + c, true, lane);
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else {
    // ============================== builders: the im2col tile, in shared memory ==============================
    const int r = (threadIdx.x - kC1Builder0 * 32) & 127;
    const int group = (warp - kC1Builder0) >> 2;     // group 0 builds the CTA's even tiles, group 1 the odd ones
    for (int it = group; static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x < g.tiles; it += 2) {
      const long long t = static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x;
      const int stage = it % kStages;
      const uint32_t phase = static_cast<uint32_t>(it / kStages) & 1u;
      mbar_wait(&a_empty[stage], phase ^ 1);
      uint8_t* a_hi = a_base + stage * kStageBytes;
      build_row(g, t * 128 + r, r, a_hi, PAIR ? a_hi + kC1Tile : nullptr);
      fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor core
      mbar_arrive(&a_full[stage]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Filter gradient.  dy maps: 2-D [pixels][64 co] bf16 (hi / lo planes), box (64, 128).
template <bool PAIR>
__global__ void __launch_bounds__(kC1Threads, 1)
conv1_wgrad_kernel(const __grid_constant__ CUtensorMap dy_hi, const __grid_constant__ CUtensorMap dy_lo,
                   const Conv1Args g) {
  pdl_launch_dependents();
  constexpr int kStages = PAIR ? 3 : 5;
  constexpr int kABytes = PAIR ? 2 * kC1Tile : kC1Tile;      // im2col tile: hi | lo
  constexpr int kStageBytes = 2 * kABytes;                   // + the dY tile(s): hi | lo
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* a_full = bars;                 // builders (128 arrivals)
  uint64_t* b_full = a_full + kStages;     // TMA bytes of the dY tile(s)
  uint64_t* empty = b_full + kStages;      // MMA commit
  uint64_t* acc_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int s = 0; s < kStages; ++s)   // zero the im2col halves once (columns 27..63 stay zero)
    for (int i = threadIdx.x; i < kABytes / 16; i += kC1Threads)
      reinterpret_cast<uint4*>(smem + s * kStageBytes)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&dy_hi);
    if (PAIR) tma_prefetch_desc(&dy_lo);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&a_full[s], 128);
      mbar_init(&b_full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<64>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int my_tiles = g.tiles > static_cast<int>(blockIdx.x) ? (g.tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < g.tiles; t += gridDim.x) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sb = smem + stage * kStageBytes + kABytes;
        mbar_expect_tx(&b_full[stage], PAIR ? 2 * kC1Tile : kC1Tile);
        tma_load_2d(&dy_hi, &b_full[stage], sb, 0, t * 128);
        if (PAIR) tma_load_2d(&dy_lo, &b_full[stage], sb + kC1Tile, 0, t * 128);
      }
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(1u, 1u, 1u, 128u, 64u);   // both operands MN-major (pixels are the K dimension)
    const uint32_t s0 = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc = 0;
    for (int it = 0; it < my_tiles; ++it) {
      mbar_wait(&a_full[stage], phase);
      mbar_wait(&b_full[stage], phase);
      tc_fence_after();
      const uint32_t sa = s0 + stage * kStageBytes;
      const uint32_t sb = sa + kABytes;
      if (elect_one()) {
        // A': M = 128 = [64 k of the hi tile | 64 k of the lo tile] (bf16 mode: the hi tile twice), LBO = their distance
        const uint64_t adesc = make_smem_desc_sw128(sa, PAIR ? kC1Tile : 0, 1024, 2u);
#pragma unroll
        for (int part = PAIR ? 1 : 0; part >= 0; --part) {      // dY_lo first, then dY_hi
          const uint64_t bdesc = make_smem_desc_sw128(sb + part * kC1Tile, 0, 1024, 2u);
#pragma unroll
          for (int k = 0; k < 8; ++k) {     // 16 pixels (16 x 128 B) per MMA
            umma_f16(tmem_base, adesc + ((k * 2048u) >> 4), bdesc + ((k * 2048u) >> 4), idesc, acc);
            acc = 1;
          }
        }
        umma_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  } else if (warp < 6) {
    // partial [(2 *) CTA][64 k][64 co]: rows 0..63 (hi tile) and, for pairs, rows 64..127 (lo tile) as a second partial
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int nparts = PAIR ? 2 : 1;
    if (row < 64 * nparts) {
      float* dst = g.partial + (static_cast<size_t>(blockIdx.x) * nparts + (row >> 6)) * 4096 + (row & 63) * 64;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + c + (static_cast<uint32_t>(quarter * 32) << 16), v);
        tmem_ld_wait();
        float4* o4 = reinterpret_cast<float4*>(dst + c);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o4[i] = my_tiles > 0 ? make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                             __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  } else if (warp >= kC1Builder0) {
    const int r = (threadIdx.x - kC1Builder0 * 32) & 127;
    const int group = (warp - kC1Builder0) >> 2;
    for (int it = group; it < my_tiles; it += 2) {
      const long long t = static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x;
      const int stage = it % kStages;
      const uint32_t phase = static_cast<uint32_t>(it / kStages) & 1u;
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* a_hi = smem + stage * kStageBytes;
      build_row(g, t * 128 + r, r, a_hi, PAIR ? a_hi + kC1Tile : nullptr);
      fence_proxy_async_smem();
      mbar_arrive(&a_full[stage]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
template <bool PAIR>
constexpr int conv1_fwd_smem() {
  return 16384 + 4 * (PAIR ? 2 * kC1Tile : kC1Tile) + 1024 + 8 * kStoreWarpBytes + 1024;
}
template <bool PAIR>
constexpr int conv1_wgrad_smem() {
  return (PAIR ? 3 : 5) * 2 * (PAIR ? 2 * kC1Tile : kC1Tile) + 1024 + 1024;
}

cudaError_t launch_conv1_fwd(const CUtensorMap& w_hi, const CUtensorMap& w_lo, const Conv1Args& a, int grid, int dev,
                             cudaStream_t st) {
  static bool done[64][2] = {};
  if (a.pair) {
    if (!done[dev][1]) {
      cudaError_t e = cudaFuncSetAttribute(conv1_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           conv1_fwd_smem<true>());
      if (e != cudaSuccess) return e;
      done[dev][1] = true;
    }
    (void)launch_k(conv1_fwd_kernel<true>, dim3(grid), dim3(kC1Threads), conv1_fwd_smem<true>(), st, w_hi, w_lo, a);
  } else {
    if (!done[dev][0]) {
      cudaError_t e = cudaFuncSetAttribute(conv1_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           conv1_fwd_smem<false>());
      if (e != cudaSuccess) return e;
      done[dev][0] = true;
    }
    (void)launch_k(conv1_fwd_kernel<false>, dim3(grid), dim3(kC1Threads), conv1_fwd_smem<false>(), st, w_hi, w_lo, a);
  }
  return cudaGetLastError();
}

cudaError_t launch_conv1_wgrad(const CUtensorMap& dy_hi, const CUtensorMap& dy_lo, const Conv1Args& a, int grid,
                               int dev, cudaStream_t st) {
  static bool done[64][2] = {};
  if (a.pair) {
    if (!done[dev][1]) {
      cudaError_t e = cudaFuncSetAttribute(conv1_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           conv1_wgrad_smem<true>());
      if (e != cudaSuccess) return e;
      done[dev][1] = true;
    }
    (void)launch_k(conv1_wgrad_kernel<true>, dim3(grid), dim3(kC1Threads), conv1_wgrad_smem<true>(), st, dy_hi, dy_lo, a);
  } else {
    if (!done[dev][0]) {
      cudaError_t e = cudaFuncSetAttribute(conv1_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           conv1_wgrad_smem<false>());
      if (e != cudaSuccess) return e;
      done[dev][0] = true;
    }
    (void)launch_k(conv1_wgrad_kernel<false>, dim3(grid), dim3(kC1Threads), conv1_wgrad_smem<false>(), st, dy_hi, dy_lo, a);
  }
  return cudaGetLastError();
}

}  // namespace fcn8
