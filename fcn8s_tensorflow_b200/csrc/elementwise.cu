// HBM-bound helper kernels of the FCN-8s step: feed pre-processing, pooling, bias gradient, weight packing,
// split-K reduction, Adam.  All are coalesced, 128-bit vectorised grid-stride kernels; none has data reuse that
// would justify shared-memory staging except the packing transpose.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_gemm.cuh"
#include "kernels.h"

namespace fcn8 {

static inline int grid_for(size_t work, int threads, int max_blocks = 148 * 16) {
  size_t b = (work + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > static_cast<size_t>(max_blocks)) b = max_blocks;
  return static_cast<int>(b);
}

// ------------------------------------------------------------------------------------------------ storage formats
// Act<FMT>: 16-byte vector access to one pixel's channels in the three activation storage formats
// (include/fcn8s_b200.h): N channels per vector, `ld(C)` elements between pixels, load -> fp32, store <- fp32.
// FMT 2 (bf16 hi/lo pair, [.., 2C]): value = hi + lo exactly (fp32), stored as hi = bf16(v), lo = bf16(v - hi).
template <int FMT>
struct Act;
template <>
struct Act<0> {
  using T = __nv_bfloat16;
  static constexpr int N = 8;
  static __device__ __forceinline__ int ld(int C) { return C; }
  // load = load_raw (the memory access, no dependent arithmetic: several can be in flight) + unpack
  using Raw = uint4;
  static __device__ __forceinline__ void load_raw(const T* p, int, Raw& q) { q = *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& q, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 t = __bfloat1622float2(h[j]);
      f[2 * j] = t.x;
      f[2 * j + 1] = t.y;
    }
  }
  static __device__ __forceinline__ void load(const T* p, int C, float (&f)[8]) {
    Raw q;
    load_raw(p, C, q);
    unpack(q, f);
  }
  static __device__ __forceinline__ void store(T* p, int, const float (&f)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                              pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
};
template <>
struct Act<1> {
  using T = float;
  static constexpr int N = 4;
  static __device__ __forceinline__ int ld(int C) { return C; }
  using Raw = float4;
  static __device__ __forceinline__ void load_raw(const T* p, int, Raw& q) { q = *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& q, float (&f)[4]) {
    f[0] = q.x;
    f[1] = q.y;
    f[2] = q.z;
    f[3] = q.w;
  }
  static __device__ __forceinline__ void load(const T* p, int C, float (&f)[4]) {
    Raw q;
    load_raw(p, C, q);
    unpack(q, f);
  }
  static __device__ __forceinline__ void store(T* p, int, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <>
struct Act<2> {
  using T = __nv_bfloat16;
  static constexpr int N = 8;
  static __device__ __forceinline__ int ld(int C) { return 2 * C; }
  struct Raw {
    uint4 hi, lo;
  };
  static __device__ __forceinline__ void load_raw(const T* p, int C, Raw& q) {
    q.hi = *reinterpret_cast<const uint4*>(p);
    q.lo = *reinterpret_cast<const uint4*>(p + C);
  }
  static __device__ __forceinline__ void unpack(const Raw& q, float (&f)[8]) {
    float l[8];
    Act<0>::unpack(q.hi, f);
    Act<0>::unpack(q.lo, l);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += l[j];
  }
  static __device__ __forceinline__ void load(const T* p, int C, float (&f)[8]) {
    Raw q;
    load_raw(p, C, q);
    unpack(q, f);
  }
  static __device__ __forceinline__ void store(T* p, int C, const float (&f)[8]) {
    float l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) l[j] = f[j] - __bfloat162float(__float2bfloat16_rn(f[j]));
    Act<0>::store(p, C, f);
    Act<0>::store(p + C, C, l);
  }
};
// launch helper: FN<FMT>(args...) for a runtime format
#define FCN8_FMT_DISPATCH(fmt, CALL)      \
  do {                                    \
    if ((fmt) == 0) { CALL(0); }          \
    else if ((fmt) == 1) { CALL(1); }     \
    else { CALL(2); }                     \
  } while (0)

// ------------------------------------------------------------------------------------------------ preprocess
// Feed kernel: uint8 RGB -> mean-subtracted BGR, emitted as conv1_1's im2col row per pixel (27 columns padded to KP =
// one 128-byte operand row: 64 bf16 / 32 fp32 columns).  A CTA handles 128 consecutive pixels of one image row: the
// three input rows (with a one-pixel halo, zeros outside the image = SAME padding applied AFTER the mean subtraction)
// are staged in shared memory as floats; each thread then assembles its pixel's row with compile-time column indices.
template <int FMT>
__global__ void __launch_bounds__(128)
preprocess_im2col_kernel(const uint8_t* __restrict__ img, typename Act<FMT>::T* __restrict__ out, int N, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  using A = Act<FMT>;
  constexpr int VEC = A::N;
  constexpr int KP = FMT == 1 ? 32 : 64;
  constexpr int TW = 128;
  __shared__ float tile[3][TW + 2][3];
  const int tiles_x = (W + TW - 1) / TW;
  const long long total_tiles = static_cast<long long>(N) * H * tiles_x;
  for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int x0 = static_cast<int>(t % tiles_x) * TW;
    const int y = static_cast<int>((t / tiles_x) % H);
    const int n = static_cast<int>(t / (static_cast<long long>(tiles_x) * H));
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * (TW + 2) * 3; i += TW) {
      const int c = i % 3, xx = (i / 3) % (TW + 2), r = i / (3 * (TW + 2));
      const int gy = y + r - 1, gx = x0 + xx - 1;
      float v = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        // c = 0,1,2 -> B,G,R = RGB channel 2,1,0 minus the ImageNet means
        const float mean = (c == 0) ? 103.939f : (c == 1 ? 116.779f : 123.68f);
        v = static_cast<float>(img[((static_cast<size_t>(n) * H + gy) * W + gx) * 3 + (2 - c)]) - mean;
      }
      tile[r][xx][c] = v;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x < W) {
      const size_t pix = (static_cast<size_t>(n) * H + y) * W + x;
#pragma unroll
      for (int v = 0; v < KP / VEC; ++v) {
        float f[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int col = v * VEC + e;   // compile-time after unrolling
          f[e] = col < 27 ? tile[col / 9][threadIdx.x + (col / 3) % 3][col % 3] : 0.f;
        }
        A::store(out + pix * A::ld(KP) + v * VEC, KP, f);
      }
    }
  }
}

cudaError_t launch_preprocess(const uint8_t* img, void* out, int N, int H, int W, int dtype, cudaStream_t st) {
  const long long tiles = static_cast<long long>(N) * H * ((W + 127) / 128);
  const int blocks = static_cast<int>(tiles < 148 * 16 ? tiles : 148 * 16);
#define CALL(F) { (void)launch_k(preprocess_im2col_kernel<F>, dim3(blocks), dim3(128), 0, st, img, static_cast<typename Act<F>::T*>(out), N, H, W); }
  FCN8_FMT_DISPATCH(dtype, CALL);
#undef CALL
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ max pool
template <int FMT>
__global__ void maxpool_fwd_kernel(const typename Act<FMT>::T* __restrict__ x, typename Act<FMT>::T* __restrict__ y,
                                   int N, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  using V = Act<FMT>;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, CV = C / V::N, LD = V::ld(C);
  const size_t total = static_cast<size_t>(N) * Ho * Wo * CV;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % CV);
    size_t r = i / CV;
    const size_t opix = r;
    const int xo = static_cast<int>(r % Wo);
    r /= Wo;
    const int yo = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    float m[V::N];
#pragma unroll
    for (int e = 0; e < V::N; ++e) m[e] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = 2 * yo + dy, xx = 2 * xo + dx;
        if (yy < H && xx < W) {
          float f[V::N];
          V::load(x + ((static_cast<size_t>(n) * H + yy) * W + xx) * LD + cv * V::N, C, f);
#pragma unroll
          for (int e = 0; e < V::N; ++e) m[e] = fmaxf(m[e], f[e]);
        }
      }
    V::store(y + opix * LD + cv * V::N, C, m);
  }
}

template <int NV>
__device__ __forceinline__ float bm_pos(const float (&f)[4][NV], int e) {
  return fmaxf(fmaxf(f[0][e], f[1][e]), fmaxf(f[2][e], f[3][e]));
}
// dx[window position] = dy if it is the first maximum of the window (scan order) and x > 0, else 0.
//
// bf16 formats (FMT 0 / 2): the window logic runs on PACKED bf16 pairs -- compare-to-mask (__hgt2_mask / __heq2_mask)
// and bitwise selects, two channels per 32-bit operation, and dy's bits are copied, never converted.  (The fp32
// formulation -- unpack 5 x 8 values, arg-max with selects, re-pack 4 x 8 values -- was ~1000 instructions per
// iteration and made the bf16 kernel issue-bound at 3.1 TB/s.)  A hi/lo pair (FMT 2) compares lexicographically:
// hi = bf16(v) and lo = bf16(v - hi) are both monotone in v, so (hi, lo) orders exactly like hi + lo.
__device__ __forceinline__ uint32_t bf2_gt(uint32_t a, uint32_t b) {
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
}
__device__ __forceinline__ uint32_t bf2_eq(uint32_t a, uint32_t b) {
  return __heq2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
}
__device__ __forceinline__ uint32_t bit_sel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }
constexpr uint32_t kBf2NegInf = 0xff80ff80u;
// One 32-bit word = two channels of the four window positions.  x?: values (hi words; l? = lo words when PAIR),
// oob bit k: position k lies outside the image.  Returns w[k] = mask of the channels that position k wins (first
// strict maximum in scan order, and the maximum is > 0).
template <bool PAIR>
__device__ __forceinline__ void pool_winner_masks(const uint32_t (&xh)[4], const uint32_t (&xl)[4], uint32_t (&w)[4]) {
  uint32_t bh = xh[0], bl = PAIR ? xl[0] : 0u;
  uint32_t g[4];
  g[0] = 0u;
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    uint32_t gt = bf2_gt(xh[k], bh);
    if (PAIR) gt |= bf2_eq(xh[k], bh) & bf2_gt(xl[k], bl);
    g[k] = gt;
    bh = bit_sel(gt, xh[k], bh);
    if (PAIR) bl = bit_sel(gt, xl[k], bl);
  }
  uint32_t pos = bf2_gt(bh, 0u);
  if (PAIR) pos |= bf2_eq(bh, 0u) & bf2_gt(bl, 0u);
  w[3] = g[3] & pos;
  w[2] = g[2] & ~g[3] & pos;
  w[1] = g[1] & ~(g[2] | g[3]) & pos;
  w[0] = ~(g[1] | g[2] | g[3]) & pos;
}
__device__ __forceinline__ uint32_t u4_get(const uint4& q, int j) { return j == 0 ? q.x : (j == 1 ? q.y : (j == 2 ? q.z : q.w)); }
__device__ __forceinline__ void u4_set(uint4& q, int j, uint32_t v) {
  if (j == 0) q.x = v;
  else if (j == 1) q.y = v;
  else if (j == 2) q.z = v;
  else q.w = v;
}
template <int FMT>
__global__ void __launch_bounds__(256, FMT == 2 ? 2 : 4) maxpool_bwd_kernel(const typename Act<FMT>::T* __restrict__ x,
                                   const typename Act<FMT>::T* __restrict__ dy, typename Act<FMT>::T* __restrict__ dx,
                                   float* __restrict__ db, int N, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  using V = Act<FMT>;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, CV = C / V::N, LD = V::ld(C);
  const unsigned int total = static_cast<unsigned int>(N) * Ho * Wo * CV;   // < 2^31, checked by the launcher
  // bias gradient of the producing conv: the grid stride is a multiple of CV, so a thread always works on the same
  // channel vector and can keep its column sums in registers
  float bsum[V::N];
#pragma unroll
  for (int e = 0; e < V::N; ++e) bsum[e] = 0.f;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cv = static_cast<int>(i % CV);
    unsigned int r = i / CV;
    const unsigned int opix = r;
    const int xo = static_cast<int>(r % Wo);
    r /= Wo;
    const int yo = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    // All five loads of the thread (dy and the window's four positions) are issued before anything depends on them:
    // a position outside the image (odd H / W, SAME padding) loads the clamped -- valid -- address and is replaced by
    // -inf afterwards, so that no load sits behind a branch.  (With one conditional load per position the loads of an
    // iteration were five dependent round trips and the kernel ran at 4.2 TB/s on 25 % occupancy.)
    typename V::Raw rg, rx[4];
    V::load_raw(dy + static_cast<size_t>(opix) * LD + cv * V::N, C, rg);
    bool inb[4];
    size_t off[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = 2 * yo + (k >> 1), xx = 2 * xo + (k & 1);
      inb[k] = (yy < H) && (xx < W);
      off[k] = (static_cast<size_t>(n * H + yy) * W + xx) * LD + cv * V::N;
      const int yc = yy < H ? yy : H - 1, xc = xx < W ? xx : W - 1;
      V::load_raw(x + (static_cast<size_t>(n * H + yc) * W + xc) * LD + cv * V::N, C, rx[k]);
    }
    if constexpr (FMT == 0 || FMT == 2) {
      constexpr bool PAIR = FMT == 2;
      typename V::Raw ro[4];
      uint32_t keep[4];    // channels (as bf16x2 masks) whose window maximum is positive: what the bias gradient sums
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t xh[4], xl[4], w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (PAIR) {
            xh[k] = inb[k] ? u4_get(rx[k].hi, j) : kBf2NegInf;
            xl[k] = inb[k] ? u4_get(rx[k].lo, j) : 0u;
          } else {
            xh[k] = inb[k] ? u4_get(rx[k], j) : kBf2NegInf;
            xl[k] = 0u;
          }
        }
        pool_winner_masks<PAIR>(xh, xl, w);
        keep[j] = w[0] | w[1] | w[2] | w[3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (PAIR) {
            u4_set(ro[k].hi, j, u4_get(rg.hi, j) & w[k]);
            u4_set(ro[k].lo, j, u4_get(rg.lo, j) & w[k]);
          } else {
            u4_set(ro[k], j, u4_get(rg, j) & w[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (inb[k]) {
          if constexpr (PAIR) {
            *reinterpret_cast<uint4*>(dx + off[k]) = ro[k].hi;
            *reinterpret_cast<uint4*>(dx + off[k] + C) = ro[k].lo;
          } else {
            *reinterpret_cast<uint4*>(dx + off[k]) = ro[k];
          }
        }
      typename V::Raw rk = rg;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if constexpr (PAIR) {
          u4_set(rk.hi, j, u4_get(rg.hi, j) & keep[j]);
          u4_set(rk.lo, j, u4_get(rg.lo, j) & keep[j]);
        } else {
          u4_set(rk, j, u4_get(rg, j) & keep[j]);
        }
      }
      float gk[V::N];
      V::unpack(rk, gk);
#pragma unroll
      for (int e = 0; e < V::N; ++e) bsum[e] += gk[e];
    } else {
      float g[V::N];
      V::unpack(rg, g);
      float f[4][V::N];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        V::unpack(rx[k], f[k]);
        if (!inb[k]) {
#pragma unroll
          for (int e = 0; e < V::N; ++e) f[k][e] = -INFINITY;
        }
      }
      float o[4][V::N];
#pragma unroll
      for (int e = 0; e < V::N; ++e) {
        int best = 0;
        float bm = f[0][e];
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (f[k][e] > bm) {
            bm = f[k][e];
            best = k;
          }
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k][e] = (k == best && bm > 0.f) ? g[e] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (inb[k]) V::store(dx + off[k], C, o[k]);
#pragma unroll
      for (int e = 0; e < V::N; ++e) bsum[e] += (bm_pos(f, e) > 0.f) ? g[e] : 0.f;
    }
  }
  if (db) {
    extern __shared__ float sdb[];  // [C]
    for (int c = threadIdx.x; c < C; c += blockDim.x) sdb[c] = 0.f;
    __syncthreads();
    const int cv = static_cast<int>((blockIdx.x * blockDim.x + threadIdx.x) % CV);
    // lanes l, l + CV, ... of a warp work on the same channel vector when CV divides 32 (C = 64, 128): sum them with
    // shuffles and let one lane per vector do the shared-memory atomics (fp32 shared atomics are CAS loops)
    bool writer = true;
    if (CV < 32 && (32 % CV) == 0) {
      for (int s = 16; s >= CV; s >>= 1) {
#pragma unroll
        for (int e = 0; e < V::N; ++e) bsum[e] += __shfl_xor_sync(0xffffffffu, bsum[e], s);
      }
      writer = (threadIdx.x & 31) < CV;
    }
    if (writer) {
#pragma unroll
      for (int e = 0; e < V::N; ++e) atomicAdd(&sdb[cv * V::N + e], bsum[e]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      if (sdb[c] != 0.f) atomicAdd(db + c, sdb[c]);
  }
}

static inline int act_vec(int dtype) { return dtype == 1 ? 4 : 8; }

cudaError_t launch_maxpool_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, cudaStream_t st) {
  const size_t total = static_cast<size_t>(N) * ((H + 1) / 2) * ((W + 1) / 2) * (C / act_vec(dtype));
  const int blocks = grid_for(total, 256);
#define CALL(F) { (void)launch_k(maxpool_fwd_kernel<F>, dim3(blocks), dim3(256), 0, st, static_cast<const typename Act<F>::T*>(x), static_cast<typename Act<F>::T*>(y), N, H, W, C); }
  FCN8_FMT_DISPATCH(dtype, CALL);
#undef CALL
  return cudaGetLastError();
}
cudaError_t launch_maxpool_bwd(const void* x, const void* dy, void* dx, float* db, int N, int H, int W, int C,
                               int dtype, cudaStream_t st) {
  const size_t total = static_cast<size_t>(N) * ((H + 1) / 2) * ((W + 1) / 2) * (C / act_vec(dtype));
  if (total >= (1ull << 31) || static_cast<size_t>(N) * H >= (1ull << 31)) return cudaErrorInvalidValue;   // 32-bit indices
  // 256 % CV == 0 for every VGG width, so the grid stride keeps cv fixed.  One wave of resident CTAs (launch bounds: 4
  // per SM, 2 for the pair format): every CTA ends with C shared + C global atomics for the bias gradient, which cost
  // as much as the main loop of a short-lived CTA on the small pools (pool3..pool5 ran at 2.4-3.5 TB/s with 2368 CTAs)
  const int blocks = grid_for(total, 256, 148 * (dtype == 2 ? 2 : 4));
  const size_t sm = db ? static_cast<size_t>(C) * sizeof(float) : 0;
#define CALL(F) { (void)launch_k(maxpool_bwd_kernel<F>, dim3(blocks), dim3(256), sm, st, static_cast<const typename Act<F>::T*>(x), static_cast<const typename Act<F>::T*>(dy), static_cast<typename Act<F>::T*>(dx), db, N, H, W, C); }
  FCN8_FMT_DISPATCH(dtype, CALL);
#undef CALL
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ label feed
// onehot[p*C + c] = (ids[p] == c): the one-hot label tensor the loss epilogue reads, rebuilt on the device from the
// class ids that crossed PCIe (fcn8_pack_labels).  One thread per 4 output bytes.
__global__ void expand_labels_kernel(const uint8_t* __restrict__ ids, uint8_t* __restrict__ onehot, size_t nbytes, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t nwords = (nbytes + 3) / 4;
  for (size_t w = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; w < nwords;
       w += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b0 = 4 * w;
    size_t p = b0 / C;
    int c = static_cast<int>(b0 - p * C);
    int id = (b0 < nbytes) ? __ldg(ids + p) : -1;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (b0 + k < nbytes && c == id) out |= 1u << (8 * k);
      if (++c == C) {
        c = 0;
        ++p;
        id = (b0 + k + 1 < nbytes) ? __ldg(ids + p) : -1;
      }
    }
    if (b0 + 4 <= nbytes) {
      reinterpret_cast<uint32_t*>(onehot)[w] = out;
    } else {
      for (int k = 0; b0 + k < nbytes; ++k) onehot[b0 + k] = static_cast<uint8_t>((out >> (8 * k)) & 0xffu);
    }
  }
}
cudaError_t launch_expand_labels(const uint8_t* ids, uint8_t* onehot, long long pixels, int C, cudaStream_t st) {
  const size_t nbytes = static_cast<size_t>(pixels) * C;
  { (void)launch_k(expand_labels_kernel, dim3(grid_for((nbytes + 3) / 4, 256)), dim3(256), 0, st, ids, onehot, nbytes, C); }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ bias gradient
// Stage 1: block (bx, by) sums rows [bx*rpb, (bx+1)*rpb) of column-vector group by into ws[bx][C].
template <int FMT>
__global__ void bias_grad_stage1(const typename Act<FMT>::T* __restrict__ dy, float* __restrict__ ws, long long P,
                                 int C, int cvb, long long rpb) {
  pdl_launch_dependents();
  pdl_wait();
  using V = Act<FMT>;
  extern __shared__ float sred[];  // [RL][cvb*VN]
  const int RL = blockDim.x / cvb;
  const int cvl = threadIdx.x % cvb;
  const int rl = threadIdx.x / cvb;
  const int cv = blockIdx.y * cvb + cvl;
  const int LD = V::ld(C);
  float acc[V::N];
#pragma unroll
  for (int e = 0; e < V::N; ++e) acc[e] = 0.f;
  const long long r0 = blockIdx.x * rpb;
  const long long r1 = (r0 + rpb < P) ? r0 + rpb : P;
  if (rl < RL) {
    for (long long r = r0 + rl; r < r1; r += RL) {
      float f[V::N];
      V::load(dy + static_cast<size_t>(r) * LD + cv * V::N, C, f);
#pragma unroll
      for (int e = 0; e < V::N; ++e) acc[e] += f[e];
    }
#pragma unroll
    for (int e = 0; e < V::N; ++e) sred[(rl * cvb + cvl) * V::N + e] = acc[e];
  }
  __syncthreads();
  if (rl == 0) {
    for (int k = 1; k < RL; ++k)
#pragma unroll
      for (int e = 0; e < V::N; ++e) acc[e] += sred[(k * cvb + cvl) * V::N + e];
#pragma unroll
    for (int e = 0; e < V::N; ++e) ws[static_cast<size_t>(blockIdx.x) * C + cv * V::N + e] = acc[e];
  }
}
// out[c] = scale * sum_b ws[b][c]: one warp per column group so that the nb partial rows are read in parallel.
__global__ void colsum_stage2(const float* __restrict__ ws, float* __restrict__ out, int nb, int C, float scale,
                              int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int part = threadIdx.x >> 5, nparts = blockDim.x >> 5;
  __shared__ float red[8][33];
  float s = 0.f;
  if (c < C)
    for (int b = part; b < nb; b += nparts) s += ws[static_cast<size_t>(b) * C + c];
  red[part][threadIdx.x & 31] = s;
  __syncthreads();
  if (part == 0 && c < C) {
    for (int k = 1; k < nparts; ++k) s += red[k][threadIdx.x & 31];
    s *= scale;
    out[c] = accumulate ? out[c] + s : s;
  }
}

int bias_grad_blocks(long long P, int C) {
  long long nb = (P + 31) / 32;
  if (nb > 592) nb = 592;
  if (nb < 1) nb = 1;
  return static_cast<int>(nb);
}
cudaError_t launch_colsum(const float* ws, float* out, int nb, int C, float scale, int accumulate, cudaStream_t st) {
  { (void)launch_k(colsum_stage2, dim3((C + 31) / 32), dim3(256), 0, st, ws, out, nb, C, scale, accumulate); }
  return cudaGetLastError();
}
cudaError_t launch_bias_grad(const void* dy, float* db, long long P, int C, int dtype, float* ws, cudaStream_t st) {
  const int vec = act_vec(dtype);
  const int CV = C / vec;
  const int cvb = CV < 256 ? CV : 256;
  const int nb = bias_grad_blocks(P, C);
  const long long rpb = (P + nb - 1) / nb;
  dim3 grid(nb, CV / cvb);
  const int RL = 256 / cvb;
  const size_t sm = static_cast<size_t>(RL) * cvb * vec * sizeof(float);
#define CALL(F) { (void)launch_k(bias_grad_stage1<F>, dim3(grid), dim3(256), sm, st, static_cast<const typename Act<F>::T*>(dy), ws, P, C, cvb, rpb); }
  FCN8_FMT_DISPATCH(dtype, CALL);
#undef CALL
  return launch_colsum(ws, db, nb, C, 1.f, 0, st);
}

// ------------------------------------------------------------------------------------------------ weight packing
// fp32 -> nearest tf32 (10-bit mantissa, low 13 bits zero).  The tensor core TRUNCATES fp32 operands to tf32
// (measured: scripts/bringup.py tf32_truncation_probe), which is a biased error; rounding here first makes the
// hardware's truncation exact and the split hi + lo unbiased.
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

template <typename T>
__device__ __forceinline__ void store_packed(T* out, T* out_lo, size_t idx, float v) {
  if constexpr (sizeof(T) == 2) {
    const T h = __float2bfloat16_rn(v);
    out[idx] = h;
    if (out_lo) out_lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));  // bf16 hi/lo pair operand
  } else {
    if (out_lo) {
      // 3xTF32 operands: the MMA itself truncates `out` to its tf32 high part; lo = the exact remainder, rounded
      out[idx] = v;
      out_lo[idx] = tf32_hi(v - tf32_trunc(v));
    } else {
      out[idx] = tf32_hi(v);  // single-pass tf32: pre-round so that the MMA's truncation is exact
    }
  }
}

// mode 0: out[co][tap*CinPad + ci] = w[(tap*Cin + ci)*Cout + co]; 32x32 smem transpose per (tap, ci-tile, co-tile).
template <typename T>
__global__ void pack_fprop_kernel(const float* __restrict__ w, T* __restrict__ out, T* __restrict__ out_lo, int taps,
                                  int Cin, int Cout, int CinPad) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[32][33];
  const int tap = blockIdx.z;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    tile[r][threadIdx.x] = (ci < Cin && co < Cout) ? w[(static_cast<size_t>(tap) * Cin + ci) * Cout + co] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    if (co < Cout && ci < CinPad)
      store_packed<T>(out, out_lo, static_cast<size_t>(co) * taps * CinPad + static_cast<size_t>(tap) * CinPad + ci,
                      tile[threadIdx.x][r]);
  }
}
// mode 1: out[ci][tap'*Cout + co] = w[((taps-1-tap')*Cin + ci)*Cout + co]
template <typename T>
__global__ void pack_dgrad_kernel(const float* __restrict__ w, T* __restrict__ out, T* __restrict__ out_lo, int taps,
                                  int Cin, int Cout) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t total = static_cast<size_t>(taps) * Cin * Cout;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % Cout);
    const size_t r = i / Cout;
    const int tp = static_cast<int>(r % taps);
    const int ci = static_cast<int>(r / taps);
    store_packed<T>(out, out_lo, i, w[(static_cast<size_t>(taps - 1 - tp) * Cin + ci) * Cout + co]);
  }
}

cudaError_t launch_pack(const float* w, void* out, void* out_lo, int ksize, int Cin, int Cout, int CinPad, int mode,
                        int dtype, cudaStream_t st) {
  const int taps = ksize * ksize;
  if (mode == 0) {
    dim3 grid((Cout + 31) / 32, (CinPad + 31) / 32, taps), block(32, 8);
    if (dtype == 0)
      { (void)launch_k(pack_fprop_kernel<__nv_bfloat16>, dim3(grid), dim3(block), 0, st, w, static_cast<__nv_bfloat16*>(out), static_cast<__nv_bfloat16*>(out_lo), taps, Cin,
                                                               Cout, CinPad); }
    else
      { (void)launch_k(pack_fprop_kernel<float>, dim3(grid), dim3(block), 0, st, w, static_cast<float*>(out), static_cast<float*>(out_lo), taps,
                                                       Cin, Cout, CinPad); }
  } else {
    const size_t total = static_cast<size_t>(taps) * Cin * Cout;
    const int blocks = grid_for(total, 256);
    if (dtype == 0)
      { (void)launch_k(pack_dgrad_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, w, static_cast<__nv_bfloat16*>(out), static_cast<__nv_bfloat16*>(out_lo), taps, Cin,
                                                               Cout); }
    else
      { (void)launch_k(pack_dgrad_kernel<float>, dim3(blocks), dim3(256), 0, st, w, static_cast<float*>(out), static_cast<float*>(out_lo), taps,
                                                       Cin, Cout); }
  }
  return cudaGetLastError();
}

// hi == nullptr: lo = rna(x - trunc(x))  (the MMA truncates x itself, so x is its own "hi" operand)
// lo == nullptr: hi = rna(x)               (single-pass tf32 operand rounding)
// both:          hi = trunc(x), lo = rna(x - hi)
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                  size_t n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    if (!lo) {
      reinterpret_cast<float4*>(hi)[i] = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      continue;
    }
    const float4 h = make_float4(tf32_trunc(v.x), tf32_trunc(v.y), tf32_trunc(v.z), tf32_trunc(v.w));
    if (hi) reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] =
        make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
  }
}
cudaError_t launch_split_tf32(const float* x, float* hi, float* lo, size_t n, cudaStream_t st) {
  { (void)launch_k(split_tf32_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, st, x, hi, lo, n / 4); }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ split-K reduce
// conv: out = epilogue(sum_s partial[s][i]); same flag semantics as the in-kernel epilogue (epilogue_row32).
// FMT 0 / 1: bf16 / fp32 output; FMT 2: bf16 with the low half also written to g.out_lo (hi/lo pair output).
template <int FMT>
__global__ void conv_splitk_reduce_kernel(const float* __restrict__ partial, int splits, size_t n_elems, int ldc,
                                          ConvGemmArgs g) {
  pdl_launch_dependents();
  pdl_wait();
  using V = Act<FMT == 2 ? 0 : FMT>;
  using T = typename V::T;
  const size_t nv = n_elems / V::N;
  const uint32_t seed = g.seed_ptr ? (__ldg(g.seed_ptr) * 2u + g.seed) : g.seed;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t idx = i * V::N;
    float f[V::N];
#pragma unroll
    for (int e = 0; e < V::N; ++e) f[e] = 0.f;
    for (int s = 0; s < splits; ++s) {
#pragma unroll
      for (int q = 0; q < V::N / 4; ++q) {
        const float4 p = *reinterpret_cast<const float4*>(partial + static_cast<size_t>(s) * n_elems + idx + 4 * q);
        f[4 * q] += p.x;
        f[4 * q + 1] += p.y;
        f[4 * q + 2] += p.z;
        f[4 * q + 3] += p.w;
      }
    }
    const int c0 = static_cast<int>(idx % ldc);
    const size_t oidx = (idx / ldc) * static_cast<size_t>(g.osW) + c0;  // pixel stride of the output tensors
    if (g.flags & EPI_BIAS) {
#pragma unroll
      for (int e = 0; e < V::N; ++e) f[e] += g.bias[c0 + e];
    }
    if (g.flags & EPI_RESIDUAL) {
      float r[V::N];
      V::load(reinterpret_cast<const T*>(g.residual) + oidx, 0, r);
#pragma unroll
      for (int e = 0; e < V::N; ++e) f[e] += r[e];
      if (FMT == 2 && g.residual_lo) {
        V::load(reinterpret_cast<const T*>(g.residual_lo) + oidx, 0, r);
#pragma unroll
        for (int e = 0; e < V::N; ++e) f[e] += r[e];
      }
    }
    if (g.flags & EPI_RELU) {
#pragma unroll
      for (int e = 0; e < V::N; ++e) f[e] = fmaxf(f[e], 0.f);
    }
    if (g.flags & EPI_DROPOUT) {
#pragma unroll
      for (int e = 0; e < V::N; ++e)
        f[e] = dropout_keep(seed, static_cast<uint64_t>(idx) + e, g.keep_threshold) ? f[e] * g.inv_keep : 0.f;
    }
    if (g.flags & EPI_MASK) {
      float m[V::N];
      V::load(reinterpret_cast<const T*>(g.mask_src) + oidx, 0, m);
#pragma unroll
      for (int e = 0; e < V::N; ++e) f[e] = m[e] > 0.f ? f[e] * g.mask_scale : 0.f;
    }
    if constexpr (FMT == 1) {
      if (g.flags & EPI_ROUND_TF32) {
#pragma unroll
        for (int e = 0; e < V::N; ++e) f[e] = round_tf32(f[e]);
      }
    }
    V::store(reinterpret_cast<T*>(g.out) + oidx, 0, f);
    if constexpr (FMT == 2) {
      float l[V::N];
#pragma unroll
      for (int e = 0; e < V::N; ++e) l[e] = f[e] - __bfloat162float(__float2bfloat16_rn(f[e]));
      V::store(reinterpret_cast<T*>(g.out_lo) + oidx, 0, l);
    }
  }
}
cudaError_t launch_conv_splitk_reduce(const float* partial, int splits, size_t n_elems, int ldc,
                                      const ConvGemmArgs& g, int dtype, cudaStream_t st) {
  const int fmt = dtype == 1 ? 1 : (g.out_lo ? 2 : 0);
  const int blocks = grid_for(n_elems / act_vec(dtype), 256);
#define CALL(F) { (void)launch_k(conv_splitk_reduce_kernel<F>, dim3(blocks), dim3(256), 0, st, partial, splits, n_elems, ldc, g); }
  FCN8_FMT_DISPATCH(fmt, CALL);
#undef CALL
  return cudaGetLastError();
}

// wgrad: out[r][c] = scale * sum_s partial[s][r][c] for r < rows_valid, c < out_cols (partial has rows_pad rows of ldc
// columns per split; out is [rows_valid][out_cols]).
__global__ void wgrad_splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int splits,
                                           size_t rows_pad, int rows_valid, int ldc) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n4 = static_cast<size_t>(rows_valid) * ldc / 4;
  const size_t split_stride = rows_pad * ldc;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 p = *reinterpret_cast<const float4*>(partial + s * split_stride + i * 4);
      a.x += p.x;
      a.y += p.y;
      a.z += p.z;
      a.w += p.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
  }
}
// Small outputs with many splits (conv1_x / conv2_x filter gradients: 1.7 K ... 150 K elements, up to 148 splits): one
// thread per element would leave most SMs idle in a long serial loop, so 8 threads share an element's splits and a
// shared-memory tree adds their sums in a fixed order.
__global__ void __launch_bounds__(256)
wgrad_splitk_reduce_wide_kernel(const float* __restrict__ partial, float* __restrict__ out, int splits,
                                size_t rows_pad, int rows_valid, int ldc) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float4 red[8][32];
  const size_t n4 = static_cast<size_t>(rows_valid) * ldc / 4;
  const size_t split_stride = rows_pad * ldc;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const size_t i = static_cast<size_t>(blockIdx.x) * 32 + tx;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4)
    for (int s = ty; s < splits; s += 8) {
      const float4 p = *reinterpret_cast<const float4*>(partial + s * split_stride + i * 4);
      a.x += p.x;
      a.y += p.y;
      a.z += p.z;
      a.w += p.w;
    }
  red[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && i < n4) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 p = red[k][tx];
      a.x += p.x;
      a.y += p.y;
      a.z += p.z;
      a.w += p.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
  }
}
// Column-clipped / scaled variant (score heads: 64 padded class columns -> out [rows][out_cols], scale = the skip
// scale): one thread per output element, splits summed in order.
__global__ void wgrad_splitk_reduce_clip_kernel(const float* __restrict__ partial, float* __restrict__ out, int splits,
                                                size_t rows_pad, int rows_valid, int ldc, int out_cols, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n = static_cast<size_t>(rows_valid) * out_cols;
  const size_t split_stride = rows_pad * ldc;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / out_cols;
    const int c = static_cast<int>(i - r * out_cols);
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += partial[s * split_stride + r * ldc + c];
    out[i] = a * scale;
  }
}
cudaError_t launch_wgrad_splitk_reduce(const float* partial, float* out, int splits, size_t rows_pad, int rows_valid,
                                       int ldc, int out_cols, float scale, cudaStream_t st) {
  if (out_cols != ldc || scale != 1.f) {
    const size_t n = static_cast<size_t>(rows_valid) * out_cols;
    { (void)launch_k(wgrad_splitk_reduce_clip_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, partial, out, splits, rows_pad, rows_valid, ldc, out_cols, scale); }
    return cudaGetLastError();
  }
  const size_t n4 = static_cast<size_t>(rows_valid) * ldc / 4;
  if (splits >= 16 && n4 <= static_cast<size_t>(148) * 256 * 2) {
    { (void)launch_k(wgrad_splitk_reduce_wide_kernel, dim3(static_cast<unsigned>((n4 + 31) / 32)), dim3(256), 0, st, partial, out, splits, rows_pad, rows_valid, ldc); }
    return cudaGetLastError();
  }
  { (void)launch_k(wgrad_splitk_reduce_kernel, dim3(grid_for(n4, 256)), dim3(256), 0, st, partial, out, splits, rows_pad, rows_valid, ldc); }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ Adam / L2
// 28 B/param of HBM traffic (read p,g,m,v; write p,m,v): the floor for an fp32 Adam step; + 2 (4) B/param when the
// bf16 (hi/lo) tensor-core shadow of the parameters is refreshed in the same pass (w_hi / w_lo != nullptr), which
// replaces every per-step weight re-packing kernel: the GEMMs read the shadow in TF layout directly.
__device__ __forceinline__ void store_shadow4(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t i4, const float4& p) {
  const uint32_t h0 = pack_bf16x2(p.x, p.y), h1 = pack_bf16x2(p.z, p.w);
  reinterpret_cast<uint2*>(hi)[i4] = make_uint2(h0, h1);
  if (lo) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h0));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h1));
    reinterpret_cast<uint2*>(lo)[i4] = make_uint2(pack_bf16x2(p.x - a.x, p.y - a.y), pack_bf16x2(p.z - b.x, p.w - b.y));
  }
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr_t, float b1, float b2, float eps, float gscale,
                            __nv_bfloat16* __restrict__ w_hi, __nv_bfloat16* __restrict__ w_lo,
                            const float* __restrict__ lr_ptr, const __nv_bfloat16* __restrict__ g16) {
  pdl_launch_dependents();
  pdl_wait();
  if (lr_ptr) lr_t = __ldg(lr_ptr);  // per-step value in device memory (CUDA-graph replays keep kernel arguments)
  const size_t n4 = n / 4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg;
    if (g16) {   // gradients that went through a bf16 all-reduce (data-parallel option): 8 B instead of 16 B per 4
      const uint2 q = reinterpret_cast<const uint2*>(g16)[i];
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
      gg = make_float4(a.x, a.y, b.x, b.y);
    } else {
      gg = reinterpret_cast<const float4*>(g)[i];
    }
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define FCN8_ADAM1(c)                                  \
  {                                                    \
    const float gr = gg.c * gscale;                    \
    mm.c = b1 * mm.c + (1.f - b1) * gr;                \
    vv.c = b2 * vv.c + (1.f - b2) * gr * gr;           \
    pp.c -= lr_t * mm.c / (sqrtf(vv.c) + eps);         \
  }
    FCN8_ADAM1(x) FCN8_ADAM1(y) FCN8_ADAM1(z) FCN8_ADAM1(w)
#undef FCN8_ADAM1
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (w_hi) store_shadow4(w_hi, w_lo, i, pp);
  }
  // tail (n not a multiple of 4)
  if (blockIdx.x == 0) {
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      const float gr = (g16 ? __bfloat162float(g16[i]) : g[i]) * gscale;
      m[i] = b1 * m[i] + (1.f - b1) * gr;
      v[i] = b2 * v[i] + (1.f - b2) * gr * gr;
      p[i] -= lr_t * m[i] / (sqrtf(v[i]) + eps);
      if (w_hi) {
        w_hi[i] = __float2bfloat16_rn(p[i]);
        if (w_lo) w_lo[i] = __float2bfloat16_rn(p[i] - __bfloat162float(w_hi[i]));
      }
    }
  }
}
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, size_t n, float lr_t, float b1, float b2,
                        float eps, float gscale, void* w_hi, void* w_lo, const float* lr_ptr, const void* g16,
                        cudaStream_t st) {
  { (void)launch_k(adam_kernel, dim3(grid_for(n / 4 + 1, 256, 148 * 8)), dim3(256), 0, st, p, g, m, v, n, lr_t, b1, b2, eps, gscale, static_cast<__nv_bfloat16*>(w_hi), static_cast<__nv_bfloat16*>(w_lo), lr_ptr, static_cast<const __nv_bfloat16*>(g16)); }
  return cudaGetLastError();
}
// out = bf16(x): the flat gradient buffer in the wire format of the bf16 all-reduce option (n multiple of 4 handled
// vectorised, the tail scalar).
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n4 = n / 4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) out[i] = __float2bfloat16_rn(x[i]);
}
cudaError_t launch_cast_bf16(const float* x, void* out, size_t n, cudaStream_t st) {
  { (void)launch_k(cast_bf16_kernel, dim3(grid_for(n / 4 + 1, 256, 148 * 8)), dim3(256), 0, st, x, static_cast<__nv_bfloat16*>(out), n); }
  return cudaGetLastError();
}
// scalars[0] = lr_t (float), scalars[1] = dropout seed (uint32 bits): written by a kernel (arguments by value) so that
// a host running many steps ahead of the device never races with a staging buffer.
__global__ void set_step_scalars_kernel(float* scalars, float lr_t, uint32_t seed) {
  pdl_launch_dependents();
  pdl_wait();
  scalars[0] = lr_t;
  reinterpret_cast<uint32_t*>(scalars)[1] = seed;
}
cudaError_t launch_set_step_scalars(float* scalars, float lr_t, uint32_t seed, cudaStream_t st) {
  { (void)launch_k(set_step_scalars_kernel, dim3(1), dim3(1), 0, st, scalars, lr_t, seed); }
  return cudaGetLastError();
}
// w_hi = bf16(p), w_lo = bf16(p - w_hi): the tensor-core shadow of the fp32 parameters (after load_weights).
__global__ void shadow_kernel(const float* __restrict__ p, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                              size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n4 = n / 4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    store_shadow4(hi, lo, i, reinterpret_cast<const float4*>(p)[i]);
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      hi[i] = __float2bfloat16_rn(p[i]);
      if (lo) lo[i] = __float2bfloat16_rn(p[i] - __bfloat162float(hi[i]));
    }
}
cudaError_t launch_shadow(const float* p, void* hi, void* lo, size_t n, cudaStream_t st) {
  { (void)launch_k(shadow_kernel, dim3(grid_for(n / 4 + 1, 256, 148 * 8)), dim3(256), 0, st, p, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), n); }
  return cudaGetLastError();
}

__global__ void l2_reg_kernel(const float* __restrict__ w, float* __restrict__ g, float* __restrict__ loss, size_t n,
                              float rate) {
  pdl_launch_dependents();
  pdl_wait();
  float s = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float x = w[i];
    if (g) g[i] += rate * x;
    s += x * x;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    float t = 0.f;
    for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) t += ws[k];
    atomicAdd(loss, 0.5f * rate * t);
  }
}
cudaError_t launch_l2_reg(const float* w, float* g, float* loss, size_t n, float rate, cudaStream_t st) {
  { (void)launch_k(l2_reg_kernel, dim3(grid_for(n, 256, 64)), dim3(256), 0, st, w, g, loss, n, rate); }
  return cudaGetLastError();
}

}  // namespace fcn8
