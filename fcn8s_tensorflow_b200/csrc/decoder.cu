// FCN-8s decoder kernels (fcn8s_tensorflow.py:154-237), loss / predictor (:253, :268-269) and metrics (:280-301).
// The decoder is 0.4 % of the step's FLOPs and HBM-bound (C = num_classes channels per pixel), so these are fp32
// CUDA-core kernels organised for coalescing and shared-memory reuse of the small filter tensors.
// Limit: num_classes <= 32 (register tiles); larger C returns FCN8_ERR_UNSUPPORTED at the C ABI.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace fcn8 {

constexpr int CMAX = 32;

static inline int grid_for(size_t work, int threads, int max_blocks = 148 * 16) {
  size_t b = (work + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > static_cast<size_t>(max_blocks)) b = max_blocks;
  return static_cast<int>(b);
}

// Scalar access to channel c of pixel p in the three activation storage formats (include/fcn8s_b200.h):
// FMT 0 bf16 [P][C], 1 fp32 [P][C], 2 bf16 hi/lo pair [P][2C] (value = hi + lo).
template <int FMT>
__device__ __forceinline__ float ld_px(const void* x, long long p, int C, int c) {
  if constexpr (FMT == 1) {
    return static_cast<const float*>(x)[p * C + c];
  } else if constexpr (FMT == 0) {
    return __bfloat162float(static_cast<const __nv_bfloat16*>(x)[p * C + c]);
  } else {
    const __nv_bfloat16* b = static_cast<const __nv_bfloat16*>(x) + p * 2 * C + c;
    return __bfloat162float(b[0]) + __bfloat162float(b[C]);
  }
}
// 8 consecutive channels c .. c+7 of pixel p (c and C multiples of 8): 16-byte loads
template <int FMT>
__device__ __forceinline__ void ld_px8(const void* x, long long p, int C, int c, float (&f)[8]) {
  if constexpr (FMT == 1) {
    const float4* q = reinterpret_cast<const float4*>(static_cast<const float*>(x) + p * C + c);
    const float4 a = __ldg(q), b = __ldg(q + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(x) + p * (FMT == 2 ? 2 : 1) * C + c;
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(base));
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __bfloat1622float2(h2[j]);
      f[2 * j] = t.x;
      f[2 * j + 1] = t.y;
    }
    if constexpr (FMT == 2) {
      const uint4 l = __ldg(reinterpret_cast<const uint4*>(base + C));
      const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&l);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __bfloat1622float2(l2[j]);
        f[2 * j] += t.x;
        f[2 * j + 1] += t.y;
      }
    }
  }
}
template <int FMT>
__device__ __forceinline__ void st_px(void* x, long long p, int C, int c, float v) {
  if constexpr (FMT == 1) {
    static_cast<float*>(x)[p * C + c] = v;
  } else if constexpr (FMT == 0) {
    static_cast<__nv_bfloat16*>(x)[p * C + c] = __float2bfloat16_rn(v);
  } else {
    __nv_bfloat16* b = static_cast<__nv_bfloat16*>(x) + p * 2 * C + c;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    b[0] = h;
    b[C] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}
#define FCN8_FMT_DISPATCH(fmt, CALL)      \
  do {                                    \
    if ((fmt) == 0) { CALL(0); }          \
    else if ((fmt) == 1) { CALL(1); }     \
    else { CALL(2); }                     \
  } while (0)

// ------------------------------------------------------------------------------------------------ score heads
// 1x1 convolutions Cin -> C <= 32 classes (fcn8s_tensorflow.py:171-200) and their gradients.  0.84 GFLOP per c2 step:
// HBM-bound on the activation tensors, so these are CUDA-core kernels whose job is to touch x / dx exactly once,
// coalesced, with K and ds staged in shared memory.
constexpr int kHeadTP = 32;    // pixels per CTA tile (bwd_x)
constexpr int kHeadFP = 128;   // pixels per CTA tile (fwd)
constexpr int kHeadKC = 64;    // input channels per shared-memory chunk (fwd)

// fwd: CTA = 128 threads = 32 pixel groups x 4 class groups; a thread owns 4 pixels (pg, pg+32, pg+64, pg+96) x the
// classes q, q+4, ... : 4 + C/4 shared-memory reads feed 4*C/4 FMAs per input channel (the one-pixel version was bound
// by its 6 reads per 5 FMAs).  grid.y splits Cin into slices of kslice channels (fc7: 4096 channels but only 2048
// pixels); the slices write partial sums [slice][P][C] that head_fwd_reduce_kernel adds in a fixed order (the forward
// pass stays bit-reproducible).
template <int FMT>
__global__ void __launch_bounds__(128)
head_fwd_kernel(const void* __restrict__ x, const float* __restrict__ K, const float* __restrict__ b,
                float* __restrict__ s, long long P, int Cin, int C, float scale, int kslice) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float xs[kHeadFP][kHeadKC + 1];
  __shared__ float Ks[kHeadKC][CMAX];
  const long long p0 = static_cast<long long>(blockIdx.x) * kHeadFP;
  const int pg = threadIdx.x >> 2, q = threadIdx.x & 3;
  const int kbeg = blockIdx.y * kslice, kend = min(Cin, kbeg + kslice);
  float acc[4][CMAX / 4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int j = 0; j < CMAX / 4; ++j) acc[u][j] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += kHeadKC) {
    __syncthreads();
    if ((Cin & 7) == 0 && k0 + kHeadKC <= kend) {
      // 8 channels (16 bytes of bf16) per load: the activation tile is the only large operand of this kernel
      for (int i = threadIdx.x; i < kHeadFP * (kHeadKC / 8); i += 128) {
        const int r = i / (kHeadKC / 8), c = (i % (kHeadKC / 8)) * 8;
        float f[8];
        if (p0 + r < P) {
          ld_px8<FMT>(x, p0 + r, Cin, k0 + c, f);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) xs[r][c + j] = f[j];
      }
    } else {
      for (int i = threadIdx.x; i < kHeadFP * kHeadKC; i += 128) {
        const int r = i / kHeadKC, c = i % kHeadKC;
        xs[r][c] = (p0 + r < P && k0 + c < kend) ? ld_px<FMT>(x, p0 + r, Cin, k0 + c) : 0.f;
      }
    }
    for (int i = threadIdx.x; i < kHeadKC * CMAX; i += 128) {
      const int r = i / CMAX, c = i % CMAX;
      Ks[r][c] = (k0 + r < kend && c < C) ? __ldg(K + static_cast<size_t>(k0 + r) * C + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int ci = 0; ci < kHeadKC; ++ci) {
      float xv[4], kv[CMAX / 4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = xs[pg + 32 * u][ci];
#pragma unroll
      for (int j = 0; j < CMAX / 4; ++j) kv[j] = (q + 4 * j < C) ? Ks[ci][q + 4 * j] : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < CMAX / 4; ++j)
          if (q + 4 * j < C) acc[u][j] = fmaf(xv[u], kv[j], acc[u][j]);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long p = p0 + pg + 32 * u;
    if (p < P) {
#pragma unroll
      for (int j = 0; j < CMAX / 4; ++j)
        if (q + 4 * j < C) {
          const float v = scale * acc[u][j] + (blockIdx.y == 0 ? b[q + 4 * j] : 0.f);
          s[(static_cast<size_t>(blockIdx.y) * P + p) * C + q + 4 * j] = v;   // gridDim.y == 1: s is the output
        }
    }
  }
}
__global__ void head_fwd_reduce_kernel(const float* __restrict__ part, float* __restrict__ s, size_t n, int slices) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < slices; ++k) a += part[k * n + i];
    s[i] = a;
  }
}

// dK / db partials: CTA (bx, by) covers pixels [bx*ppb, ...) and input channels [by*256, ...); thread = one input
// channel holding C accumulators; ds rows are broadcast from shared memory, x is read coalesced, 4 pixels in flight.
template <int FMT>
__global__ void __launch_bounds__(256)
head_bwd_w_kernel(const void* __restrict__ x, const float* __restrict__ ds, float* __restrict__ ws, long long P,
                  int Cin, int C, long long ppb) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float sds[64][CMAX];
  const int ci = blockIdx.y * blockDim.x + threadIdx.x;
  for (int i = threadIdx.x; i < 64 * CMAX; i += blockDim.x) (&sds[0][0])[i] = 0.f;   // columns >= C stay zero
  const long long p0 = blockIdx.x * ppb;
  const long long p1 = (p0 + ppb < P) ? p0 + ppb : P;
  float acc[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
  for (long long pc = p0; pc < p1; pc += 64) {
    const int np = static_cast<int>((p1 - pc < 64) ? (p1 - pc) : 64);
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * C; i += blockDim.x) sds[i / C][i % C] = (i < np * C) ? ds[pc * C + i] : 0.f;
    __syncthreads();
    if (ci < Cin) {
      for (int q = 0; q < np; q += 4) {
        float xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = (q + u < np) ? ld_px<FMT>(x, pc + q + u, Cin, ci) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4* d4 = reinterpret_cast<const float4*>(&sds[q + u][0]);   // broadcast reads, 16 B each
#pragma unroll
          for (int c4 = 0; c4 < CMAX / 4; ++c4)
            if (4 * c4 < C) {
              const float4 d = d4[c4];
              acc[4 * c4 + 0] = fmaf(xv[u], d.x, acc[4 * c4 + 0]);
              acc[4 * c4 + 1] = fmaf(xv[u], d.y, acc[4 * c4 + 1]);
              acc[4 * c4 + 2] = fmaf(xv[u], d.z, acc[4 * c4 + 2]);
              acc[4 * c4 + 3] = fmaf(xv[u], d.w, acc[4 * c4 + 3]);
            }
        }
      }
    }
  }
  if (ci < Cin) {
    float* o = ws + (static_cast<size_t>(blockIdx.x) * Cin + ci) * C;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) o[c] = acc[c];
  }
}
// column sums of ds [P][C] -> partial [nb][C]
__global__ void rows_colsum_kernel(const float* __restrict__ ds, float* __restrict__ ws, long long P, int C,
                                   long long ppb) {
  pdl_launch_dependents();
  pdl_wait();
  const long long p0 = blockIdx.x * ppb;
  const long long p1 = (p0 + ppb < P) ? p0 + ppb : P;
  __shared__ float red[256];
  // thread t handles column t % C, row lane t / C
  const int c = threadIdx.x % C;
  const int rl = threadIdx.x / C;
  const int RL = blockDim.x / C;
  float a = 0.f;
  if (rl < RL)
    for (long long p = p0 + rl; p < p1; p += RL) a += ds[p * C + c];
  red[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int k = 0; k < RL; ++k) t += red[k * C + threadIdx.x];
    ws[static_cast<size_t>(blockIdx.x) * C + threadIdx.x] = t;
  }
}
// dx[p][ci] = scale * sum_c ds[p][c] * K[ci][c]  (* relu/dropout mask of x).  CTA (bx, by) = 32 pixels x 256 input
// channels; thread = one input channel with its K row in registers, ds rows broadcast from shared memory; every
// x / dx access is a coalesced row of 256 channels.
template <int FMT>
__global__ void __launch_bounds__(256)
head_bwd_x_kernel(const void* __restrict__ x, const float* __restrict__ K, const float* __restrict__ ds,
                  void* __restrict__ dx, long long P, int Cin, int C, float scale, int mask, float mask_scale) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float sds[kHeadTP][CMAX];
  const long long p0 = static_cast<long long>(blockIdx.x) * kHeadTP;
  const int np = static_cast<int>((P - p0 < kHeadTP) ? (P - p0) : kHeadTP);
  const int ci = blockIdx.y * blockDim.x + threadIdx.x;
  for (int i = threadIdx.x; i < np * CMAX; i += blockDim.x) {
    const int r = i / CMAX, c = i % CMAX;
    sds[r][c] = c < C ? ds[(p0 + r) * C + c] : 0.f;
  }
  float kr[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) kr[c] = (c < C && ci < Cin) ? __ldg(K + static_cast<size_t>(ci) * C + c) * scale : 0.f;
  __syncthreads();
  if (ci >= Cin) return;
  for (int q = 0; q < np; ++q) {
    float a = 0.f;
    const float4* d4 = reinterpret_cast<const float4*>(&sds[q][0]);   // broadcast reads, 16 B each
#pragma unroll
    for (int c4 = 0; c4 < CMAX / 4; ++c4)
      if (4 * c4 < C) {
        const float4 d = d4[c4];
        a = fmaf(d.x, kr[4 * c4 + 0], a);
        a = fmaf(d.y, kr[4 * c4 + 1], a);
        a = fmaf(d.z, kr[4 * c4 + 2], a);
        a = fmaf(d.w, kr[4 * c4 + 3], a);
      }
    if (mask) a = (ld_px<FMT>(x, p0 + q, Cin, ci) > 0.f) ? a * mask_scale : 0.f;
    st_px<FMT>(dx, p0 + q, Cin, ci, a);
  }
}

// enough CTAs to fill the GPU: split Cin when there are few pixel tiles (fc7 at 1/32 resolution)
int head_fwd_slices(long long P, int Cin) {
  const long long blocks = (P + kHeadFP - 1) / kHeadFP;
  int ksplit = 1;
  while (blocks * ksplit < 2 * 148 && Cin / (ksplit * 2) >= 2 * kHeadKC) ksplit *= 2;
  return ksplit;
}
cudaError_t launch_head_fwd(const void* x, const float* K, const float* b, float* s, long long P, int Cin, int C,
                            float scale, int dtype, float* ws, cudaStream_t st) {
  const int blocks = static_cast<int>((P + kHeadFP - 1) / kHeadFP);
  const int ksplit = head_fwd_slices(P, Cin);
  const int kslice = (Cin + ksplit - 1) / ksplit;
  float* dst = ksplit > 1 ? ws : s;
  dim3 grid(blocks, ksplit);
#define CALL(F) { (void)launch_k(head_fwd_kernel<F>, dim3(grid), dim3(128), 0, st, x, K, b, dst, P, Cin, C, scale, kslice); }
  FCN8_FMT_DISPATCH(dtype, CALL);
#undef CALL
  if (ksplit > 1) {
    const size_t n = static_cast<size_t>(P) * C;
    { (void)launch_k(head_fwd_reduce_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, ws, s, n, ksplit); }
  }
  return cudaGetLastError();
}
int head_bwd_blocks(long long P) {
  long long nb = (P + 63) / 64;
  if (nb > 512) nb = 512;
  if (nb < 1) nb = 1;
  return static_cast<int>(nb);
}
cudaError_t launch_head_bwd(const void* x, const float* K, const float* ds, float* dK, float* db, void* dx,
                            long long P, int Cin, int C, float scale, int dtype, int mask, float mask_scale, float* ws,
                            cudaStream_t st) {
  const int nb = head_bwd_blocks(P);
  const long long ppb = (P + nb - 1) / nb;
  float* ws_k = ws;                                          // [nb][Cin][C]
  float* ws_b = ws + static_cast<size_t>(nb) * Cin * C;      // [nb][C]
  dim3 grid(nb, (Cin + 255) / 256);
#define CALL(F) { (void)launch_k(head_bwd_w_kernel<F>, dim3(grid), dim3(256), 0, st, x, ds, ws_k, P, Cin, C, ppb); }
  FCN8_FMT_DISPATCH(dtype, CALL);
#undef CALL
  { (void)launch_k(rows_colsum_kernel, dim3(nb), dim3(256), 0, st, ds, ws_b, P, C, ppb); }
  cudaError_t e = launch_colsum(ws_k, dK, nb, Cin * C, scale, 0, st);
  if (e != cudaSuccess) return e;
  e = launch_colsum(ws_b, db, nb, C, 1.f, 0, st);
  if (e != cudaSuccess) return e;
  if (dx) {
    dim3 gx(static_cast<unsigned>((P + kHeadTP - 1) / kHeadTP), (Cin + 255) / 256);
#define CALL(F) { (void)launch_k(head_bwd_x_kernel<F>, dim3(gx), dim3(256), 0, st, x, K, ds, dx, P, Cin, C, scale, mask, mask_scale); }
    FCN8_FMT_DISPATCH(dtype, CALL);
#undef CALL
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ transposed conv
// CUDA-core transposed convolution for the two small 2x stages (fcn8s_tensorflow.py:204-224; 0.13 GFLOP per c2
// step, launch-latency-bound) -- the 8x stage runs on the tensor cores (fcn8_upscore_tc_*).  Output pixel (oy, ox)
// = (s*J - p + dy, s*I - p + dx) depends on the 2x2 inputs (J-1+ty, I-1+tx) through taps a = dy + s*(1-ty),
// b = dx + s*(1-tx).  One thread per output element (pixel, co); the filter sits in shared memory with
// the ci rows padded to C+1 floats so that the co-strided reads are bank-conflict free.
__global__ void upscore_fwd_kernel(const float* __restrict__ x, const float* __restrict__ T,
                                   const float* __restrict__ bias, const float* __restrict__ skip,
                                   float* __restrict__ y, int N, int h, int w, int C, int s) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sT[];  // [k*k][C][C+1]  (a*k+b, co, ci)
  const int p = s / 2, k = 2 * s, CP1 = C + 1;
  for (int i = threadIdx.x; i < k * k * C * C; i += blockDim.x) {
    const int ci = i % C, r = i / C;
    sT[r * CP1 + ci] = T[i];
  }
  __syncthreads();
  const int H = h * s, W = w * s;
  const size_t total = static_cast<size_t>(N) * H * W * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % C);
    size_t r = i / C;
    const int ox = static_cast<int>(r % W);
    r /= W;
    const int oy = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    const int J = (oy + p) / s, dy = (oy + p) % s, I = (ox + p) / s, dx = (ox + p) % s;
    float acc = bias[co];
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const int ty = tap >> 1, tx = tap & 1;
      const int iy = J - 1 + ty, ix = I - 1 + tx;
      if (iy < 0 || iy >= h || ix < 0 || ix >= w) continue;
      const float* xp = x + ((static_cast<size_t>(n) * h + iy) * w + ix) * C;
      const float* tp = sT + (((dy + s * (1 - ty)) * k + dx + s * (1 - tx)) * C + co) * CP1;
      for (int ci = 0; ci < C; ++ci) acc = fmaf(__ldg(xp + ci), tp[ci], acc);
    }
    y[i] = acc + (skip ? skip[i] : 0.f);
  }
}

// dx[n,i,j,ci] = sum_{a,b,co} dy[n, s*i+a-p, s*j+b-p, co] * T[a,b,co,ci]; one thread per input element (pixel, ci),
// the filter in shared memory (ci fastest: conflict-free), dy rows broadcast across the C threads of a pixel.
__global__ void upscore_bwd_x_kernel(const float* __restrict__ dy, const float* __restrict__ T, float* __restrict__ dx,
                                     int N, int h, int w, int C, int s) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sT[];  // [k*k][C][C]
  const int p = s / 2, k = 2 * s;
  for (int i = threadIdx.x; i < k * k * C * C; i += blockDim.x) sT[i] = T[i];
  __syncthreads();
  const int H = h * s, W = w * s;
  const size_t total = static_cast<size_t>(N) * h * w * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % C);
    size_t r = i / C;
    const int jx = static_cast<int>(r % w);
    r /= w;
    const int iy = static_cast<int>(r % h);
    const int n = static_cast<int>(r / h);
    float acc = 0.f;
    for (int a = 0; a < k; ++a) {
      const int oy = s * iy + a - p;
      if (oy < 0 || oy >= H) continue;
      for (int b = 0; b < k; ++b) {
        const int ox = s * jx + b - p;
        if (ox < 0 || ox >= W) continue;
        const float* dp = dy + ((static_cast<size_t>(n) * H + oy) * W + ox) * C;
        const float* tp = sT + (a * k + b) * C * C + ci;
        for (int co = 0; co < C; ++co) acc = fmaf(__ldg(dp + co), tp[co * C], acc);
      }
    }
    dx[i] = acc;
  }
}

// dT[a,b,co,ci] partials: CTA (tap, split) reduces its share of the input pixels; thread owns a 2x2 (co,ci) tile.
__global__ void upscore_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ ws,
                                     int N, int h, int w, int C, int s, int nsplit) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sx[32][CMAX];
  __shared__ float sg[32][CMAX];
  const int p = s / 2, k = 2 * s;
  const int H = h * s, W = w * s;
  const int tap = blockIdx.x;
  const int a = tap / k, b = tap % k;
  const int CT = (C + 1) / 2;
  const int tco = (threadIdx.x / CT) * 2, tci = (threadIdx.x % CT) * 2;
  const bool owner = threadIdx.x < CT * CT;
  float acc00 = 0.f, acc01 = 0.f, acc10 = 0.f, acc11 = 0.f;
  const long long total = static_cast<long long>(N) * h * w;
  const long long per = (total + nsplit - 1) / nsplit;
  const long long q0 = blockIdx.y * per;
  const long long q1 = (q0 + per < total) ? q0 + per : total;
  for (long long qc = q0; qc < q1; qc += 32) {
    __syncthreads();
    for (int t = threadIdx.x; t < 32 * C; t += blockDim.x) {
      const int r = t / C, c = t % C;
      const long long q = qc + r;
      float xv = 0.f, gv = 0.f;
      if (q < q1) {
        const int jx = static_cast<int>(q % w);
        const int iy = static_cast<int>((q / w) % h);
        const int n = static_cast<int>(q / (static_cast<long long>(w) * h));
        const int oy = s * iy + a - p, ox = s * jx + b - p;
        if (oy >= 0 && oy < H && ox >= 0 && ox < W) {
          xv = x[q * C + c];
          gv = dy[((static_cast<size_t>(n) * H + oy) * W + ox) * C + c];
        }
      }
      sx[r][c] = xv;
      sg[r][c] = gv;
    }
    __syncthreads();
    if (owner) {
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const float g0 = sg[r][tco], g1 = sg[r][tco + 1];
        const float x0 = sx[r][tci], x1 = sx[r][tci + 1];
        acc00 = fmaf(g0, x0, acc00);
        acc01 = fmaf(g0, x1, acc01);
        acc10 = fmaf(g1, x0, acc10);
        acc11 = fmaf(g1, x1, acc11);
      }
    }
  }
  if (owner) {
    float* o = ws + (static_cast<size_t>(blockIdx.y) * k * k + tap) * C * C;
    o[tco * C + tci] = acc00;
    if (tci + 1 < C) o[tco * C + tci + 1] = acc01;
    if (tco + 1 < C) {
      o[(tco + 1) * C + tci] = acc10;
      if (tci + 1 < C) o[(tco + 1) * C + tci + 1] = acc11;
    }
  }
}

// the whole filter must fit in shared memory: k*k*C*(C+1) floats (27 KB for the 4x4 stages at C = 20; the 16x16 stage
// at C = 20 needs 430 KB and is therefore only reachable through the tensor-core path or with few classes)
static size_t upscore_smem_bytes(int C, int s, bool padded) {
  return static_cast<size_t>(4) * s * s * C * (padded ? C + 1 : C) * sizeof(float);
}
cudaError_t launch_upscore_fwd(const float* x, const float* T, const float* bias, const float* skip, float* y, int N,
                               int h, int w, int C, int s, cudaStream_t st) {
  const size_t sm = upscore_smem_bytes(C, s, true);
  if (sm > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(upscore_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(sm));
  if (e != cudaSuccess) return e;
  const size_t total = static_cast<size_t>(N) * h * s * w * s * C;
  { (void)launch_k(upscore_fwd_kernel, dim3(grid_for(total, 256, 148 * 4)), dim3(256), sm, st, x, T, bias, skip, y, N, h, w, C, s); }
  return cudaGetLastError();
}
int upscore_bwd_splits(int N, int h, int w, int s) {
  const long long total = static_cast<long long>(N) * h * w;
  const int taps = 4 * s * s;
  long long want = (148 * 4 + taps - 1) / taps;
  long long maxs = (total + 255) / 256;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}
cudaError_t launch_upscore_bwd(const float* x, const float* T, const float* dy, float* dx, float* dT, float* dbias,
                               int N, int h, int w, int C, int s, float* ws, cudaStream_t st) {
  const int k = 2 * s;
  const long long Pout = static_cast<long long>(N) * h * s * w * s;
  // dbias
  {
    const int nb = head_bwd_blocks(Pout);
    const long long ppb = (Pout + nb - 1) / nb;
    { (void)launch_k(rows_colsum_kernel, dim3(nb), dim3(256), 0, st, dy, ws, Pout, C, ppb); }
    cudaError_t e = launch_colsum(ws, dbias, nb, C, 1.f, 0, st);
    if (e != cudaSuccess) return e;
  }
  // dT
  {
    const int nsplit = upscore_bwd_splits(N, h, w, s);
    dim3 grid(k * k, nsplit);
    { (void)launch_k(upscore_bwd_w_kernel, dim3(grid), dim3(256), 0, st, x, dy, ws + 128 * CMAX, N, h, w, C, s, nsplit); }
    cudaError_t e = launch_colsum(ws + 128 * CMAX, dT, nsplit, k * k * C * C, 1.f, 0, st);
    if (e != cudaSuccess) return e;
  }
  if (dx) {
    const size_t total = static_cast<size_t>(N) * h * w * C;
    const size_t sm = upscore_smem_bytes(C, s, false);
    if (sm > 200 * 1024) return cudaErrorInvalidConfiguration;
    cudaFuncSetAttribute(upscore_bwd_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
    { (void)launch_k(upscore_bwd_x_kernel, dim3(grid_for(total, 128, 148 * 8)), dim3(128), sm, st, dy, T, dx, N, h, w, C, s); }
  }
  return cudaGetLastError();
}
size_t upscore_bwd_ws_floats(int N, int h, int w, int C, int s) {
  return static_cast<size_t>(128) * CMAX + static_cast<size_t>(upscore_bwd_splits(N, h, w, s)) * 4 * s * s * C * C;
}

// ------------------------------------------------------------------------------------------------ softmax / xent
// One thread per pixel, no shared memory: a pixel's logits are C <= 32 consecutive floats of one row of a (possibly
// padded) tensor [N, H+2*pad, W+2*pad, CP] (the blocked output of the tensor-core upscore8 stage), so a thread's
// float4 loads / stores touch only the sectors of its own row: the traffic is the minimum, every access is
// independent of every other thread's, and many of them are in flight per thread.  (The earlier version staged
// 128-pixel runs through shared memory behind two block barriers; it sat at 1.5 TB/s, bound by exposed latency.)
// Labels (bool one-hot, dense), softmax and argmax outputs are dense.  Per-class sums of dlogits (the bias gradient of
// the last transposed convolution) are accumulated in registers and reduced once per CTA.
constexpr int kLossThreads = 128;
// CM = compile-time bound on the class count (register arrays are sized by it: 4, 20 or 32)
template <int CM>
__global__ void __launch_bounds__(kLossThreads)
softmax_xent_kernel(const float* __restrict__ z, const uint8_t* __restrict__ labels, float* __restrict__ loss_sum,
                    float* __restrict__ dz, float* __restrict__ dbias, float* __restrict__ sm,
                    long long* __restrict__ amax, int N, int H, int W, int C, int CP, int pad, float gscale) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sred[kLossThreads / 32];
  __shared__ float sdb[CMAX];
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const long long P = static_cast<long long>(N) * H * W;
  const bool vec = (C & 3) == 0 && (CP & 3) == 0;   // float4 path (C = 20: 5 loads of the row's 8 float4)
  const bool want_db = dz && dbias;
  float local_loss = 0.f;
  float dbr[CM];
#pragma unroll
  for (int c = 0; c < CM; ++c) dbr[c] = 0.f;
  if (threadIdx.x < CMAX) sdb[threadIdx.x] = 0.f;
  for (long long p = blockIdx.x * static_cast<long long>(kLossThreads) + threadIdx.x; p < P;
       p += static_cast<long long>(gridDim.x) * kLossThreads) {
    const int x = static_cast<int>(p % W);
    const int y = static_cast<int>((p / W) % H);
    const int n = static_cast<int>(p / (static_cast<long long>(W) * H));
    const size_t zoff = ((static_cast<size_t>(n) * Hp + y + pad) * Wp + x + pad) * CP;
    float v[CM];
    if (vec) {
      const float4* z4 = reinterpret_cast<const float4*>(z + zoff);
#pragma unroll
      for (int m = 0; m < CM / 4; ++m)
        if (4 * m < C) {
          const float4 q = __ldg(z4 + m);
          v[4 * m] = q.x;
          v[4 * m + 1] = q.y;
          v[4 * m + 2] = q.z;
          v[4 * m + 3] = q.w;
        }
    } else {
#pragma unroll
      for (int c = 0; c < CM; ++c)
        if (c < C) v[c] = __ldg(z + zoff + c);
    }
    float yv[CM];
    if (labels) {
      const uint8_t* lp = labels + p * C;
      if (vec) {   // C % 4 == 0: the pixel's C label bytes are 4-byte aligned
        const uint32_t* l4 = reinterpret_cast<const uint32_t*>(lp);
#pragma unroll
        for (int m = 0; m < CM / 4; ++m)
          if (4 * m < C) {
            const uint32_t w = __ldg(l4 + m);
            yv[4 * m] = static_cast<float>(w & 0xffu);
            yv[4 * m + 1] = static_cast<float>((w >> 8) & 0xffu);
            yv[4 * m + 2] = static_cast<float>((w >> 16) & 0xffu);
            yv[4 * m + 3] = static_cast<float>(w >> 24);
          }
      } else {
#pragma unroll
        for (int c = 0; c < CM; ++c)
          if (c < C) yv[c] = static_cast<float>(lp[c]);
      }
    }
    float mx = -INFINITY;
    int am = 0;
#pragma unroll
    for (int c = 0; c < CM; ++c)
      if (c < C && v[c] > mx) {
        mx = v[c];
        am = c;
      }
    float se = 0.f;
    float e[CM];
#pragma unroll
    for (int c = 0; c < CM; ++c) {
      e[c] = (c < C) ? expf(v[c] - mx) : 0.f;
      se += e[c];
    }
    const float inv = 1.f / se;
    if (amax) amax[p] = am;
    bool have_dz = false;
    if (labels) {
      const float lse = mx + logf(se);
      float ysum = 0.f, yz = 0.f;
#pragma unroll
      for (int c = 0; c < CM; ++c)
        if (c < C) {
          ysum += yv[c];
          yz += yv[c] * v[c];
        }
      local_loss += ysum * lse - yz;
      if (dz) {
#pragma unroll
        for (int c = 0; c < CM; ++c) {
          e[c] = (c < C) ? (e[c] * inv * ysum - yv[c]) * gscale : 0.f;
          if (want_db && c < C) dbr[c] += e[c];
        }
        have_dz = true;
      }
    }
    if (have_dz) {
      // the whole padded row is written (zeros beyond C): the transposed-convolution gradient GEMMs read CP channels
      if ((CP & 3) == 0) {
        float4* d4 = reinterpret_cast<float4*>(dz + zoff);
#pragma unroll
        for (int m = 0; m < CMAX / 4; ++m)
          if (4 * m < CP) {
            if (m < CM / 4)
              d4[m] = make_float4(e[4 * (m < CM / 4 ? m : 0)], e[4 * (m < CM / 4 ? m : 0) + 1],
                                  e[4 * (m < CM / 4 ? m : 0) + 2], e[4 * (m < CM / 4 ? m : 0) + 3]);
            else
              d4[m] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
      } else {
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < CP) dz[zoff + c] = c < CM ? e[c < CM ? c : 0] : 0.f;
      }
    } else if (sm) {
      float* dst = sm + p * C;   // dense [.., C]
      if (vec) {
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int m = 0; m < CM / 4; ++m)
          if (4 * m < C)
            d4[m] = make_float4(e[4 * m] * inv, e[4 * m + 1] * inv, e[4 * m + 2] * inv, e[4 * m + 3] * inv);
      } else {
#pragma unroll
        for (int c = 0; c < CM; ++c)
          if (c < C) dst[c] = e[c] * inv;
      }
    }
  }
  __syncthreads();
  if (loss_sum) {
    for (int o = 16; o > 0; o >>= 1) local_loss += __shfl_xor_sync(0xffffffffu, local_loss, o);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = local_loss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int k = 0; k < kLossThreads / 32; ++k) t += sred[k];
      atomicAdd(loss_sum, t);
    }
  }
  if (want_db) {
#pragma unroll
    for (int c = 0; c < CM; ++c)
      if (c < C) {
        float a = dbr[c];
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sdb[c], a);
      }
    __syncthreads();
    if (threadIdx.x < C) atomicAdd(dbias + threadIdx.x, sdb[threadIdx.x]);
  }
}
cudaError_t launch_softmax_xent(const float* z, const uint8_t* labels, float* loss_sum, float* dz, float* dbias,
                                float* sm, long long* amax, int N, int H, int W, int C, int CP, int pad, float gscale,
                                cudaStream_t st) {
  const long long P = static_cast<long long>(N) * H * W;
  const long long want = (P + kLossThreads - 1) / kLossThreads;
  const int blocks = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
#define CALL(CM) { (void)launch_k(softmax_xent_kernel<CM>, dim3(blocks), dim3(kLossThreads), 0, st, z, labels, loss_sum, dz, dbias, sm, amax, N, H, W, C, CP, pad, gscale); }
  if (C <= 4) CALL(4)
  else if (C <= 20) CALL(20)
  else CALL(32)
#undef CALL
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ phase-GEMM operands
// Operand packing for the tensor-core transposed convolution (capi.cu: fcn8_upscore_tc_*).  T[a][b][co][ci] (TF
// layout), a = dy + s*(1-ty), b = dx + s*(1-tx):
//   w_fwd[(dy,dx,co)][(ty,tx,ci32)]   rows s*s*CP, row length 128   (zero for co >= C or ci >= C)
//   w_dx [ci64][(ty,tx,dy,dx,co)]     rows 64,     row length 4*s*s*CP
//   bias_big[(dy,dx,co)] = bias[co]
// *_lo != nullptr: 3xTF32 split (hi operand = the fp32 value itself, the MMA truncates it; lo = rna(w - trunc(w)));
// otherwise the single-pass operand rna_tf32(w).
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ void put_split(float* hi, float* lo, size_t i, float v) {
  if (lo) {
    hi[i] = v;
    lo[i] = rna_tf32(v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u));
  } else {
    hi[i] = rna_tf32(v);
  }
}
__global__ void upscore_pack_kernel(const float* __restrict__ T, const float* __restrict__ bias, float* w_fwd,
                                    float* w_fwd_lo, float* w_dx, float* w_dx_lo, float* bias_big, int C, int CP,
                                    int s) {
  pdl_launch_dependents();
  pdl_wait();
  const int k = 2 * s;
  const int ncols = s * s * CP;
  const size_t n_fwd = static_cast<size_t>(ncols) * 128;
  const size_t n_dx = static_cast<size_t>(64) * 4 * ncols;
  const size_t total = n_fwd + n_dx + ncols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (i < n_fwd) {
      const int kk = static_cast<int>(i % 128), col = static_cast<int>(i / 128);
      const int ci = kk % 32, tap = kk / 32, ty = tap >> 1, tx = tap & 1;
      const int co = col % CP, dx = (col / CP) % s, dy = col / (CP * s);
      float v = 0.f;
      if (co < C && ci < C) v = T[((static_cast<size_t>(dy + s * (1 - ty)) * k + dx + s * (1 - tx)) * C + co) * C + ci];
      put_split(w_fwd, w_fwd_lo, i, v);
    } else if (i < n_fwd + n_dx) {
      const size_t j = i - n_fwd;
      const int kk = static_cast<int>(j % (4 * ncols)), ci = static_cast<int>(j / (4 * ncols));
      const int tap = kk / ncols, col = kk % ncols, ty = tap >> 1, tx = tap & 1;
      const int co = col % CP, dx = (col / CP) % s, dy = col / (CP * s);
      float v = 0.f;
      if (co < C && ci < C) v = T[((static_cast<size_t>(dy + s * (1 - ty)) * k + dx + s * (1 - tx)) * C + co) * C + ci];
      put_split(w_dx, w_dx_lo, j, v);
    } else {
      const int col = static_cast<int>(i - n_fwd - n_dx);
      const int co = col % CP;
      bias_big[col] = co < C ? bias[co] : 0.f;
    }
  }
}
cudaError_t launch_upscore_pack(const float* T, const float* bias, float* w_fwd, float* w_fwd_lo, float* w_dx,
                                float* w_dx_lo, float* bias_big, int C, int CP, int s, cudaStream_t st) {
  const size_t total = static_cast<size_t>(s) * s * CP * (128 + 256 + 1);
  { (void)launch_k(upscore_pack_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, T, bias, w_fwd, w_fwd_lo, w_dx, w_dx_lo, bias_big, C, CP, s); }
  return cudaGetLastError();
}
// Interior of a padded blocked transposed-conv output (+ the skip tensor): f[n,y,x,c] = zp[n,y+pad,x+pad,c] + skip[n,y,x,c]
// for c < C, 0 for C <= c < ldf (fcn8s_tensorflow.py:213,224: the tf.add of the skip connections).
__global__ void upscore_gather_kernel(const float* __restrict__ zp, const float* __restrict__ skip,
                                      float* __restrict__ f, int N, int H, int W, int C, int CP, int pad, int ldf,
                                      int ld_skip) {
  pdl_launch_dependents();
  pdl_wait();
  const int Wp = W + 2 * pad, Hp = H + 2 * pad;
  const size_t total = static_cast<size_t>(N) * H * W * ldf;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % ldf);
    size_t r = i / ldf;
    const int x = static_cast<int>(r % W);
    r /= W;
    const int y = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    float v = 0.f;
    if (c < C) {
      v = zp[((static_cast<size_t>(n) * Hp + y + pad) * Wp + x + pad) * CP + c];
      if (skip) v += skip[((static_cast<size_t>(n) * H + y) * W + x) * ld_skip + c];
    }
    f[i] = v;
  }
}
// The reverse for the gradient: dzp interior = g (border and channels >= C stay zero), db[c] += sum over pixels of g.
__global__ void upscore_scatter_kernel(const float* __restrict__ g, float* __restrict__ dzp, float* __restrict__ db,
                                       int N, int H, int W, int C, int CP, int pad, int ldg) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sdb[CMAX];
  if (threadIdx.x < CMAX) sdb[threadIdx.x] = 0.f;
  __syncthreads();
  const int Wp = W + 2 * pad, Hp = H + 2 * pad;
  const size_t total = static_cast<size_t>(N) * H * W * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    size_t r = i / C;
    const int x = static_cast<int>(r % W);
    r /= W;
    const int y = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    const float v = g[((static_cast<size_t>(n) * H + y) * W + x) * ldg + c];
    dzp[((static_cast<size_t>(n) * Hp + y + pad) * Wp + x + pad) * CP + c] = v;
    if (db) atomicAdd(&sdb[c], v);
  }
  __syncthreads();
  if (db && threadIdx.x < C && sdb[threadIdx.x] != 0.f) atomicAdd(db + threadIdx.x, sdb[threadIdx.x]);
}
cudaError_t launch_upscore_gather(const float* zp, const float* skip, float* f, int N, int H, int W, int C, int CP,
                                  int pad, int ldf, int ld_skip, cudaStream_t st) {
  const size_t total = static_cast<size_t>(N) * H * W * ldf;
  { (void)launch_k(upscore_gather_kernel, dim3(grid_for(total, 256, 148 * 8)), dim3(256), 0, st, zp, skip, f, N, H, W, C, CP, pad, ldf, ld_skip); }
  return cudaGetLastError();
}
cudaError_t launch_upscore_scatter(const float* g, float* dzp, float* db, int N, int H, int W, int C, int CP, int pad,
                                   int ldg, cudaStream_t st) {
  const size_t total = static_cast<size_t>(N) * H * W * C;
  { (void)launch_k(upscore_scatter_kernel, dim3(grid_for(total, 256, 148 * 2)), dim3(256), 0, st, g, dzp, db, N, H, W, C, CP, pad, ldg); }
  return cudaGetLastError();
}

// dT[a][b][co][ci] = sum_split src[split][(ty,tx,ci)][(dy,dx,co)]
__global__ void upscore_unpack_dw_kernel(const float* __restrict__ src, int nsplit, float* __restrict__ dT, int C,
                                         int CP, int s) {
  pdl_launch_dependents();
  pdl_wait();
  const int k = 2 * s;
  const int ncols = s * s * CP;
  const int total = k * k * C * C;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ci = i % C, co = (i / C) % C, b = (i / (C * C)) % k, a = i / (C * C * k);
  const int ty = a < s ? 1 : 0, tx = b < s ? 1 : 0;
  const int dy = a - s * (1 - ty), dx = b - s * (1 - tx);
  const size_t off = static_cast<size_t>((ty * 2 + tx) * 32 + ci) * ncols + (dy * s + dx) * CP + co;
  float acc = 0.f;
  for (int sp = 0; sp < nsplit; ++sp) acc += src[static_cast<size_t>(sp) * 128 * ncols + off];
  dT[i] = acc;
}
cudaError_t launch_upscore_unpack_dw(const float* src, int nsplit, float* dT, int C, int CP, int s, cudaStream_t st) {
  const int total = 4 * s * s * C * C;
  { (void)launch_k(upscore_unpack_dw_kernel, dim3((total + 255) / 256), dim3(256), 0, st, src, nsplit, dT, C, CP, s); }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ confusion matrix
__global__ void confusion_kernel(const long long* __restrict__ pred, const uint8_t* __restrict__ onehot,
                                 unsigned long long* __restrict__ conf, long long P, int C) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ unsigned int hist[CMAX * CMAX];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < P;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    int gt = 0;
    uint8_t best = 0;
    for (int c = 0; c < C; ++c) {
      const uint8_t v = onehot[p * C + c];
      if (v > best) {
        best = v;
        gt = c;
      }
    }
    const int pr = static_cast<int>(pred[p]);
    atomicAdd(&hist[gt * C + pr], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x)
    if (hist[i]) atomicAdd(&conf[i], static_cast<unsigned long long>(hist[i]));
}
cudaError_t launch_confusion(const long long* pred, const uint8_t* onehot, unsigned long long* conf, long long P, int C,
                             cudaStream_t st) {
  { (void)launch_k(confusion_kernel, dim3(grid_for(P, 256, 148 * 4)), dim3(256), 0, st, pred, onehot, conf, P, C); }
  return cudaGetLastError();
}

}  // namespace fcn8
