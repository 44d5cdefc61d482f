// Operand packing for the FCN-8s decoder GEMMs (fcn8s_tensorflow.py:164-235) and the stand-alone confusion-matrix
// kernel (:280-301).  The decoder's arithmetic itself -- score heads, transposed convolutions, loss, predictor -- runs
// in the tcgen05 GEMM kernels of conv_gemm.cuh (entry points: capi.cu, "Decoder on the tensor cores").
// Limit: num_classes <= 32 (one 32-column chunk of the loss epilogue holds a pixel's classes).
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

__device__ __forceinline__ void put_pair(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t i, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ------------------------------------------------------------------------------------------------ phase-GEMM operands
// T[a][b][co][ci] (TF layout of tf.layers.conv2d_transpose kernels), a = dy + s*(1-ty), b = dx + s*(1-tx):
//   w_fwd[(dy,dx,co)][(ty,tx,ci64)]   rows s*s*CP, row length 256   (zero for co >= C or ci >= C)
//   w_dx [ci64][(ty,tx,dy,dx,co)]     rows 64,     row length 4*s*s*CP
//   bias_big[(dy,dx,co)] = bias[co]
// bf16 hi (and lo = bf16(v - hi) when the lo pointers are given).
__global__ void deconv_pack_kernel(const float* __restrict__ T, const float* __restrict__ bias, __nv_bfloat16* w_fwd,
                                   __nv_bfloat16* w_fwd_lo, __nv_bfloat16* w_dx, __nv_bfloat16* w_dx_lo,
                                   float* bias_big, int C, int CP, int s) {
  pdl_launch_dependents();
  pdl_wait();
  const int k = 2 * s;
  const int ncols = s * s * CP;
  const size_t n_fwd = static_cast<size_t>(ncols) * 256;
  const size_t n_dx = static_cast<size_t>(64) * 4 * ncols;
  const size_t total = n_fwd + n_dx + ncols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (i < n_fwd) {
      const int kk = static_cast<int>(i % 256), col = static_cast<int>(i / 256);
      const int ci = kk % 64, tap = kk / 64, ty = tap >> 1, tx = tap & 1;
      const int co = col % CP, dx = (col / CP) % s, dy = col / (CP * s);
      float v = 0.f;
      if (co < C && ci < C) v = T[((static_cast<size_t>(dy + s * (1 - ty)) * k + dx + s * (1 - tx)) * C + co) * C + ci];
      put_pair(w_fwd, w_fwd_lo, i, v);
    } else if (i < n_fwd + n_dx) {
      const size_t j = i - n_fwd;
      const int kk = static_cast<int>(j % (4 * ncols)), ci = static_cast<int>(j / (4 * ncols));
      const int tap = kk / ncols, col = kk % ncols, ty = tap >> 1, tx = tap & 1;
      const int co = col % CP, dx = (col / CP) % s, dy = col / (CP * s);
      float v = 0.f;
      if (co < C && ci < C) v = T[((static_cast<size_t>(dy + s * (1 - ty)) * k + dx + s * (1 - tx)) * C + co) * C + ci];
      put_pair(w_dx, w_dx_lo, j, v);
    } else {
      const int col = static_cast<int>(i - n_fwd - n_dx);
      const int co = col % CP;
      bias_big[col] = co < C ? bias[co] : 0.f;
    }
  }
}
cudaError_t launch_deconv_pack(const float* T, const float* bias, void* w_fwd, void* w_fwd_lo, void* w_dx, void* w_dx_lo,
                               float* bias_big, int C, int CP, int s, cudaStream_t st) {
  const size_t total = static_cast<size_t>(s) * s * CP * (256 + 256 + 1);
  { (void)launch_k(deconv_pack_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, T, bias, static_cast<__nv_bfloat16*>(w_fwd), static_cast<__nv_bfloat16*>(w_fwd_lo), static_cast<__nv_bfloat16*>(w_dx), static_cast<__nv_bfloat16*>(w_dx_lo), bias_big, C, CP, s); }
  return cudaGetLastError();
}

// Score-head kernels K [1,1,Cin,C] (tf.layers.conv2d, fcn8s_tensorflow.py:173-200) as a TF-layout weight tensor with the
// class dimension zero-padded to 64: w[ci][c64] (bf16 hi / lo), read in place by fcn8_conv_gemm (w_mode 1: forward,
// w_mode 2: input gradient); bias64[c] = bias[c], zero beyond C.
__global__ void head_pack_kernel(const float* __restrict__ K, const float* __restrict__ bias, int Cin, int C,
                                 __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, float* bias64) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t total = static_cast<size_t>(Cin) * 64;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % 64);
    const size_t ci = i / 64;
    put_pair(w_hi, w_lo, i, c < C ? K[ci * C + c] : 0.f);
  }
  if (bias64 && blockIdx.x == 0 && threadIdx.x < 64) bias64[threadIdx.x] = (bias && threadIdx.x < C) ? bias[threadIdx.x] : 0.f;
}
cudaError_t launch_head_pack(const float* K, const float* bias, int Cin, int C, void* w_hi, void* w_lo, float* bias64,
                             cudaStream_t st) {
  { (void)launch_k(head_pack_kernel, dim3(grid_for(static_cast<size_t>(Cin) * 64, 256, 148 * 4)), dim3(256), 0, st, K, bias, Cin, C, static_cast<__nv_bfloat16*>(w_hi), static_cast<__nv_bfloat16*>(w_lo), bias64); }
  return cudaGetLastError();
}

// dT[a][b][co][ci] = sum_split src[split][(ty,tx,ci64)][(dy,dx,co)]
__global__ void deconv_unpack_dw_kernel(const float* __restrict__ src, int nsplit, float* __restrict__ dT, int C,
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
  const size_t off = static_cast<size_t>((ty * 2 + tx) * 64 + ci) * ncols + (dy * s + dx) * CP + co;
  float acc = 0.f;
  for (int sp = 0; sp < nsplit; ++sp) acc += src[static_cast<size_t>(sp) * 256 * ncols + off];
  dT[i] = acc;
}
cudaError_t launch_deconv_unpack_dw(const float* src, int nsplit, float* dT, int C, int CP, int s, cudaStream_t st) {
  const int total = 4 * s * s * C * C;
  { (void)launch_k(deconv_unpack_dw_kernel, dim3((total + 255) / 256), dim3(256), 0, st, src, nsplit, dT, C, CP, s); }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ confusion matrix
// Stand-alone conf[label * C + prediction] += 1 (the device analogue of the reference's only native code,
// cityscapesscripts/evaluation/addToConfusionMatrix_impl.c:3-16); the evaluation step itself accumulates the matrix in
// the epilogue of the upscore8 kernel (fcn8_deconv_loss).
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
