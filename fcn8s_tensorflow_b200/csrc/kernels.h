// Internal launcher declarations shared by the translation units of libfcn8s_sm100.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace fcn8 {

struct ConvGemmArgs;

// Every kernel launch of the library goes through this counter (fcn8_launch_count() in the C ABI): it is what
// bench.py reports as gpu_launches.
extern unsigned long long g_launch_count;
inline void count_launch() { __atomic_fetch_add(&g_launch_count, 1ull, __ATOMIC_RELAXED); }   // the feed thread launches too

// Programmatic dependent launch (OFF by default; fcn8_debug_set(6, 1) / FCN8_DEBUG=6=1 switches it on).  Every kernel
// of the library starts with pdl_launch_dependents() (the NEXT kernel of the stream may be scheduled as soon as all
// CTAs of this one have started, so that its launch latency and prologue fill the SMs this kernel's tail leaves idle)
// and executes pdl_wait() before its first access to global memory (= until the previous kernel has completed and its
// writes are visible); both are no-ops in a kernel launched with plain stream order.  Measured on the replayed CUDA
// graph of the c2 step: 15.65 ms with the attribute against 15.53 ms without (bf16 6.67 / 6.59 ms) -- the graph
// already removes the launch gaps, and early CTAs of the next kernel compete with the tail of the running one.
extern int g_pdl_off;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_off ? 0 : 1;
  count_launch();
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// elementwise.cu
cudaError_t launch_preprocess(const uint8_t* img, void* out, int N, int H, int W, int dtype, cudaStream_t st);
cudaError_t launch_maxpool_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, cudaStream_t st);
cudaError_t launch_maxpool_bwd(const void* x, const void* dy, void* dx, float* db, int N, int H, int W, int C,
                               int dtype, cudaStream_t st);
cudaError_t launch_expand_labels(const uint8_t* ids, uint8_t* onehot, long long pixels, int C, cudaStream_t st);
int bias_grad_blocks(long long P, int C);
cudaError_t launch_bias_grad(const void* dy, float* db, long long P, int C, int dtype, float* ws, cudaStream_t st);
cudaError_t launch_colsum(const float* ws, float* out, int nb, int C, float scale, int accumulate, cudaStream_t st);
cudaError_t launch_pack(const float* w, void* out, void* out_lo, int ksize, int Cin, int Cout, int CinPad, int mode,
                        int dtype, cudaStream_t st);
cudaError_t launch_split_tf32(const float* x, float* hi, float* lo, size_t n, cudaStream_t st);
cudaError_t launch_conv_splitk_reduce(const float* partial, int splits, size_t n_elems, int ldc,
                                      const ConvGemmArgs& g, int dtype, cudaStream_t st);
cudaError_t launch_wgrad_splitk_reduce(const float* partial, float* out, int splits, size_t rows_pad, int rows_valid,
                                       int ldc, int out_cols, float scale, cudaStream_t st);
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, size_t n, float lr_t, float b1, float b2,
                        float eps, float gscale, void* w_hi, void* w_lo, const float* lr_ptr, const void* g16,
                        cudaStream_t st);
cudaError_t launch_cast_bf16(const float* x, void* out, size_t n, cudaStream_t st);
cudaError_t launch_set_step_scalars(float* scalars, float lr_t, uint32_t seed, cudaStream_t st);
cudaError_t launch_shadow(const float* p, void* hi, void* lo, size_t n, cudaStream_t st);
cudaError_t launch_l2_reg(const float* w, float* g, float* loss, size_t n, float rate, cudaStream_t st);

// decoder.cu
cudaError_t launch_deconv_pack(const float* T, const float* bias, void* w_fwd, void* w_fwd_lo, void* w_dx, void* w_dx_lo,
                               float* bias_big, int C, int CP, int s, cudaStream_t st);
cudaError_t launch_head_pack(const float* K, const float* bias, int Cin, int C, void* w_hi, void* w_lo, float* bias64,
                             cudaStream_t st);
cudaError_t launch_deconv_unpack_dw(const float* src, int nsplit, float* dT, int C, int CP, int s, cudaStream_t st);
cudaError_t launch_confusion(const long long* pred, const uint8_t* onehot, unsigned long long* conf, long long P, int C,
                             cudaStream_t st);

}  // namespace fcn8
