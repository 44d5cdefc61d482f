// conv1_1 from the uint8 image (conv1.cu): argument block and launchers shared with capi.cu.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fcn8 {

constexpr int kC1Threads = 576;      // 18 warps: TMA, MMA, 8 epilogue, 8 builders (conv1.cu)
constexpr int kC1Builder0 = 10;      // first builder warp
constexpr int kC1Tile = 128 * 128;   // bytes of one [128 rows][64 bf16] operand tile

struct Conv1Args {
  const uint8_t* img;   // [N,H,W,3] RGB
  int N, H, W;
  int tiles;            // N*H*W / 128
  int pair;             // 1: hi / lo operands (fp32-equivalent), 0: bf16
  // forward
  __nv_bfloat16* out;   // [N*H*W][out_ld]
  __nv_bfloat16* out_lo;
  const float* bias;
  int out_ld;
  float rz_c;           // per k-block compensation of the hi*hi segment (ConvGemmArgs::rz_c)
  // wgrad
  float* partial;       // [gridDim.x * (pair ? 2 : 1)][64][64]
};

cudaError_t launch_conv1_fwd(const CUtensorMap& w_hi, const CUtensorMap& w_lo, const Conv1Args& a, int grid, int dev,
                             cudaStream_t st);
cudaError_t launch_conv1_wgrad(const CUtensorMap& dy_hi, const CUtensorMap& dy_lo, const Conv1Args& a, int grid,
                               int dev, cudaStream_t st);

}  // namespace fcn8
