// Issue-loop probe: how fast can ONE warp feed tcgen05.mma (M=128, N=256, bf16, K-major, floor = 128 cycles per MMA)?
// mma_rate_probe.cu shows the tensor core retires such an MMA every 128.1 cycles when the issue loop is nothing but
// MMAs.  The GEMM kernels' loops also wait on a full barrier, fence, elect a lane, build descriptors from the stage
// index and commit per k-block; this probe adds those pieces one at a time (operands resident in shared memory, every
// barrier already complete) to price them.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fcn8s_tensorflow_b200/csrc \
//        scripts/issue_loop_probe.cu -o scripts/_build/issue_loop_probe && scripts/_build/issue_loop_probe
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace fcn8;


constexpr uint32_t kStages = 4;
constexpr uint32_t kStageBytes = 48 * 1024;   // A 16 KB + B 32 KB
constexpr uint32_t kABytes = 16 * 1024;

// V0: 16 MMAs + 4 commits per elected region, descriptors in registers (the reference point)
// V1: one elected region per k-block (4 MMAs + commit), descriptors in registers, stage loop unrolled
// V2: V1 + try_wait on a (complete) full barrier + tcgen05.fence::after_thread_sync per k-block
// V3: rolled stage loop, descriptors built from the runtime stage index (the kernels' shape), no barrier wait
// V4: V3 + barrier wait + fence  = the loop of conv_gemm_kernel today
// V5: V4 with the stage loop unrolled by kStages (compile-time stage offsets), still one elected region per k-block
// V6: V5, but the barrier wait of k-block i+1 is issued BEFORE the MMAs of k-block i (software-pipelined wait)
template <int V>
__global__ void __launch_bounds__(128, 1) issue_kernel(uint32_t idesc, uint32_t kblocks, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done_bar, empty_bar[kStages], full_bar[kStages];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x * 4; i < kStages * kStageBytes; i += blockDim.x * 4) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    h ^= h >> 13;
    *reinterpret_cast<uint32_t*>(smem + i) = 0x3c003c00u | (h & 0x807f807fu);
  }
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    for (uint32_t s = 0; s < kStages; ++s) {
      mbar_init(&empty_bar[s], 1);
      mbar_init(&full_bar[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t sbase = smem_u32(smem);
    const uint64_t desc0 = make_smem_desc_sw128(0, 16, 1024);
    t0 = clock64();
    if (V <= 2) {
      uint64_t ad[kStages], bd[kStages];
#pragma unroll
      for (uint32_t s = 0; s < kStages; ++s) {
        ad[s] = desc0 | static_cast<uint64_t>(((sbase + s * kStageBytes) & 0x3FFFF) >> 4);
        bd[s] = desc0 | static_cast<uint64_t>(((sbase + s * kStageBytes + kABytes) & 0x3FFFF) >> 4);
      }
      for (uint32_t kb = 0; kb < kblocks; kb += kStages) {
        if (V == 0) {
          if (elect_one()) {
#pragma unroll
            for (uint32_t s = 0; s < kStages; ++s) {
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k) umma_f16(tmem_base, ad[s] + 2 * k, bd[s] + 2 * k, idesc, kb | s | k);
              umma_commit(&empty_bar[s]);
            }
          }
          __syncwarp();
        } else {
#pragma unroll
          for (uint32_t s = 0; s < kStages; ++s) {
            if (V == 2) {
              mbar_wait(&full_bar[s], 1);   // parity of the phase before the first: completes at once
              tc_fence_after();
            }
            if (elect_one()) {
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k) umma_f16(tmem_base, ad[s] + 2 * k, bd[s] + 2 * k, idesc, kb | s | k);
              umma_commit(&empty_bar[s]);
            }
            __syncwarp();
          }
        }
      }
    } else if (V <= 4) {
      uint32_t stage = 0, phase = 0;
#pragma unroll 1
      for (uint32_t kb = 0; kb < kblocks; ++kb) {
        if (V == 4) {
          mbar_wait(&full_bar[stage], phase ^ 1 ^ phase);   // always the already-complete parity
          tc_fence_after();
        }
        const uint32_t sa = sbase + stage * kStageBytes;
        const uint64_t adesc = desc0 | static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
        const uint64_t bdesc = desc0 | static_cast<uint64_t>(((sa + kABytes) & 0x3FFFF) >> 4);
        if (elect_one()) {
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, kb | k);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else {
      uint32_t phase = 0;
      if (V == 6) mbar_wait(&full_bar[0], 1);
      for (uint32_t kb = 0; kb < kblocks; kb += kStages) {
#pragma unroll
        for (uint32_t s = 0; s < kStages; ++s) {
          if (V == 5) mbar_wait(&full_bar[s], 1);
          tc_fence_after();
          const uint32_t sa = sbase + s * kStageBytes;
          const uint64_t adesc = desc0 | static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
          const uint64_t bdesc = desc0 | static_cast<uint64_t>(((sa + kABytes) & 0x3FFFF) >> 4);
          if (elect_one()) {
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, kb | s | k);
            umma_commit(&empty_bar[s]);
          }
          __syncwarp();
          if (V == 6) mbar_wait(&full_bar[(s + 1) % kStages], 1);
        }
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
    mbar_wait(&done_bar, 0);
    t1 = clock64();
    tc_fence_after();
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

template <int V>
static void run(const char* name, int sms, long long* d_cycles) {
  const int smem = kStages * kStageBytes + 1024;
  const uint32_t kblocks = 2048;
  cudaFuncSetAttribute(issue_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const uint32_t idesc = make_idesc(1u, 0u, 0u, 128u, 256u);
  issue_kernel<V><<<sms, 128, smem>>>(idesc, kblocks, d_cycles);
  issue_kernel<V><<<sms, 128, smem>>>(idesc, kblocks, d_cycles);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: CUDA error %s\n", name, cudaGetErrorString(e));
    exit(2);
  }
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double sum = 0;
  for (long long v : h) sum += double(v);
  printf("%-100s %8.1f cycles per MMA (floor 128.0)\n", name, sum / sms / (kblocks * 4.0));
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  if (prop.major != 10) {
    fprintf(stderr, "needs sm_100\n");
    return 1;
  }
  const int sms = prop.multiProcessorCount;
  long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * sms);
  printf("device %s; one warp issues 128x256x16 bf16 MMAs, 4 per k-block + commit\n", prop.name);
  run<0>("V0 16 MMAs + 4 commits per elected region, descriptors in registers", sms, d_cycles);
  run<1>("V1 one elected region per k-block, descriptors in registers, stage loop unrolled", sms, d_cycles);
  run<2>("V2 = V1 + full-barrier try_wait + tcgen05.fence::after_thread_sync per k-block", sms, d_cycles);
  run<3>("V3 rolled stage loop, descriptors from the runtime stage index, no barrier wait", sms, d_cycles);
  run<4>("V4 = V3 + barrier wait + fence (the kernels' loop today)", sms, d_cycles);
  run<5>("V5 = V4 with the stage loop unrolled by the stage count", sms, d_cycles);
  run<6>("V6 = V5 with the wait for k-block i+1 issued before... after the MMAs of k-block i are issued", sms, d_cycles);
  return 0;
}
