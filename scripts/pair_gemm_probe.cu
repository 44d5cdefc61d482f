// Bring-up + rate probe for the CTA-pair (cta_group::2) GEMM main loop on sm_100a.
//
// Why: mma_rate_probe / tma_feed_probe / issue_loop_probe show that a single SM can issue a 128x256x16 bf16 MMA every
// 128 cycles (the floor) and that TMA can deliver a 48 KB k-block every 431 cycles, yet the 128x256 GEMM kernels
// retire an MMA only every ~175 cycles.  Both streams go through the same shared memory: per MMA the tensor core reads
// 4 KB of A + 8 KB of B and TMA writes another 12 KB = 24 KB per 128 cycles = 192 B/clk against ~128 B/clk of
// shared-memory bandwidth.  A CTA pair computes a 256x256 tile with each CTA holding its own 128 rows of A and HALF of
// B: 8 KB read + 8 KB written per MMA per SM = 128 B/clk.  This file is the smallest complete kernel with that
// protocol (TMA producer / MMA issuer / epilogue warps, persistent, double-buffered accumulators), verified against
// a CPU product and timed against the same kernel in single-CTA form.
//
//   D[M][N] (fp32) = A[M][K] (bf16, K-major) x B[N][K]^T (bf16, K-major), N = 256.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fcn8s_tensorflow_b200/csrc \
//        scripts/pair_gemm_probe.cu -o scripts/_build/pair_gemm_probe && scripts/_build/pair_gemm_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace fcn8;

constexpr int BN = 256;
constexpr int kThreads = 192;   // warp 0 producer, warp 1 MMA, warps 2-5 epilogue


// Barrier wait without the suspend-time hint of ptx.cuh's mbar_wait: a tight test_wait spin.  -DSPIN_WAIT=1 uses it for
// every wait of this file (question: do waits that are completed by a REMOTE CTA wake up late when the warp sleeps?)
#ifndef SPIN_WAIT
#define SPIN_WAIT 0
#endif
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
#if SPIN_WAIT
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 4000000000ll) __trap();
  }
#else
  mbar_wait(bar, parity);
#endif
}

template <bool PAIR>
struct Cfg {
  static constexpr int kBRows = PAIR ? BN / 2 : BN;            // rows of B per CTA
  static constexpr int kABytes = 128 * 128;                    // 128 rows x 64 bf16
  static constexpr int kBBytes = kBRows * 128;
  static constexpr int kStage = kABytes + kBBytes;             // 48 KB / 32 KB
  static constexpr int kStages = PAIR ? 6 : 4;
  static constexpr int kSmem = kStages * kStage + 1024 + 256;
};

template <bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ D,
            int m_tiles, int kblocks, int a_kblocks, int store) {
  using C = Cfg<PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStage);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* acc_full = empty_bar + C::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  const int unit = PAIR ? blockIdx.x >> 1 : blockIdx.x;        // scheduling unit: a CTA or a CTA pair
  const int units = PAIR ? gridDim.x >> 1 : gridDim.x;
  const int unit_tiles = PAIR ? (m_tiles + 1) / 2 : m_tiles;   // a pair owns two consecutive M tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], PAIR ? 2 : 1);   // pair: the leader's expect_tx arrive + the peer's plain arrive
      mbar_init(&empty_bar[s], 1);             // one (multicast) commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], PAIR ? 8 : 4);  // epilogue warps of both CTAs release the leader's accumulator
    }
    fence_mbar_init();
  }
  if (PAIR) cluster_sync_all();   // barriers of both CTAs exist before anybody signals across
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc<512>(tmem_slot);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (both CTAs of a pair)
    int stage = 0;
    uint32_t phase = 0;
    for (int u = unit; u < unit_tiles; u += units) {
      const int mt = PAIR ? 2 * u + rank : u;
      for (int kb = 0; kb < kblocks; ++kb) {
        wait_bar(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * C::kStage;
          uint8_t* sb = sa + C::kABytes;
          if (!PAIR) {
            mbar_expect_tx(&full_bar[stage], C::kStage);
            tma_load_2d(&map_a, &full_bar[stage], sa, (kb % a_kblocks) * 64, mt * 128);
            tma_load_2d(&map_b, &full_bar[stage], sb, kb * 64, 0);
          } else {
            const uint32_t lead_bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
            if (rank == 0)
              mbar_expect_tx(&full_bar[stage], 2 * C::kStage);
            else
              mbar_arrive_cluster(lead_bar);
            tma_load_2d_pair(&map_a, lead_bar, sa, (kb % a_kblocks) * 64, mt * 128);   // rows past M: zero-filled
            tma_load_2d_pair(&map_b, lead_bar, sb, kb * 64, rank * (BN / 2));
          }
        }
        __syncwarp();
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA only in a pair)
    if (rank == 0) {
      const uint32_t idesc = make_idesc(1u, 0u, 0u, PAIR ? 256u : 128u, BN);
      const uint64_t desc0 = make_smem_desc_sw128(0, 16, 1024);
      const uint32_t sbase = smem_u32(smem);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int u = unit; u < unit_tiles; u += units) {
        wait_bar(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          wait_bar(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = sbase + stage * C::kStage;
          const uint64_t adesc = desc0 | static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
          const uint64_t bdesc = desc0 | static_cast<uint64_t>(((sa + C::kABytes) & 0x3FFFF) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (PAIR)
                umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
              else
                umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
            }
            if (PAIR)
              umma_commit_pair(&empty_bar[stage], 3);
            else
              umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) {
          if (PAIR)
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
    }
  } else {
    // ------------------------------------------------ epilogue: TMEM -> registers -> global (each CTA its own rows)
    const int quarter = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t lead_acc_empty0 = PAIR ? map_to_cta(smem_u32(&acc_empty[0]), 0) : 0;
    for (int u = unit; u < unit_tiles; u += units) {
      const int mt = PAIR ? 2 * u + rank : u;
      wait_bar(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      float* dst = D + (static_cast<size_t>(mt) * 128 + quarter * 32 + lane) * BN;
      float keep = 0.f;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        if (store && mt < m_tiles) {
          float4* o4 = reinterpret_cast<float4*>(dst + c);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o4[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                __uint_as_float(v[4 * i + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) keep += __uint_as_float(v[i]);
        }
      }
      if (!store && keep == 123.456f && mt < m_tiles) dst[0] = keep;   // keeps the loads alive
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR)
          mbar_arrive_cluster(lead_acc_empty0 + as * 8);
        else
          mbar_arrive(&acc_empty[as]);
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // nobody leaves while the partner may still read its smem / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else
      tmem_dealloc<512>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn enc, void* ptr, int rows, int k, int box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(3);
  }
  return m;
}

template <bool PAIR>
static float launch(EncodeTiledFn enc, void* dA, void* dB, float* dD, int M, int K, int Ka, int grid, int store,
                    int reps) {
  using C = Cfg<PAIR>;
  const CUtensorMap ma = make_map(enc, dA, M, Ka, 128);
  const CUtensorMap mb = make_map(enc, dB, BN, K, C::kBRows);
  cudaFuncSetAttribute(gemm_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int m_tiles = (M + 127) / 128;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) cudaLaunchKernelEx(&cfg, gemm_kernel<PAIR>, ma, mb, dD, m_tiles, K / 64, Ka / 64, store);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s kernel: CUDA error %s\n", PAIR ? "pair" : "single", cudaGetErrorString(e));
    exit(2);
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  if (prop.major != 10) {
    fprintf(stderr, "needs sm_100\n");
    return 1;
  }
  const int sms = prop.multiProcessorCount & ~1;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 3;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);

  // ---- correctness: M = 640 (5 tiles: odd, the last pair has a dummy partner), K = 320, against a CPU product
  {
    const int M = 640, K = 320;
    std::vector<__nv_bfloat16> hA(size_t(M) * K), hB(size_t(BN) * K);
    std::vector<float> fA(hA.size()), fB(hB.size());
    uint32_t s = 12345;
    auto rnd = [&]() {
      s = s * 1664525u + 1013904223u;
      return float(int((s >> 9) & 0xFFF) - 2048) / 2048.f;
    };
    for (size_t i = 0; i < hA.size(); ++i) {
      hA[i] = __float2bfloat16(rnd());
      fA[i] = __bfloat162float(hA[i]);
    }
    for (size_t i = 0; i < hB.size(); ++i) {
      hB[i] = __float2bfloat16(rnd());
      fB[i] = __bfloat162float(hB[i]);
    }
    void *dA, *dB;
    float* dD;
    cudaMalloc(&dA, hA.size() * 2);
    cudaMalloc(&dB, hB.size() * 2);
    cudaMalloc(&dD, size_t(M) * BN * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    std::vector<double> ref(size_t(M) * BN);
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < BN; ++j) {
        double a = 0;
        for (int k = 0; k < K; ++k) a += double(fA[size_t(i) * K + k]) * fB[size_t(j) * K + k];
        ref[size_t(i) * BN + j] = a;
      }
    for (int pair = 0; pair < 2; ++pair) {
      cudaMemset(dD, 0xFF, size_t(M) * BN * 4);
      // small grids so that every CTA / pair walks several tiles (exercises the accumulator ring)
      if (pair)
        launch<true>(enc, dA, dB, dD, M, K, K, 4, 1, 1);
      else
        launch<false>(enc, dA, dB, dD, M, K, K, 2, 1, 1);
      std::vector<float> out(size_t(M) * BN);
      cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
      double worst = 0;
      for (size_t i = 0; i < out.size(); ++i) {
        const double d = std::fabs(double(out[i]) - ref[i]);
        if (!(d <= worst)) worst = d;   // NaN-propagating max
      }
      printf("%-7s kernel, M=%d K=%d: max |D - ref| = %.3e  %s\n", pair ? "pair" : "single", M, K, worst,
             worst < 1e-3 ? "OK" : "MISMATCH");
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
  }

  // ---- rate: every SM busy for 8 full waves; K = 2304 as a 3x3 convolution over 256 channels has it: the A tile is
  // 256 deep and re-read 9 times (L2 hits after the first pass), B (the weights) is shared by all CTAs
  {
    const int M = sms * 128 * 8, K = 2304, Ka = 256;
    void *dA, *dB;
    float* dD;
    cudaMalloc(&dA, size_t(M) * Ka * 2);
    cudaMalloc(&dB, size_t(BN) * K * 2);
    cudaMalloc(&dD, size_t(M) * BN * 4);
    cudaMemset(dA, 0x3c, size_t(M) * Ka * 2);   // bf16 0x3c3c = 0.0115
    cudaMemset(dB, 0x3c, size_t(BN) * K * 2);
    const double flop = 2.0 * M * BN * K;
    for (int store = 1; store >= 0; --store) {
      launch<false>(enc, dA, dB, dD, M, K, Ka, sms, store, 2);
      const float t1 = launch<false>(enc, dA, dB, dD, M, K, Ka, sms, store, 5);
      launch<true>(enc, dA, dB, dD, M, K, Ka, sms, store, 2);
      const float t2 = launch<true>(enc, dA, dB, dD, M, K, Ka, sms, store, 5);
      printf("M=%d N=256 K=%d %s: single CTA 128x256 tiles %.3f ms = %.0f TFLOP/s | CTA pairs 256x256 %.3f ms = %.0f "
             "TFLOP/s\n",
             M, K, store ? "fp32 stores   " : "no stores     ", t1, flop / t1 / 1e9, t2, flop / t2 / 1e9);
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
  }
  return 0;
}
