// L2 -> shared-memory feed-rate probe for the 128 x 256 GEMM tile on sm_100a (companion of mma_rate_probe.cu).
//
// mma_rate_probe shows that one SM retires a 128x256x16 bf16 MMA every 172 cycles whatever the operand majors, i.e.
// a 64-deep k-block (16 KB of A + 32 KB of B) every 687 cycles = 70 B/clk of operand traffic per SM.  The GEMM kernels
// sit at 61-68 % of that rate.  This probe measures what the TMA path delivers when NOTHING consumes the data:
// every CTA streams k-blocks through a 4-stage ring and a single thread recycles the stages.
//   unicast        : each CTA loads its own A box (64 x 128) and the whole B box (64 x 256)            48 KB from L2
//   unicast-32K    : A box + half a B box (what a 2-CTA pair would need per CTA)                       32 KB from L2
//   multicast pair : clusters of 2 CTAs with the same B tile; each CTA loads its A box and multicasts one half of
//                    the B box to both CTAs -> 48 KB arrive per CTA, 32 KB are read from L2 per CTA
// Variants: A tiles shared by many CTAs or one per CTA, re-read from L2 or streamed from HBM; 2-4 stages; the
// consumer releasing a stage at once or holding it for the 687 cycles its MMAs would take.  B (one 256-row tile x
// K = 4608) is shared by all CTAs, as the weights of a 256-channel layer are.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fcn8s_tensorflow_b200/csrc \
//        scripts/tma_feed_probe.cu -o scripts/_build/tma_feed_probe && scripts/_build/tma_feed_probe
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "ptx.cuh"

using namespace fcn8;

constexpr int kABytes = 128 * 128;   // 128 rows x 64 bf16
constexpr int kBBytes = 256 * 128;
constexpr int kStage = kABytes + kBBytes;
constexpr int kKBlocks = 72;         // K = 4608

__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// remote arrive with a CLUSTER-scope release (ptx.cuh's mbar_arrive_cluster uses the default CTA scope): mode 5
__device__ __forceinline__ void mbar_arrive_remote_release_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

// MODE 0 unicast 48 KB, 1 unicast 32 KB, 2 multicast pair,
// 3 multicast pair WITHOUT the cross-CTA release of the stages (unsafe for real data; prices the multicast alone),
// 4 pair with unicast loads but WITH the cross-CTA stage release (prices the remote barrier arrives alone),
// 5 as 4, the remote arrive with .release.cluster instead of the default CTA-scope release
template <int MODE, int kStages>
__global__ void __launch_bounds__(64, 1)
feed_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_bh, int reps, int a_tiles, int a_kblocks, int delay,
            long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
  const int warp = threadIdx.x >> 5;
  constexpr bool kCluster = MODE >= 2;
  constexpr bool kMulticast = MODE == 2 || MODE == 3;
  constexpr bool kRemoteRelease = MODE == 2 || MODE >= 4;
  const uint32_t rank = kCluster ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kRemoteRelease ? 2 : 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (kCluster) cluster_sync_all();
  const int row_a = (blockIdx.x % a_tiles) * 128;
  const uint32_t rx_bytes = MODE == 1 ? kABytes + kBBytes / 2 : kStage;
  long long t0 = clock64(), t1 = t0;
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r)
      for (int kb = 0; kb < kKBlocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * kStage;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&full_bar[stage], rx_bytes);
          tma_load_2d(&map_a, &full_bar[stage], sa, (kb % a_kblocks) * 64, row_a);
          if (MODE == 0 || MODE >= 4) {
            tma_load_2d(&map_b, &full_bar[stage], sb, kb * 64, 0);
          } else if (MODE == 1) {
            tma_load_2d(&map_bh, &full_bar[stage], sb, kb * 64, 0);
          } else {
            tma_load_2d_mc(&map_bh, &full_bar[stage], sb + rank * (kBBytes / 2), kb * 64, rank * 128, 0x3);
          }
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
  } else {
    int stage = 0;
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r)
      for (int kb = 0; kb < kKBlocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        if (delay) {   // stand-in for the MMAs that read this stage
          const long long t = clock64();
          while (clock64() - t < delay) {}
        }
        if (elect_one()) {
          mbar_arrive(&empty_bar[stage]);
          if (kRemoteRelease) {
            if (MODE == 5)
              mbar_arrive_remote_release_cluster(&empty_bar[stage], rank ^ 1);
            else
              mbar_arrive_cluster(map_to_cta(smem_u32(&empty_bar[stage]), rank ^ 1));
          }
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (kCluster) cluster_sync_all();
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

struct Variant {
  const char* name;
  int a_tiles;     // distinct A row tiles (148 = one per CTA, 16 = heavily shared)
  int a_kblocks;   // K extent of A in k-blocks; the k loop wraps over it (4 = a 256-channel activation re-read per tap)
  int delay;       // cycles the consumer holds a stage (687 = the MMAs of one k-block)
};

template <int MODE, int kStages>
static void run(const Variant& v, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mbh, int sms,
                long long* d_cycles) {
  const double rx_kb = MODE == 1 ? 32 : 48, l2_kb = (MODE == 0 || MODE >= 4) ? 48 : 32;
  const int reps = 10;
  const int smem = kStages * kStage + 1024;
  cudaFuncSetAttribute(feed_kernel<MODE, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MODE >= 2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, feed_kernel<MODE, kStages>, ma, mb, mbh, reps, v.a_tiles, v.a_kblocks, v.delay, d_cycles);
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, feed_kernel<MODE, kStages>, ma, mb, mbh, reps, v.a_tiles, v.a_kblocks, v.delay, d_cycles);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e));
    exit(2);
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double sum = 0;
  for (long long c : h) sum += double(c);
  const double blocks = double(reps) * kKBlocks;
  const double cyc = sum / sms / blocks;
  static const char* mode_name[6] = {"unicast 48K", "unicast 32K", "multicast pair", "mcast, local rel", "ucast, remote rel",
                                     "ucast, rel.cluster"};
  printf("%-14s %d stages  %-44s hold %4d: %6.0f cyc/k-block  arrive %6.1f B/clk/SM  L2 %6.2f TB/s chip (%.3f ms)\n",
         mode_name[MODE], kStages, v.name, v.delay, cyc, rx_kb * 1024 / cyc,
         l2_kb * 1024 * blocks * sms / (ms * 1e-3) / 1e12, ms);
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
  const int K = kKBlocks * 64;
  const size_t a_bytes = size_t(sms) * 128 * K * 2;   // 175 MB: larger than L2 when every CTA streams its own rows
  void *dA, *dB;
  cudaMalloc(&dA, a_bytes);
  cudaMalloc(&dB, size_t(256) * K * 2);
  cudaMemset(dA, 0x11, a_bytes);
  cudaMemset(dB, 0x11, size_t(256) * K * 2);
  long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * sms);
  const CUtensorMap ma = make_map(enc, dA, sms * 128, K, 128);
  const CUtensorMap mb = make_map(enc, dB, 256, K, 256);
  const CUtensorMap mbh = make_map(enc, dB, 256, K, 128);
  printf("device %s, %d CTAs (1 per SM), k-block = 64 bf16 deep (A 16 KB + B 32 KB); the MMAs of a k-block take 687 "
         "cycles\n", prop.name, sms);
  const Variant shared16 = {"A: 16 shared tiles, L2-resident", 16, kKBlocks, 0};
  const Variant own_l2 = {"A: one tile per CTA, 256 deep, re-read (L2)", sms, 4, 0};
  const Variant own_hbm = {"A: one tile per CTA, streamed once (HBM)", sms, kKBlocks, 0};
  Variant v;
  run<0, 4>(shared16, ma, mb, mbh, sms, d_cycles);
  run<0, 4>(own_l2, ma, mb, mbh, sms, d_cycles);
  run<0, 4>(own_hbm, ma, mb, mbh, sms, d_cycles);
  run<0, 2>(own_l2, ma, mb, mbh, sms, d_cycles);
  run<0, 3>(own_l2, ma, mb, mbh, sms, d_cycles);
  v = own_l2; v.delay = 687;
  run<0, 2>(v, ma, mb, mbh, sms, d_cycles);
  run<0, 3>(v, ma, mb, mbh, sms, d_cycles);
  run<0, 4>(v, ma, mb, mbh, sms, d_cycles);
  v = own_hbm; v.delay = 687;
  run<0, 3>(v, ma, mb, mbh, sms, d_cycles);
  run<0, 4>(v, ma, mb, mbh, sms, d_cycles);
  run<1, 4>(own_l2, ma, mb, mbh, sms, d_cycles);
  run<2, 4>(own_l2, ma, mb, mbh, sms, d_cycles);
  run<3, 4>(own_l2, ma, mb, mbh, sms, d_cycles);
  run<4, 4>(own_l2, ma, mb, mbh, sms, d_cycles);
  run<5, 4>(own_l2, ma, mb, mbh, sms, d_cycles);
  return 0;
}
