// Issue-rate probe for tcgen05.mma kind::f16 (bf16 operands, fp32 accumulate, M = 128) on sm_100a.
//
// Question it answers: the filter-gradient GEMMs have pixels as the reduction
// dimension of NHWC tensors, so BOTH shared-memory operands are MN-major; they plateau at 55-60 % of the tensor
// pipe while the K-major forward / dgrad GEMMs reach 80-90 %.  Is that the hardware's rate for MN-major operands
// or the kernel's pipeline?  This probe takes loads, barriers and the epilogue out of the picture: one CTA per SM,
// operands resident in shared memory (filled once), one thread issues a long train of MMAs over the same descriptor
// pattern the kernels use and the CTA measures clock64() from the first issue to the commit's arrival.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fcn8s_tensorflow_b200/csrc \
//        scripts/mma_rate_probe.cu -o gpurun_out/mma_rate_probe && gpurun_out/mma_rate_probe
//
// Output: one line per configuration with cycles per MMA (mean / max over CTAs) and the implied dense TFLOP/s.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace fcn8;

constexpr uint32_t kMmasPerBlk = 4;
constexpr uint32_t kBlks = 3;

struct ProbeArgs {
  uint32_t idesc;
  uint64_t adesc0, bdesc0;    // descriptors without the start address
  uint32_t a_adv, b_adv;      // start-address advance per MMA inside a block, bytes
  uint32_t a_blk, b_blk;      // start-address advance per block (stage), bytes
  uint32_t blks;              // blocks cycled through (stages)
  uint32_t iters;             // passes over all blocks
  uint32_t a_bytes;           // size of the A region (B follows)
  uint32_t a_tmem;            // 1: A operand read from tensor memory (columns 256..), 8 columns per MMA
  uint32_t commit_every_blk;  // 1: tcgen05.commit after every block (as the pipelined kernels do)
};

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// pair == false: one CTA per SM, cta_group::1, M = 128.
// pair == true : clusters of two CTAs (one SM pair), cta_group::2, M = 256 = 128 rows of A per CTA, each CTA holds
//                half of B (N/2 rows); the leader CTA issues, both tensor cores run, D = 128 lanes x N columns per CTA.
template <bool A_TMEM, bool PAIR, bool TF32 = false>
__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs g, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done_bar, blk_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  // operands: small finite bf16 values (0x3c00 +- noise ~ 0.0078), written by all threads
  const uint32_t total = g.a_bytes + g.blks * g.b_blk;
  for (uint32_t i = threadIdx.x * 4; i < total; i += blockDim.x * 4) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    h ^= h >> 13;
    *reinterpret_cast<uint32_t*>(smem + i) = 0x3c003c00u | (h & 0x807f807fu);
  }
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    mbar_init(&blk_bar, 1);
    fence_mbar_init();
  }
  if (PAIR) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc<512>(&tmem_slot);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (PAIR) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && rank == 0) {
    const uint32_t sa = smem_u32(smem);
    const uint32_t sb = sa + g.a_bytes;
    // descriptors of one pass (kBlks blocks x 4 MMAs) are loop-invariant: keep them in registers so that the
    // issue loop is nothing but the MMAs (the probe must not be bound by address arithmetic)
    uint64_t ad[kBlks * kMmasPerBlk], bd[kBlks * kMmasPerBlk];
#pragma unroll
    for (uint32_t b = 0; b < kBlks; ++b)
#pragma unroll
      for (uint32_t k = 0; k < kMmasPerBlk; ++k) {
        ad[b * kMmasPerBlk + k] = g.adesc0 | static_cast<uint64_t>(((sa + b * g.a_blk + k * g.a_adv) & 0x3FFFF) >> 4);
        bd[b * kMmasPerBlk + k] = g.bdesc0 | static_cast<uint64_t>(((sb + b * g.b_blk + k * g.b_adv) & 0x3FFFF) >> 4);
      }
    t0 = clock64();
    // the whole warp runs the uniform loop and one elected lane issues (a loop inside `if (lane == 0)` makes the
    // compiler re-elect and broadcast before every MMA: profiles/r01_issue_loop.md)
    for (uint32_t it = 0; it < g.iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (uint32_t b = 0; b < kBlks; ++b) {
#pragma unroll
          for (uint32_t k = 0; k < kMmasPerBlk; ++k) {
            const uint32_t i = b * kMmasPerBlk + k;
            const uint32_t acc = it | i;
            if (PAIR) {
              asm volatile(
                  "{\n\t"
                  ".reg .pred p;\n\t"
                  "setp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
                  "}\n" ::"r"(tmem_base),
                  "l"(ad[i]), "l"(bd[i]), "r"(g.idesc), "r"(acc), "r"(0u)
                  : "memory");
            } else if (A_TMEM) {
              umma_f16_ts(tmem_base, tmem_base + 256 + (i & 15) * 8, bd[i], g.idesc, acc);
            } else {
              if (TF32)
                umma_tf32(tmem_base, ad[i], bd[i], g.idesc, acc);
              else
                umma_f16(tmem_base, ad[i], bd[i], g.idesc, acc);
            }
          }
          if (g.commit_every_blk) {
            if (PAIR)
              asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                               smem_u32(&blk_bar))
                           : "memory");
            else
              umma_commit(&blk_bar);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      if (PAIR)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&done_bar))
                     : "memory");
      else
        umma_commit(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);
    t1 = clock64();
    tc_fence_after();
  }
  __syncthreads();
  if (PAIR) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (warp == 0) {
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else
      tmem_dealloc<512>(tmem_base);
  }
}

struct Config {
  const char* name;
  int a_mn, b_mn, N, a_tmem, commit;
  int M = 128;
  int pair = 0;
  int tf32 = 0;   // kind::tf32 (fp32-stored operands, K = 8 per MMA: the same 32 bytes per row)
};

typedef void (*KernelFn)(ProbeArgs, long long*);

static void launch(KernelFn fn, int grid, bool pair, int smem, const ProbeArgs& g, long long* d_cycles) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, fn, g, d_cycles);
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  if (prop.major != 10) {
    fprintf(stderr, "needs sm_100 (got %d.%d)\n", prop.major, prop.minor);
    return 1;
  }
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev);
  const int sms = prop.multiProcessorCount & ~1;
  long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * sms);
  cudaMemset(d_cycles, 0, sizeof(long long) * sms);
  const int kSmem = 200 * 1024;
  KernelFn k_ss = probe_kernel<false, false>, k_ts = probe_kernel<true, false>, k_pair = probe_kernel<false, true>;
  KernelFn k_tf32 = probe_kernel<false, false, true>;
  cudaFuncSetAttribute(k_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  cudaFuncSetAttribute(k_ss, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  cudaFuncSetAttribute(k_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);

  const Config configs[] = {
      {"M=128  A K-major   B K-major   N=256 (forward / dgrad)     ", 0, 0, 256, 0, 1},
      {"M=128  A K-major   B MN-major  N=256 (forward, HWIO weights)", 0, 1, 256, 0, 1},
      {"M=128  A MN-major  B K-major   N=256                        ", 1, 0, 256, 0, 1},
      {"M=128  A MN-major  B MN-major  N=256 (filter gradient)      ", 1, 1, 256, 0, 1},
      {"M=128  A MN-major  B MN-major  N=256, one commit at the end ", 1, 1, 256, 0, 0},
      {"M=128  A in TMEM   B MN-major  N=256                        ", 0, 1, 256, 1, 1},
      {"M=128  A in TMEM   B K-major   N=256                        ", 0, 0, 256, 1, 1},
      {"M=128  A K-major   B K-major   N=128                        ", 0, 0, 128, 0, 1},
      {"M=128  A MN-major  B MN-major  N=128                        ", 1, 1, 128, 0, 1},
      {"M=128  A in TMEM   B K-major   N=128                        ", 0, 0, 128, 1, 1},
      {"M=128  A K-major   B K-major   N=64                         ", 0, 0, 64, 0, 1},
      {"M=128  A MN-major  B MN-major  N=64 (halo filter gradient)  ", 1, 1, 64, 0, 1},
      {"M=128  A in TMEM   B K-major   N=64                         ", 0, 0, 64, 1, 1},
      {"M=64   A K-major   B K-major   N=256 (channels as rows)     ", 0, 0, 256, 0, 1, 64},
      {"M=64   A K-major   B K-major   N=128                        ", 0, 0, 128, 0, 1, 64},
      {"M=64   A K-major   B K-major   N=64                         ", 0, 0, 64, 0, 1, 64},
      {"PAIR M=256  A K-major   B K-major   N=256 (cta_group::2)    ", 0, 0, 256, 0, 1, 256, 1},
      {"PAIR M=256  A K-major   B MN-major  N=256                   ", 0, 1, 256, 0, 1, 256, 1},
      {"PAIR M=256  A MN-major  B MN-major  N=256                   ", 1, 1, 256, 0, 1, 256, 1},
      {"PAIR M=256  A K-major   B K-major   N=128                   ", 0, 0, 128, 0, 1, 256, 1},
      {"PAIR M=256  A K-major   B K-major   N=64                    ", 0, 0, 64, 0, 1, 256, 1},
      {"PAIR M=256  A MN-major  B MN-major  N=128                   ", 1, 1, 128, 0, 1, 256, 1},
      {"M=128  A K-major   B K-major   N=32                         ", 0, 0, 32, 0, 1},
      {"M=128  A in TMEM   B K-major   N=32                         ", 0, 0, 32, 1, 1},
      {"TF32  M=128  A K-major  B K-major  N=256 (K=8 per MMA)      ", 0, 0, 256, 0, 1, 128, 0, 1},
      {"TF32  M=128  A K-major  B K-major  N=128                    ", 0, 0, 128, 0, 1, 128, 0, 1},
      {"TF32  M=128  A K-major  B K-major  N=64 (decoder dgrad)     ", 0, 0, 64, 0, 1, 128, 0, 1},
  };
  printf("device %s, %d SMs, nominal SM clock %d MHz; bf16 x bf16 -> fp32, K=16 per MMA, 64-deep blocks, every SM busy\n",
         prop.name, sms, clock_khz / 1000);
  printf("floor = M*N/256 cycles per MMA per tensor core at the nominal 8192 MAC/clk/SM\n");
  printf("%-60s %10s %10s %10s %12s\n", "operands", "floor cyc", "mean cyc", "max cyc", "TFLOP/s");
  for (const Config& c : configs) {
    // one block = 64 reduction elements = 4 MMAs; 128-byte-swizzled canonical layouts:
    //   K-major : rows x 128 B (64 bf16 of K), 8-row atoms 1024 B apart (SBO), +32 B per MMA
    //   MN-major: per 64-wide MN chunk 64 K-rows x 128 B = 8192 B (LBO between chunks), 8-row atoms 1024 B apart
    //             (SBO), +2048 B per MMA (16 K-rows)
    // pair: per CTA 128 rows of A and N/2 rows of B
    const int rows_a = c.pair ? 128 : c.M;
    const int rows_b = c.pair ? c.N / 2 : c.N;
    ProbeArgs g{};
    g.idesc = make_idesc(c.tf32 ? 2u : 1u, c.a_mn, c.b_mn, c.M, c.N);
    g.adesc0 = c.a_mn ? make_smem_desc_sw128(0, 8192, 1024) : make_smem_desc_sw128(0, 16, 1024);
    g.bdesc0 = c.b_mn ? make_smem_desc_sw128(0, 8192, 1024) : make_smem_desc_sw128(0, 16, 1024);
    g.a_adv = c.a_mn ? 2048 : 32;
    g.b_adv = c.b_mn ? 2048 : 32;
    g.a_blk = rows_a * 128;
    g.b_blk = rows_b * 128;
    g.blks = kBlks;
    g.iters = 512;
    g.a_bytes = g.blks * g.a_blk;
    g.a_tmem = c.a_tmem;
    g.commit_every_blk = c.commit;
    KernelFn fn = c.tf32 ? k_tf32 : (c.pair ? k_pair : (c.a_tmem ? k_ts : k_ss));
    const int grid = sms;
    launch(fn, grid, c.pair, kSmem, g, d_cycles);   // warm-up
    launch(fn, grid, c.pair, kSmem, g, d_cycles);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e));
      return 2;
    }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double sum = 0, mx = 0;
    int timed = 0;
    for (int i = 0; i < grid; ++i) {
      if (c.pair && (i & 1)) continue;   // only the leader CTA of a pair issues and times
      sum += double(h[i]);
      if (double(h[i]) > mx) mx = double(h[i]);
      ++timed;
    }
    const double n_mma = double(g.iters) * g.blks * kMmasPerBlk;
    const double mean = sum / timed / n_mma, worst = mx / n_mma;
    // time the same launch with events for an absolute rate that does not depend on the SM clock reading
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) launch(fn, grid, c.pair, kSmem, g, d_cycles);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double issuers = c.pair ? grid / 2 : grid;
    const double flops = 10.0 * issuers * n_mma * 2.0 * c.M * c.N * (c.tf32 ? 8 : 16);
    const double floor_cyc = (c.pair ? 128.0 : double(c.M < 128 ? 128 : c.M)) * c.N / 256.0;   // tf32: same cycles, half the K
    printf("%-60s %10.1f %10.1f %10.1f %12.1f\n", c.name, floor_cyc, mean, worst, flops / (ms * 1e-3) / 1e12);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  cudaFree(d_cycles);
  return 0;
}
