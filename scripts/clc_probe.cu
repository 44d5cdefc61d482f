// Probe (GPU): dynamic persistent scheduling with cluster launch control (clusterlaunchcontrol.try_cancel), the
// mechanism behind the DYN tile loops of csrc/conv_gemm.cuh.  A grid of one cluster (1 or 2 CTAs) per work unit is
// launched; every running cluster processes its own unit and then cancels not-yet-launched clusters and takes their
// units, one query in flight ahead of the unit being processed.  Checks that every unit is processed exactly once by
// each CTA rank, reports how many CTAs actually ran, and the cost of a query.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/clc_probe scripts/clc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, 0x989680;\nselp.u32 %0, 1, 0, P;\n}\n"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try(b, parity))
    if (clock64() - t0 > 2000000000ll) __trap();
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta));
  return r;
}
__device__ __forceinline__ uint32_t ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int CL>
__global__ void clc_kernel(int* count, int* ran, long long* qcycles, int spin) {
  constexpr int S = 2, WARPS = 4;
  __shared__ __align__(16) uint32_t resp[S][4];
  __shared__ uint64_t full[S], empty[S];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CL > 1 ? ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], WARPS * CL);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    atomicAdd(ran, 1);
  }
  __syncthreads();
  if (CL > 1) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  int tile = blockIdx.x / CL;
  long long qc = 0;
  for (int it = 0;; ++it) {
    const int s = it % S;                 // query number `it` asks for the unit after unit number `it`
    const uint32_t ph = (it / S) & 1;
    if (warp == 0 && rank == 0) {   // scheduler: query for the unit after this one
      mbar_wait(&empty[s], ph ^ 1);
      const long long t0 = clock64();
      if (lane < CL) {
        const uint32_t remote = CL > 1 ? mapa(smem_u32(&full[s]), lane) : smem_u32(&full[s]);
        if (CL > 1)
          asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], 16;" ::"r"(remote) : "memory");
        else
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(remote) : "memory");
      }
      __syncwarp();
      if (lane == 0) {
        if (CL > 1)
          asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
                       ::"r"(smem_u32(resp[s])), "r"(smem_u32(&full[s])) : "memory");
        else
          asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
                       ::"r"(smem_u32(resp[s])), "r"(smem_u32(&full[s])) : "memory");
      }
      qc += clock64() - t0;
    }
    // ---- the unit's work
    if (lane == 0 && warp == 1) atomicAdd(&count[tile * CL + rank], 1);
    const long long w0 = clock64();
    while (clock64() - w0 < spin) {
    }
    // ---- every warp of every CTA reads the answer
    const long long t1 = clock64();
    mbar_wait(&full[s], ph);
    if (warp == 0 && rank == 0) qc += clock64() - t1;
    uint32_t x, valid;
    asm volatile(
        "{\n.reg .pred p1;\n.reg .b128 r;\nld.shared.b128 r, [%2];\n"
        "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n"
        "selp.u32 %1, 1, 0, p1;\n"
        "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, _, _, _}, r;\n}\n"
        : "=r"(x), "=r"(valid) : "r"(smem_u32(resp[s])) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      if (CL > 1)
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(mapa(smem_u32(&empty[s]), 0)) : "memory");
      else
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
    if (!valid) break;
    tile = x / CL;
  }
  if (warp == 0 && rank == 0 && lane == 0) atomicAdd((unsigned long long*)qcycles, (unsigned long long)qc);
  if (CL > 1) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CL>
int run(int units, int spin) {
  int *count, *ran;
  long long* qc;
  cudaMalloc(&count, units * CL * sizeof(int));
  cudaMalloc(&ran, sizeof(int));
  cudaMalloc(&qc, sizeof(long long));
  cudaMemset(count, 0, units * CL * sizeof(int));
  cudaMemset(ran, 0, sizeof(int));
  cudaMemset(qc, 0, sizeof(long long));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units * CL);
  cfg.blockDim = dim3(128);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  cudaError_t e = cudaLaunchKernelEx(&cfg, clc_kernel<CL>, count, ran, qc, spin);
  cudaEventRecord(e1);
  cudaError_t e2 = cudaDeviceSynchronize();
  if (e != cudaSuccess || e2 != cudaSuccess) {
    printf("cluster %d: launch %s / sync %s\n", CL, cudaGetErrorString(e), cudaGetErrorString(e2));
    return 1;
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int* h = (int*)malloc(units * CL * sizeof(int));
  int hran;
  long long hq;
  cudaMemcpy(h, count, units * CL * sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemcpy(&hran, ran, sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemcpy(&hq, qc, sizeof(long long), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < units * CL; ++i) bad += h[i] != 1;
  printf("cluster %d, %6d units, %5d spin cycles: %5d CTAs ran, units processed exactly once by every rank: %s (%d bad), "
         "%.3f ms, scheduler cycles per unit (acquire + issue + wait for the answer) %.0f\n",
         CL, units, spin, hran, bad ? "NO" : "yes", bad, ms, (double)hq / units);
  free(h);
  return bad != 0;
}

int main() {
  int rc = 0;
  rc |= run<1>(148, 2000);
  rc |= run<1>(10000, 2000);
  rc |= run<1>(10000, 20000);
  rc |= run<2>(74, 2000);
  rc |= run<2>(5000, 2000);
  rc |= run<2>(5001, 20000);
  printf(rc ? "CLC_PROBE_FAILED\n" : "CLC_PROBE_OK\n");
  return rc;
}
