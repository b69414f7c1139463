// tcgen05.mma issue-rate probe (B200 only): how many cycles does one bf16 MMA occupy the tensor
// pipe with NO operand loads (shared memory holds whatever it holds), for
//   cta_group::1  M = 128,            N = 64 / 128 / 256
//   cta_group::2  M = 256 (CTA pair), N = 64 / 128 / 256
// first on one CTA (pair) -> pipe cycles per instruction, then on every SM of the chip for ~50 ms
// -> the MMA-only chip throughput under the 1000 W power cap, i.e. the ceiling of any kernel
// built from that instruction shape.  Design input for the implicit-GEMM kernel.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/umma_rate_probe tools/umma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../mri_image_generation_b200/csrc/ptx.cuh"

using namespace mri;

template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc) : "memory");
}

template <int CG>
__global__ void __launch_bounds__(128, 1)
rate_kernel(int n, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    if (CG == 1) {
      tmem_alloc(smem_u32(&tmem_holder), 512);
      tmem_relinquish();
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(&tmem_holder)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(CG == 1 ? 128 : 256, (uint32_t)n);
    // A: 128 rows x 64 K (16 KB) at base; B: n (or n/2 per CTA) rows x 64 K at base + 16 KB
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t a_desc = umma_desc_k_sw128(base, 1024) + 2u * k;
        const uint64_t b_desc = umma_desc_k_sw128(base + 16384, 1024) + 2u * k;
        mma<CG>(tmem + (uint32_t)((it & 1) * 256), a_desc, b_desc, idesc, (it | k) != 0 ? 1u : 0u);
      }
    }
    if (CG == 1)
      umma_commit(smem_u32(&bar));
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(smem_u32(&bar), 0);
  if (threadIdx.x == 0 && rank == 0) {
    t1 = clock64();
    cycles[blockIdx.x / CG] = t1 - t0;
  }
  tc_fence_before();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (warp == 0) {
    if (CG == 1) tmem_dealloc(tmem, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int CG>
static void run(int n, int grid, int iters, const char* tag) {
  long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * 256);
  cudaMemset(d_cycles, 0, sizeof(long long) * 256);
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {  // first launch warms up
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<CG>, n, iters, d_cycles);
    cudaEventRecord(e1);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: CUDA error %s\n", tag, cudaGetErrorString(e));
      exit(1);
    }
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c = 0;
  cudaMemcpy(&c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
  const double n_mma = 4.0 * iters;
  const double flops_per_mma_per_sm = 2.0 * 128 * n * 16;  // per SM: 128 rows x n x 16
  const double ideal = flops_per_mma_per_sm / 8192.0;       // 8192 dense bf16 FLOP / clk / SM
  printf("%-34s grid %3d: %7.1f cycles / MMA (ideal %5.1f -> %5.1f%% of pipe peak)", tag, grid, c / n_mma,
         ideal, 100.0 * ideal / (c / n_mma));
  printf("   chip: %8.1f TFLOP/s over %.1f ms\n", grid * n_mma * flops_per_mma_per_sm / (ms * 1e-3) / 1e12, ms);
  cudaFree(d_cycles);
}

int main() {
  const int sizes[3] = {64, 128, 256};
  for (int n : sizes) run<1>(n, 1, 20000, n == 64 ? "cta_group::1 M=128 N=64" : n == 128 ? "cta_group::1 M=128 N=128" : "cta_group::1 M=128 N=256");
  for (int n : sizes) run<2>(n, 2, 20000, n == 64 ? "cta_group::2 M=256 N=64" : n == 128 ? "cta_group::2 M=256 N=128" : "cta_group::2 M=256 N=256");
  // whole chip, ~50-100 ms each: sustained MMA-only throughput under the power cap
  for (int n : sizes) if (n >= 128) run<1>(n, 148, 300000, n == 128 ? "chip cta_group::1 M=128 N=128" : "chip cta_group::1 M=128 N=256");
  for (int n : sizes) if (n >= 128) run<2>(n, 148, 300000, n == 128 ? "chip cta_group::2 M=256 N=128" : "chip cta_group::2 M=256 N=256");
  return 0;
}
