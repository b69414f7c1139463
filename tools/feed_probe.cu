// Operand-feed probe (B200 only): the main loop of the implicit-GEMM kernel stripped to its two
// warps -- a TMA producer filling a shared-memory ring from L2-resident data and an MMA issuer
// running one 128 x N x 64 k-step (64 / 16 tcgen05.mma of N = 256 or 2 x N = 128) per ring slot --
// with NO epilogue.  Per k-step the producer loads a 16 KB weight slab and, every `reuse`-th
// step, a 32 KB activation tile (reuse = 1: today's plain taps; 3: the kw-sharing tiles of
// `xreuse`; 9 / 27: a 2-D / 3-D halo tile; 0: never).  All 148 SMs run it for ~50 ms, so the
// result is the sustained chip rate under the power cap: what the feed path alone allows.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/feed_probe tools/feed_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../mri_image_generation_b200/csrc/ptx.cuh"

using namespace mri;

constexpr int kMaxWStages = 10;
constexpr int kXStages = 2;
constexpr int kWBytes = 16384;  // 128 rows x 64 bf16
constexpr int kXBytes = 32768;  // 256 rows x 64 bf16

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __launch_bounds__(192, 1)
feed_kernel(const __grid_constant__ CUtensorMap w_map, const __grid_constant__ CUtensorMap x_map,
            int steps, int reuse, int n_per_mma, int rows_per_cta, long long* cycles, int kWStages,
            int no_tma, int ksteps_per_slot, int flags) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxWStages + kXStages + 1];
  __shared__ uint32_t tmem_holder;
  __shared__ uint32_t done_flag;
  __shared__ float scratch[192];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_ring = base + kWStages * kWBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxWStages + s); };
  auto xempty_bar = [&](int s) { return bar0 + 8u * (2 * kMaxWStages + s); };
  const uint32_t done_bar = bar0 + 8u * (2 * kMaxWStages + kXStages);
  if (threadIdx.x == 0) {
    done_flag = 0u;
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < kXStages; ++s) mbar_init(xempty_bar(s), 1);
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_holder), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const int row0 = blockIdx.x * rows_per_cta;  // this CTA cycles over its own L2-resident rows
  long long t0 = clock64();
  if (warp == 0 && (flags & 32)) {
    // variant: one elected thread owns the producer loop
    if (elect_one_sync()) {
      int ws = 0, xs = 0, since_x = 0;
      uint32_t wphase = 0, xphase = 0;
      for (int i = 0; i < steps; ++i) {
        const bool need_x = reuse > 0 && since_x == 0;
        mbar_wait(empty_bar(ws), wphase ^ 1u);
        if (need_x) mbar_wait(xempty_bar(xs), xphase ^ 1u);
        if (no_tma) {
          mbar_arrive_expect_tx(full_bar(ws), 0u);
        } else {
          mbar_arrive_expect_tx(full_bar(ws), (uint32_t)kWBytes + (need_x ? (uint32_t)kXBytes : 0u));
          const int r = row0 + (i * 384) % (rows_per_cta - 384);
          tma_load_2d(base + ws * kWBytes, &w_map, full_bar(ws), 0, r);
          if (need_x) tma_load_2d(x_ring + xs * kXBytes, &x_map, full_bar(ws), 0, r + 128);
        }
        if (++ws == kWStages) { ws = 0; wphase ^= 1u; }
        if (reuse > 0 && ++since_x == reuse) {
          since_x = 0;
          if (++xs == kXStages) { xs = 0; xphase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    int ws = 0, xs = 0;
    uint32_t wphase = 0, xphase = 0;
    for (int i = 0; i < steps; ++i) {
      const bool need_x = reuse > 0 && (i % reuse) == 0;
      const bool next_needs_x = reuse > 0 && ((i + 1) % reuse) == 0;
      if (!(flags & 4)) mbar_wait(empty_bar(ws), wphase ^ 1u);
      if (need_x) mbar_wait(xempty_bar(xs), xphase ^ 1u);
      if (elect_one_sync()) {
        if (no_tma) {  // synchronisation only: the slot is declared full without moving data
          mbar_arrive_expect_tx(full_bar(ws), 0u);
        } else {
          mbar_arrive_expect_tx(full_bar(ws), (uint32_t)kWBytes + (need_x ? (uint32_t)kXBytes : 0u));
          const int r = row0 + (i * 384) % (rows_per_cta - 384);
          tma_load_2d(base + ws * kWBytes, &w_map, full_bar(ws), 0, r);
          if (need_x) tma_load_2d(x_ring + xs * kXBytes, &x_map, full_bar(ws), 0, r + 128);
        }
      }
      __syncwarp();
      if (++ws == kWStages) { ws = 0; wphase ^= 1u; }
      if (next_needs_x && ++xs == kXStages) { xs = 0; xphase ^= 1u; }
    }
  } else if (warp >= 2) {
    // "epilogue-like" co-resident warps (flags bit 6): dependent FMA / shared-memory work on all
    // four SM sub-partitions until the MMA warp is done -- do they slow the issue threads down?
    if (flags & 64) {
      float acc = (float)threadIdx.x;
      volatile uint32_t* flag = &done_flag;
      uint32_t it = 0;
      while (*flag == 0u) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaf(acc, 1.0001f, 0.5f);
        scratch[threadIdx.x] = acc + (float)(it++);
      }
      if (acc == 1.2345f) cycles[0] = 0;
    }
  } else if (flags & 8) {
    // variant: ONE elected thread runs the whole issue loop (waits, MMAs, commits); the ring
    // walk is unrolled over `kU` slots with compile-time slot offsets
    const uint32_t idesc = umma_idesc_bf16(128, 256u);
    if (elect_one_sync()) {
      constexpr int kU = 2;
      uint32_t wphase = 0;
      int ws = 0, xs = 0, since_x = 0;
      const uint64_t a_desc0 = umma_desc_k_sw128(base, 1024);
      const uint64_t b_desc0 = umma_desc_k_sw128(x_ring, 1024);
      for (int i = 0; i < steps; i += kU) {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          mbar_wait(full_bar(ws + u), wphase);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + (uint64_t)(((ws + u) * kWBytes) >> 4);
          const uint64_t b_desc = b_desc0 + (uint64_t)((xs * kXBytes) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (i | u | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(ws + u));
          if (reuse > 0 && ++since_x == reuse) {  // last tap of this activation tile
            since_x = 0;
            umma_commit(xempty_bar(xs));
            if (++xs == kXStages) xs = 0;
          }
        }
        ws += kU;
        if (ws == kWStages) { ws = 0; wphase ^= 1u; }
      }
      umma_commit(done_bar);
    }
    __syncwarp();
  } else {
    const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)n_per_mma);
    int ws = 0, xs = 0;
    uint32_t wphase = 0;
    for (int i = 0; i < steps; ++i) {
      const bool next_needs_x = reuse > 0 && ((i + 1) % reuse) == 0;
      if (!(flags & 1)) mbar_wait(full_bar(ws), wphase);   // flags bit 0: skip the full-barrier wait
      if (!(flags & 2)) tc_fence_after();                   // bit 1: skip the tcgen05 fence
      const uint64_t a_desc = umma_desc_k_sw128(base + ws * kWBytes, 1024);
      const uint32_t xa = x_ring + xs * kXBytes;
      if (elect_one_sync()) {
        if (flags & 16) {   // the original, fully unrolled body (N = 256, one k-step per slot)
          const uint64_t b_desc = umma_desc_k_sw128(xa, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (i | k) != 0 ? 1u : 0u);
        } else {
          for (int rep = 0; rep < ksteps_per_slot; ++rep)
            for (int h = 0; h < 256 / n_per_mma; ++h) {
              const uint64_t b_desc = umma_desc_k_sw128(xa + h * n_per_mma * 128, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem + (uint32_t)(h * n_per_mma), a_desc + 2u * k, b_desc + 2u * k, idesc,
                          (i | k) != 0 ? 1u : 0u);
            }
        }
        if (!(flags & 4)) umma_commit(empty_bar(ws));       // bit 2: no per-slot commit
        if (next_needs_x) umma_commit(xempty_bar(xs));
        if (i == steps - 1) umma_commit(done_bar);
      }
      __syncwarp();
      if (++ws == kWStages) { ws = 0; wphase ^= 1u; }
      if (next_needs_x && ++xs == kXStages) xs = 0;
    }
  }
  if (warp == 1) {
    mbar_wait(done_bar, 0);
    if (lane == 0) {
      cycles[blockIdx.x] = clock64() - t0;
      *reinterpret_cast<volatile uint32_t*>(&done_flag) = 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  const int rows_per_cta = 4096;            // 512 KB per CTA, 76 MB in all: L2 resident
  const size_t rows = (size_t)148 * rows_per_cta;
  std::vector<__nv_bfloat16> h(rows * 64);
  srand(1);
  for (auto& v : h) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);  // real data: real switching power
  __nv_bfloat16* d;
  cudaMalloc(&d, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  auto mk = [&](int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
  };
  CUtensorMap wm = mk(128), xm = mk(256);
  long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * 148);
  const int smem = kMaxWStages * kWBytes + kXStages * kXBytes + 1024;
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int steps = 50000;
  struct Case { int n, reuse, stages, no_tma, kps, flags; const char* what; };
  const Case cases[] = {
      {256, 0, 10, 1, 1, 16, "warp-converged loops (revisions B/C), barriers only"},
      {256, 0, 10, 1, 1, 8 | 32, "single-thread MMA + producer loops, barriers only"},
      {256, 0, 10, 0, 1, 8 | 32, "  ... + weight TMA"},
      {256, 27, 10, 0, 1, 8 | 32, "  ... + activation tile every 27 k-steps"},
      {256, 9, 10, 0, 1, 8 | 32, "  ... + activation tile every 9 k-steps"},
      {256, 3, 10, 0, 1, 8 | 32, "  ... + activation tile every 3 k-steps"},
      {256, 3, 6, 0, 1, 8 | 32, "  ... + activation tile every 3 k-steps, ring of 6"},
      {256, 3, 6, 0, 1, 8 | 32 | 64, "  ... ring of 6, four busy co-resident warps"},
      {256, 3, 10, 0, 1, 8 | 32 | 64, "  ... ring of 10, four busy co-resident warps"},
      {256, 1, 10, 0, 1, 8 | 32, "  ... + activation tile every k-step"},
      {256, 3, 4, 0, 1, 16, "revisions B/C: activation tile every 3 k-steps, ring of 4"},
      {256, 1, 4, 0, 1, 16, "revisions B/C: activation tile every k-step, ring of 4"},
  };
  for (const Case& cs : cases) {
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      feed_kernel<<<148, 192, smem>>>(wm, xm, steps, cs.reuse, cs.n, rows_per_cta, d_cycles, cs.stages, cs.no_tma,
                                     cs.kps, cs.flags);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      cudaEventElapsedTime(&ms, e0, e1);
    }
    long long c[148];
    cudaMemcpy(c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)c[i] / 148;
    const double ksteps = (double)steps * cs.kps;
    const double flops = 148.0 * ksteps * 2.0 * 128 * 256 * 64;
    printf("%-58s %7.1f cycles per 128x256x64 k-step (ideal 512, %5.1f%% of pipe peak)  chip %7.1f TFLOP/s, %5.0f MHz\n",
           cs.what, avg / ksteps, 100.0 * 512.0 * ksteps / avg, flops / (ms * 1e-3) / 1e12, avg / (ms * 1e-3) / 1e6);
  }
  return 0;
}
