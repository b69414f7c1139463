// Probe of tcgen05 shared-memory descriptor semantics (B200 only):
// can the N operand of one MMA be a STRIDED, UNALIGNED view of a 128B-swizzled tile that TMA
// wrote?  Rows of the operand = row (start + g * group_stride + r), g = 0..31, r = 0..7, taken
// from a linear buffer of 128-byte rows.  This is what a "halo tile" convolution needs: the 27
// taps of a 3x3x3 filter would then read shifted views of ONE activation tile kept in shared
// memory instead of re-loading it from L2 for every tap.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu -lcuda
//   ./umma_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../mri_image_generation_b200/csrc/ptx.cuh"

using namespace mri;

constexpr int kRows = 512;  // rows of the linear activation buffer in smem (64 KB)

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap w_map, const __grid_constant__ CUtensorMap x_map,
             float* out, int start_row, int group_stride_rows, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_holder;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_s = base;               // 128 rows x 128 B
  const uint32_t x_s = base + 16384;       // kRows rows x 128 B
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_holder), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(smem_u32(&bar[0]), 16384 + kRows * 128);
    tma_load_2d(w_s, &w_map, smem_u32(&bar[0]), 0, 0);
    for (int r0 = 0; r0 < kRows; r0 += 256) tma_load_2d(x_s + r0 * 128, &x_map, smem_u32(&bar[0]), 0, r0);
    mbar_wait(smem_u32(&bar[0]), 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 256);
    const uint32_t xa = x_s + (uint32_t)start_row * 128u;
    for (int k = 0; k < 4; ++k) {
      uint64_t a_desc = umma_desc_k_sw128(w_s, 1024) + 2u * k;
      uint64_t b_desc = umma_desc_k_sw128(xa, (uint32_t)group_stride_rows * 128u) + 2u * k;
      if (use_base_offset) b_desc |= (uint64_t)((xa >> 7) & 7u) << 49;
      umma_bf16(tmem, a_desc, b_desc, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar[1]));
  }
  mbar_wait(smem_u32(&bar[1]), 0);
  tc_fence_after();
  // lane = output row (weight row), columns = 256 positions
  const uint32_t tacc = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 256; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tacc + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 256 + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  std::vector<float> W(128 * 64), X(kRows * 64);
  srand(1);
  for (auto& v : W) v = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : X) v = bf((rand() % 2001 - 1000) / 1000.f);
  std::vector<__nv_bfloat16> Wb(W.size()), Xb(X.size());
  for (size_t i = 0; i < W.size(); ++i) Wb[i] = __float2bfloat16(W[i]);
  for (size_t i = 0; i < X.size(); ++i) Xb[i] = __float2bfloat16(X[i]);
  __nv_bfloat16 *dW, *dX;
  float* dO;
  cudaMalloc(&dW, Wb.size() * 2);
  cudaMalloc(&dX, Xb.size() * 2);
  cudaMalloc(&dO, 128 * 256 * 4);
  cudaMemcpy(dW, Wb.data(), Wb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, Xb.data(), Xb.size() * 2, cudaMemcpyHostToDevice);
  auto mk = [&](void* p, int rows, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
  };
  CUtensorMap wm = mk(dW, 128, 128), xm = mk(dX, kRows, 256);
  const int smem = 16384 + kRows * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * 256);
  const int cases[][2] = {{0, 8}, {0, 10}, {1, 8}, {1, 10}, {3, 10}, {11, 10}, {5, 16}, {21, 12}};
  for (auto& c : cases) {
    for (int ubo = 0; ubo < 2; ++ubo) {
      cudaMemset(dO, 0, O.size() * 4);
      probe_kernel<<<1, 128, smem>>>(wm, xm, dO, c[0], c[1], ubo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("start %d stride %d base_off %d: CUDA error %s\n", c[0], c[1], ubo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 256; ++n) {
          const int row = c[0] + (n / 8) * c[1] + (n % 8);
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)W[m * 64 + k] * X[row * 64 + k];
          maxerr = fmax(maxerr, fabs(ref - O[m * 256 + n]));
        }
      printf("start_row %2d group_stride %2d rows, base_offset %s: max |err| = %.4g  %s\n", c[0], c[1],
             ubo ? "set  " : "unset", maxerr, maxerr < 1e-3 ? "MATCH" : "mismatch");
    }
  }
  return 0;
}
