// Probe (B200 only): can ONE MN-major tcgen05 B operand cover the three kw taps of a wgrad group
// as three shifted views of a single 10-wide activation tile?  Operand B of the weight-gradient
// MMA is [K = 16 positions] x [N = 3 taps x 64 channels], MN-major, 128B swizzle.  The three
// 64-channel blocks of N are the SAME rows shifted by one position each, so the descriptor's
// leading-dimension byte offset (distance between 64-element MN blocks) would be 128 B -- one row --
// and its stride byte offset (distance between 8-row K atoms) 1280 B, one 10-position line.
//   D[co][j*64 + c] = sum_{y < 16, x < 8} dY[8y + x][co] * X[10y + x + j][c],   j = 0, 1, 2
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/umma_mn_probe tools/umma_mn_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../mri_image_generation_b200/csrc/ptx.cuh"

using namespace mri;

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap dy_map, const __grid_constant__ CUtensorMap x_map,
             float* out, int n_cols, int lbo_bytes, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_holder;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t dy_s = base;            // 2 x [128 rows x 64 co]
  const uint32_t x_s = base + 32768;     // 160 rows x 64 ch
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
    mbar_arrive_expect_tx(smem_u32(&bar[0]), 32768 + 160 * 128);
    tma_load_2d(dy_s, &dy_map, smem_u32(&bar[0]), 0, 0);
    tma_load_2d(dy_s + 16384, &dy_map, smem_u32(&bar[0]), 64, 0);
    tma_load_2d(x_s, &x_map, smem_u32(&bar[0]), 0, 0);
    mbar_wait(smem_u32(&bar[0]), 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)n_cols) | (1u << 15) | (1u << 16);
    for (int ks = 0; ks < 8; ++ks) {  // 16 positions = 2 lines per MMA
      const uint64_t a_desc = desc_mn_sw128(dy_s + ks * 2048, 16384, 1024);
      const uint32_t xa = x_s + ks * 2560;
      uint64_t b_desc = desc_mn_sw128(xa, (uint32_t)lbo_bytes, 1280);
      if (use_base_offset) b_desc |= (uint64_t)((xa >> 7) & 7u) << 49;
      umma_bf16(tmem, a_desc, b_desc, idesc, ks != 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar[1]));
  }
  mbar_wait(smem_u32(&bar[1]), 0);
  tc_fence_after();
  const uint32_t tacc = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < n_cols; c0 += 16) {
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
  std::vector<float> DY(128 * 128), X(160 * 64);
  srand(3);
  for (auto& v : DY) v = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : X) v = bf((rand() % 2001 - 1000) / 1000.f);
  std::vector<__nv_bfloat16> DYb(DY.size()), Xb(X.size());
  for (size_t i = 0; i < DY.size(); ++i) DYb[i] = __float2bfloat16(DY[i]);
  for (size_t i = 0; i < X.size(); ++i) Xb[i] = __float2bfloat16(X[i]);
  __nv_bfloat16 *dDY, *dX;
  float* dO;
  cudaMalloc(&dDY, DYb.size() * 2);
  cudaMalloc(&dX, Xb.size() * 2);
  cudaMalloc(&dO, 128 * 256 * 4);
  cudaMemcpy(dDY, DYb.data(), DYb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, Xb.data(), Xb.size() * 2, cudaMemcpyHostToDevice);
  auto mk = [&](void* p, int cols, int rows, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
  };
  CUtensorMap dym = mk(dDY, 128, 128, 128), xm = mk(dX, 64, 160, 160);
  const int smem = 32768 + 160 * 128 + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * 256);
  const int cases[][3] = {{64, 128, 0}, {128, 128, 0}, {192, 128, 0}, {192, 128, 1}};
  for (auto& c : cases) {
    cudaMemset(dO, 0, O.size() * 4);
    probe_kernel<<<1, 128, smem>>>(dym, xm, dO, c[0], c[1], c[2]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d: CUDA error %s\n", c[0], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int co = 0; co < 128; ++co)
      for (int n = 0; n < c[0]; ++n) {
        const int j = n / 64, ch = n % 64;
        double ref = 0;
        for (int y = 0; y < 16; ++y)
          for (int x = 0; x < 8; ++x) ref += (double)DY[(8 * y + x) * 128 + co] * X[(10 * y + x + j) * 64 + ch];
        maxerr = fmax(maxerr, fabs(ref - O[co * 256 + n]));
      }
    printf("N = %3d (%d taps in one operand), LBO %d B, SBO 1280 B, base_offset %s: max |err| = %.4g  %s\n", c[0], c[0] / 64,
           c[1], c[2] ? "set" : "unset", maxerr, maxerr < 2e-3 ? "MATCH" : "mismatch");
  }
  return 0;
}
