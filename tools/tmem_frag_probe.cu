// Probe (B200 only): register layout of tcgen05.ld.16x256b.x2 and of stmatrix.x4.trans -- the
// building blocks of an epilogue that moves a D^T accumulator (TMEM lane = channel, column =
// position) to a channels-last bf16 staging tile without per-element shared-memory stores.
// TMEM is filled with value(lane, col) = lane * 256 + col through the 32x32b shape; every thread
// then prints what the 16x256b.x2 load hands it, and the stmatrix result is dumped.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tmem_frag_probe tools/tmem_frag_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../mri_image_generation_b200/csrc/ptx.cuh"

using namespace mri;

__global__ void __launch_bounds__(128, 1) probe(float* out_regs, uint16_t* out_smem) {
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(128) uint16_t tile[16 * 64];  // 16 rows (positions) x 64 channels (128 B rows)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_holder), 32);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 16 * 64; i += 128) tile[i] = 0xffff;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  // fill: lane L (= warp * 32 + lane), columns 0..15
  uint32_t v[16];
  for (int c = 0; c < 16; ++c) v[c] = __float_as_uint((float)((warp * 32 + lane) * 256 + c));
  tmem_st16(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem));
    tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out_regs[lane * 8 + i] = __uint_as_float(r[i]);
    // pack (row, 2 columns) pairs to bf16x2 of SMALL ids so that they are exact: id = channel * 16 + position
    uint32_t m[4];
    for (int i = 0; i < 4; ++i) {
      const float a = __uint_as_float(r[2 * i]), b = __uint_as_float(r[2 * i + 1]);
      const int la = (int)a / 256, ca = (int)a % 256, lb = (int)b / 256, cb = (int)b % 256;
      __nv_bfloat162 h2 = __floats2bfloat162_rn((float)(la * 16 + ca), (float)(lb * 16 + cb));
      m[i] = *reinterpret_cast<uint32_t*>(&h2);
    }
    // stmatrix.x4.trans: lanes 8i .. 8i+7 give the row addresses of matrix i.  Matrices in register
    // order i = 0..3; we place matrix i at rows (i >> 1) * 8 + (lane & 7), 16-byte unit (i & 1).
    const int i = lane >> 3, rr = lane & 7;
    const uint32_t addr = smem_u32(tile) + (uint32_t)(((i >> 1) * 8 + rr) * 128 + (i & 1) * 16);
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(m[0]),
                 "r"(m[1]), "r"(m[2]), "r"(m[3]) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * 64; i += 128) out_smem[i] = tile[i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  float* d_regs;
  uint16_t* d_smem;
  cudaMalloc(&d_regs, 32 * 8 * 4);
  cudaMalloc(&d_smem, 16 * 64 * 2);
  probe<<<1, 128>>>(d_regs, d_smem);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  float regs[32 * 8];
  uint16_t sm[16 * 64];
  cudaMemcpy(regs, d_regs, sizeof(regs), cudaMemcpyDeviceToHost);
  cudaMemcpy(sm, d_smem, sizeof(sm), cudaMemcpyDeviceToHost);
  printf("tcgen05.ld.16x256b.x2: thread -> (lane, column) of registers r0..r7\n");
  for (int t = 0; t < 32; ++t) {
    printf("T%2d:", t);
    for (int i = 0; i < 8; ++i) printf(" (%2d,%2d)", (int)regs[t * 8 + i] / 256, (int)regs[t * 8 + i] % 256);
    printf("\n");
  }
  printf("stmatrix.x4.trans result, rows = 16 positions, first 16 channels (value = channel*16 + position):\n");
  for (int r = 0; r < 16; ++r) {
    printf("row %2d:", r);
    for (int c = 0; c < 16; ++c) {
      uint32_t bits = (uint32_t)sm[r * 64 + c] << 16;
      float f;
      memcpy(&f, &bits, 4);
      printf(" %3d", sm[r * 64 + c] == 0xffff ? -1 : (int)f);
    }
    printf("\n");
  }
  return 0;
}
