// First convolution of the UNets (in_conv / init_conv: 1..4 input channels -> 64 / 128 channels,
// 3^d taps; ddpm_3d_ldm/unet_attention.py:114,161, slice_cond_2d_ddpm/unet.py:137,185), inference:
// the patch matrix is built IN SHARED MEMORY from the fp32 NC[D]HW input and fed to tcgen05
// directly -- no im2col buffer, no layout pass, one launch.  K = taps * 4 <= 128 is far too thin
// for the generic implicit-GEMM kernel (its tiles are epilogue-bound there); this kernel is bound
// by writing the output (HBM) and by building the patches (issue slots).
//
// One CTA = 128 threads = 4 warps (warp w owns TMEM lanes 32w..32w+31 = positions of the tile);
// up to three CTAs per SM hide each other's phases.  Per tile of 128 consecutive positions:
//   build   thread = position: 27 (9) taps x cin loads from x (L1 / L2 resident), zero outside the
//           volume, packed to bf16 and stored as the K-major, 128B-swizzled A operand (K index =
//           tap * 4 + channel, the column order of the packed weights)
//   mma     one thread: 8 x tcgen05.mma (M = 128 positions, N = Cout, K = 16), weights resident in
//           shared memory for the whole kernel, accumulator in TMEM
//   store   thread = position: + bias, GroupNorm partial sums (registers across tiles, reduced
//           when the sample changes), bf16, 16-byte global stores of the position's channel row
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/mri_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace mri {

constexpr int kTcThreads = 128;
constexpr int kTcSlab = 128 * 128;   // 128 rows x 64 bf16

// n / d and n % d for 0 <= n < 2^24 through one float multiply and a fix-up
__device__ __forceinline__ void fdiv_small(int n, int d, float rcp, int& q, int& r) {
  q = __float2int_rz(__int2float_rn(n) * rcp);
  r = n - q * d;
  if (r < 0) {
    --q;
    r += d;
  } else if (r >= d) {
    ++q;
    r -= d;
  }
}

struct ThinConvArgs {
  const float* x;            // fp32 [samples][cin][D][H][W]
  const uint4* w;            // bf16 [Cout][128], column = tap * 4 + channel (zero padded)
  const float* bias;         // fp32 [Cout] or null
  __nv_bfloat16* y;          // bf16 [samples][D*H*W][Cout]
  double* stats;             // fp64 [samples][stats_ld][2] or null
  int stats_ld, samples, cin, D, H, W, ndim;
  long long spatial;
  int tiles_per_sample;
  long long n_tiles;
};

template <int COUT, int CIN, int NDIM>
__global__ void __launch_bounds__(kTcThreads, 3)
thin_in_conv_kernel(const ThinConvArgs a) {
  constexpr int CPG = COUT / 8;                 // GroupNorm(8, COUT): channels per group
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float s_bias[COUT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_smem = smem_base;                       // 2 slabs of COUT rows
  const uint32_t a_smem = smem_base + 2u * COUT * 128u;    // 2 slabs of 128 rows
  const uint32_t mbar = smem_u32(&bar);

  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_holder), 128);
    tmem_relinquish();
  }
  // weights -> shared memory, K-major 128B swizzle: 16-byte unit u of row n goes to slab u / 8,
  // unit (u % 8) ^ (n % 8)
  for (int i = tid; i < COUT * 16; i += kTcThreads) {
    const int n = i >> 4, u = i & 15;
    const uint4 v = __ldg(a.w + i);
    const uint32_t dst = w_smem + (uint32_t)(u >> 3) * (COUT * 128u) + (uint32_t)n * 128u +
                         (uint32_t)(((u & 7) ^ (n & 7)) << 4);
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
  }
  for (int i = tid; i < COUT; i += kTcThreads) s_bias[i] = a.bias != nullptr ? __ldg(a.bias + i) : 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_holder;
  const uint32_t idesc = umma_idesc_bf16(128, COUT);

  // contiguous range of tiles per CTA (tiles of one sample stay together: few statistics flushes)
  const long long t0 = a.n_tiles * blockIdx.x / gridDim.x;
  const long long t1 = a.n_tiles * (blockIdx.x + 1) / gridDim.x;
  constexpr int taps = NDIM == 3 ? 27 : 9;
  const long long plane = (long long)a.H * a.W;
  const float rcp_w = 1.0f / (float)a.W;
  float2 gs[8], gq[8];   // per group: sums / sums of squares of the even and the odd channels
#pragma unroll
  for (int g = 0; g < 8; ++g) gs[g] = gq[g] = make_float2(0.f, 0.f);
  const int sp32 = (int)a.spatial, plane32 = (int)plane;
  int acc_sample = -1;
  uint32_t phase = 0;
  const uint32_t row_smem = a_smem + (uint32_t)tid * 128u;
  const uint32_t xr = (uint32_t)(tid & 7);

  auto flush_stats = [&]() {  // all threads of the warp
    if (a.stats != nullptr && acc_sample >= 0) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float s = gs[g].x + gs[g].y, q = gq[g].x + gq[g].y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o);
          q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane == 0) {
          double* dst = a.stats + ((size_t)acc_sample * a.stats_ld + g) * 2;
          atomicAdd(dst, (double)s);
          atomicAdd(dst + 1, (double)q);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) gs[g] = gq[g] = make_float2(0.f, 0.f);
  };

  for (long long t = t0; t < t1; ++t) {
    const int n = (int)(t / a.tiles_per_sample);
    const long long s = (t - (long long)n * a.tiles_per_sample) * 128 + tid;   // my position
    const bool valid = s < a.spatial;
    if (n != acc_sample) {
      flush_stats();
      acc_sample = n;
    }
    // ---- build my row of the patch matrix --------------------------------------------------
    int d0 = 0, h0 = 0, w0 = 0;
    {
      long long rest = s;
      if (NDIM == 3) {
        d0 = (int)(rest / plane);
        rest -= (long long)d0 * plane;
      }
      int r = (int)rest, qh, rw;
      fdiv_small(r, a.W, rcp_w, qh, rw);
      h0 = qh;
      w0 = rw;
    }
    // x[n][c][s + off]: 32-bit offsets from my own voxel; per-axis validity computed once
    const float* xs = a.x + (size_t)n * CIN * a.spatial + (valid ? s : 0);
    bool vd[3], vh[3], vw[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      vd[k] = valid && (NDIM != 3 ? k == 1 : (unsigned)(d0 + k - 1) < (unsigned)a.D);
      vh[k] = (unsigned)(h0 + k - 1) < (unsigned)a.H;
      vw[k] = (unsigned)(w0 + k - 1) < (unsigned)a.W;
    }
    int tap = 0;   // compile-time after unrolling
#pragma unroll
    for (int kd = (NDIM == 3 ? 0 : 1); kd < (NDIM == 3 ? 3 : 2); ++kd) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const bool okdh = vd[kd] && vh[kh];
        const int base = (kd - 1) * plane32 + (kh - 1) * a.W - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw, ++tap) {
          const bool ok = okdh && vw[kw];
          float v[4] = {0.f, 0.f, 0.f, 0.f};
          if (ok) {
            const float* px = xs + (base + kw);
#pragma unroll
            for (int c = 0; c < CIN; ++c) v[c] = __ldg(px + c * sp32);
          }
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]);
          const __nv_bfloat162 p1 = __floats2bfloat162_rn(v[2], v[3]);
          // K index tap * 4: slab (tap >> 4), 16-byte unit ((tap >> 1) & 7) ^ (row & 7), half tap & 1
          const uint32_t dst = row_smem + (uint32_t)(tap >> 4) * kTcSlab +
                               (((uint32_t)((tap >> 1) & 7) ^ xr) << 4) + (uint32_t)(tap & 1) * 8u;
          asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(dst),
                       "r"(*reinterpret_cast<const uint32_t*>(&p0)),
                       "r"(*reinterpret_cast<const uint32_t*>(&p1))
                       : "memory");
        }
      }
    }
    // zero the K padding [taps * 4, 128)
#pragma unroll
    for (int k8 = taps; k8 < 32; ++k8) {
      const uint32_t dst = row_smem + (uint32_t)(k8 >> 4) * kTcSlab +
                           (((uint32_t)((k8 >> 1) & 7) ^ xr) << 4) + (uint32_t)(k8 & 1) * 8u;
      asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(dst), "r"(0u), "r"(0u) : "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();   // rows complete; everybody has finished reading the previous accumulator
    // ---- MMA -----------------------------------------------------------------------------------
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const uint64_t a_desc = umma_desc_k_sw128(a_smem + sl * kTcSlab, 1024);
        const uint64_t b_desc = umma_desc_k_sw128(w_smem + sl * (COUT * 128), 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_acc, a_desc + 2u * k, b_desc + 2u * k, idesc, (sl != 0 || k != 0) ? 1u : 0u);
      }
      umma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    tc_fence_after();
    // ---- epilogue: my position's COUT channels ---------------------------------------------------
    const uint32_t tacc = tmem_acc + ((uint32_t)(warp * 32) << 16);
    __nv_bfloat16* yrow = a.y + ((size_t)n * a.spatial + (size_t)(valid ? s : 0)) * COUT;
    uint32_t v[32];
    tmem_ld32(tacc, v);
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 32) {
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
      if (c0 + 32 < COUT) tmem_ld32(tacc + (uint32_t)(c0 + 32), v);
      float2 f2[16];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {   // bias from shared memory, channel pairs on the packed fp32 pipe
        const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c0 + 2 * i]);
        f2[i] = __fadd2_rn(make_float2(f[2 * i], f[2 * i + 1]), make_float2(b4.x, b4.y));
        f2[i + 1] = __fadd2_rn(make_float2(f[2 * i + 2], f[2 * i + 3]), make_float2(b4.z, b4.w));
      }
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int g = (c0 + 2 * i) / CPG;     // compile-time after unrolling; CPG is even
          gs[g] = __fadd2_rn(gs[g], f2[i]);
          gq[g] = __ffma2_rn(f2[i], f2[i], gq[g]);
          f[2 * i] = f2[i].x;
          f[2 * i + 1] = f2[i].y;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t wv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[u * 8 + 2 * e], f[u * 8 + 2 * e + 1]);
            wv[e] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          *reinterpret_cast<uint4*>(yrow + c0 + u * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
      }
    }
    tc_fence_before();   // the next tile's MMA overwrites the accumulator after the __syncthreads above
  }
  flush_stats();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_acc, 128);
}

}  // namespace mri

using namespace mri;

extern "C" int mri_thin_in_conv(const float* x, const void* w_packed, const float* bias, void* y,
                                double* stats, int stats_ld, int samples, int cin, int D, int H, int W,
                                int ndim, int cout, void* stream) {
  if (x == nullptr || w_packed == nullptr || y == nullptr)
    return set_error(-1, "mri_thin_in_conv: null argument");
  if (ndim != 2 && ndim != 3) return set_error(-2, "mri_thin_in_conv: ndim must be 2 or 3");
  if (cin < 1 || cin > 4) return set_error(-2, "mri_thin_in_conv: 1..4 input channels");
  if (cout != 64 && cout != 128) return set_error(-2, "mri_thin_in_conv: 64 or 128 output channels");
  if (samples < 1 || D < 1 || H < 1 || W < 1) return set_error(-2, "mri_thin_in_conv: empty input");
  if (ndim == 2 && D != 1) return set_error(-2, "mri_thin_in_conv: 2-D problems have D = 1");
  if (stats != nullptr && stats_ld < 8) return set_error(-2, "mri_thin_in_conv: statistics need 8 groups");
  if ((long long)H * W >= (1LL << 24)) return set_error(-2, "mri_thin_in_conv: plane too large");
  ThinConvArgs a;
  a.x = x;
  a.w = reinterpret_cast<const uint4*>(w_packed);
  a.bias = bias;
  a.y = reinterpret_cast<__nv_bfloat16*>(y);
  a.stats = stats;
  a.stats_ld = stats_ld;
  a.samples = samples;
  a.cin = cin;
  a.D = D;
  a.H = H;
  a.W = W;
  a.ndim = ndim;
  a.spatial = (long long)D * H * W;
  a.tiles_per_sample = (int)((a.spatial + 127) / 128);
  a.n_tiles = (long long)samples * a.tiles_per_sample;
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaDeviceGetAttribute(SM count)");
  }
  long long grid = 3LL * n_sms;
  if (grid > a.n_tiles) grid = a.n_tiles;
  const int smem = 2 * cout * 128 + 2 * kTcSlab + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  if (a.spatial * cin >= (1LL << 31)) return set_error(-2, "mri_thin_in_conv: sample too large");
  typedef void (*Kern)(const ThinConvArgs);
#define MRI_TC_ROW(CO, ND) \
  {thin_in_conv_kernel<CO, 1, ND>, thin_in_conv_kernel<CO, 2, ND>, thin_in_conv_kernel<CO, 3, ND>, \
   thin_in_conv_kernel<CO, 4, ND>}
  static const Kern kerns[2][2][4] = {{MRI_TC_ROW(64, 2), MRI_TC_ROW(128, 2)},
                                      {MRI_TC_ROW(64, 3), MRI_TC_ROW(128, 3)}};
#undef MRI_TC_ROW
  static int configured[2][2][4] = {{{0, 0, 0, 0}, {0, 0, 0, 0}}, {{0, 0, 0, 0}, {0, 0, 0, 0}}};
  const int ci = cout == 128 ? 1 : 0;
  Kern kern = kerns[ndim - 2][ci][cin - 1];
  if (!configured[ndim - 2][ci][cin - 1]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(thin_in_conv_kernel)");
    configured[ndim - 2][ci][cin - 1] = 1;
  }
  kern<<<(unsigned)grid, kTcThreads, smem, st>>>(a);
  return check_launch("thin_in_conv_kernel");
}
