// GroupNorm statistics and the fused GroupNorm-apply + SiLU + time-embedding add + residual add
// pass.  HBM-bound: every element is read once and written once as 16-byte bf16x8 vectors.
// Reference call sites: see include/mri_b200.h (mri_gn_stats / mri_gn_apply).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/mri_b200.h"
#include "common.h"

namespace mri {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&h2);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------------------------------
// Statistics: grid (chunks, samples).  Each block reduces rows [chunk*rows_per, ...) of one
// sample; thread -> fixed channel vector (8 channels), strided over rows, so the per-thread
// partial belongs to exactly one fine group.  Block partials are combined in shared memory
// and added to stats with one atomic pair per fine group.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_stats_kernel(const uint4* __restrict__ x, float* __restrict__ stats, int64_t spatial, int C,
                int stats_ld, int stats_g0, int stats_cpg, int rows_per_block) {
  extern __shared__ float red[];  // [n_fine][2]
  const int vec_per_row = C >> 3;
  const int n_fine = C / stats_cpg;
  const int sample = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > spatial) r1 = spatial;
  for (int i = threadIdx.x; i < n_fine * 2; i += blockDim.x) red[i] = 0.f;
  __syncthreads();

  const int rows_step = blockDim.x / vec_per_row;  // host guarantees blockDim % vec_per_row == 0
  const int cv = threadIdx.x % vec_per_row;
  const int rsub = threadIdx.x / vec_per_row;
  float s = 0.f, ss = 0.f;
  const uint4* base = x + (size_t)sample * spatial * vec_per_row;
  for (int64_t r = r0 + rsub; r < r1; r += rows_step) {
    const uint4 v = __ldg(base + r * vec_per_row + cv);
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s += f[i];
      ss = fmaf(f[i], f[i], ss);
    }
  }
  const int g = (cv * 8) / stats_cpg;
  atomicAdd(&red[2 * g], s);
  atomicAdd(&red[2 * g + 1], ss);
  __syncthreads();
  for (int i = threadIdx.x; i < n_fine * 2; i += blockDim.x) {
    const int gi = i >> 1;
    atomicAdd(stats + ((size_t)sample * stats_ld + stats_g0 + gi) * 2 + (i & 1), red[i]);
  }
}

// ---------------------------------------------------------------------------------------
// Apply.  One thread per 8-channel vector; grid-stride.  mean / rstd are derived on the fly
// from the fine-group partial sums (a handful of cached loads per vector).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_apply_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, const float* __restrict__ stats,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ rowbias, int rowbias_ld, const uint4* __restrict__ residual,
                int64_t spatial, int C, int groups, int stats_ld, int stats_g0, int stats_cpg,
                float eps, int silu, int64_t total_vec) {
  const int vec_per_row = C >> 3;
  const int cpg = C / groups;           // channels per normalisation group
  const int comb = cpg / stats_cpg;     // fine groups per normalisation group
  const float inv_cnt = 1.0f / ((float)cpg * (float)spatial);
  const int64_t vec_per_sample = spatial * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int sample = (int)(i / vec_per_sample);
    const int cv = (int)(i % vec_per_row);
    const int c = cv * 8;
    const int g = c / cpg;
    const float* st = stats + ((size_t)sample * stats_ld + stats_g0 + g * comb) * 2;
    float s = 0.f, ss = 0.f;
    for (int j = 0; j < comb; ++j) {
      s += __ldg(st + 2 * j);
      ss += __ldg(st + 2 * j + 1);
    }
    const float mean = s * inv_cnt;
    float var = ss * inv_cnt - mean * mean;
    var = var < 0.f ? 0.f : var;
    const float rstd = rsqrtf(var + eps);

    float f[8];
    unpack8(__ldg(x + i), f);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = (f[k] - mean) * rstd * gm[k] + bt[k];
      if (silu) v = v / (1.0f + __expf(-v));
      f[k] = v;
    }
    if (rowbias != nullptr) {
      const float* rb = rowbias + (size_t)sample * rowbias_ld + c;
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(rb));
      const float4 r1 = __ldg(reinterpret_cast<const float4*>(rb + 4));
      f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
      f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
    }
    if (residual != nullptr) {
      float rr[8];
      unpack8(__ldg(residual + i), rr);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += rr[k];
    }
    y[i] = pack8(f);
  }
}

}  // namespace mri

using namespace mri;

extern "C" int mri_gn_stats(const void* x, float* stats, int samples, int64_t spatial, int C,
                            int stats_ld, int stats_g0, int stats_cpg, void* stream) {
  if (C % 8 != 0 || stats_cpg % 8 != 0 || C % stats_cpg != 0)
    return set_error(-2, "mri_gn_stats: C and stats_cpg must be multiples of 8");
  const int vec_per_row = C / 8;
  int threads = 256;
  if (vec_per_row > 256) return set_error(-2, "mri_gn_stats: C > 2048 unsupported");
  threads = (256 / vec_per_row) * vec_per_row;
  // aim for >= 4 blocks per SM across the grid
  int64_t want_blocks = 148 * 4 / (samples > 0 ? samples : 1) + 1;
  int64_t rows_per = (spatial + want_blocks - 1) / want_blocks;
  const int rows_step = threads / vec_per_row;
  if (rows_per < rows_step * 4) rows_per = rows_step * 4;
  const int chunks = (int)((spatial + rows_per - 1) / rows_per);
  const int n_fine = C / stats_cpg;
  gn_stats_kernel<<<dim3(chunks, samples), threads, n_fine * 2 * sizeof(float),
                    (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(x), stats, spatial, C,
                                            stats_ld, stats_g0, stats_cpg, (int)rows_per);
  return check_launch("gn_stats_kernel");
}

extern "C" int mri_gn_apply(const void* x, void* y, const float* stats, const float* gamma,
                            const float* beta, const float* rowbias, int rowbias_ld,
                            const void* residual, int samples, int64_t spatial, int C, int groups,
                            int stats_ld, int stats_g0, int stats_cpg, float eps, int silu,
                            void* stream) {
  if (C % 8 != 0 || groups < 1 || C % groups != 0 || (C / groups) % stats_cpg != 0)
    return set_error(-2, "mri_gn_apply: bad channel / group configuration");
  const int64_t total_vec = (int64_t)samples * spatial * (C / 8);
  int64_t blocks = (total_vec + 255) / 256;
  const int64_t cap = 148 * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  gn_apply_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), stats, gamma, beta, rowbias,
      rowbias_ld, reinterpret_cast<const uint4*>(residual), spatial, C, groups, stats_ld, stats_g0,
      stats_cpg, eps, silu, total_vec);
  return check_launch("gn_apply_kernel");
}
