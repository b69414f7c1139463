// GroupNorm statistics and the fused GroupNorm-apply + SiLU + time-embedding add + residual add
// pass.  HBM-bound: every element is read once and written once as 16-byte bf16x8 vectors.
// Reference call sites: see include/mri_b200.h (mri_gn_stats / mri_gn_apply).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>

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
gn_stats_kernel(const uint4* __restrict__ x, double* __restrict__ stats, int64_t spatial, int C,
                int stats_ld, int stats_g0, int stats_cpg, int rows_per_block) {
  extern __shared__ double red[];  // [n_fine][2] (fp64: summation order does not matter)
  const int vec_per_row = C >> 3;
  const int n_fine = C / stats_cpg;
  const int sample = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > spatial) r1 = spatial;
  for (int i = threadIdx.x; i < n_fine * 2; i += blockDim.x) red[i] = 0.0;
  __syncthreads();

  const int rows_step = blockDim.x / vec_per_row;  // host guarantees blockDim % vec_per_row == 0
  const int cv = threadIdx.x % vec_per_row;
  const int rsub = threadIdx.x / vec_per_row;
  // two accumulators per thread: channels [0,4) and [4,8) of its vector may belong to different
  // fine groups when stats_cpg == 4 (GroupNorm(8, 32) of the VAE, ddpm_3d_ldm/vae.py:8)
  float s[2] = {0.f, 0.f}, ss[2] = {0.f, 0.f};
  const uint4* base = x + (size_t)sample * spatial * vec_per_row;
  for (int64_t r = r0 + rsub; r < r1; r += rows_step) {
    const uint4 v = __ldg(base + r * vec_per_row + cv);
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i >> 2] += f[i];
      ss[i >> 2] = fmaf(f[i], f[i], ss[i >> 2]);
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int g = (cv * 8 + 4 * h) / stats_cpg;
    atomicAdd(&red[2 * g], (double)s[h]);
    atomicAdd(&red[2 * g + 1], (double)ss[h]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_fine * 2; i += blockDim.x) {
    const int gi = i >> 1;
    atomicAdd(stats + ((size_t)sample * stats_ld + stats_g0 + gi) * 2 + (i & 1), red[i]);
  }
}

// ---------------------------------------------------------------------------------------
// Apply.  ONE resident wave of blocks (148 x occupancy, mri_gn_apply) over the flattened
// (sample, row) space: block b owns rows [b * rows_per_block, ...) of that space, i.e. a range of
// one sample or the tail of one and the head of the next -- every block has the same amount of
// work for ANY batch size (a (chunks, samples) grid leaves a ragged second wave whenever the slot
// count is not a multiple of the batch: 640 blocks on 592 slots at batch 64).  A thread owns one
// 8-channel vector position (fixed channel group), so mean / rstd / gamma / beta collapse into 8
// (scale, shift) pairs computed once per sample segment; the thread then streams rows with 4
// independent 16-byte loads in flight.
// ---------------------------------------------------------------------------------------
template <bool kSilu, bool kResidual>
__global__ void __launch_bounds__(256, kResidual ? 3 : 4)
gn_apply_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, const double* __restrict__ stats,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ rowbias, int rowbias_ld, const uint4* __restrict__ residual,
                int samples, int64_t spatial, int C, int groups, int stats_ld, int stats_g0,
                int stats_cpg, float eps, int64_t rows_per_block) {
  const int vec_per_row = C >> 3;
  const int cpg = C / groups;        // channels per normalisation group
  const int comb = cpg / stats_cpg;  // fine groups per normalisation group
  const int cv = threadIdx.x % vec_per_row;
  const int rsub = threadIdx.x / vec_per_row;
  const int rows_step = blockDim.x / vec_per_row;
  const int c = cv * 8;
  const double inv_cnt = 1.0 / ((double)cpg * (double)spatial);
  const int64_t total_rows = (int64_t)samples * spatial;
  int64_t g0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t g1 = g0 + rows_per_block;
  if (g1 > total_rows) g1 = total_rows;
  while (g0 < g1) {
  const int sample = (int)(g0 / spatial);
  const int r0 = (int)(g0 - (int64_t)sample * spatial);
  const int r1 = (g1 - g0) < (spatial - r0) ? r0 + (int)(g1 - g0) : (int)spatial;
  g0 += r1 - r0;
  float sc[8], sh[8], rb[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // the two halves of the vector may lie in different groups
    const int g = (c + 4 * h) / cpg;
    const double* st = stats + ((size_t)sample * stats_ld + stats_g0 + g * comb) * 2;
    double s = 0.0, ss = 0.0;
    for (int j = 0; j < comb; ++j) {
      s += __ldg(st + 2 * j);
      ss += __ldg(st + 2 * j + 1);
    }
    const double mean_d = s * inv_cnt;
    double var_d = ss * inv_cnt - mean_d * mean_d;  // fp64: no cancellation problem
    var_d = var_d < 0.0 ? 0.0 : var_d;
    const float mean = (float)mean_d;
    const float rstd = rsqrtf((float)var_d + eps);
#pragma unroll
    for (int k = 4 * h; k < 4 * h + 4; ++k) {
      const float gm = __ldg(gamma + c + k), bt = __ldg(beta + c + k);
      sc[k] = rstd * gm;
      sh[k] = bt - mean * rstd * gm;
      rb[k] = rowbias != nullptr ? __ldg(rowbias + (size_t)sample * rowbias_ld + c + k) : 0.f;
    }
  }

  // channel PAIRS on the packed fp32x2 pipe (FFMA2 / FMUL2 / FADD2): the pass is co-limited by
  // instruction issue, not only by HBM; SiLU's sigmoid is 0.5 * tanh.approx(t / 2) + 0.5 (one
  // MUFU per element, relative error 2^-11: below the bf16 resolution of the stored result)
  float2 sc2[4], sh2[4], rb2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc2[j] = make_float2(sc[2 * j], sc[2 * j + 1]);
    sh2[j] = make_float2(sh[2 * j], sh[2 * j + 1]);
    rb2[j] = make_float2(rb[2 * j], rb[2 * j + 1]);
  }
  auto apply = [&](const uint4& xq, const uint4& rq) -> uint4 {
    const uint32_t xw[4] = {xq.x, xq.y, xq.z, xq.w};
    const uint32_t rw[4] = {rq.x, rq.y, rq.z, rq.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = make_float2(__uint_as_float(xw[j] << 16), __uint_as_float(xw[j] & 0xffff0000u));
      float2 t = __ffma2_rn(f, sc2[j], sh2[j]);
      if (kSilu) {
        const float2 h = __fmul2_rn(t, make_float2(0.5f, 0.5f));
        float2 th;
        asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(h.x));
        asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(h.y));
        t = __fmul2_rn(t, __ffma2_rn(th, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f)));
      }
      t = __fadd2_rn(t, rb2[j]);
      if (kResidual)
        t = __fadd2_rn(t, make_float2(__uint_as_float(rw[j] << 16), __uint_as_float(rw[j] & 0xffff0000u)));
      __nv_bfloat162 h2 = __floats2bfloat162_rn(t.x, t.y);
      ow[j] = *reinterpret_cast<uint32_t*>(&h2);
    }
    return make_uint4(ow[0], ow[1], ow[2], ow[3]);
  };
  // rows [r0, r1) of the sample; whole batches of U rows without predicates, then the ragged end
  const size_t base = (size_t)sample * spatial * vec_per_row + cv;
  const uint4* xp = x + base;
  const uint4* rp = kResidual ? residual + base : xp;
  uint4* yp = y + base;
  constexpr int U = 4;
  const int stepv = rows_step * vec_per_row;
  int r = r0 + rsub;
  for (; r + (U - 1) * rows_step < r1; r += U * rows_step) {
    uint4 v[U], rr[U];
    const int o = r * vec_per_row;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = __ldg(xp + o + u * stepv);
      if (kResidual) rr[u] = __ldg(rp + o + u * stepv);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) yp[o + u * stepv] = apply(v[u], kResidual ? rr[u] : v[u]);
  }
  for (; r < r1; r += rows_step) {
    const uint4 v = __ldg(xp + r * vec_per_row);
    const uint4 q = kResidual ? __ldg(rp + r * vec_per_row) : v;
    yp[r * vec_per_row] = apply(v, q);
  }
  }  // sample segments of this block's row range
}

template <typename K>
static int blocks_per_sm(K kernel, int threads) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, 0) != cudaSuccess || n < 1) n = 1;
  return n;
}

}  // namespace mri

using namespace mri;

extern "C" int mri_gn_stats(const void* x, double* stats, int samples, int64_t spatial, int C,
                            int stats_ld, int stats_g0, int stats_cpg, void* stream) {
  if (C % 8 != 0 || stats_cpg % 4 != 0 || C % stats_cpg != 0)
    return set_error(-2, "mri_gn_stats: C must be a multiple of 8 and stats_cpg of 4");
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
  gn_stats_kernel<<<dim3(chunks, samples), threads, n_fine * 2 * sizeof(double),
                    (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(x), stats, spatial, C,
                                            stats_ld, stats_g0, stats_cpg, (int)rows_per);
  return check_launch("gn_stats_kernel");
}

extern "C" int mri_gn_apply(const void* x, void* y, const double* stats, const float* gamma,
                            const float* beta, const float* rowbias, int rowbias_ld,
                            const void* residual, int samples, int64_t spatial, int C, int groups,
                            int stats_ld, int stats_g0, int stats_cpg, float eps, int silu,
                            void* stream) {
  if (C % 8 != 0 || groups < 1 || C % groups != 0 || (C / groups) % stats_cpg != 0 ||
      (C / groups) % 4 != 0)
    return set_error(-2, "mri_gn_apply: bad channel / group configuration");
  const int vec_per_row = C / 8;
  if (vec_per_row > 256) return set_error(-2, "mri_gn_apply: C > 2048 unsupported");
  const int threads = (256 / vec_per_row) * vec_per_row;
  const int rows_step = threads / vec_per_row;
  if (spatial * vec_per_row > 0x7fffffffLL) return set_error(-2, "mri_gn_apply: sample too large");
  static const int env_bps = [] {
    const char* e = getenv("MRI_GN_BLOCKS_PER_SM");  // tuning probe (tools/gn_probe.py)
    return e != nullptr ? atoi(e) : 0;
  }();
  const uint4* xp = reinterpret_cast<const uint4*>(x);
  uint4* yp = reinterpret_cast<uint4*>(y);
  const uint4* rp = reinterpret_cast<const uint4*>(residual);
  cudaStream_t st = (cudaStream_t)stream;
  // exactly one resident wave (62 registers: 4 blocks per SM, 70-74 with a residual: 3), every
  // block the same multiple of 4 * rows_step rows of the flattened (sample, row) space
  const int64_t total_rows = (int64_t)samples * spatial;
  const int64_t quantum = (int64_t)rows_step * 4;
#define MRI_GN_LAUNCH(S, R)                                                                       \
  do {                                                                                            \
    static int occ = 0; /* per instantiation; the block size only depends on C / 8 <= 256 */       \
    if (occ == 0) occ = blocks_per_sm(gn_apply_kernel<S, R>, 256);                                \
    const int bps = env_bps > 0 ? env_bps : occ;                                                  \
    int64_t blocks = 148LL * bps;                                                                 \
    int64_t rows_per = (total_rows + blocks - 1) / blocks;                                        \
    rows_per = (rows_per + quantum - 1) / quantum * quantum;                                      \
    blocks = (total_rows + rows_per - 1) / rows_per;                                              \
    gn_apply_kernel<S, R><<<(unsigned)blocks, threads, 0, st>>>(                                  \
        xp, yp, stats, gamma, beta, rowbias, rowbias_ld, rp, samples, spatial, C, groups,         \
        stats_ld, stats_g0, stats_cpg, eps, rows_per);                                            \
  } while (0)
  if (silu) {
    if (residual) MRI_GN_LAUNCH(true, true); else MRI_GN_LAUNCH(true, false);
  } else {
    if (residual) MRI_GN_LAUNCH(false, true); else MRI_GN_LAUNCH(false, false);
  }
#undef MRI_GN_LAUNCH
  return check_launch("gn_apply_kernel");
}
