// "Split" precision mode (parity with the reference's fp32 / TF32 sampling path,
// ddpm_3d_ldm/show_model.py:254 runs the UNet without autocast).  The tensor-core kernels only take
// bf16 operands, so an fp32 value v travels as TWO bf16 numbers, hi = bf16(v) and lo = bf16(v - hi)
// (16 mantissa bits together), and a product a * w is evaluated as
//     a_hi * w_hi + a_lo * w_hi + a_hi * w_lo        (the dropped a_lo * w_lo term is ~2^-18)
// by widening the contraction: an activation with C channels is stored channels-last as
// [hi | lo | hi] (3C channels) and the weights as [w_hi | w_hi | w_lo] along their input-channel
// axis -- an ordinary convolution with 3C input channels for the implicit-GEMM kernel, which
// accumulates in fp32 and (out_f32) writes fp32.  The kernels here are the HBM-bound glue:
// GroupNorm(+SiLU) that reads the fp32 convolution output and writes the widened bf16 operand, a
// generic strided fp32 -> widened bf16 split, a row softmax that writes widened probabilities,
// and the low part of the fp32 NC[D]HW input for the first convolution.  Exact sigmoid (expf),
// no tanh.approx: this mode exists for accuracy, not speed.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/mri_b200.h"
#include "common.h"

namespace mri {

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// x fp32 [samples, spatial, C] -> v = act(GroupNorm(x)) (+ rowbias) (+ residual, fp32 [.., C]);
// stats == nullptr: v = x (+ ...).  Written as y bf16 [samples, spatial, 3C] = [hi | lo | hi] of v
// (y != nullptr) and / or as fp32 y32 [samples, spatial, C] (y32 != nullptr).
__global__ void __launch_bounds__(256)
gn_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, float* __restrict__ y32,
                const double* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ rowbias, int rowbias_ld,
                const float* __restrict__ residual, int64_t spatial, int C, int groups, int stats_ld,
                int stats_cpg, float eps, int silu, int64_t rows_per_block) {
  const int sample = blockIdx.y;
  const int c = threadIdx.x % C;           // host: blockDim.x is a multiple of C (C <= 256) or
  const int rsub = threadIdx.x / C;        // C is a multiple of blockDim.x handled by the c loop
  const int rows_step = blockDim.x / C > 0 ? blockDim.x / C : 1;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > spatial) r1 = spatial;
  for (int cc = c; cc < C; cc += blockDim.x) {
    float sc = 1.f, sh = 0.f, rb = 0.f;
    if (stats != nullptr) {
      const int cpg = C / groups, comb = cpg / stats_cpg, g = cc / cpg;
      const double* st = stats + ((size_t)sample * stats_ld + (size_t)g * comb) * 2;
      double s = 0.0, ss = 0.0;
      for (int j = 0; j < comb; ++j) {
        s += st[2 * j];
        ss += st[2 * j + 1];
      }
      const double inv_cnt = 1.0 / ((double)cpg * (double)spatial);
      const double mean = s * inv_cnt;
      double var = ss * inv_cnt - mean * mean;
      var = var < 0.0 ? 0.0 : var;
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      sc = rstd * gamma[cc];
      sh = beta[cc] - (float)mean * sc;
    }
    if (rowbias != nullptr) rb = rowbias[(size_t)sample * rowbias_ld + cc];
    for (int64_t r = r0 + rsub; r < r1; r += rows_step) {
      const size_t row = (size_t)sample * spatial + r;
      float v = x[row * C + cc];
      if (stats != nullptr) {
        v = fmaf(v, sc, sh);
        if (silu) v = v / (1.f + expf(-v));
      }
      v += rb;
      if (residual != nullptr) v += residual[row * C + cc];
      if (y32 != nullptr) y32[row * C + cc] = v;
      if (y != nullptr) {
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        __nv_bfloat16* o = y + row * (size_t)(3 * C);
        o[cc] = hi;
        o[C + cc] = lo;
        o[2 * C + cc] = hi;
      }
    }
  }
}

// generic: dst[o, i, p * seg + j] = part_p(src[o, i, j]), j < width, p = 0..2;
// pattern 0: (hi, lo, hi) -- the activation side; pattern 1: (hi, hi, lo) -- the weight side
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t outer, int inner,
              int width, int64_t src_outer_ld, int64_t src_inner_ld, int64_t dst_outer_ld,
              int64_t dst_inner_ld, int seg, int pattern) {
  const int64_t total = outer * inner * (int64_t)width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % width);
    const int64_t r = i / width;
    const int in = (int)(r % inner);
    const int64_t o = r / inner;
    const float v = src[o * src_outer_ld + in * src_inner_ld + j];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    __nv_bfloat16* d = dst + o * dst_outer_ld + in * dst_inner_ld + j;
    d[0] = hi;
    d[seg] = pattern == 0 ? lo : hi;
    d[2 * (int64_t)seg] = pattern == 0 ? hi : lo;
  }
}

// P = softmax(scale * S) row-wise in fp32 (one warp per row), written widened: [hi | lo | hi]
__global__ void __launch_bounds__(256)
softmax_split_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ P, int64_t rows, int cols,
                     int ld_s, int ld_p, int seg, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* s = S + row * ld_s;
  float m = -INFINITY;
  for (int j = lane; j < cols; j += 32) m = fmaxf(m, s[j] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int j = lane; j < cols; j += 32) sum += expf(s[j] * scale - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  __nv_bfloat16* p = P + row * ld_p;
  for (int j = lane; j < cols; j += 32) {
    const float v = expf(s[j] * scale - m) * inv;
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    p[j] = hi;
    p[seg + j] = lo;
    p[2 * seg + j] = hi;
  }
}

// x fp32 [B, C, S] (per_sample = C * S) -> out[b] (samples out_sample_stride elements apart):
// channels [0, C) = x - bf16(x), [C, 2C) = x
__global__ void __launch_bounds__(256)
nchw_lo_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t per_sample,
               int64_t out_sample_stride, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per_sample, r = i - b * per_sample;
    const float v = x[i];
    float* o = out + b * out_sample_stride + r;
    o[0] = v - __bfloat162float(__float2bfloat16_rn(v));
    o[per_sample] = v;
  }
}

// GroupNorm partial sums of an fp32 channels-last tensor: stats[sample][group] += (sum, sum of squares)
__global__ void __launch_bounds__(256)
stats_f32_kernel(const float* __restrict__ x, double* __restrict__ stats, int64_t spatial, int C,
                 int groups, int64_t rows_per_block) {
  extern __shared__ double red[];  // [groups][2]
  const int sample = blockIdx.y;
  const int cpg = C / groups;
  for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) red[i] = 0.0;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > spatial) r1 = spatial;
  const float* base = x + (size_t)sample * spatial * C;
  for (int g = 0; g < groups; ++g) {
    double s = 0.0, ss = 0.0;
    const int64_t n = (r1 - r0) * cpg;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const int64_t r = r0 + i / cpg;
      const int c = g * cpg + (int)(i % cpg);
      const double v = (double)base[r * C + c];
      s += v;
      ss += v * v;
    }
    atomicAdd(&red[2 * g], s);
    atomicAdd(&red[2 * g + 1], ss);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x)
    atomicAdd(stats + ((size_t)sample * groups) * 2 + i, red[i]);
}

static unsigned blocks_for(int64_t total) {
  int64_t b = (total + 255) / 256;
  if (b > 148 * 32) b = 148 * 32;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace mri

using namespace mri;

extern "C" int mri_gn_split(const float* x, void* y, float* y32, const double* stats, const float* gamma,
                            const float* beta, const float* rowbias, int rowbias_ld,
                            const float* residual, int samples, int64_t spatial, int C, int groups,
                            int stats_ld, int stats_cpg, float eps, int silu, void* stream) {
  if (C < 1 || samples < 1 || spatial < 1) return set_error(-2, "mri_gn_split: bad shape");
  if (y == nullptr && y32 == nullptr) return set_error(-2, "mri_gn_split: no output");
  if (stats != nullptr && (groups < 1 || C % groups != 0 || stats_cpg < 1 || (C / groups) % stats_cpg != 0))
    return set_error(-2, "mri_gn_split: bad group configuration");
  const int threads = C >= 256 ? 256 : (256 / C) * C;
  const int rows_step = threads / C > 0 ? threads / C : 1;
  int64_t want = (148 * 8 + samples - 1) / samples;
  int64_t rows_per = (spatial + want - 1) / want;
  if (rows_per < rows_step) rows_per = rows_step;
  const int64_t chunks = (spatial + rows_per - 1) / rows_per;
  if (samples > 65535) return set_error(-2, "mri_gn_split: too many samples");
  gn_split_kernel<<<dim3((unsigned)chunks, (unsigned)samples), threads, 0, (cudaStream_t)stream>>>(
      x, reinterpret_cast<__nv_bfloat16*>(y), y32, stats, gamma, beta, rowbias, rowbias_ld, residual,
      spatial, C, groups, stats_ld, stats_cpg, eps, silu, rows_per);
  return check_launch("gn_split_kernel");
}

extern "C" int mri_split3(const float* src, void* dst, int64_t outer, int inner, int width,
                          int64_t src_outer_ld, int64_t src_inner_ld, int64_t dst_outer_ld,
                          int64_t dst_inner_ld, int seg, int pattern, void* stream) {
  if (outer < 1 || inner < 1 || width < 1 || seg < width || (pattern != 0 && pattern != 1))
    return set_error(-2, "mri_split3: bad arguments");
  split3_kernel<<<blocks_for(outer * inner * (int64_t)width), 256, 0, (cudaStream_t)stream>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), outer, inner, width, src_outer_ld, src_inner_ld,
      dst_outer_ld, dst_inner_ld, seg, pattern);
  return check_launch("split3_kernel");
}

extern "C" int mri_softmax_rows_split(const float* S, void* P, int64_t rows, int cols, int ld_s, int ld_p,
                                      int seg, float scale, void* stream) {
  if (rows < 1 || cols < 1 || seg < cols || ld_p < 3 * seg) return set_error(-2, "mri_softmax_rows_split: bad arguments");
  const int64_t blocks = (rows + 7) / 8;
  if (blocks > 0x7fffffffLL) return set_error(-2, "mri_softmax_rows_split: too many rows");
  softmax_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      S, reinterpret_cast<__nv_bfloat16*>(P), rows, cols, ld_s, ld_p, seg, scale);
  return check_launch("softmax_split_kernel");
}

extern "C" int mri_bf16_residual_nchw(const float* x, float* out, int samples, int64_t per_sample,
                                      int64_t out_sample_stride, void* stream) {
  if (samples < 1 || per_sample < 1 || out_sample_stride < 2 * per_sample)
    return set_error(-2, "mri_bf16_residual_nchw: bad shape");
  const int64_t total = (int64_t)samples * per_sample;
  nchw_lo_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(x, out, per_sample,
                                                                      out_sample_stride, total);
  return check_launch("nchw_lo_kernel");
}

extern "C" int mri_stats_f32(const float* x, double* stats, int samples, int64_t spatial, int C,
                             int groups, void* stream) {
  if (C < 1 || samples < 1 || spatial < 1 || groups < 1 || C % groups != 0 || groups > 512 || samples > 65535)
    return set_error(-2, "mri_stats_f32: bad shape");
  int64_t want = (148 * 4 + samples - 1) / samples;
  int64_t rows_per = (spatial + want - 1) / want;
  if (rows_per < 1) rows_per = 1;
  const int64_t chunks = (spatial + rows_per - 1) / rows_per;
  stats_f32_kernel<<<dim3((unsigned)chunks, (unsigned)samples), 256, 2 * groups * sizeof(double),
                     (cudaStream_t)stream>>>(x, stats, spatial, C, groups, rows_per);
  return check_launch("stats_f32_kernel");
}
