// Small kernels around the tensor-core path: timestep embedding + the tiny fp32 linears,
// the explicit patch matrix for thin-channel first convolutions, layout converters at the
// public NC[D]HW fp32 edge, and the attention row softmax.
// Reference call sites: see include/mri_b200.h.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/mri_b200.h"
#include "common.h"

namespace mri {

// emb[b, j] = sin(t_b * f_j), emb[b, half + j] = cos(t_b * f_j); the frequency table
// f_j = exp(j * -(ln(1e4)/(half-1))) is built once on the host exactly as the reference builds
// it (fp32), so the products t*f are bit-identical; odd dim is zero padded.
__global__ void sinusoidal_kernel(const int64_t* __restrict__ t, const float* __restrict__ freqs,
                                  float* __restrict__ out, int batch, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * dim) return;
  const int b = i / dim, j = i % dim;
  if (j >= 2 * half) {
    out[i] = 0.f;
    return;
  }
  const int jj = j < half ? j : j - half;
  const float arg = __fmul_rn((float)t[b], __ldg(freqs + jj));
  out[i] = j < half ? sinf(arg) : cosf(arg);
}

// y[b, o] = act(bias[o] + sum_i x[b, i] W[o, i]) + addend[b, o]; one warp per output feature.
template <int kMaxB>
__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
              const float* __restrict__ bias, const float* __restrict__ addend,
              float* __restrict__ y, int batch, int in_f, int out_f, int act, int b0) {
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= out_f) return;
  float acc[kMaxB];
#pragma unroll
  for (int b = 0; b < kMaxB; ++b) acc[b] = 0.f;
  const float* w = W + (size_t)o * in_f;
  for (int i = lane; i < in_f; i += 32) {
    const float wv = __ldg(w + i);
#pragma unroll
    for (int b = 0; b < kMaxB; ++b)
      if (b0 + b < batch) acc[b] = fmaf(wv, __ldg(x + (size_t)(b0 + b) * in_f + i), acc[b]);
  }
#pragma unroll
  for (int b = 0; b < kMaxB; ++b) {
    float v = acc[b];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0 && b0 + b < batch) {
      v += bias != nullptr ? __ldg(bias + o) : 0.f;
      if (act == 1) v = v / (1.0f + expf(-v));
      if (addend != nullptr) v += addend[(size_t)(b0 + b) * out_f + o];
      y[(size_t)(b0 + b) * out_f + o] = v;
    }
  }
}

// Patch matrix of a thin-channel input (fp32 NC[D]HW) for a ksize^ndim, pad ksize/2 convolution.
// dst[m][((kd*k+kh)*k+kw)*cin_total + c]; one thread per (m, 8 output columns) -> one 16-byte
// store; the (tiny) source stays L1/L2 resident.  The column -> (tap shift, channel) decode is a
// shared-memory table built once per block, so the inner loop has no divisions.
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ src, const float* __restrict__ src2,
              uint4* __restrict__ dst, int samples, int cin, int cin2, int D, int H, int W,
              int k, int ndim, int kpad) {
  extern __shared__ int col_tab[];  // per column: (dd+8) | (dh+8)<<4 | (dw+8)<<8 | c<<12, or -1
  const int taps = ndim == 3 ? k * k * k : k * k;
  const int ct = cin + cin2;
  const int kvalid = taps * ct;
  const int pad = k / 2;
  for (int col = threadIdx.x; col < kpad; col += blockDim.x) {
    int e = -1;
    if (col < kvalid) {
      const int tap = col / ct, c = col - tap * ct;
      const int kw = tap % k, kh = (tap / k) % k, kd = ndim == 3 ? tap / (k * k) : pad;
      e = (kd - pad + 8) | ((kh - pad + 8) << 4) | ((kw - pad + 8) << 8) | (c << 12);
    }
    col_tab[col] = e;
  }
  __syncthreads();
  const int64_t spatial = (int64_t)D * H * W;
  const int64_t M = (int64_t)samples * spatial;
  const int vec_per_row = kpad >> 3;
  const int64_t total = M * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / vec_per_row;
    const int col0 = (int)(i - m * vec_per_row) * 8;
    const int n = (int)(m / spatial);
    int s = (int)(m - (int64_t)n * spatial);
    const int w0 = s % W;
    s /= W;
    const int h0 = s % H;
    const int d0 = s / H;
    const float* s1 = src + (size_t)n * cin * spatial;
    const float* s2 = src2 != nullptr ? src2 + (size_t)n * cin2 * spatial : nullptr;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = col_tab[col0 + j];
      float v = 0.f;
      if (e >= 0) {
        const int d = d0 + (e & 15) - 8, h = h0 + ((e >> 4) & 15) - 8, w = w0 + ((e >> 8) & 15) - 8;
        const int c = e >> 12;
        if ((unsigned)w < (unsigned)W && (unsigned)h < (unsigned)H && (unsigned)d < (unsigned)D) {
          const int64_t sp = ((int64_t)d * H + h) * W + w;
          v = c < cin ? __ldg(s1 + (size_t)c * spatial + sp) : __ldg(s2 + (size_t)(c - cin) * spatial + sp);
        }
      }
      f[j] = v;
    }
    uint32_t wd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
      wd[j] = *reinterpret_cast<uint32_t*>(&h2);
    }
    dst[i] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
  }
}

// Patch matrix from a channels-last bf16 copy of the thin input ([n][pos][cp], cp = channels padded
// to a multiple of 4): column (tap * cp + c).  One thread per 16-byte destination vector = two
// "quads" (4 channels of one tap): two 8-byte loads, bounds checks from a shared-memory table,
// no divisions in the loop.  ~5x fewer instructions per byte than im2col_kernel.
__global__ void __launch_bounds__(256)
im2col4_kernel(const uint2* __restrict__ src, uint4* __restrict__ dst, int samples, int cp, int D,
               int H, int W, int k, int ndim, int kpad) {
  extern __shared__ int quad_tab[];  // per quad: (dd+8) | (dh+8)<<4 | (dw+8)<<8 | cq<<12, or -1
  const int taps = ndim == 3 ? k * k * k : k * k;
  const int qpt = cp >> 2;           // quads per tap
  const int nquad = kpad >> 2;
  const int pad = k / 2;
  for (int q = threadIdx.x; q < nquad; q += blockDim.x) {
    int e = -1;
    const int tap = q / qpt;
    if (tap < taps) {
      const int kw = tap % k, kh = (tap / k) % k, kd = ndim == 3 ? tap / (k * k) : pad;
      e = (kd - pad + 8) | ((kh - pad + 8) << 4) | ((kw - pad + 8) << 8) | ((q - tap * qpt) << 12);
    }
    quad_tab[q] = e;
  }
  __syncthreads();
  const int64_t spatial = (int64_t)D * H * W;
  const int64_t M = (int64_t)samples * spatial;
  const int vec_per_row = kpad >> 3;
  const int64_t total = M * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / vec_per_row;
    const int v = (int)(i - m * vec_per_row);
    const int n = (int)(m / spatial);
    int s = (int)(m - (int64_t)n * spatial);
    const int w0 = s % W;
    s /= W;
    const int h0 = s % H;
    const int d0 = s / H;
    const uint2* sn = src + (size_t)n * spatial * qpt;
    uint2 r[2];
#pragma unroll
    for (int hq = 0; hq < 2; ++hq) {
      const int e = quad_tab[2 * v + hq];
      uint2 val = make_uint2(0u, 0u);
      if (e >= 0) {
        const int d = d0 + (e & 15) - 8, h = h0 + ((e >> 4) & 15) - 8, w = w0 + ((e >> 8) & 15) - 8;
        if ((unsigned)w < (unsigned)W && (unsigned)h < (unsigned)H && (unsigned)d < (unsigned)D)
          val = __ldg(sn + (((size_t)d * H + h) * W + w) * qpt + (e >> 12));
      }
      r[hq] = val;
    }
    dst[i] = make_uint4(r[0].x, r[0].y, r[1].x, r[1].y);
  }
}

// Thin-Cout convolution as "GEMM over taps, then gather": Y[q][tap*cout + co] = W[tap][co] . x[q]
// was computed for every position q by the tensor-core kernel; the convolution output is
//   out[o][co] = bias[co] + sum_tap Y[o + tap - pad][tap*cout + co]      (zero outside the volume)
// One thread per (output position, channel): reads 2 bytes from each of the k^ndim neighbouring
// rows of Y (L2-resident), writes bf16 channels-last with row pitch ldo (padding channels = 0).
__global__ void __launch_bounds__(256)
tap_gather_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out,
                  const float* __restrict__ bias, int samples, int D, int H, int W, int k, int ndim,
                  int cout, int ldy, int ldo) {
  const int64_t spatial = (int64_t)D * H * W;
  const int64_t total = (int64_t)samples * spatial * ldo;
  const int pad = k / 2;
  const int kd_n = ndim == 3 ? k : 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / ldo;
    const int co = (int)(i - m * ldo);
    float acc = 0.f;
    if (co < cout) {
      const int n = (int)(m / spatial);
      int s = (int)(m - (int64_t)n * spatial);
      const int w0 = s % W;
      s /= W;
      const int h0 = s % H;
      const int d0 = s / H;
      acc = bias != nullptr ? __ldg(bias + co) : 0.f;
      const __nv_bfloat16* yn = y + (size_t)n * spatial * ldy;
      int tap = 0;
      for (int kd = 0; kd < kd_n; ++kd) {
        const int d = ndim == 3 ? d0 + kd - pad : d0;
        for (int kh = 0; kh < k; ++kh) {
          const int h = h0 + kh - pad;
          for (int kw = 0; kw < k; ++kw, ++tap) {
            const int w = w0 + kw - pad;
            if ((unsigned)w < (unsigned)W && (unsigned)h < (unsigned)H && (unsigned)d < (unsigned)D)
              acc += __bfloat162float(yn[(((size_t)d * H + h) * W + w) * ldy + tap * cout + co]);
          }
        }
      }
    }
    out[i] = __float2bfloat16_rn(acc);
  }
}

// dst[n][c][s] (fp32) = src[n][s][c] (bf16, row pitch ldc); small C (<= 32) at the output edge.
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int samples,
                    int64_t spatial, int C, int ldc) {
  const int64_t total = (int64_t)samples * C * spatial;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i % spatial;
    const int c = (int)((i / spatial) % C);
    const int n = (int)(i / (spatial * C));
    dst[i] = __bfloat162float(src[((size_t)n * spatial + s) * ldc + c]);
  }
}

__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int samples,
                    int64_t spatial, int C, int ldc) {
  const int64_t total = (int64_t)samples * spatial * ldc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % ldc);
    const int64_t s = (i / ldc) % spatial;
    const int n = (int)(i / ((int64_t)ldc * spatial));
    const float v = c < C ? __ldg(src + ((size_t)n * C + c) * spatial + s) : 0.f;
    dst[i] = __float2bfloat16(v);
  }
}

// One warp per row: P = softmax(S * scale) in fp32, stored bf16.  The row lives in registers
// (up to kMaxPerLane * 32 columns), so S is read exactly once.
constexpr int kSoftmaxMaxPerLane = 64;
template <int kPerLane>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ P, int64_t rows,
                    int cols, int ld_s, int ld_p, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* s = S + row * ld_s;
  __nv_bfloat16* p = P + row * ld_p;
  float v[kPerLane];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const int c = lane + 32 * j;
    v[j] = c < cols ? __ldg(s + c) * scale : -INFINITY;
    mx = fmaxf(mx, v[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    v[j] = __expf(v[j] - mx);  // exp(-inf) = 0 for the padding columns
    sum += v[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const int c = lane + 32 * j;
    if (c < ld_p) p[c] = __float2bfloat16(v[j] * inv);
  }
}

// Re-pack of derived weight buffers after an optimizer step: every packed buffer (K-major bf16
// conv matrices, transposed / flipped dgrad matrices, concatenated projections ...) is a pure
// gather of parameter elements, so ONE launch refreshes them all from a table of segments.
// idx: (source slot << 28) | element offset inside that source, or -1 for a zero (padding).
__global__ void __launch_bounds__(256)
gather_pack_kernel(const MriGatherSeg* __restrict__ segs, int n_segs) {
  // binary search: last segment whose first block <= blockIdx.x
  int lo = 0, hi = n_segs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].block0 <= (int64_t)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const MriGatherSeg sg = segs[lo];
  const int64_t i0 = ((int64_t)blockIdx.x - sg.block0) * 2048 + threadIdx.x;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int64_t i = i0 + (int64_t)u * 256;
    if (i >= sg.n) break;
    const int32_t e = __ldg(sg.idx + i);
    float v = 0.f;
    if (e >= 0) {  // select, not a dynamically indexed copy of the array (that would live in local memory)
      const int sl = (e >> 28) & 3;
      const float* sp = sl == 0 ? sg.src[0] : (sl == 1 ? sg.src[1] : (sl == 2 ? sg.src[2] : sg.src[3]));
      v = __ldg(sp + (e & 0x0fffffff));
    }
    if (sg.dst_bf16) reinterpret_cast<__nv_bfloat16*>(sg.dst)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(sg.dst)[i] = v;
  }
}

// ---- strided copy with dtype conversion (small layout shuffles of the backward launch list) ----
struct Copy4 {
  int64_t shape[4], ss[4], ds[4];
};
template <typename T> __device__ __forceinline__ float ld_as_float(const void* p, int64_t i);
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const void* p, int64_t i) {
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
template <> __device__ __forceinline__ float ld_as_float<float>(const void* p, int64_t i) {
  return reinterpret_cast<const float*>(p)[i];
}
template <> __device__ __forceinline__ float ld_as_float<double>(const void* p, int64_t i) {
  return (float)reinterpret_cast<const double*>(p)[i];
}
template <typename TS>
__global__ void __launch_bounds__(256)
copy_cast_kernel(const void* __restrict__ src, void* __restrict__ dst, int dst_dtype, Copy4 c,
                 int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int64_t i3 = r % c.shape[3]; r /= c.shape[3];
    const int64_t i2 = r % c.shape[2]; r /= c.shape[2];
    const int64_t i1 = r % c.shape[1]; r /= c.shape[1];
    const float v = ld_as_float<TS>(src, r * c.ss[0] + i1 * c.ss[1] + i2 * c.ss[2] + i3 * c.ss[3]);
    const int64_t o = r * c.ds[0] + i1 * c.ds[1] + i2 * c.ds[2] + i3 * c.ds[3];
    if (dst_dtype == 0) reinterpret_cast<__nv_bfloat16*>(dst)[o] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(dst)[o] = v;
  }
}

static inline unsigned grid_for(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = 148 * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace mri

using namespace mri;

extern "C" int mri_sinusoidal(const int64_t* t, const float* freqs, float* out, int batch, int dim,
                              void* stream) {
  if (batch < 1 || dim < 4 || freqs == nullptr) return set_error(-2, "mri_sinusoidal: bad arguments");
  const int total = batch * dim;
  sinusoidal_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t, freqs, out, batch, dim);
  return check_launch("sinusoidal_kernel");
}

extern "C" int mri_linear(const float* x, const float* W, const float* bias, const float* addend,
                          float* y, int batch, int in_f, int out_f, int act, void* stream) {
  if (batch < 1 || in_f < 1 || out_f < 1) return set_error(-2, "mri_linear: bad shape");
  constexpr int kB = 8;
  const int warps = 8;
  const unsigned grid = (out_f + warps - 1) / warps;
  for (int b0 = 0; b0 < batch; b0 += kB) {
    linear_kernel<kB><<<grid, warps * 32, 0, (cudaStream_t)stream>>>(x, W, bias, addend, y, batch,
                                                                     in_f, out_f, act, b0);
    int rc = check_launch("linear_kernel");
    if (rc) return rc;
  }
  return 0;
}

extern "C" int mri_im2col(const float* src, const float* src2, void* dst, int samples, int cin,
                          int cin2, int D, int H, int W, int ksize, int ndim, int kpad,
                          void* stream) {
  const int taps = ndim == 3 ? ksize * ksize * ksize : ksize * ksize;
  if (ndim != 2 && ndim != 3) return set_error(-2, "mri_im2col: ndim must be 2 or 3");
  if ((cin + cin2) * taps > kpad || kpad % 64 != 0)
    return set_error(-2, "mri_im2col: kpad must be a multiple of 64 and >= taps*cin");
  if (cin2 > 0 && src2 == nullptr) return set_error(-2, "mri_im2col: src2 missing");
  const int64_t total = (int64_t)samples * D * H * W * (kpad / 8);
  if (ksize > 15 || cin + cin2 >= (1 << 18)) return set_error(-2, "mri_im2col: kernel size / channels too large");
  im2col_kernel<<<grid_for(total), 256, kpad * sizeof(int), (cudaStream_t)stream>>>(
      src, src2, reinterpret_cast<uint4*>(dst), samples, cin, cin2, D, H, W, ksize, ndim, kpad);
  return check_launch("im2col_kernel");
}

extern "C" int mri_im2col4(const void* src, void* dst, int samples, int cp, int D, int H, int W, int ksize,
                           int ndim, int kpad, void* stream) {
  if (ndim != 2 && ndim != 3) return set_error(-2, "mri_im2col4: ndim must be 2 or 3");
  const int taps = ndim == 3 ? ksize * ksize * ksize : ksize * ksize;
  if (cp < 4 || cp % 4 != 0 || kpad % 8 != 0 || taps * cp > kpad || ksize > 15)
    return set_error(-2, "mri_im2col4: channels must be padded to a multiple of 4 and fit kpad");
  const int64_t total = (int64_t)samples * D * H * W * (kpad / 8);
  im2col4_kernel<<<grid_for(total), 256, (kpad / 4) * sizeof(int), (cudaStream_t)stream>>>(
      reinterpret_cast<const uint2*>(src), reinterpret_cast<uint4*>(dst), samples, cp, D, H, W, ksize,
      ndim, kpad);
  return check_launch("im2col4_kernel");
}

extern "C" int mri_gather_pack(const MriGatherSeg* segs_dev, int n_segs, int64_t total_blocks,
                               void* stream) {
  if (n_segs < 1 || total_blocks < 1) return 0;
  if (total_blocks > 0x7fffffffLL) return set_error(-2, "mri_gather_pack: too many blocks");
  gather_pack_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(segs_dev, n_segs);
  return check_launch("gather_pack_kernel");
}

extern "C" int mri_tap_gather(const void* y, void* out, const float* bias, int samples, int D, int H,
                              int W, int ksize, int ndim, int cout, int ldy, int ldo, void* stream) {
  const int taps = ndim == 3 ? ksize * ksize * ksize : ksize * ksize;
  if (ndim != 2 && ndim != 3) return set_error(-2, "mri_tap_gather: ndim must be 2 or 3");
  if (cout < 1 || cout > ldo || taps * cout > ldy) return set_error(-2, "mri_tap_gather: bad channel layout");
  const int64_t total = (int64_t)samples * D * H * W * ldo;
  tap_gather_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<__nv_bfloat16*>(out), bias, samples,
      D, H, W, ksize, ndim, cout, ldy, ldo);
  return check_launch("tap_gather_kernel");
}

extern "C" int mri_nhwc_to_nchw(const void* src, float* dst, int samples, int64_t spatial, int C,
                                int ldc, void* stream) {
  const int64_t total = (int64_t)samples * C * spatial;
  nhwc_to_nchw_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), dst, samples, spatial, C, ldc);
  return check_launch("nhwc_to_nchw_kernel");
}

extern "C" int mri_nchw_to_nhwc(const float* src, void* dst, int samples, int64_t spatial, int C,
                                int ldc, void* stream) {
  const int64_t total = (int64_t)samples * spatial * ldc;
  nchw_to_nhwc_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), samples, spatial, C, ldc);
  return check_launch("nchw_to_nhwc_kernel");
}

extern "C" int mri_softmax_rows(const float* S, void* P, int64_t rows, int cols, int ld_s, int ld_p,
                                float scale, void* stream) {
  if (rows < 1 || cols < 1 || ld_s < cols || ld_p < cols)
    return set_error(-2, "mri_softmax_rows: bad shape");
  if (ld_p > kSoftmaxMaxPerLane * 32)
    return set_error(-2, "mri_softmax_rows: more than 2048 columns unsupported");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  __nv_bfloat16* Pp = reinterpret_cast<__nv_bfloat16*>(P);
  cudaStream_t st = (cudaStream_t)stream;
  const int per_lane = (ld_p + 31) / 32;
  if (per_lane <= 8) softmax_rows_kernel<8><<<grid, 256, 0, st>>>(S, Pp, rows, cols, ld_s, ld_p, scale);
  else if (per_lane <= 16) softmax_rows_kernel<16><<<grid, 256, 0, st>>>(S, Pp, rows, cols, ld_s, ld_p, scale);
  else if (per_lane <= 40) softmax_rows_kernel<40><<<grid, 256, 0, st>>>(S, Pp, rows, cols, ld_s, ld_p, scale);
  else softmax_rows_kernel<64><<<grid, 256, 0, st>>>(S, Pp, rows, cols, ld_s, ld_p, scale);
  return check_launch("softmax_rows_kernel");
}

extern "C" int mri_copy_cast(const void* src, int src_dtype, const int64_t* src_strides, void* dst,
                             int dst_dtype, const int64_t* dst_strides, const int64_t* shape,
                             void* stream) {
  if (src_dtype < 0 || src_dtype > 2 || dst_dtype < 0 || dst_dtype > 1)
    return set_error(-2, "mri_copy_cast: dtypes are 0 = bf16, 1 = f32, 2 = f64 (source only)");
  Copy4 c;
  int64_t total = 1;
  for (int i = 0; i < 4; ++i) {
    c.shape[i] = shape[i];
    c.ss[i] = src_strides[i];
    c.ds[i] = dst_strides[i];
    if (shape[i] < 1) return set_error(-2, "mri_copy_cast: empty shape");
    total *= shape[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for(total);
  if (src_dtype == 0) copy_cast_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, dst, dst_dtype, c, total);
  else if (src_dtype == 1) copy_cast_kernel<float><<<grid, 256, 0, st>>>(src, dst, dst_dtype, c, total);
  else copy_cast_kernel<double><<<grid, 256, 0, st>>>(src, dst, dst_dtype, c, total);
  return check_launch("copy_cast_kernel");
}

extern "C" int mri_memset_zero(void* ptr, int64_t bytes, void* stream) {
  if (bytes <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : set_cuda_error(e, "cudaMemsetAsync");
}
