// The data path in front of the UNets: what the reference's Dataset classes do to a loaded NIfTI
// volume on the CPU with numpy, one slice / volume per __getitem__, done here on the device for a
// whole batch of slices (or a whole volume) per launch.  All three kernels are HBM-bound
// element-wise / reduction work: coalesced reads along whichever axis is contiguous, fp64
// accumulation of the statistics, no intermediate tensors (normalise + clip + rescale are applied
// on the fly to the four bilinear neighbours / to the transposed tile).
// Reference call sites: see include/mri_b200.h.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mri_b200.h"
#include "common.h"

namespace mri {

__device__ __forceinline__ void warp_reduce3(double& c, double& s, double& q) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
}

// ---- statistics over the non-zero elements of each item ------------------------------------
// Two passes, as numpy's ndarray.std() makes them: PASS 0 accumulates (count, sum), the mean is
// rounded to fp32, PASS 1 accumulates sum (x - mean)^2 around that rounded mean -- no cancellation
// however small the deviation is beside the mean.  The second read of a <= 36 MB volume is served
// by the 126 MB L2.  acc[item] = {count, sum, sum of squared deviations} in fp64.
template <int PASS>
__device__ __forceinline__ void accumulate(float v, float mean, double& c, double& s, double& q) {
  if (v != 0.f) {
    if (PASS == 0) {
      c += 1.0;
      s += (double)v;
    } else {
      const double d = (double)v - (double)mean;
      q += d * d;
    }
  }
}

template <int PASS>
__device__ __forceinline__ void publish(double* a, double c, double s, double q) {
  if (PASS == 0) {
    if (c != 0.0) {
      atomicAdd(a, c);
      atomicAdd(a + 1, s);
    }
  } else if (q != 0.0) {
    atomicAdd(a + 2, q);
  }
}

// Variant A: the column axis is contiguous (col_stride == 1): a CTA walks a chunk of one item's
// rows with the threads along the columns.  grid = (chunks, items).
template <int PASS>
__global__ void __launch_bounds__(256)
masked_stats_cols_kernel(const float* __restrict__ x, int64_t item_stride, int64_t rows,
                         int64_t cols, int64_t row_stride, double* __restrict__ acc,
                         const float* __restrict__ mean_std) {
  const float* base = x + (int64_t)blockIdx.y * item_stride;
  const int64_t total = rows * cols;
  const float mean = PASS ? __ldg(mean_std + 2 * blockIdx.y) : 0.f;
  double c = 0.0, s = 0.0, q = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols, col = e - r * cols;
    accumulate<PASS>(__ldg(base + r * row_stride + col), mean, c, s, q);
  }
  warp_reduce3(c, s, q);
  __shared__ double sh[8][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[warp][0] = c; sh[warp][1] = s; sh[warp][2] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { c += sh[w][0]; s += sh[w][1]; q += sh[w][2]; }
    publish<PASS>(acc + 3 * (int64_t)blockIdx.y, c, s, q);
  }
}

// Variant B: the ITEM axis is contiguous (item_stride == 1, slices vol[:, :, z] of an (H, W, D)
// NIfTI array): lane <-> item, so a warp reads 32 consecutive floats; the 8 warps of a CTA take
// different pixels.  grid = (pixel chunks, ceil(items / 32)).
template <int PASS>
__global__ void __launch_bounds__(256)
masked_stats_items_kernel(const float* __restrict__ x, int items, int64_t rows, int64_t cols,
                          int64_t row_stride, int64_t col_stride, double* __restrict__ acc,
                          const float* __restrict__ mean_std) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.y * 32 + lane;
  const int64_t total = rows * cols;
  double c = 0.0, s = 0.0, q = 0.0;
  if (item < items) {
    const float mean = PASS ? __ldg(mean_std + 2 * item) : 0.f;
    for (int64_t p = (int64_t)blockIdx.x * 8 + warp; p < total; p += (int64_t)gridDim.x * 8) {
      const int64_t r = p / cols, col = p - r * cols;
      accumulate<PASS>(__ldg(x + r * row_stride + col * col_stride + item), mean, c, s, q);
    }
  }
  __shared__ double sh[8][32][3];
  sh[warp][lane][0] = c; sh[warp][lane][1] = s; sh[warp][lane][2] = q;
  __syncthreads();
  if (warp == 0 && item < items) {
    for (int w = 1; w < 8; ++w) { c += sh[w][lane][0]; s += sh[w][lane][1]; q += sh[w][lane][2]; }
    publish<PASS>(acc + 3 * (int64_t)item, c, s, q);
  }
}

// after pass 0: mean = fp32(sum / n).  After pass 1: std = fp32(sqrt(sum sq dev / n)), replaced
// by 1 when it is not positive or below eps.  An item without non-zero elements gets (0, 1).
__global__ void stats_finalize_kernel(const double* __restrict__ acc, float* __restrict__ mean_std,
                                      int items, float eps, int pass) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= items) return;
  const double n = acc[3 * i];
  if (pass == 0) {
    mean_std[2 * i] = n > 0.0 ? (float)(acc[3 * i + 1] / n) : 0.f;
    return;
  }
  float sd = 1.f;
  if (n > 0.0) {
    sd = (float)sqrt(acc[3 * i + 2] / n);
    if (!(sd > 0.f) || sd < eps) sd = 1.f;
  }
  mean_std[2 * i + 1] = sd;
}

// z-score of the non-zero elements, clip to +-clip, map to [0, 1]  (fp32, the reference's
// operation order: (v - mean) / std ; clip ; (v + clip) / (2 clip))
__device__ __forceinline__ float norm01(float v, float mean, float sd, float clip, float two_clip) {
  if (v != 0.f) v = __fdiv_rn(__fsub_rn(v, mean), sd);
  v = fminf(fmaxf(v, -clip), clip);
  return __fdiv_rn(__fadd_rn(v, clip), two_clip);
}

// ---- 2-D slices: normalise + bilinear resize (align_corners = False) + map to [-1, 1] --------
// One thread per output pixel, ox fastest.  The source index / weight arithmetic follows ATen's
// area_pixel_compute_source_index: src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in/out.
__global__ void __launch_bounds__(256)
slice_resize_kernel(const float* __restrict__ x, int64_t item_stride, int H, int W,
                    int64_t row_stride, int64_t col_stride, const float* __restrict__ mean_std,
                    int out_h, int out_w, float* __restrict__ out, int64_t out_item_stride,
                    float scale_h, float scale_w) {
  const int item = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= out_h * out_w) return;
  const int oy = o / out_w, ox = o - oy * out_w;
  const float mean = __ldg(mean_std + 2 * item), sd = __ldg(mean_std + 2 * item + 1);
  float sy = __fsub_rn(__fmul_rn(scale_h, (float)oy + 0.5f), 0.5f);
  float sx = __fsub_rn(__fmul_rn(scale_w, (float)ox + 0.5f), 0.5f);
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  int y0 = (int)sy, x0 = (int)sx;
  y0 = y0 > H - 1 ? H - 1 : y0;
  x0 = x0 > W - 1 ? W - 1 : x0;
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const float* b = x + (int64_t)item * item_stride;
  const float a00 = norm01(__ldg(b + y0 * row_stride + x0 * col_stride), mean, sd, 5.f, 10.f);
  const float a01 = norm01(__ldg(b + y0 * row_stride + x1 * col_stride), mean, sd, 5.f, 10.f);
  const float a10 = norm01(__ldg(b + y1 * row_stride + x0 * col_stride), mean, sd, 5.f, 10.f);
  const float a11 = norm01(__ldg(b + y1 * row_stride + x1 * col_stride), mean, sd, 5.f, 10.f);
  const float top = __fadd_rn(__fmul_rn(lx0, a00), __fmul_rn(lx1, a01));
  const float bot = __fadd_rn(__fmul_rn(lx0, a10), __fmul_rn(lx1, a11));
  const float v = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
  out[(int64_t)item * out_item_stride + o] = __fsub_rn(__fmul_rn(v, 2.f), 1.f);
}

// ---- 3-D volumes: normalise + zero pad + crop (+ the (H, W, D) -> (D, H, W) transposition) ---
struct PatchGeom {
  int D, H, W;            // logical extents of the source volume
  int64_t sd, sh, sw;     // element strides of the source along D, H, W
  int od, oh, ow;         // origin of the patch in source coordinates (negative = inside padding)
  int pd, ph, pw;         // patch extents
};

__device__ __forceinline__ float norm11(float v, float mean, float sd, float clip, float two_clip) {
  return __fsub_rn(__fmul_rn(norm01(v, mean, sd, clip, two_clip), 2.f), 1.f);
}

// source contiguous along W (already (D, H, W)): thread per output voxel, x fastest
__global__ void __launch_bounds__(256)
volume_patch_direct_kernel(const float* __restrict__ vol, PatchGeom g,
                           const float* __restrict__ mean_std, float clip,
                           float* __restrict__ out) {
  const int64_t total = (int64_t)g.pd * g.ph * g.pw;
  const float mean = __ldg(mean_std), sd = __ldg(mean_std + 1);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(e % g.pw);
    const int64_t r = e / g.pw;
    const int y = (int)(r % g.ph), z = (int)(r / g.ph);
    const int d = z + g.od, h = y + g.oh, w = x + g.ow;
    float v = 0.f;
    if (d >= 0 && d < g.D && h >= 0 && h < g.H && w >= 0 && w < g.W)
      v = norm11(__ldg(vol + d * g.sd + h * g.sh + w * g.sw), mean, sd, clip, 2.f * clip);
    out[e] = v;
  }
}

// source contiguous along D (nibabel's (H, W, D) array): a 32 x 32 (x, z) tile per CTA and y,
// read with z fastest, written with x fastest.  grid = (ceil(pw/32), ceil(pd/32), ph), block (32, 8)
__global__ void __launch_bounds__(256)
volume_patch_transpose_kernel(const float* __restrict__ vol, PatchGeom g,
                              const float* __restrict__ mean_std, float clip,
                              float* __restrict__ out) {
  __shared__ float tile[32][33];
  const float mean = __ldg(mean_std), sd = __ldg(mean_std + 1);
  const int y = blockIdx.z, h = y + g.oh;
  const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
  const bool h_in = h >= 0 && h < g.H;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xl = threadIdx.y + 8 * k, zl = threadIdx.x;
    const int w = x0 + xl + g.ow, d = z0 + zl + g.od;
    float v = 0.f;
    if (h_in && x0 + xl < g.pw && z0 + zl < g.pd && w >= 0 && w < g.W && d >= 0 && d < g.D)
      v = norm11(__ldg(vol + d * g.sd + h * g.sh + w * g.sw), mean, sd, clip, 2.f * clip);
    tile[xl][zl] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int zl = threadIdx.y + 8 * k, xl = threadIdx.x;
    const int z = z0 + zl, x = x0 + xl;
    if (z < g.pd && x < g.pw) out[((int64_t)z * g.ph + y) * g.pw + x] = tile[xl][zl];
  }
}

}  // namespace mri

using namespace mri;

extern "C" int mri_masked_stats(const float* x, int items, int64_t item_stride, int64_t rows,
                                int64_t cols, int64_t row_stride, int64_t col_stride, float eps,
                                double* acc, float* mean_std, void* stream) {
  if (items <= 0 || rows <= 0 || cols <= 0) return set_error(-1, "mri_masked_stats: empty input");
  if (items > 65535 * 32) return set_error(-1, "mri_masked_stats: too many items");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * 3 * (size_t)items, st);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync");
  const int64_t total = rows * cols;
  const bool by_cols = col_stride == 1;
  if (!by_cols && item_stride != 1)
    return set_error(-1, "mri_masked_stats: need col_stride == 1 or item_stride == 1");
  if (by_cols && items > 65535) return set_error(-1, "mri_masked_stats: too many items");
  const int groups = by_cols ? items : (items + 31) / 32;
  const int64_t per_cta = by_cols ? 256 * 16 : 8 * 16;
  int64_t chunks = (total + per_cta - 1) / per_cta;
  const int64_t cap = (148 * 8 + groups - 1) / groups;     // ~8 CTAs per SM over all items
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  const dim3 grid((unsigned)chunks, (unsigned)groups);
  for (int pass = 0; pass < 2; ++pass) {
    if (by_cols) {
      if (pass == 0)
        masked_stats_cols_kernel<0><<<grid, 256, 0, st>>>(x, item_stride, rows, cols, row_stride,
                                                          acc, mean_std);
      else
        masked_stats_cols_kernel<1><<<grid, 256, 0, st>>>(x, item_stride, rows, cols, row_stride,
                                                          acc, mean_std);
    } else {
      if (pass == 0)
        masked_stats_items_kernel<0><<<grid, 256, 0, st>>>(x, items, rows, cols, row_stride,
                                                           col_stride, acc, mean_std);
      else
        masked_stats_items_kernel<1><<<grid, 256, 0, st>>>(x, items, rows, cols, row_stride,
                                                           col_stride, acc, mean_std);
    }
    if (int rc = check_launch("masked_stats_kernel")) return rc;
    stats_finalize_kernel<<<(items + 127) / 128, 128, 0, st>>>(acc, mean_std, items, eps, pass);
    if (int rc = check_launch("stats_finalize_kernel")) return rc;
  }
  return 0;
}

extern "C" int mri_slice_normalize_resize(const float* x, int items, int64_t item_stride, int H,
                                          int W, int64_t row_stride, int64_t col_stride,
                                          const float* mean_std, int out_h, int out_w, float* out,
                                          int64_t out_item_stride, void* stream) {
  if (items <= 0 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0)
    return set_error(-1, "mri_slice_normalize_resize: empty input");
  if (items > 65535) return set_error(-1, "mri_slice_normalize_resize: too many items");
  const int n = out_h * out_w;
  slice_resize_kernel<<<dim3((n + 255) / 256, items), 256, 0, (cudaStream_t)stream>>>(
      x, item_stride, H, W, row_stride, col_stride, mean_std, out_h, out_w, out, out_item_stride,
      (float)H / (float)out_h, (float)W / (float)out_w);
  return check_launch("slice_resize_kernel");
}

extern "C" int mri_volume_normalize_patch(const float* vol, int D, int H, int W, int64_t stride_d,
                                          int64_t stride_h, int64_t stride_w, const float* mean_std,
                                          float clip, int off_d, int off_h, int off_w, int pd,
                                          int ph, int pw, float* out, void* stream) {
  if (D <= 0 || H <= 0 || W <= 0 || pd <= 0 || ph <= 0 || pw <= 0)
    return set_error(-1, "mri_volume_normalize_patch: empty input");
  if (!(clip > 0.f)) return set_error(-1, "mri_volume_normalize_patch: clip must be positive");
  PatchGeom g{D, H, W, stride_d, stride_h, stride_w, off_d, off_h, off_w, pd, ph, pw};
  cudaStream_t st = (cudaStream_t)stream;
  if (stride_d == 1 && stride_w != 1 && ph <= 65535 && (pd + 31) / 32 <= 65535) {
    volume_patch_transpose_kernel<<<dim3((pw + 31) / 32, (pd + 31) / 32, ph), dim3(32, 8), 0, st>>>(
        vol, g, mean_std, clip, out);
    return check_launch("volume_patch_transpose_kernel");
  }
  const int64_t total = (int64_t)pd * ph * pw;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  volume_patch_direct_kernel<<<(unsigned)blocks, 256, 0, st>>>(vol, g, mean_std, clip, out);
  return check_launch("volume_patch_direct_kernel");
}
