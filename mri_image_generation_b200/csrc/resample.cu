// Image sizes that the down-sampling factor does not divide (slice_cond_2d_ddpm/unet.py:95-99:
// `if x.shape[-2:] != skip.shape[-2:]: x = F.interpolate(x, size=skip.shape[-2:], mode="bilinear",
// align_corners=False)`).  Two HBM-bound helpers on channels-last bf16 tensors:
//   * a window copy (zero padding an odd level to an even size in front of the stride-2
//     convolution, cropping its output back; the adjoints are the same copy with the roles
//     swapped), 16-byte vectors;
//   * the bilinear resize with ATen's source-index rule and its adjoint in GATHER form (each input
//     pixel sums the few output pixels that read it: deterministic, no atomics).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

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

// dst[b, dh0 + i, dw0 + j, :] = src[b, sh0 + i, sw0 + j, :] (+ add at the dst index)
__global__ void __launch_bounds__(256)
copy_window_kernel(const uint4* __restrict__ src, uint4* dst, const uint4* add, int B, int sH,
                   int sW, int dH, int dW, int vpr, int sh0, int sw0, int dh0, int dw0, int nH,
                   int nW) {
  const int64_t total = (int64_t)B * nH * nW * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vpr);
    int64_t r = i / vpr;
    const int x = (int)(r % nW);
    r /= nW;
    const int y = (int)(r % nH);
    const int b = (int)(r / nH);
    const int64_t so = (((int64_t)b * sH + sh0 + y) * sW + sw0 + x) * vpr + cv;
    const int64_t dofs = (((int64_t)b * dH + dh0 + y) * dW + dw0 + x) * vpr + cv;
    uint4 v = __ldg(src + so);
    if (add != nullptr) {
      float a[8], c[8];
      unpack8(v, a);
      unpack8(add[dofs], c);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += c[j];
      v = pack8(a);
    }
    dst[dofs] = v;
  }
}

// ATen area_pixel_compute_source_index(scale, dst, align_corners = false, cubic = false)
struct Tap {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Tap source_taps(int d, float scale, int in_size) {
  float s = scale * ((float)d + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  Tap t;
  t.i0 = (int)s;
  if (t.i0 > in_size - 1) t.i0 = in_size - 1;
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.l1 = s - (float)t.i0;
  t.l0 = 1.f - t.l1;
  return t;
}

__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int B, int sH,
                       int sW, int dH, int dW, int vpr, float scale_h, float scale_w) {
  const int64_t total = (int64_t)B * dH * dW * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vpr);
    int64_t r = i / vpr;
    const int x = (int)(r % dW);
    r /= dW;
    const int y = (int)(r % dH);
    const int b = (int)(r / dH);
    const Tap th = source_taps(y, scale_h, sH), tw = source_taps(x, scale_w, sW);
    const uint4* base = src + (int64_t)b * sH * sW * vpr + cv;
    float p00[8], p01[8], p10[8], p11[8], o[8];
    unpack8(__ldg(base + ((int64_t)th.i0 * sW + tw.i0) * vpr), p00);
    unpack8(__ldg(base + ((int64_t)th.i0 * sW + tw.i1) * vpr), p01);
    unpack8(__ldg(base + ((int64_t)th.i1 * sW + tw.i0) * vpr), p10);
    unpack8(__ldg(base + ((int64_t)th.i1 * sW + tw.i1) * vpr), p11);
#pragma unroll
    for (int j = 0; j < 8; ++j)  // ATen's association: h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d)
      o[j] = th.l0 * (tw.l0 * p00[j] + tw.l1 * p01[j]) + th.l1 * (tw.l0 * p10[j] + tw.l1 * p11[j]);
    dst[i] = pack8(o);
  }
}

// range of output indices that can read input index i (conservative; the taps decide)
__device__ __forceinline__ void reader_range(int i, float scale, int in_size, int out_size, int& lo,
                                             int& hi) {
  const float inv = 1.f / scale;
  lo = (int)floorf(((float)i - 0.5f) * inv - 0.5f) - 1;
  hi = (int)ceilf(((float)i + 1.5f) * inv - 0.5f) + 1;
  if (i == 0 || lo < 0) lo = 0;                         // negative source indices clamp to 0
  if (i == in_size - 1 || hi > out_size - 1) hi = out_size - 1;
}

__global__ void __launch_bounds__(256)
resize_bilinear_bwd_kernel(const uint4* __restrict__ dy, uint4* dx, const uint4* add, int B, int sH,
                           int sW, int dH, int dW, int vpr, float scale_h, float scale_w) {
  const int64_t total = (int64_t)B * sH * sW * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vpr);
    int64_t r = i / vpr;
    const int x = (int)(r % sW);
    r /= sW;
    const int y = (int)(r % sH);
    const int b = (int)(r / sH);
    int ylo, yhi, xlo, xhi;
    reader_range(y, scale_h, sH, dH, ylo, yhi);
    reader_range(x, scale_w, sW, dW, xlo, xhi);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const uint4* base = dy + (int64_t)b * dH * dW * vpr + cv;
    for (int oy = ylo; oy <= yhi; ++oy) {
      const Tap th = source_taps(oy, scale_h, sH);
      const float wy = (th.i0 == y ? th.l0 : 0.f) + (th.i1 == y ? th.l1 : 0.f);
      if (wy == 0.f) continue;
      for (int ox = xlo; ox <= xhi; ++ox) {
        const Tap tw = source_taps(ox, scale_w, sW);
        const float wx = (tw.i0 == x ? tw.l0 : 0.f) + (tw.i1 == x ? tw.l1 : 0.f);
        if (wx == 0.f) continue;
        float g[8];
        unpack8(__ldg(base + ((int64_t)oy * dW + ox) * vpr), g);
        const float w = wy * wx;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, g[j], acc[j]);
      }
    }
    if (add != nullptr) {
      float c[8];
      unpack8(add[i], c);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += c[j];
    }
    dx[i] = pack8(acc);
  }
}

static unsigned grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace mri

using namespace mri;

extern "C" int mri_copy_window_nhwc(const void* src, void* dst, const void* add, int B, int sH,
                                    int sW, int dH, int dW, int C, int sh0, int sw0, int dh0,
                                    int dw0, int nH, int nW, void* stream) {
  if (C % 8 != 0 || B < 1 || nH < 1 || nW < 1) return set_error(-2, "mri_copy_window_nhwc: bad shape");
  if (sh0 < 0 || sw0 < 0 || dh0 < 0 || dw0 < 0 || sh0 + nH > sH || sw0 + nW > sW || dh0 + nH > dH ||
      dw0 + nW > dW)
    return set_error(-2, "mri_copy_window_nhwc: window outside a tensor");
  const int64_t total = (int64_t)B * nH * nW * (C / 8);
  copy_window_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst),
      reinterpret_cast<const uint4*>(add), B, sH, sW, dH, dW, C / 8, sh0, sw0, dh0, dw0, nH, nW);
  return check_launch("copy_window_kernel");
}

extern "C" int mri_resize_bilinear_nhwc(const void* src, void* dst, int B, int sH, int sW, int dH,
                                        int dW, int C, void* stream) {
  if (C % 8 != 0 || B < 1 || sH < 1 || sW < 1 || dH < 1 || dW < 1)
    return set_error(-2, "mri_resize_bilinear_nhwc: bad shape");
  const int64_t total = (int64_t)B * dH * dW * (C / 8);
  resize_bilinear_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), B, sH, sW, dH, dW, C / 8,
      (float)sH / (float)dH, (float)sW / (float)dW);
  return check_launch("resize_bilinear_kernel");
}

extern "C" int mri_resize_bilinear_nhwc_bwd(const void* dy, void* dx, const void* add, int B, int sH,
                                            int sW, int dH, int dW, int C, void* stream) {
  if (C % 8 != 0 || B < 1 || sH < 1 || sW < 1 || dH < 1 || dW < 1)
    return set_error(-2, "mri_resize_bilinear_nhwc_bwd: bad shape");
  const int64_t total = (int64_t)B * sH * sW * (C / 8);
  resize_bilinear_bwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(dy), reinterpret_cast<uint4*>(dx),
      reinterpret_cast<const uint4*>(add), B, sH, sW, dH, dW, C / 8, (float)sH / (float)dH,
      (float)sW / (float)dW);
  return check_launch("resize_bilinear_bwd_kernel");
}
