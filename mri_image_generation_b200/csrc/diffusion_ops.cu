// DDPM / DDIM arithmetic as single fused passes.  fp32, written with explicit round-to-nearest
// intrinsics (never contracted into FMAs) in the reference's exact association order, so the
// result is bit-identical to eager torch on the same inputs.
// Reference call sites: see include/mri_b200.h.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/mri_b200.h"
#include "common.h"

namespace mri {

__device__ __forceinline__ float load_eps(const void* eps, int ldc, int channels, int64_t per_sample,
                                          int sample, int64_t j) {
  if (ldc == 0) return __ldg(reinterpret_cast<const float*>(eps) + (size_t)sample * per_sample + j);
  // channels-last bf16 UNet output: element (c, s) of NC[D]HW index j = c*spatial + s
  const int64_t spatial = per_sample / channels;
  const int c = (int)(j / spatial);
  const int64_t s = j - (int64_t)c * spatial;
  const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(eps);
  return __bfloat162float(e[((size_t)sample * spatial + s) * ldc + c]);
}

__global__ void __launch_bounds__(256)
q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                const int64_t* __restrict__ t, const float* __restrict__ sqrt_ac,
                const float* __restrict__ sqrt_1mac, float* __restrict__ out, int64_t per_sample) {
  const int sample = blockIdx.y;
  const int64_t ts = t[sample];
  const float a = __ldg(sqrt_ac + ts), b = __ldg(sqrt_1mac + ts);
  const size_t base = (size_t)sample * per_sample;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < per_sample;
       j += (int64_t)gridDim.x * blockDim.x) {
    out[base + j] = __fadd_rn(__fmul_rn(a, x0[base + j]), __fmul_rn(b, noise[base + j]));
  }
}

__global__ void __launch_bounds__(256)
ddpm_step_kernel(const float* __restrict__ x, const void* __restrict__ eps, int ldc, int channels,
                 const float* __restrict__ noise, const int64_t* __restrict__ t,
                 const float* __restrict__ betas, const float* __restrict__ sqrt_1mac,
                 const float* __restrict__ sqrt_recip_alphas, const float* __restrict__ post_var,
                 float* __restrict__ out, int64_t per_sample) {
  const int sample = blockIdx.y;
  const int64_t ts = t[sample];
  // q = beta/s ; sd = sqrt(pv) ; w = mask*sd   (per-sample scalars, same ops as the reference)
  const float q = __fdiv_rn(__ldg(betas + ts), __ldg(sqrt_1mac + ts));
  const float c1 = __ldg(sqrt_recip_alphas + ts);
  const float mask = ts != 0 ? 1.0f : 0.0f;
  const float w = __fmul_rn(mask, __fsqrt_rn(__ldg(post_var + ts)));
  const size_t base = (size_t)sample * per_sample;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < per_sample;
       j += (int64_t)gridDim.x * blockDim.x) {
    const float e = load_eps(eps, ldc, channels, per_sample, sample, j);
    const float mean = __fmul_rn(c1, __fsub_rn(x[base + j], __fmul_rn(q, e)));
    out[base + j] = __fadd_rn(mean, __fmul_rn(w, noise[base + j]));
  }
}

__global__ void __launch_bounds__(256)
ddim_step_kernel(const float* __restrict__ x, const void* __restrict__ eps, int ldc, int channels,
                 const int64_t* __restrict__ t, const int64_t* __restrict__ t_prev,
                 const float* __restrict__ ac, float* __restrict__ out, int64_t per_sample) {
  const int sample = blockIdx.y;
  const float a_t = __ldg(ac + t[sample]);
  const float a_p = __ldg(ac + t_prev[sample]);
  const float sqrt_a_t = __fsqrt_rn(a_t);
  const float s1m_t = __fsqrt_rn(__fsub_rn(1.0f, a_t));
  const float denom = fmaxf(sqrt_a_t, 1e-8f);
  const float sqrt_a_p = __fsqrt_rn(a_p);
  const float s1m_p = __fsqrt_rn(__fsub_rn(1.0f, a_p));
  const size_t base = (size_t)sample * per_sample;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < per_sample;
       j += (int64_t)gridDim.x * blockDim.x) {
    const float e = load_eps(eps, ldc, channels, per_sample, sample, j);
    const float x0 = __fdiv_rn(__fsub_rn(x[base + j], __fmul_rn(s1m_t, e)), denom);
    out[base + j] = __fadd_rn(__fmul_rn(sqrt_a_p, x0), __fmul_rn(s1m_p, e));
  }
}

// per-sample sum of squared differences; grid (blocks, samples)
__global__ void __launch_bounds__(256)
sqdiff_kernel(const float* __restrict__ pred, const float* __restrict__ noise,
              float* __restrict__ acc, int64_t per_sample) {
  const int sample = blockIdx.y;
  const size_t base = (size_t)sample * per_sample;
  float s = 0.f;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < per_sample;
       j += (int64_t)gridDim.x * blockDim.x) {
    const float d = pred[base + j] - noise[base + j];
    s = fmaf(d, d, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += part[i];
    atomicAdd(acc + sample, tot);
  }
}

__global__ void loss_finalize_kernel(float* __restrict__ per_sample_acc, const int64_t* __restrict__ t,
                                     const float* __restrict__ snr, float gamma,
                                     float* __restrict__ loss_out, int samples, int64_t per_sample) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float tot = 0.f;
  for (int b = 0; b < samples; ++b) {
    const float mse = per_sample_acc[b] / (float)per_sample;
    float w = 1.0f;
    if (gamma > 0.f) {
      const float s = snr[t[b]];
      w = fminf(s, gamma) / s;
    }
    per_sample_acc[b] = mse;
    tot += w * mse;
  }
  loss_out[0] = tot / (float)samples;
}

__global__ void add_i64_kernel(int64_t* t, int n, int64_t delta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] += delta;
}

static inline dim3 ew_grid(int samples, int64_t per_sample) {
  int64_t bx = (per_sample + 255) / 256;
  int64_t cap = (148 * 8 + samples - 1) / samples;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)samples);
}

}  // namespace mri

using namespace mri;

extern "C" int mri_q_sample(const float* x0, const float* noise, const int64_t* t,
                            const float* sqrt_ac, const float* sqrt_1mac, float* out, int samples,
                            int64_t per_sample, void* stream) {
  if (samples < 1 || per_sample < 1) return set_error(-2, "mri_q_sample: empty input");
  q_sample_kernel<<<ew_grid(samples, per_sample), 256, 0, (cudaStream_t)stream>>>(
      x0, noise, t, sqrt_ac, sqrt_1mac, out, per_sample);
  return check_launch("q_sample_kernel");
}

extern "C" int mri_ddpm_step(const float* x, const void* eps, int eps_nhwc_ldc, int channels,
                             const float* noise, const int64_t* t, const float* betas,
                             const float* sqrt_1mac, const float* sqrt_recip_alphas,
                             const float* post_var, float* out, int samples, int64_t per_sample,
                             void* stream) {
  if (samples < 1 || per_sample < 1) return set_error(-2, "mri_ddpm_step: empty input");
  if (eps_nhwc_ldc != 0 && (channels < 1 || per_sample % channels != 0))
    return set_error(-2, "mri_ddpm_step: per_sample must be channels * spatial");
  ddpm_step_kernel<<<ew_grid(samples, per_sample), 256, 0, (cudaStream_t)stream>>>(
      x, eps, eps_nhwc_ldc, channels, noise, t, betas, sqrt_1mac, sqrt_recip_alphas, post_var, out,
      per_sample);
  return check_launch("ddpm_step_kernel");
}

extern "C" int mri_ddim_step(const float* x, const void* eps, int eps_nhwc_ldc, int channels,
                             const int64_t* t, const int64_t* t_prev, const float* alphas_cumprod,
                             float* out, int samples, int64_t per_sample, void* stream) {
  if (samples < 1 || per_sample < 1) return set_error(-2, "mri_ddim_step: empty input");
  if (eps_nhwc_ldc != 0 && (channels < 1 || per_sample % channels != 0))
    return set_error(-2, "mri_ddim_step: per_sample must be channels * spatial");
  ddim_step_kernel<<<ew_grid(samples, per_sample), 256, 0, (cudaStream_t)stream>>>(
      x, eps, eps_nhwc_ldc, channels, t, t_prev, alphas_cumprod, out, per_sample);
  return check_launch("ddim_step_kernel");
}

extern "C" int mri_minsnr_loss(const float* pred, const float* noise, const int64_t* t,
                               const float* snr, float gamma, float* per_sample_out,
                               float* loss_out, int samples, int64_t per_sample, void* stream) {
  if (samples < 1 || per_sample < 1) return set_error(-2, "mri_minsnr_loss: empty input");
  cudaError_t e = cudaMemsetAsync(per_sample_out, 0, sizeof(float) * samples, (cudaStream_t)stream);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(per_sample_out)");
  sqdiff_kernel<<<ew_grid(samples, per_sample), 256, 0, (cudaStream_t)stream>>>(
      pred, noise, per_sample_out, per_sample);
  int rc = check_launch("sqdiff_kernel");
  if (rc) return rc;
  loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(per_sample_out, t, snr, gamma, loss_out,
                                                          samples, per_sample);
  return check_launch("loss_finalize_kernel");
}

extern "C" int mri_add_i64(int64_t* t, int n, int64_t delta, void* stream) {
  if (n < 1) return 0;
  add_i64_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t, n, delta);
  return check_launch("add_i64_kernel");
}
