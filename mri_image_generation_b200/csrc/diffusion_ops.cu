// DDPM / DDIM arithmetic as single fused passes.  fp32, written with explicit round-to-nearest
// intrinsics (never contracted into FMAs) in the reference's exact association order, so the
// result is bit-identical to eager torch on the same inputs.
// Reference call sites: see include/mri_b200.h.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/mri_b200.h"
#include "common.h"

namespace mri {

// ------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller with ATen's element mapping, so that the draws are bit-identical to
// torch.randn / randn_like on the same device for the same (seed, offset):
//   ATen launches G = min(#SM * (maxThreadsPerSM / 256), ceil(numel / 256)) blocks of 256 threads;
//   thread j (subsequence j) makes one curand_normal4 call per grid-stride iteration `it`
//   (Philox counter offset / 4 + it) and component ii of that call goes to element
//   e = it * 4 * Tt + ii * Tt + j with Tt = 256 * G; the generator offset advances by
//   4 * ceil(numel / (4 * Tt))  (ATen/native/cuda/DistributionTemplates.h: calc_execution_policy,
//   distribution_elementwise_grid_stride_kernel, normal_and_transform; curand_normal.h:
//   _curand_box_muller).  Here one thread owns FOUR consecutive subsequences, so that for each
//   component it holds four consecutive elements: 16-byte loads and stores.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    if (i < 9) {
      k.x += 0x9E3779B9u;
      k.y += 0xBB67AE85u;
    }
  }
  return c;
}

// two standard normals from two 32-bit draws: the operations (and their roundings) of
// _curand_box_muller followed by ATen's `rand * std + mean` with std = 1, mean = 0
__device__ __forceinline__ float2 box_muller(uint32_t x, uint32_t y) {
  const float kInv = 2.3283064e-10f;
  const float kInv2Pi = 2.3283064e-10f * 6.2831855f;
  const float u = fmaf((float)x, kInv, kInv / 2);
  const float v = fmaf((float)y, kInv2Pi, kInv2Pi / 2);
  const float s = sqrtf(-2.0f * logf(u));
  float sn, cs;
  __sincosf(v, &sn, &cs);
  return make_float2(fmaf(sn * s, 1.0f, 0.0f), fmaf(cs * s, 1.0f, 0.0f));
}

struct AtenGrid {
  long long Tt;     // threads of ATen's launch
  int n_iter;       // grid-stride iterations = curand_normal4 calls per thread
};

static int g_sm_count = 0, g_threads_per_sm = 0;
static int device_props() {
  if (g_sm_count != 0) return 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess)
    e = cudaDeviceGetAttribute(&g_threads_per_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
  if (e != cudaSuccess) {
    g_sm_count = 0;
    return set_cuda_error(e, "cudaDeviceGetAttribute(SM count / threads per SM)");
  }
  return 0;
}

static AtenGrid aten_grid(long long numel) {
  long long grid = (numel + 255) / 256;
  const long long cap = (long long)g_sm_count * (g_threads_per_sm / 256);
  if (grid > cap) grid = cap;
  AtenGrid g;
  g.Tt = grid * 256;
  g.n_iter = (int)((numel - 1) / (g.Tt * 4) + 1);
  return g;
}

// z[m][ii]: component ii of subsequence j0 + m at counter c
__device__ __forceinline__ void draw16(uint64_t seed, uint64_t c, uint64_t j0, float (&z)[4][4]) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const uint64_t sub = j0 + (uint64_t)m;
    const uint4 r = philox4x32_10(
        make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)sub, (uint32_t)(sub >> 32)), key);
    const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
    z[m][0] = a.x;
    z[m][1] = a.y;
    z[m][2] = b.x;
    z[m][3] = b.y;
  }
}

// per-sample coefficients of one reverse step; the same fp32 operations as the reference:
// q = beta / s ; c1 = sqrt_recip_alphas ; w = mask * sqrt(pv)
struct StepCoef {
  float q, c1, w;
};
struct StepTables {
  const int64_t* t;
  const float *betas, *sqrt_1mac, *sqrt_recip_alphas, *post_var;
};
__device__ __forceinline__ StepCoef step_coef(const StepTables& T, long long n) {
  const int64_t ts = T.t[n];
  StepCoef c;
  c.q = __fdiv_rn(__ldg(T.betas + ts), __ldg(T.sqrt_1mac + ts));
  c.c1 = __ldg(T.sqrt_recip_alphas + ts);
  c.w = __fmul_rn(ts != 0 ? 1.0f : 0.0f, __fsqrt_rn(__ldg(T.post_var + ts)));
  return c;
}

// eps of NC[D]HW element (sample n, in-sample index j): fp32 in x's layout (ldc = 0) or the bf16
// channels-last UNet output [n][spatial][ldc]
__device__ __forceinline__ float load_eps(const void* eps, int ldc, long long spatial,
                                          long long per_sample, long long n, long long j) {
  if (ldc == 0) return reinterpret_cast<const float*>(eps)[n * per_sample + j];
  const long long c = j / spatial;
  const long long s = j - c * spatial;
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(eps)[(n * spatial + s) * ldc + c]);
}

// The element loop shared by the kernels below.  MODE_RNG: z comes from Philox (ATen mapping);
// otherwise from `noise` (may be null when F ignores z).  F(e, z) handles ONE element; the
// loop hands it four consecutive elements at a time, so the compiler keeps 16-byte accesses where
// F's loads / stores are expressed through the Vec4 helpers.
template <bool RNG, class F4, class F1>
__device__ __forceinline__ void element_loop(long long numel, long long Tt, int n_iter,
                                             const uint64_t* rng, const float* noise, F4 f4, F1 f1) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long j0 = 4 * k;
  if (j0 >= Tt) return;
  uint64_t seed = 0, ctr0 = 0;
  if (RNG) {
    seed = rng[0];
    ctr0 = rng[1] >> 2;
  }
  const bool n_al = (reinterpret_cast<uintptr_t>(noise) & 15) == 0;
  for (int it = 0; it < n_iter; ++it) {
    const long long base = (long long)it * 4 * Tt + j0;
    if (base >= numel) break;
    float z[4][4];
    if (RNG) draw16(seed, ctr0 + (uint64_t)it, (uint64_t)j0, z);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const long long e0 = base + (long long)ii * Tt;
      if (e0 >= numel) break;
      float zz[4] = {0.f, 0.f, 0.f, 0.f};
      const bool full = e0 + 3 < numel;
      if (RNG) {
#pragma unroll
        for (int m = 0; m < 4; ++m) zz[m] = z[m][ii];
      } else if (noise != nullptr) {
        if (full && n_al) {
          const float4 v = *reinterpret_cast<const float4*>(noise + e0);
          zz[0] = v.x; zz[1] = v.y; zz[2] = v.z; zz[3] = v.w;
        } else {
#pragma unroll
          for (int m = 0; m < 4; ++m)
            if (e0 + m < numel) zz[m] = noise[e0 + m];
        }
      }
      if (full) {
        f4(e0, zz);
      } else {
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (e0 + m < numel) f1(e0 + m, zz[m]);
      }
    }
  }
}

__device__ __forceinline__ bool aligned16(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}
__device__ __forceinline__ void load4(const float* p, bool al, float (&v)[4]) {
  if (al) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int m = 0; m < 4; ++m) v[m] = p[m];
  }
}
__device__ __forceinline__ void store4(float* p, bool al, const float (&v)[4]) {
  if (al) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int m = 0; m < 4; ++m) p[m] = v[m];
  }
}

template <bool RNG>
__global__ void __launch_bounds__(256)
randn_kernel(float* out, long long numel, long long Tt, int n_iter, const uint64_t* rng) {
  const bool al = aligned16(out);
  element_loop<RNG>(numel, Tt, n_iter, rng, nullptr,
      [&](long long e0, const float (&z)[4]) { store4(out + e0, al, z); },
      [&](long long e, float z) { out[e] = z; });
}

// x_t = sqrt_ac[t] * x0 + sqrt_1mac[t] * z ; z optionally written out (the loss needs it).
// x0 / out may alias (no __restrict__).
template <bool RNG>
__global__ void __launch_bounds__(256)
q_sample_kernel(const float* x0, const float* noise, const uint64_t* rng, const int64_t* t,
                const float* sqrt_ac, const float* sqrt_1mac, float* out, float* noise_out,
                long long per_sample, long long numel, long long Tt, int n_iter) {
  const bool al = aligned16(x0) && aligned16(out) && (noise_out == nullptr || aligned16(noise_out));
  auto one = [&](long long e, float z) {
    const int64_t ts = t[e / per_sample];
    out[e] = __fadd_rn(__fmul_rn(__ldg(sqrt_ac + ts), x0[e]), __fmul_rn(__ldg(sqrt_1mac + ts), z));
    if (noise_out != nullptr) noise_out[e] = z;
  };
  element_loop<RNG>(numel, Tt, n_iter, rng, noise,
      [&](long long e0, const float (&z)[4]) {
        const long long n = e0 / per_sample;
        if ((e0 + 3) / per_sample != n) {
#pragma unroll
          for (int m = 0; m < 4; ++m) one(e0 + m, z[m]);
          return;
        }
        const int64_t ts = t[n];
        const float a = __ldg(sqrt_ac + ts), b = __ldg(sqrt_1mac + ts);
        float xv[4], o[4];
        load4(x0 + e0, al, xv);
#pragma unroll
        for (int m = 0; m < 4; ++m) o[m] = __fadd_rn(__fmul_rn(a, xv[m]), __fmul_rn(b, z[m]));
        store4(out + e0, al, o);
        if (noise_out != nullptr) store4(noise_out + e0, al, z);
      },
      one);
}

// x_{t-1} = c1 * (x - q * eps) + w * z.  x / out may alias (the sampler updates its state in place).
template <bool RNG>
__global__ void __launch_bounds__(256)
ddpm_step_kernel(const float* x, const void* eps, int ldc, int channels, const float* noise,
                 const uint64_t* rng, StepTables T, float* out, long long per_sample,
                 long long numel, long long Tt, int n_iter) {
  const long long spatial = per_sample / (channels > 0 ? channels : 1);
  const bool al = aligned16(x) && aligned16(out);
  const bool eps_al = ldc == 0 && aligned16(eps);
  auto one = [&](long long e, float z) {
    const long long n = e / per_sample;
    const StepCoef c = step_coef(T, n);
    const float ev = load_eps(eps, ldc, spatial, per_sample, n, e - n * per_sample);
    out[e] = __fadd_rn(__fmul_rn(c.c1, __fsub_rn(x[e], __fmul_rn(c.q, ev))), __fmul_rn(c.w, z));
  };
  element_loop<RNG>(numel, Tt, n_iter, rng, noise,
      [&](long long e0, const float (&z)[4]) {
        const long long n = e0 / per_sample;
        const long long j = e0 - n * per_sample;
        // four elements inside one sample and, for the channels-last eps, inside one channel
        const bool same = (j + 3 < per_sample) && (ldc == 0 || (j % spatial) + 3 < spatial);
        if (!same) {
#pragma unroll
          for (int m = 0; m < 4; ++m) one(e0 + m, z[m]);
          return;
        }
        const StepCoef c = step_coef(T, n);
        float xv[4], ev[4], o[4];
        load4(x + e0, al, xv);
        if (ldc == 0) {
          load4(reinterpret_cast<const float*>(eps) + e0, eps_al, ev);
        } else {
          const long long ch = j / spatial;
          const long long s = j - ch * spatial;
          const __nv_bfloat16* ep =
              reinterpret_cast<const __nv_bfloat16*>(eps) + (n * spatial + s) * ldc + ch;
#pragma unroll
          for (int m = 0; m < 4; ++m) ev[m] = __bfloat162float(ep[(long long)m * ldc]);
        }
#pragma unroll
        for (int m = 0; m < 4; ++m)
          o[m] = __fadd_rn(__fmul_rn(c.c1, __fsub_rn(xv[m], __fmul_rn(c.q, ev[m]))),
                           __fmul_rn(c.w, z[m]));
        store4(out + e0, al, o);
      },
      one);
}

// deterministic DDIM step (eta = 0); x / out may alias
__global__ void __launch_bounds__(256)
ddim_step_kernel(const float* x, const void* eps, int ldc, int channels, const int64_t* t,
                 const int64_t* t_prev, const float* ac, float* out, long long per_sample,
                 long long numel, long long Tt, int n_iter) {
  const long long spatial = per_sample / (channels > 0 ? channels : 1);
  auto one = [&](long long e, float) {
    const long long n = e / per_sample;
    const float a_t = __ldg(ac + t[n]);
    const float a_p = __ldg(ac + t_prev[n]);
    const float sqrt_a_t = __fsqrt_rn(a_t);
    const float s1m_t = __fsqrt_rn(__fsub_rn(1.0f, a_t));
    const float denom = fmaxf(sqrt_a_t, 1e-8f);
    const float sqrt_a_p = __fsqrt_rn(a_p);
    const float s1m_p = __fsqrt_rn(__fsub_rn(1.0f, a_p));
    const float ev = load_eps(eps, ldc, spatial, per_sample, n, e - n * per_sample);
    const float x0 = __fdiv_rn(__fsub_rn(x[e], __fmul_rn(s1m_t, ev)), denom);
    out[e] = __fadd_rn(__fmul_rn(sqrt_a_p, x0), __fmul_rn(s1m_p, ev));
  };
  element_loop<false>(numel, Tt, n_iter, nullptr, nullptr,
      [&](long long e0, const float (&z)[4]) {
#pragma unroll
        for (int m = 0; m < 4; ++m) one(e0 + m, z[m]);
      },
      one);
}

// ------------------------------------------------------------------------------------------
// Thin output convolution finished INSIDE the reverse step.  The k^d convolution with 1..4 output
// channels (out_conv) is computed as one GEMM over all taps, Y[q][tap * cout + co] = W[tap][co] . a[q]
// (the activation is read once); this kernel sums the shifted taps -- eps[q][co] = bias[co] +
// sum_tap Y[q + off(tap)][tap][co], in the same order and with the same bf16 rounding of the result
// as mri_tap_gather -- and applies the DDPM (noise drawn in-kernel, ATen mapping) or DDIM update to
// the fp32 NC[D]HW sampler state in place, so the predicted noise never goes to HBM.
// One CTA = a tile of (TD x TH x TW) output positions; the Y rows of the tile plus a one-voxel halo
// are staged in shared memory (odd word pitch: conflict-free), one thread per output position.
// ------------------------------------------------------------------------------------------
struct FusedStepArgs {
  const __nv_bfloat16* y;
  const float* bias;
  float* x;
  __nv_bfloat16* eps_out;   // optional bf16 [positions][ldo]
  const uint64_t* rng;
  const int64_t* t;
  const int64_t* t_prev;    // DDIM
  const float *betas, *sqrt_1mac, *sqrt_recip_alphas, *post_var, *alphas_cumprod;
  int samples, D, H, W, ndim, cout, ldy, ldo, mode;  // mode 0: DDPM, 1: DDIM, 2: eps only
  int TD, TH, TW, pitch_w;  // tile and shared-memory row pitch in 32-bit words
  long long Tt;             // ATen launch geometry of a randn over the whole state
};

template <int CH>   // 16-byte chunks per Y row that hold the tap products (0: run-time value)
__global__ void __launch_bounds__(256)
tap_gather_step_kernel(const FusedStepArgs a) {
  extern __shared__ uint32_t ysm[];
  const int hd = a.ndim == 3 ? 1 : 0;             // halo along depth only for 3-D filters
  const int RD = a.TD + 2 * hd, RH = a.TH + 2, RW = a.TW + 2;
  const int rows = RD * RH * RW;
  const int taps = a.ndim == 3 ? 27 : 9;
  const int row_elems = taps * a.cout;            // bf16 per Y row that matter
  const int row_words = (row_elems + 1) >> 1;
  const int chunks = CH > 0 ? CH : (row_words + 3) >> 2;
  // tile origin
  const int tiles_w = (a.W + a.TW - 1) / a.TW, tiles_h = (a.H + a.TH - 1) / a.TH;
  const int tiles_d = (a.D + a.TD - 1) / a.TD;
  int tix = (int)blockIdx.x;
  const int tw0 = (tix % tiles_w) * a.TW;
  tix /= tiles_w;
  const int th0 = (tix % tiles_h) * a.TH;
  tix /= tiles_h;
  const int td0 = (tix % tiles_d) * a.TD;
  const int n = tix / tiles_d;
  const long long spatial = (long long)a.D * a.H * a.W;
  const uint4* yn = reinterpret_cast<const uint4*>(a.y + (size_t)n * spatial * a.ldy);
  const int ldq = a.ldy >> 3;                     // Y row pitch in 16-byte units (ldy % 8 == 0)

  // ---- stage the halo tile -----------------------------------------------------------------------
  // (1) per halo row: its offset in Y (16-byte units) or -1 outside the tensor
  int* row_off = reinterpret_cast<int*>(ysm + (size_t)rows * a.pitch_w);
  for (int r = threadIdx.x; r < rows; r += 256) {
    int rr = r;
    const int w = tw0 + rr % RW - 1;
    rr /= RW;
    const int h = th0 + rr % RH - 1;
    const int d = td0 + rr / RH - hd;
    const bool ok = (unsigned)w < (unsigned)a.W && (unsigned)h < (unsigned)a.H && (unsigned)d < (unsigned)a.D;
    row_off[r] = ok ? (int)((((long long)d * a.H + h) * a.W + w) * ldq) : -1;
  }
  __syncthreads();
  // (2) 16-byte loads, four in flight per thread, scattered into rows of odd word pitch
  const int total = rows * chunks;
  for (int base = 0; base < total; base += 256 * 4) {
    uint4 v[4];
    int dst[4], jw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * 256 + (int)threadIdx.x;
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      dst[u] = -1;
      jw[u] = 0;
      if (i < total) {
        const int r = i / chunks;
        const int j = i - r * chunks;
        const int off = row_off[r];
        if (off >= 0) v[u] = __ldg(yn + off + j);
        dst[u] = r * a.pitch_w + 4 * j;
        jw[u] = 4 * j;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (dst[u] >= 0) {
        const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (jw[u] + k < row_words) ysm[dst[u] + k] = w4[k];
      }
    }
  }
  __syncthreads();

  // ---- one output position per thread -------------------------------------------------------------
  int q = threadIdx.x;
  const int lw = q % a.TW;
  q /= a.TW;
  const int lh = q % a.TH;
  const int ld_ = q / a.TH;
  const int w0 = tw0 + lw, h0 = th0 + lh, d0 = td0 + ld_;
  if (ld_ >= a.TD || w0 >= a.W || h0 >= a.H || d0 >= a.D) return;
  float acc[4];
#pragma unroll
  for (int co = 0; co < 4; ++co) acc[co] = (co < a.cout && a.bias != nullptr) ? __ldg(a.bias + co) : 0.f;
  const __nv_bfloat16* ys = reinterpret_cast<const __nv_bfloat16*>(ysm);
  int tap = 0;
  const int kd_n = a.ndim == 3 ? 3 : 1;
  for (int kd = 0; kd < kd_n; ++kd)
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw, ++tap) {
        const int r = ((ld_ + kd) * RH + lh + kh) * RW + lw + kw;
        const __nv_bfloat16* row = ys + (size_t)r * a.pitch_w * 2 + tap * a.cout;
#pragma unroll
        for (int co = 0; co < 4; ++co)
          if (co < a.cout) acc[co] += __bfloat162float(row[co]);   // halo rows outside the tensor are 0
      }
  const long long s = ((long long)d0 * a.H + h0) * a.W + w0;
  const long long per_sample = (long long)a.cout * spatial;
  // per-sample coefficients
  float q_c = 0.f, c1 = 0.f, wn = 0.f, sqrt_a_t = 0.f, s1m_t = 0.f, denom = 1.f, sqrt_a_p = 0.f, s1m_p = 0.f;
  if (a.mode == 0) {
    StepTables T{a.t, a.betas, a.sqrt_1mac, a.sqrt_recip_alphas, a.post_var};
    const StepCoef c = step_coef(T, n);
    q_c = c.q;
    c1 = c.c1;
    wn = c.w;
  } else if (a.mode == 1) {
    const float a_t = __ldg(a.alphas_cumprod + a.t[n]);
    const float a_p = __ldg(a.alphas_cumprod + a.t_prev[n]);
    sqrt_a_t = __fsqrt_rn(a_t);
    s1m_t = __fsqrt_rn(__fsub_rn(1.0f, a_t));
    denom = fmaxf(sqrt_a_t, 1e-8f);
    sqrt_a_p = __fsqrt_rn(a_p);
    s1m_p = __fsqrt_rn(__fsub_rn(1.0f, a_p));
  }
  uint64_t seed = 0, ctr0 = 0;
  if (a.mode == 0) {
    seed = a.rng[0];
    ctr0 = a.rng[1] >> 2;
  }
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    if (co >= a.cout) break;
    const __nv_bfloat16 eb = __float2bfloat16_rn(acc[co]);
    if (a.eps_out != nullptr) a.eps_out[((size_t)n * spatial + s) * a.ldo + co] = eb;
    if (a.mode == 2) continue;
    const float e = __bfloat162float(eb);
    const long long idx = (long long)n * per_sample + (long long)co * spatial + s;   // NC[D]HW element
    const float xv = a.x[idx];
    if (a.mode == 0) {
      // z of element idx under ATen's mapping: call it, component ii of subsequence j
      const long long it = idx / (4 * a.Tt);
      const long long rem = idx - it * 4 * a.Tt;
      const int ii = (int)(rem / a.Tt);
      const long long j = rem - (long long)ii * a.Tt;
      const uint64_t cc = ctr0 + (uint64_t)it;
      const uint4 rr = philox4x32_10(
          make_uint4((uint32_t)cc, (uint32_t)(cc >> 32), (uint32_t)j, (uint32_t)((uint64_t)j >> 32)),
          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      const float2 bm = (ii < 2) ? box_muller(rr.x, rr.y) : box_muller(rr.z, rr.w);
      const float z = (ii & 1) ? bm.y : bm.x;
      a.x[idx] = __fadd_rn(__fmul_rn(c1, __fsub_rn(xv, __fmul_rn(q_c, e))), __fmul_rn(wn, z));
    } else {
      const float x0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(s1m_t, e)), denom);
      a.x[idx] = __fadd_rn(__fmul_rn(sqrt_a_p, x0), __fmul_rn(s1m_p, e));
    }
  }
}

// one thread: t -= 1 for every sample (and t_prev), generator offset += rng_inc -- the bookkeeping
// between two replays of the captured reverse step
__global__ void step_advance_kernel(int64_t* t, int64_t* t_prev, int n, int64_t delta,
                                    uint64_t* rng, uint64_t rng_inc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    t[i] += delta;
    if (t_prev != nullptr) {  // strided DDIM: the last step lands on t_prev = 0
      const int64_t v = t_prev[i] + delta;
      t_prev[i] = v < 0 ? 0 : v;
    }
  }
  if (i == 0 && rng != nullptr) rng[1] += rng_inc;
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

__global__ void rng_seed_kernel(uint64_t* rng, uint64_t seed, uint64_t offset) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    rng[0] = seed;
    rng[1] = offset;
  }
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

static inline unsigned my_blocks(const AtenGrid& g) { return (unsigned)((g.Tt / 4 + 255) / 256); }

static int check_rng_args(const char* who, long long numel) {
  if (numel < 1) return set_error(-2, who);
  return device_props();
}

extern "C" int mri_randn_offset_increment(int64_t numel, uint64_t* increment_out) {
  if (numel < 1 || increment_out == nullptr)
    return set_error(-2, "mri_randn_offset_increment: empty tensor / null output");
  int rc = device_props();
  if (rc) return rc;
  *increment_out = 4ull * (uint64_t)aten_grid(numel).n_iter;
  return 0;
}

extern "C" int mri_randn(float* out, int64_t numel, const uint64_t* rng, void* stream) {
  int rc = check_rng_args("mri_randn: empty tensor", numel);
  if (rc) return rc;
  if (rng == nullptr) return set_error(-2, "mri_randn: null generator state");
  const AtenGrid g = aten_grid(numel);
  randn_kernel<true><<<my_blocks(g), 256, 0, (cudaStream_t)stream>>>(out, numel, g.Tt, g.n_iter, rng);
  return check_launch("randn_kernel");
}

static int q_sample_impl(const float* x0, const float* noise, const uint64_t* rng, const int64_t* t,
                         const float* sqrt_ac, const float* sqrt_1mac, float* out, float* noise_out,
                         int samples, int64_t per_sample, void* stream) {
  if (samples < 1 || per_sample < 1) return set_error(-2, "mri_q_sample: empty input");
  const long long numel = (long long)samples * per_sample;
  int rc = device_props();
  if (rc) return rc;
  const AtenGrid g = aten_grid(numel);
  if (rng != nullptr)
    q_sample_kernel<true><<<my_blocks(g), 256, 0, (cudaStream_t)stream>>>(
        x0, nullptr, rng, t, sqrt_ac, sqrt_1mac, out, noise_out, per_sample, numel, g.Tt, g.n_iter);
  else
    q_sample_kernel<false><<<my_blocks(g), 256, 0, (cudaStream_t)stream>>>(
        x0, noise, nullptr, t, sqrt_ac, sqrt_1mac, out, noise_out, per_sample, numel, g.Tt, g.n_iter);
  return check_launch("q_sample_kernel");
}

extern "C" int mri_q_sample(const float* x0, const float* noise, const int64_t* t,
                            const float* sqrt_ac, const float* sqrt_1mac, float* out, int samples,
                            int64_t per_sample, void* stream) {
  if (noise == nullptr) return set_error(-2, "mri_q_sample: null noise");
  return q_sample_impl(x0, noise, nullptr, t, sqrt_ac, sqrt_1mac, out, nullptr, samples, per_sample,
                       stream);
}

extern "C" int mri_q_sample_rng(const float* x0, const uint64_t* rng, const int64_t* t,
                                const float* sqrt_ac, const float* sqrt_1mac, float* out,
                                float* noise_out, int samples, int64_t per_sample, void* stream) {
  if (rng == nullptr) return set_error(-2, "mri_q_sample_rng: null generator state");
  return q_sample_impl(x0, nullptr, rng, t, sqrt_ac, sqrt_1mac, out, noise_out, samples, per_sample,
                       stream);
}

static int ddpm_step_impl(const float* x, const void* eps, int ldc, int channels, const float* noise,
                          const uint64_t* rng, const int64_t* t, const float* betas,
                          const float* sqrt_1mac, const float* sqrt_recip_alphas,
                          const float* post_var, float* out, int samples, int64_t per_sample,
                          void* stream) {
  if (samples < 1 || per_sample < 1) return set_error(-2, "mri_ddpm_step: empty input");
  if (ldc != 0 && (channels < 1 || per_sample % channels != 0))
    return set_error(-2, "mri_ddpm_step: per_sample must be channels * spatial");
  const long long numel = (long long)samples * per_sample;
  int rc = device_props();
  if (rc) return rc;
  const AtenGrid g = aten_grid(numel);
  StepTables T{t, betas, sqrt_1mac, sqrt_recip_alphas, post_var};
  if (rng != nullptr)
    ddpm_step_kernel<true><<<my_blocks(g), 256, 0, (cudaStream_t)stream>>>(
        x, eps, ldc, channels, nullptr, rng, T, out, per_sample, numel, g.Tt, g.n_iter);
  else
    ddpm_step_kernel<false><<<my_blocks(g), 256, 0, (cudaStream_t)stream>>>(
        x, eps, ldc, channels, noise, nullptr, T, out, per_sample, numel, g.Tt, g.n_iter);
  return check_launch("ddpm_step_kernel");
}

extern "C" int mri_ddpm_step(const float* x, const void* eps, int eps_nhwc_ldc, int channels,
                             const float* noise, const int64_t* t, const float* betas,
                             const float* sqrt_1mac, const float* sqrt_recip_alphas,
                             const float* post_var, float* out, int samples, int64_t per_sample,
                             void* stream) {
  if (noise == nullptr) return set_error(-2, "mri_ddpm_step: null noise");
  return ddpm_step_impl(x, eps, eps_nhwc_ldc, channels, noise, nullptr, t, betas, sqrt_1mac,
                        sqrt_recip_alphas, post_var, out, samples, per_sample, stream);
}

extern "C" int mri_ddpm_step_rng(const float* x, const void* eps, int eps_nhwc_ldc, int channels,
                                 const uint64_t* rng, const int64_t* t, const float* betas,
                                 const float* sqrt_1mac, const float* sqrt_recip_alphas,
                                 const float* post_var, float* out, int samples, int64_t per_sample,
                                 void* stream) {
  if (rng == nullptr) return set_error(-2, "mri_ddpm_step_rng: null generator state");
  return ddpm_step_impl(x, eps, eps_nhwc_ldc, channels, nullptr, rng, t, betas, sqrt_1mac,
                        sqrt_recip_alphas, post_var, out, samples, per_sample, stream);
}

extern "C" int mri_ddim_step(const float* x, const void* eps, int eps_nhwc_ldc, int channels,
                             const int64_t* t, const int64_t* t_prev, const float* alphas_cumprod,
                             float* out, int samples, int64_t per_sample, void* stream) {
  if (samples < 1 || per_sample < 1) return set_error(-2, "mri_ddim_step: empty input");
  if (eps_nhwc_ldc != 0 && (channels < 1 || per_sample % channels != 0))
    return set_error(-2, "mri_ddim_step: per_sample must be channels * spatial");
  const long long numel = (long long)samples * per_sample;
  int rc = device_props();
  if (rc) return rc;
  const AtenGrid g = aten_grid(numel);
  ddim_step_kernel<<<my_blocks(g), 256, 0, (cudaStream_t)stream>>>(
      x, eps, eps_nhwc_ldc, channels, t, t_prev, alphas_cumprod, out, per_sample, numel, g.Tt,
      g.n_iter);
  return check_launch("ddim_step_kernel");
}

extern "C" int mri_rng_seed(uint64_t* rng, uint64_t seed, uint64_t offset, void* stream) {
  if (rng == nullptr) return set_error(-2, "mri_rng_seed: null state");
  if (offset % 4 != 0) return set_error(-2, "mri_rng_seed: philox offset must be a multiple of 4");
  rng_seed_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(rng, seed, offset);
  return check_launch("rng_seed_kernel");
}

extern "C" int mri_step_advance(int64_t* t, int64_t* t_prev, int n, int64_t delta, uint64_t* rng,
                                uint64_t rng_increment, void* stream) {
  if (n < 1) return 0;
  step_advance_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t, t_prev, n, delta, rng,
                                                                        rng_increment);
  return check_launch("step_advance_kernel");
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

extern "C" int mri_tap_gather_step(const void* y, const float* bias, int samples, int D, int H, int W,
                                   int ndim, int cout, int ldy, float* x, void* eps_out, int ldo,
                                   int mode, const uint64_t* rng, const int64_t* t,
                                   const int64_t* t_prev, const float* betas, const float* sqrt_1mac,
                                   const float* sqrt_recip_alphas, const float* post_var,
                                   const float* alphas_cumprod, void* stream) {
  if (ndim != 2 && ndim != 3) return set_error(-2, "mri_tap_gather_step: ndim must be 2 or 3");
  if (cout < 1 || cout > 4) return set_error(-2, "mri_tap_gather_step: 1..4 output channels");
  const int taps = ndim == 3 ? 27 : 9;
  if (ldy % 2 != 0 || taps * cout > ldy) return set_error(-2, "mri_tap_gather_step: bad Y row layout");
  if (samples < 1 || D < 1 || H < 1 || W < 1) return set_error(-2, "mri_tap_gather_step: empty input");
  if (ndim == 2 && D != 1) return set_error(-2, "mri_tap_gather_step: 2-D problems have D = 1");
  if (mode < 0 || mode > 2) return set_error(-2, "mri_tap_gather_step: mode 0 (DDPM) / 1 (DDIM) / 2 (eps only)");
  if (mode == 0 && (rng == nullptr || t == nullptr || betas == nullptr || sqrt_1mac == nullptr ||
                    sqrt_recip_alphas == nullptr || post_var == nullptr || x == nullptr))
    return set_error(-2, "mri_tap_gather_step: DDPM mode needs x, rng, t and the four schedule tables");
  if (mode == 1 && (t == nullptr || t_prev == nullptr || alphas_cumprod == nullptr || x == nullptr))
    return set_error(-2, "mri_tap_gather_step: DDIM mode needs x, t, t_prev and alphas_cumprod");
  if (mode == 2 && eps_out == nullptr) return set_error(-2, "mri_tap_gather_step: eps-only mode needs eps_out");
  if (eps_out != nullptr && ldo < cout) return set_error(-2, "mri_tap_gather_step: ldo < cout");
  int rc = device_props();
  if (rc) return rc;
  FusedStepArgs a;
  a.y = reinterpret_cast<const __nv_bfloat16*>(y);
  a.bias = bias;
  a.x = x;
  a.eps_out = reinterpret_cast<__nv_bfloat16*>(eps_out);
  a.rng = rng;
  a.t = t;
  a.t_prev = t_prev;
  a.betas = betas;
  a.sqrt_1mac = sqrt_1mac;
  a.sqrt_recip_alphas = sqrt_recip_alphas;
  a.post_var = post_var;
  a.alphas_cumprod = alphas_cumprod;
  a.samples = samples; a.D = D; a.H = H; a.W = W; a.ndim = ndim; a.cout = cout; a.ldy = ldy; a.ldo = ldo;
  a.mode = mode;
  a.TD = ndim == 3 ? 4 : 1;
  a.TH = ndim == 3 ? 8 : 16;
  a.TW = ndim == 3 ? 8 : 16;
  const int row_words = (taps * cout + 1) / 2;
  a.pitch_w = row_words | 1;                         // odd pitch: consecutive rows hit different banks
  const long long numel = (long long)samples * cout * D * H * W;
  a.Tt = aten_grid(numel).Tt;
  const int rows = (a.TD + (ndim == 3 ? 2 : 0)) * (a.TH + 2) * (a.TW + 2);
  const int smem = rows * a.pitch_w * 4 + rows * 4;   // staged rows + the row-offset table
  if (ldy % 8 != 0) return set_error(-2, "mri_tap_gather_step: ldy must be a multiple of 8");
  if ((long long)D * H * W * (ldy / 8) > 0x7fffffffLL)
    return set_error(-2, "mri_tap_gather_step: sample too large for 32-bit row offsets");
  const long long tiles = (long long)samples * ((D + a.TD - 1) / a.TD) * ((H + a.TH - 1) / a.TH) *
                          ((W + a.TW - 1) / a.TW);
  if (tiles > 0x7fffffffLL) return set_error(-2, "mri_tap_gather_step: too many tiles");
  const int chunks = (row_words + 3) / 4;
  void (*kern)(const FusedStepArgs) = chunks == 11 ? tap_gather_step_kernel<11>
                                      : chunks == 5 ? tap_gather_step_kernel<5>
                                      : chunks == 2 ? tap_gather_step_kernel<2>
                                                    : tap_gather_step_kernel<0>;
  static int configured[4] = {0, 0, 0, 0};   // per instantiation (not re-done inside a graph capture)
  const int slot = chunks == 11 ? 0 : chunks == 5 ? 1 : chunks == 2 ? 2 : 3;
  if (smem > configured[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(tap_gather_step_kernel)");
    configured[slot] = smem;
  }
  kern<<<(unsigned)tiles, 256, smem, (cudaStream_t)stream>>>(a);
  return check_launch("tap_gather_step_kernel");
}
