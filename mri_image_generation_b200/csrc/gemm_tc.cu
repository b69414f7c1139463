// Implicit-GEMM convolution / GEMM on Blackwell tensor cores (tcgen05 + TMEM + TMA).
//
//   D[m, n] = sum_k A[m, k] * B[n, k]      A, B bf16 (K-major, 128B swizzle), D fp32 in TMEM
//
// One CTA = one 128 x block_n output tile.  Warp roles (192 threads):
//   warp 0     TMA producer: walks the k-table, one 64-channel A slab + B slab per stage
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-5  epilogue: tcgen05.ld -> bias / time-embedding / residual / GroupNorm partial
//              sums -> bf16|fp32 -> swizzled smem staging -> TMA store
// The 128 rows of a tile are a box of output positions; every filter tap is the same box
// shifted by (o1..o4), loaded by TMA with hardware zero fill outside the tensor (= padding).
// See include/mri_b200.h (MriGemmArgs) for the contract and the reference call sites.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <string.h>

#include "../../include/mri_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace mri {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 -> 128-byte rows = one swizzle span
constexpr int kUmmaK = 16;
constexpr int kAStageBytes = kBlockM * 128;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

__host__ __device__ inline int stage_bytes(int block_n) { return kAStageBytes + block_n * 128; }

__global__ void __launch_bounds__(kThreads, 2)
gemm_tc_kernel(const __grid_constant__ MriGemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 2];
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  const int block_n = p.block_n;
  const int sbytes = stage_bytes(block_n);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * kMaxStages);
  const uint32_t resid_bar = bar0 + 8u * (2 * kMaxStages + 1);

  // ---- tile decode -------------------------------------------------------------------
  int tile = blockIdx.x;
  const int nt = tile % p.n_tiles_n;
  tile /= p.n_tiles_n;
  int tix[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    tix[i] = tile % p.tiles[i];
    tile /= p.tiles[i];
  }
  const int cls = tile;
  int org[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) org[i] = tix[i] * p.box[i];
  const int n0 = nt * block_n;
  const int rows_in_box = p.box[0] * p.box[1] * p.box[2] * p.box[3];

  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < block_n) tmem_cols <<= 1;

  // ---- one-time setup ------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(resid_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_holder), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;

  const CUtensorMap* a_maps = reinterpret_cast<const CUtensorMap*>(p.a_maps);
  const CUtensorMap* b_map = reinterpret_cast<const CUtensorMap*>(p.b_map);

  if (warp == 0) {
    // ================================ TMA producer ==================================
    if (lane == 0) {
      auto sel = [&](int s) { return s == 0 ? 0 : (s == 1 ? cls : tix[s - 2]); };
      const int bz1 = sel(p.bz_sel[0]);
      const int bz2 = sel(p.bz_sel[1]);
      const uint32_t tx_bytes = (uint32_t)rows_in_box * 128u + (uint32_t)block_n * 128u;
      const int4* kt = reinterpret_cast<const int4*>(p.ktable) + (size_t)cls * p.n_kb * 2;
      int4 e0 = __ldg(kt), e1 = __ldg(kt + 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.n_kb; ++kb) {
        int4 f0 = e0, f1 = e1;
        if (kb + 1 < p.n_kb) {  // prefetch the next table entry ahead of the barrier wait
          f0 = __ldg(kt + 2 * (kb + 1));
          f1 = __ldg(kt + 2 * (kb + 1) + 1);
        }
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t a_dst = smem_base + stage * sbytes;
        const uint32_t b_dst = a_dst + kAStageBytes;
        mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
        tma_load_5d(a_dst, a_maps + e0.x, full_bar(stage), e0.y, org[0] + e0.z, org[1] + e0.w,
                    org[2] + e1.x, org[3] + e1.y);
        tma_load_4d(b_dst, b_map, full_bar(stage), e1.z, n0, bz1, bz2);
        e0 = f0;
        e1 = f1;
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ====================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kBlockM, (uint32_t)block_n);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.n_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * sbytes;
        const uint32_t b_addr = a_addr + kAStageBytes;
        const uint64_t a_desc = umma_desc_k_sw128(a_addr, 1024);
        const uint64_t b_desc = umma_desc_k_sw128(b_addr, 1024);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          // advance 16 bf16 = 32 bytes inside the 128B swizzle span: +2 in 16-byte units
          umma_bf16(tmem_base, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));  // frees this smem stage once the MMAs have read it
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(tmem_full_bar);  // accumulator complete -> epilogue
    }
  } else {
    // ================================ epilogue ======================================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;
    const int epi_tid = (warp - 2) * 32 + lane;
    const int esize = p.out_f32 ? 4 : 2;
    const int chunk_cols_full = p.out_f32 ? 32 : 64;
    const int chunk_cols = block_n < chunk_cols_full ? block_n : chunk_cols_full;
    const int chunk_bytes = chunk_cols * esize;
    const bool swz = (chunk_bytes == 128);
    const int n_chunks = block_n / chunk_cols;
    const uint32_t stag = smem_base;  // staging aliases the (drained) pipeline stages

    // row -> box-local coordinates
    int rl[4];
    {
      int rr = r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rl[i] = rr % p.box[i];
        rr /= p.box[i];
      }
    }
    bool valid = r < rows_in_box;
#pragma unroll
    for (int i = 0; i < 4; ++i) valid = valid && (org[i] + rl[i] < p.ext[i]);
    const int sd = p.sample_dim;
    const int sample = sd > 0 ? org[sd - 1] + rl[sd - 1] : 0;
    const bool uniform_sample = (sd == 0) || (p.box[sd - 1] == 1);

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();

    const CUtensorMap* o_map = reinterpret_cast<const CUtensorMap*>(p.o_maps) + cls;
    if (p.r_maps != nullptr) {
      if (epi_tid == 0) {
        const CUtensorMap* r_map = reinterpret_cast<const CUtensorMap*>(p.r_maps) + cls;
        mbar_arrive_expect_tx(resid_bar, (uint32_t)(rows_in_box * chunk_bytes * n_chunks));
        for (int c = 0; c < n_chunks; ++c)
          tma_load_5d(stag + c * (kBlockM * chunk_bytes), r_map, resid_bar, n0 + c * chunk_cols,
                      org[0], org[1], org[2], org[3]);
      }
      mbar_wait(resid_bar, 0);
    }

    const float* bias = p.bias;
    const float* rowbias = (p.rowbias != nullptr && valid)
                               ? p.rowbias + (size_t)sample * p.rowbias_ld
                               : nullptr;
    const float bias_m = (p.bias_m != nullptr && valid) ? __ldg(p.bias_m + org[0] + rl[0]) : 0.f;
    double* stats = p.stats;
    const int cpg = p.stats_cpg;
    int cur_g = -1;
    float s_sum = 0.f, s_sq = 0.f;

    auto flush_stats = [&]() {
      if (cur_g < 0) return;
      if (uniform_sample) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s_sum += __shfl_xor_sync(0xffffffffu, s_sum, o);
          s_sq += __shfl_xor_sync(0xffffffffu, s_sq, o);
        }
        if (lane == 0) {
          const int smp = sd > 0 ? org[sd - 1] : 0;
          double* dst = stats + ((size_t)smp * p.stats_ld + cur_g) * 2;
          atomicAdd(dst, (double)s_sum);
          atomicAdd(dst + 1, (double)s_sq);
        }
      } else if (valid) {
        double* dst = stats + ((size_t)sample * p.stats_ld + cur_g) * 2;
        atomicAdd(dst, (double)s_sum);
        atomicAdd(dst + 1, (double)s_sq);
      }
      s_sum = 0.f;
      s_sq = 0.f;
    };

    for (int c0 = 0; c0 < block_n; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + bias_m;
      const int nb = n0 + c0;
      if (nb < p.n_total) {  // n_total is a multiple of 8; 16-col groups may straddle the end
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (nb + h * 8 < p.n_total) {
            if (bias != nullptr) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + nb + h * 8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + nb + h * 8 + 4));
              f[h * 8 + 0] += b0.x; f[h * 8 + 1] += b0.y; f[h * 8 + 2] += b0.z; f[h * 8 + 3] += b0.w;
              f[h * 8 + 4] += b1.x; f[h * 8 + 5] += b1.y; f[h * 8 + 6] += b1.z; f[h * 8 + 7] += b1.w;
            }
            if (rowbias != nullptr) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(rowbias + nb + h * 8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(rowbias + nb + h * 8 + 4));
              f[h * 8 + 0] += b0.x; f[h * 8 + 1] += b0.y; f[h * 8 + 2] += b0.z; f[h * 8 + 3] += b0.w;
              f[h * 8 + 4] += b1.x; f[h * 8 + 5] += b1.y; f[h * 8 + 6] += b1.z; f[h * 8 + 7] += b1.w;
            }
          }
        }
      }
      // staging address of this thread's 16 columns
      const int chunk = c0 / chunk_cols;
      const int within = c0 - chunk * chunk_cols;
      const uint32_t row_base = stag + chunk * (kBlockM * chunk_bytes) + r * chunk_bytes;
      const int unit0 = (within * esize) >> 4;
      const int xr = swz ? (r & 7) : 0;

      if (p.r_maps != nullptr) {  // residual tile (bf16) sits in the staging buffer
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t addr = row_base + (uint32_t)(((unit0 + u) ^ xr) << 4);
          uint32_t w0, w1, w2, w3;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                       : "r"(addr));
          const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            f[u * 8 + 2 * j] += __uint_as_float(w[j] << 16);
            f[u * 8 + 2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
          }
        }
      }

      if (stats != nullptr) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n = nb + h * 8;
          if (n < p.n_total) {
            const int g = n / cpg;
            if (g != cur_g) {
              flush_stats();
              cur_g = g;
            }
            if (valid) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float x = f[h * 8 + i];
                s_sum += x;
                s_sq = fmaf(x, x, s_sq);
              }
            }
          }
        }
      }

      if (p.out_f32) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t addr = row_base + (uint32_t)(((unit0 + u) ^ xr) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(f[u * 4 + 0]),
                       "f"(f[u * 4 + 1]), "f"(f[u * 4 + 2]), "f"(f[u * 4 + 3])
                       : "memory");
        }
      } else {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[u * 8 + 2 * j], f[u * 8 + 2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&h2);
          }
          const uint32_t addr = row_base + (uint32_t)(((unit0 + u) ^ xr) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(w[0]), "r"(w[1]),
                       "r"(w[2]), "r"(w[3])
                       : "memory");
        }
      }
    }
    if (stats != nullptr) flush_stats();

    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (epi_tid == 0) {
      for (int c = 0; c < n_chunks; ++c) {
        if (n0 + c * chunk_cols < p.n_total)
          tma_store_5d(o_map, stag + c * (kBlockM * chunk_bytes), n0 + c * chunk_cols, org[0],
                       org[1], org[2], org[3]);
      }
      tma_store_commit();
      tma_store_wait_all();
    }
  }

  // ---- teardown ------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mri

// =============================== host side =============================================
using namespace mri;

extern "C" int mri_gemm_smem_bytes(int block_n, int stages) {
  return stages * stage_bytes(block_n) + 1024;
}

extern "C" int mri_gemm_occupancy(int block_n, int stages) {
  const int smem = mri_gemm_smem_bytes(block_n, stages);
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gemm_tc_kernel)");
  (void)cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
  int nb = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gemm_tc_kernel, kThreads, smem);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  return nb;
}

extern "C" int mri_gemm_launch(const MriGemmArgs* a, void* stream) {
  if (a == nullptr) return set_error(-1, "mri_gemm_launch: null args");
  const int bn = a->block_n;
  if (!(bn == 16 || bn == 32 || bn == 64 || bn == 128 || bn == 256))
    return set_error(-2, "mri_gemm_launch: block_n must be 16/32/64/128/256");
  if (a->stages < 2 || a->stages > kMaxStages)
    return set_error(-2, "mri_gemm_launch: stages must be in [2, 8]");
  if (a->n_kb < 1 || a->n_class < 1) return set_error(-2, "mri_gemm_launch: empty K loop");
  long rows = 1;
  long grid = (long)a->n_tiles_n * a->n_class;
  for (int i = 0; i < 4; ++i) {
    if (a->box[i] < 1 || a->tiles[i] < 1) return set_error(-2, "mri_gemm_launch: bad box/tiles");
    rows *= a->box[i];
    grid *= a->tiles[i];
  }
  if (rows > kBlockM) return set_error(-2, "mri_gemm_launch: box has more than 128 rows");
  if (a->n_total % 8 != 0) return set_error(-2, "mri_gemm_launch: n_total must be a multiple of 8");
  if (a->stats != nullptr && (a->stats_cpg < 8 || a->stats_cpg % 8 != 0))
    return set_error(-2, "mri_gemm_launch: statistics need channel groups in multiples of 8");
  if (a->r_maps != nullptr && a->out_f32)
    return set_error(-2, "mri_gemm_launch: residual input requires bf16 output");
  const int smem = mri_gemm_smem_bytes(bn, a->stages);
  // staging (aliases the stages) must fit
  const int stag_bytes = kBlockM * bn * (a->out_f32 ? 4 : 2);
  if (stag_bytes > a->stages * stage_bytes(bn))
    return set_error(-2, "mri_gemm_launch: staging does not fit the stage ring");
  if (smem > 227 * 1024) return set_error(-2, "mri_gemm_launch: shared memory over 227 KB");
  static int configured_smem = 0;
  if (smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gemm_tc_kernel)");
    // ask for the full shared-memory carveout so that two CTAs (2 x ~100 KB) can share an SM
    e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(carveout)");
    configured_smem = smem;
  }
  gemm_tc_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(*a);
  return check_launch("gemm_tc_kernel");
}

// ---- TMA descriptor encoding through the driver entry point (no link-time libcuda) -------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int mri_tmap_encode(void* out_map_host, uint64_t global_addr, int dtype, int rank,
                               const uint64_t* dims, const uint64_t* strides_bytes,
                               const uint32_t* box, int swizzle) {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || sym == nullptr || qres != cudaDriverEntryPointSuccess)
      return set_error(-3, "mri_tmap_encode: cuTensorMapEncodeTiled not available (no driver?)");
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  if (rank < 1 || rank > 5) return set_error(-2, "mri_tmap_encode: rank must be 1..5");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  const CUtensorMapDataType dt =
      dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = swizzle == 3   ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                               : CU_TENSOR_MAP_SWIZZLE_NONE;
  alignas(64) CUtensorMap tmp;
  CUresult r = fn(&tmp, dt, (cuuint32_t)rank, reinterpret_cast<void*>(global_addr), gdim, gstr, bx,
                  es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[256];
    snprintf(msg, sizeof msg,
             "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu %llu] "
             "box [%u %u %u %u %u]",
             (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
             (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
             (unsigned long long)(rank > 4 ? gdim[4] : 0), bx[0], rank > 1 ? bx[1] : 0,
             rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
    return set_error(-4, msg);
  }
  memcpy(out_map_host, &tmp, sizeof(CUtensorMap));
  return 0;
}
