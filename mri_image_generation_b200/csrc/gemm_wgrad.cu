// Weight-gradient GEMM on tcgen05:   dW[co, kb*64 + c] += sum_m dY[m, co] * A_kb[m, c]
//
// It re-uses the forward plan's k-table and activation TMA maps: for every k-table entry (one
// filter tap x 64 input channels) the matching 64-wide slab of dW is the product of the dY tile
// transposed and the same shifted activation box the forward pass multiplied.  The reduction
// runs over the output positions m, i.e. over the ROWS of both TMA tiles, so both operands are
// "MN-major" UMMA operands -- exactly what the 128-byte-swizzled TMA tiles already are when
// read along the other axis.  No transposition pass exists anywhere.
//
// One CTA owns (class, 128-wide co block, group of `group` <= 4 k-table entries) and a strided
// share (`splits`) of the M tiles.  The `group` activation boxes of a stage sit back to back in
// shared memory and form ONE MN-major B operand of N = 64 * group columns, so every 16 positions
// cost a single tcgen05.mma of shape 128 x (64*group) x 16 (the M = 128 weight-gradient operand
// is read from shared memory once per instruction, so wide N halves the operand traffic per
// FLOP; tools/umma_rate_probe.cu: the pipe itself runs N = 128 and N = 256 at full rate).
// Partial sums are added to the fp32 dW matrix with
// vectorised reductions (red.global.add.v4.f32).  Replaces autograd's conv weight gradients for every
// nn.Conv*/ConvTranspose* of the UNets (reference sites: include/mri_b200.h, MriGemmArgs) and,
// with per-class maps, the dV / dK products of the attention backward.
//
// xgroup mode (stride-1 3x3(x3) convolutions whose forward plan shares activation tiles, boxes
// of 8 x 16 positions): the k-table comes in triples (kw = 0, 1, 2 of one (kd, kh, channel slab)).
// One box loaded 10 positions wide serves all three taps: the three 64-channel blocks of the
// MN-major B operand are the SAME rows shifted by one position each, i.e. the descriptor's
// leading-dimension offset is 128 B (one row) and its stride offset 1280 B (one 10-position
// line) -- tools/umma_mn_probe.cu verifies that tcgen05 accepts this.  A CTA owns two triples:
// per 128 positions it loads 32 KB of dY + 2 x 20 KB of activations for SIX taps (12 KB per tap
// instead of 24 KB) and issues two 128 x 192 x 16 MMAs per 16 positions.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/mri_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace mri {

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 6;
constexpr int kTileBytes = 128 * 128;  // one [128 rows x 64 ch] bf16 box
constexpr int kWgMaxGroup = 4;
constexpr int kWgXTileBytes = 160 * 128;  // xgroup: one box loaded 10 positions wide (x - 1 .. x + 8) x 16 lines
constexpr int kWgMaxTriples = 2;

// MN-major operand, 128B swizzle: 64 MN elements contiguous (128 B) per K row, 8 K rows per
// 1024-byte atom; lbo = distance between 64-element MN blocks, sbo = distance between atoms.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                       uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ MriWgradArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kWgMaxStages + 1];
  __shared__ uint32_t tmem_holder;

  const int warp = uniform((int)(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  const int G = p.group;
  const bool xg = p.xgroup != 0;                       // G = 3 * triples (3 or 6)
  const int stage_bytes = xg ? 2 * kTileBytes + kWgMaxTriples * kWgXTileBytes : (2 + G) * kTileBytes;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kWgMaxStages + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * kWgMaxStages);

  // ---- work decode: blockIdx.x = ((cls * co_blocks + cb) * kb_groups + kg) * splits + split
  const int kb_groups = (p.n_kb + G - 1) / G;
  int w = blockIdx.x;
  const int split = w % p.splits;
  w /= p.splits;
  const int kg = w % kb_groups;
  w /= kb_groups;
  const int cb = w % p.co_blocks;
  const int cls = w / p.co_blocks;
  const int kb0 = kg * G;
  const int gact = (p.n_kb - kb0) < G ? (p.n_kb - kb0) : G;
  const int co0 = cb * 128;
  const int rows_in_box = p.box[0] * p.box[1] * p.box[2] * p.box[3];
  const int total_mt = p.tiles[0] * p.tiles[1] * p.tiles[2] * p.tiles[3];
  const int my_tiles = split < total_mt ? (total_mt - split + p.splits - 1) / p.splits : 0;

  uint32_t tmem_cols = 64;
  while ((int)tmem_cols < 64 * G) tmem_cols <<= 1;  // 6 entries -> 384 columns -> 512
  const int ntri = gact / 3;                          // xgroup: triples this CTA owns (1 or 2)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_holder), tmem_cols);
    tmem_relinquish();
  }
  // rows beyond the box are never written by TMA: they must read as zero (they are summed)
  if (rows_in_box < 128) {
    const int tiles_total = S * (2 + G);
    const int tail_bytes = (128 - rows_in_box) * 128;
    for (int t = 0; t < tiles_total; ++t) {
      const uint32_t base = smem_base + t * kTileBytes + rows_in_box * 128;
      for (int o = threadIdx.x * 16; o < tail_bytes; o += kWgThreads * 16)
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(base + o), "r"(0u) : "memory");
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;

  const CUtensorMap* a_maps = reinterpret_cast<const CUtensorMap*>(p.a_maps);
  const CUtensorMap* dy_map = reinterpret_cast<const CUtensorMap*>(p.dy_maps) + cls;

  if (warp == 0) {
    // one elected thread owns the producer loop (see gemm_tc.cu / tools/feed_probe.cu)
    if (my_tiles > 0 && elect_one_sync()) {
      // k-table entries of this group (registers; every lane loads the same values)
      int em[kWgMaxGroup], ec[kWgMaxGroup], eo[kWgMaxGroup][4];
      const int4* kt = reinterpret_cast<const int4*>(p.ktable) + ((size_t)cls * p.n_kb + kb0) * 2;
      const int n_src = xg ? ntri : gact;   // xgroup: one load per triple, at its first entry (kw = 0, o1 = -1)
      const int e_step = xg ? 3 : 1;
#pragma unroll
      for (int g = 0; g < kWgMaxGroup; ++g) {
        if (g < n_src) {
          const int4 e0 = __ldg(kt + 2 * g * e_step), e1 = __ldg(kt + 2 * g * e_step + 1);
          em[g] = e0.x;
          ec[g] = e0.y;
          eo[g][0] = e0.z;
          eo[g][1] = e0.w;
          eo[g][2] = e1.x;
          eo[g][3] = e1.y;
        }
      }
      const uint32_t tx = xg ? (uint32_t)(2 * kTileBytes + ntri * kWgXTileBytes)
                             : (uint32_t)rows_in_box * 128u * (2u + (uint32_t)gact);
      const uint32_t x_bytes = xg ? (uint32_t)kWgXTileBytes : (uint32_t)kTileBytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        int mt = split + it * p.splits;
        int org[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          org[i] = (mt % p.tiles[i]) * p.box[i];
          mt /= p.tiles[i];
        }
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t base = smem_base + stage * stage_bytes;
        {
          mbar_arrive_expect_tx(full_bar(stage), tx);
          tma_load_5d(base, dy_map, full_bar(stage), co0, org[0], org[1], org[2], org[3]);
          tma_load_5d(base + kTileBytes, dy_map, full_bar(stage), co0 + 64, org[0], org[1], org[2],
                      org[3]);
#pragma unroll
          for (int g = 0; g < kWgMaxGroup; ++g) {
            if (g < n_src)
              tma_load_5d(base + 2 * kTileBytes + g * x_bytes, a_maps + em[g], full_bar(stage), ec[g],
                          org[0] + eo[g][0], org[1] + eo[g][1], org[2] + eo[g][2], org[3] + eo[g][3]);
          }
        }
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (my_tiles > 0 && elect_one_sync()) {
      // D = f32, A = B = bf16, both MN-major (bits 15, 16), M = 128 (co), N = 64 * gact (channels)
      const uint32_t idesc =
          umma_idesc_bf16(128, xg ? 192u : (uint32_t)(64 * gact)) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t base = smem_base + stage * stage_bytes;
        const uint32_t b_addr = base + 2 * kTileBytes;
        {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // 16 rows (K) per MMA = 2 swizzle atoms = 2048 B
            const uint64_t a_desc = umma_desc_mn_sw128(base + ks * 2048, kTileBytes, 1024);
            if (xg) {
              // 16 positions = two 10-row lines of the tile; blocks of N = the taps, one row apart
              for (int t = 0; t < ntri; ++t) {
                const uint64_t b_desc =
                    umma_desc_mn_sw128(b_addr + t * kWgXTileBytes + ks * 2560, 128, 1280);
                umma_bf16(tmem_base + (uint32_t)(t * 192), a_desc, b_desc, idesc, (it | ks) != 0 ? 1u : 0u);
              }
            } else {
              const uint64_t b_desc = umma_desc_mn_sw128(b_addr + ks * 2048, kTileBytes, 1024);
              umma_bf16(tmem_base, a_desc, b_desc, idesc, (it | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));
          if (it == my_tiles - 1) umma_commit(tmem_full_bar);
        }
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (my_tiles > 0) {
    // ---- epilogue: TMEM (lane = co row) -> fp32 reductions into dW ----
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int4* kt = reinterpret_cast<const int4*>(p.ktable) + ((size_t)cls * p.n_kb + kb0) * 2;
    float* dw_row = p.dw + ((size_t)cls * p.dw_rows + co) * p.dw_ld;
    for (int g = 0; g < gact; ++g) {
      const int bk = __ldg(kt + 2 * g + 1).z;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 64 + c0), v);  // triples sit 192 apart: 3 * 64
        tmem_ld_wait();
        if (co < p.n_total) {
          float* dst = dw_row + bk + c0;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4 * u),
                         "f"(__uint_as_float(v[4 * u])), "f"(__uint_as_float(v[4 * u + 1])),
                         "f"(__uint_as_float(v[4 * u + 2])), "f"(__uint_as_float(v[4 * u + 3]))
                         : "memory");
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mri

using namespace mri;

extern "C" int mri_wgrad_launch(const MriWgradArgs* a, void* stream) {
  if (a == nullptr) return set_error(-1, "mri_wgrad_launch: null args");
  if (a->xgroup) {
    if (a->group != 3 && a->group != 6) return set_error(-2, "mri_wgrad_launch: xgroup needs group 3 or 6");
    if (a->n_kb % 3 != 0 || a->n_class != 1 || a->box[0] != 8 || a->box[1] != 16 ||
        a->box[0] * a->box[1] * a->box[2] * a->box[3] != 128)
      return set_error(-2, "mri_wgrad_launch: xgroup needs whole kw triples, one class and boxes of 8 x 16 positions");
  } else if (a->group < 1 || a->group > kWgMaxGroup) {
    return set_error(-2, "mri_wgrad_launch: group must be 1..4");
  }
  if (a->stages < 2 || a->stages > kWgMaxStages) return set_error(-2, "mri_wgrad_launch: stages must be 2..6");
  if (a->n_kb < 1 || a->n_class < 1 || a->splits < 1 || a->co_blocks < 1)
    return set_error(-2, "mri_wgrad_launch: empty problem");
  long rows = 1;
  for (int i = 0; i < 4; ++i) {
    if (a->box[i] < 1 || a->tiles[i] < 1) return set_error(-2, "mri_wgrad_launch: bad box/tiles");
    rows *= a->box[i];
  }
  if (rows > 128) return set_error(-2, "mri_wgrad_launch: box has more than 128 rows");
  if (a->dw_ld % 4 != 0) return set_error(-2, "mri_wgrad_launch: dw_ld must be a multiple of 4");
  const int smem = a->stages * (a->xgroup ? 2 * kTileBytes + kWgMaxTriples * kWgXTileBytes
                                            : (2 + a->group) * kTileBytes) + 1024;
  if (smem > 227 * 1024) return set_error(-2, "mri_wgrad_launch: shared memory over 227 KB");
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gemm_wgrad_kernel)");
    configured = smem;
  }
  const int kb_groups = (a->n_kb + a->group - 1) / a->group;
  const long grid = (long)a->n_class * a->co_blocks * kb_groups * a->splits;
  gemm_wgrad_kernel<<<(unsigned)grid, kWgThreads, smem, (cudaStream_t)stream>>>(*a);
  return check_launch("gemm_wgrad_kernel");
}
