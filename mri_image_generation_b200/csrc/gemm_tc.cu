// Implicit-GEMM convolution / GEMM on Blackwell tensor cores (tcgen05 + TMEM + TMA), revision D:
// persistent CTAs, stream-K work split, epilogue overlapped with the next tile's main loop,
// single-thread producer / MMA-issue loops.
//
//   D[m, n] = sum_k A[m, k] * B[n, k]      A, B bf16 (K-major, 128B swizzle), D fp32 in TMEM
//   m = output position (a box of up to 128 positions), n = output channel, k = (tap, channel)
//
// One CTA per SM (grid <= #SMs), each walking its share of the tiles (see "schedule" below):
// whole tiles, plus -- for the tiles that would leave the last wave under-filled -- an equal
// share of their K loops ("stream-K").
//
// Two tile shapes (the M = 128 operand is read from shared memory once per instruction, so N = 256
// halves that operand's traffic per FLOP; tools/umma_rate_probe.cu: the pipe itself runs N = 128
// and N = 256 instructions at full rate):
//   swap_ab = 1  (convolutions with Cout % 128 == 0): the WEIGHTS are the M = 128 operand and
//                two boxes of positions are the N = 256 operand; the accumulator holds D^T
//                (lane = channel, column = position) and the epilogue transposes on its way
//                to the channels-last output;
//   swap_ab = 0  (thin / odd shapes, fp32 outputs): positions are M, block_n channels are N.
// Warp roles (192 threads):
//   warp 0     TMA producer: walks the k-table, one stage = activation slab(s) + weight slab;
//              ONE elected thread owns the whole loop (tools/feed_probe.cu: a warp kept converged
//              around a per-step elect.sync costs ~130 cycles per k-step on the issue path)
//   warp 1     TMEM allocator + tcgen05.mma issuer, also a single elected thread (waits, MMAs,
//              commits).  TMEM holds two tile buffers of 256 columns, so tile i+1 accumulates
//              while tile i drains
//   warps 2-5  epilogue: tcgen05.ld -> (+ partial sums of the CTAs that share the tile) ->
//              bias / time-embedding / residual / GroupNorm partial sums -> bf16|fp32 ->
//              swizzled staging (2 x 16 KB) -> TMA store
// Split tiles: every CTA but the one holding a tile's first k-step writes its fp32 partial
// tile to a workspace and raises a flag; the head CTA (which reaches that tile LAST in its own
// range) adds the partials in fixed order and runs the fused epilogue -> deterministic.
// Every filter tap is the same box of positions shifted by (o1..o4), loaded by TMA with
// hardware zero fill outside the tensor (= the convolution's zero padding).
// See include/mri_b200.h (MriGemmArgs) for the contract and the reference call sites.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mri_b200.h"
#include "common.h"
#include "ptx.cuh"

// Fine-grained epilogue trace (tools/gemm_probe.py ft=1; compile with -DMRI_GEMM_FINE_TRACE): per-phase
// cycle sums of epilogue threads 0 and 96 in trace rows [gridDim + block] and [2 gridDim + block].
#ifdef MRI_GEMM_FINE_TRACE
#define FT_START() do { if (trace != nullptr) ft_t = clock64(); } while (0)
#define FT_ADD(k) do { if (trace != nullptr) { const long long n_ = clock64(); ft[k] += (unsigned long long)(n_ - ft_t); ft_t = n_; } } while (0)
#else
#define FT_START() do { } while (0)
#define FT_ADD(k) do { } while (0)
#endif

namespace mri {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 -> 128-byte rows = one swizzle span
constexpr int kUmmaK = 16;
constexpr int kSlabBytes = kBlockM * 128;        // 128 rows x 64 bf16
// Warp roles.  Default (MRI_EPI_WARPS = 4): 6 warps -- TMA producer, MMA issuer, four epilogue warps
// (a warp may only read the TMEM lane quadrant warp % 4).  MRI_EPI_WARPS = 8 builds the variant with
// 12 warps = 3 warpgroups: WG0 = {producer, MMA issuer, 2 idle warps}, WG1 + WG2 = 8 epilogue warps
// in PAIRS per quadrant that split the columns of every tile, registers re-balanced with setmaxnreg
// (the CTA launches at 168 per thread; 104 + 2 * 200 = 3 * 168 is what one sub-partition's three
// warps own).  Measured on B200 (profiles/README.md, round 2): the pairs hide each other's latencies
// only on partial-box tiles; everywhere else the variant LOSES 4 % on the cfg4 step, because the
// values shared by all roles end up in a stack frame and the single-thread producer / MMA loops
// re-load them every k-step.  What made the epilogue ~1.6x faster instead was doing less per box.
#ifndef MRI_EPI_WARPS
#define MRI_EPI_WARPS 4
#endif
constexpr int kEpiWarps = MRI_EPI_WARPS;
static_assert(kEpiWarps == 4 || kEpiWarps == 8, "one or two epilogue warps per TMEM lane quadrant");
constexpr int kSplit = kEpiWarps / 4;                   // warps sharing a quadrant split a tile's columns
constexpr int kEpiWarp0 = kEpiWarps == 8 ? 4 : 2;       // first epilogue warp
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;  // 384 / 192
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kRegsProducer = 104, kRegsEpilogue = 200;  // 104 + 2 * 200 = 3 * 168: what one sub-partition's three warps own
constexpr int kMaxStages = 8;
constexpr int kChunkBytes = kBlockM * 128;       // one staging buffer: 128 rows x 128 bytes
constexpr int kStagingBytes = 2 * kChunkBytes;
constexpr int kPartialLd = 256;                  // floats per row of a stream-K partial tile
constexpr int kMaxStatGroups = 32;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kXTileBytes = 2 * 160 * 128;       // xreuse 1: two boxes of (8+2) x 16 positions x 64 channels
constexpr int kXYTileBytes = 2 * 180 * 128;      // xreuse 2: two boxes of (8+2) x (16+2) positions
constexpr int kXStages = 3;                       // activation-tile ring (2 when a second staging set is in use)

__host__ __device__ inline int stage_bytes(int block_n, int swap_ab) {
  return swap_ab ? 3 * kSlabBytes : kSlabBytes + block_n * 128;
}

// One tile of work: 1 (normal) or 2 (swap_ab) boxes of positions x one block of channels.
struct Work {
  int cls, n0, nbox;
  int tix[2][4], org[2][4];
};

struct Geom {
  int boxes_per_class;   // boxes of positions per output class
  int groups_per_class;  // tiles along the position axis per class (box pairs when swap_ab)
  long long n_tiles;
  // reciprocals for fdivmod (tile index -> class / channel block / box coordinates): a generic
  // 32-bit division is ~25 instructions and decode_tile needs a dozen of them per tile
  float rcp_ntn, rcp_gpc, rcp_te[4];
  int te[4], od[4];      // box-grid extents / dims in enumeration order
};

// n / d and n % d for 0 <= n < 2^24, d >= 1 through one float multiply and a fix-up
__device__ __forceinline__ void fdivmod(int n, int d, float rcp, int& q, int& r) {
  q = __float2int_rz(__int2float_rn(n) * rcp);
  r = n - q * d;
  if (r < 0) {
    --q;
    r += d;
  } else if (r >= d) {
    ++q;
    r -= d;
  }
}

__device__ __forceinline__ Geom make_geom(const MriGemmArgs& p) {
  Geom g;
  g.boxes_per_class = p.tiles[0] * p.tiles[1] * p.tiles[2] * p.tiles[3];
  g.groups_per_class = p.swap_ab ? (g.boxes_per_class + 1) / 2 : g.boxes_per_class;
  g.n_tiles = (long long)p.n_tiles_n * p.n_class * g.groups_per_class;
  g.rcp_ntn = 1.0f / (float)p.n_tiles_n;
  g.rcp_gpc = 1.0f / (float)g.groups_per_class;
  // boxes advance along dim f first, then along x1..x4 in order: enumeration step k visits dim
  // od[k] = f, then the others ascending
  const int f = p.tile_fast_dim;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    g.od[k] = (f == 0) ? k : (k == 0 ? f : (k <= f ? k - 1 : k));
    g.te[k] = g.od[k] == 0 ? p.tiles[0] : (g.od[k] == 1 ? p.tiles[1] : (g.od[k] == 2 ? p.tiles[2] : p.tiles[3]));
    g.rcp_te[k] = 1.0f / (float)g.te[k];
  }
  return g;
}

// a[i] for a run-time i WITHOUT indexing the array: a dynamically indexed local array lives in
// local memory (a stack frame), and with 227 KB of shared memory per CTA the L1 that is left cannot
// hold 192 stack frames -- every access then costs an L2 round trip (measured: ~10 k cycles of
// epilogue per tile with NO work in it, profiles/r02b_epi_trace.txt).  Selects keep it in registers.
__device__ __forceinline__ int pick4(const int (&a)[4], int i) {
  return i == 0 ? a[0] : (i == 1 ? a[1] : (i == 2 ? a[2] : a[3]));
}

__device__ __forceinline__ void decode_tile(const MriGemmArgs& p, const Geom& g, int tile, Work& w) {
  int nt, pr;
  fdivmod(tile, p.n_tiles_n, g.rcp_ntn, pr, nt);
  int rem;
  fdivmod(pr, g.groups_per_class, g.rcp_gpc, w.cls, rem);
  w.n0 = nt * p.block_n;
  int b0 = rem;
  w.nbox = 1;
  if (p.swap_ab) {
    b0 = 2 * rem;
    w.nbox = (b0 + 1 < g.boxes_per_class) ? 2 : 1;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int b = b0 + h;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int qn;
      fdivmod(b, g.te[k], g.rcp_te[k], qn, v[k]);
      b = qn;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      w.tix[h][i] = g.od[0] == i ? v[0] : (g.od[1] == i ? v[1] : (g.od[2] == i ? v[2] : v[3]));
      w.org[h][i] = w.tix[h][i] * p.box[i];
    }
  }
}

// ---- schedule ------------------------------------------------------------------------------
// sched 0: CTA with range index r owns the whole tiles [n_tiles*r/G, n_tiles*(r+1)/G).
// sched 1: n_tiles = q*G + rem.  Phase A ("stream-K"): the K loops of the first `rem` tiles are
//          one linear space of rem*n_kb k-steps cut into GA equal ranges (GA <= G CTAs take
//          part), so a tile may be shared by neighbouring ranges.  Phase B: q whole tiles per
//          CTA.  The split tiles come FIRST so that the exchange of partial sums is hidden
//          behind the whole tiles that follow.
struct Sched {
  int n_kb, G, GA, q;
  long long rem_units;  // rem * n_kb
  long long n_tiles, rem;
};

__device__ __forceinline__ Sched make_sched(const MriGemmArgs& p, long long n_tiles, int G) {
  Sched s;
  s.n_kb = p.n_kb;
  s.G = G;
  s.n_tiles = n_tiles;
  if (p.sched == 0) {
    s.q = 0;
    s.rem = 0;
    s.rem_units = 0;
    s.GA = 0;
  } else {
    s.q = (int)(n_tiles / G);
    s.rem = n_tiles - (long long)s.q * G;
    s.rem_units = s.rem * p.n_kb;
    // a tile is shared by at most ~4 CTAs (its head CTA reads every other share back from L2)
    // and a share is at least 8 k-steps
    const int min_share = (p.n_kb + 3) / 4 > 8 ? (p.n_kb + 3) / 4 : 8;
    const long long ga = s.rem_units / min_share;
    s.GA = (int)(ga < 1 ? (s.rem > 0 ? 1 : 0) : (ga > G ? G : ga));
  }
  return s;
}

// phase-A unit range of range index r
__device__ __forceinline__ void range_a(const Sched& s, int r, long long& u0, long long& u1) {
  if (r >= s.GA) {
    u0 = u1 = s.rem_units;
    return;
  }
  u0 = s.rem_units * r / s.GA;
  u1 = s.rem_units * (r + 1) / s.GA;
}

struct SegIter {
  long long u, u1;  // phase A
  long long tb, tb1;  // phase B (whole tiles)
  int n_kb;
  __device__ __forceinline__ void init(const MriGemmArgs& p, const Sched& s, int r) {
    n_kb = s.n_kb;
    if (p.sched == 0) {
      u = u1 = 0;
      tb = s.n_tiles * r / s.G;
      tb1 = s.n_tiles * (r + 1) / s.G;
    } else {
      range_a(s, r, u, u1);
      tb = s.rem + (long long)r * s.q;
      tb1 = tb + s.q;
    }
  }
  // next segment: tile, first k-step, number of k-steps; false when the CTA's work is done
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& len, long long& seg_end) {
    if (u < u1) {
      tile = (int)(u / n_kb);
      kb0 = (int)(u - (long long)tile * n_kb);
      const long long left = u1 - u;
      len = (int)(left < (long long)(n_kb - kb0) ? left : (long long)(n_kb - kb0));
      u += len;
      seg_end = u;
      return true;
    }
    if (tb < tb1) {
      tile = (int)tb++;
      kb0 = 0;
      len = n_kb;
      seg_end = -1;
      return true;
    }
    return false;
  }
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const int32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Contributors of a split tile: the phase-A ranges after `my_r` that cover [pos, tile_end).
__device__ __forceinline__ int count_contributors(const Sched& s, int my_r, long long pos,
                                                  long long tile_end) {
  int n = 0;
  int rr = my_r + 1;
  while (pos < tile_end) {
    long long a, b;
    range_a(s, rr, a, b);
    pos = b < tile_end ? b : tile_end;
    ++rr;
    ++n;
  }
  return n;
}

__device__ __forceinline__ void wait_contributors(const MriGemmArgs& p, int first_cta, int n) {
  for (int j = 0; j < n; ++j) {
    const int32_t* f = p.sk_flags + (first_cta - j);
    const uint64_t t0 = globaltimer_ns();
    while (ld_acquire_gpu(f) == 0u) {
      if (globaltimer_ns() - t0 > 4000000000ull) {
        printf("mri_b200: stream-K partial wait timeout (block %d waits for %d)\n",
               (int)blockIdx.x, first_cta - j);
        __trap();
      }
    }
  }
}

// Partial tiles live in the workspace as [cta][column / 4][lane 0..127][4 floats]: the 32 lanes
// of a warp touch 512 contiguous bytes per access.
__device__ __forceinline__ float4* partial_ptr(const MriGemmArgs& p, int cta, int lane128, int col) {
  return reinterpret_cast<float4*>(p.sk_partials) +
         ((size_t)cta * (kPartialLd / 4) + (size_t)(col >> 2)) * kBlockM + lane128;
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ MriGemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 5 + kXStages];
  __shared__ uint32_t tmem_holder;
  __shared__ double s_stats[kMaxStatGroups * 2];
  __shared__ int s_pos_info[kBlockM];       // swap_ab: sample index, or -1 for an invalid position
  __shared__ uint32_t s_vmask[4];           // swap_ab: validity bits of the box's 128 positions
  __shared__ uint32_t s_winfo[kMaxStages];  // xreuse: per weight stage, (dx + 1) | next-needs-tile << 8

  const int warp = uniform((int)(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  const int block_n = p.block_n;
  const bool swap = p.swap_ab != 0;
  const int sbytes = stage_bytes(block_n, p.swap_ab);
  const uint32_t stag = smem_base + (uint32_t)(p.xreuse ? S * kSlabBytes + (p.xreuse == 2 ? 2 * kXYTileBytes : ((p.staging2 || p.stages > 4) ? 2 : kXStages) * kXTileBytes)
                                                        : S * sbytes);  // staging follows the ring(s)
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto tmem_full_bar = [&](int b) { return bar0 + 8u * (2 * kMaxStages + b); };
  auto tmem_empty_bar = [&](int b) { return bar0 + 8u * (2 * kMaxStages + 2 + b); };
  const uint32_t resid_bar = bar0 + 8u * (2 * kMaxStages + 4);
  auto xempty_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + 5 + s); };
  const bool xreuse = p.xreuse != 0;
  // xreuse 1: the three kw taps of a (kd, kh, slab) group share a 10-wide tile (ring of 3 tiles,
  // 2 beside a second staging set); xreuse 2: the nine (kh, kw) taps of a (kd, slab) group share
  // a 10 x 18 tile (ring of 2 tiles = 18 k-steps of lookahead)
  const bool xy = p.xreuse == 2;
  const int XS = (xy || p.staging2 || p.stages > 4) ? 2 : kXStages;  // weight ring of 6 leaves room for 2 tiles
  const uint32_t x_tile_bytes = xy ? (uint32_t)kXYTileBytes : (uint32_t)kXTileBytes;
  const int x_box_rows = xy ? 180 : 160;
  const uint32_t x_ring = smem_base + (uint32_t)(S * kSlabBytes);  // xreuse: after the weight ring

  const int n_kb = p.n_kb;
  const int G = (int)gridDim.x;
  const int my_r = G - 1 - (int)blockIdx.x;  // waiters (tile heads) wait on LOWER block indices
  // tile geometry and schedule live in shared memory, not in registers: they are read once per
  // tile (decode_tile, SegIter), and the single-thread producer / MMA loops run on a small
  // register budget after setmaxnreg -- values kept live across their k-loops would be spilled
  // to local memory and re-loaded every k-step
  __shared__ Geom s_geom;
  __shared__ Sched s_sch;
  if (threadIdx.x == 0) {
    s_geom = make_geom(p);
    s_sch = make_sched(p, s_geom.n_tiles, G);
  }
  const Geom& geom = s_geom;
  const Sched& sch = s_sch;
  const int rows_in_box = p.box[0] * p.box[1] * p.box[2] * p.box[3];
  const bool dual = !swap && block_n <= 128;  // two accumulators per tile (k-step parity)

  // ---- one-time setup ------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), kEpiThreads / 32);  // one arrive per epilogue warp
    }
    mbar_init(resid_bar, 1);
    for (int xs = 0; xs < kXStages; ++xs) mbar_init(xempty_bar(xs), 1);
    mbar_fence_init();
    tma_prefetch_desc(p.b_map);
    tma_prefetch_desc(p.a_maps);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_holder), 512);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * kMaxStatGroups) s_stats[threadIdx.x - 64] = 0.0;
  static_assert(2 * kMaxStatGroups <= 128, "statistics table is zeroed by the epilogue threads");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  // profiling: [0] kernel entry (globaltimer ns), [1] setup done, [2] first stage landed,
  // [3] last MMA issued, [4] last accumulator complete, [5] epilogue done (clock64 cycles),
  // [6] exit (ns); cycles spent waiting: [7] MMA thread on full barriers (operand starvation),
  // [8] MMA thread on tmem_empty (epilogue back-pressure), [9] producer on empty barriers
  // (ring full: healthy), [10] producer on the activation ring, [11] segments run by the CTA
  uint64_t* trace = p.trace != nullptr ? p.trace + (size_t)blockIdx.x * 16 : nullptr;
  if (trace != nullptr && threadIdx.x == 0) {
    trace[0] = globaltimer_ns();
    trace[1] = (uint64_t)clock64();
  }

  const CUtensorMap* a_maps = reinterpret_cast<const CUtensorMap*>(p.a_maps);
  const CUtensorMap* b_map = reinterpret_cast<const CUtensorMap*>(p.b_map);
  // barrier wait that charges its duration to a trace slot (profiling runs only)
  auto wait_t = [&](uint32_t bar, uint32_t parity, int slot) {
    if (trace == nullptr) {
      mbar_wait(bar, parity);
    } else {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      trace[slot] += (uint64_t)(clock64() - t0);
    }
  };

  if (warp == 0) {
    if (kEpiWarps == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer));
    // ================================ TMA producer ==================================
    // ONE elected thread owns the whole loop (barrier waits, table reads, TMA issue).  Measured
    // with tools/feed_probe.cu: keeping the warp converged around a per-step elect.sync costs
    // ~130 cycles per k-step on the issue path (644 -> 512 cycles per 128x256x64 k-step once both
    // the producer and the MMA loop are single-thread loops); the elect tells the compiler the
    // region has exactly one thread, so TMA operands still move through uniform registers.
    if (!elect_one_sync()) {
      // the other 31 lanes have nothing to do until the teardown barrier
    } else if (xreuse) {
      // k-table entries come in groups that share ONE activation tile: the group leader (entry[7]
      // = 1, or the first entry of a segment) loads both boxes 10 positions wide (x - 1 .. x + 8),
      // every entry loads its own 64-channel weight slab; the MMA reads the tap dx as the view that
      // starts dx + 1 rows into the tile (group stride 10 rows)
      int ws = 0, xs = 0;
      uint32_t wphase = 0, xphase = 0;
      SegIter it;
      it.init(p, sch, my_r);
      int tile, kb0, len;
      long long seg_end;
      while (it.next(tile, kb0, len, seg_end)) {
        Work t;
        decode_tile(p, geom, tile, t);
        const int bz1 = p.bz_sel[0] == 1 ? t.cls : 0;
        const int bz2 = p.bz_sel[1] == 1 ? t.cls : 0;
        const int4* kt = reinterpret_cast<const int4*>(p.ktable) + ((size_t)t.cls * n_kb + kb0) * 2;
        int4 e0 = __ldg(kt), e1 = __ldg(kt + 1);
        for (int i = 0; i < len; ++i) {
          int4 f0 = e0, f1 = e1;
          if (i + 1 < len) {
            f0 = __ldg(kt + 2 * (i + 1));
            f1 = __ldg(kt + 2 * (i + 1) + 1);
          }
          const int am = e0.x, c0 = e0.y, o2 = e0.w;
          const int o3 = e1.x, o4 = e1.y, bk = e1.z;
          const bool need_x = (e1.w != 0) || i == 0;
          const int o1 = e0.z;
          // the tile stays current until the next leader (or the end of the segment)
          const bool next_needs_x = (i + 1 < len) ? (f1.w != 0) : true;
          wait_t(empty_bar(ws), wphase ^ 1u, 9);
          if (need_x) wait_t(xempty_bar(xs), xphase ^ 1u, 10);
          const uint32_t w_dst = smem_base + ws * kSlabBytes;
          const uint32_t x_dst = x_ring + xs * x_tile_bytes;
          {
            const uint32_t tx =
                (uint32_t)block_n * 128u + (need_x ? (uint32_t)(x_box_rows * t.nbox) * 128u : 0u);
            // row offset of this tap's view inside the tile (read by the MMA thread)
            const uint32_t row_off = xy ? (uint32_t)((o2 + 1) * 10 + o1 + 1) : (uint32_t)(o1 + 1);
            s_winfo[ws] = row_off | (next_needs_x ? 256u : 0u);
            mbar_arrive_expect_tx(full_bar(ws), tx);
            if (need_x) {
              const int y0 = xy ? -1 : o2;  // the 18-line tile starts one line above the box
              tma_load_5d(x_dst, a_maps + am, full_bar(ws), c0, t.org[0][0] - 1, t.org[0][1] + y0,
                          t.org[0][2] + o3, t.org[0][3] + o4);
              if (t.nbox == 2)
                tma_load_5d(x_dst + x_tile_bytes / 2, a_maps + am, full_bar(ws), c0, t.org[1][0] - 1,
                            t.org[1][1] + y0, t.org[1][2] + o3, t.org[1][3] + o4);
            }
            tma_load_4d(w_dst, b_map, full_bar(ws), bk, t.n0, bz1, bz2);
          }
          e0 = f0;
          e1 = f1;
          if (++ws == S) {
            ws = 0;
            wphase ^= 1u;
          }
          if (next_needs_x) {
            if (++xs == XS) {
              xs = 0;
              xphase ^= 1u;
            }
          }
        }
      }
    } else {
      int stage = 0;
      uint32_t phase = 0;
      SegIter it;
      it.init(p, sch, my_r);
      int tile, kb0, len;
      long long seg_end;
      while (it.next(tile, kb0, len, seg_end)) {
        Work t;
        decode_tile(p, geom, tile, t);
        auto sel = [&](int s) { return s == 0 ? 0 : (s == 1 ? t.cls : pick4(t.tix[0], s - 2)); };
        const int bz1 = sel(p.bz_sel[0]);
        const int bz2 = sel(p.bz_sel[1]);
        const uint32_t tx_bytes = (uint32_t)(rows_in_box * t.nbox) * 128u + (uint32_t)block_n * 128u;
        const int4* kt = reinterpret_cast<const int4*>(p.ktable) + ((size_t)t.cls * n_kb + kb0) * 2;
        int4 e0 = __ldg(kt), e1 = __ldg(kt + 1);
        for (int i = 0; i < len; ++i) {
          int4 f0 = e0, f1 = e1;
          if (i + 1 < len) {  // prefetch the next table entry ahead of the barrier wait
            f0 = __ldg(kt + 2 * (i + 1));
            f1 = __ldg(kt + 2 * (i + 1) + 1);
          }
          const int am = e0.x, c0 = e0.y, o1 = e0.z, o2 = e0.w;
          const int o3 = e1.x, o4 = e1.y, bk = e1.z;
          wait_t(empty_bar(stage), phase ^ 1u, 9);
          // stage layout: normal [positions 16K][weights block_n x 128B];
          //               swap_ab [weights 16K][positions box0 16K][positions box1 16K]
          const uint32_t s0 = smem_base + stage * sbytes;
          const uint32_t x_dst = swap ? s0 + kSlabBytes : s0;
          const uint32_t w_dst = swap ? s0 : s0 + kSlabBytes;
          {
            mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
            tma_load_5d(x_dst, a_maps + am, full_bar(stage), c0, t.org[0][0] + o1, t.org[0][1] + o2,
                        t.org[0][2] + o3, t.org[0][3] + o4);
            if (t.nbox == 2)
              tma_load_5d(x_dst + kSlabBytes, a_maps + am, full_bar(stage), c0, t.org[1][0] + o1,
                          t.org[1][1] + o2, t.org[1][2] + o3, t.org[1][3] + o4);
            tma_load_4d(w_dst, b_map, full_bar(stage), bk, t.n0, bz1, bz2);
          }
          e0 = f0;
          e1 = f1;
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (kEpiWarps == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer));
    // ================================ MMA issuer ====================================
    // one elected thread owns the loop as well (see the producer): waits, tcgen05.mma, commits
    if (!elect_one_sync()) {
    } else if (xreuse) {
      const uint32_t idesc = umma_idesc_bf16(kBlockM, 256u);
      const uint32_t idesc128 = umma_idesc_bf16(kBlockM, 128u);
      int ws = 0, xs = 0;
      uint32_t wphase = 0;
      int seg = 0;
      SegIter it;
      it.init(p, sch, my_r);
      int tile, kb0, len;
      long long seg_end;
      while (it.next(tile, kb0, len, seg_end)) {
        const int buf = seg & 1;
        wait_t(tmem_empty_bar(buf), (((uint32_t)seg >> 1) & 1u) ^ 1u, 8);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * 256);
        for (int i = 0; i < len; ++i) {
          wait_t(full_bar(ws), wphase, 7);
          tc_fence_after();
          const uint32_t info = *reinterpret_cast<volatile uint32_t*>(&s_winfo[ws]);  // written before the arrive
          const bool next_needs_x = (info & 256u) != 0u;
          const uint32_t w_addr = smem_base + ws * kSlabBytes;
          // view of the activation tile for tap dx: rows shifted by dx + 1, groups 10 rows apart
          const uint32_t x_addr = x_ring + xs * x_tile_bytes + (info & 255u) * 128u;
          const uint64_t a_desc = umma_desc_k_sw128(w_addr, 1024);
          const uint64_t b_desc = umma_desc_k_sw128(x_addr, 1280);
          if (xy) {
            // 18-line tiles: the second box starts 180 rows in, which breaks the uniform group
            // stride of a 256-row operand -> one N = 128 instruction per box (same pipe rate)
            const uint64_t b_desc1 = umma_desc_k_sw128(x_addr + (uint32_t)kXYTileBytes / 2u, 1280);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              umma_bf16(d0, a_desc + 2u * k, b_desc + 2u * k, idesc128, (k != 0 || i != 0) ? 1u : 0u);
              umma_bf16(d0 + 128u, a_desc + 2u * k, b_desc1 + 2u * k, idesc128, (k != 0 || i != 0) ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_bf16(d0, a_desc + 2u * k, b_desc + 2u * k, idesc, (k != 0 || i != 0) ? 1u : 0u);
          }
          {
            umma_commit(empty_bar(ws));                        // weight slab free
            if (next_needs_x) umma_commit(xempty_bar(xs));     // activation tile free
            if (i == len - 1) umma_commit(tmem_full_bar(buf));
          }
          if (++ws == S) {
            ws = 0;
            wphase ^= 1u;
          }
          if (next_needs_x) {
            if (++xs == XS) xs = 0;
          }
        }
        ++seg;
      }
      if (trace != nullptr) {
        trace[3] = (uint64_t)clock64();
        trace[11] = (uint64_t)seg;
      }
    } else {
      const uint32_t idesc = umma_idesc_bf16(kBlockM, swap ? 256u : (uint32_t)block_n);
      int stage = 0;
      uint32_t phase = 0;
      int seg = 0;
      SegIter it;
      it.init(p, sch, my_r);
      int tile, kb0, len;
      long long seg_end;
      while (it.next(tile, kb0, len, seg_end)) {
        const int buf = seg & 1;
        wait_t(tmem_empty_bar(buf), (((uint32_t)seg >> 1) & 1u) ^ 1u, 8);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * 256);
        for (int i = 0; i < len; ++i) {
          wait_t(full_bar(stage), phase, 7);
          if (trace != nullptr && seg == 0 && i == 0) trace[2] = (uint64_t)clock64();
          tc_fence_after();
          // the M = 128 operand is always the first slab of the stage
          const uint32_t m_addr = smem_base + stage * sbytes;
          const uint32_t n_addr = m_addr + kSlabBytes;
          const uint64_t a_desc = umma_desc_k_sw128(m_addr, 1024);
          const uint64_t b_desc = umma_desc_k_sw128(n_addr, 1024);
          const uint32_t d = d0 + ((dual && (i & 1)) ? 128u : 0u);
          const uint32_t acc_first = (dual ? (i >= 2) : (i >= 1)) ? 1u : 0u;
          {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              // advance 16 bf16 = 32 bytes inside the 128B swizzle span: +2 in 16-byte units
              umma_bf16(d, a_desc + 2u * k, b_desc + 2u * k, idesc, (k != 0) ? 1u : acc_first);
            }
            umma_commit(empty_bar(stage));  // frees this smem stage once the MMAs have read it
            if (i == len - 1) umma_commit(tmem_full_bar(buf));  // segment complete -> epilogue
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ++seg;
      }
      if (trace != nullptr) {
        trace[3] = (uint64_t)clock64();
        trace[11] = (uint64_t)seg;
      }
    }
  } else if (warp < kEpiWarp0) {
    if (kEpiWarps == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer));  // idle warps of WG0
  } else {
    if (kEpiWarps == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue));
    // ================================ epilogue ======================================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int ew = warp - kEpiWarp0;  // epilogue warp 0..7
    const int half = ew >> 2;         // which half of a tile's columns this warp of the pair takes
    const int r = q * 32 + lane;  // TMEM lane: position (normal) or channel (swap_ab) of this thread
    const int epi_tid = ew * 32 + lane;   // 0..255; threads 0..127 also describe positions
    const int esize = p.out_f32 ? 4 : 2;
    const int chunk_cols_full = p.out_f32 ? 32 : 64;
    const int chunk_cols = block_n < chunk_cols_full ? block_n : chunk_cols_full;
    const int chunk_bytes = chunk_cols * esize;
    const bool swz = (chunk_bytes == 128);
    const int n_chunks = block_n / chunk_cols;
    const int sd = p.sample_dim;
    const bool uniform_sample = (sd == 0) || (p.box[sd - 1] == 1);
    const bool smem_stats = p.stats != nullptr && uniform_sample && p.stats_ld <= kMaxStatGroups;
    const int cpg = p.stats_cpg;
    const float* bias = p.bias;
    int cur_sample = -1;       // sample whose statistics s_stats currently accumulates
    float rb_val = 0.f;        // swap_ab: cached rowbias[(rb_sample, rb_ch)]
    int rb_sample = -1, rb_ch = -1;
    // swap_ab, uniform sample: fp32 statistics sums of channel acc_ch over sample acc_sample
    float s_sum = 0.f, s_sq = 0.f;
    int acc_sample = -1, acc_ch = -1;
    const bool pow2 = cpg > 0 && (cpg & (cpg - 1)) == 0 && cpg <= 32;
    const int span = pow2 ? cpg : ((cpg > 0 && cpg % 32 == 0) ? 32 : 1);  // lanes sharing a stats group
    uint32_t chunk_ctr = 0;    // staging buffer alternation across tiles (normal mode)
    uint32_t resid_phase = 0;  // swap_ab: parity of the residual-tile barrier

    // TMEM-lane index -> box-local coordinates (normal mode: fixed for the whole kernel)
    int rl[4];
    {
      int rr = r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rl[i] = rr % p.box[i];
        rr /= p.box[i];
      }
    }

    // swap_ab: position `epi_tid` of a box -> box-local coordinates (fixed for the whole kernel)
    int pl[4];
    {
      int rr = epi_tid;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pl[i] = rr % p.box[i];
        rr /= p.box[i];
      }
    }
    // swap_ab with short K loops: two sets of staging buffers so that box h+1 is written while
    // the TMA store of box h still reads its set (host picks 3 pipeline stages to make room)
    const bool staging2 = p.staging2 != 0;
    uint32_t box_ctr = 0;

    auto flush_smem_stats = [&]() {  // all 128 epilogue threads
      named_bar_sync(1, kEpiThreads);
      if (cur_sample >= 0 && epi_tid < 2 * p.stats_ld) {
        const double v = s_stats[epi_tid];
        if (v != 0.0) atomicAdd(p.stats + (size_t)cur_sample * p.stats_ld * 2 + epi_tid, v);
        s_stats[epi_tid] = 0.0;
      }
      named_bar_sync(1, kEpiThreads);
    };

    // reduce the lanes that share a statistics group, then one fp64 atomic pair per group
    // (warp-uniform call: acc_sample and the validity of a warp's 32 channels are warp-uniform)
    auto reduce_acc_stats = [&]() {
      if (p.stats != nullptr && acc_sample >= 0 && acc_ch >= 0) {
        for (int o = span >> 1; o > 0; o >>= 1) {
          s_sum += __shfl_xor_sync(0xffffffffu, s_sum, o);
          s_sq += __shfl_xor_sync(0xffffffffu, s_sq, o);
        }
        if (acc_ch < p.n_total && (lane & (span - 1)) == 0) {
          const int g = acc_ch / cpg;
          double* dstp = smem_stats ? &s_stats[g * 2]
                                    : p.stats + ((size_t)acc_sample * p.stats_ld + g) * 2;
          atomicAdd(dstp, (double)s_sum);
          atomicAdd(dstp + 1, (double)s_sq);
        }
      }
      s_sum = 0.f;
      s_sq = 0.f;
    };

    int seg = 0;
    SegIter it;
    it.init(p, sch, my_r);
    int tile, kb0, len;
    long long seg_end;
#ifdef MRI_GEMM_FINE_TRACE
    long long ft_t = 0;
    unsigned long long ft[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
    FT_START();
    while (it.next(tile, kb0, len, seg_end)) {
      const int buf = seg & 1;
      const bool head = (kb0 == 0);
      const bool two_acc = dual && len >= 2;
      const int acc_cols = swap ? 256 : block_n;
      Work t;
      decode_tile(p, geom, tile, t);
      FT_ADD(0);
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256);

      if (trace != nullptr && epi_tid == 0) {   // [12]: epilogue waiting for accumulators
        const long long t0 = clock64();
        mbar_wait(tmem_full_bar(buf), ((uint32_t)seg >> 1) & 1u);
        trace[12] += (uint64_t)(clock64() - t0);
      } else {
        mbar_wait(tmem_full_bar(buf), ((uint32_t)seg >> 1) & 1u);
      }
      tc_fence_after();
      if (trace != nullptr && epi_tid == 0) trace[4] = (uint64_t)clock64();
      FT_ADD(1);

      if (!head) {
        // ---------- partial tile: raw fp32 sums -> workspace, flag -> the tile's head CTA ------
        for (int c0 = half * 16; c0 < acc_cols; c0 += 16 * kSplit) {   // a pair interleaves 16-column groups
          float4* dst = partial_ptr(p, (int)blockIdx.x, r, c0);
          uint32_t v[16], w[16];
          tmem_ld16(tacc + (uint32_t)c0, v);
          if (two_acc) tmem_ld16(tacc + 128u + (uint32_t)c0, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 o;
            o.x = __uint_as_float(v[4 * j + 0]);
            o.y = __uint_as_float(v[4 * j + 1]);
            o.z = __uint_as_float(v[4 * j + 2]);
            o.w = __uint_as_float(v[4 * j + 3]);
            if (two_acc) {
              o.x += __uint_as_float(w[4 * j + 0]);
              o.y += __uint_as_float(w[4 * j + 1]);
              o.z += __uint_as_float(w[4 * j + 2]);
              o.w += __uint_as_float(w[4 * j + 3]);
            }
            __stcg(dst + (size_t)j * kBlockM, o);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar(buf));
        __threadfence();
        named_bar_sync(1, kEpiThreads);
        if (epi_tid == 0) st_release_gpu(p.sk_flags + blockIdx.x, 1u);
      } else {
        // ---------- head of the tile: gather partials (if split), fused epilogue --------------
        int n_contrib = 0;
        const int contrib_cta0 = G - 2 - my_r;  // block index of range my_r + 1
        if (len < n_kb) {
          n_contrib = count_contributors(sch, my_r, seg_end, (long long)(tile + 1) * n_kb);
          if (epi_tid == 0) wait_contributors(p, contrib_cta0, n_contrib);
          named_bar_sync(1, kEpiThreads);
          // add the partial tiles into accumulator 0 (fixed order -> deterministic); loads of the
          // next 16 columns are in flight while the current ones are folded into TMEM
          for (int j = 0; j < n_contrib; ++j) {
            const int cta = contrib_cta0 - j;
            float4 nx[4];
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              nx[k4] = __ldcg(partial_ptr(p, cta, r, half * 16) + (size_t)k4 * kBlockM);
            for (int c0 = half * 16; c0 < acc_cols; c0 += 16 * kSplit) {
              float4 cur[4];
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) cur[k4] = nx[k4];
              if (c0 + 16 * kSplit < acc_cols) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                  nx[k4] = __ldcg(partial_ptr(p, cta, r, c0 + 16 * kSplit) + (size_t)k4 * kBlockM);
              }
              uint32_t v[16];
              tmem_ld16(tacc + (uint32_t)c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                v[4 * k4 + 0] = __float_as_uint(__uint_as_float(v[4 * k4 + 0]) + cur[k4].x);
                v[4 * k4 + 1] = __float_as_uint(__uint_as_float(v[4 * k4 + 1]) + cur[k4].y);
                v[4 * k4 + 2] = __float_as_uint(__uint_as_float(v[4 * k4 + 2]) + cur[k4].z);
                v[4 * k4 + 3] = __float_as_uint(__uint_as_float(v[4 * k4 + 3]) + cur[k4].w);
              }
              tmem_st16(tacc + (uint32_t)c0, v);
            }
            tmem_st_wait();
          }
          // the fused epilogue below splits the columns differently from the interleaved fold:
          // every warp must see its partner's TMEM stores
          tc_fence_before();
          named_bar_sync(1, kEpiThreads);
          tc_fence_after();
        }
        const CUtensorMap* o_map = reinterpret_cast<const CUtensorMap*>(p.o_maps) + t.cls;

        if (!swap) {
          // ======================= normal: lane = position, columns = channels ================
          bool valid = r < rows_in_box;
#pragma unroll
          for (int i = 0; i < 4; ++i) valid = valid && (t.org[0][i] + rl[i] < p.ext[i]);
          const int tile_sample = sd > 0 ? pick4(t.org[0], sd - 1) : 0;
          const int sample = sd > 0 ? tile_sample + pick4(rl, sd - 1) : 0;
          if (smem_stats && tile_sample != cur_sample) {
            flush_smem_stats();
            cur_sample = tile_sample;
          }
          const float* rowbias = (p.rowbias != nullptr && valid)
                                     ? p.rowbias + (size_t)sample * p.rowbias_ld
                                     : nullptr;
          const float bias_m =
              (p.bias_m != nullptr && valid) ? __ldg(p.bias_m + t.org[0][0] + rl[0]) : 0.f;
          const __nv_bfloat16* resid = nullptr;
          if (p.r_base != nullptr && valid) {
            long long off = p.r_cls_off[t.cls];
#pragma unroll
            for (int i = 0; i < 4; ++i) off += (long long)(t.org[0][i] + rl[i]) * p.r_stride[i];
            resid = reinterpret_cast<const __nv_bfloat16*>(p.r_base) + off + t.n0;
          }

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
                if (smem_stats) {
                  atomicAdd(&s_stats[cur_g * 2], (double)s_sum);
                  atomicAdd(&s_stats[cur_g * 2 + 1], (double)s_sq);
                } else {
                  double* dstp = p.stats + ((size_t)tile_sample * p.stats_ld + cur_g) * 2;
                  atomicAdd(dstp, (double)s_sum);
                  atomicAdd(dstp + 1, (double)s_sq);
                }
              }
            } else if (valid) {
              double* dstp = p.stats + ((size_t)sample * p.stats_ld + cur_g) * 2;
              atomicAdd(dstp, (double)s_sum);
              atomicAdd(dstp + 1, (double)s_sq);
            }
            s_sum = 0.f;
            s_sq = 0.f;
          };

          for (int ch = 0; ch < n_chunks; ++ch) {
            const int cbase = ch * chunk_cols;
            if (t.n0 + cbase >= p.n_total) break;  // whole chunk beyond the valid columns
            const uint32_t sbuf = stag + (chunk_ctr & 1u) * kChunkBytes;
            // the TMA store that last used this buffer (two chunks ago) must have read it out
            if (epi_tid == 0) tma_store_wait_read1();
            named_bar_sync(1, kEpiThreads);
            for (int c0 = cbase + half * 16; c0 < cbase + chunk_cols; c0 += 16 * kSplit) {
              uint32_t v[16], w[16];
              tmem_ld16(tacc + (uint32_t)c0, v);
              if (two_acc) tmem_ld16(tacc + 128u + (uint32_t)c0, w);
              tmem_ld_wait();
              float f[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
              if (two_acc) {
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] += __uint_as_float(w[i]);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += bias_m;
              const int nb = t.n0 + c0;
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
                    if (resid != nullptr) {
                      const uint4 rv = *reinterpret_cast<const uint4*>(resid + c0 + h * 8);
                      const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                      for (int j = 0; j < 4; ++j) {
                        f[h * 8 + 2 * j] += __uint_as_float(rw[j] << 16);
                        f[h * 8 + 2 * j + 1] += __uint_as_float(rw[j] & 0xffff0000u);
                      }
                    }
                  }
                }
              }

              if (p.stats != nullptr) {
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

              // staging address of this thread's 16 columns
              const int within = c0 - cbase;
              const uint32_t row_base = sbuf + r * chunk_bytes;
              const int unit0 = (within * esize) >> 4;
              const int xr = swz ? (r & 7) : 0;
              if (p.out_f32) {
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                  const uint32_t addr = row_base + (uint32_t)(((unit0 + uu) ^ xr) << 4);
                  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(f[uu * 4 + 0]),
                               "f"(f[uu * 4 + 1]), "f"(f[uu * 4 + 2]), "f"(f[uu * 4 + 3])
                               : "memory");
                }
              } else {
#pragma unroll
                for (int uu = 0; uu < 2; ++uu) {
                  uint32_t wv[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[uu * 8 + 2 * j], f[uu * 8 + 2 * j + 1]);
                    wv[j] = *reinterpret_cast<uint32_t*>(&h2);
                  }
                  const uint32_t addr = row_base + (uint32_t)(((unit0 + uu) ^ xr) << 4);
                  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(wv[0]),
                               "r"(wv[1]), "r"(wv[2]), "r"(wv[3])
                               : "memory");
                }
              }
            }
            fence_proxy_async_smem();
            named_bar_sync(1, kEpiThreads);
            if (epi_tid == 0) {
              tma_store_5d(o_map, sbuf, t.n0 + cbase, t.org[0][0], t.org[0][1], t.org[0][2],
                           t.org[0][3]);
              tma_store_commit();
            }
            ++chunk_ctr;
          }
          if (p.stats != nullptr) flush_stats();
        } else {
          // ============ swap_ab: lane = channel, columns = positions of box 0 | box 1 ============
          const int ch = t.n0 + r;              // this thread's output channel
          // n_total % 64 == 0: a warp (32 channels) is entirely inside or outside the valid range;
          // outside, the weight rows were zero-filled by TMA and nothing is added or stored
          const bool ch_ok = ch < p.n_total;
          const bool chunk1 = t.n0 + 64 < p.n_total;
          const float bias_c = (bias != nullptr && ch_ok) ? __ldg(bias + ch) : 0.f;
          const int grp = (p.stats != nullptr && ch_ok) ? ch / cpg : 0;
          double* const stats_p = ch_ok ? p.stats : nullptr;
          const bool has_res = p.r_maps != nullptr;
          const bool per_pos_sample = !uniform_sample && (p.rowbias != nullptr || p.stats != nullptr);
          // staging: chunk buffer (q >> 1) holds channels [64*(q>>1), +64); 128B rows, swizzled.
          // Position c0 + i (c0 % 16 == 0) lives in row c0 + i, 16-byte unit (cunit ^ (i & 7)).
          const uint32_t cbyte = (uint32_t)(((q & 1) * 32 + lane) * 2);
          // statistics sums live in registers ACROSS boxes and tiles for as long as the thread's
          // channel and the sample stay the same (with one channel block per position -- Cout <= 128 --
          // that is until the sample changes): the lane shuffles and the two fp64 shared-memory
          // atomics of the reduction cost ~1 k cycles, which used to be paid per box
          if (acc_ch != ch) {
            reduce_acc_stats();
            acc_ch = ch;
          }
          for (int h = 0; h < t.nbox; ++h, ++box_ctr) {
            const uint32_t sset = stag + ((staging2 && (box_ctr & 1u)) ? (uint32_t)kStagingBytes : 0u);
            const uint32_t sbuf = sset + (uint32_t)(q >> 1) * kChunkBytes;
            uint32_t swz8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) swz8[j] = sbuf + ((((cbyte >> 4) ^ (uint32_t)j) << 4) | (cbyte & 15u));
            int oh[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) oh[i] = h ? t.org[1][i] : t.org[0][i];
            // is the box entirely inside the tensor?  (Then every position is valid and, with a
            // uniform sample, no per-position table is needed.)
            bool full = rows_in_box == kBlockM;
#pragma unroll
            for (int i = 0; i < 4; ++i) full = full && (oh[i] + p.box[i] <= p.ext[i]);
            const bool tables = !full || per_pos_sample;
            if (tables) {  // per-position tables of this box (thread epi_tid describes position epi_tid)
              bool ok = epi_tid < rows_in_box;
#pragma unroll
              for (int i = 0; i < 4; ++i) ok = ok && (oh[i] + pl[i] < p.ext[i]);
              const int smp = sd > 0 ? pick4(oh, sd - 1) + pick4(pl, sd - 1) : 0;
              if (ew < 4) {  // threads 0..127 describe the box's 128 positions
                s_pos_info[epi_tid] = ok ? smp : -1;
                const uint32_t bal = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) s_vmask[ew] = bal;  // positions 32 * ew .. + 31
              }
            }
            const int tile_sample = sd > 0 ? pick4(oh, sd - 1) : 0;
            if (uniform_sample && tile_sample != acc_sample) {
              reduce_acc_stats();  // sums of the previous box belong to another sample
              acc_sample = tile_sample;
            }
            if (smem_stats && tile_sample != cur_sample) {
              flush_smem_stats();
              cur_sample = tile_sample;
            }
            // this set of staging buffers must have been read out by the store that last used it
            if (epi_tid == 0) {
              const long long t0 = trace != nullptr ? clock64() : 0;
              if (staging2) tma_store_wait_read1(); else tma_store_wait_read0();
              if (trace != nullptr) trace[13] += (uint64_t)(clock64() - t0);  // [13]: staging busy
            }
            named_bar_sync(1, kEpiThreads);  // also publishes the position tables
            FT_ADD(2);
            if (has_res) {
              // residual box -> the staging buffers (same layout as the output), by TMA
              if (epi_tid == 0) {
                const CUtensorMap* r_map = reinterpret_cast<const CUtensorMap*>(p.r_maps) + t.cls;
                mbar_arrive_expect_tx(resid_bar, (uint32_t)rows_in_box * (chunk1 ? 256u : 128u));
                tma_load_5d(sset, r_map, resid_bar, t.n0, oh[0], oh[1], oh[2], oh[3]);
                if (chunk1)
                  tma_load_5d(sset + kChunkBytes, r_map, resid_bar, t.n0 + 64, oh[0], oh[1], oh[2], oh[3]);
              }
              mbar_wait(resid_bar, resid_phase);
              resid_phase ^= 1u;
            }
            // bias + time-embedding projection of (sample, channel): reloaded only when either changes
            if (p.rowbias != nullptr && uniform_sample && ch_ok &&
                (tile_sample != rb_sample || ch != rb_ch)) {
              rb_val = __ldg(p.rowbias + (size_t)tile_sample * p.rowbias_ld + ch);
              rb_sample = tile_sample;
              rb_ch = ch;
            }
            const float add_c = bias_c + ((p.rowbias != nullptr && uniform_sample && ch_ok) ? rb_val : 0.f);
            int pp_sample = -1;  // per_pos_sample: sample the register sums pp_sum / pp_sq belong to
            float pp_sum = 0.f, pp_sq = 0.f;
            auto flush_pp = [&](float& a, float& b, int smp) {  // warp-uniform call
              if (stats_p != nullptr && smp >= 0) {
                for (int o = span >> 1; o > 0; o >>= 1) {
                  a += __shfl_xor_sync(0xffffffffu, a, o);
                  b += __shfl_xor_sync(0xffffffffu, b, o);
                }
                if ((lane & (span - 1)) == 0) {
                  double* dstp = p.stats + ((size_t)smp * p.stats_ld + grp) * 2;
                  atomicAdd(dstp, (double)a);
                  atomicAdd(dstp + 1, (double)b);
                }
              }
              a = 0.f;
              b = 0.f;
            };
            FT_ADD(3);
            // this warp's half of the box's positions: [c_lo, c_hi)
            const int c_lo = half * (kBlockM / kSplit);
            const int c_hi = rows_in_box < c_lo + kBlockM / kSplit ? rows_in_box : c_lo + kBlockM / kSplit;
            if (full && !has_res && !per_pos_sample) {
              // ---- fast path: every position valid, one sample, nothing to add from memory.  32
              // positions per step (one 32-column TMEM load, the next one in flight): long
              // straight-line blocks give the two warps of a sub-partition independent work to issue
              const bool want_stats = stats_p != nullptr;
              uint32_t v[32];
              tmem_ld32(tacc + (uint32_t)(h * 128 + c_lo), v);
#pragma unroll 1
              for (int c0 = c_lo; c0 < c_hi; c0 += 32) {   // full box: c_hi - c_lo = 128 / kSplit
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + add_c;
                if (c0 + 32 < c_hi) tmem_ld32(tacc + (uint32_t)(h * 128 + c0 + 32), v);
                if (want_stats) {
                  float ps[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};  // short dependency chains
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    ps[i & 3] += f[i];
                    pq[i & 3] = fmaf(f[i], f[i], pq[i & 3]);
                  }
                  s_sum += (ps[0] + ps[1]) + (ps[2] + ps[3]);
                  s_sq += (pq[0] + pq[1]) + (pq[2] + pq[3]);
                }
                const uint32_t rowb = (uint32_t)c0 * 128u;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const __nv_bfloat16 o = __float2bfloat16_rn(f[i]);
                  asm volatile("st.shared.u16 [%0], %1;" ::"r"(swz8[i & 7] + rowb + (uint32_t)i * 128u),
                               "h"(*reinterpret_cast<const uint16_t*>(&o))
                               : "memory");
                }
              }
            } else {
              // ---- general path: partial boxes, residual tile, boxes spanning several samples ----
              uint32_t v[16];
              if (c_lo < c_hi) tmem_ld16(tacc + (uint32_t)(h * 128 + c_lo), v);  // software pipeline
              for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
                tmem_ld_wait();
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + add_c;
                if (c0 + 16 < c_hi) tmem_ld16(tacc + (uint32_t)(h * 128 + c0 + 16), v);
                const uint32_t vm = tables ? ((s_vmask[c0 >> 5] >> (c0 & 31)) & 0xffffu) : 0xffffu;  // valid positions
                const uint32_t rowb = (uint32_t)c0 * 128u;
                if (has_res) {  // TMA zero-filled the positions outside the tensor
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    uint16_t rv;
                    asm volatile("ld.shared.u16 %0, [%1];"
                                 : "=h"(rv)
                                 : "r"(swz8[i & 7] + rowb + (uint32_t)i * 128u));
                    f[i] += __uint_as_float((uint32_t)rv << 16);
                  }
                }
                if (per_pos_sample) {
                  // the box spans several samples (sample = slowest box coordinate): a 16-position
                  // chunk normally lies inside one sample -> register sums, flushed on change
                  int smp = -1;
                  bool same = true;
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const int info = s_pos_info[c0 + i];
                    if (info >= 0) {
                      same = same && (smp < 0 || smp == info);
                      smp = info;
                    }
                  }
                  if (smp >= 0 && same) {
                    if (smp != pp_sample) {
                      flush_pp(pp_sum, pp_sq, pp_sample);
                      pp_sample = smp;
                    }
                    const float rbv = (p.rowbias != nullptr && ch_ok)
                                          ? __ldg(p.rowbias + (size_t)smp * p.rowbias_ld + ch)
                                          : 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                      f[i] += rbv;
                      const float x = ((vm >> i) & 1u) ? f[i] : 0.f;
                      pp_sum += x;
                      pp_sq = fmaf(x, x, pp_sq);
                    }
                  } else if (smp >= 0) {  // chunk straddles a sample boundary: element-wise
                    for (int i = 0; i < 16; ++i) {
                      const int info = s_pos_info[c0 + i];
                      if (info < 0) continue;
                      if (p.rowbias != nullptr && ch_ok) f[i] += __ldg(p.rowbias + (size_t)info * p.rowbias_ld + ch);
                      if (stats_p != nullptr) {
                        double* dstp = p.stats + ((size_t)info * p.stats_ld + grp) * 2;
                        atomicAdd(dstp, (double)f[i]);
                        atomicAdd(dstp + 1, (double)f[i] * (double)f[i]);
                      }
                    }
                  }
                } else if (stats_p != nullptr) {
                  if (vm == 0xffffu) {
                    float ps[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                      ps[i & 3] += f[i];
                      pq[i & 3] = fmaf(f[i], f[i], pq[i & 3]);
                    }
                    s_sum += (ps[0] + ps[1]) + (ps[2] + ps[3]);
                    s_sq += (pq[0] + pq[1]) + (pq[2] + pq[3]);
                  } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                      const float x = ((vm >> i) & 1u) ? f[i] : 0.f;
                      s_sum += x;
                      s_sq = fmaf(x, x, s_sq);
                    }
                  }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const __nv_bfloat16 o = __float2bfloat16_rn(f[i]);
                  asm volatile("st.shared.u16 [%0], %1;" ::"r"(swz8[i & 7] + rowb + (uint32_t)i * 128u),
                               "h"(*reinterpret_cast<const uint16_t*>(&o))
                               : "memory");
                }
              }
            }
            FT_ADD(4);
            if (per_pos_sample) flush_pp(pp_sum, pp_sq, pp_sample);
            if (uniform_sample && !smem_stats) reduce_acc_stats();  // global atomics: no CTA-level table
            FT_ADD(5);
            fence_proxy_async_smem();
            FT_ADD(9);
            named_bar_sync(1, kEpiThreads);
            FT_ADD(6);
            if (epi_tid == 0) {
              tma_store_5d(o_map, sset, t.n0, oh[0], oh[1], oh[2], oh[3]);
              if (chunk1)
                tma_store_5d(o_map, sset + kChunkBytes, t.n0 + 64, oh[0], oh[1], oh[2], oh[3]);
              tma_store_commit();
            }
            FT_ADD(7);
          }
        }
        // TMEM buffer drained -> the MMA warp may start the tile after next in it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar(buf));
        if (n_contrib > 0 && epi_tid == 0) {  // consume the flags: the next launch starts clean
          for (int j = 0; j < n_contrib; ++j) p.sk_flags[contrib_cta0 - j] = 0;
        }
      }
      ++seg;
      FT_ADD(8);
    }
#ifdef MRI_GEMM_FINE_TRACE
    if (trace != nullptr && (epi_tid == 0 || epi_tid == 96)) {
      uint64_t* row = p.trace + (size_t)((epi_tid == 0 ? 1 : 2) * gridDim.x + blockIdx.x) * 16;
      for (int k = 0; k < 12; ++k) row[k] = ft[k];
    }
#endif
    if (swap) reduce_acc_stats();
    if (smem_stats) flush_smem_stats();
    if (epi_tid == 0) tma_store_wait_all();
    if (trace != nullptr && epi_tid == 0) {
      trace[5] = (uint64_t)clock64();
      trace[6] = globaltimer_ns();
    }
  }

  // ---- teardown ------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace mri

// =============================== host side =============================================
using namespace mri;

static int pick_stages(int block_n, int swap_ab, int requested, int staging2 = 0) {
  const int budget = kSmemLimit - 1024 - kStagingBytes * (staging2 ? 2 : 1);
  int s = budget / stage_bytes(block_n, swap_ab);
  if (s > kMaxStages) s = kMaxStages;
  if (requested >= 2 && requested < s) s = requested;
  return s;
}

extern "C" int mri_gemm_smem_bytes(int block_n, int swap_ab, int stages) {
  return pick_stages(block_n, swap_ab, stages) * stage_bytes(block_n, swap_ab) + kStagingBytes + 1024;
}

// Tuning probe (profiles/README.md, "Where the multi-GPU training step loses time"): MRI_GEMM_SMS=n
// makes the persistent GEMM grids n CTAs wide instead of one per SM, which leaves SMs to the
// all-reduce kernels that run under a DDP backward.  Unset: every SM.
static int usable_sms(int hw) {
  static const int env = [] {
    const char* e = getenv("MRI_GEMM_SMS");
    return e != nullptr ? atoi(e) : 0;
  }();
  return (env >= 8 && env < hw) ? env : hw;
}

extern "C" int mri_gemm_workspace_bytes(int* n_ctas_out) {
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return set_cuda_error(e, "mri_gemm_workspace_bytes");
  sms = usable_sms(sms);
  if (n_ctas_out) *n_ctas_out = sms;
  // [sms][128][256] fp32 partial tiles, then [sms] int32 flags (padded to 1 KB)
  const long long bytes = (long long)sms * kBlockM * kPartialLd * 4 + 1024;
  return (int)bytes;
}

extern "C" int mri_gemm_launch(const MriGemmArgs* a, void* stream) {
  if (a == nullptr) return set_error(-1, "mri_gemm_launch: null args");
  const int bn = a->block_n;
  if (!(bn == 16 || bn == 32 || bn == 64 || bn == 128 || bn == 256))
    return set_error(-2, "mri_gemm_launch: block_n must be 16/32/64/128/256");
  if (a->n_kb < 1 || a->n_class < 1) return set_error(-2, "mri_gemm_launch: empty K loop");
  long long rows = 1;
  long long boxes = 1;
  for (int i = 0; i < 4; ++i) {
    if (a->box[i] < 1 || a->tiles[i] < 1) return set_error(-2, "mri_gemm_launch: bad box/tiles");
    rows *= a->box[i];
    boxes *= a->tiles[i];
  }
  const int swap = a->swap_ab != 0;
  if (swap && (bn != 128 || a->n_total % 64 != 0 || a->out_f32 || a->bias_m != nullptr ||
               a->bz_sel[0] > 1 || a->bz_sel[1] > 1))
    return set_error(-2, "mri_gemm_launch: swap_ab needs block_n 128, n_total % 64 == 0, bf16 "
                         "output, no bias_m and class-only weight selection");
  if (a->r_base != nullptr) {  // 16-byte residual loads / 8-element offset units
    bool ok = true;
    for (int i = 0; i < 4; ++i) ok = ok && (a->r_stride[i] % 8 == 0);
    for (int c = 0; c < a->n_class && c < 8; ++c) ok = ok && (a->r_cls_off[c] % 8 == 0);
    if (!ok) return set_error(-2, "mri_gemm_launch: residual strides/offsets must be multiples of 8");
  }
  const long long tiles = (long long)a->n_tiles_n * a->n_class * (swap ? (boxes + 1) / 2 : boxes);
  if (rows > kBlockM) return set_error(-2, "mri_gemm_launch: box has more than 128 rows");
  if (tiles >= (1LL << 24) || boxes + 1 >= (1LL << 24))
    return set_error(-2, "mri_gemm_launch: more than 2^24 tiles (tile decoding uses float reciprocals)");
  if (tiles > 0x7fffffffLL / a->n_kb) return set_error(-2, "mri_gemm_launch: too many work units");
  if (a->n_total % 8 != 0) return set_error(-2, "mri_gemm_launch: n_total must be a multiple of 8");
  if (a->stats != nullptr && !a->swap_ab && (a->stats_cpg < 8 || a->stats_cpg % 8 != 0))
    return set_error(-2, "mri_gemm_launch: statistics need channel groups in multiples of 8");
  if (a->stats != nullptr && a->stats_cpg < 1) return set_error(-2, "mri_gemm_launch: stats_cpg < 1");
  if (a->r_base != nullptr && a->out_f32)
    return set_error(-2, "mri_gemm_launch: residual input requires bf16 output");
  if (a->r_base != nullptr && a->n_class > 8)
    return set_error(-2, "mri_gemm_launch: residual supports at most 8 output classes");
  if ((a->r_base != nullptr) != (a->r_maps != nullptr))
    return set_error(-2, "mri_gemm_launch: residual needs both r_maps and r_base");
  if (a->out_f32 && bn > 128)
    return set_error(-2, "mri_gemm_launch: fp32 output needs block_n <= 128");
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaDeviceGetAttribute(SM count)");
    n_sms = usable_sms(n_sms);
  }
  MriGemmArgs k = *a;
  // short K loops (2D convolutions: 9 taps): the epilogue, not the main loop, paces the CTA ->
  // trade one pipeline stage for a second set of staging buffers
  k.staging2 = (swap && a->n_kb <= 40) ? 1 : 0;
  k.stages = pick_stages(bn, swap, a->stages, k.staging2);
  int smem = k.stages * stage_bytes(bn, swap) + kStagingBytes * (k.staging2 ? 2 : 1) + 1024;
  if (a->tile_fast_dim < 0 || a->tile_fast_dim > 3) return set_error(-2, "mri_gemm_launch: tile_fast_dim must be 0..3");
  if (a->xreuse) {
    if (!swap || a->box[0] != 8 || rows != kBlockM || a->bz_sel[0] > 1 || a->bz_sel[1] > 1)
      return set_error(-2, "mri_gemm_launch: xreuse needs swap_ab and boxes of 8 x 16 positions");
    // per-CTA wait trace (tools/gemm_probe.py trace=1): the MMA thread waits for operands ~24 %
    // of the time with either ring split -- TMA latency under load is ~2 k cycles and the rings
    // hold ~3 k cycles of work
    static const int env_xring = [] {
      const char* e = getenv("MRI_GEMM_XRING");  // tuning probe: 3 = weight ring 4 + 3 tiles
      return e != nullptr ? atoi(e) : 0;
    }();
    // weight ring 6 + 2 activation tiles vs 4 + 3: within 1 % of each other on the same GPU
    // (the step is power-bound, not latency-bound); 6 + 2 is the default
    k.stages = (env_xring == 3 || k.staging2) ? 4 : 6;
    smem = k.stages * kSlabBytes + ((k.staging2 || k.stages > 4) ? 2 : kXStages) * kXTileBytes +
           kStagingBytes * (k.staging2 ? 2 : 1) + 1024;
    if (a->xreuse == 2) {  // 10 x 18 tiles shared by nine taps: 2 tiles, weight ring of 6 (4)
      if (a->box[1] != 16) return set_error(-2, "mri_gemm_launch: xreuse 2 needs boxes of 8 x 16 positions");
      k.stages = k.staging2 ? 4 : 6;
      smem = k.stages * kSlabBytes + 2 * kXYTileBytes + kStagingBytes * (k.staging2 ? 2 : 1) + 1024;
    }
  }
  long long grid = tiles < n_sms ? tiles : n_sms;
  if (k.sched != 0) {
    if (k.sk_partials == nullptr || k.sk_flags == nullptr)
      return set_error(-2, "mri_gemm_launch: stream-K schedule needs the workspace");
    if (tiles < n_sms) {
      // fewer tiles than SMs: split K loops into shares of >= max(8, n_kb / 4) k-steps
      const long long units = tiles * k.n_kb;
      const int min_share = (k.n_kb + 3) / 4 > 8 ? (k.n_kb + 3) / 4 : 8;
      const long long want = units / min_share > tiles ? units / min_share : tiles;
      grid = want < n_sms ? want : n_sms;
    }
    if (grid > k.sk_ctas) return set_error(-2, "mri_gemm_launch: stream-K workspace too small");
  }
  static int configured_smem = 0;
  if (smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gemm_tc_kernel)");
    configured_smem = smem;
  }
  gemm_tc_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(k);
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
