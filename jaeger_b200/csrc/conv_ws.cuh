// Stage 3 hot kernel, weights-stationary version for the 128-channel layers that carry the time.
//
// The CTA-pair kernel (conv_tc2.cuh) is bound by shared-memory bandwidth: every tcgen05.mma reads its
// activation operand (4 KB) AND its half of the weights (2 KB) from shared memory, 96 of the SM's 128 B/clk,
// and the epilogue's staging traffic has to share the rest (profiles/conv_kernel_r1.md).  Here the operand
// roles are swapped:
//
//     D^T[Cout = 128 lanes, rows] = sum_taps  W_t^T[Cout, Cin]  x  X^T[Cin, rows + shift_t]
//
//   * A operand = the layer's weights, transposed, RESIDENT IN TENSOR MEMORY for the life of the CTA
//     (5 taps x 128 input channels = 320 of the 512 TMEM columns; two fp16 per 32-bit column).  A from TMEM
//     costs no shared-memory bandwidth at all.
//   * B operand = activation rows, K-major SWIZZLE_128B straight from the pre-swizzled HBM layout (conv_common.cuh);
//     a conv tap is still a row shift of the descriptor.  N = 64 rows per MMA: 2 KB per 32 cycles = 64 B/clk.
//   * D^T: fp32, lanes = output channels, columns = rows; a ring of 3 x 64 columns in the remaining TMEM.
//   * No weights in shared memory: all 227 KB hold operand stages and output / shortcut staging tiles.
//
// The transposed accumulator also moves every per-channel quantity (BatchNorm shift, second affine, masked-row
// constants, NMD column sums, pooled maxima) from "per register, fetched from shared memory" to "per thread,
// kept in registers": the NMD tap becomes a running per-thread sum that is reduced over four lanes once per
// window instead of a 19-shuffle butterfly per 32 channels per tile.
// The epilogue reads the accumulator with tcgen05.ld.16x256b (the mma C-fragment layout), so that row pairs of
// one channel sit in one register as a half2, and moves tiles between registers and the [row][channel] staging
// tile with stmatrix / ldmatrix .trans -- the transposition costs no instructions.
//
// Roles: warp 0 TMA producer (operand stages), warps 1 and 3 MMA issuers (first / second sub-tile of every tile),
// warp 2 TMEM allocator + helper (shortcut-tile loads, row validity), warps 4..15 three epilogue groups (TMEM lane quarter = warp % 4) draining the
// 64-row sub-tiles round-robin.  An N = 64 MMA lasts 32 cycles, so the issue loop is fully unrolled for the layer
// shapes that carry the time (kTaps x kGroups compile-time) and runs on one lane: two uniform adds per MMA.  Reference layers: nnlib/v2/layers.py:1217-1280, 918-941, 1882-1915, 517-529; nnlib/v2/nmd.py:52-77.
#pragma once
#include "conv_tc2.cuh"

namespace jg {
namespace ws {

using namespace jg::tc;
using jg::tc2::bulk_commit;
using jg::tc2::bulk_s2g;
using jg::tc2::bulk_wait_all;
using jg::tc2::bulk_wait_read;
using jg::tc2::fence_async_smem;

constexpr int kEpiGroupsWs = 3;       // epilogue groups of 4 warps draining sub-tiles round-robin (group g <-> accumulator g)
constexpr int kThreadsWs = 128 + 128 * kEpiGroupsWs;   // 4 control warps + the epilogue warps
constexpr int kStagesWs = 4;          // operand stages of [lead + 128 + tail rows][64 channels]
constexpr int kSubRows = 64;          // rows per accumulator = MMA N
constexpr int kAccWs = 3;             // accumulators: 3 x 64 TMEM columns behind the weights
constexpr int kSlotsWs = 9;           // output / shortcut staging slots (3 per epilogue group), each 64 rows x 128 channels (16 KB)
constexpr int kWColsMax = 320;        // TMEM columns that may hold weights (320 + 3 * 64 = 512)
#ifdef JG_WS_VSLOTS4
constexpr int kVSlotsWs = 4;
#else
constexpr int kVSlotsWs = 8;          // validity ring (sub-tiles): the helper's global loads run this far ahead
#endif
constexpr uint32_t kSlotBytes = kSubRows * 128u * 2u;

enum WsMode { WS_LIGHT = 0, WS_LIGHT_SC = 1, WS_FINAL = 2, WS_FINAL_POOL = 3 };

struct SmemLayoutWs {
  uint32_t stage_off, slot_off, bar_off, val_off, total;
  uint32_t stage_bytes, stage_pitch, rows_a, lead, groups;
};

__host__ __device__ inline SmemLayoutWs smem_layout_ws(int cin, int halo_l, int halo_r) {
  SmemLayoutWs L;
  L.lead = static_cast<uint32_t>((halo_l + 7) / 8 * 8);
  L.rows_a = L.lead + kTileM + static_cast<uint32_t>((halo_r + 7) / 8 * 8);
  L.groups = cin / 64;
  L.stage_bytes = L.rows_a * 128;
  L.stage_pitch = (L.stage_bytes + 1023u) & ~1023u;
  L.stage_off = 0;
  L.slot_off = kStagesWs * L.stage_pitch;
  L.bar_off = L.slot_off + kSlotsWs * kSlotBytes;
  L.val_off = L.bar_off + 512u;
  L.total = L.val_off + kVSlotsWs * 16u + 1024u;
  return L;
}

// weights image of this kernel: Wt[cout][tap * cin + ci] fp16 (row = one TMEM lane, two K elements per column)
__host__ __device__ __forceinline__ long long w3_index(int t, int ci, int co, int cin, int ntaps) {
  return static_cast<long long>(co) * (ntaps * cin) + t * cin + ci;
}

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 16 TMEM lanes x 32 columns in the mma C-fragment layout: v[4u + 2h + e] = (lane base + 8h + t/4, column 8u + 2(t%4) + e)
__device__ __forceinline__ void tmem_ld16x256_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<const __half2*>(&u); }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<const uint32_t*>(&h); }
// 0xFFFF per valid row of the pair (bits 0 / 1 of b)
__device__ __forceinline__ uint32_t pair_mask(uint32_t b) { return (b & 1u) * 0xFFFFu | ((b >> 1) & 1u) * 0xFFFF0000u; }

#ifdef JG_WS_DEBUG_HANG
// probe builds only: a wait that gives up after ~2^21 probes, records who was stuck where and lets the kernel run on
__device__ __forceinline__ void mbar_wait_dbg(uint32_t bar, uint32_t parity, int* err, int code, int it) {
  for (uint32_t n = 0; n < (1u << 21); ++n)
    if (mbar_try_wait(bar, parity)) return;
  if ((threadIdx.x & 31) == 0 && atomicCAS(err, 0, code) == 0) { err[1] = it; err[2] = blockIdx.x; err[3] = static_cast<int>(parity); err[4] = threadIdx.x >> 5; }
}
#define WS_WAIT(bar, parity, code, it) mbar_wait_dbg(bar, parity, p.err, code, it)
#else
#define WS_WAIT(bar, parity, code, it) mbar_wait(bar, parity)
#endif

// Row validity of one 128-row tile = two 64-row sub-tiles, as bit masks (bit = row within its 32-row chunk):
// dst_x[0..1] output rows valid, dst_x[2..3] shortcut rows valid, for sub-tile a (rows 0..63) and b (rows 64..127).
// Every global load of the tile is issued before the first dependent instruction: one memory round trip per tile.
// With fuse_mask this is also the layer's mask propagation (conv_epilogue.cuh:tile_validity).
__device__ __forceinline__ void tile_validity_bits(const ConvParams& p, long long row0, int lane, volatile uint32_t* dst_a,
                                                   volatile uint32_t* dst_b) {
  const int win = static_cast<int>(row0 / p.rows_per_window);
  uint32_t any[4] = {0u, 0u, 0u, 0u}, scm[4];
  int lp = 0;
  if (p.fuse_mask) {
    lp = p.lpad[win];
    if (p.masking) {
      for (int t0 = 0; t0 < p.ntaps; t0 += 8) {
        uint32_t m[4][8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool on = t0 + u < p.ntaps;
          const int sh = on ? p.shifts[t0 + u] : 0;
#pragma unroll
          for (int c = 0; c < 4; ++c) m[c][u] = on ? p.in_mask[row0 + c * 32 + lane + sh] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int c = 0; c < 4; ++c) any[c] |= m[c][u];
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) any[c] = p.out_mask[row0 + c * 32 + lane];
  }
  const bool has_scm = p.sc != nullptr && p.sc_mask != nullptr;
#pragma unroll
  for (int c = 0; c < 4; ++c) scm[c] = has_scm ? p.sc_mask[row0 + c * 32 + lane] : (p.sc != nullptr ? 1u : 0u);
  const int limit = ((lp - p.shrink_in + p.len_round) >> p.halvings) - p.shrink;
  int n_valid = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const long long row = row0 + c * 32 + lane;
    bool ok;
    if (p.fuse_mask) {
      const int rw = static_cast<int>(row - static_cast<long long>(win) * p.rows_per_window);
      const int f = rw / p.period, j = rw - f * p.period;
      ok = f < p.frames && j < limit && (!p.masking || any[c] != 0u);
      p.out_mask_w[row] = static_cast<uint8_t>(ok);
    } else {
      ok = any[c] != 0u;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const unsigned sbal = __ballot_sync(0xffffffffu, scm[c] != 0u);
    n_valid += __popc(bal);
    if (lane == 0) {
      volatile uint32_t* dst = c < 2 ? dst_a : dst_b;
      dst[c & 1] = bal;
      dst[2 + (c & 1)] = sbal;
    }
  }
  if (p.fuse_mask && lane == 0 && n_valid) atomicAdd(p.count + win, n_valid);
}

// kTaps / kGroups > 0: the layer's tap count and Cin / 64 at compile time (unrolled MMA issue); 0 = run-time loops.
template <int kMode, int kTaps, int kGroups>
__global__ void __launch_bounds__(kThreadsWs, 1) conv_ws_kernel(const __grid_constant__ ConvParams p) {
  constexpr bool kSc = kMode != WS_LIGHT;
  constexpr bool kFinal = kMode == WS_FINAL || kMode == WS_FINAL_POOL;
  constexpr bool kPool = kMode == WS_FINAL_POOL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const SmemLayoutWs L = smem_layout_ws(p.cin, p.halo_l, p.halo_r);

  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  const uint32_t bar0 = smem_u32(s_bar);
  // barrier slots: FULL[S] EMPTY[S] TFULL[3] TEMPTY[3][2] VFULL[V] VEMPTY[V] SFULL[9] SFREE[9], then the TMEM base word.
  // TEMPTY[a][w]: accumulator a handed back to MMA issuer w.  The two issuers use an accumulator alternately; with ONE barrier
  // per accumulator each issuer would wait for every other phase of it, and a parity wait cannot tell phase u from phase u + 2
  // (an issuer that ran ahead passed it early and overwrote a live accumulator).  The epilogue signals the issuer of the
  // accumulator's NEXT sub-tile, so every barrier has one waiter whose waits are for consecutive phases.
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (kStagesWs + s); };
  auto TFULL = [&](int a) { return bar0 + 8u * (2 * kStagesWs + a); };
  auto TEMPTY = [&](int a, int w) { return bar0 + 8u * (2 * kStagesWs + kAccWs + 2 * a + w); };
  auto VFULL = [&](int v) { return bar0 + 8u * (2 * kStagesWs + 3 * kAccWs + v); };
  auto VEMPTY = [&](int v) { return bar0 + 8u * (2 * kStagesWs + 3 * kAccWs + kVSlotsWs + v); };
  auto SFULL = [&](int s) { return bar0 + 8u * (2 * kStagesWs + 3 * kAccWs + 2 * kVSlotsWs + s); };
  auto SFREE = [&](int s) { return bar0 + 8u * (2 * kStagesWs + 3 * kAccWs + 2 * kVSlotsWs + kSlotsWs + s); };
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * kStagesWs + 3 * kAccWs + 2 * kVSlotsWs + 2 * kSlotsWs);
  volatile uint32_t* s_valid = reinterpret_cast<volatile uint32_t*>(smem + L.val_off);

  const uint32_t st_base = smem_u32(smem + L.stage_off);
  const uint32_t slot_base = smem_u32(smem + L.slot_off);
  const int groups = kGroups > 0 ? kGroups : static_cast<int>(L.groups);
  const int ntaps = kTaps > 0 ? kTaps : p.ntaps;
  const int ktot = ntaps * p.cin;                 // K of the implicit GEMM; ktot / 2 TMEM columns hold the weights
  const uint32_t w_cols = static_cast<uint32_t>(ktot / 2);
  const uint32_t acc_col0 = kWColsMax;              // accumulators always sit behind the 320-column weight region

  // contiguous range of 128-row tiles for this CTA (neighbouring tiles share halo rows in L2)
  const int tile_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * p.n_tiles / gridDim.x);
  const int tile_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.n_tiles / gridDim.x);
  const int n_sub = 2 * (tile_end - tile_begin);

  const bool dbg0 = p.dbg != nullptr && blockIdx.x == 0;      // probe only: per-role wait counters of CTA 0
  if (dbg0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.dbg[0] = clock64(); p.dbg[15] = n_sub; p.dbg[16] = static_cast<long long>(gt);
  }
  // ---- one-time setup ------------------------------------------------------------------
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStagesWs; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 2); }   // both MMA issuers release a stage
    for (int a = 0; a < kAccWs; ++a) { mbar_init(TFULL(a), 1); mbar_init(TEMPTY(a, 0), 4); mbar_init(TEMPTY(a, 1), 4); }
    for (int v = 0; v < kVSlotsWs; ++v) { mbar_init(VFULL(v), 1); mbar_init(VEMPTY(v), 4); }
    for (int s = 0; s < kSlotsWs; ++s) { mbar_init(SFULL(s), 1); mbar_init(SFREE(s), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(smem_u32(s_tmem), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (warp >= 4 && warp < 8) {
    // weights -> tensor memory: thread = output channel = TMEM lane, 32 columns (64 K elements) per store
    const int ch = (warp & 3) * 32 + lane;
    const uint4* src = reinterpret_cast<const uint4*>(p.w + static_cast<size_t>(ch) * ktot);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (uint32_t c0 = 0; c0 < w_cols; c0 += 32) {
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 q4 = src[c0 / 4 + j];
        v[4 * j + 0] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w;
      }
      tmem_st32(t_lane + c0, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ===== TMA producer: per tile, one halo'd [lead + 128 + tail][64] slab per channel group, then (layers with a
    // shortcut) the block-input rows of the tile's two sub-tiles into their staging slots =====
    const bool leader = elect_one();
    int cnt = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const long long r_first = static_cast<long long>(tile) * kTileM - L.lead;
      for (int g = 0; g < groups; ++g, ++cnt) {
        const int s = cnt % kStagesWs;
        const uint32_t ph = static_cast<uint32_t>(cnt / kStagesWs) & 1u;
        WS_WAIT(EMPTY(s), ph ^ 1u, 1, tile);
        if (leader) {
          mbar_expect_tx(FULL(s), L.stage_bytes);
          const act_t* src = p.x + (static_cast<long long>(g) * p.x_plane + r_first) * 64;
          bulk_g2s(st_base + s * L.stage_pitch, src, L.stage_bytes, FULL(s));
        }
      }
      if (kSc) {
        const int it = 2 * (tile - tile_begin);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int slot = (it + h) % kSlotsWs;
          const uint32_t use = static_cast<uint32_t>((it + h) / kSlotsWs);
          WS_WAIT(SFREE(slot), (use & 1u) ^ 1u, 4, it + h);
          if (leader) {
            mbar_expect_tx(SFULL(slot), kSlotBytes);
            for (int cg = 0; cg < 2; ++cg)
              bulk_g2s(slot_base + slot * kSlotBytes + cg * (kSlotBytes / 2),
                       p.sc + (static_cast<long long>(cg) * p.y_plane + static_cast<long long>(tile) * kTileM + h * kSubRows) * 64,
                       kSlotBytes / 2, SFULL(slot));
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===== MMA issuers: A = weights in TMEM, B = rows from the stage, D^T accumulators of 64 columns =====
    // tcgen05.mma issue is close to synchronous with execution (the time a lone issuer spends in barrier waits shows up as
    // idle tensor cycles, measured with tools/conv_probe.cu), so TWO warps issue: warp 1 the first, warp 3 the second 64-row
    // sub-tile of every tile.  Their MMAs go to different accumulators and interleave freely in the pipe; while one warp
    // sits in its waits the other keeps the pipe fed.
    const int sub = warp == 1 ? 0 : 1;
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(kSubRows >> 3) << 17) | (static_cast<uint32_t>(kTileM >> 4) << 24);
    uint32_t boff[kTaps > 0 ? kTaps : 1];                      // tap row shifts in descriptor units (128 B per row / 16)
    if (kTaps > 0) {
#pragma unroll
      for (int t = 0; t < kTaps; ++t) boff[t] = static_cast<uint32_t>(p.shifts[t] * 8);
    }
    const uint32_t tap_cols = static_cast<uint32_t>(p.cin >> 1);   // TMEM columns per tap
    const uint32_t sub_off = (L.lead + static_cast<uint32_t>(sub * kSubRows)) * 128u;
    long long d_wacc = 0, d_wfull = 0, d_issue = 0;               // probe only (p.dbg): kept in registers, written once
    int as = sub;                                                  // accumulator of sub-tile it = it % 3, it = 2 * tile + sub
    int it = sub;
    int s0 = 0;                                                    // stage of the tile's first channel group
    uint32_t sph = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const long long tw0 = dbg0 ? clock64() : 0;
      // the accumulator's previous sub-tile (it - 3) has been read out; this issuer's (n + 1)-th hand-back of this accumulator
      if (it >= kAccWs) WS_WAIT(TEMPTY(as, sub), static_cast<uint32_t>((it - kAccWs) / (2 * kAccWs)) & 1u, 2, tile);
      tc_fence_after();
      if (dbg0) d_wacc += clock64() - tw0;
      const uint32_t d_tmem = tmem_base + acc_col0 + static_cast<uint32_t>(as * kSubRows);
      int s = s0;
      uint32_t ph = sph;
#pragma unroll
      for (int g = 0; g < (kGroups > 0 ? kGroups : 4); ++g) {
        if (kGroups == 0 && g >= groups) break;
        const long long tw1 = dbg0 ? clock64() : 0;
        WS_WAIT(FULL(s), ph, 3, tile);
        tc_fence_after();
        const long long ti0 = dbg0 ? clock64() : 0;
        if (dbg0) d_wfull += ti0 - tw1;
        const uint32_t b_lo = desc_lo_sw128(st_base + s * L.stage_pitch + sub_off);
        const uint32_t a_g = tmem_base + static_cast<uint32_t>(g * 32);
        if (leader) {                                            // one lane issues; nothing in here diverges
          if (kTaps > 0) {
#pragma unroll
            for (int t = 0; t < kTaps; ++t) {
#pragma unroll
              for (uint32_t k16 = 0; k16 < 4; ++k16)
                umma_f16_ts(d_tmem, a_g + static_cast<uint32_t>(t) * tap_cols + k16 * 8u,
                            desc_pack(b_lo + boff[t] + k16 * 2u, kDescHiSw128), idesc, (g | t | static_cast<int>(k16)) != 0);
            }
          } else {
#pragma unroll 1
            for (int t = 0; t < ntaps; ++t) {
              const uint32_t b_tap = b_lo + static_cast<uint32_t>(p.shifts[t] * 8);
#pragma unroll
              for (uint32_t k16 = 0; k16 < 4; ++k16)
                umma_f16_ts(d_tmem, a_g + static_cast<uint32_t>(t) * tap_cols + k16 * 8u, desc_pack(b_tap + k16 * 2u, kDescHiSw128),
                            idesc, (g | t | static_cast<int>(k16)) != 0);
            }
          }
          umma_commit(EMPTY(s));                                 // this issuer's reads of the stage are done when these retire
        }
        __syncwarp();
        if (dbg0) d_issue += clock64() - ti0;
        if (++s == kStagesWs) { s = 0; ph ^= 1u; }
      }
      if (leader) umma_commit(TFULL(as));
      __syncwarp();
      s0 = s; sph = ph;
      it += 2;                                                   // next sub-tile of this issuer
      as += 2;
      if (as >= kAccWs) as -= kAccWs;
    }
    if (dbg0 && leader && sub == 0) { p.dbg[1] = d_wacc; p.dbg[2] = d_wfull; p.dbg[13] = d_issue; }
  } else if (warp == 2) {
    // ===== helper: row masks / window counts of a tile's two sub-tiles, up to kVSlotsWs sub-tiles ahead of the epilogue =====
    for (int it = 0; it < n_sub; it += 2) {                       // one 128-row tile = sub-tiles it, it + 1 per iteration
      const long long row0 = static_cast<long long>(tile_begin + (it >> 1)) * kTileM;
      const int va = it & (kVSlotsWs - 1), vb = (it + 1) & (kVSlotsWs - 1);
      const uint32_t vph = ((static_cast<uint32_t>(it) / kVSlotsWs) & 1u) ^ 1u;      // it is even and the ring size even: same phase for both
      WS_WAIT(VEMPTY(va), vph, 5, it);
      WS_WAIT(VEMPTY(vb), vph, 5, it + 1);
      tile_validity_bits(p, row0, lane, s_valid + va * 4, s_valid + vb * 4);
      __syncwarp();
      if (lane == 0) { mbar_arrive(VFULL(va)); mbar_arrive(VFULL(vb)); }
    }
  } else {
    // ===== epilogue: thread = 4 channels (TMEM lanes) x row pairs; everything per-channel is a register =====
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int cg = q >> 1;                              // 64-channel group this warp's channels belong to
    const int t4 = lane >> 2, tq = lane & 3;
    // the bulk stores of (grp, cg) are issued by lanes 0..3 of the even warp in turn: a lane only ever waits for a store
    // of its own that is two sub-tiles of this group old (bulk groups are per thread), never for the most recent one
    const bool issuer_warp = (q & 1) == 0;
    const int bar_id = 1 + grp * 2 + cg;
    // channel of (slab s, half h): 32 q + 16 s + 8 h + t4
    float sh1[2][2];
    __half2 a2[2][2], b2[2][2], scc[2][2];
    float tapacc[2][2], poolacc[2][2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ch = 32 * q + 16 * s + 8 * h + t4;
        sh1[s][h] = p.shift1[ch];
        a2[s][h] = __float2half2_rn(kFinal ? p.scale2[ch] : 1.0f);
        b2[s][h] = __float2half2_rn(kFinal ? p.shift2[ch] : 0.0f);
        scc[s][h] = __float2half2_rn((kSc && p.sc_const) ? p.sc_const[ch] : 0.0f);
        tapacc[s][h] = 0.0f;
        poolacc[s][h] = -3.0e38f;
      }
    // per-thread part of the ldmatrix / stmatrix address: thread k supplies row (k & 7) of matrix (k >> 3) = (h, u & 1)
    const int m_h = (lane >> 3) & 1, m_u = lane >> 4, m_i = lane & 7;
    int cur_win = -1;
    long long e_v = 0, e_s = 0, e_t = 0, e_m = 0, e_b = 0, e_tot = 0;   // probe only
    auto flush = [&](int w) {
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = 32 * q + 16 * s + 8 * h + t4;
          if (kFinal) {
            float v = tapacc[s][h];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (tq == 0) atomicAdd(p.tap_sum + static_cast<long long>(w) * red_pitch_of(p) + ch, v);
            tapacc[s][h] = 0.0f;
          }
          if (kPool) {
            float v = poolacc[s][h];
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
            if (tq == 0 && v > -1.0e38f) atomic_max_f32(p.pool + static_cast<long long>(w) * red_pitch_of(p) + ch, v);
            poolacc[s][h] = -3.0e38f;
          }
        }
    };

    const int tiles_per_win = p.rows_per_window / kTileM;      // rows_per_window is a multiple of 128
    for (int it = grp; it < n_sub; it += kEpiGroupsWs) {
      const int as = it % kAccWs;
      const uint32_t aph = static_cast<uint32_t>(it / kAccWs) & 1u;
      const int slot = it % kSlotsWs;
      const uint32_t use = static_cast<uint32_t>(it / kSlotsWs);
      const long long row0 = static_cast<long long>(tile_begin + (it >> 1)) * kTileM + (it & 1) * kSubRows;
      const int win = (tile_begin + (it >> 1)) / tiles_per_win;
      if (kFinal && win != cur_win) {
        if (cur_win >= 0) flush(cur_win);
        cur_win = win;
      }
      // this group's stores before the most recent one have finished reading their slots: hand the oldest slot on
      // (it is the one this group uses after the next sub-tile, so the shortcut loader runs a whole period ahead)
      const int k_it = it / kEpiGroupsWs;                     // this group's iteration index
      if (issuer_warp && p.y && k_it >= 2 && lane == ((k_it - 2) & 3)) {
        bulk_wait_read<0>();
        mbar_arrive(SFREE((it + kEpiGroupsWs) % kSlotsWs));
      }
      const int vslot = it & (kVSlotsWs - 1);
      const bool dbgw = dbg0 && q == 0 && lane == 0;
      const long long tv0 = dbgw ? clock64() : 0;
      WS_WAIT(VFULL(vslot), (static_cast<uint32_t>(it) / kVSlotsWs) & 1u, 6, it);
      if (dbgw) e_v += clock64() - tv0;
      uint32_t vmask[2], smask[2];
      vmask[0] = s_valid[vslot * 4 + 0]; vmask[1] = s_valid[vslot * 4 + 1];
      smask[0] = s_valid[vslot * 4 + 2]; smask[1] = s_valid[vslot * 4 + 3];
      __syncwarp();
      if (lane == 0) mbar_arrive(VEMPTY(vslot));
      const long long ts0 = dbgw ? clock64() : 0;
      if (kSc) WS_WAIT(SFULL(slot), use & 1u, 7, it);            // the shortcut rows have landed in the slot
      else WS_WAIT(SFREE(slot), (use & 1u) ^ 1u, 8, it);          // the slot's previous store has been read
      const long long ts1 = dbgw ? clock64() : 0;
      WS_WAIT(TFULL(as), aph, 9, it);
      tc_fence_after();
      const long long ts2 = dbgw ? clock64() : 0;
      if (dbgw) { e_s += ts1 - ts0; e_t += ts2 - ts1; }
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_col0 + static_cast<uint32_t>(as * kSubRows);
      uint32_t acc[2][2][16];                                 // [chunk][slab][fragment]
      auto release_acc = [&]() {                              // every TMEM read of the sub-tile has landed in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(TEMPTY(as, (it + kAccWs) & 1));      // to the issuer of this accumulator's next sub-tile
      };
#ifdef JG_WS_FINAL_UPFRONT
      constexpr bool kChunked = false;
#else
      constexpr bool kChunked = kFinal;
#endif
      if (!kChunked) {          // light epilogues: the whole sub-tile fits the register budget, the accumulator is freed at once
        tmem_ld16x256_x4(t_addr, acc[0][0]);
        tmem_ld16x256_x4(t_addr + (16u << 16), acc[0][1]);
        tmem_ld16x256_x4(t_addr + 32u, acc[1][0]);
        tmem_ld16x256_x4(t_addr + (16u << 16) + 32u, acc[1][1]);
        tmem_ld_wait();
        release_acc();
      }
      const uint32_t slot_addr = slot_base + slot * kSlotBytes + cg * (kSlotBytes / 2);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (kChunked) {         // block-final epilogues hold more per-channel state: one 32-row chunk in registers at a time
          tmem_ld16x256_x4(t_addr + static_cast<uint32_t>(32 * c), acc[c][0]);
          tmem_ld16x256_x4(t_addr + (16u << 16) + static_cast<uint32_t>(32 * c), acc[c][1]);
          tmem_ld_wait();
          if (c == 1) release_acc();
        }
        const uint32_t vm = vmask[c] >> (2 * tq), sm = smask[c] >> (2 * tq);
        const bool all_valid = vmask[c] == 0xFFFFFFFFu, all_sc = smask[c] == 0xFFFFFFFFu;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const uint32_t (&r)[16] = acc[c][s];
          // address of this thread's row in the two matrix quads (units 0,1 and 2,3)
          const int chunk = ((q & 1) * 4 + 2 * s + m_h) ^ m_i;
          const uint32_t addr0 = slot_addr + static_cast<uint32_t>((32 * c + 8 * m_u + m_i) * 128 + chunk * 16);
          const uint32_t addr1 = addr0 + 16u * 128u;
          __half2 x[8];                                       // x[2u + h]: rows 8u + 2 tq, +1 of channel (s, h)
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int h = 0; h < 2; ++h)
              x[2 * u + h] = cvt_sat_h2(__uint_as_float(r[4 * u + 2 * h]) + sh1[s][h], __uint_as_float(r[4 * u + 2 * h + 1]) + sh1[s][h]);
          if (kSc) {
            uint32_t sc[8];
            ldmatrix_x4_trans(addr0, sc[0], sc[1], sc[2], sc[3]);
            ldmatrix_x4_trans(addr1, sc[4], sc[5], sc[6], sc[7]);
            if (!all_sc) {                                     // rows the shortcut tensor's own mask zeroed carry a per-channel constant
#pragma unroll
              for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint32_t pm = pair_mask(sm >> (8 * u));
                  sc[2 * u + h] = (sc[2 * u + h] & pm) | (h2u(scc[s][h]) & ~pm);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __hadd2(x[i], u2h(sc[i]));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = gelu_tanh_h2(x[i]);
          if (kFinal) {
            // NMD tap on the block output: running per-thread sums over the valid rows (nmd.py:52-77)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              __half2 t0 = x[h], t1 = x[2 + h], t2 = x[4 + h], t3 = x[6 + h];
              if (!all_valid) {
                t0 = u2h(h2u(t0) & pair_mask(vm)); t1 = u2h(h2u(t1) & pair_mask(vm >> 8));
                t2 = u2h(h2u(t2) & pair_mask(vm >> 16)); t3 = u2h(h2u(t3) & pair_mask(vm >> 24));
              }
              const float2 f = __half22float2(__hadd2(__hadd2(t0, t1), __hadd2(t2, t3)));
              tapacc[s][h] += f.x + f.y;
            }
            // stand-alone norm + activation after the stack
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int h = 0; h < 2; ++h) x[2 * u + h] = gelu_tanh_h2(__hfma2(x[2 * u + h], a2[s][h], b2[s][h]));
            if (kPool) {
              const uint32_t ninf = 0xFC00FC00u;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                __half2 t0 = x[h], t1 = x[2 + h], t2 = x[4 + h], t3 = x[6 + h];
                if (!all_valid) {
                  uint32_t pm;
                  pm = pair_mask(vm); t0 = u2h((h2u(t0) & pm) | (ninf & ~pm));
                  pm = pair_mask(vm >> 8); t1 = u2h((h2u(t1) & pm) | (ninf & ~pm));
                  pm = pair_mask(vm >> 16); t2 = u2h((h2u(t2) & pm) | (ninf & ~pm));
                  pm = pair_mask(vm >> 24); t3 = u2h((h2u(t3) & pm) | (ninf & ~pm));
                }
                const float2 f = __half22float2(__hmax2(__hmax2(t0, t1), __hmax2(t2, t3)));
                poolacc[s][h] = fmaxf(poolacc[s][h], fmaxf(f.x, f.y));
              }
            }
          }
          if (p.y) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = h2u(x[i]);
            if (!all_valid) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint32_t pm = pair_mask(vm >> (8 * u));
                o[2 * u] &= pm; o[2 * u + 1] &= pm;
              }
            }
            stmatrix_x4_trans(addr0, o[0], o[1], o[2], o[3]);
            stmatrix_x4_trans(addr1, o[4], o[5], o[6], o[7]);
          }
        }
      }
      // both warps of the channel group have written (or, without an output tensor, read) the slot
      if (p.y) fence_async_smem();
      const long long tb0 = dbgw ? clock64() : 0;
      named_bar_sync(bar_id, 64);
      if (dbgw) { e_b += clock64() - tb0; e_m += tb0 - ts2; }
      if (issuer_warp && lane == (k_it & 3)) {
        if (p.y) {
          bulk_s2g(p.y + (static_cast<long long>(cg) * p.y_plane + row0) * 64, slot_addr, kSlotBytes / 2);
          bulk_commit();
        } else {
          mbar_arrive(SFREE(slot));
        }
      }
    }
    if (kFinal && cur_win >= 0) flush(cur_win);
    if (issuer_warp && lane < 4 && p.y) bulk_wait_all();
    if (dbg0 && q == 0 && lane == 0 && grp == 0) {
      p.dbg[3] = e_t; p.dbg[5] = e_s; p.dbg[7] = e_v; p.dbg[9] = e_m; p.dbg[11] = e_b;
      p.dbg[14] = clock64() - p.dbg[0];
    }
  }

  // ---- teardown ------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (dbg0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.dbg[0] = clock64() - p.dbg[0]; p.dbg[16] = static_cast<long long>(gt) - p.dbg[16];
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace ws
}  // namespace jg
