// Window-resident conv stack for NARROW models (every activation tensor <= 32 channels; BASELINE config 3: the 500 bp / 32-filter
// baseline, train_config/nn_config_500bp_baseline.yaml; graph = nnlib/builder.py:982-1193 over nnlib/v2/layers.py:1110-1265,
// 1836-1915).
//
// The per-layer kernels stream every activation tensor through HBM with its 32 channels zero-padded to a 64-channel group; at
// 32 channels a whole window (6 frames x ~166 codons = 1 024 rows x 64 B) is only 64 KB per tensor, so here ONE CTA keeps the
// window in shared memory through ALL layers of the stack:
//
//   tokens (1 KB per window, the only HBM read)  ->  one-hot stem operand built per 128-row tile in shared memory
//   ->  stem conv  ->  conv1 / conv2 of every residual block (shortcut added in place)  ->  trailing norm + activation
//   ->  masked sum / max pool  ->  32 floats per window (the only HBM write)
//
// Layout of an activation tensor in shared memory: fp16 [4 channel chunks of 8][guard + rows + guard][8] ("chunk planes"): the
// canonical K-major NO-SWIZZLE operand layout of tcgen05.mma with core matrices of consecutive row groups contiguous (SBO = 128 B)
// and the two K chunks of one MMA a plane apart (LBO = plane bytes).  Rows are linear inside a plane, so a conv tap is the
// descriptor's start address advanced by `shift` x 16 B, an MMA operand fetch reads 8 rows x 16 B = 128 contiguous bytes, and the
// epilogue's thread-per-row 16-byte accesses are conflict-free.  Guard rows stay zero = the SAME padding at the window ends; the
// gap rows between frames are stored as zeros by the epilogue like in the HBM layout (conv_common.cuh).
//
// MMA: M = 128 rows (A, shared memory), N = 32 output channels (B = the layer's weights, resident in shared memory for the CTA's
// lifetime), K = 16, fp32 accumulators of all 8 tiles of a layer in tensor memory (8 x 32 columns), two sets alternating between
// layers.  Dependencies are per 128-row tile, not per layer: the MMAs of layer l+1, tile i wait for the epilogues of layer l, tiles
// i-1 .. i+1 (mbarrier TILE_DONE), so the tensor pipe and the epilogue warps overlap across the layer boundary, and the next
// window's stem overlaps the last layer's epilogue.
//
// Roles (608 threads, one persistent CTA per SM): warps 0-15 epilogue (group g = warp / 4 owns tiles g and g + 4; thread = row, the
// same two rows of the window for every layer, so frame / position are computed once), warp 16 MMA issuer (+ TMEM allocation),
// warps 17 / 18 builders (tokens -> token mask + one-hot tiles of their tile parity; builder 0 prefetches the next window's tokens
// into registers and prepares its mask while the current window runs).
// The epilogue arithmetic is the generic epilogue of conv_epilogue.cuh (fp32 affine, packed fp16 from there on), so the numerics
// equal the per-layer kernels'.
#pragma once
#include "conv_ws.cuh"

namespace jg {
namespace rs {

using namespace jg::tc;
using jg::tc2::fence_async_smem;
using jg::ws::h2u;
using jg::ws::named_bar_sync;

constexpr int kMaxLayersRs = 12;
constexpr int kMaxTapsRs = 8;
constexpr int kGuardRs = 8;                              // zero rows before / after a window's rows in every chunk plane
constexpr int kEpiWarpsRs = 16;
constexpr int kBuildersRs = 2;                           // builder warps: stem tile i is built by builder i % 2
constexpr int kThreadsRs = (kEpiWarpsRs + 1 + kBuildersRs) * 32;
constexpr int kMaxTilesRs = 8;                           // 8 tiles x 32 columns x 2 sets = 512 TMEM columns
constexpr int kParBytesRs = 640;                         // per layer: shift1 f32[32] | scale1 f32[32] | scale2 h[32] | shift2 h[32] | scc h[32] | pad | bias f32[32] @ 512
constexpr int kTokWordsRs = 12;                          // token words a builder lane prefetches: frames x pitch <= 32 x 12 x 4 bytes

struct LayerRs {
  int ntaps;
  int shifts[kMaxTapsRs];
  int in_arr;              // 0 = tokens (stem: one-hot operand), 1 = X, 2 = H
  int out_arr;             // 1 = X, 2 = H, 0 = none (pool-only last layer)
  int has_sc;              // residual shortcut
  int sc_arr;              // the buffer it is read from (1 = X, 2 = H): the output buffer's current content (added in place), or X / H
                           // for the pool-only last layer
  int sc_all_valid;        // the shortcut tensor carries no mask
  int act1, act2, has_aff2, pool_mode, masking, folded;
  int shrink_in, shrink;
  int tap_mode;            // NMD tap taken by this kernel: 0 none, 1 raw conv output (acc + bias), 2 after the first activation, 3 on the
                           // launch output (nmd.py:52-77); a stem whose tap is taken from token counts (stem_tap_kernel) has 0 here
  int tap_slot;
  int count_id;            // >= 0: add the layer's valid rows to count[count_id][window] (layers with a tap or the pool)
  int mode;                // EpiModeRs: the compile-time epilogue shape this layer matches (0 = generic)
  int zero_tap;            // one of the taps has shift 0: in a window without masked codons every in-frame output row is then valid
  int kc;                  // input channel chunks of 8: 8 for the one-hot stem operand, 4 otherwise
  uint32_t w_off;          // byte offset of the layer's weight image in the weights block: [tap][kc][32 out channels][8] fp16
  // byte offsets from the (aligned) shared-memory base, filled by the launcher from smem_rs() so that the epilogue warps' per-layer
  // setup is a handful of independent constant loads; kNoneRs = absent.  e_mask_in of the stem is that of even windows
  // (odd windows: + 3 mask arrays).
  uint32_t e_par, e_mask_in, e_mask_out, e_out, e_sc, e_sc_mask;
  // the MMA warp's per-layer constants, also from the launcher: byte offset of the A operand of tile 0 / tap 0 (guard rows and the
  // first tap's shift included; stem: of staging slot 0), its chunk-plane pitch, byte offset of the weight image, tap-to-tap row
  // step, and which unrolled issue routine fits (0 generic, 1 = 3 taps x 2 K pairs, 2 = 5 x 2, 3 = stem 7 x 4)
  uint32_t i_a_off, i_plane, i_b_off, i_dil, i_shape;
};
constexpr uint32_t kNoneRs = 0xFFFFFFFFu;

struct ResidentParams {
  const uint8_t* tokens;   // [n_windows][frames][pitch]
  const int* lpad;         // [n_windows] padded frame length
  long long n_windows;
  int lc, pitch, tok_offset, period, frames, rpw, n_layers;
  const uint8_t* wblock;   // all weight images, then kParBytesRs per layer
  uint32_t w_bytes;
  float* pool;             // [n_windows][pool_pitch], pre-filled (0 for the sum, -1e9 for the max)
  int pool_pitch;
  int* count;              // [n_masks][cap_windows] valid rows per mask and window, pre-zeroed
  long long cap_windows;
  float* tap_sum;          // [n_taps][n_windows][tap_width] masked column sums, pre-zeroed
  int tap_width;
  int* err;
  long long* dbg;          // probe only (JG_RS_TRACE=1): clock64 timeline of CTA 0's fourth window, [layer][tile][4] + [1024..) misc
  LayerRs layer[kMaxLayersRs];
};

struct SmemRs {
  uint32_t plane_bytes, buf_bytes, buf_off[2], w_off, par_off, mask_bytes, mask_off[4], tok_off, tok_bytes, bar_off, total;
  uint32_t slot_rows, slot_plane, slot_bytes, n_slots;
};

__host__ __device__ inline SmemRs smem_rs(int rpw, int n_layers, uint32_t w_bytes, int frames, int pitch, int stem_span) {
  SmemRs S;
  S.plane_bytes = static_cast<uint32_t>(rpw + 2 * kGuardRs) * 16u;
  S.buf_bytes = 4u * S.plane_bytes;
  S.buf_off[0] = 0;
  S.buf_off[1] = S.buf_bytes;
  S.w_off = 2u * S.buf_bytes;
  S.par_off = S.w_off + ((w_bytes + 127u) & ~127u);
  S.mask_bytes = (static_cast<uint32_t>(rpw + 2 * kGuardRs) + 15u) & ~15u;
  S.mask_off[0] = S.par_off + static_cast<uint32_t>(n_layers) * kParBytesRs;
  S.mask_off[1] = S.mask_off[0] + S.mask_bytes;
  S.mask_off[2] = S.mask_off[1] + S.mask_bytes;
  S.mask_off[3] = S.mask_off[2] + S.mask_bytes;             // token masks are double-buffered by window parity: arrays 0 and 3
  S.tok_off = S.mask_off[3] + S.mask_bytes;
  S.tok_bytes = (static_cast<uint32_t>(frames * pitch) + 15u) & ~15u;
  S.bar_off = S.tok_off + 2u * S.tok_bytes;
  S.total = S.bar_off + 256u + static_cast<uint32_t>(kMaxLayersRs) * 32u + 128u;   // barriers + TMEM pointer, the MMA warp's per-layer
                                                                                    // table, + slack to align the base to 128 B
  S.slot_rows = static_cast<uint32_t>(128 + stem_span + 7) / 8u * 8u;
  S.slot_plane = S.slot_rows * 16u;
  S.slot_bytes = 8u * S.slot_plane;
  S.n_slots = S.buf_bytes / S.slot_bytes >= 3u ? 3u : 2u;    // one-hot stem tiles staged in H
  return S;
}

// The launcher's half of the per-layer setup: shared-memory byte offsets of everything a layer's epilogue touches.
inline void fill_layer_offsets(ResidentParams& p, const SmemRs& S) {
  for (int l = 0; l < p.n_layers; ++l) {
    LayerRs& L = p.layer[l];
    auto mask = [&](int arr) { return S.mask_off[0] + static_cast<uint32_t>(arr) * S.mask_bytes + kGuardRs; };
    L.e_par = S.par_off + static_cast<uint32_t>(l) * kParBytesRs;
    L.e_mask_in = mask(L.in_arr);
    L.e_mask_out = L.out_arr ? mask(L.out_arr) : kNoneRs;
    L.e_out = L.out_arr ? static_cast<uint32_t>(L.out_arr - 1) * S.buf_bytes : kNoneRs;
    L.e_sc = L.sc_arr ? static_cast<uint32_t>(L.sc_arr - 1) * S.buf_bytes : kNoneRs;
    L.e_sc_mask = (L.sc_arr && !L.sc_all_valid) ? mask(L.sc_arr) : kNoneRs;
    const bool stem = L.in_arr == 0;
    int smin = 0;
    for (int t = 0; t < L.ntaps; ++t) smin = L.shifts[t] < smin ? L.shifts[t] : smin;
    const int rs0 = stem ? L.shifts[0] - smin : L.shifts[0];
    L.i_plane = stem ? S.slot_plane : S.plane_bytes;
    L.i_a_off = (stem ? S.buf_off[1] : static_cast<uint32_t>(L.in_arr - 1) * S.buf_bytes + kGuardRs * 16u) + static_cast<uint32_t>(rs0 * 16);
    L.i_b_off = S.w_off + L.w_off;
    L.i_dil = L.ntaps > 1 ? static_cast<uint32_t>(L.shifts[1] - L.shifts[0]) : 0u;      // taps are equally spaced
    L.i_shape = stem ? ((L.ntaps == 7 && L.kc == 8) ? 3u : 0u) : (L.ntaps == 3 && L.kc == 4) ? 1u : (L.ntaps == 5 && L.kc == 4) ? 2u : 0u;
  }
}

// K-major no-swizzle matrix descriptor (cute/arch/mma_sm100_desc.hpp; canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units)
__device__ __forceinline__ uint64_t desc_ns(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return desc_pack(lo, hi);
}

// tcgen05.mma issue, SS mode.  No "memory" clobber: the instruction touches no memory the compiler knows about, and its ordering
// against the barrier waits / commits around it is that of the volatile asm statements themselves.
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
// The MMAs of one 128-row tile, fully unrolled for the layer shapes that carry the time (the rolled version spends ~180 cycles per
// MMA in a dependent chain of uniform-datapath address arithmetic; an MMA is ~40).  a_lo / b_lo: low descriptor words (start
// address and LBO) of the tile's first row, tap 0, K chunk 0; `dil`: tap-to-tap row shift, `a_kk`: two chunk planes, both in
// 16-byte units.  The weight image is contiguous in (tap, K chunk pair): 1 KB per MMA.
template <int kT, int kK>
__device__ __forceinline__ void issue_tile(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t dil, uint32_t a_kk, uint32_t hi, uint32_t idesc) {
#pragma unroll
  for (int t = 0; t < kT; ++t) {
#pragma unroll
    for (int kk = 0; kk < kK; ++kk)
      umma_ss(d, desc_pack(a_lo + static_cast<uint32_t>(t) * dil + static_cast<uint32_t>(kk) * a_kk, hi),
              desc_pack(b_lo + static_cast<uint32_t>(t * kK + kk) * 64u, hi), idesc, (t | kk) != 0);
  }
}

// barrier slots (8 bytes each) behind SmemRs::bar_off
constexpr uint32_t kBarAccFull = 0, kBarTileDone = 8, kBarOhFull = 16, kBarOhFree = 19, kBarWinDone = 22;

struct LayerIssue {
  long long* dbg;
  uint32_t idesc, hi, bar0, slot_step, n_slots;
  uint32_t a_lo, b_lo, dil, a_kk, d0, prev;
  int n_tiles, ntaps, kk;
  bool wait_prev, last;
};

// Epilogue shapes specialised at compile time (the runtime-flag version spends most of its issue slots on flag loads and branches):
//   LIGHT      folded affine -> tanh-GELU -> store                        (stem, conv1 of a residual block)
//   LIGHT_SC   folded affine -> + shortcut -> tanh-GELU -> store          (conv2 of a residual block)
//   FINAL_SUM  folded affine -> + shortcut -> tanh-GELU -> second affine -> tanh-GELU -> masked sum pool, no store
//   GENERIC    every feature behind its runtime flag
enum EpiModeRs { EPI_RS_GENERIC = 0, EPI_RS_LIGHT = 1, EPI_RS_LIGHT_SC = 2, EPI_RS_FINAL_SUM = 3 };

struct EpiLayer {
  const uint8_t* par;
  const uint8_t* mask_in;
  uint8_t* mask_out;
  uint8_t* out;
  const uint8_t* scb;
  const uint8_t* sc_mask;      // nullptr: the shortcut tensor carries no mask
  float* pool;
  float* tap;                  // this window's row of the layer's tap slot, at the thread's lane (nullptr: no tap)
  int* count;                  // this window's counter of the layer's mask (nullptr: nobody needs it)
  uint32_t plane_bytes, acc, parity, bar0;
  long long* dbg;              // trace slot of this layer (nullptr = off)
  int limit, frames, rpw, n_tiles;
  bool slow_mask;              // masked codons in this window (or no tap at shift 0): evaluate the "any" rule per row
};

// The clean-window flag is the builder's; it is published by the barrier chain that ends in the window's first ACC_FULL, which this
// thread has not waited for yet at the top of the stem layer -- so the stem layer waits for its first tile's accumulator here (the
// wait is repeated, already satisfied, inside epi_layer).
__device__ __forceinline__ bool s_clean_of(volatile int* s_clean, uint32_t itw, uint32_t bar0, int g, uint32_t gl) {
  mbar_wait(bar0 + (kBarAccFull + static_cast<uint32_t>(g)) * 8u, gl & 1u);
  return s_clean[itw & 1u] != 0;
}

// runtime-selected activation of the generic epilogue shape: tanh-GELU, ReLU or none (the launcher admits nothing else), without the
// erf-GELU branch of act_apply_h2 -- the kernel's code size matters, the roles and epilogue shapes share one instruction cache
__device__ __forceinline__ void act_rs(__half2 (&h)[16], int act) {
  if (act == ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = gelu_tanh_h2(h[j]);
  } else if (act == ACT_RELU) {
    const __half2 z = __float2half2_rn(0.0f);
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = __hmax2(h[j], z);
  }
}

template <int kMode>
__device__ __forceinline__ void epi_layer(const EpiLayer& E, const LayerRs& L, int g, int lane, int rw0, int fr0, int jj0, int rw1, int fr1, int jj1) {
  constexpr bool kGen = kMode == EPI_RS_GENERIC;
  const bool folded = kGen ? L.folded != 0 : true;
  const bool has_sc = kGen ? L.has_sc != 0 : (kMode == EPI_RS_LIGHT_SC || kMode == EPI_RS_FINAL_SUM);
  const bool has_aff2 = kGen ? L.has_aff2 != 0 : kMode == EPI_RS_FINAL_SUM;
  const int pool_mode = kGen ? L.pool_mode : (kMode == EPI_RS_FINAL_SUM ? 2 : 0);
  const bool has_out = kGen ? E.out != nullptr : kMode != EPI_RS_FINAL_SUM;
  const int act1 = kGen ? L.act1 : ACT_GELU_TANH, act2 = kGen ? L.act2 : ACT_GELU_TANH;
  const int tap_mode = kGen ? L.tap_mode : 0;
  const float4* bias = reinterpret_cast<const float4*>(E.par + 512);
  const float4* shift1 = reinterpret_cast<const float4*>(E.par);
  const float4* scale1 = reinterpret_cast<const float4*>(E.par + 128);
  const uint4* scale2 = reinterpret_cast<const uint4*>(E.par + 256);
  const uint4* shift2 = reinterpret_cast<const uint4*>(E.par + 320);
  const uint4* scc = reinterpret_cast<const uint4*>(E.par + 384);
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const int i = g + 4 * h;
    if (i >= E.n_tiles) break;
    mbar_wait(E.bar0 + (kBarAccFull + static_cast<uint32_t>(i)) * 8u, E.parity);
    if (E.dbg && (threadIdx.x & 127) == 0) E.dbg[i * 4 + 2] = clock64();
    tc_fence_after();
    uint32_t raw[32];
    tmem_ld32(E.acc + static_cast<uint32_t>(i) * 32u, raw);
    const int r = h ? rw1 : rw0;
    bool valid = (h ? fr1 : fr0) < E.frames && (h ? jj1 : jj0) < E.limit;
    // Keras mask propagation, mode "any" (layers.py:1245-1252).  In a window without masked codons every in-frame row of every layer
    // is valid as soon as the layer has a tap at shift 0, so the tap loop only runs for windows with N runs.
    if (E.slow_mask) {
      uint32_t any = 0u;
      for (int t = 0; t < L.ntaps; ++t) any |= E.mask_in[r + L.shifts[t]];
      valid = valid && any != 0u;
    }
    const uint32_t row_off = static_cast<uint32_t>(kGuardRs + r) * 16u;
    const bool sc_valid = has_sc && (E.sc_mask == nullptr || E.sc_mask[r] != 0);
    tmem_ld_wait();
    tc_fence_before();
    if (kGen && tap_mode == 1) {      // NMD tap on the raw conv output
      float tv[32];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 b = bias[j4];
        tv[j4 * 4 + 0] = valid ? __uint_as_float(raw[j4 * 4 + 0]) + b.x : 0.0f;
        tv[j4 * 4 + 1] = valid ? __uint_as_float(raw[j4 * 4 + 1]) + b.y : 0.0f;
        tv[j4 * 4 + 2] = valid ? __uint_as_float(raw[j4 * 4 + 2]) + b.z : 0.0f;
        tv[j4 * 4 + 3] = valid ? __uint_as_float(raw[j4 * 4 + 3]) + b.w : 0.0f;
      }
      warp_cols_reduce<false>(tv, lane);
      atomicAdd(E.tap, tv[0]);
    }
    __half2 hv[16];
    if (folded) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 b = shift1[j4];
        hv[j4 * 2 + 0] = cvt_sat_h2(__uint_as_float(raw[j4 * 4 + 0]) + b.x, __uint_as_float(raw[j4 * 4 + 1]) + b.y);
        hv[j4 * 2 + 1] = cvt_sat_h2(__uint_as_float(raw[j4 * 4 + 2]) + b.z, __uint_as_float(raw[j4 * 4 + 3]) + b.w);
      }
    } else {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 a = scale1[j4], b = shift1[j4];
        hv[j4 * 2 + 0] = cvt_sat_h2(fmaf(__uint_as_float(raw[j4 * 4 + 0]), a.x, b.x), fmaf(__uint_as_float(raw[j4 * 4 + 1]), a.y, b.y));
        hv[j4 * 2 + 1] = cvt_sat_h2(fmaf(__uint_as_float(raw[j4 * 4 + 2]), a.z, b.z), fmaf(__uint_as_float(raw[j4 * 4 + 3]), a.w, b.w));
      }
    }
    if (has_sc) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 sv = sc_valid ? *reinterpret_cast<const uint4*>(E.scb + c * E.plane_bytes + row_off) : scc[c];
        const __half2* s2 = reinterpret_cast<const __half2*>(&sv);
#pragma unroll
        for (int k = 0; k < 4; ++k) hv[c * 4 + k] = __hadd2(hv[c * 4 + k], s2[k]);
      }
    }
    if (kGen) act_rs(hv, act1); else act_rs(hv, ACT_GELU_TANH);
    if (kGen && tap_mode == 2) {      // NMD tap after the first activation
      __half2 tv[16];
      const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
      for (int k = 0; k < 16; ++k) tv[k] = valid ? hv[k] : zero;
      atomicAdd(E.tap, warp_cols_reduce_h2<false>(tv, lane));
    }
    if (has_aff2) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 a = scale2[c], b = shift2[c];
        const __half2* a2 = reinterpret_cast<const __half2*>(&a);
        const __half2* b2 = reinterpret_cast<const __half2*>(&b);
#pragma unroll
        for (int k = 0; k < 4; ++k) hv[c * 4 + k] = __hfma2(hv[c * 4 + k], a2[k], b2[k]);
      }
      if (kGen) act_rs(hv, act2); else act_rs(hv, ACT_GELU_TANH);
    }
    if (kGen && tap_mode == 3) {      // NMD tap on the launch output
      __half2 tv[16];
      const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
      for (int k = 0; k < 16; ++k) tv[k] = valid ? hv[k] : zero;
      atomicAdd(E.tap, warp_cols_reduce_h2<false>(tv, lane));
    }
    if (E.count != nullptr) {
      const unsigned bal = __ballot_sync(0xffffffffu, valid);
      if (lane == 0 && bal) atomicAdd(E.count, __popc(bal));
    }
    if (pool_mode != 0) {
      __half2 tv[16];
      const bool pool_max = pool_mode == 1;
      const uint32_t fill_bits = pool_max ? 0xFC00FC00u : 0u;
      const __half2 fill = *reinterpret_cast<const __half2*>(&fill_bits);
#pragma unroll
      for (int k = 0; k < 16; ++k) tv[k] = valid ? hv[k] : fill;
      if (pool_max) {
        const float m = warp_cols_reduce_h2<true>(tv, lane);
        if (m > -1.0e38f) atomic_max_f32(E.pool, m);
      } else {
        atomicAdd(E.pool, warp_cols_reduce_h2<false>(tv, lane));
      }
    }
    if (has_out) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 o;
        o.x = h2u(hv[c * 4 + 0]); o.y = h2u(hv[c * 4 + 1]); o.z = h2u(hv[c * 4 + 2]); o.w = h2u(hv[c * 4 + 3]);
        if (!valid) o = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(E.out + c * E.plane_bytes + row_off) = o;
      }
      E.mask_out[r] = static_cast<uint8_t>(valid);
      // the window's guard rows: H doubles as the staging area of the one-hot stem tiles, so they are re-zeroed by every layer
      if (r < kGuardRs || r >= E.rpw - kGuardRs) {
        const uint32_t goff = static_cast<uint32_t>(r < kGuardRs ? r : r + 2 * kGuardRs) * 16u;
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(E.out + c * E.plane_bytes + goff) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(E.bar0 + (kBarTileDone + static_cast<uint32_t>(i)) * 8u);
    if (E.dbg && (threadIdx.x & 127) == 0) E.dbg[i * 4 + 3] = clock64();
  }
}

// The MMA warp's work for one layer of one window: per 128-row tile the barrier waits, the tile's MMAs and the commits.
// kT, kK > 0: taps / K-chunk pairs known at compile time (unrolled issue); 0: generic loops.
template <int kT, int kK, bool kStem>
__device__ __forceinline__ void layer_tiles(const LayerIssue& I, bool leader, uint32_t& oh_slot, uint32_t& oh_phase) {
  for (int i = 0; i < I.n_tiles; ++i) {
    // Every phase of every TILE_DONE barrier is awaited exactly once, in order (a parity wait cannot tell phase u from u + 2, so
    // no phase may be skipped -- the stem waits for the previous window's last layer too, although it only needs it for the
    // accumulator set it is about to overwrite).  Tile i needs tiles i-1 .. i+1 of the layer before; i-1 and i were awaited for
    // the tiles before this one.
    if (I.wait_prev) {
      if (i == 0) {
        mbar_wait(I.bar0 + (kBarTileDone + 0) * 8u, I.prev);
        if (I.n_tiles > 1) mbar_wait(I.bar0 + (kBarTileDone + 1) * 8u, I.prev);
      } else if (i + 1 < I.n_tiles) {
        mbar_wait(I.bar0 + (kBarTileDone + static_cast<uint32_t>(i) + 1u) * 8u, I.prev);
      }
    }
    uint32_t slot = 0;
    if (kStem) {
      slot = oh_slot;
      mbar_wait(I.bar0 + (kBarOhFull + slot) * 8u, oh_phase);
      if (++oh_slot == I.n_slots) { oh_slot = 0; oh_phase ^= 1u; }
    }
    tc_fence_after();
    if (I.dbg && (threadIdx.x & 31) == 0) I.dbg[i * 4 + 0] = clock64();
    const uint32_t d = I.d0 + static_cast<uint32_t>(i) * 32u;
    const uint32_t a_lo = I.a_lo + (kStem ? slot * I.slot_step : static_cast<uint32_t>(i) * (kTileM * 16u >> 4));
    if (leader) {
      if (kT > 0) {
        issue_tile<kT, kK>(d, a_lo, I.b_lo, I.dil, I.a_kk, I.hi, I.idesc);
      } else {
        for (int t = 0; t < I.ntaps; ++t)
          for (int kk = 0; kk < I.kk; ++kk)
            umma_ss(d, desc_pack(a_lo + static_cast<uint32_t>(t) * I.dil + static_cast<uint32_t>(kk) * I.a_kk, I.hi),
                    desc_pack(I.b_lo + static_cast<uint32_t>(t * I.kk + kk) * 64u, I.hi), I.idesc, (t | kk) != 0);
      }
      umma_commit(I.bar0 + (kBarAccFull + static_cast<uint32_t>(i)) * 8u);
      if (kStem) umma_commit(I.bar0 + (kBarOhFree + slot) * 8u);
      if (I.last && i == I.n_tiles - 1) umma_commit(I.bar0 + kBarWinDone * 8u);
    }
    __syncwarp();
    if (I.dbg && (threadIdx.x & 31) == 0) I.dbg[i * 4 + 1] = clock64();
  }
}

__global__ void __launch_bounds__(kThreadsRs, 1) stack_resident_kernel(const __grid_constant__ ResidentParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.rpw / kTileM;
  const int n_layers = p.n_layers;
  int stem_min = 0, stem_max = 0;
  for (int t = 0; t < p.layer[0].ntaps; ++t) {
    stem_min = min(stem_min, p.layer[0].shifts[t]);
    stem_max = max(stem_max, p.layer[0].shifts[t]);
  }
  const SmemRs S = smem_rs(p.rpw, n_layers, p.w_bytes, p.frames, p.pitch, stem_max - stem_min);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + S.bar_off);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 24);
  volatile int* s_clean = reinterpret_cast<volatile int*>(s_bar + 25);   // [2], by window parity: no in-frame codon of the window is masked
  auto ACC_FULL = [&](int i) { return smem_u32(s_bar + i); };
  auto TILE_DONE = [&](int i) { return smem_u32(s_bar + 8 + i); };
  auto OH_FULL = [&](int s) { return smem_u32(s_bar + kBarOhFull + s); };
  auto OH_FREE = [&](int s) { return smem_u32(s_bar + kBarOhFree + s); };
  const uint32_t WIN_DONE = smem_u32(s_bar + kBarWinDone);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxTilesRs; ++i) { mbar_init(ACC_FULL(i), 1); mbar_init(TILE_DONE(i), 4); }
    for (int s = 0; s < 3; ++s) { mbar_init(OH_FULL(s), 1); mbar_init(OH_FREE(s), 1); }
    mbar_init(WIN_DONE, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarpsRs) tmem_alloc(smem_u32(s_tmem), 512u);
  {   // zero the activation buffers and masks, bring in the weights and per-layer parameters
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = threadIdx.x; i < 2u * S.buf_bytes / 16u; i += kThreadsRs) z[i] = make_uint4(0u, 0u, 0u, 0u);
    uint4* mz = reinterpret_cast<uint4*>(smem + S.mask_off[0]);
    for (uint32_t i = threadIdx.x; i < 4u * S.mask_bytes / 16u; i += kThreadsRs) mz[i] = make_uint4(0u, 0u, 0u, 0u);
    const uint4* src = reinterpret_cast<const uint4*>(p.wblock);
    uint4* wd = reinterpret_cast<uint4*>(smem + S.w_off);
    for (uint32_t i = threadIdx.x; i < p.w_bytes / 16u; i += kThreadsRs) wd[i] = src[i];
    const uint4* psrc = reinterpret_cast<const uint4*>(p.wblock + p.w_bytes);
    uint4* pd = reinterpret_cast<uint4*>(smem + S.par_off);
    for (uint32_t i = threadIdx.x; i < static_cast<uint32_t>(n_layers) * kParBytesRs / 16u; i += kThreadsRs) pd[i] = psrc[i];
  }
  // the MMA warp's per-layer constants as a shared-memory table (two 16-byte loads per layer instead of a chain of indexed constant
  // loads: the gap between the last tile of a layer and the first of the next was ~800 cycles of this one warp's serial time)
  uint4* s_lay = reinterpret_cast<uint4*>(smem + S.bar_off + 256u);
  if (threadIdx.x < static_cast<unsigned>(n_layers)) {
    const LayerRs& L = p.layer[threadIdx.x];
    const uint32_t smem16 = smem_u32(smem) >> 4, plane16 = L.i_plane >> 4;
    // K-major no-swizzle descriptors: LBO (bits 16-29 of the low word) = distance of the two K chunks of an MMA = one chunk
    // plane (A) / 512 B (B); SBO (high word) = 128 B: consecutive 8-row core matrices are contiguous.
    s_lay[threadIdx.x * 2 + 0] = make_uint4(((smem16 + (L.i_a_off >> 4)) & 0x3FFFu) | ((plane16 & 0x3FFFu) << 16),
                                            ((smem16 + (L.i_b_off >> 4)) & 0x3FFFu) | ((512u >> 4) << 16), L.i_dil, 2u * plane16);
    s_lay[threadIdx.x * 2 + 1] = make_uint4(static_cast<uint32_t>(L.ntaps), static_cast<uint32_t>(L.kc / 2), L.i_shape, L.in_arr == 0 ? 1u : 0u);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const long long w0 = blockIdx.x, wstep = gridDim.x;

  if (warp < kEpiWarpsRs) {
    // ------------------------------------------------------------------ epilogue: thread = row ------------------------------
    const int q = warp & 3, g = warp >> 2;
    const int rw0 = g * kTileM + q * 32 + lane, rw1 = rw0 + 4 * kTileM;       // the two rows of the window this thread owns
    const int fr0 = rw0 / p.period, fr1 = rw1 / p.period;
    const int jj0 = rw0 - fr0 * p.period, jj1 = rw1 - fr1 * p.period;
    uint32_t gl = 0, itw = 0;
    int lp_next = w0 < p.n_windows ? p.lpad[w0] : 0;
    for (long long w = w0; w < p.n_windows && g < n_tiles; w += wstep, ++itw) {      // a group beyond the tile count has no rows
      const int lp = lp_next;
      if (w + wstep < p.n_windows) lp_next = p.lpad[w + wstep];
      bool win_clean = false;          // the builder's flag for this window, read once (after the stem's first accumulator is up)
      for (int l = 0; l < n_layers; ++l, ++gl) {
        const LayerRs& L = p.layer[l];
        EpiLayer E;
        E.par = smem + L.e_par;
        E.mask_in = smem + L.e_mask_in + ((l == 0 && (itw & 1u)) ? 3u * S.mask_bytes : 0u);
        E.mask_out = L.e_mask_out != kNoneRs ? smem + L.e_mask_out : nullptr;
        E.out = L.e_out != kNoneRs ? smem + L.e_out : nullptr;
        E.scb = L.e_sc != kNoneRs ? smem + L.e_sc : nullptr;
        E.sc_mask = L.e_sc_mask != kNoneRs ? smem + L.e_sc_mask : nullptr;
        E.limit = lp - L.shrink_in - L.shrink;
        E.plane_bytes = S.plane_bytes;
        E.acc = tmem + (static_cast<uint32_t>(q * 32) << 16) + (gl & 1u) * 256u;
        E.parity = gl & 1u;
        E.bar0 = smem_u32(s_bar);
        if (l == 0) win_clean = s_clean_of(s_clean, itw, smem_u32(s_bar), g, gl);
        E.slow_mask = L.masking && !(L.zero_tap && win_clean);
        E.pool = p.pool + w * p.pool_pitch + lane;
        E.count = L.count_id >= 0 ? p.count + static_cast<long long>(L.count_id) * p.cap_windows + w : nullptr;
        E.tap = L.tap_mode ? p.tap_sum + (static_cast<long long>(L.tap_slot) * p.n_windows + w) * p.tap_width + lane : nullptr;
        E.frames = p.frames;
        E.rpw = p.rpw;
        E.n_tiles = n_tiles;
        E.dbg = (p.dbg != nullptr && blockIdx.x == 0 && itw == 3) ? p.dbg + l * 32 : nullptr;
        switch (L.mode) {
          case EPI_RS_LIGHT: epi_layer<EPI_RS_LIGHT>(E, L, g, lane, rw0, fr0, jj0, rw1, fr1, jj1); break;
          case EPI_RS_LIGHT_SC: epi_layer<EPI_RS_LIGHT_SC>(E, L, g, lane, rw0, fr0, jj0, rw1, fr1, jj1); break;
          case EPI_RS_FINAL_SUM: epi_layer<EPI_RS_FINAL_SUM>(E, L, g, lane, rw0, fr0, jj0, rw1, fr1, jj1); break;
          default: epi_layer<EPI_RS_GENERIC>(E, L, g, lane, rw0, fr0, jj0, rw1, fr1, jj1); break;
        }
      }
    }
  } else if (warp == kEpiWarpsRs) {
    // ------------------------------------------------------------------ MMA issuer -------------------------------------------
    const bool leader = elect_one();
    LayerIssue I;
    I.idesc = (1u << 4) | (static_cast<uint32_t>(32 >> 3) << 17) | (static_cast<uint32_t>(kTileM >> 4) << 24);
    I.hi = (128u >> 4) | (1u << 14);
    I.n_tiles = n_tiles;
    I.bar0 = smem_u32(s_bar);
    I.slot_step = S.slot_bytes >> 4;
    I.n_slots = S.n_slots;
    uint32_t gl = 0, oh_slot = 0, oh_phase = 0;
    for (long long w = w0; w < p.n_windows; w += wstep) {
      for (int l = 0; l < n_layers; ++l, ++gl) {
        // Everything that is constant over the layer's tiles comes ready-made from the launcher (fill_layer_offsets): the per-layer
        // and per-tile code of this one warp is the serial section of the whole kernel (a layer's period = its eight tiles here).
        // K-major no-swizzle descriptors: LBO (bits 16-29 of the low word) = distance of the two K chunks of an MMA = one chunk
        // plane (A) / 512 B (B); SBO (high word) = 128 B: consecutive 8-row core matrices are contiguous.
        const uint4 q0 = s_lay[l * 2], q1 = s_lay[l * 2 + 1];
        const uint32_t shape = q1.z;
        const bool stem = q1.w != 0u;
        I.a_lo = q0.x;
        I.b_lo = q0.y;
        I.dil = q0.z;
        I.a_kk = q0.w;
        I.d0 = tmem + (gl & 1u) * 256u;
        I.prev = (gl - 1u) & 1u;
        I.wait_prev = gl > 0u;
        I.last = l == n_layers - 1;
        I.ntaps = static_cast<int>(q1.x);
        I.kk = static_cast<int>(q1.y);
        I.dbg = (p.dbg != nullptr && blockIdx.x == 0 && w == w0 + 3 * wstep) ? p.dbg + l * 32 : nullptr;
        if (shape == 1u) layer_tiles<3, 2, false>(I, leader, oh_slot, oh_phase);
        else if (shape == 3u) layer_tiles<7, 4, true>(I, leader, oh_slot, oh_phase);
        else if (shape == 2u) layer_tiles<5, 2, false>(I, leader, oh_slot, oh_phase);
        else if (stem) layer_tiles<0, 0, true>(I, leader, oh_slot, oh_phase);
        else layer_tiles<0, 0, false>(I, leader, oh_slot, oh_phase);
      }
    }
  } else {
    // ------------------------------------------------------------------ builder: tokens -> mask + one-hot stem tiles ----------
    const int n_words = p.frames * p.pitch / 4;
    uint32_t tk[kTokWordsRs];
    auto fetch = [&](long long w) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(p.tokens + w * p.frames * p.pitch);
#pragma unroll
      for (int u = 0; u < kTokWordsRs; ++u) tk[u] = (u * 32 + lane < n_words) ? src[u * 32 + lane] : 0u;
    };
    // tokens (from the prefetch registers) + token mask + clean flag of window number `itp` into the buffers of its parity; runs
    // while the window before it is still in its conv layers, so only the one-hot tiles are built after WIN_DONE
    auto prepare = [&](uint32_t itp, int lim) {
      uint8_t* tok = smem + S.tok_off + (itp & 1u) * S.tok_bytes;
      uint8_t* mask_t = smem + S.mask_off[0] + (itp & 1u) * 3u * S.mask_bytes + kGuardRs;
#pragma unroll
      for (int u = 0; u < kTokWordsRs; ++u)
        if (u * 32 + lane < n_words) reinterpret_cast<uint32_t*>(tok)[u * 32 + lane] = tk[u];
      __syncwarp();
      bool clean = true;
      int f = lane / p.period, j = lane - f * p.period;
      for (int r = lane; r < p.rpw; r += 32) {
        const bool in = f < p.frames && j < lim;
        const bool on = in && static_cast<int>(tok[f * p.pitch + j]) - p.tok_offset >= 0;
        mask_t[r] = static_cast<uint8_t>(on);
        clean = clean && (on || !in);
        j += 32;
        while (j >= p.period) { j -= p.period; ++f; }
      }
      clean = __all_sync(0xffffffffu, clean);
      if (lane == 0) s_clean[itp & 1u] = clean ? 1 : 0;     // published by the OH_FULL arrive of the window's first stem tile
    };
    // Two builder warps (one 128-row one-hot tile takes a single warp ~2 000 cycles, the stem's MMAs of a tile ~1 200): builder b
    // builds the tiles of parity b; builder 0 also prepares the next window.  Both walk the same slot ring over ALL tiles.
    const int bld = warp - (kEpiWarpsRs + 1);
    uint8_t* stage = smem + S.buf_off[1];
    uint32_t it = 0, slot = 0, sphase = 0;
    bool first_lap = true;
    if (bld == 0 && w0 < p.n_windows) {
      fetch(w0);
      const int lp0 = p.lpad[w0];
      prepare(0u, lp0 < p.lc ? lp0 : p.lc);
      if (w0 + wstep < p.n_windows) fetch(w0 + wstep);
    }
    for (long long w = w0; w < p.n_windows; w += wstep, ++it) {
      // builder 0 has written this window's tokens / mask / flag (at the end of its previous iteration); builder 1 reads the tokens
      named_bar_sync(1, 32 * kBuildersRs);
      const int lpw = p.lpad[w];
      const int lim = lpw < p.lc ? lpw : p.lc;
      const uint8_t* tok = smem + S.tok_off + (it & 1u) * S.tok_bytes;
      // the previous window's last MMAs (the last readers of H, where the one-hot tiles are staged) are done
      if (it > 0) mbar_wait(WIN_DONE, (it - 1u) & 1u);
      for (int i = 0; i < n_tiles; ++i) {
        if ((i & 1) != bld) {            // the other builder's tile: only advance the ring
          if (++slot == S.n_slots) { slot = 0; sphase ^= 1u; first_lap = false; }
          continue;
        }
        if (!first_lap) mbar_wait(OH_FREE(slot), sphase ^ 1u);
        uint8_t* sl = stage + slot * S.slot_bytes;
        int r = i * kTileM + stem_min + lane;
        int f = r >= 0 ? r / p.period : 0, j = r >= 0 ? r - f * p.period : r;      // j < 0: a row before the window
        for (int sr = lane; sr < static_cast<int>(S.slot_rows); sr += 32) {
          int ch = -1;
          if (j >= 0 && r < p.rpw && f < p.frames && j < lim) ch = static_cast<int>(tok[f * p.pitch + j]) - p.tok_offset;
          const uint32_t one = ch >= 0 ? (0x3C00u << (16 * (ch & 1))) : 0u;      // fp16 1.0 in the channel's half of its word
          const int word = ch >= 0 ? (ch & 7) >> 1 : -1, chunk = ch >= 0 ? ch >> 3 : -1;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (c == chunk) { o.x = word == 0 ? one : 0u; o.y = word == 1 ? one : 0u; o.z = word == 2 ? one : 0u; o.w = word == 3 ? one : 0u; }
            *reinterpret_cast<uint4*>(sl + c * S.slot_plane + sr * 16) = o;
          }
          r += 32; j += 32;
          while (j >= p.period) { j -= p.period; ++f; }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(OH_FULL(slot));
        if (++slot == S.n_slots) { slot = 0; sphase ^= 1u; first_lap = false; }
      }
      if (bld == 0 && w + wstep < p.n_windows) {       // the next window's tokens, mask and flag, while this one runs through its layers
        const int lpn = p.lpad[w + wstep];
        prepare(it + 1u, lpn < p.lc ? lpn : p.lc);
        if (w + 2 * wstep < p.n_windows) fetch(w + 2 * wstep);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarpsRs) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

}  // namespace rs
}  // namespace jg
