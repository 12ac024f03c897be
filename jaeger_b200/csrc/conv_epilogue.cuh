// The fused epilogue of one 32-channel batch of a conv tile, shared by the single-CTA and the
// CTA-pair kernels (thread = output row, lane = row within the warp's 32 rows).
//
//   acc (fp32, from TMEM) -> [NMD tap on acc + bias] -> fp32 affine (BatchNorm folded)
//   -> packed fp16 from here on: [+ shortcut] -> activation -> [NMD tap] -> [affine -> activation]
//   -> [masked global max / sum pool] -> 32 fp16 values, zeroed on masked rows
//
// Reference layers: nnlib/v2/layers.py:918-941 (BatchNorm inference), 1882-1915 (ResidualBlock),
// 27-29 (tanh GELU), 517-529 / 460-480 (masked pooling); nnlib/v2/nmd.py:52-77 (NMD).
#pragma once
#include "conv_common.cuh"

namespace jg {

// Column-wise reduction across the 32 lanes of a warp of a [32 lanes][32 columns] register
// tile.  Afterwards v[0] on lane l holds the reduction of column l.
template <bool kMax>
__device__ __forceinline__ void warp_cols_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      float send = hi ? v[j] : v[j + off];
      float keep = hi ? v[j + off] : v[j];
      float r = __shfl_xor_sync(0xffffffffu, send, off);
      v[j] = kMax ? fmaxf(keep, r) : (keep + r);
    }
  }
}

// The same reduction for a tile held as 16 packed fp16 pairs (h[i] = columns 2i, 2i+1).  The two
// widest exchange levels run on the packed registers (half the shuffles and selects; a max is
// exact in fp16, a sum rounds two of its 31 additions to fp16), the narrow ones in fp32.
// Returns the reduction of column `lane`.
template <bool kMax>
__device__ __forceinline__ float warp_cols_reduce_h2(const __half2 (&h)[16], int lane) {
  uint32_t v[8];
  {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t a = *reinterpret_cast<const uint32_t*>(&h[j]), b = *reinterpret_cast<const uint32_t*>(&h[j + 8]);
      const uint32_t r = __shfl_xor_sync(0xffffffffu, hi ? a : b, 16);
      const uint32_t k = hi ? b : a;
      const __half2 x = *reinterpret_cast<const __half2*>(&k), y = *reinterpret_cast<const __half2*>(&r);
      const __half2 z = kMax ? __hmax2(x, y) : __hadd2(x, y);
      v[j] = *reinterpret_cast<const uint32_t*>(&z);
    }
  }
  {
    const bool hi = (lane & 8) != 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t r = __shfl_xor_sync(0xffffffffu, hi ? v[j] : v[j + 4], 8);
      const uint32_t k = hi ? v[j + 4] : v[j];
      const __half2 x = *reinterpret_cast<const __half2*>(&k), y = *reinterpret_cast<const __half2*>(&r);
      const __half2 z = kMax ? __hmax2(x, y) : __hadd2(x, y);
      v[j] = *reinterpret_cast<const uint32_t*>(&z);
    }
  }
  float f[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&v[j]));
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
#pragma unroll
  for (int off = 4; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float r = __shfl_xor_sync(0xffffffffu, hi ? f[j] : f[j + off], off);
      const float k = hi ? f[j + off] : f[j];
      f[j] = kMax ? fmaxf(k, r) : (k + r);
    }
  }
  return f[0];
}

// Per-channel parameters staged in shared memory: fp32 for the stage that touches the fp32
// accumulator, fp16 pairs for the packed stages.
struct EpiParams {
  const float4* scale1;   // [cout/4]
  const float4* shift1;
  const float4* bias;
  const uint4* scale2;    // [cout/8] 8 halves each
  const uint4* shift2;
  const uint4* scc;       // shortcut value at masked rows
  const uint4* g1;        // MaskedDYT gamma / beta of the first and second norm (fp16, only read when p.dyt1 / p.dyt2)
  const uint4* b1;
  const uint4* g2;
  const uint4* b2;
};
constexpr int kEpiParFloats = 10;   // shared-memory parameter block: kEpiParFloats * cout floats

__device__ __forceinline__ EpiParams epi_params(float* s_par, int cout) {
  EpiParams e;
  e.scale1 = reinterpret_cast<const float4*>(s_par);
  e.shift1 = reinterpret_cast<const float4*>(s_par + cout);
  e.bias = reinterpret_cast<const float4*>(s_par + 2 * cout);
  e.scale2 = reinterpret_cast<const uint4*>(s_par + 3 * cout);
  e.shift2 = reinterpret_cast<const uint4*>(s_par + 4 * cout);
  e.scc = reinterpret_cast<const uint4*>(s_par + 5 * cout);
  e.g1 = reinterpret_cast<const uint4*>(s_par + 6 * cout);
  e.b1 = reinterpret_cast<const uint4*>(s_par + 7 * cout);
  e.g2 = reinterpret_cast<const uint4*>(s_par + 8 * cout);
  e.b2 = reinterpret_cast<const uint4*>(s_par + 9 * cout);
  return e;
}

// cooperative fill of the parameter block (all threads of the CTA)
__device__ __forceinline__ void epi_params_fill(float* s_par, const ConvParams& p, int tid, int nthreads) {
  __half* h2 = reinterpret_cast<__half*>(s_par + 3 * p.cout);
  __half* t2 = reinterpret_cast<__half*>(s_par + 4 * p.cout);
  __half* sc = reinterpret_cast<__half*>(s_par + 5 * p.cout);
  for (int i = tid; i < p.cout; i += nthreads) {
    s_par[i] = p.scale1[i];
    s_par[p.cout + i] = p.shift1[i];
    s_par[2 * p.cout + i] = p.bias ? p.bias[i] : 0.0f;
    if (p.epi_f32) {        // the fp32 epilogue keeps the second affine and the shortcut constant in fp32 (same slots, full width)
      s_par[3 * p.cout + i] = p.has_affine2 ? p.scale2[i] : 1.0f;
      s_par[4 * p.cout + i] = p.has_affine2 ? p.shift2[i] : 0.0f;
      s_par[5 * p.cout + i] = p.sc_const ? p.sc_const[i] : 0.0f;
      continue;
    }
    h2[i] = __float2half_rn(p.has_affine2 ? p.scale2[i] : 1.0f);
    t2[i] = __float2half_rn(p.has_affine2 ? p.shift2[i] : 0.0f);
    sc[i] = __float2half_rn(p.sc_const ? p.sc_const[i] : 0.0f);
    if (p.dyt1) {
      reinterpret_cast<__half*>(s_par + 6 * p.cout)[i] = __float2half_rn(p.dyt_g1[i]);
      reinterpret_cast<__half*>(s_par + 7 * p.cout)[i] = __float2half_rn(p.dyt_b1[i]);
    }
    if (p.dyt2) {
      reinterpret_cast<__half*>(s_par + 8 * p.cout)[i] = __float2half_rn(p.dyt_g2[i]);
      reinterpret_cast<__half*>(s_par + 9 * p.cout)[i] = __float2half_rn(p.dyt_b2[i]);
    }
  }
}

// Row validity travels through shared memory: a helper warp (otherwise idle) evaluates it a few
// tiles ahead of the epilogue warps, so the global loads behind it (window length, input masks,
// shortcut mask) never stall the warps the kernel is bound by.  bit 0: output row valid,
// bit 1: shortcut row valid.  Ring of kVSlots tiles x 128 rows, one byte per row.
constexpr int kVSlots = 4;
// One warp evaluates the 128 rows of the tile starting at tile_row0 (lane = row within each of the
// four 32-row chunks) and writes the codes to dst[128].  Every global load of the tile is issued
// before the first dependent instruction, so the tile costs one memory round trip, not one per
// chunk and tap.  With fuse_mask this is also the mask propagation of the layer: it publishes the
// output mask (one byte per row) and adds the valid rows to the window's counter.
// The mask buffers carry kGuardRows bytes on either side, so row + shift is always in bounds.
__device__ __forceinline__ void tile_validity(const ConvParams& p, long long tile_row0, int lane, volatile uint8_t* dst) {
  const int win = static_cast<int>(tile_row0 / p.rows_per_window);
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
          for (int c = 0; c < 4; ++c) m[c][u] = on ? p.in_mask[tile_row0 + c * 32 + lane + sh] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int c = 0; c < 4; ++c) any[c] += m[c][u];         // mask bytes are 0 / 1: the number of valid taps
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) any[c] = p.out_mask[tile_row0 + c * 32 + lane];
  }
  const bool has_scm = p.sc != nullptr && p.sc_mask != nullptr;
#pragma unroll
  for (int c = 0; c < 4; ++c) scm[c] = has_scm ? p.sc_mask[tile_row0 + c * 32 + lane] : (p.sc != nullptr ? 1u : 0u);
  const int limit = ((lp - p.shrink_in + p.len_round) >> p.halvings) - p.shrink;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const long long row = tile_row0 + c * 32 + lane;
    bool ok;
    if (p.fuse_mask) {
      const int rw = static_cast<int>(row - static_cast<long long>(win) * p.rows_per_window);
      const int f = rw / p.period, j = rw - f * p.period;
      ok = f < p.frames && j < limit && (!p.masking || any[c] >= static_cast<uint32_t>(p.mask_thr > 1 ? p.mask_thr : 1));
      p.out_mask_w[row] = static_cast<uint8_t>(ok);
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (lane == 0 && bal) atomicAdd(p.count + win, __popc(bal));
    } else {
      ok = any[c] != 0u;
    }
    dst[c * 32 + lane] = static_cast<uint8_t>((ok ? 1u : 0u) | (scm[c] != 0u ? 2u : 0u));
  }
}

// raw: 32 fp32 accumulators (as bits) of channels [32*cb, 32*cb+32) of this thread's row.
// scc: the shortcut's 32 fp16 values for the same channels (4 x uint4, logical chunk order).
// out: the 32 fp16 results (4 x uint4, logical chunk order), zero when the row is masked.
// MaskedDYT after an affine (nnlib/v2/layers.py:432-435): h = gamma * tanh(h) + beta, packed fp16.
__device__ __forceinline__ void dyt_apply_h2(__half2 (&h)[16], const uint4* g, const uint4* b, int cb) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 gg = g[cb * 4 + j], bb = b[cb * 4 + j];
    const __half2* g2 = reinterpret_cast<const __half2*>(&gg);
    const __half2* b2 = reinterpret_cast<const __half2*>(&bb);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t u = *reinterpret_cast<const uint32_t*>(&h[j * 4 + k]), t;
      asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(u));
      h[j * 4 + k] = __hfma2(*reinterpret_cast<const __half2*>(&t), g2[k], b2[k]);
    }
  }
}

// The generic epilogue in fp32 from the accumulator to the store (p.epi_f32): conv + bias -> [tap] -> affine -> [+ shortcut] ->
// activation -> [tap] -> [affine -> activation] -> [tap] -> [pool] -> ONE rounding to fp16.  No MaskedDYT on this path.
__device__ __forceinline__ void epilogue_batch_f32(const ConvParams& p, const float* s_par, int cb, const uint32_t (&raw)[32],
                                                const uint4 (&scc)[4], bool has_sc, bool sc_valid, bool valid, int lane, int win,
                                                uint4 (&out)[4]) {
  const float* sc1 = s_par + cb * 32;
  const float* sh1 = s_par + p.cout + cb * 32;
  const float* bias = s_par + 2 * p.cout + cb * 32;
  const float* sc2 = s_par + 3 * p.cout + cb * 32;
  const float* sh2 = s_par + 4 * p.cout + cb * 32;
  const float* scc_const = s_par + 5 * p.cout + cb * 32;
  float v[32];
  if (p.tap_mode == 1) {
    float tv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) tv[j] = valid ? __uint_as_float(raw[j]) + bias[j] : 0.0f;
    warp_cols_reduce<false>(tv, lane);
    atomicAdd(p.tap_sum + static_cast<long long>(win) * red_pitch_of(p) + cb * 32 + lane, tv[0]);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(raw[j]), sc1[j], sh1[j]);
  if (has_sc) {
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const __half2* s2 = reinterpret_cast<const __half2*>(&scc[j4]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(s2[k]);
        v[j4 * 8 + 2 * k] += sc_valid ? f.x : scc_const[j4 * 8 + 2 * k];
        v[j4 * 8 + 2 * k + 1] += sc_valid ? f.y : scc_const[j4 * 8 + 2 * k + 1];
      }
    }
  }
  act_apply_vec(v, p.act1);
  if (p.tap_mode == 2) {
    float tv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) tv[j] = valid ? v[j] : 0.0f;
    warp_cols_reduce<false>(tv, lane);
    atomicAdd(p.tap_sum + static_cast<long long>(win) * red_pitch_of(p) + cb * 32 + lane, tv[0]);
  }
  if (p.has_affine2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], sc2[j], sh2[j]);
    act_apply_vec(v, p.act2);
  }
  if (p.tap_mode == 3) {
    float tv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) tv[j] = valid ? v[j] : 0.0f;
    warp_cols_reduce<false>(tv, lane);
    atomicAdd(p.tap_sum + static_cast<long long>(win) * red_pitch_of(p) + cb * 32 + lane, tv[0]);
  }
  if (p.pool_mode != 0) {
    float tv[32];
    const bool pool_max = p.pool_mode == 1;
#pragma unroll
    for (int j = 0; j < 32; ++j) tv[j] = valid ? v[j] : (pool_max ? -3.0e38f : 0.0f);
    if (pool_max) {
      warp_cols_reduce<true>(tv, lane);
      if (tv[0] > -1.0e38f) atomic_max_f32(p.pool + static_cast<long long>(win) * red_pitch_of(p) + cb * 32 + lane, tv[0]);
    } else {
      warp_cols_reduce<false>(tv, lane);
      atomicAdd(p.pool + static_cast<long long>(win) * red_pitch_of(p) + cb * 32 + lane, tv[0]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 o;
    const __half2 a = cvt_sat_h2(valid ? v[j * 8 + 0] : 0.0f, valid ? v[j * 8 + 1] : 0.0f);
    const __half2 b = cvt_sat_h2(valid ? v[j * 8 + 2] : 0.0f, valid ? v[j * 8 + 3] : 0.0f);
    const __half2 c = cvt_sat_h2(valid ? v[j * 8 + 4] : 0.0f, valid ? v[j * 8 + 5] : 0.0f);
    const __half2 d = cvt_sat_h2(valid ? v[j * 8 + 6] : 0.0f, valid ? v[j * 8 + 7] : 0.0f);
    o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&b);
    o.z = *reinterpret_cast<const uint32_t*>(&c); o.w = *reinterpret_cast<const uint32_t*>(&d);
    out[j] = o;
  }
}

// kMode specialises the epilogue at compile time for the layer shapes that carry the time, so that
// the flag tests disappear and the whole batch is one scheduling region (the runtime-flag version
// splits it into basic blocks the compiler cannot overlap):
//   EPI_GENERIC  every feature behind its runtime flag
//   EPI_LIGHT    tanh-GELU only: no NMD tap, no second affine, no pooling (conv1 / conv2 of a residual block)
//   EPI_FINAL    shortcut + tanh-GELU + NMD tap on the block output + second affine + tanh-GELU
//   EPI_FINAL_POOL  the same + masked global max pool (last layer)
// The specialised modes also assume the first affine's scale is folded into the weights (p.folded).
// The caller checks that the layer matches the mode it picks.
// kD1 / kD2: in a specialised mode, the first / second norm is a MaskedDYT (tanh + gamma / beta after the affine).
enum EpiMode { EPI_GENERIC = 0, EPI_LIGHT = 1, EPI_FINAL = 2, EPI_FINAL_POOL = 3 };
template <int kMode_, bool kD1_ = false, bool kD2_ = false>
struct EpiTag { static constexpr int kMode = kMode_; static constexpr bool kD1 = kD1_, kD2 = kD2_; };
template <int kMode = EPI_GENERIC, bool kD1 = false, bool kD2 = false>
__device__ __forceinline__ void epilogue_batch(const ConvParams& p, const EpiParams& e, int cb, const uint32_t (&raw)[32],
                                               const uint4 (&scc)[4], bool has_sc, bool sc_valid, bool valid, int lane,
                                               int win, uint4 (&out)[4], float ln_mu = 0.0f, float ln_rs = 1.0f) {
  constexpr bool kGen = kMode == EPI_GENERIC, kFinal = kMode == EPI_FINAL || kMode == EPI_FINAL_POOL;
  if (kGen && p.epi_f32) {
    epilogue_batch_f32(p, reinterpret_cast<const float*>(e.scale1), cb, raw, scc, has_sc, sc_valid, valid, lane, win, out);
    return;
  }
  if (kGen && p.tap_mode == 1) {   // NMD tap on the raw conv output (acc + bias), stem only
    float tv[32];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 b = e.bias[cb * 8 + j4];
      tv[j4 * 4 + 0] = valid ? __uint_as_float(raw[j4 * 4 + 0]) + b.x : 0.0f;
      tv[j4 * 4 + 1] = valid ? __uint_as_float(raw[j4 * 4 + 1]) + b.y : 0.0f;
      tv[j4 * 4 + 2] = valid ? __uint_as_float(raw[j4 * 4 + 2]) + b.z : 0.0f;
      tv[j4 * 4 + 3] = valid ? __uint_as_float(raw[j4 * 4 + 3]) + b.w : 0.0f;
    }
    warp_cols_reduce<false>(tv, lane);
    atomicAdd(p.tap_sum + static_cast<long long>(win) * red_pitch_of(p) + cb * 32 + lane, tv[0]);
  }
  __half2 h[16];
  if (kGen && p.ln1) {       // MaskedLayerNormalization: (acc + bias - mean) / sqrt(var + eps) * gamma + beta, row statistics from the caller
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 a = e.scale1[cb * 8 + j4], b = e.shift1[cb * 8 + j4], c = e.bias[cb * 8 + j4];
      h[j4 * 2 + 0] = cvt_sat_h2(fmaf((__uint_as_float(raw[j4 * 4 + 0]) + c.x - ln_mu) * ln_rs, a.x, b.x),
                                 fmaf((__uint_as_float(raw[j4 * 4 + 1]) + c.y - ln_mu) * ln_rs, a.y, b.y));
      h[j4 * 2 + 1] = cvt_sat_h2(fmaf((__uint_as_float(raw[j4 * 4 + 2]) + c.z - ln_mu) * ln_rs, a.z, b.z),
                                 fmaf((__uint_as_float(raw[j4 * 4 + 3]) + c.w - ln_mu) * ln_rs, a.w, b.w));
    }
  } else if (kGen) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 a = e.scale1[cb * 8 + j4], b = e.shift1[cb * 8 + j4];
      h[j4 * 2 + 0] = cvt_sat_h2(fmaf(__uint_as_float(raw[j4 * 4 + 0]), a.x, b.x), fmaf(__uint_as_float(raw[j4 * 4 + 1]), a.y, b.y));
      h[j4 * 2 + 1] = cvt_sat_h2(fmaf(__uint_as_float(raw[j4 * 4 + 2]), a.z, b.z), fmaf(__uint_as_float(raw[j4 * 4 + 3]), a.w, b.w));
    }
  } else {          // specialised modes run on layers whose scale is folded into the weights
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 b = e.shift1[cb * 8 + j4];
      h[j4 * 2 + 0] = cvt_sat_h2(__uint_as_float(raw[j4 * 4 + 0]) + b.x, __uint_as_float(raw[j4 * 4 + 1]) + b.y);
      h[j4 * 2 + 1] = cvt_sat_h2(__uint_as_float(raw[j4 * 4 + 2]) + b.z, __uint_as_float(raw[j4 * 4 + 3]) + b.w);
    }
  }
  if (kD1 || (kGen && p.dyt1)) dyt_apply_h2(h, e.g1, e.b1, cb);
  if (kFinal || has_sc) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 s = sc_valid ? scc[j] : e.scc[cb * 4 + j];
      const __half2* s2 = reinterpret_cast<const __half2*>(&s);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[j * 4 + k] = __hadd2(h[j * 4 + k], s2[k]);
    }
  }
  if (kFinal || kMode == EPI_LIGHT) act_apply_h2(h, ACT_GELU_TANH); else act_apply_h2(h, p.act1);
  // Rows outside the mask are rare (frame ends and the window tail): the masking selects only run
  // in warps that contain one.
  const bool any_masked = __any_sync(0xffffffffu, !valid);
  if (kFinal) {
    // second affine first (it is the last reader of h), so the tap can mask h in place
    __half2 x3[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 a = e.scale2[cb * 4 + j], b = e.shift2[cb * 4 + j];
      const __half2* a2 = reinterpret_cast<const __half2*>(&a);
      const __half2* b2 = reinterpret_cast<const __half2*>(&b);
#pragma unroll
      for (int k = 0; k < 4; ++k) x3[j * 4 + k] = __hfma2(h[j * 4 + k], a2[k], b2[k]);
    }
    if (kD2) dyt_apply_h2(x3, e.g2, e.b2, cb);
    if (any_masked) {
      const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = valid ? h[i] : zero;
    }
    atomicAdd(p.tap_sum + static_cast<long long>(win) * red_pitch_of(p) + cb * 32 + lane, warp_cols_reduce_h2<false>(h, lane));
#pragma unroll
    for (int i = 0; i < 16; ++i) h[i] = x3[i];
    act_apply_h2(h, ACT_GELU_TANH);
  } else {
    if (kGen && p.tap_mode == 2) {   // NMD tap on the block output
      __half2 tv[16];
      const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
      for (int i = 0; i < 16; ++i) tv[i] = valid ? h[i] : zero;
      atomicAdd(p.tap_sum + static_cast<long long>(win) * red_pitch_of(p) + cb * 32 + lane, warp_cols_reduce_h2<false>(tv, lane));
    }
    if (kGen && p.has_affine2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 a = e.scale2[cb * 4 + j], b = e.shift2[cb * 4 + j];
        const __half2* a2 = reinterpret_cast<const __half2*>(&a);
        const __half2* b2 = reinterpret_cast<const __half2*>(&b);
#pragma unroll
        for (int k = 0; k < 4; ++k) h[j * 4 + k] = __hfma2(h[j * 4 + k], a2[k], b2[k]);
      }
      if (p.dyt2) dyt_apply_h2(h, e.g2, e.b2, cb);
      act_apply_h2(h, p.act2);
    }
    if (kGen && p.tap_mode == 3) {   // NMD tap on the launch output (after the stand-alone norm + activation)
      __half2 tv[16];
      const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
      for (int i = 0; i < 16; ++i) tv[i] = valid ? h[i] : zero;
      atomicAdd(p.tap_sum + static_cast<long long>(win) * red_pitch_of(p) + cb * 32 + lane, warp_cols_reduce_h2<false>(tv, lane));
    }
  }
  if (kMode == EPI_FINAL_POOL || (kGen && p.pool_mode != 0)) {
    __half2 tv[16];
    const bool pool_max = kMode == EPI_FINAL_POOL || p.pool_mode == 1;
    const uint32_t fill_bits = pool_max ? 0xFC00FC00u : 0u;      // -inf for the max, 0 for the sum
    const __half2 fill = *reinterpret_cast<const __half2*>(&fill_bits);
#pragma unroll
    for (int i = 0; i < 16; ++i) tv[i] = valid ? h[i] : fill;
    if (pool_max) {
      const float m = warp_cols_reduce_h2<true>(tv, lane);
      if (m > -1.0e38f) atomic_max_f32(p.pool + static_cast<long long>(win) * red_pitch_of(p) + cb * 32 + lane, m);
    } else {
      atomicAdd(p.pool + static_cast<long long>(win) * red_pitch_of(p) + cb * 32 + lane, warp_cols_reduce_h2<false>(tv, lane));
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t*>(&h[j * 4 + 0]);
    o.y = *reinterpret_cast<const uint32_t*>(&h[j * 4 + 1]);
    o.z = *reinterpret_cast<const uint32_t*>(&h[j * 4 + 2]);
    o.w = *reinterpret_cast<const uint32_t*>(&h[j * 4 + 3]);
    out[j] = o;
  }
  if (any_masked) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = valid ? out[j] : make_uint4(0u, 0u, 0u, 0u);
  }
}

}  // namespace jg
