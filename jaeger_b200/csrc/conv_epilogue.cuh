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

// Per-channel parameters staged in shared memory: fp32 for the stage that touches the fp32
// accumulator, fp16 pairs for the packed stages.
struct EpiParams {
  const float4* scale1;   // [cout/4]
  const float4* shift1;
  const float4* bias;
  const uint4* scale2;    // [cout/8] 8 halves each
  const uint4* shift2;
  const uint4* scc;       // shortcut value at masked rows
};

__device__ __forceinline__ EpiParams epi_params(float* s_par, int cout) {
  EpiParams e;
  e.scale1 = reinterpret_cast<const float4*>(s_par);
  e.shift1 = reinterpret_cast<const float4*>(s_par + cout);
  e.bias = reinterpret_cast<const float4*>(s_par + 2 * cout);
  e.scale2 = reinterpret_cast<const uint4*>(s_par + 3 * cout);
  e.shift2 = reinterpret_cast<const uint4*>(s_par + 4 * cout);
  e.scc = reinterpret_cast<const uint4*>(s_par + 5 * cout);
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
    h2[i] = __float2half_rn(p.has_affine2 ? p.scale2[i] : 1.0f);
    t2[i] = __float2half_rn(p.has_affine2 ? p.shift2[i] : 0.0f);
    sc[i] = __float2half_rn(p.sc_const ? p.sc_const[i] : 0.0f);
  }
}

// Validity of this thread's output row.  With fuse_mask the propagate_mask_kernel launch is folded
// in here: one thread owns one row of the tile, so it also publishes the mask and the warp adds
// its valid-row count to the window's counter.
__device__ __forceinline__ bool row_valid(const ConvParams& p, long long row, int win, int lane) {
  if (!p.fuse_mask) return p.out_mask[row] != 0;
  const int rw = static_cast<int>(row - static_cast<long long>(win) * p.rows_per_window);
  const int f = rw / p.period, j = rw - f * p.period;
  bool ok = f < p.frames && j < ((p.lpad[win] - p.shrink_in) >> p.halvings) - p.shrink;
  if (ok && p.masking) {
    int any = 0;
    for (int t = 0; t < p.ntaps; ++t) any |= p.in_mask[row + p.shifts[t]];
    ok = any != 0;
  }
  p.out_mask_w[row] = static_cast<uint8_t>(ok);
  const unsigned b = __ballot_sync(0xffffffffu, ok);
  if (lane == 0 && b) atomicAdd(p.count + win, __popc(b));
  return ok;
}

// raw: 32 fp32 accumulators (as bits) of channels [32*cb, 32*cb+32) of this thread's row.
// scc: the shortcut's 32 fp16 values for the same channels (4 x uint4, logical chunk order).
// out: the 32 fp16 results (4 x uint4, logical chunk order), zero when the row is masked.
__device__ __forceinline__ void epilogue_batch(const ConvParams& p, const EpiParams& e, int cb, const uint32_t (&raw)[32],
                                               const uint4 (&scc)[4], bool has_sc, bool sc_valid, bool valid, int lane,
                                               int win, uint4 (&out)[4]) {
  if (p.tap_mode == 1) {   // NMD tap on the raw conv output (acc + bias), stem only
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
    atomicAdd(p.tap_sum + static_cast<long long>(win) * p.cout + cb * 32 + lane, tv[0]);
  }
  __half2 h[16];
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 a = e.scale1[cb * 8 + j4], b = e.shift1[cb * 8 + j4];
    h[j4 * 2 + 0] = cvt_sat_h2(fmaf(__uint_as_float(raw[j4 * 4 + 0]), a.x, b.x), fmaf(__uint_as_float(raw[j4 * 4 + 1]), a.y, b.y));
    h[j4 * 2 + 1] = cvt_sat_h2(fmaf(__uint_as_float(raw[j4 * 4 + 2]), a.z, b.z), fmaf(__uint_as_float(raw[j4 * 4 + 3]), a.w, b.w));
  }
  if (has_sc) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 s = sc_valid ? scc[j] : e.scc[cb * 4 + j];
      const __half2* s2 = reinterpret_cast<const __half2*>(&s);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[j * 4 + k] = __hadd2(h[j * 4 + k], s2[k]);
    }
  }
  act_apply_h2(h, p.act1);
  if (p.tap_mode == 2) {   // NMD tap on the block output
    float tv[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 f = __half22float2(h[i]);
      tv[2 * i] = valid ? f.x : 0.0f;
      tv[2 * i + 1] = valid ? f.y : 0.0f;
    }
    warp_cols_reduce<false>(tv, lane);
    atomicAdd(p.tap_sum + static_cast<long long>(win) * p.cout + cb * 32 + lane, tv[0]);
  }
  if (p.has_affine2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 a = e.scale2[cb * 4 + j], b = e.shift2[cb * 4 + j];
      const __half2* a2 = reinterpret_cast<const __half2*>(&a);
      const __half2* b2 = reinterpret_cast<const __half2*>(&b);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[j * 4 + k] = __hfma2(h[j * 4 + k], a2[k], b2[k]);
    }
    act_apply_h2(h, p.act2);
  }
  if (p.pool_mode != 0) {
    float tv[32];
    const float fill = p.pool_mode == 1 ? -3.0e38f : 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 f = __half22float2(h[i]);
      tv[2 * i] = valid ? f.x : fill;
      tv[2 * i + 1] = valid ? f.y : fill;
    }
    if (p.pool_mode == 1) {
      warp_cols_reduce<true>(tv, lane);
      if (tv[0] > -1.0e38f) atomic_max_f32(p.pool + static_cast<long long>(win) * p.cout + cb * 32 + lane, tv[0]);
    } else {
      warp_cols_reduce<false>(tv, lane);
      atomicAdd(p.pool + static_cast<long long>(win) * p.cout + cb * 32 + lane, tv[0]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t*>(&h[j * 4 + 0]);
    o.y = *reinterpret_cast<const uint32_t*>(&h[j * 4 + 1]);
    o.z = *reinterpret_cast<const uint32_t*>(&h[j * 4 + 2]);
    o.w = *reinterpret_cast<const uint32_t*>(&h[j * 4 + 3]);
    out[j] = valid ? o : make_uint4(0u, 0u, 0u, 0u);
  }
}

}  // namespace jg
