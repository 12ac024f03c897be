// CUDA-core restatement of the conv layer (same ConvParams contract as conv_tc.cuh).
// It exists so that the tensor-core kernel can be validated ON THE GPU at sizes no CPU
// oracle finishes in reasonable time (tests call both through the C ABI and compare);
// it is never on the product path.
#pragma once
#include "conv_common.cuh"

namespace jg {
namespace ref {

__global__ void conv_ref_kernel(const ConvParams p) {
  const long long total = static_cast<long long>(p.n_tiles) * kTileM * p.cout;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = idx / p.cout;
    const int co = static_cast<int>(idx % p.cout);
    float acc = 0.0f;
    for (int t = 0; t < p.ntaps; ++t) {
      const long long r = row + p.shifts[t];
      for (int ci = 0; ci < p.cin; ++ci) {
        const float xv = __half2float(p.x[act_index(r, ci, p.x_plane)]);
        const float wv = __half2float(p.w[w_index(t, ci, co, p.cin, p.cout)]);
        acc = fmaf(xv, wv, acc);
      }
    }
    const bool valid = p.out_mask[row] != 0;
    const int win = static_cast<int>(row / p.rows_per_window);
    if (p.tap_mode == 1 && valid)
      atomicAdd(p.tap_sum + static_cast<long long>(win) * red_pitch_of(p) + co, acc + p.bias[co]);
    float v = fmaf(acc, p.scale1[co], p.shift1[co]);
    if (p.ln1) {      // MaskedLayerNormalization: the row's channel statistics, recomputed by every thread of the row (test path only)
      float s1 = 0.0f, s2 = 0.0f;
      for (int c2 = 0; c2 < p.cout; ++c2) {
        float a2 = 0.0f;
        for (int t = 0; t < p.ntaps; ++t) {
          const long long r = row + p.shifts[t];
          for (int ci = 0; ci < p.cin; ++ci)
            a2 = fmaf(__half2float(p.x[act_index(r, ci, p.x_plane)]), __half2float(p.w[w_index(t, ci, c2, p.cin, p.cout)]), a2);
        }
        a2 += p.bias[c2];
        s1 += a2;
        s2 = fmaf(a2, a2, s2);
      }
      const float mu = s1 * p.ln_inv_c;
      const float rs = 1.0f / sqrtf(fmaxf(s2 * p.ln_inv_c - mu * mu, 0.0f) + p.ln_eps);
      v = fmaf((acc + p.bias[co] - mu) * rs, p.scale1[co], p.shift1[co]);
    }
    if (p.dyt1) v = fmaf(tanhf(v), p.dyt_g1[co], p.dyt_b1[co]);
    if (p.sc) {
      const bool scv = p.sc_mask ? p.sc_mask[row] != 0 : true;
      v += scv ? __half2float(p.sc[act_index(row, co, p.y_plane)])
               : (p.sc_const ? p.sc_const[co] : 0.0f);
    }
    v = act_apply(v, p.act1);
    if (p.tap_mode == 2 && valid) atomicAdd(p.tap_sum + static_cast<long long>(win) * red_pitch_of(p) + co, v);
    if (p.has_affine2) {
      v = fmaf(v, p.scale2[co], p.shift2[co]);
      if (p.dyt2) v = fmaf(tanhf(v), p.dyt_g2[co], p.dyt_b2[co]);
      v = act_apply(v, p.act2);
    }
    if (p.tap_mode == 3 && valid) atomicAdd(p.tap_sum + static_cast<long long>(win) * red_pitch_of(p) + co, v);
    if (p.pool_mode == 1 && valid) atomic_max_f32(p.pool + static_cast<long long>(win) * red_pitch_of(p) + co, v);
    if (p.pool_mode == 2 && valid) atomicAdd(p.pool + static_cast<long long>(win) * red_pitch_of(p) + co, v);
    if (p.y)
      p.y[act_index(row, co, p.y_plane)] = __float2half_rn(valid ? v : 0.0f);
  }
}

}  // namespace ref
}  // namespace jg
