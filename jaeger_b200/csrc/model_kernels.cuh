// Small kernels around the conv stack: token expansion into the stem's one-hot operand,
// per-layer mask propagation, and the dense heads.  All HBM-bound byte / row work.
// Reference: nnlib/builder.py:844-894 (Embedding mask_zero / Masking), nnlib/v2/layers.py:
// 1245-1252 (mask propagation), 517-529 / 460-480 (masked pooling), nnlib/v2/nmd.py:43-77.
#pragma once
#include "conv_common.cuh"

namespace jg {

// Row geometry shared by the kernels: window w owns rows [w*rpw, (w+1)*rpw); frame f of the
// window starts at f*period; position j of the frame is valid input when j < lpad[w] - shrink.
struct RowGeom {
  int rpw, period, frames;
};

// tokens [W][6][pitch] -> one-hot rows (64 channels, g64sw) + input mask (token != 0).
// One thread per 16-byte chunk, 8 threads per row; every row of the window block is written
// (gap / tail rows get zeros) so the buffer can be recycled between chunks of windows.
__global__ void expand_tokens_kernel(const uint8_t* __restrict__ tokens, const int* __restrict__ lpad,
                                     long long n_rows, int lc, int pitch, RowGeom g,
                                     __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ mask,
                                     int* __restrict__ count) {
  const long long total = n_rows * 8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = idx >> 3;
    const int pc = static_cast<int>(idx & 7);            // physical chunk position
    const int chunk = pc ^ static_cast<int>(row & 7);    // logical chunk = channels 8*chunk..
    const long long w = row / g.rpw;
    const int rw = static_cast<int>(row - w * g.rpw);
    const int f = rw / g.period, j = rw - f * g.period;
    int tok = 0;
    if (f < g.frames && j < lc && j < lpad[w]) tok = tokens[(w * g.frames + f) * pitch + j];
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    const int ch = tok - 1;                               // token t -> one-hot channel t-1
    if (tok > 0 && (ch >> 3) == chunk) {
      const uint32_t one = 0x3F80u << (16 * (ch & 1));
      const int word = (ch & 7) >> 1;
      if (word == 0) o.x = one; else if (word == 1) o.y = one; else if (word == 2) o.z = one; else o.w = one;
    }
    *reinterpret_cast<uint4*>(x + row * 64 + pc * 8) = o;
    if (pc == 0) {
      mask[row] = tok > 0;
      if (tok > 0) atomicAdd(count + w, 1);
    }
  }
}

// out_mask[row] = in-frame(row, L_out) && (masking ? OR_t in_mask[row + shift_t] : 1)
// (nnlib/v2/layers.py:1245-1252, mask_mode "any").  Also accumulates the per-window count of
// valid rows that the NMD taps and the pooling need.
__global__ void propagate_mask_kernel(const uint8_t* __restrict__ in_mask, const int* __restrict__ lpad,
                                      long long n_rows, RowGeom g, int shrink_out, int ntaps,
                                      const int* __restrict__ shifts, int masking,
                                      uint8_t* __restrict__ out_mask, int* __restrict__ count) {
  for (long long row = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; row < n_rows;
       row += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long w = row / g.rpw;
    const int rw = static_cast<int>(row - w * g.rpw);
    const int f = rw / g.period, j = rw - f * g.period;
    int ok = (f < g.frames) && (j < lpad[w] - shrink_out);
    if (ok && masking) {
      int any = 0;
      for (int t = 0; t < ntaps; ++t) any |= in_mask[row + shifts[t]];
      ok = any;
    }
    out_mask[row] = static_cast<uint8_t>(ok);
    // warp-aggregated count (rows of a warp almost always share the window)
    const unsigned ballot = __ballot_sync(__activemask(), ok);
    const long long w0 = __shfl_sync(__activemask(), w, 0);
    if (__all_sync(__activemask(), w == w0)) {
      if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(count + w, __popc(ballot));
    } else if (ok) {
      atomicAdd(count + w, 1);
    }
  }
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) p[i] = v;
}

struct HeadParams {
  const float* pool;        // [W][feat] masked max (sentinel -1e9) or masked sum
  const int* pool_count;    // [W] valid rows under the final mask
  const float* tap_sum;     // [n_taps][W][tap_width]
  const int* const* tap_count;  // n_taps pointers to [W]
  const float* tap_mean;    // [n_taps][tap_width] moving means
  const float* cls_w; const float* cls_b;
  const float* rel_w1; const float* rel_b1; const float* rel_w2; const float* rel_b2;
  float* logits; float* rel; float* emb; float* nmd;
  int n_windows, feat, n_classes, pool_mode, n_taps, tap_width, rel_hidden, masking;
};

// One warp per window: finalise the pooled features, classifier dense, NMD vector,
// reliability head (Dense(gelu) -> Dense(1)).  fp32 throughout (builder.py:589-596,705-713).
__global__ void heads_kernel(const HeadParams p) {
  extern __shared__ float s_feat[];   // per warp: feat + n_taps*tap_width floats
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nmd_dim = p.n_taps * p.tap_width;
  float* feat = s_feat + static_cast<size_t>(warp) * (p.feat + nmd_dim);
  float* nmdv = feat + p.feat;
  for (int w = blockIdx.x * warps_per_block + warp; w < p.n_windows; w += gridDim.x * warps_per_block) {
    const int cnt = p.pool_count[w];
    for (int c = lane; c < p.feat; c += 32) {
      float v = p.pool[static_cast<long long>(w) * p.feat + c];
      if (p.pool_mode == 1) v = cnt > 0 ? v : 0.0f;   // all-masked sample pools to zeros
      else v = p.masking ? (cnt > 0 ? v / fmaxf(static_cast<float>(cnt), 1e-7f) : 0.0f)
                         : v / static_cast<float>(cnt);
      feat[c] = v;
      if (p.emb) p.emb[static_cast<long long>(w) * p.feat + c] = v;
    }
    for (int i = lane; i < nmd_dim; i += 32) {
      const int t = i / p.tap_width, c = i - t * p.tap_width;
      const float s = p.tap_sum[(static_cast<long long>(t) * p.n_windows + w) * p.tap_width + c];
      const float n = static_cast<float>(p.tap_count[t][w]);
      const float mean = p.masking ? s / (n + 1e-5f) : s / n;
      const float v = mean - p.tap_mean[t * p.tap_width + c];
      nmdv[i] = v;
      if (p.nmd) p.nmd[static_cast<long long>(w) * nmd_dim + i] = v;
    }
    __syncwarp();
    for (int k = 0; k < p.n_classes; ++k) {
      float acc = 0.0f;
      for (int c = lane; c < p.feat; c += 32) acc = fmaf(feat[c], p.cls_w[c * p.n_classes + k], acc);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) p.logits[static_cast<long long>(w) * p.n_classes + k] = acc + p.cls_b[k];
    }
    if (p.rel && p.n_taps > 0) {
      float out = 0.0f;
      for (int h = 0; h < p.rel_hidden; ++h) {
        float acc = 0.0f;
        for (int i = lane; i < nmd_dim; i += 32) acc = fmaf(nmdv[i], p.rel_w1[i * p.rel_hidden + h], acc);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        const float a = acc + p.rel_b1[h];
        // Keras Dense(activation="gelu") is the tanh approximation; full-precision tanhf here
        const float g = 0.5f * a * (1.0f + tanhf(0.7978845608028654f * (a + 0.044715f * a * a * a)));
        out = fmaf(g, p.rel_w2[h], out);
      }
      if (lane == 0) p.rel[w] = out + p.rel_b2[0];
    }
    __syncwarp();
  }
}

}  // namespace jg
