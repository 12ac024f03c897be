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
// One CTA per (window, frame) segment -- plus one per window for the zero tail rows -- so the row
// geometry costs no per-thread division; one thread per 16-byte chunk, consecutive threads write
// consecutive chunks.  Every row of the window block is written (gap / tail rows get zeros) so the
// buffer can be recycled between chunks of windows.
__global__ void expand_tokens_kernel(const uint8_t* __restrict__ tokens, const int* __restrict__ lpad,
                                     long long n_windows, int lc, int pitch, RowGeom g, int tok_offset,
                                     act_t* __restrict__ x, uint8_t* __restrict__ mask,
                                     int* __restrict__ count) {
  const int n_part = g.frames + 1;
  const long long n_seg = n_windows * n_part;
  for (long long sg = blockIdx.x; sg < n_seg; sg += gridDim.x) {
    const long long w = sg / n_part;
    const int f = static_cast<int>(sg - w * n_part);
    const int row_lo = f * g.period;
    const int n_rows = (f < g.frames) ? g.period : g.rpw - row_lo;
    int lim = 0;
    if (f < g.frames) { lim = lpad[w]; lim = lim < lc ? lim : lc; }
    const uint8_t* tk = tokens + (w * g.frames + (f < g.frames ? f : 0)) * pitch;
    const long long row0 = w * g.rpw + row_lo;
    int cnt = 0;
    for (int t = threadIdx.x; t < n_rows * 8; t += blockDim.x) {
      const int j = t >> 3, pc = t & 7;                      // row of the segment, physical chunk position
      const long long row = row0 + j;
      const int chunk = pc ^ static_cast<int>(row & 7);      // logical chunk = channels 8*chunk..
      // token t -> one-hot channel t - tok_offset (v2: offset 1, token 0 = unknown = zero row;
      // legacy: offset 0, every in-frame position has a channel)
      const int ch = (j < lim ? static_cast<int>(tk[j]) : -1) - tok_offset;
      const bool on = ch >= 0 && j < lim;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (on && (ch >> 3) == chunk) {
        const uint32_t one = 0x3C00u << (16 * (ch & 1));     // fp16 1.0
        const int word = (ch & 7) >> 1;
        if (word == 0) o.x = one; else if (word == 1) o.y = one; else if (word == 2) o.z = one; else o.w = one;
      }
      *reinterpret_cast<uint4*>(x + row * 64 + pc * 8) = o;
      if (pc == 0) {
        mask[row] = on;
        cnt += on ? 1 : 0;
      }
    }
    for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(count + w, cnt);
  }
}

// out_mask[row] = in-frame(row, L_out) && (masking ? OR_t in_mask[row + shift_t] : 1)
// (nnlib/v2/layers.py:1245-1252, mask_mode "any").  Also accumulates the per-window count of
// valid rows that the NMD taps and the pooling need.
__global__ void propagate_mask_kernel(const uint8_t* __restrict__ in_mask, const int* __restrict__ lpad,
                                      long long n_rows, RowGeom g, int shrink_in, int halvings, int len_round, int shrink,
                                      int ntaps, const int* __restrict__ shifts, int masking, int mask_thr,
                                      uint8_t* __restrict__ out_mask, int* __restrict__ count) {
  for (long long row = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; row < n_rows;
       row += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long w = row / g.rpw;
    const int rw = static_cast<int>(row - w * g.rpw);
    const int f = rw / g.period, j = rw - f * g.period;
    int ok = (f < g.frames) && (j < ((lpad[w] - shrink_in + len_round) >> halvings) - shrink);
    if (ok && masking) {
      int any = 0;
      for (int t = 0; t < ntaps; ++t) any += in_mask[row + shifts[t]];
      ok = any >= (mask_thr > 1 ? mask_thr : 1);
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

// Rows -> (even, odd) row planes inside every frame: out[w,f,j] = (in[w,f,2j] | in[w,f,2j+1]) along the channel axis for
// j < ceil(L_in / 2), zero elsewhere (a missing odd row of an odd-length frame is zero = the SAME padding of the strided conv).
// A stride-2 conv over `in` is then a stride-1 conv over `out` with twice the input channels (plan.py:split_phases;
// reference: the strided conv1 / bypass conv of a ResidualBlock, nnlib/v2/layers.py:1840-1864).  One thread per 16-byte chunk of
// an output row; both tensors are g64sw.  groups = channel groups of the INPUT; output group p * groups + g = phase p, group g.
__global__ void rows_to_phases_kernel(const act_t* __restrict__ x, const int* __restrict__ lpad, long long n_rows, RowGeom g,
                                      int shrink_in, int halvings, int len_round, int groups, long long plane,
                                      act_t* __restrict__ y) {
  const long long total = n_rows * 8 * groups * 2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pc = static_cast<int>(idx & 7);
    const long long t = idx >> 3;
    const long long row = t % n_rows;
    const int og = static_cast<int>(t / n_rows);               // output group
    const int phase = og / groups, grp = og - phase * groups;
    const long long w = row / g.rpw;
    const int rw = static_cast<int>(row - w * g.rpw);
    const int f = rw / g.period, j = rw - f * g.period;
    const int l_in = (lpad[w] - shrink_in + len_round) >> halvings;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (f < g.frames && 2 * j + phase < l_in) {
      const int chunk = pc ^ static_cast<int>(row & 7);        // logical chunk held at this position of the output row
      const long long r = w * g.rpw + static_cast<long long>(f) * g.period + 2 * j + phase;
      o = *reinterpret_cast<const uint4*>(x + (grp * plane + r) * 64 + ((chunk ^ static_cast<int>(r & 7)) * 8));
    }
    *reinterpret_cast<uint4*>(y + (og * plane + row) * 64 + pc * 8) = o;
  }
}

// MaxPooling1D(pool_size=2) inside every frame (legacy graph, nnlib/v1/layers.py:65-69):
// out[w,f,j] = max(in[w,f,2j], in[w,f,2j+1]) for j < L_in/2, zero elsewhere.  One thread per
// 16-byte chunk of an output row; both tensors are g64sw, so the chunk position is re-swizzled
// for every row it touches.
__global__ void maxpool2_kernel(const act_t* __restrict__ x, const int* __restrict__ lpad, long long n_rows,
                                RowGeom g, int shrink_in, int halvings, int groups, long long plane,
                                act_t* __restrict__ y, uint8_t* __restrict__ out_mask, int* __restrict__ count) {
  const long long total = n_rows * 8 * groups;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pc = static_cast<int>(idx & 7);
    const long long t = idx >> 3;
    const long long row = t % n_rows;
    const int grp = static_cast<int>(t / n_rows);
    const long long w = row / g.rpw;
    const int rw = static_cast<int>(row - w * g.rpw);
    const int f = rw / g.period, j = rw - f * g.period;
    const int l_out = ((lpad[w] - shrink_in) >> halvings) >> 1;
    const bool ok = f < g.frames && j < l_out;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (ok) {
      const int chunk = pc ^ static_cast<int>(row & 7);
      const long long r0 = w * g.rpw + static_cast<long long>(f) * g.period + 2 * j, r1 = r0 + 1;
      const uint4 a = *reinterpret_cast<const uint4*>(x + (grp * plane + r0) * 64 + ((chunk ^ static_cast<int>(r0 & 7)) * 8));
      const uint4 b = *reinterpret_cast<const uint4*>(x + (grp * plane + r1) * 64 + ((chunk ^ static_cast<int>(r1 & 7)) * 8));
      const __half2* pa = reinterpret_cast<const __half2*>(&a);
      const __half2* pb = reinterpret_cast<const __half2*>(&b);
      __half2* po = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) po[e] = __hmax2(pa[e], pb[e]);
    }
    *reinterpret_cast<uint4*>(y + (grp * plane + row) * 64 + pc * 8) = o;
    if (grp == 0 && pc == 0) {
      out_mask[row] = ok;
      if (ok) atomicAdd(count + w, 1);
    }
  }
}

// Add over the six frames followed by GlobalMaxPool1D over the positions (legacy graph,
// nnlib/v1/layers.py:207, 413).  One CTA per window; thread = (8-channel chunk, position lane).
__global__ void framesum_globalmax_kernel(const act_t* __restrict__ x, const int* __restrict__ lpad,
                                          int n_windows, RowGeom g, int shrink_in, int halvings, int channels,
                                          long long plane, float* __restrict__ pool) {
  extern __shared__ float s_max[];     // [pos lanes][channels]
  const int chunks = channels / 8;
  const int lanes = blockDim.x / chunks;
  const int chunk = threadIdx.x % chunks, lane = threadIdx.x / chunks;
  for (int w = blockIdx.x; w < n_windows; w += gridDim.x) {
    const int L = (lpad[w] - shrink_in) >> halvings;
    float best[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) best[e] = -3.0e38f;
    for (int j = lane; j < L; j += lanes) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int f = 0; f < g.frames; ++f) {
        const long long row = static_cast<long long>(w) * g.rpw + static_cast<long long>(f) * g.period + j;
        const uint4 v = *reinterpret_cast<const uint4*>(x + ((chunk >> 3) * plane + row) * 64 + (((chunk & 7) ^ static_cast<int>(row & 7)) * 8));
        const __half2* pv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 q = __half22float2(pv[e]);
          acc[2 * e] += q.x;
          acc[2 * e + 1] += q.y;
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) best[e] = fmaxf(best[e], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) s_max[lane * channels + chunk * 8 + e] = best[e];
    __syncthreads();
    for (int c = threadIdx.x; c < channels; c += blockDim.x) {
      float m = -3.0e38f;
      for (int l = 0; l < lanes; ++l) m = fmaxf(m, s_max[l * channels + c]);
      pool[static_cast<long long>(w) * channels + c] = L > 0 ? m : 0.0f;
    }
    __syncthreads();
  }
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) p[i] = v;
}

constexpr int kMaxSignalClasses = 16;

struct HeadParams {
  const float* pool;        // [W][feat] masked max (sentinel -1e9) or masked sum
  const int* pool_count;    // [W] valid rows under the final mask
  const float* tap_sum;     // [n_taps][W][tap_width]
  const int* const* tap_count;  // n_taps pointers to [W]
  const float* tap_mean;    // [n_taps][tap_width] moving means
  const float* cls_w; const float* cls_b;
  const float* rel_w1; const float* rel_b1; const float* rel_w2; const float* rel_b2;
  const float* mlp_w1; const float* mlp_b1; const float* mlp_w2; const float* mlp_b2;   // optional Dense, Dense before the classifier
  float* logits; float* rel; float* emb; float* nmd;
  int n_windows, feat, n_classes, pool_mode, n_taps, tap_width, rel_hidden, masking;
  int mlp_hidden, mlp_act, pool_final;   // pool_final: the pool buffer already holds finished features
  int signals;   // reliability_model.mode "nmd_plus_signals" (builder.py:644-657): n | id0 << 3 | id1 << 6 ...; ids:
                 // 1 max_prob, 2 entropy, 3 energy, 4 margin, 5 nmd_norm (OODSignalLayer, nnlib/v2/layers.py:1632-1666)
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
  const int mlp_layers = ((p.signals >> 20) & 3) ? ((p.signals >> 20) & 3) : 2;
  const bool emb_pooled = ((p.signals >> 22) & 1) != 0;
  for (int w = blockIdx.x * warps_per_block + warp; w < p.n_windows; w += gridDim.x * warps_per_block) {
    const int cnt = p.pool_final ? 1 : p.pool_count[w];
    for (int c = lane; c < p.feat; c += 32) {
      float v = p.pool[static_cast<long long>(w) * p.feat + c];
      if (p.pool_final) { /* finished features */ }
      else if (p.pool_mode == 1) v = cnt > 0 ? v : 0.0f;   // all-masked sample pools to zeros
      else v = p.masking ? (cnt > 0 ? v / fmaxf(static_cast<float>(cnt), 1e-7f) : 0.0f)
                         : v / static_cast<float>(cnt);
      feat[c] = v;
      if (p.emb && (p.mlp_hidden == 0 || emb_pooled)) p.emb[static_cast<long long>(w) * p.feat + c] = v;
    }
    if (p.mlp_hidden > 0) {
      // hidden Dense layers in front of the classifier, zero-padded to mlp_hidden == feat so the buffer can be reused:
      // the legacy head Dense(h, gelu) -> Dense(h, gelu) = "embedding" (nnlib/v1/layers.py:414-419), or the one / two hidden
      // layers of a v2 classification head (builder.py:589-596), whose "embedding" stays the pooled features
      __syncwarp();
      for (int layer = 0; layer < mlp_layers; ++layer) {
        const float* W = layer == 0 ? p.mlp_w1 : p.mlp_w2;
        const float* B = layer == 0 ? p.mlp_b1 : p.mlp_b2;
        float outv[8];                                   // up to 256 hidden units per warp
        for (int h = lane, i = 0; h < p.mlp_hidden; h += 32, ++i) {
          float acc = 0.0f;
          for (int c = 0; c < p.feat; ++c) acc = fmaf(feat[c], W[c * p.mlp_hidden + h], acc);
          outv[i] = act_apply(acc + B[h], p.mlp_act);
        }
        __syncwarp();
        for (int h = lane, i = 0; h < p.mlp_hidden; h += 32, ++i) feat[h] = outv[i];
        __syncwarp();
      }
      if (p.emb && !emb_pooled)
        for (int c = lane; c < p.feat; c += 32) p.emb[static_cast<long long>(w) * p.feat + c] = feat[c];
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
    // OOD signals of the window's logits (fp32 like the reference layer): running softmax statistics, all lanes alike
    const int n_sig = p.signals & 7;
    float z_max = -3.0e38f, z_second = -3.0e38f;
    float zs[kMaxSignalClasses];
    for (int k = 0; k < p.n_classes; ++k) {
      float acc = 0.0f;
      for (int c = lane; c < p.feat; c += 32) acc = fmaf(feat[c], p.cls_w[c * p.n_classes + k], acc);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      const float z = acc + p.cls_b[k];
      if (lane == 0) p.logits[static_cast<long long>(w) * p.n_classes + k] = z;
      if (n_sig && k < kMaxSignalClasses) {
        zs[k] = z;
        if (z > z_max) { z_second = z_max; z_max = z; } else if (z > z_second) { z_second = z; }
      }
    }
    float sig[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (n_sig) {
      float denom = 0.0f;
      for (int k = 0; k < p.n_classes && k < kMaxSignalClasses; ++k) denom += expf(zs[k] - z_max);
      float ent = 0.0f;
      for (int k = 0; k < p.n_classes && k < kMaxSignalClasses; ++k) {
        const float pr = fmaxf(expf(zs[k] - z_max) / denom, 1e-10f);
        ent -= pr * logf(pr);
      }
      float nn = 0.0f;
      for (int i = lane; i < nmd_dim; i += 32) nn = fmaf(nmdv[i], nmdv[i], nn);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, off);
      const float all[6] = {0.0f, 1.0f / denom, ent, z_max + logf(denom), (1.0f - expf(z_second - z_max)) / denom, sqrtf(nn)};
      for (int i = 0; i < n_sig; ++i) sig[i] = all[(p.signals >> (3 * (i + 1))) & 7];
    }
    if (p.rel && p.n_taps > 0) {
      float out = 0.0f;
      for (int h = 0; h < p.rel_hidden; ++h) {
        float acc = 0.0f;
        for (int i = lane; i < nmd_dim; i += 32) acc = fmaf(nmdv[i], p.rel_w1[i * p.rel_hidden + h], acc);
        if (lane < n_sig) acc = fmaf(sig[lane], p.rel_w1[(nmd_dim + lane) * p.rel_hidden + h], acc);
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


// NMD tap of the stem without touching the conv output.  The tap is the sum over a window's valid
// output rows of (conv output + bias) per channel (nnlib/v2/nmd.py:52-77 applied to the first
// masked_conv1d).  The stem's input is a one-hot token row, so that sum is linear in token counts:
//   tap[c] = sum_t sum_v H[t][v] * W[t][v][c] + (number of valid rows) * bias[c],
//   H[t][v] = #{positions i of the window's frames holding symbol v with 0 <= i - shift_t < limit}.
// A token that exists makes its rows valid in mask mode `any`, so H does not depend on the mask.
// Positions inside [max shift, limit + min shift) feed every tap (one histogram, weights summed
// over the taps); the few positions at the frame ends are counted per tap.  One CTA per window.
struct StemTapParams {
  const uint8_t* tokens; const int* lpad; const int* count;
  const float* w;      // [ntaps][64][cout], the fp16-rounded weights the MMA uses
  const float* wsum;   // [64][cout] = sum over taps
  const float* bias;   // [cout]
  float* tap;          // [n_windows][cout]
  long long n_windows;
  int lc, pitch, frames, tok_offset, ntaps, shrink, cout;
  int shifts[16];
};
__global__ void stem_tap_kernel(const StemTapParams p) {
  __shared__ int s_total[64];
  __shared__ int s_edge[16 * 64];
  int lo = p.shifts[0], mn = p.shifts[0];
  for (int t = 1; t < p.ntaps; ++t) { lo = p.shifts[t] > lo ? p.shifts[t] : lo; mn = p.shifts[t] < mn ? p.shifts[t] : mn; }
  for (long long w = blockIdx.x; w < p.n_windows; w += gridDim.x) {
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_total[i] = 0;
    for (int i = threadIdx.x; i < p.ntaps * 64; i += blockDim.x) s_edge[i] = 0;
    __syncthreads();
    const int lp = p.lpad[w];
    const int len = lp < p.lc ? lp : p.lc;
    const int limit = lp - p.shrink;
    const int hi = limit + mn;
    for (int idx = threadIdx.x; idx < p.frames * len; idx += blockDim.x) {
      const int f = idx / len, i = idx - f * len;
      const int ch = static_cast<int>(p.tokens[(w * p.frames + f) * p.pitch + i]) - p.tok_offset;
      if (ch < 0 || ch >= 64) continue;
      if (i >= lo && i < hi) {
        atomicAdd(&s_total[ch], 1);
      } else {
        for (int t = 0; t < p.ntaps; ++t) {
          const int j = i - p.shifts[t];
          if (j >= 0 && j < limit) atomicAdd(&s_edge[t * 64 + ch], 1);
        }
      }
    }
    __syncthreads();
    const float n_valid = static_cast<float>(p.count[w]);
    for (int c = threadIdx.x; c < p.cout; c += blockDim.x) {
      float acc = 0.0f;
      for (int v = 0; v < 64; ++v) {
        const int n = s_total[v];
        if (n) acc = fmaf(static_cast<float>(n), p.wsum[v * p.cout + c], acc);
      }
      for (int tv = 0; tv < p.ntaps * 64; ++tv) {
        const int n = s_edge[tv];
        if (n) acc = fmaf(static_cast<float>(n), p.w[static_cast<long long>(tv) * p.cout + c], acc);
      }
      p.tap[w * p.cout + c] = fmaf(n_valid, p.bias[c], acc);
    }
    __syncthreads();
  }
}

}  // namespace jg
