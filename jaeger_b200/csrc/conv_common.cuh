// Shared definitions for the stage-3 (masked dilated Conv1D stack) kernels.
//
// Activation layout in HBM ("g64sw": channel groups of 64, rows pre-swizzled):
//     act[g][row][64]  fp16,   g = C/64,  row in [-GUARD, R + GUARD)
// inside a row's 128 bytes the eight 16-byte chunks are stored XOR-swizzled by the row
// number (chunk c of row r sits at chunk position c ^ (r & 7)).  That is exactly the
// shared-memory image the tensor core's SWIZZLE_128B K-major operand layout expects when
// a span of rows starting at a multiple of 8 is copied linearly to a 1024-byte aligned
// shared-memory address, so ONE bulk-TMA copy per 64-channel stage stages a whole
// halo'd A tile and no tensor map is needed.
// A "row" is one codon position of one frame of one window.  Frame f of window w
// lives at rows  w*rows_per_window + f*period + [0, L) ; every other row (the gap
// between frames, the tail of the window, the guard rows) is kept at zero, so a
// dilated tap that reaches past a frame edge reads the zero padding the reference
// gets from TF "SAME" padding (reference: src/jaeger/nnlib/v2/layers.py:1217-1280).
// With this layout a conv tap is a pure row shift of the operand descriptor.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace jg {

// Activations and conv weights are stored as IEEE fp16 (11 significand bits; fp32 accumulation in
// the tensor core, fp32 BatchNorm affine).  fp16 rather than bf16 because the epilogue can then
// run its GELU / residual / second affine as packed half2 arithmetic directly on the stored
// format, which halves its instruction count; conversions saturate at +-65504.
using act_t = __half;

constexpr int kTileM = 128;        // rows (positions) per MMA tile == TMEM lanes
constexpr int kMaxTaps = 16;
constexpr int kGuardRows = 64;     // zero rows before row 0 / after the last row of every plane

enum Act : int { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_RELU = 2, ACT_GELU_ERF = 3 };

// One conv layer launch.  All pointers are device pointers.
struct ConvParams {
  // tensors -------------------------------------------------------------
  const act_t* x;       // input, g64sw, points at row 0 of group 0
  act_t* y;             // output, g64sw (may alias sc), or nullptr (pool-only last layer)
  const act_t* sc;      // residual shortcut tensor (Cout channels) or nullptr
  const uint8_t* sc_mask;       // row mask the shortcut tensor was stored with (nullptr = all valid)
  const float* sc_const;        // [Cout] value of the shortcut at rows its mask zeroed
  const uint8_t* out_mask;      // [R] 1 = valid output row (in frame and mask-propagated)
  const act_t* w;       // weights, smem image (see w_index)
  const float* bias;            // [Cout] conv bias (used only when tap_raw)
  const float* scale1;          // [Cout] affine after conv (norm folded; bias folded unless tap_raw)
  const float* shift1;
  const float* scale2;          // [Cout] optional second affine (stand-alone norm after a stack)
  const float* shift2;
  float* tap_sum;               // [n_windows][Cout] masked column sums (NMD tap) or nullptr
  float* pool;                  // [n_windows][Cout] masked max (or sum) or nullptr
  // geometry ------------------------------------------------------------
  long long x_plane;            // plane stride of x in rows
  long long y_plane;            // plane stride of y / sc in rows
  int n_tiles;                  // R / 128
  int rows_per_window;          // multiple of 128
  int cin, cout, ntaps;
  int shifts[kMaxTaps];         // row shift of every tap (t*dilation - pad_left)
  int halo_l, halo_r;           // max(0,-min shift), max(0,max shift)
  // epilogue ------------------------------------------------------------
  int act1, act2;               // Act
  int has_affine2;
  int tap_mode;                 // 0 none, 1 raw conv output (acc+bias), 2 after act1, 3 after the second affine + act2
  int pool_mode;                // 0 none, 1 masked max, 2 masked sum
  // fused mask propagation (tensor-core kernels): when fuse_mask != 0 the epilogue derives the row's
  // validity itself -- in-frame test + OR over the taps of the INPUT mask (layers.py:1245-1252,
  // mode "any") -- writes it to out_mask_w and adds the tile's valid rows to count[window]
  const uint8_t* in_mask;       // [R] mask the input tensor was stored with (guard rows readable)
  uint8_t* out_mask_w;          // [R] mask of this launch's output
  const int* lpad;              // [n_windows] padded frame length of every window
  int* count;                   // [n_windows] valid output rows
  int fuse_mask, masking, period, frames, shrink_in, halvings, shrink;
  int len_round;           // frame length = (lpad - shrink_in + len_round) >> halvings: 0 = floor (MaxPool(2) stages), 2^halvings - 1 =
                           // ceil (SAME stride-2 convs, layers.py:1318-1320)
  int red_pitch;           // channel pitch of tap_sum / pool when this launch computes a SLICE of a wider layer's output channels
                           // (0 = cout); tap_sum / pool then point at the slice's first channel
  const float *dyt_g1, *dyt_b1, *dyt_g2, *dyt_b2;   // MaskedDYT gamma / beta after the tanh (alpha rides in scale1 / scale2)
  int dyt1, dyt2;          // the first / second norm is a MaskedDYT: y = gamma * tanh(scale * x + shift) + beta
  int folded;              // scale1 is folded into the weights (== 1): the specialised epilogues add shift1 only
  int epi_f32;             // generic epilogue entirely in fp32 with ONE rounding at the store (conv_epilogue.cuh): for graphs
                           // whose norm FOLLOWS the activation with a large scale (the legacy `default` model: BatchNorm after
                           // GELU with gamma / sigma up to 26), where a half-precision activation would be amplified
  int mask_thr;            // valid taps an output row needs under mask propagation: <= 1 "any", (k + 1) / 2 "majority", k "strict"
                           // (layers.py:1245-1252); honoured by the single-CTA / CTA-pair kernels and the CUDA-core path
  int ln1;                 // the first norm is a MaskedLayerNormalization (layers.py:337-367): per ROW, v = acc + bias is normalised over
                           // the layer's real channels, then scale1 (gamma) / shift1 (beta); single-CTA kernel, generic epilogue only
  float ln_eps, ln_inv_c;  // its epsilon and 1 / (real channel count): padded channels carry acc = bias = gamma = beta = 0
  int* err;                     // device int, set non-zero on a barrier time-out
  long long* dbg;               // optional per-tile clock64 trace of CTA 0 (probe only)
};

// element index of (row, channel c) inside a g64sw tensor whose groups are `plane` rows apart
__host__ __device__ __forceinline__ long long act_index(long long row, int c, long long plane) {
  const int g = c >> 6, cl = c & 63;
  const int chunk = (cl >> 3) ^ static_cast<int>(row & 7);
  return (static_cast<long long>(g) * plane + row) * 64 + chunk * 8 + (cl & 7);
}
// element index of weight (tap t, in-channel ci, out-channel co) inside the smem image:
// blocks [t][ci/64] of [cout rows][64 k] fp16, 16-byte chunks swizzled by (co & 7)
__host__ __device__ __forceinline__ long long w_index(int t, int ci, int co, int cin, int cout) {
  const int g = ci >> 6, cl = ci & 63;
  const int chunk = (cl >> 3) ^ (co & 7);
  return (static_cast<long long>(t * (cin >> 6) + g) * cout + co) * 64 + chunk * 8 + (cl & 7);
}

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_GELU_TANH) {
    // tf.nn.gelu(approximate=True): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float u = k0 * (v + k1 * v * v * v);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    return 0.5f * v * (1.0f + t);
  } else if (act == ACT_RELU) {
    return fmaxf(v, 0.0f);
  } else if (act == ACT_GELU_ERF) {
    return 0.5f * v * (1.0f + erff(v * 0.7071067811865476f));
  }
  return v;
}

// ---- packed half2 epilogue math ------------------------------------------------------------
__device__ __forceinline__ __half2 cvt_sat_h2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // hi <- b, lo <- a
  return *reinterpret_cast<__half2*>(&r);
}
// tanh-approximate GELU on two values: hx + hx * tanh(x * (k0 + k0*k1 * x^2)), hx = x/2
__device__ __forceinline__ __half2 gelu_tanh_h2(__half2 x) {
  const __half2 k0 = __float2half2_rn(0.7978845608028654f);
  const __half2 k01 = __float2half2_rn(0.7978845608028654f * 0.044715f);
  const __half2 u = __hmul2(x, __hfma2(__hmul2(x, x), k01, k0));
  uint32_t ui = *reinterpret_cast<const uint32_t*>(&u), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(ui));
  const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
  return __hfma2(hx, *reinterpret_cast<const __half2*>(&ti), hx);
}
template <int N>
__device__ __forceinline__ void act_apply_h2(__half2 (&h)[N], int act) {
  if (act == ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < N; ++j) h[j] = gelu_tanh_h2(h[j]);
  } else if (act == ACT_RELU) {
    const __half2 z = __float2half2_rn(0.0f);
#pragma unroll
    for (int j = 0; j < N; ++j) h[j] = __hmax2(h[j], z);
  } else if (act == ACT_GELU_ERF) {      // legacy graph only: exact erf form in fp32
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float2 f = __half22float2(h[j]);
      h[j] = cvt_sat_h2(act_apply(f.x, ACT_GELU_ERF), act_apply(f.y, ACT_GELU_ERF));
    }
  }
}

// Activation over a register array with the (warp-uniform) selector hoisted out of the loop,
// so only the selected formula is executed.
template <int N>
__device__ __forceinline__ void act_apply_vec(float (&v)[N], int act) {
  if (act == ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = act_apply(v[j], ACT_GELU_TANH);
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = fmaxf(v[j], 0.0f);
  } else if (act == ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = act_apply(v[j], ACT_GELU_ERF);
  }
}

__host__ __device__ __forceinline__ int red_pitch_of(const ConvParams& p) { return p.red_pitch ? p.red_pitch : p.cout; }

// float atomic max through the sign-split integer trick (no NaNs on this path)
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

}  // namespace jg
