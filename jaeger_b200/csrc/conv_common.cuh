// Shared definitions for the stage-3 (masked dilated Conv1D stack) kernels.
//
// Activation layout in HBM ("planar-8", channels split in chunks of 8):
//     act[c8][row][8]  bf16,   c8 = C/8,  row in [-GUARD, R + GUARD)
// A "row" is one codon position of one frame of one window.  Frame f of window w
// lives at rows  w*rows_per_window + f*period + [0, L) ; every other row (the gap
// between frames, the tail of the window, the guard rows) is kept at zero, so a
// dilated tap that reaches past a frame edge reads the zero padding the reference
// gets from TF "SAME" padding (reference: src/jaeger/nnlib/v2/layers.py:1217-1280).
// With this layout a conv tap is a pure row shift, each 8-channel plane of an
// A tile is one contiguous span (one bulk-TMA copy), and the epilogue's 16-byte
// stores are fully coalesced across the 32 lanes of a warp (lane == row).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace jg {

constexpr int kTileM = 128;        // rows (positions) per MMA tile == TMEM lanes
constexpr int kMaxTaps = 16;
constexpr int kGuardRows = 64;     // zero rows before row 0 / after the last row of every plane

enum Act : int { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_RELU = 2, ACT_GELU_ERF = 3 };

// One conv layer launch.  All pointers are device pointers.
struct ConvParams {
  // tensors -------------------------------------------------------------
  const __nv_bfloat16* x;       // input, planar-8, points at row 0 of plane 0
  __nv_bfloat16* y;             // output, planar-8 (may alias sc), or nullptr (pool-only last layer)
  const __nv_bfloat16* sc;      // residual shortcut tensor (Cout channels) or nullptr
  const uint8_t* sc_mask;       // row mask the shortcut tensor was stored with (nullptr = all valid)
  const float* sc_const;        // [Cout] value of the shortcut at rows its mask zeroed
  const uint8_t* out_mask;      // [R] 1 = valid output row (in frame and mask-propagated)
  const __nv_bfloat16* w;       // weights, smem image [ntaps*Cin/8][Cout][8]
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
  int tap_mode;                 // 0 none, 1 raw conv output (acc+bias), 2 after act1
  int pool_mode;                // 0 none, 1 masked max, 2 masked sum
  // tcgen05 operand descriptors (host-computed so a probe can vary them)
  unsigned a_lbo, a_sbo, b_lbo, b_sbo;   // bytes
  int* err;                     // device int, set non-zero on a barrier time-out
};

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

// float atomic max through the sign-split integer trick (no NaNs on this path)
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

}  // namespace jg
