// Host-side launchers for the conv layer kernels.
#pragma once
#include "conv_ref.cuh"
#include "conv_tc.cuh"
#include "conv_tc2.cuh"
#include "conv_ws.cuh"

namespace jg {

constexpr unsigned kMaxSmem = 232448;  // 227 KB opt-in dynamic shared memory per CTA on sm_100

// Returns 0 on success, a negative code when the layer does not fit the kernel's constraints.
inline int conv_tc_stages(const ConvParams& p) {
  if (p.cin % 64 != 0 || p.cout % 32 != 0 || p.cout > 256 || p.ntaps > kMaxTaps) return -1;
  for (int s = 4; s >= 2; --s) {
    tc::SmemLayout L = tc::smem_layout(p.cin, p.cout, p.ntaps, p.halo_l, p.halo_r, s);
    if (L.total <= kMaxSmem) return s;
  }
  return -2;  // weights + 2 stages exceed shared memory
}

inline cudaError_t launch_conv_tc(const ConvParams& p, int num_sms, cudaStream_t stream) {
  const int stages = conv_tc_stages(p);
  if (stages < 0) return cudaErrorInvalidConfiguration;
  tc::SmemLayout L = tc::smem_layout(p.cin, p.cout, p.ntaps, p.halo_l, p.halo_r, stages);
  const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
  if (grid <= 0) return cudaSuccess;
  cudaError_t e;
#define JG_LAUNCH(S)                                                                            \
  e = cudaFuncSetAttribute(tc::conv_tc_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                           static_cast<int>(L.total));                                          \
  if (e != cudaSuccess) return e;                                                               \
  tc::conv_tc_kernel<S><<<grid, tc::kThreads, L.total, stream>>>(p);
  if (stages == 4) { JG_LAUNCH(4) }
  else if (stages == 3) { JG_LAUNCH(3) }
  else { JG_LAUNCH(2) }
#undef JG_LAUNCH
  return cudaGetLastError();
}

// CTA-pair kernel: eligible when Cout is a multiple of 64, the tile count is even and the
// per-CTA weights half + 4 stages (+ the output staging tiles of the 2-group variant) fit in
// shared memory.  groups = 2: staged bulk stores (light layers); groups = 3: direct stores.
inline bool conv_tc2_eligible(const ConvParams& p, int groups = 2) {
  if (p.cin % 64 != 0 || p.cout % 64 != 0 || p.cout > 256 || p.ntaps > kMaxTaps || (p.n_tiles & 1)) return false;
  if (groups == 3 && 4 * p.cout > 512) return false;      // three groups need at least three accumulators
  return tc2::smem_layout2(p.cin, p.cout, p.ntaps, p.halo_l, p.halo_r, groups == 2 ? 2 : 0).total <= kMaxSmem;
}

inline cudaError_t launch_conv_tc2(const ConvParams& p, int num_sms, cudaStream_t stream, int groups = 2) {
  if (!conv_tc2_eligible(p, groups)) return cudaErrorInvalidConfiguration;
  const tc2::SmemLayout2 L = tc2::smem_layout2(p.cin, p.cout, p.ntaps, p.halo_l, p.halo_r, groups == 2 ? 2 : 0);
  int pairs = num_sms / 2;
  if (pairs > p.n_tiles / 2) pairs = p.n_tiles / 2;
  if (pairs <= 0) return cudaSuccess;
  cudaError_t e;
  if (groups == 2) {
    e = cudaFuncSetAttribute(tc2::conv_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total));
    if (e != cudaSuccess) return e;
    tc2::conv_tc2_kernel<2><<<2 * pairs, tc2::kThreads2, L.total, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(tc2::conv_tc2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total));
    if (e != cudaSuccess) return e;
    tc2::conv_tc2_kernel<3><<<2 * pairs, tc2::kThreads2H, L.total, stream>>>(p);
  }
  return cudaGetLastError();
}

// Weights-stationary kernel (conv_ws.cuh): 128 output channels, weights within 320 TMEM columns, and one of the epilogue
// shapes it is specialised for.  Returns the ws::WsMode, or -1 when the layer has to run on another kernel.
inline int conv_ws_mode(const ConvParams& p) {
  if (p.ln1) return -1;      // MaskedLayerNormalization runs on the single-CTA kernel's generic epilogue
  if (p.mask_thr > 1) return -1;      // mask modes strict / majority: the round-1 kernels' validity helper counts the taps
  if (p.cout != 128 || (p.cin != 64 && p.cin != 128) || p.ntaps * p.cin / 2 > ws::kWColsMax || p.ntaps > kMaxTaps) return -1;
  if (!p.folded || p.dyt1 || p.dyt2 || p.act1 != ACT_GELU_TANH || p.rows_per_window % ws::kSubRows != 0 || p.n_tiles < 1) return -1;
  if (p.halo_l > 56 || p.halo_r > 56) return -1;
  if (ws::smem_layout_ws(p.cin, p.halo_l, p.halo_r).total > kMaxSmem) return -1;
  const bool has_sc = p.sc != nullptr;
  if (p.tap_mode == 0 && p.pool_mode == 0 && !p.has_affine2 && p.y != nullptr) return has_sc ? ws::WS_LIGHT_SC : ws::WS_LIGHT;
  if (has_sc && p.tap_mode == 2 && p.has_affine2 && p.act2 == ACT_GELU_TANH && p.tap_sum != nullptr) {
    if (p.pool_mode == 0 && p.y != nullptr) return ws::WS_FINAL;
    if (p.pool_mode == 1 && p.pool != nullptr) return ws::WS_FINAL_POOL;
  }
  return -1;
}

template <int kMode, int kTaps, int kGroups>
inline cudaError_t launch_conv_ws_inst(const ConvParams& p, int grid, unsigned smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(ws::conv_ws_kernel<kMode, kTaps, kGroups>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  ws::conv_ws_kernel<kMode, kTaps, kGroups><<<grid, ws::kThreadsWs, smem, stream>>>(p);
  return cudaGetLastError();
}

inline cudaError_t launch_conv_ws(const ConvParams& p, int num_sms, cudaStream_t stream) {
  const int mode = conv_ws_mode(p);
  if (mode < 0) return cudaErrorInvalidConfiguration;
  const ws::SmemLayoutWs L = ws::smem_layout_ws(p.cin, p.halo_l, p.halo_r);
  const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
  // the shapes that carry the time get an unrolled MMA issue loop: k5 x 128 channels (residual convs), k7 x 64 (stem)
  if (p.ntaps == 5 && p.cin == 128) {
    if (mode == ws::WS_LIGHT) return launch_conv_ws_inst<ws::WS_LIGHT, 5, 2>(p, grid, L.total, stream);
    if (mode == ws::WS_LIGHT_SC) return launch_conv_ws_inst<ws::WS_LIGHT_SC, 5, 2>(p, grid, L.total, stream);
    if (mode == ws::WS_FINAL) return launch_conv_ws_inst<ws::WS_FINAL, 5, 2>(p, grid, L.total, stream);
    return launch_conv_ws_inst<ws::WS_FINAL_POOL, 5, 2>(p, grid, L.total, stream);
  }
  if (p.ntaps == 7 && p.cin == 64 && mode == ws::WS_LIGHT) return launch_conv_ws_inst<ws::WS_LIGHT, 7, 1>(p, grid, L.total, stream);
  if (mode == ws::WS_LIGHT) return launch_conv_ws_inst<ws::WS_LIGHT, 0, 0>(p, grid, L.total, stream);
  if (mode == ws::WS_LIGHT_SC) return launch_conv_ws_inst<ws::WS_LIGHT_SC, 0, 0>(p, grid, L.total, stream);
  if (mode == ws::WS_FINAL) return launch_conv_ws_inst<ws::WS_FINAL, 0, 0>(p, grid, L.total, stream);
  return launch_conv_ws_inst<ws::WS_FINAL_POOL, 0, 0>(p, grid, L.total, stream);
}

inline cudaError_t launch_conv_ref(const ConvParams& p, cudaStream_t stream) {
  const long long total = static_cast<long long>(p.n_tiles) * kTileM * p.cout;
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  ref::conv_ref_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace jg
