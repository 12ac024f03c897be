// Host-side launchers for the conv layer kernels.
#pragma once
#include "conv_ref.cuh"
#include "conv_tc.cuh"
#include "conv_tc2.cuh"

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

inline cudaError_t launch_conv_ref(const ConvParams& p, cudaStream_t stream) {
  const long long total = static_cast<long long>(p.n_tiles) * kTileM * p.cout;
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  ref::conv_ref_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace jg
