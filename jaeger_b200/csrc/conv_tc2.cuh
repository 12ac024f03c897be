// Stage 3 hot kernel, CTA-pair version: the same masked dilated Conv1D implicit GEMM as
// conv_tc.cuh, but two CTAs of a cluster (one TPC) cooperate on a 256-row tile with
// tcgen05.mma.cta_group::2, and the epilogue leaves through shared memory + bulk-TMA stores.
//
// Why (measured on B200, profiles/conv_kernel_r1.md): with cta_group::1 an M=128 x N=128 MMA
// reads 8 KB of operands per 64 cycles -- all of the SM's 128 B/clk shared-memory bandwidth -- so
// any epilogue LSU traffic (row-strided 16-byte stores, parameter loads) stalls the tensor
// pipe (117 instead of 64 cycles per MMA).  In the pair each CTA feeds its own 128 rows of A but
// only HALF of the weights (B is split along N), i.e. 6 KB per MMA, the weights shrink to 80 KB
// per CTA, and the freed shared memory holds output staging rows so the stores become 4 KB bulk
// copies per warp instead of row-strided sector writes.
//
// Per CTA (rank r of the pair, rows (2*pt + r)*128 .. +128 of pair-tile pt):
//   warp 0  producer: bulk-TMA loads of its own halo'd A stage
//   warp 1  MMA issuer (leader CTA only): tcgen05.mma.cta_group::2, M=256 x N=Cout x K=16
//   warp 2  TMEM allocator (cta_group::2, both CTAs), then the validity helper (conv_epilogue.cuh)
//   warp 3  relay (peer CTA only): forwards "my stage landed" to the leader's FULL barriers
//   warps 4..  kG epilogue groups draining alternate tiles of this CTA's accumulator ring:
//              kG = 2 (light layers, staged bulk stores) or kG = 3 (tap / second-affine layers)
// Barriers: FULL/EMPTY per stage, TFULL/TEMPTY per accumulator, VFULL/VEMPTY per validity slot;
// tcgen05.commit multicasts EMPTY and TFULL to both CTAs, the peer's epilogue and relay arrive
// remotely on the leader's barriers.
#pragma once
#include <type_traits>
#include "conv_tc.cuh"

namespace jg {
namespace tc2 {

using namespace jg::tc;

constexpr int kThreads2 = 384;        // kG = 2 epilogue groups (light layers, staged bulk stores)
constexpr int kEpiGroups2 = 2;
constexpr int kThreads2H = 512;       // kG = 3 epilogue groups (heavy layers, direct stores, no staging tiles)
constexpr int kStages2 = 4;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) inside CTA `rank`
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(rank));
  return out;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct SmemLayout2 {
  uint32_t w_off, stage_off, out_off, par_off, bar_off, val_off, total;
  uint32_t stage_bytes, stage_pitch, rows_a, lead, groups, w_bytes, out_group_bytes, out_groups;
};

__host__ __device__ inline SmemLayout2 smem_layout2(int cin, int cout, int ntaps, int halo_l, int halo_r,
                                                    int staged_groups = kEpiGroups2) {
  SmemLayout2 L;
  L.lead = static_cast<uint32_t>((halo_l + 7) / 8 * 8);
  L.rows_a = L.lead + kTileM + static_cast<uint32_t>((halo_r + 7) / 8 * 8);
  L.groups = cin / 64;
  L.stage_bytes = L.rows_a * 128;
  L.stage_pitch = (L.stage_bytes + 1023u) & ~1023u;
  L.w_bytes = static_cast<uint32_t>(ntaps) * cin * (cout / 2) * 2;     // this CTA's half of the weights
  L.out_groups = cout / 64;
  L.out_group_bytes = kTileM * 128;                                     // 128 rows x 64 channels fp16
  L.w_off = 0;
  L.stage_off = (L.w_bytes + 1023u) & ~1023u;
  L.out_off = L.stage_off + kStages2 * L.stage_pitch;
  L.par_off = L.out_off + static_cast<uint32_t>(staged_groups) * L.out_groups * L.out_group_bytes;
  L.bar_off = L.par_off + static_cast<uint32_t>(kEpiParFloats) * cout * 4u;
  L.val_off = L.bar_off + 256u;                                         // validity ring: kVSlots x 128 bytes
  L.total = L.val_off + kVSlots * 128u + 1024u;
  return L;
}

// element index inside the pair weight image: half h (= CTA rank) holds output channels
// [h*cout/2, (h+1)*cout/2) as blocks [tap][cin/64] of [cout/2 rows][64 k], 16-byte chunks swizzled
// by the row; the two halves are stored back to back.
__host__ __device__ __forceinline__ long long w2_index(int t, int ci, int co, int cin, int cout, int ntaps) {
  const int half = cout / 2;
  const int h = co / half, n = co % half;
  const int g = ci >> 6, cl = ci & 63;
  const int chunk = (cl >> 3) ^ (n & 7);
  const long long half_elems = static_cast<long long>(ntaps) * cin * half;
  return h * half_elems + ((static_cast<long long>(t) * (cin >> 6) + g) * half + n) * 64 + chunk * 8 + (cl & 7);
}

// kG = 2: two epilogue groups, output staged in shared memory and bulk-stored (light layers).
// kG = 3: three epilogue groups at 128 registers, direct global stores and no staging tiles, rolled
//         epilogue only: for layers bound by epilogue instruction issue (NMD tap / second affine / pool).
template <int kG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 128 * kG, 1)
conv_tc2_kernel(const __grid_constant__ ConvParams p) {
  constexpr bool kStage = kG == 2;
  constexpr int kNThreads = 128 + 128 * kG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const SmemLayout2 L = smem_layout2(p.cin, p.cout, p.ntaps, p.halo_l, p.halo_r, kStage ? kG : 0);

  float* s_par = reinterpret_cast<float*>(smem + L.par_off);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  // barrier slots: [0,S) full, [S,2S) empty, 2S wbar, 2S+1..4 tmem_full, 2S+5..8 tmem_empty
  const uint32_t bar0 = smem_u32(s_bar);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (kStages2 + s); };
  const uint32_t WBAR = bar0 + 8u * (2 * kStages2);
  auto TFULL = [&](int a) { return bar0 + 8u * (2 * kStages2 + 1 + a); };
  auto TEMPTY = [&](int a) { return bar0 + 8u * (2 * kStages2 + 5 + a); };
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * kStages2 + 9);
  auto VFULL = [&](int v) { return bar0 + 8u * (2 * kStages2 + 10 + v); };
  auto VEMPTY = [&](int v) { return bar0 + 8u * (2 * kStages2 + 10 + kVSlots + v); };
  volatile uint8_t* s_valid = smem + L.val_off;
  const int n_acc = (4 * p.cout <= 512) ? 4 : 2;

  const uint32_t w_base = smem_u32(smem + L.w_off);
  const uint32_t st_base = smem_u32(smem + L.stage_off);
  const uint32_t out_base = smem_u32(smem + L.out_off);

  // pair-tiles (256 rows) handled by this pair: a contiguous range
  const int n_pairs = gridDim.x / 2, pair = blockIdx.x / 2;
  const int n_pt = p.n_tiles / 2;
  const int pt_begin = static_cast<int>(static_cast<long long>(pair) * n_pt / n_pairs);
  const int pt_end = static_cast<int>(static_cast<long long>(pair + 1) * n_pt / n_pairs);

  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[519] = clock64();
  // ---- one-time setup ------------------------------------------------------------------
  epi_params_fill(s_par, p, threadIdx.x, kNThreads);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(FULL(s), is_leader ? 2 : 1);     // leader: own producer + the peer's relay
      mbar_init(EMPTY(s), 1);
    }
    mbar_init(WBAR, is_leader ? 2 : 1);
    for (int a = 0; a < 4; ++a) {
      mbar_init(TFULL(a), 1);
      mbar_init(TEMPTY(a), 8);                   // one elected lane per epilogue warp (4) in each CTA
    }
    for (int v = 0; v < kVSlots; ++v) {
      mbar_init(VFULL(v), 1);
      mbar_init(VEMPTY(v), 4);                   // the 4 warps of the group that drains the tile
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t acc_cols = static_cast<uint32_t>(n_acc * p.cout);
  const uint32_t tmem_cols = (acc_cols <= 32) ? 32u : (acc_cols <= 64) ? 64u
                           : (acc_cols <= 128) ? 128u : (acc_cols <= 256) ? 256u : 512u;
  if (warp == 2) tmem_alloc2(smem_u32(s_tmem), tmem_cols);
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== producer: this CTA's weights half once, then its own A stages =====
    const bool leader_lane = elect_one();
    if (leader_lane) {
      mbar_expect_tx(WBAR, L.w_bytes);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + static_cast<size_t>(rank) * L.w_bytes;
      for (uint32_t off = 0; off < L.w_bytes; off += 32768u) {
        const uint32_t n = L.w_bytes - off < 32768u ? L.w_bytes - off : 32768u;
        bulk_g2s(w_base + off, wsrc + off, n, WBAR);
      }
    }
    int s = 0;
    uint32_t ph = 0;
    for (int pt = pt_begin; pt < pt_end; ++pt) {
      const long long r_first = (static_cast<long long>(pt) * 2 + rank) * kTileM - L.lead;
      for (uint32_t g = 0; g < L.groups; ++g) {
        const long long tw0 = p.dbg ? clock64() : 0;
        mbar_wait(EMPTY(s), ph ^ 1u);
        if (p.dbg && blockIdx.x == 0 && leader_lane) p.dbg[520] += clock64() - tw0;
        if (leader_lane) {
          mbar_expect_tx(FULL(s), L.stage_bytes);
          const act_t* src = p.x + (static_cast<long long>(g) * p.x_plane + r_first) * 64;
          bulk_g2s(st_base + s * L.stage_pitch, src, L.stage_bytes, FULL(s));
        }
        if (++s == kStages2) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (is_leader) {
      // ===== MMA issuer (leader CTA): one instruction drives both SMs' tensor cores =====
      const bool leader_lane = elect_one();
      const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(p.cout >> 3) << 17) |   // D f32, A/B f16
                             (static_cast<uint32_t>((2 * kTileM) >> 4) << 24);          // M = 256
      const uint32_t b_group_step = static_cast<uint32_t>(p.cout / 2) * 8u;              // one [cout/2][64] block (>>4)
      const uint32_t b_tap_step = b_group_step * L.groups;
      const uint32_t b_lo0 = desc_lo_sw128(w_base);
      mbar_wait(WBAR, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt, ++it) {
        const int as = it % n_acc;
        const uint32_t aph = static_cast<uint32_t>(it / n_acc) & 1u;
        const long long tw1 = p.dbg ? clock64() : 0;
        mbar_wait(TEMPTY(as), aph ^ 1u);
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && leader_lane) p.dbg[521] += clock64() - tw1;
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.cout);
        uint32_t accumulate = 0;
        for (uint32_t g = 0; g < L.groups; ++g) {
          const long long tw2 = p.dbg ? clock64() : 0;
          mbar_wait(FULL(s), ph);
          tc_fence_after();
          if (p.dbg && blockIdx.x == 0 && leader_lane) p.dbg[522] += clock64() - tw2;
          const uint32_t a_lo = desc_lo_sw128(st_base + s * L.stage_pitch + L.lead * 128u);
          const uint32_t b_lo = b_lo0 + g * b_group_step;
#pragma unroll 1
          for (int t = 0; t < p.ntaps; ++t) {
            const uint32_t a_tap = a_lo + static_cast<uint32_t>(p.shifts[t] * 8);
            const uint32_t b_tap = b_lo + static_cast<uint32_t>(t) * b_tap_step;
#pragma unroll
            for (uint32_t k16 = 0; k16 < 4; ++k16) {
              if (leader_lane)
                umma_bf16_2cta(d_tmem, desc_pack(a_tap + k16 * 2u, kDescHiSw128),
                               desc_pack(b_tap + k16 * 2u, kDescHiSw128), idesc, accumulate);
              accumulate = 1;
            }
          }
          if (leader_lane) umma_commit_pair(EMPTY(s));
          if (++s == kStages2) { s = 0; ph ^= 1u; }
        }
        if (leader_lane) umma_commit_pair(TFULL(as));
      }
    }
  } else if (warp == 2) {
    // ===== validity helper: row masks / window counts of this CTA's tiles, kVSlots tiles ahead =====
    int it = 0;
    for (int pt = pt_begin; pt < pt_end; ++pt, ++it) {
      const int slot = it & (kVSlots - 1);
      mbar_wait(VEMPTY(slot), ((static_cast<uint32_t>(it) / kVSlots) & 1u) ^ 1u);
      const long long tile_row0 = (static_cast<long long>(pt) * 2 + rank) * kTileM;
      tile_validity(p, tile_row0, lane, s_valid + slot * 128);
      __syncwarp();
      if (lane == 0) mbar_arrive(VFULL(slot));
    }
  } else if (warp == 3) {
    if (!is_leader) {
      // ===== relay (peer CTA): tell the leader when this CTA's operands have landed =====
      const bool leader_lane = elect_one();
      mbar_wait(WBAR, 0);
      if (leader_lane) mbar_arrive_cluster(map_to_cta(WBAR, 0));
      int s = 0;
      uint32_t ph = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        for (uint32_t g = 0; g < L.groups; ++g) {
          mbar_wait(FULL(s), ph);
          if (leader_lane) mbar_arrive_cluster(map_to_cta(FULL(s), 0));
          if (++s == kStages2) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue: TMEM -> registers -> fused math -> smem staging -> bulk-TMA store =====
    // Every warp works alone: it owns 32 rows of the tile, i.e. 4 KB contiguous per 64-channel
    // group both in the staging tile and in HBM, stages them and issues its own bulk stores.
    // TMEM loads run one batch ahead of the math.
    const int q = warp & 3;
    const int grp = (warp - kEpiWarp0) >> 2;
    const int n_cb = p.cout / 32;
    const EpiParams ep = epi_params(s_par, p.cout);
    const bool has_sc = p.sc != nullptr;
    const bool light = p.folded && p.tap_mode == 0 && p.pool_mode == 0 && !p.has_affine2 && p.act1 == ACT_GELU_TANH;
    const bool final_shape = p.folded && p.dyt1 == p.dyt2 && has_sc && p.tap_mode == 2 && p.has_affine2 && p.act1 == ACT_GELU_TANH && p.act2 == ACT_GELU_TANH;
    const int epi_mode = final_shape ? (p.pool_mode == 0 ? EPI_FINAL : (p.pool_mode == 1 ? EPI_FINAL_POOL : EPI_GENERIC)) : EPI_GENERIC;
    const uint32_t warp_stage = out_base + grp * L.out_groups * L.out_group_bytes + static_cast<uint32_t>(q) * 32u * 128u;
    for (int pt = pt_begin + grp, it = grp; pt < pt_end; pt += kG, it += kG) {
      const int as = it % n_acc;
      const uint32_t aph = static_cast<uint32_t>(it / n_acc) & 1u;
      const bool tr_on = p.dbg && blockIdx.x == 0 && grp == 0 && q == 0 && lane == 0 && (it / kG) < 24;
      long long* tr = tr_on ? p.dbg + 600 + 16 * (it / kG) : nullptr;
      if (tr_on) tr[0] = clock64();
      const long long tile_row0 = (static_cast<long long>(pt) * 2 + rank) * kTileM;
      const int row_in_tile = q * 32 + lane;
      const long long row = tile_row0 + row_in_tile;
      const int sw = static_cast<int>(row & 7);
      const int win = static_cast<int>(tile_row0 / p.rows_per_window);
      const int vslot = it & (kVSlots - 1);
      mbar_wait(VFULL(vslot), (static_cast<uint32_t>(it) / kVSlots) & 1u);
      const uint32_t vcode = s_valid[vslot * 128 + row_in_tile];
      __syncwarp();
      if (lane == 0) mbar_arrive(VEMPTY(vslot));
      const bool valid = (vcode & 1u) != 0, sc_valid = (vcode & 2u) != 0;
      uint4 scv[4];
      if (sc_valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) scv[j] = *reinterpret_cast<const uint4*>(p.sc + row * 64 + ((j ^ sw) * 8));
      }
      const long long tw3 = p.dbg ? clock64() : 0;
      if (tr_on) tr[1] = tw3;
      mbar_wait(TFULL(as), aph);
      tc_fence_after();
      if (tr_on) tr[2] = clock64();
      if (p.dbg && blockIdx.x == 0 && q == 0 && lane == 0) p.dbg[524 + grp] += clock64() - tw3;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.cout);

      // one 32-channel batch whose accumulators are already in `raw`
      auto batch = [&](int cb, const uint32_t (&raw)[32], auto mode_tag) {
        using Tag = decltype(mode_tag);
        uint4 scc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) scc[j] = scv[j];
        if (sc_valid && cb + 1 < n_cb) {
          const int nb = cb + 1;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            scv[j] = *reinterpret_cast<const uint4*>(p.sc + (static_cast<long long>(nb >> 1) * p.y_plane + row) * 64 +
                                                     ((((nb & 1) * 4 + j) ^ sw) * 8));
        }
        uint4 out[4];
        epilogue_batch<Tag::kMode, Tag::kD1, Tag::kD2>(p, ep, cb, raw, scc, has_sc, sc_valid, valid, lane, win, out);
        if constexpr (kStage) {
          if (p.y) {
            const int og = cb >> 1;
            if ((cb & 1) == 0) {
              // this warp's previous store from the same staging rows must have finished READING them
              if (lane == 0) {
                if (L.out_groups == 2) bulk_wait_read<1>();
                else if (L.out_groups == 4) bulk_wait_read<3>();
                else bulk_wait_read<0>();
              }
              __syncwarp();
            }
            // staging rows = the exact g64sw image of rows [tile_row0 + 32 q, +32) of channel group og
            const uint32_t srow = warp_stage + og * L.out_group_bytes + static_cast<uint32_t>(lane) * 128u;
  #pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t a = srow + ((((cb & 1) * 4 + j) ^ sw) * 16);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(out[j].x), "r"(out[j].y), "r"(out[j].z), "r"(out[j].w) : "memory");
            }
            if (cb & 1) {   // 32 rows x 64 channels staged: 4 KB, contiguous in HBM
              fence_async_smem();
              __syncwarp();
              if (lane == 0) {
                bulk_s2g(p.y + (static_cast<long long>(og) * p.y_plane + tile_row0 + q * 32) * 64, warp_stage + og * L.out_group_bytes, 32u * 128u);
                bulk_commit();
              }
            }
          }
        } else if (p.y) {
          act_t* yrow = p.y + (static_cast<long long>(cb >> 1) * p.y_plane + row) * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(yrow + ((((cb & 1) * 4 + j) ^ sw) * 8)) = out[j];
        }
        if (tr_on && cb < 4) tr[4 + 2 * cb] = clock64();
      };
      // all TMEM reads of this tile landed: hand the accumulator back to the leader.  No generic-proxy
      // data is published with it, so the arrive is relaxed (a release here costs a full membar).
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(map_to_cta(TEMPTY(as), 0));
      };

      // light layers (conv1 / conv2 of a residual block): unrolled, TMEM loads one batch ahead
      auto light_tile = [&](auto tag) {
        uint32_t r0[32], r1[32], r2[32], r3[32];
        tmem_ld32(t_addr, r0);
        tmem_ld_wait();
        if (tr_on) tr[3] = clock64();
        tmem_ld32(t_addr + 32u, r1);
        batch(0, r0, tag);
        tmem_ld_wait();
        if (tr_on) tr[5] = clock64();
        tmem_ld32(t_addr + 64u, r2);
        batch(1, r1, tag);
        tmem_ld_wait();
        if (tr_on) tr[7] = clock64();
        tmem_ld32(t_addr + 96u, r3);
        batch(2, r2, tag);
        tmem_ld_wait();
        if (tr_on) tr[9] = clock64();
        release_acc();
        batch(3, r3, tag);
      };
      if (kStage && n_cb == 4 && light) {
        if (p.dyt1) light_tile(EpiTag<EPI_LIGHT, true>{}); else light_tile(EpiTag<EPI_LIGHT>{});
      } else {
        // everything else: rolled copies of the full epilogue (unrolled it would not fit the instruction
        // cache), specialised for the two block-final shapes, generic otherwise
        auto rolled = [&](auto mode_tag) {
          for (int cb = 0; cb < n_cb; ++cb) {
            uint32_t r0[32];
            tmem_ld32(t_addr + static_cast<uint32_t>(cb * 32), r0);
            tmem_ld_wait();
            if (tr_on && cb < 4) tr[3 + 2 * cb] = clock64();
            if (cb + 1 == n_cb) release_acc();
            batch(cb, r0, mode_tag);
          }
        };
        if (epi_mode == EPI_FINAL) { if (p.dyt1) rolled(EpiTag<EPI_FINAL, true, true>{}); else rolled(EpiTag<EPI_FINAL>{}); }
        else if (epi_mode == EPI_FINAL_POOL) { if (p.dyt1) rolled(EpiTag<EPI_FINAL_POOL, true, true>{}); else rolled(EpiTag<EPI_FINAL_POOL>{}); }
        else rolled(EpiTag<EPI_GENERIC>{});
      }
      if (tr_on) tr[11] = clock64();
    }
    if (kStage && p.y && lane == 0) bulk_wait_all();      // stores complete before the kernel ends
  }

  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[523] = clock64();
  // ---- teardown ------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                 // the peer may still be signalling our barriers / TMEM pair
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, tmem_cols);
  }
}

}  // namespace tc2
}  // namespace jg
