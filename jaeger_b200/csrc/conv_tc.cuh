// Stage 3 hot kernel: masked dilated Conv1D as an implicit GEMM on the sm_100a
// tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by bulk TMA).
//
// Replaces, per layer, the reference's  MaskedConv1D -> MaskedBatchNorm -> Activation
// [-> MaskedAdd -> Activation] [-> NMDLayer] [-> MaskedBatchNorm -> Activation]
// [-> MaskedGlobalMaxPooling]  chain (reference: src/jaeger/nnlib/v2/layers.py:1217-1280,
// 918-941, 1882-1915, 517-529; src/jaeger/nnlib/v2/nmd.py:43-77).
//
// GEMM view of one tile:  D[128 rows, Cout] = sum_t  X[rows + shift_t, Cin] * W_t[Cin, Cout]
//   * A operand: one halo'd activation tile [8+128+8 rows][64 ch] (SWIZZLE_128B K-major
//     image, pre-swizzled in HBM) is brought into shared memory by ONE bulk-TMA copy per
//     64-channel stage; every tap re-reads it through a shared-memory descriptor whose
//     start address is advanced by shift_t rows (128 B each), so the k-fold im2col
//     re-read never leaves the SM.
//   * B operand: the whole layer's weights (blocks [tap][Cin/64] of [Cout][64], swizzled)
//     stay resident in shared memory for the life of the (persistent) CTA.
//   * D: fp32 in TMEM, a ring of 4 accumulators (4 x Cout columns) so the MMAs run up to
//     three tiles ahead of the epilogues.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warp 3 = validity helper (row masks a few tiles ahead, conv_epilogue.cuh),
// warps 4.. = kEpiGroups epilogue groups (TMEM lane quarter = warp % 4)
// that drain tiles round-robin, so the latency chain of one tile's epilogue (mask/shortcut
// loads, TMEM reads, stores) overlaps the next tile's.
#pragma once
#include "conv_epilogue.cuh"

namespace jg {
namespace tc {

#ifndef JG_EPI_GROUPS
#define JG_EPI_GROUPS 3
#endif
constexpr int kEpiGroups = JG_EPI_GROUPS;                // epilogue warpgroups (4 warps each)
constexpr int kThreads = 128 + 128 * kEpiGroups;         // 4 control warps + the epilogue warps
constexpr int kEpiWarp0 = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait, written as ONE opaque asm block so the compiler keeps treating the calling
// warp as converged (a C++ spin loop makes every later value "divergent" and forces the
// MMA descriptors through per-lane registers).  A protocol bug must surface as a launch
// error, never as a hung GPU: after ~2^24 failed (hardware-suspended) probes the warp traps.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .u32 n;\n"
      "mov.u32 n, 0;\n"
      "JG_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra JG_DONE;\n"
      "add.u32 n, n, 1;\n"
      "setp.lt.u32 p, n, 16777216;\n"
      "@p bra JG_WAIT;\n"
      "trap;\n"
      "JG_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows are 128 bytes (64 fp16) apart,
// 8-row groups 1024 bytes apart (SBO); the 16-byte chunks of a row are XOR-swizzled with
// address bits [7,10).  The hardware applies the XOR to the absolute shared-memory address,
// so advancing the start address by whole rows (a conv tap) or by 32 bytes (a K=16 step)
// keeps addressing the data a linear copy of the pre-swizzled HBM rows put there.
// hi word is constant; lo word = (addr >> 4) | LBO field, so stepping an operand by `bytes`
// is  lo += bytes >> 4  (the 14-bit address field cannot overflow below 256 KB).
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t saddr) {
  return ((saddr >> 4) & 0x3FFFu) | (1u << 16);
}
__device__ __forceinline__ uint64_t desc_pack(uint32_t lo, uint32_t hi) {
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
struct SmemLayout {
  uint32_t w_off, stage_off, par_off, bar_off, val_off, total;
  uint32_t stage_bytes, stage_pitch, rows_a, lead, groups, w_bytes;
};

__host__ __device__ inline SmemLayout smem_layout(int cin, int cout, int ntaps, int halo_l,
                                                  int halo_r, int n_stages) {
  SmemLayout L;
  L.lead = static_cast<uint32_t>((halo_l + 7) / 8 * 8);   // tile loads start at a multiple of 8 rows
  L.rows_a = L.lead + kTileM + static_cast<uint32_t>((halo_r + 7) / 8 * 8);
  L.groups = cin / 64;                                      // one stage per 64 input channels
  L.stage_bytes = L.rows_a * 128;
  L.stage_pitch = (L.stage_bytes + 1023u) & ~1023u;
  L.w_bytes = static_cast<uint32_t>(ntaps) * cin * cout * 2;
  L.w_off = 0;
  L.stage_off = (L.w_bytes + 1023u) & ~1023u;
  L.par_off = L.stage_off + n_stages * L.stage_pitch;
  L.bar_off = L.par_off + static_cast<uint32_t>(kEpiParFloats) * cout * 4u;
  L.val_off = L.bar_off + 256u;                             // validity ring: kVSlots x 128 bytes
  L.total = L.val_off + kVSlots * 128u + 1024u;             // + slack to align the base to 1024 B
  return L;
}

template <int kStages>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const SmemLayout L = smem_layout(p.cin, p.cout, p.ntaps, p.halo_l, p.halo_r, kStages);
  const uint32_t stage_pitch = L.stage_pitch;

  float* s_par = reinterpret_cast<float*>(smem + L.par_off);  // scale1,shift1,scale2,shift2,bias,sc_const
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  // barrier slots: [0,S) full, [S,2S) empty, 2S wbar, 2S+1..4 tmem_full, 2S+5..8 tmem_empty
  const uint32_t bar0 = smem_u32(s_bar);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (kStages + s); };
  const uint32_t WBAR = bar0 + 8u * (2 * kStages);
  auto TFULL = [&](int a) { return bar0 + 8u * (2 * kStages + 1 + a); };
  auto TEMPTY = [&](int a) { return bar0 + 8u * (2 * kStages + 5 + a); };
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * kStages + 9);
  auto VFULL = [&](int v) { return bar0 + 8u * (2 * kStages + 10 + v); };
  auto VEMPTY = [&](int v) { return bar0 + 8u * (2 * kStages + 10 + kVSlots + v); };
  volatile uint8_t* s_valid = smem + L.val_off;
  // accumulator ring in TMEM: 4 x cout columns when they fit the 512 columns, else 2
  const int n_acc = (4 * p.cout <= 512) ? 4 : 2;

  const uint32_t w_base = smem_u32(smem + L.w_off);
  const uint32_t st_base = smem_u32(smem + L.stage_off);

  // contiguous tile range for this CTA (neighbouring tiles share halo rows in L2)
  const int tile_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * p.n_tiles / gridDim.x);
  const int tile_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.n_tiles / gridDim.x);

  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[519] = clock64();
  // ---- one-time setup ------------------------------------------------------------------
  epi_params_fill(s_par, p, threadIdx.x, kThreads);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(FULL(s), 1);
      mbar_init(EMPTY(s), 1);
    }
    mbar_init(WBAR, 1);
    for (int a = 0; a < 4; ++a) {
      mbar_init(TFULL(a), 1);
      mbar_init(TEMPTY(a), 128);  // one epilogue group (4 warps) drains a tile
    }
    for (int v = 0; v < kVSlots; ++v) {
      mbar_init(VFULL(v), 1);
      mbar_init(VEMPTY(v), 4);    // one lane per warp of the draining group
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t acc_cols = static_cast<uint32_t>(n_acc * p.cout);
  const uint32_t tmem_cols = (acc_cols <= 32) ? 32u : (acc_cols <= 64) ? 64u
                           : (acc_cols <= 128) ? 128u : (acc_cols <= 256) ? 256u : 512u;
  if (warp == 2) tmem_alloc(smem_u32(s_tmem), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== TMA producer (whole warp stays converged; one elected lane issues) =====
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(WBAR, L.w_bytes);
      for (uint32_t off = 0; off < L.w_bytes; off += 32768u) {
        uint32_t n = L.w_bytes - off < 32768u ? L.w_bytes - off : 32768u;
        bulk_g2s(w_base + off, reinterpret_cast<const uint8_t*>(p.w) + off, n, WBAR);
      }
    }
    int s = 0;
    uint32_t ph = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const long long r_first = static_cast<long long>(tile) * kTileM - L.lead;
      for (uint32_t g = 0; g < L.groups; ++g) {
        const long long tw0 = p.dbg ? clock64() : 0;
        mbar_wait(EMPTY(s), ph ^ 1u);
        if (p.dbg && blockIdx.x == 0 && leader) p.dbg[520] += clock64() - tw0;      // producer: wait for a free stage
        if (leader) {
          mbar_expect_tx(FULL(s), L.stage_bytes);
          const act_t* src = p.x + (static_cast<long long>(g) * p.x_plane + r_first) * 64;
          bulk_g2s(st_base + s * stage_pitch, src, L.stage_bytes, FULL(s));
        }
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp converged, tcgen05.mma predicated on one elected lane) =====
    const bool leader = elect_one();
    // instruction descriptor: D=f32, A=B=f16, both K-major, N=cout, M=128
    const uint32_t idesc = (1u << 4) |                    // D = f32, A = B = f16 (format 0)
                           (static_cast<uint32_t>(p.cout >> 3) << 17) |
                           (static_cast<uint32_t>(kTileM >> 4) << 24);
    const uint32_t b_group_step = static_cast<uint32_t>(p.cout) * 8u;  // one [cout][64] block
    const uint32_t b_tap_step = b_group_step * L.groups;
    const uint32_t b_lo0 = desc_lo_sw128(w_base);
    mbar_wait(WBAR, 0);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int as = it % n_acc;
      const uint32_t aph = static_cast<uint32_t>(it / n_acc) & 1u;
      const long long tw1 = p.dbg ? clock64() : 0;
      mbar_wait(TEMPTY(as), aph ^ 1u);
      tc_fence_after();
      if (p.dbg && blockIdx.x == 0 && leader) p.dbg[521] += clock64() - tw1;        // MMA: wait for a free accumulator
      if (p.dbg && blockIdx.x == 0 && leader && it < 64) p.dbg[it * 8 + 0] = clock64();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.cout);
      uint32_t accumulate = 0;
      for (uint32_t g = 0; g < L.groups; ++g) {
        const long long tw2 = p.dbg ? clock64() : 0;
        mbar_wait(FULL(s), ph);
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && leader) p.dbg[522] += clock64() - tw2;      // MMA: wait for operands
        const uint32_t a_lo = desc_lo_sw128(st_base + s * stage_pitch + L.lead * 128u);
        const uint32_t b_lo = b_lo0 + g * b_group_step;
#pragma unroll 1
        for (int t = 0; t < p.ntaps; ++t) {
          const uint32_t a_tap = a_lo + static_cast<uint32_t>(p.shifts[t] * 8);  // 128 B / 16 per row
          const uint32_t b_tap = b_lo + static_cast<uint32_t>(t) * b_tap_step;
#pragma unroll
          for (uint32_t k16 = 0; k16 < 4; ++k16) {
            if (leader)
              umma_bf16(d_tmem, desc_pack(a_tap + k16 * 2u, kDescHiSw128),
                        desc_pack(b_tap + k16 * 2u, kDescHiSw128), idesc, accumulate);
            accumulate = 1;
          }
        }
        if (leader) umma_commit(EMPTY(s));  // frees the smem stage when these MMAs retire
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
      if (leader) umma_commit(TFULL(as));   // accumulator complete -> epilogue
      if (p.dbg && blockIdx.x == 0 && leader && it < 64) p.dbg[it * 8 + 1] = clock64();
    }
  } else if (warp == 3) {
    // ===== validity helper: row masks / window counts of this CTA's tiles, kVSlots tiles ahead =====
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int slot = it & (kVSlots - 1);
      mbar_wait(VEMPTY(slot), ((static_cast<uint32_t>(it) / kVSlots) & 1u) ^ 1u);
      const long long tile_row0 = static_cast<long long>(tile) * kTileM;
      tile_validity(p, tile_row0, lane, s_valid + slot * 128);
      __syncwarp();
      if (lane == 0) mbar_arrive(VFULL(slot));
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue: TMEM -> registers -> fused affine/residual/activation/taps -> HBM =====
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int grp = (warp - kEpiWarp0) >> 2;   // epilogue group: drains tiles with it % kEpiGroups == grp
    const int n_cb = p.cout / 32;
    const EpiParams ep = epi_params(s_par, p.cout);
    const bool has_sc = p.sc != nullptr;
    // A group may only wait on an accumulator barrier whose previous phase has surely completed:
    // that holds when no more groups rotate over the tiles than there are accumulators (Cout = 256
    // has two), so the surplus groups sit such a layer out.
    const int n_grp = n_acc < kEpiGroups ? n_acc : kEpiGroups;
    for (int tile = tile_begin + grp, it = grp; grp < n_grp && tile < tile_end; tile += n_grp, it += n_grp) {
      const int as = it % n_acc;
      const uint32_t aph = static_cast<uint32_t>(it / n_acc) & 1u;
      const long long row = static_cast<long long>(tile) * kTileM + q * 32 + lane;
      const int sw = static_cast<int>(row & 7);
      const int win = static_cast<int>((static_cast<long long>(tile) * kTileM) / p.rows_per_window);
      const int vslot = it & (kVSlots - 1);
      mbar_wait(VFULL(vslot), (static_cast<uint32_t>(it) / kVSlots) & 1u);
      const uint32_t vcode = s_valid[vslot * 128 + q * 32 + lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(VEMPTY(vslot));
      const bool valid = (vcode & 1u) != 0, sc_valid = (vcode & 2u) != 0;
      // the shortcut does not depend on the MMAs: fetch the first batch before waiting on them
      uint4 scv[4];
      if (sc_valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          scv[j] = *reinterpret_cast<const uint4*>(p.sc + row * 64 + ((j ^ sw) * 8));
      }
      if (p.dbg && blockIdx.x == 0 && q == 0 && lane == 0 && it < 64) p.dbg[it * 8 + 2] = clock64();
      const long long tw3 = p.dbg ? clock64() : 0;
      mbar_wait(TFULL(as), aph);
      tc_fence_after();
      if (p.dbg && blockIdx.x == 0 && q == 0 && lane == 0) p.dbg[524 + grp] += clock64() - tw3;   // epilogue group: wait for MMAs
      if (p.dbg && blockIdx.x == 0 && q == 0 && lane == 0 && it < 64) p.dbg[it * 8 + 3] = clock64();
      float ln_mu = 0.0f, ln_rs = 1.0f;
      if (p.ln1) {     // MaskedLayerNormalization: a first pass over the row's accumulators for the channel mean / variance
        float s1 = 0.0f, s2 = 0.0f;
        for (int cb = 0; cb < n_cb; ++cb) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.cout + cb * 32), raw);
          tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 c = ep.bias[cb * 8 + j4];
            const float v0 = __uint_as_float(raw[j4 * 4 + 0]) + c.x, v1 = __uint_as_float(raw[j4 * 4 + 1]) + c.y;
            const float v2 = __uint_as_float(raw[j4 * 4 + 2]) + c.z, v3 = __uint_as_float(raw[j4 * 4 + 3]) + c.w;
            s1 += (v0 + v1) + (v2 + v3);
            s2 = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, s2))));
          }
        }
        ln_mu = s1 * p.ln_inv_c;
        const float var = fmaxf(s2 * p.ln_inv_c - ln_mu * ln_mu, 0.0f);
        ln_rs = 1.0f / sqrtf(var + p.ln_eps);
      }
      for (int cb = 0; cb < n_cb; ++cb) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                      static_cast<uint32_t>(as * p.cout + cb * 32), raw);
        uint4 scc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) scc[j] = scv[j];
        if (sc_valid && cb + 1 < n_cb) {  // prefetch the next batch's shortcut
          const int nb = cb + 1;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            scv[j] = *reinterpret_cast<const uint4*>(
                p.sc + (static_cast<long long>(nb >> 1) * p.y_plane + row) * 64 +
                ((((nb & 1) * 4 + j) ^ sw) * 8));
        }
        tmem_ld_wait();
        if (cb + 1 == n_cb) {  // every TMEM read of this tile has landed: hand the accumulator back
          tc_fence_before();
          mbar_arrive(TEMPTY(as));
        }
        uint4 out[4];
        epilogue_batch(p, ep, cb, raw, scc, has_sc, sc_valid, valid, lane, win, out, ln_mu, ln_rs);
        if (p.y) {
          act_t* yrow = p.y + (static_cast<long long>(cb >> 1) * p.y_plane + row) * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(yrow + ((((cb & 1) * 4 + j) ^ sw) * 8)) = out[j];
        }
      }
      if (p.dbg && blockIdx.x == 0 && q == 0 && lane == 0 && it < 64) p.dbg[it * 8 + 4] = clock64();
    }
  }

  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[523] = clock64();
  // ---- teardown ------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace tc
}  // namespace jg
