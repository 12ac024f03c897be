// Stage 3 hot kernel: masked dilated Conv1D as an implicit GEMM on the sm_100a
// tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by bulk TMA).
//
// Replaces, per layer, the reference's  MaskedConv1D -> MaskedBatchNorm -> Activation
// [-> MaskedAdd -> Activation] [-> NMDLayer] [-> MaskedBatchNorm -> Activation]
// [-> MaskedGlobalMaxPooling]  chain (reference: src/jaeger/nnlib/v2/layers.py:1217-1280,
// 918-941, 1882-1915, 517-529; src/jaeger/nnlib/v2/nmd.py:43-77).
//
// GEMM view of one tile:  D[128 rows, Cout] = sum_t  X[rows + shift_t, Cin] * W_t[Cin, Cout]
//   * A operand: one halo'd activation tile [Cin/8 planes][128+halo rows][8 ch] is brought
//     into shared memory once per 64-channel stage; every tap re-reads it through a
//     shared-memory descriptor whose start address is advanced by shift_t rows (the
//     no-swizzle K-major canonical layout makes a row shift a 16-byte address shift),
//     so the k-fold im2col re-read never leaves the SM.
//   * B operand: the whole layer's weights [ntaps*Cin/8][Cout][8] stay resident in
//     shared memory for the life of the (persistent) CTA.
//   * D: fp32 in TMEM, double buffered (2 x Cout columns) so the epilogue of tile i
//     overlaps the MMAs of tile i+1.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM lane quarter = warp % 4).
#pragma once
#include "conv_common.cuh"

namespace jg {
namespace tc {

constexpr int kThreads = 256;
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
// Bounded wait: a protocol bug must surface as a launch error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (err) atomicExch(err, code);
      __trap();
    }
  }
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

// K-major, no-swizzle ("interleaved") shared-memory matrix descriptor.
//   element (row r, k-chunk j of 8 bf16) lives at  start + (r%8)*16 + (r/8)*SBO + j*LBO
// With SBO = 128 the rows are linear at a 16-byte pitch, which is what lets a conv tap
// be expressed as  start += shift*16.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version 1 (sm_100)
  return d;         // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

// Column-wise reduction across the 32 lanes of a warp of a [32 lanes][32 columns]
// register tile.  Afterwards v[0] on lane l holds the reduction of column l.
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

struct SmemLayout {
  uint32_t w_off, stage_off, par_off, bar_off, total;
  uint32_t stage_bytes, plane_a, rows_a, ch_stage, stages_per_tile, w_bytes;
};

__host__ __device__ inline SmemLayout smem_layout(int cin, int cout, int ntaps, int halo_l,
                                                  int halo_r, int n_stages) {
  SmemLayout L;
  L.rows_a = kTileM + halo_l + halo_r;
  L.plane_a = L.rows_a * 16;
  L.ch_stage = (cin >= 64) ? 8 : (cin / 8);
  L.stages_per_tile = (cin / 8) / L.ch_stage;
  L.stage_bytes = L.ch_stage * L.plane_a;
  L.w_bytes = static_cast<uint32_t>(ntaps) * cin * cout * 2;
  L.w_off = 0;
  L.stage_off = (L.w_bytes + 127u) & ~127u;
  L.par_off = L.stage_off + n_stages * ((L.stage_bytes + 127u) & ~127u);
  L.bar_off = L.par_off + 6u * cout * 4u;
  L.total = L.bar_off + 256u;
  return L;
}

template <int kStages>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const SmemLayout L = smem_layout(p.cin, p.cout, p.ntaps, p.halo_l, p.halo_r, kStages);
  const uint32_t stage_pitch = (L.stage_bytes + 127u) & ~127u;

  float* s_par = reinterpret_cast<float*>(smem + L.par_off);  // scale1,shift1,scale2,shift2,bias,sc_const
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  // barrier slots: [0,S) full, [S,2S) empty, 2S wbar, 2S+1..2 tmem_full, 2S+3..4 tmem_empty
  const uint32_t bar0 = smem_u32(s_bar);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (kStages + s); };
  const uint32_t WBAR = bar0 + 8u * (2 * kStages);
  auto TFULL = [&](int a) { return bar0 + 8u * (2 * kStages + 1 + a); };
  auto TEMPTY = [&](int a) { return bar0 + 8u * (2 * kStages + 3 + a); };
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * kStages + 5);

  const uint32_t w_base = smem_u32(smem + L.w_off);
  const uint32_t st_base = smem_u32(smem + L.stage_off);

  // contiguous tile range for this CTA (neighbouring tiles share halo rows in L2)
  const int tile_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * p.n_tiles / gridDim.x);
  const int tile_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.n_tiles / gridDim.x);

  // ---- one-time setup ------------------------------------------------------------------
  for (int i = threadIdx.x; i < p.cout; i += kThreads) {
    s_par[i] = p.scale1[i];
    s_par[p.cout + i] = p.shift1[i];
    s_par[2 * p.cout + i] = p.has_affine2 ? p.scale2[i] : 1.0f;
    s_par[3 * p.cout + i] = p.has_affine2 ? p.shift2[i] : 0.0f;
    s_par[4 * p.cout + i] = p.bias ? p.bias[i] : 0.0f;
    s_par[5 * p.cout + i] = p.sc_const ? p.sc_const[i] : 0.0f;
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(FULL(s), 1);
      mbar_init(EMPTY(s), 1);
    }
    mbar_init(WBAR, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(TFULL(a), 1);
      mbar_init(TEMPTY(a), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem_cols = (2 * p.cout <= 32) ? 32u : (2 * p.cout <= 64) ? 64u
                           : (2 * p.cout <= 128) ? 128u : (2 * p.cout <= 256) ? 256u : 512u;
  if (warp == 2) tmem_alloc(smem_u32(s_tmem), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(WBAR, L.w_bytes);
      for (uint32_t off = 0; off < L.w_bytes; off += 32768u) {
        uint32_t n = L.w_bytes - off < 32768u ? L.w_bytes - off : 32768u;
        bulk_g2s(w_base + off, reinterpret_cast<const uint8_t*>(p.w) + off, n, WBAR);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const long long r_first = static_cast<long long>(tile) * kTileM - p.halo_l;
        for (uint32_t h = 0; h < L.stages_per_tile; ++h) {
          mbar_wait(EMPTY(s), ph ^ 1u, p.err, 1);
          mbar_expect_tx(FULL(s), L.stage_bytes);
          for (uint32_t c = 0; c < L.ch_stage; ++c) {
            const __nv_bfloat16* src =
                p.x + (static_cast<long long>(h * L.ch_stage + c) * p.x_plane + r_first) * 8;
            bulk_g2s(st_base + s * stage_pitch + c * L.plane_a, src, L.plane_a, FULL(s));
          }
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=cout, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                             (static_cast<uint32_t>(p.cout >> 3) << 17) |
                             (static_cast<uint32_t>(kTileM >> 4) << 24);
      mbar_wait(WBAR, 0, p.err, 2);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1u;
        mbar_wait(TEMPTY(as), aph ^ 1u, p.err, 3);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.cout);
        uint32_t accumulate = 0;
        for (uint32_t h = 0; h < L.stages_per_tile; ++h) {
          mbar_wait(FULL(s), ph, p.err, 4);
          tc_fence_after();
          const uint32_t a_stage = st_base + s * stage_pitch;
          for (int t = 0; t < p.ntaps; ++t) {
            const uint32_t a_tap = a_stage + static_cast<uint32_t>(p.halo_l + p.shifts[t]) * 16u;
            const uint32_t b_tap =
                w_base + static_cast<uint32_t>(t * (p.cin >> 3) + h * L.ch_stage) * (p.cout * 16u);
            for (uint32_t j = 0; j < L.ch_stage / 2; ++j) {
              const uint64_t adesc = make_desc(a_tap + 2 * j * L.plane_a, p.a_lbo, p.a_sbo);
              const uint64_t bdesc = make_desc(b_tap + 2 * j * (p.cout * 16u), p.b_lbo, p.b_sbo);
              umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit(EMPTY(s));  // frees the smem stage when these MMAs retire
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
        umma_commit(TFULL(as));   // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue: TMEM -> registers -> fused affine/residual/activation/taps -> HBM =====
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const float* s_scale1 = s_par;
    const float* s_shift1 = s_par + p.cout;
    const float* s_scale2 = s_par + 2 * p.cout;
    const float* s_shift2 = s_par + 3 * p.cout;
    const float* s_bias = s_par + 4 * p.cout;
    const float* s_scc = s_par + 5 * p.cout;
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1u;
      const long long row = static_cast<long long>(tile) * kTileM + q * 32 + lane;
      const int win = static_cast<int>((static_cast<long long>(tile) * kTileM) / p.rows_per_window);
      const bool valid = p.out_mask[row] != 0;
      const bool sc_valid = p.sc ? (p.sc_mask ? p.sc_mask[row] != 0 : true) : false;
      mbar_wait(TFULL(as), aph, p.err, 5);
      tc_fence_after();
      for (int cb = 0; cb < p.cout / 32; ++cb) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                      static_cast<uint32_t>(as * p.cout + cb * 32), raw);
        uint4 scv[4];
        if (p.sc && sc_valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            scv[j] = *reinterpret_cast<const uint4*>(
                p.sc + (static_cast<long long>(cb * 4 + j) * p.y_plane + row) * 8);
        }
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);

        if (p.tap_mode == 1) {
          float tv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) tv[j] = valid ? v[j] + s_bias[cb * 32 + j] : 0.0f;
          warp_cols_reduce<false>(tv, lane);
          atomicAdd(p.tap_sum + static_cast<long long>(win) * p.cout + cb * 32 + lane, tv[0]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], s_scale1[cb * 32 + j], s_shift1[cb * 32 + j]);
        if (p.sc) {
          if (sc_valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&scv[j]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = __bfloat1622float2(h2[e]);
                v[j * 8 + 2 * e] += f.x;
                v[j * 8 + 2 * e + 1] += f.y;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += s_scc[cb * 32 + j];
          }
        }
        if (p.act1 != ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = act_apply(v[j], p.act1);
        }
        if (p.tap_mode == 2) {
          float tv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) tv[j] = valid ? v[j] : 0.0f;
          warp_cols_reduce<false>(tv, lane);
          atomicAdd(p.tap_sum + static_cast<long long>(win) * p.cout + cb * 32 + lane, tv[0]);
        }
        if (p.has_affine2) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = act_apply(fmaf(v[j], s_scale2[cb * 32 + j], s_shift2[cb * 32 + j]), p.act2);
        }
        if (p.pool_mode != 0) {
          float tv[32];
          if (p.pool_mode == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) tv[j] = valid ? v[j] : -3.0e38f;
            warp_cols_reduce<true>(tv, lane);
            if (tv[0] > -1.0e38f)
              atomic_max_f32(p.pool + static_cast<long long>(win) * p.cout + cb * 32 + lane, tv[0]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) tv[j] = valid ? v[j] : 0.0f;
            warp_cols_reduce<false>(tv, lane);
            atomicAdd(p.pool + static_cast<long long>(win) * p.cout + cb * 32 + lane, tv[0]);
          }
        }
        if (p.y) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = valid ? v[j * 8 + 2 * e] : 0.0f;
              float b = valid ? v[j * 8 + 2 * e + 1] : 0.0f;
              h2[e] = __floats2bfloat162_rn(a, b);
            }
            *reinterpret_cast<uint4*>(p.y + (static_cast<long long>(cb * 4 + j) * p.y_plane + row) * 8) = o;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(TEMPTY(as));
    }
  }

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
