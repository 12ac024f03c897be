// Stages 1-2 of the hot path: ASCII -> 2-bit pack, and fragment windowing + N-masking +
// six-frame translation + codon encoding + base counts in one kernel.  HBM-bound integer work:
// coalesced vector loads, shared-memory staging of the window's packed bases, word-wide token
// stores.  Reference behaviour: seqops/io.py:103-133, seqops/encode.py:229-302,
// preprocess/v1/convert.py:75-99 (paths relative to /root/reference/src/jaeger).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jg {

// ---- stage 1 -------------------------------------------------------------------------------
// One thread packs 32 bases: two 16-byte loads -> one 64-bit code word + one validity word.
// A=0 C=1 T=2 G=3 is (ascii >> 1) & 3 on the case-folded letter; complement = code ^ 2.
__global__ void pack_bases_kernel(const uint8_t* __restrict__ ascii, long long n,
                                  uint32_t* __restrict__ codes, uint32_t* __restrict__ valid) {
  const long long n_groups = (n + 31) / 32;
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < n_groups;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long base = g * 32;
    uint32_t bytes[8];
    if (base + 32 <= n && ((reinterpret_cast<uintptr_t>(ascii) + base) & 15) == 0) {
      const uint4 a = *reinterpret_cast<const uint4*>(ascii + base);
      const uint4 b = *reinterpret_cast<const uint4*>(ascii + base + 16);
      bytes[0] = a.x; bytes[1] = a.y; bytes[2] = a.z; bytes[3] = a.w;
      bytes[4] = b.x; bytes[5] = b.y; bytes[6] = b.z; bytes[7] = b.w;
    } else {
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const long long i = base + w * 4 + k;
          v |= static_cast<uint32_t>(i < n ? ascii[i] : 0) << (8 * k);
        }
        bytes[w] = v;
      }
    }
    uint32_t lo = 0, hi = 0, vmask = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const uint32_t c = (bytes[i >> 2] >> (8 * (i & 3))) & 0xDFu;  // fold case
      const uint32_t code = (c >> 1) & 3u;
      const bool ok = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
      if (i < 16) lo |= code << (2 * i); else hi |= code << (2 * (i - 16));
      vmask |= static_cast<uint32_t>(ok) << i;
    }
    codes[2 * g] = lo;
    codes[2 * g + 1] = hi;
    valid[g] = vmask;
  }
}

// ---- stage 2 -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t get_code(const uint32_t* s_codes, int i) {
  return (s_codes[i >> 4] >> (2 * (i & 15))) & 3u;
}
__device__ __forceinline__ uint32_t get_bit(const uint32_t* s_bits, int i) {
  return (s_bits[i >> 5] >> (i & 31)) & 1u;
}

// Exact emulation of Python's round(x, 2) for x = n/d (d > 0), returning hundredths.
// Python rounds the double x = fl(n/d) correctly (half-even on the exact binary value); the
// product 100*x is recovered exactly as p + e with an FMA, so the comparison with k + 0.5 is
// exact.  (reference: utils/misc.py:117-123 safe_divide, seqops/io.py:128,133)
__device__ __forceinline__ int round_hundredths(int num, int den) {
  if (den == 0) return 0;
  const double x = fabs(static_cast<double>(num) / static_cast<double>(den));
  const double p = x * 100.0;
  const double e = fma(x, 100.0, -p);
  const double k = floor(p);
  double r = (p - k) - 0.5;  // exact: both are multiples of ulp(p) and close together
  // sign of (r + e) decides; a tie needs r + e == 0 exactly
  const double s = r + e;
  int kk = static_cast<int>(k);
  if (s > 0.0) kk += 1;
  else if (s == 0.0 && (r == -e)) kk += (kk & 1);  // half -> even
  return num < 0 ? -kk : kk;
}

constexpr int kEncThreads = 128;
struct CodonLut { uint8_t v[64]; };   // codon (first<<4 | second<<2 | third, 2-bit codes) -> token

// One WARP per window (four windows in flight per CTA, no block-wide barrier): the warp's slice of
// shared memory holds the window's packed codes / validity / soft-mask re-based to bit 0, so a
// codon is one funnel shift.  Tokens leave as 32-bit words, the lanes of a warp writing consecutive
// words of a frame.
constexpr int kEncWarps = kEncThreads / 32;
__global__ void __launch_bounds__(kEncThreads)
encode_windows_kernel(const uint32_t* __restrict__ codes, const uint32_t* __restrict__ valid,
                      const uint32_t* __restrict__ soft, const long long* __restrict__ win_base,
                      const int* __restrict__ win_nbases, long long n_windows, int crop, int lc,
                      int pitch, const CodonLut lut64, int case_sensitive,
                      uint8_t* __restrict__ tokens, int* __restrict__ counts,
                      short* __restrict__ skew100) {
  extern __shared__ uint32_t s_mem[];
  const int words_c = (crop + 15) / 16 + 1;
  const int words_b = (crop + 31) / 32 + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* s_codes = s_mem + warp * (words_c + 2 * words_b);
  uint32_t* s_valid = s_codes + words_c;   // valid for tokens (soft-masked removed if case_sensitive)
  uint32_t* s_count = s_valid + words_b;   // valid & ~soft (what the G/C/A/T counts see)
  __shared__ uint8_t s_lut2[128];     // [0, 64): forward frames, index = first | (second << 2) | (third << 4)
  uint8_t* const s_lut_fwd = s_lut2;  // [64, 128): reverse frames, index = (first base << 4) | (second << 2) | third
  uint8_t* const s_lut = s_lut2 + 64;
  if (threadIdx.x < 64) {
    const uint32_t i = threadIdx.x;
    s_lut[i] = lut64.v[i];
    s_lut_fwd[i] = lut64.v[((i & 3u) << 4) | (i & 12u) | (i >> 4)];
  }
  __syncthreads();
  // encode.py:232-236: the slice offset comes from crop % 3, not from the true length.
  const int off = (crop % 3 == 0) ? -2 : (crop % 3 == 1 ? -1 : 0);
  const int words_per_frame = pitch / 4;       // pitch is a multiple of 4: word-wide stores

  for (long long w = static_cast<long long>(blockIdx.x) * kEncWarps + warp; w < n_windows;
       w += static_cast<long long>(gridDim.x) * kEncWarps) {
    const long long base = win_base[w];
    int n = win_nbases[w];
    if (n > crop) n = crop;
    // re-based copies: code word j holds bases [16j, 16j+16) of the window
    const int sh_c = static_cast<int>(base & 15) * 2;
    const long long w0_c = base >> 4;
    // loads are bounded by the window's own length (a short contig at the end of the packed buffer must not read
    // crop bases past it): word j is only fetched when it holds a base of the window, the rest is zero-filled
    for (int j = lane; j < words_c; j += 32) {
      uint32_t a = 0u, b = 0u;
      if (j * 16 < n) { a = codes[w0_c + j]; if (sh_c) b = codes[w0_c + j + 1]; }
      s_codes[j] = sh_c ? ((a >> sh_c) | (b << (32 - sh_c))) : a;
    }
    const int sh_b = static_cast<int>(base & 31);
    const long long w0_b = base >> 5;
    for (int j = lane; j < words_b; j += 32) {
      const bool in_window = j * 32 < n;
      uint32_t a = 0u, b = 0u;
      if (in_window) { a = valid[w0_b + j]; if (sh_b) b = valid[w0_b + j + 1]; }
      uint32_t v = sh_b ? ((a >> sh_b) | (b << (32 - sh_b))) : a;
      uint32_t sm = 0;
      if (soft && in_window) {
        const uint32_t c = soft[w0_b + j], d = sh_b ? soft[w0_b + j + 1] : 0u;
        sm = sh_b ? ((c >> sh_b) | (d << (32 - sh_b))) : c;
      }
      // clip to the window length
      const int first = j * 32;
      const uint32_t keep = (first + 32 <= n) ? 0xFFFFFFFFu : (first >= n ? 0u : ((1u << (n - first)) - 1u));
      v &= keep;
      s_count[j] = v & ~sm;
      s_valid[j] = case_sensitive ? (v & ~sm) : v;
    }
    __syncwarp();

    // ---- base counts (upper-case, un-masked A/C/G/T) ------------------------------------
    int cg = 0, cc = 0, ca = 0, ct = 0;
    for (int j = lane; j * 16 < n; j += 32) {
      const uint32_t cw = s_codes[j];
      const uint32_t vb = (s_count[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
      // spread the 16 validity bits to the low bit of each 2-bit lane
      uint32_t m = vb;
      m = (m | (m << 8)) & 0x00FF00FFu;
      m = (m | (m << 4)) & 0x0F0F0F0Fu;
      m = (m | (m << 2)) & 0x33333333u;
      m = (m | (m << 1)) & 0x55555555u;
      const uint32_t lo = cw & 0x55555555u, hi = (cw >> 1) & 0x55555555u;
      ca += __popc(m & ~lo & ~hi);  // 00
      cc += __popc(m & lo & ~hi);   // 01
      ct += __popc(m & ~lo & hi);   // 10
      cg += __popc(m & lo & hi);    // 11
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      cg += __shfl_xor_sync(0xffffffffu, cg, o);
      cc += __shfl_xor_sync(0xffffffffu, cc, o);
      ca += __shfl_xor_sync(0xffffffffu, ca, o);
      ct += __shfl_xor_sync(0xffffffffu, ct, o);
    }

    // ---- tokens: frames f1,f2,f3 (forward) r1,r2,r3 (reverse complement) ------------------
    int nc = (n - 5 + off + 2) / 3;          // ceil((n - 5 + off) / 3)
    if (n - 5 + off <= 0) nc = 0;
    if (nc > lc) nc = lc;
    uint32_t* out_words = reinterpret_cast<uint32_t*>(tokens + w * 6ll * pitch);
    // the 6 x words_per_frame token words of the window are one index space for the warp's lanes
    for (int idx = lane, f = 0, wd = lane; idx < 6 * words_per_frame; idx += 32, wd += 32) {
      while (wd >= words_per_frame) { wd -= words_per_frame; ++f; }
      {
        const int j0 = wd * 4;
        uint32_t packed = 0;
        if (j0 + 3 < nc) {
          // four whole codons = 12 contiguous bases: one funnel shift brings their 24 code bits, one their 12
          // validity bits; forward frames hold the word's first codon lowest, reverse frames highest
          const int b0 = (f < 3) ? f + 3 * j0 : n - 3 - (f - 3) - 3 * (j0 + 3);
          const int cb = 2 * b0, cw = cb >> 5, vw = b0 >> 5;
          uint32_t bits = __funnelshift_r(s_codes[cw], s_codes[cw + 1], cb & 31);
          const uint32_t vb = __funnelshift_r(s_valid[vw], s_valid[vw + 1], b0 & 31) & 0xFFFu;
          const uint8_t* lut = s_lut2 + ((f < 3) ? 0 : 64);
          if (f >= 3) bits ^= 0xAAAAAAu;                   // complement of all 12 bases
          const uint32_t t0 = lut[bits & 63u], t1 = lut[(bits >> 6) & 63u], t2 = lut[(bits >> 12) & 63u], t3 = lut[(bits >> 18) & 63u];
          uint32_t m = 0xFFFFFFFFu;
          if (vb != 0xFFFu)                                // an N / soft-masked base in the word: blank its codons
            m = ((vb & 7u) == 7u ? 0xFFu : 0u) | ((vb & 0x38u) == 0x38u ? 0xFF00u : 0u) |
                ((vb & 0x1C0u) == 0x1C0u ? 0xFF0000u : 0u) | ((vb & 0xE00u) == 0xE00u ? 0xFF000000u : 0u);
          const uint32_t asc = t0 | (t1 << 8) | (t2 << 16) | (t3 << 24);       // codons in base order
          packed = ((f < 3) ? asc : __byte_perm(asc, 0, 0x0123)) & ((f < 3) ? m : __byte_perm(m, 0, 0x0123));
        } else
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = j0 + e;
          uint32_t tok = 0;
          if (j < nc) {
            // a codon is 6 contiguous bits of the packed stream starting at its lowest base:
            // forward frames read them first-base-lowest (the LUT copy with the base order reversed),
            // reverse frames first-base-highest, complemented by xor 0b101010
            const int lowest = (f < 3) ? f + 3 * j : n - 3 - (f - 3) - 3 * j;
            const int cb = 2 * lowest, cw = cb >> 5;
            const uint32_t six = __funnelshift_r(s_codes[cw], s_codes[cw + 1], cb & 31) & 63u;
            const int vw = lowest >> 5;
            const uint32_t ok3 = __funnelshift_r(s_valid[vw], s_valid[vw + 1], lowest & 31) & 7u;
            const uint32_t t = (f < 3) ? s_lut_fwd[six] : s_lut[six ^ 0x2Au];
            tok = ok3 == 7u ? t : 0u;
          }
          packed |= tok << (8 * e);
        }
        out_words[idx] = packed;                           // pitch = 4 * words_per_frame: word f * wpf + wd is word idx
      }
    }
    if (lane == 0) {
      counts[w * 4 + 0] = cg; counts[w * 4 + 1] = cc; counts[w * 4 + 2] = ca; counts[w * 4 + 3] = ct;
      const int h = round_hundredths(cg - cc, cg + cc);
      short v = static_cast<short>(h);
      if (h == 0 && cg - cc < 0 && cg + cc != 0) v = static_cast<short>(1 << 14);  // "-0.000"
      skew100[w] = v;
    }
    __syncwarp();          // the slice is rewritten by the next window
  }
}

}  // namespace jg
