// Symmetric DUST (SDUST, Morgulis et al. 2006) low-complexity soft-masking on the packed bases.
// Replaces pydustmasker.DustMasker(seq, window_size=64, score_threshold=20).mask()
// (reference call site: seqops/io.py:105-108).  The algorithm is a sequential scan whose state is
// confined to the last W bases, so contigs are cut into chunks that are scanned independently
// (one thread per chunk) with a 2W warm-up before and a W run-out after the chunk; a chunk only
// sets mask bits inside its own core, which makes the union identical to a whole-contig scan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jg {

constexpr int kDustW = 64;        // window (bases)
constexpr int kDustWords = 62;    // words per full window: W - 3 + 1
constexpr int kDustMaxP = 2048;   // perfect intervals alive in one window (<= 62*61/2 = 1891)

struct DustState {
  uint8_t w[64];                  // ring buffer of the window's words
  uint8_t cv[64], cw[64];
  int head, size;                 // ring buffer
  int rv, rw, L;
  int n_p;
  // one perfect interval per word: start (relative to the chunk scan start, 12 bits) |
  // finish - start (7 bits) << 12 | score r (12 bits) << 19;  its length l = finish - start - 3
  uint32_t p[kDustMaxP];
};
__device__ __forceinline__ int dp_start(uint32_t e) { return static_cast<int>(e & 0xFFFu); }
__device__ __forceinline__ int dp_len(uint32_t e) { return static_cast<int>((e >> 12) & 0x7Fu); }
__device__ __forceinline__ int dp_r(uint32_t e) { return static_cast<int>(e >> 19); }

__device__ __forceinline__ void dust_set_bits(uint32_t* soft, long long a, long long b) {
  for (long long i = a; i < b;) {
    const long long word = i >> 5;
    const int lo = static_cast<int>(i & 31);
    const long long end = (word + 1) << 5;
    const int hi = static_cast<int>((b < end ? b : end) - (word << 5));   // exclusive bit index in the word
    const uint32_t m = (hi == 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
    atomicOr(soft + word, m);
    i = end;
  }
}

// one thread per chunk; all coordinates are absolute base offsets into the packed arrays
__global__ void dust_kernel(const uint32_t* __restrict__ codes, const uint32_t* __restrict__ valid,
                            const long long* __restrict__ core_begin, const long long* __restrict__ core_end,
                            const long long* __restrict__ contig_begin, const long long* __restrict__ contig_end,
                            long long n_chunks, int T, uint32_t* __restrict__ soft) {
  const long long ch = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (ch >= n_chunks) return;
  const long long cb = core_begin[ch], ce = core_end[ch];
  long long s0 = cb - 2 * kDustW;
  if (s0 < contig_begin[ch]) s0 = contig_begin[ch];
  long long e1 = ce + kDustW;
  if (e1 > contig_end[ch]) e1 = contig_end[ch];
  DustState st;
  for (int k = 0; k < 64; ++k) st.cv[k] = st.cw[k] = 0;
  st.head = st.size = st.rv = st.rw = st.L = st.n_p = 0;
  int l = 0;
  unsigned t = 0;

  auto at = [&](int i) -> int { return st.w[(st.head + i) & 63]; };
  auto save_masked = [&](long long start) {
    if (st.n_p == 0 || dp_start(st.p[st.n_p - 1]) + s0 >= start) return;
    const long long a = dp_start(st.p[st.n_p - 1]) + s0, b = a + dp_len(st.p[st.n_p - 1]);
    const long long lo = a > cb ? a : cb, hi = b < ce ? b : ce;
    if (lo < hi) dust_set_bits(soft, lo, hi);
    int i = st.n_p - 1;
    while (i >= 0 && dp_start(st.p[i]) + s0 < start) --i;
    st.n_p = i + 1;
  };

  for (long long i = s0; i <= e1; ++i) {
    int b = 4;
    if (i < e1 && ((valid[i >> 5] >> (i & 31)) & 1u)) b = static_cast<int>((codes[i >> 4] >> (2 * (i & 15))) & 3u);
    if (b < 4) {
      ++l;
      t = ((t << 2) | static_cast<unsigned>(b)) & 63u;
      if (l >= 3) {
        const long long start = (l - kDustW > 0 ? l - kDustW : 0) + (i + 1 - l);
        save_masked(start);
        // ---- shift_window ----
        if (st.size >= kDustWords) {
          const int s = st.w[st.head];
          st.head = (st.head + 1) & 63;
          --st.size;
          st.rw -= --st.cw[s];
          if (st.L > st.size) { --st.L; st.rv -= --st.cv[s]; }
        }
        st.w[(st.head + st.size) & 63] = static_cast<uint8_t>(t);
        ++st.size;
        ++st.L;
        st.rw += st.cw[t]++;
        st.rv += st.cv[t]++;
        if (st.cv[t] * 10 > T * 2) {
          int s;
          do {
            s = at(st.size - st.L);
            st.rv -= --st.cv[s];
            --st.L;
          } while (s != static_cast<int>(t));
        }
        // ---- find_perfect ----
        if (st.rw * 10 > st.L * T) {
          uint8_t c[64];
          for (int k = 0; k < 64; ++k) c[k] = st.cv[k];
          int r = st.rv, max_r = 0, max_l = 0;
          const int rel_start = static_cast<int>(start - s0);
          for (int q = st.size - st.L - 1; q >= 0; --q) {
            const int tw = at(q);
            r += c[tw]++;
            const int new_r = r, new_l = st.size - q - 1;
            if (new_r * 10 > T * new_l) {
              int j = 0;
              while (j < st.n_p && dp_start(st.p[j]) >= q + rel_start) {
                const int pr = dp_r(st.p[j]), pl = dp_len(st.p[j]) - 3;
                if (max_r == 0 || pr * max_l > max_r * pl) { max_r = pr; max_l = pl; }
                ++j;
              }
              if (max_r == 0 || new_r * max_l >= max_r * new_l) {
                max_r = new_r; max_l = new_l;
                if (st.n_p < kDustMaxP) {
                  for (int m = st.n_p; m > j; --m) st.p[m] = st.p[m - 1];
                  ++st.n_p;
                  st.p[j] = static_cast<uint32_t>(q + rel_start) | (static_cast<uint32_t>(st.size + 2 - q) << 12) |
                            (static_cast<uint32_t>(new_r) << 19);
                }
              }
            }
          }
        }
      }
    } else {
      long long start = (l - kDustW + 1 > 0 ? l - kDustW + 1 : 0) + (i + 1 - l);
      while (st.n_p) save_masked(start++);
      l = 0; t = 0;
      for (int k = 0; k < 64; ++k) st.cv[k] = st.cw[k] = 0;
      st.head = st.size = st.rv = st.rw = st.L = 0;
    }
  }
}

}  // namespace jg
