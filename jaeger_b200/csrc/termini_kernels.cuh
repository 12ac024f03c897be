// Terminal-repeat scan: the two Smith-Waterman alignments the reference runs per
// contig with parasail.sw_trace_scan_16 (utils/termini.py:103-131) -- the first `n` bases of the
// contig against its last `n` bases (direct repeat) and against their reverse complement (inverted
// repeat); substitution matrix parasail.matrix_create("ACGT", 2, -100) (any other letter scores 0
// against everything, case-insensitive), gap open 100 (cost of a gap of length 1) / extend 5.
//
// Layout: one CTA per alignment, thread t owns the kRows query rows [t*kRows, (t+1)*kRows) and
// sweeps the reference columns as a skewed wavefront (thread t is at column step - t).  The strip's
// H / E / run-length state lives in registers; only the bottom cell of a strip crosses to the next
// thread, through a double-buffered shared-memory slot, with one barrier per step.
//
// Pass 1 (sw_scan_kernel) returns score, end position (first maximum in column-major order, as a
// column-wise scan finds it) and the length of the diagonal run that ends there.  Gaps and
// mismatches cost 100, so a path with one needs >= 51 matches on both sides of it: every alignment
// scoring below 104 is a pure diagonal run and pass 1 already knows its length.  Pass 2
// (sw_trace_kernel) refills the rectangle up to the end cell for the few alignments scoring >= 104,
// writing one direction byte per cell, and walks the traceback (H: zero > diagonal > E > F; E and F
// prefer to open on ties) to count columns, gaps and identities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jg {

constexpr int kSwRows = 16;          // query rows per thread
constexpr int kSwNeg = -30000;

struct SwScores { int match, mismatch, wild, open, ext; };

struct SwSeq {
  const uint32_t* codes;   // 2-bit packed bases (A0 C1 T2 G3), 16 per word
  const uint32_t* valid;   // 1 bit per base: is A/C/G/T
};

// code 0..3, or 4 for a letter outside ACGT
__device__ __forceinline__ int sw_base(const SwSeq& s, long long b) {
  const uint32_t ok = (s.valid[b >> 5] >> (b & 31)) & 1u;
  return ok ? static_cast<int>((s.codes[b >> 4] >> (2 * (b & 15))) & 3u) : 4;
}
__device__ __forceinline__ int sw_sub(int a, int b, const SwScores& sc) {
  return (a > 3 || b > 3) ? sc.wild : (a == b ? sc.match : sc.mismatch);
}

// job = (query start, reference start, n, inverted?, nq): query[i] = seq[q0 + i], i < nq (nq = 0 means n: the
// square terminal-repeat jobs); ref[j] = seq[r0 + j] (direct) or complement(seq[r0 + n - 1 - j]) (inverted), j < n.
// Rectangular jobs are the att-site scans around a prophage region (postprocess/prophages.py:771-800).
struct SwJob { long long q0, r0; int n, inverted, nq, reserved; };
__device__ __forceinline__ int sw_rows(const SwJob& jb) { return jb.nq > 0 ? jb.nq : jb.n; }

__device__ __forceinline__ int sw_ref_base(const SwSeq& s, const SwJob& jb, int j) {
  if (!jb.inverted) return sw_base(s, jb.r0 + j);
  const int c = sw_base(s, jb.r0 + jb.n - 1 - j);
  return c > 3 ? c : (c ^ 2);
}

// kTrace = false: out[job] = {score, end_query, end_ref, diagonal run length at the end cell}
// kTrace = true : rows/cols limited to the rectangle [0, lim_i] x [0, lim_j]; writes dirs.
template <bool kTrace>
__device__ __forceinline__ void sw_fill(const SwSeq& seq, const SwJob& jb, const SwScores& sc, int n_rows, int n_cols,
                                        uint8_t* s_ref, int* s_slot, uint8_t* dirs, int (&best)[4]) {
  const int t = threadIdx.x, T = blockDim.x;
  for (int j = t; j < n_cols; j += T) s_ref[j] = static_cast<uint8_t>(sw_ref_base(seq, jb, j));
  const int i0 = t * kSwRows;
  int q[kSwRows], Hl[kSwRows], E[kSwRows], Ll[kSwRows];
#pragma unroll
  for (int k = 0; k < kSwRows; ++k) {
    q[k] = (i0 + k < n_rows) ? sw_base(seq, jb.q0 + i0 + k) : 4;
    Hl[k] = 0; E[k] = kSwNeg; Ll[k] = 0;
  }
  int* slot_h = s_slot;                 // [2][T]
  int* slot_f = s_slot + 2 * T;
  int* slot_l = s_slot + 4 * T;
  int up_h_prev = 0, up_l_prev = 0;     // H / L of (i0-1, j-1)
  int bs = 0, bi = 0, bj = 0, bl = 0;
  __syncthreads();
  const int steps = n_cols + T - 1;
  for (int s = 0; s < steps; ++s) {
    const int j = s - t;
    if (j >= 0 && j < n_cols && i0 < n_rows) {
      int hup = 0, fup = kSwNeg, lup = 0;
      if (t > 0) {
        const int rb = ((s - 1) & 1) * T + (t - 1);
        hup = slot_h[rb]; fup = slot_f[rb]; lup = slot_l[rb];
      }
      int hdiag = up_h_prev, ldiag = up_l_prev;
      up_h_prev = hup; up_l_prev = lup;
      const int rj = s_ref[j];
#pragma unroll
      for (int k = 0; k < kSwRows; ++k) {
        if (i0 + k < n_rows) {
          const int e_ext = E[k] - sc.ext, e_opn = Hl[k] - sc.open;
          const int f_ext = fup - sc.ext, f_opn = hup - sc.open;
          const int e = e_ext > e_opn ? e_ext : e_opn;
          const int f = f_ext > f_opn ? f_ext : f_opn;
          const int hd = hdiag + sw_sub(q[k], rj, sc);
          int h = hd > 0 ? hd : 0;
          h = e > h ? e : h;
          h = f > h ? f : h;
          const int l = (h > 0 && h == hd) ? ldiag + 1 : 0;
          if (kTrace) {
            const uint32_t src = h == 0 ? 0u : (h == hd ? 1u : (h == e ? 2u : 3u));
            dirs[static_cast<long long>(i0 + k) * n_cols + j] =
                static_cast<uint8_t>(src | (e_ext > e_opn ? 4u : 0u) | (f_ext > f_opn ? 8u : 0u));
          } else if (h > bs || (h == bs && h > 0 && (j < bj || (j == bj && i0 + k < bi)))) {
            bs = h; bi = i0 + k; bj = j; bl = l;
          }
          hdiag = Hl[k]; ldiag = Ll[k];
          Hl[k] = h; E[k] = e; Ll[k] = l;
          hup = h; fup = f;
        }
      }
      const int wb = (s & 1) * T + t;
      slot_h[wb] = hup; slot_f[wb] = fup; slot_l[wb] = Ll[kSwRows - 1];
    }
    __syncthreads();
  }
  best[0] = bs; best[1] = bi; best[2] = bj; best[3] = bl;
}

// jobs sorted so that every job of a launch fits blockDim.x * kSwRows rows
__global__ void sw_scan_kernel(SwSeq seq, const SwJob* __restrict__ jobs, SwScores sc, int* __restrict__ out) {
  extern __shared__ int s_sw[];
  const SwJob jb = jobs[blockIdx.x];
  const int T = blockDim.x;
  int* s_slot = s_sw;                                        // 6 * T ints
  int* s_red = s_sw + 6 * T;                                 // 4 * T ints
  uint8_t* s_ref = reinterpret_cast<uint8_t*>(s_sw + 10 * T);
  int best[4];
  sw_fill<false>(seq, jb, sc, sw_rows(jb), jb.n, s_ref, s_slot, nullptr, best);
  const int t = threadIdx.x;
  for (int k = 0; k < 4; ++k) s_red[k * T + t] = best[k];
  __syncthreads();
  if (t == 0) {
    int bs = 0, bi = 0, bj = 0, bl = 0;
    for (int u = 0; u < T; ++u) {
      const int h = s_red[u], i = s_red[T + u], j = s_red[2 * T + u];
      if (h > bs || (h == bs && h > 0 && (j < bj || (j == bj && i < bi)))) { bs = h; bi = i; bj = j; bl = s_red[3 * T + u]; }
    }
    int* o = out + 4 * static_cast<long long>(blockIdx.x);
    o[0] = bs; o[1] = bi; o[2] = bj; o[3] = bl;
  }
}

// ops_off >= 0: the traceback also writes one byte per alignment column at scratch + ops_off, from the LAST column
// backwards (1 = pair, 2 = gap in the query line, 3 = gap in the reference line); at most end_i + end_j + 2 bytes.
struct SwTraceJob { SwJob job; int end_i, end_j; long long dirs_off, ops_off; };

// out[job] = {alignment columns, gaps in the query line, gaps in the reference line, identities}
__global__ void sw_trace_kernel(SwSeq seq, const SwTraceJob* __restrict__ jobs, SwScores sc, uint8_t* __restrict__ scratch,
                                int* __restrict__ out) {
  extern __shared__ int s_sw[];
  const SwTraceJob tj = jobs[blockIdx.x];
  const int T = blockDim.x;
  int* s_slot = s_sw;
  uint8_t* s_ref = reinterpret_cast<uint8_t*>(s_sw + 6 * T);
  uint8_t* dirs = scratch + tj.dirs_off;
  const int n_rows = tj.end_i + 1, n_cols = tj.end_j + 1;
  int best[4];
  sw_fill<true>(seq, tj.job, sc, n_rows, n_cols, s_ref, s_slot, dirs, best);
  __threadfence_block();
  __syncthreads();
  if (threadIdx.x == 0) {
    int i = tj.end_i, j = tj.end_j, cols = 0, qgaps = 0, rgaps = 0, iden = 0, state = 0;   // 0 H, 1 E, 2 F
    uint8_t* ops = tj.ops_off >= 0 ? scratch + tj.ops_off : nullptr;
    while (i >= 0 && j >= 0) {
      const uint32_t d = dirs[static_cast<long long>(i) * n_cols + j];
      if (state == 0) {
        const uint32_t src = d & 3u;
        if (src == 0u) break;
        if (src == 1u) {
          const int a = sw_base(seq, tj.job.q0 + i), b = s_ref[j];
          iden += (a < 4 && a == b) ? 1 : 0;
          if (ops) ops[cols] = 1;
          ++cols; --i; --j;
        } else {
          state = src == 2u ? 1 : 2;
        }
      } else if (state == 1) {        // reference base against a gap in the query
        if (ops) ops[cols] = 2;
        ++cols; ++qgaps;
        state = (d & 4u) ? 1 : 0;
        --j;
      } else {                        // query base against a gap in the reference
        if (ops) ops[cols] = 3;
        ++cols; ++rgaps;
        state = (d & 8u) ? 2 : 0;
        --i;
      }
    }
    int* o = out + 4 * static_cast<long long>(blockIdx.x);
    o[0] = cols; o[1] = qgaps; o[2] = rgaps; o[3] = iden;
  }
}

}  // namespace jg
