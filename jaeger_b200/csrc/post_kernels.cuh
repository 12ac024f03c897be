// Stage 4: per-contig aggregation of window logits, score smoothing and change-point
// segmentation.  Reference: postprocess/collect.py:293-403, postprocess/helpers.py:175-235,
// postprocess/prophages.py:126-151, 554-595 (paths relative to /root/reference/src/jaeger).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace jg {

// ---- per-window scalars: argmax, entropy term, energy term --------------------------------
// entropy: helpers.py:175-177 applies -sum(p*log2 p) to the CLIPPED RAW LOGITS (not a softmax).
// energy : helpers.py:189-219 -- for n_cls != 2 the code falls into its "binary" branch, i.e.
//          -log(exp(z)+1) per element (float64), later averaged over windows AND classes.
__global__ void window_scalars_kernel(const float* __restrict__ logits, long long n_windows, int n_cls,
                                      int* __restrict__ frag_pred, float* __restrict__ entropy,
                                      double* __restrict__ energy_sum) {
  for (long long w = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; w < n_windows;
       w += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float* z = logits + w * n_cls;
    int best = 0;
    float bv = z[0];
    float ent = 0.0f;
    double en = 0.0;
    if (n_cls == 2) {
      // multiclass-softmax branch of energy(): -logsumexp over the class axis
      const double a = z[0], b = z[1];
      const double m = a > b ? a : b;
      en = -(m + log(exp(a - m) + exp(b - m)));
    }
    for (int k = 0; k < n_cls; ++k) {
      const float v = z[k];
      if (v > bv) { bv = v; best = k; }          // first maximum wins, like np.argmax
      const float c = fminf(fmaxf(v, 1e-12f), 1.0f);
      ent = __fadd_rn(ent, __fmul_rn(c, log2f(c)));
      if (n_cls != 2) {
        const double zd = v;
        const double m = zd > 0.0 ? zd : 0.0;    // logsumexp([z, 0])
        en += -(m + log(exp(zd - m) + exp(0.0 - m)));
      }
    }
    frag_pred[w] = best;
    entropy[w] = -ent;
    energy_sum[w] = en;
  }
}

// ---- per-contig reductions -----------------------------------------------------------------
// One thread per (contig, class) reproduces numpy's arithmetic exactly: np.mean / np.var over
// axis 0 of a C-contiguous float32 [T, C] array add the rows sequentially in float32, divide
// by T in float32, and the results are then cast to float16 (collect.py:332-337).
__global__ void contig_moments_kernel(const float* __restrict__ logits, const long long* __restrict__ offsets,
                                      long long n_contigs, int n_cls, __half* __restrict__ mean_h,
                                      __half* __restrict__ var_h) {
  const long long total = n_contigs * n_cls;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long c = idx / n_cls;
    const int k = static_cast<int>(idx - c * n_cls);
    const long long b = offsets[c], e = offsets[c + 1];
    const float cnt = static_cast<float>(e - b);
    float s = 0.0f;
    for (long long w = b; w < e; ++w) s = __fadd_rn(s, logits[w * n_cls + k]);
    const float mean = __fdiv_rn(s, cnt);
    float v = 0.0f;
    for (long long w = b; w < e; ++w) {
      const float d = __fsub_rn(logits[w * n_cls + k], mean);
      v = __fadd_rn(v, __fmul_rn(d, d));
    }
    mean_h[idx] = __float2half_rn(mean);
    var_h[idx] = __float2half_rn(__fdiv_rn(v, cnt));
  }
}

// One warp per contig: consensus (argmax of the fp16 means, first max), per-class counts of
// the window argmax, entropy / energy means, reliability fraction.
__global__ void contig_summary_kernel(const __half* __restrict__ mean_h, const int* __restrict__ frag_pred,
                                      const float* __restrict__ entropy, const double* __restrict__ energy_sum,
                                      const float* __restrict__ rel, const long long* __restrict__ offsets,
                                      long long n_contigs, int n_cls, int* __restrict__ consensus,
                                      int* __restrict__ counts, __half* __restrict__ entropy_h,
                                      __half* __restrict__ energy_h, int* __restrict__ rel_pos_out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long c = warp0; c < n_contigs; c += n_warps) {
    const long long b = offsets[c], e = offsets[c + 1];
    // per-class counts: lane k counts class k (n_cls <= 32)
    int my_count = 0;
    double ent = 0.0, en = 0.0;
    int rel_pos = 0;
    for (long long w = b + lane; w < e; w += 32) {
      ent += static_cast<double>(entropy[w]);
      en += energy_sum[w];
      if (rel) {
        // helpers.py:222-235 sigmoid in the logits' dtype, collect.py:233-244 "> 0.5"
        const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rel[w])));
        rel_pos += sg > 0.5f;
      }
    }
    for (long long w0 = b; w0 < e; w0 += 32) {
      const long long w = w0 + lane;
      const int fp = w < e ? frag_pred[w] : -1;
      for (int k = 0; k < n_cls; ++k) {
        const unsigned bal = __ballot_sync(0xffffffffu, fp == k);
        if (lane == k) my_count += __popc(bal);
      }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      ent += __shfl_xor_sync(0xffffffffu, ent, off);
      en += __shfl_xor_sync(0xffffffffu, en, off);
      rel_pos += __shfl_xor_sync(0xffffffffu, rel_pos, off);
    }
    if (lane < n_cls) counts[c * n_cls + lane] = my_count;
    if (lane == 0) {
      int best = 0;
      float bv = __half2float(mean_h[c * n_cls]);
      for (int k = 1; k < n_cls; ++k) {
        const float v = __half2float(mean_h[c * n_cls + k]);
        if (v > bv) { bv = v; best = k; }
      }
      consensus[c] = best;
      const double t = static_cast<double>(e - b);
      // np.mean of a float32 vector returns float32; the fp16 cast follows (collect.py:397-402)
      entropy_h[c] = __float2half_rn(static_cast<float>(ent / t));
      const double denom = (n_cls == 2) ? t : t * n_cls;
      energy_h[c] = __double2half(en / denom);
      if (rel_pos_out) rel_pos_out[c] = rel ? rel_pos : -1;
    }
  }
}

// ---- prophage score smoothing ----------------------------------------------------------------
// prophages.py:126-131: softmax over classes, then np.convolve(p, ones(box), mode="same") per
// class inside each contig.  For a kernel of length M <= N, "same" output i sums inputs
// i - (M-1)//2 - ... : out[i] = sum_{j} p[i + (M-1)/2 - j], j = 0..M-1 (zero outside).
__global__ void smooth_scores_kernel(const float* __restrict__ logits, const long long* __restrict__ offsets,
                                     long long n_contigs, int n_cls, int box, double* __restrict__ out) {
  for (long long c = blockIdx.y; c < n_contigs; c += gridDim.y) {
    const long long b = offsets[c], e = offsets[c + 1];
    const long long n = e - b;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      for (int k = 0; k < n_cls; ++k) {
        double acc = 0.0;
        // numpy swaps the operands when the kernel is longer than the signal; then the
        // output has the kernel's length -- not reachable here (contigs have >= box windows
        // whenever prophage mode applies, lc >= 500 kbp), so only the M <= N case is coded.
        const long long lo = i + (box - 1) / 2 - (box - 1), hi = i + (box - 1) / 2;
        for (long long j = lo; j <= hi; ++j) {
          if (j < 0 || j >= n) continue;
          // prophages.py:126: np.exp(value) / np.sum(np.exp(value), axis=1) in the logits' float32
          const float* z = logits + (b + j) * n_cls;
          float den = 0.0f;
          for (int q = 0; q < n_cls; ++q) den = __fadd_rn(den, expf(z[q]));
          acc += static_cast<double>(__fdiv_rn(expf(z[k]), den));
        }
        out[(b + i) * n_cls + k] = acc;
      }
    }
  }
}

// ---- change-point segmentation ---------------------------------------------------------------
// Optimal partitioning (the objective PELT minimises exactly) with the L2 cost
//   cost(s, t) = sum x^2 - (sum x)^2 / (t - s)   over x[s:t],
// every segment >= min_size, total = sum cost + pen * (#segments - 1)... ruptures adds `pen`
// per change point.  One CTA per penalty; the inner minimisation over the last change point
// is a block-wide min-reduction (warp shuffles), the outer loop over t is sequential.
// Batched over contigs (blockIdx.y): contig c owns signal points [offsets[c], offsets[c+1]) of the concatenated signal and
// the work / output slots at the same positions (+ c for the n + 1 sized arrays).  offsets == nullptr: one signal of n points.
struct SegSlot { long long base, wbase; int n; };
__device__ __forceinline__ SegSlot seg_slot(const long long* offsets, int n_single, int c) {
  SegSlot s;
  if (offsets) { s.base = offsets[c]; s.n = static_cast<int>(offsets[c + 1] - offsets[c]); s.wbase = s.base + c; }
  else { s.base = 0; s.n = n_single; s.wbase = 0; }
  return s;
}
// work layout for N = total points, C contigs, P penalties, T = N + C:  S1 [T] | S2 [T] | F [P][T] doubles | prev [P][T] ints
__global__ void prefix_sums_kernel(const double* __restrict__ x, const long long* __restrict__ offsets, int n_single,
                                   long long total_slots, double* __restrict__ work) {
  if (threadIdx.x == 0) {
    const SegSlot sl = seg_slot(offsets, n_single, blockIdx.x);
    double* S1 = work + sl.wbase;
    double* S2 = work + total_slots + sl.wbase;
    const double* xs = x + sl.base;
    double a = 0.0, b = 0.0;
    S1[0] = 0.0; S2[0] = 0.0;
    for (int i = 0; i < sl.n; ++i) { a += xs[i]; b += xs[i] * xs[i]; S1[i + 1] = a; S2[i + 1] = b; }
  }
}

__global__ void segment_scores_kernel(const long long* __restrict__ offsets, int n_single, long long total_points,
                                      long long total_slots, int min_size, int n_pen,
                                      double* __restrict__ work, int* __restrict__ bkps, int* __restrict__ nbkps) {
  extern __shared__ double s_red[];
  int* s_idx = reinterpret_cast<int*>(s_red + 32);
  const int pen_i = blockIdx.x;
  if (pen_i >= n_pen) return;
  const SegSlot sl = seg_slot(offsets, n_single, blockIdx.y);
  const int n = sl.n;
  const double pen = static_cast<double>(pen_i + 1);
  const double* S1 = work + sl.wbase;                 // prefix sums from prefix_sums_kernel
  const double* S2 = work + total_slots + sl.wbase;
  double* F = work + 2 * total_slots + static_cast<long long>(pen_i) * total_slots + sl.wbase;
  int* prev = reinterpret_cast<int*>(work + 2 * total_slots + static_cast<long long>(n_pen) * total_slots) +
              static_cast<long long>(pen_i) * total_slots + sl.wbase;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int t = 0; t <= n; ++t) {
    if (threadIdx.x == 0) { if (t == 0) { F[0] = -pen; prev[0] = 0; } }
    __syncthreads();
    if (t == 0) continue;
    double best = 1e300;
    int best_s = -1;
    if (t >= min_size) {
      for (int s = threadIdx.x; s <= t - min_size; s += blockDim.x) {
        if (s != 0 && s < min_size) continue;      // the first segment must also be >= min_size
        const double f = F[s];
        if (f >= 1e299) continue;
        const double sum = S1[t] - S1[s];
        const double c = (S2[t] - S2[s]) - sum * sum / static_cast<double>(t - s);
        const double v = f + c + pen;
        if (v < best || (v == best && s < best_s)) { best = v; best_s = s; }
      }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, off);
      const int os = __shfl_xor_sync(0xffffffffu, best_s, off);
      if (ov < best || (ov == best && os >= 0 && (best_s < 0 || os < best_s))) { best = ov; best_s = os; }
    }
    if (lane == 0) { s_red[warp] = best; s_idx[warp] = best_s; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int wv = 1; wv < nwarps; ++wv)
        if (s_red[wv] < best || (s_red[wv] == best && s_idx[wv] >= 0 && (best_s < 0 || s_idx[wv] < best_s))) {
          best = s_red[wv]; best_s = s_idx[wv];
        }
      F[t] = best_s >= 0 ? best : 1e300;
      prev[t] = best_s;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // backtrack: segment ends in ascending order, last = n.  Contig c's list for penalty p starts at bkps[p][offsets[c]].
    int cnt = 0;
    int t = n;
    int* out = bkps + static_cast<long long>(pen_i) * total_points + sl.base;
    while (t > 0 && prev[t] >= 0) { out[cnt++] = t; t = prev[t]; }
    for (int i = 0; i < cnt / 2; ++i) { const int tmp = out[i]; out[i] = out[cnt - 1 - i]; out[cnt - 1 - i] = tmp; }
    nbkps[static_cast<long long>(blockIdx.y) * n_pen + pen_i] = cnt;
  }
}


// ---- linear-chain CRF (Viterbi) decoding of each contig's window labels ---------------------------
// postprocess/helpers.py:393-449: emissions = log-softmax of the logits in float64, switching from
// class a to b costs costs[a][b]; delta/backpointer recursion, first index wins every argmax (as
// np.argmax).  One thread walks one contig (the recursion is sequential in t; contigs are the
// parallel axis); back-pointers go through a [W][n_cls] byte scratch.  Also returns the per-contig
// class counts of the decoded path (collect.py:349-355).
constexpr int kMaxCrfClasses = 8;
__global__ void viterbi_kernel(const float* __restrict__ logits, const long long* __restrict__ offsets, int n_contigs,
                               int n_cls, const double* __restrict__ costs, uint8_t* __restrict__ backptr,
                               int* __restrict__ path, int* __restrict__ counts) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_contigs; c += gridDim.x * blockDim.x) {
    const long long w0 = offsets[c], w1 = offsets[c + 1];
    int cnt[kMaxCrfClasses];
    for (int k = 0; k < kMaxCrfClasses; ++k) cnt[k] = 0;
    if (w1 > w0) {
      double cost[kMaxCrfClasses][kMaxCrfClasses], delta[kMaxCrfClasses], em[kMaxCrfClasses];
      for (int a = 0; a < n_cls; ++a)
        for (int b = 0; b < n_cls; ++b) cost[a][b] = costs[a * n_cls + b];
      auto emissions = [&](long long w) {
        double mx = -1.0e300;
        for (int k = 0; k < n_cls; ++k) { em[k] = static_cast<double>(logits[w * n_cls + k]); mx = em[k] > mx ? em[k] : mx; }
        double s = 0.0;
        for (int k = 0; k < n_cls; ++k) s += exp(em[k] - mx);
        const double lse = mx + log(s);
        for (int k = 0; k < n_cls; ++k) em[k] -= lse;
      };
      emissions(w0);
      for (int k = 0; k < n_cls; ++k) delta[k] = em[k];
      for (long long w = w0 + 1; w < w1; ++w) {
        emissions(w);
        double nd[kMaxCrfClasses];
        for (int cur = 0; cur < n_cls; ++cur) {
          int best = 0;
          double bs = delta[0] - cost[0][cur];
          for (int prev = 1; prev < n_cls; ++prev) {
            const double sc = delta[prev] - cost[prev][cur];
            if (sc > bs) { bs = sc; best = prev; }
          }
          backptr[w * n_cls + cur] = static_cast<uint8_t>(best);
          nd[cur] = em[cur] + bs;
        }
        for (int k = 0; k < n_cls; ++k) delta[k] = nd[k];
      }
      int cur = 0;
      for (int k = 1; k < n_cls; ++k) if (delta[k] > delta[cur]) cur = k;
      for (long long w = w1 - 1; w >= w0; --w) {
        path[w] = cur;
        ++cnt[cur];
        if (w > w0) cur = backptr[w * n_cls + cur];
      }
    }
    for (int k = 0; k < n_cls; ++k) counts[c * n_cls + k] = cnt[k];
  }
}


// ---- legacy `default` model: per-window reliability from the embedding ---------------------------
// postprocess/helpers.py:558-565 (ood_predict_default, "sklearn" variant) with the bundled model
// written out: features = l2_normalise((embedding - batch_mean) / batch_std) in float32, decision =
// features . coef + intercept in float64 (LogisticRegression), probability of class 0 after the
// prefit sigmoid calibration = 1 - 1 / (1 + exp(a * decision + b)).  One warp per window; a second
// pass takes the per-contig mean (generate_summary_legacy's reliability_score, collect.py:121-123).
__global__ void legacy_reliability_kernel(const float* __restrict__ emb, long long n_windows, int dim,
                                          const float* __restrict__ mean, const float* __restrict__ sd,
                                          const double* __restrict__ coef, double intercept, double cal_a, double cal_b,
                                          double* __restrict__ p0) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long w = warp; w < n_windows; w += n_warps) {
    float ss = 0.0f;
    for (int k = lane; k < dim; k += 32) {
      const float f = __fdiv_rn(__fsub_rn(emb[w * dim + k], mean[k]), sd[k]);
      ss = __fmaf_rn(f, f, ss);
    }
    for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = __fsqrt_rn(ss);
    double dec = 0.0;
    for (int k = lane; k < dim; k += 32) {
      const float f = __fdiv_rn(__fdiv_rn(__fsub_rn(emb[w * dim + k], mean[k]), sd[k]), nrm);
      dec += static_cast<double>(f) * coef[k];
    }
    for (int o = 16; o >= 1; o >>= 1) dec += __shfl_xor_sync(0xffffffffu, dec, o);
    if (lane == 0) p0[w] = 1.0 - 1.0 / (1.0 + exp(cal_a * (dec + intercept) + cal_b));
  }
}

__global__ void segment_mean_f64_kernel(const double* __restrict__ v, const long long* __restrict__ offsets, int n_seg,
                                        double* __restrict__ out) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_seg; c += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (long long w = offsets[c]; w < offsets[c + 1]; ++w) s += v[w];
    const long long n = offsets[c + 1] - offsets[c];
    out[c] = n > 0 ? s / static_cast<double>(n) : 0.0;
  }
}

// ---- post-hoc refinement (`--refine`): postprocess/refinement.py:39-247 ---------------------------------------
// Score columns are positional (commands/predict.py:140): 0 phage, 1 virus, 2 archaea, 3 bacteria, 4 plasmid,
// 5 eukarya.  Window labels: 0..5 = class, 6 = unknown, 7 = bacteria_or_plasmid, 8 = virus_any.
constexpr int kRefineClasses = 6;
constexpr int kRefineUnknown = 6, kRefineBactPlasmid = 7, kRefineVirusAny = 8;

// positions -1 / -2 of a stable ascending argsort of 6 values (np.argsort on a row; ties keep index order)
__device__ __forceinline__ void refine_top2(const double (&z)[kRefineClasses], int& last, int& second) {
  double v[kRefineClasses];
  int id[kRefineClasses];
#pragma unroll
  for (int k = 0; k < kRefineClasses; ++k) { v[k] = z[k]; id[k] = k; }
#pragma unroll
  for (int a = 1; a < kRefineClasses; ++a) {          // insertion sort with a strict comparison: stable
#pragma unroll
    for (int b = a; b > 0; --b) {
      if (v[b] < v[b - 1]) {
        const double tv = v[b]; v[b] = v[b - 1]; v[b - 1] = tv;
        const int ti = id[b]; id[b] = id[b - 1]; id[b - 1] = ti;
      }
    }
  }
  last = id[kRefineClasses - 1];
  second = id[kRefineClasses - 2];
}

// add_score_features + refine (refinement.py:39-73, 97-137): one thread per window, float64 like the reference's
// NumPy arithmetic on the widened logits.  tau[0..5] = per-class logit thresholds, tau[6..11] = margin thresholds.
__global__ void refine_windows_kernel(const float* __restrict__ logits, long long n_windows, int n_cls,
                                      const double* __restrict__ tau, int merge_bp, int merge_pv,
                                      uint8_t* __restrict__ label, double* __restrict__ margin_out) {
  for (long long w = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; w < n_windows;
       w += static_cast<long long>(gridDim.x) * blockDim.x) {
    double z[kRefineClasses];
#pragma unroll
    for (int k = 0; k < kRefineClasses; ++k) z[k] = static_cast<double>(logits[w * n_cls + k]);
    int top = 0;                                        // np.argmax: first maximum
#pragma unroll
    for (int k = 1; k < kRefineClasses; ++k) top = z[k] > z[top] ? k : top;
    int last, second;
    refine_top2(z, last, second);
    const double top_logit = z[last], margin = z[last] - z[second];
    const double tau_logit = tau[top], tau_margin = tau[kRefineClasses + top];
    int lab = top;
    const bool low_margin = margin < tau_margin;
    if (merge_bp && low_margin && ((top == 3 && second == 4) || (top == 4 && second == 3))) lab = kRefineBactPlasmid;
    if (merge_pv && low_margin && ((top == 0 && second == 1) || (top == 1 && second == 0))) lab = kRefineVirusAny;
    if (lab < kRefineClasses && (top_logit < tau_logit || low_margin)) lab = kRefineUnknown;
    label[w] = static_cast<uint8_t>(lab);
    margin_out[w] = margin;
  }
}

// aggregate_contig (refinement.py:140-213): one warp per contig, segmented float64 sums of score * weight *
// class multiplier.  mode 0 gated (drop unknown, weight 1), 1 weighted (drop unknown, weight = max(margin, 0)),
// 2 unweighted (all windows, weight 1).  sums [n][6], stats [n][2] = {windows used, merged-label windows},
// total_weight [n].  Lanes add their strided partial sums, then a fixed butterfly: deterministic.
__global__ void refine_contigs_kernel(const float* __restrict__ logits, int n_cls, const uint8_t* __restrict__ label,
                                      const double* __restrict__ margin, const long long* __restrict__ offsets,
                                      long long n_contigs, int mode, double merge_share, double* __restrict__ sums,
                                      int* __restrict__ stats, double* __restrict__ total_weight) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long c = warp0; c < n_contigs; c += n_warps) {
    double acc[kRefineClasses] = {0, 0, 0, 0, 0, 0};
    double wsum = 0.0;
    int used = 0, merged = 0;
    for (long long w = offsets[c] + lane; w < offsets[c + 1]; w += 32) {
      const int lab = label[w];
      if (mode != 2 && lab == kRefineUnknown) continue;
      const double m = margin[w];
      const double wt = mode == 1 ? (m > 0.0 ? m : 0.0) : 1.0;
      const bool is_merged = lab == kRefineBactPlasmid || lab == kRefineVirusAny;
#pragma unroll
      for (int k = 0; k < kRefineClasses; ++k) {
        const bool member = lab == kRefineBactPlasmid ? (k == 3 || k == 4) : (k == 0 || k == 1);
        const double mult = is_merged ? (member ? merge_share : 0.0) : 1.0;
        acc[k] += static_cast<double>(logits[w * n_cls + k]) * wt * mult;
      }
      wsum += wt; ++used; merged += is_merged;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
      for (int k = 0; k < kRefineClasses; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
      wsum += __shfl_xor_sync(0xffffffffu, wsum, off);
      used += __shfl_xor_sync(0xffffffffu, used, off);
      merged += __shfl_xor_sync(0xffffffffu, merged, off);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < kRefineClasses; ++k) sums[c * kRefineClasses + k] = acc[k];
      stats[c * 2] = used; stats[c * 2 + 1] = merged;
      total_weight[c] = wsum;
    }
  }
}

}  // namespace jg
