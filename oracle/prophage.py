"""Oracle (test infrastructure): prophage score smoothing and region calling.

Restates
  * `logits_to_df_v2`           postprocess/prophages.py:99-153  (pinned: tests/golden/smooth.npz)
  * `segment`                   postprocess/prophages.py:524-602
  * `merge_overlapping_ranges`  postprocess/helpers.py:604-632   (pinned: tests/golden/merge_ranges.json)
  * `scale_range`               postprocess/helpers.py:656-675

The flow of `segment` around its two third-party calls is pinned on the reference's own function run with both calls
stubbed by this file's restatements (tests/golden/make_segment_goldens.py -> segment_cases.json).
PARITY UNPINNED for the two third-party calls inside `segment` (neither library is installed
here, neither is vendored by the reference):
  * ruptures >= 1.1.9 `KernelCPD(kernel="linear", min_size=3, jump=1).predict(pen=p)`: restated
    from its published algorithm -- PELT minimising  sum_segments cost + pen * n_breakpoints
    with the linear-kernel cost  cost(s,t) = sum x^2 - (sum x)^2 / (t - s)  and every segment
    >= min_size.  PELT is exact, so plain optimal partitioning returns the same optimum.
  * kneed >= 0.8.5 `KneeLocator(x, y, curve="convex", direction="decreasing")` with defaults
    S=1.0, interp_method="interp1d", online=False: restated from the Kneedle paper / kneed's
    documented steps (see `knee_locator`).
"""
from __future__ import annotations

import numpy as np


def smooth_scores(logits: np.ndarray, box: int = 4) -> np.ndarray:
    """prophages.py:126-141: row softmax in the logits' dtype, then a width-`box` box SUM."""
    v = np.exp(logits) / np.sum(np.exp(logits), axis=1).reshape(-1, 1)
    return np.stack([np.convolve(v[:, k], np.ones(box), mode="same") for k in range(v.shape[1])], axis=1)


def scale_range(x, lo, hi):
    x = np.array(x, dtype=float)
    x += -np.min(x)
    x /= np.max(x) / (hi - lo)
    x += lo
    return x


def smooth_gc_skew(gc_skew, width: int = 10):
    """prophages.py:144-151."""
    return scale_range(np.convolve(np.array(gc_skew), np.ones(width) / width, mode="same"), -1, 1)


def window_x(n: int, stride: int, length: int):
    """prophages.py:133-135 (x-axis clamp; reference test_prophage_plots.py known answer)."""
    return [min(i * stride, length) for i in range(n)]


def merge_overlapping_ranges(ranges):
    """helpers.py:604-632: the input is assumed sorted (the function's own sorted() is discarded)."""
    merged = []
    for r in ranges:
        r = [int(r[0]), int(r[1])]
        if not merged or r[0] > merged[-1][1]:
            merged.append(r)
        else:
            merged[-1][1] = max(merged[-1][1], r[1])
    return merged


def optimal_partition(x: np.ndarray, pen: float, min_size: int = 3) -> list[int]:
    """Segment end indices (ascending, last = n) minimising sum cost + pen per change point."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    s1 = np.concatenate([[0.0], np.cumsum(x)])
    s2 = np.concatenate([[0.0], np.cumsum(x * x)])
    F = np.full(n + 1, np.inf)
    prev = np.full(n + 1, -1, dtype=np.int64)
    F[0] = -pen
    for t in range(min_size, n + 1):
        s = np.arange(0, t - min_size + 1)
        s = s[(s == 0) | (s >= min_size)]
        tot = s1[t] - s1[s]
        cost = (s2[t] - s2[s]) - tot * tot / (t - s)
        v = F[s] + cost + pen
        j = int(np.argmin(v))
        F[t], prev[t] = v[j], s[j]
    out, t = [], n
    while t > 0 and prev[t] >= 0:
        out.append(t)
        t = int(prev[t])
    return out[::-1]


def knee_locator(x, y, S: float = 1.0):
    """kneed.KneeLocator(x, y, curve="convex", direction="decreasing"), defaults S=1.0,
    interp_method="interp1d" (identity on the given points), online=False.  Returns knee x or None.

    1. min-max normalise x and y;  2. transform_y for convex+decreasing: y = y.max() - y;
    3. difference curve yd = y - x;  4. local maxima / minima with argrelextrema(>=, <=), whose
    default mode="clip" lets end points qualify;  5. thresholds Tmx = yd[max] - S*|mean(diff(x))|;
    6. walk the curve from the first maximum: a maximum (re)arms the threshold, a minimum resets
    it to 0, the first point whose successor drops below the threshold yields x[threshold index]."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    if len(x) < 2 or x.max() == x.min() or y.max() == y.min():
        return None
    xn = (x - x.min()) / (x.max() - x.min())
    yn = (y - y.min()) / (y.max() - y.min())
    yn = yn.max() - yn
    yd = yn - xn
    n = len(yd)
    mx, mn = [], []
    for i in range(n):
        l, r = yd[max(i - 1, 0)], yd[min(i + 1, n - 1)]
        if yd[i] >= l and yd[i] >= r:
            mx.append(i)
        if yd[i] <= l and yd[i] <= r:
            mn.append(i)
    if not mx:
        return None
    step = abs(np.diff(xn).mean())
    thr, thr_i, used = None, None, 0
    for i in range(n):
        if i < mx[0]:
            continue
        if i == n - 1:
            break
        if i in mx:
            thr, thr_i = yd[i] - S * step, i
            used += 1
        if i in mn:
            thr = 0.0
        if thr is not None and yd[i + 1] < thr:
            return x[thr_i]
    return None


def segment(phage_col: np.ndarray, sensitivity: float = 1.5):
    """prophages.py:554-595 on one smoothed phage-score column.
    Returns (selected window-index ranges, their mean scores)."""
    col = np.asarray(phage_col, dtype=np.float64)
    preds = [optimal_partition(col, float(p)) for p in range(1, 10)]
    bkpts = [b for b in preds if len(b) > 1]
    if not bkpts:
        return [], []
    lens = np.array([len(b) for b in bkpts])
    knee = knee_locator(lens, list(range(len(bkpts))))
    try:
        idx = [len(b) for b in bkpts].index(knee) if knee else int(np.searchsorted(lens, 1))
        if idx == len(lens):
            raise TypeError("bkpts[None]")            # prophages.py:574-575 then :579 raises -> caught -> no regions
        b = bkpts[idx]
        ranges = [b[i:i + 2] for i in range(len(b) - 1)]
        scores = np.array([col[s:e + 1].mean() for s, e in ranges])   # DataFrame.loc[s:e] is end-inclusive
        mask = scores > sensitivity
        sel = merge_overlapping_ranges(np.array(ranges)[mask]) if mask.any() else []
        return sel, scores[mask]
    except Exception:
        return [], []
