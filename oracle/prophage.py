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
    """kneed.KneeLocator(x, y, curve="convex", direction="decreasing").knee, library defaults S=1.0,
    interp_method="interp1d", online=False (kneed >= 0.8.5 is not installable here; call site
    /root/reference/src/jaeger/postprocess/prophages.py:563-573).  Written with the same SciPy calls kneed makes
    -- `scipy.interpolate.interp1d(x, y)(x)` for the smoothing step and `scipy.signal.argrelextrema` for the
    extrema -- so that ties in x (equal breakpoint counts at neighbouring penalties, where interp1d is NOT the
    identity) behave as in the library.  Returns the knee's x value or None."""
    from scipy.interpolate import interp1d
    from scipy.signal import argrelextrema
    x = np.asarray(x)
    y = np.asarray(y)
    if len(x) < 2:
        return None
    with np.errstate(all="ignore"):
        ds_y = interp1d(x, y)(x)                                              # step 1: "fit a smooth line"
        x_n = (x - x.min()) / (x.max() - x.min())                             # step 2: normalise
        y_n = (ds_y - ds_y.min()) / (ds_y.max() - ds_y.min())
    if not (np.isfinite(x_n).all() and np.isfinite(y_n).all()):               # flat x or y: kneed finds nothing usable
        return None
    y_n = y_n.max() - y_n                                                     # step 3: transform_y(decreasing, convex)
    y_diff = y_n - x_n
    maxima = argrelextrema(y_diff, np.greater_equal)[0]                       # step 4 (mode="clip": end points qualify)
    minima = argrelextrema(y_diff, np.less_equal)[0]
    if not maxima.size:
        return None
    tmx = y_diff[maxima] - S * np.abs(np.diff(x_n).mean())                    # step 5
    threshold = threshold_index = None                                        # step 6: find_knee
    k = 0
    for i in range(len(y_diff)):
        if i < maxima[0]:
            continue
        if i == len(y_diff) - 1:
            break
        if (maxima == i).any():
            threshold, threshold_index = tmx[k], i
            k += 1
        if (minima == i).any():
            threshold = 0.0
        if y_diff[i + 1] < threshold:
            return float(x[threshold_index])
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
