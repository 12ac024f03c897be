"""Prophage region calling on top of the stage-4b device kernels.

Counterpart of `logits_to_df_v2` + `segment` (postprocess/prophages.py:99-153, 524-602):
softmax + width-4 box sums (`jg_smooth_scores`) and the penalised change-point search for
pen = 1..9 (`jg_segment_scores`) run on the GPU; the knee selection over <= 9 points, the range
filter and the interval merge are scalar host logic.
"""
from __future__ import annotations

import numpy as np
import torch

from ._cabi import check, lib


def _interp_at_nodes(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """kneed's smoothing step `scipy.interpolate.interp1d(x, y)(x)` written out: interp1d sorts the nodes by x
    (stable merge sort) and, for 1-D linear interpolation, evaluates with numpy.interp semantics -- a query that
    equals a node value takes the LAST node of that value in sorted order.  x here are breakpoint counts at
    pen = 1..9, non-increasing and often tied, so this is not the identity: every member of a tie group gets the
    y of the group's last member (postprocess/prophages.py:563-568 passes x = counts, y = 0..n-1)."""
    order = np.argsort(x, kind="mergesort")
    xs, ys = x[order], y[order]
    last = np.searchsorted(xs, x, side="right") - 1          # last sorted node with value <= x[i]: x[i] itself is a node
    return ys[last]


def knee_point(x, y, S: float = 1.0):
    """kneed.KneeLocator(x, y, curve="convex", direction="decreasing").knee with the library defaults
    (S=1, interp_method="interp1d", online=False): the x value at the first knee, or None.
    Steps (kneed 0.8.x `KneeLocator.__init__` / `find_knee`; call site postprocess/prophages.py:563-573):
    Ds_y = interp1d(x, y)(x); min-max normalise x and Ds_y; y <- max(y) - y; difference curve; local maxima /
    minima with >= / <= against both neighbours (argrelextrema, mode="clip"); threshold of a maximum =
    its height - S * |mean(diff(x_norm))|; walk from the first maximum, a minimum resets the threshold to 0,
    the first point whose successor falls below the threshold yields x[index of the last maximum]."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    if len(x) < 2 or np.ptp(x) == 0:
        return None
    ds = _interp_at_nodes(x, y)
    if np.ptp(ds) == 0:
        return None
    xn = (x - x.min()) / (x.max() - x.min())
    yn = (ds - ds.min()) / (ds.max() - ds.min())
    yn = yn.max() - yn                                   # convex + decreasing -> knee form
    yd = yn - xn
    n = len(yd)
    left = yd[np.clip(np.arange(n) - 1, 0, n - 1)]
    right = yd[np.clip(np.arange(n) + 1, 0, n - 1)]
    is_max = (yd >= left) & (yd >= right)
    is_min = (yd <= left) & (yd <= right)
    if not is_max.any():
        return None
    step = abs(np.diff(xn).mean())
    threshold, threshold_index = None, None
    for i in range(int(np.argmax(is_max)), n - 1):
        if is_max[i]:
            threshold, threshold_index = yd[i] - S * step, i
        if is_min[i]:
            threshold = 0.0
        if yd[i + 1] < threshold:
            return x[threshold_index]
    return None


def merge_overlapping_ranges(ranges):
    """helpers.py:604-632 (input order is kept: the reference discards its own sort)."""
    merged: list[list[int]] = []
    for r in ranges:
        r = [int(r[0]), int(r[1])]
        if not merged or r[0] > merged[-1][1]:
            merged.append(r)
        else:
            merged[-1][1] = max(merged[-1][1], r[1])
    return merged


def smooth(engine, logits_dev: torch.Tensor, box: int = 4) -> torch.Tensor:
    """[T, n_cls] fp32 logits of ONE contig -> [T, n_cls] float64 smoothed class scores."""
    t, n_cls = logits_dev.shape
    off = engine._h2d(np.array([0, t], dtype=np.int64))
    out = torch.empty((t, n_cls), dtype=torch.float64, device=engine.tdev)
    check(lib.jg_smooth_scores(engine.ctx.handle, logits_dev.data_ptr(), off.data_ptr(), 1, n_cls, box, out.data_ptr()))
    return out


def breakpoints(engine, signal_dev: torch.Tensor, min_size: int = 3, n_pen: int = 9) -> list[list[int]]:
    """Segment ends for pen = 1..n_pen (what KernelCPD("linear", min_size, jump=1).predict(pen) returns)."""
    n = signal_dev.numel()
    bk = torch.zeros((n_pen, n), dtype=torch.int32, device=engine.tdev)
    nb = torch.zeros((n_pen,), dtype=torch.int32, device=engine.tdev)
    check(lib.jg_segment_scores(engine.ctx.handle, signal_dev.contiguous().data_ptr(), n, min_size, n_pen,
                                bk.data_ptr(), nb.data_ptr()))
    bk, nb = bk.cpu().numpy(), nb.cpu().numpy()
    return [bk[p, :nb[p]].tolist() for p in range(n_pen)]


def segment_contig(engine, logits: np.ndarray, phage_index: int, sensitivity: float = 1.5):
    """prophages.py:554-595 for one contig: (merged window-index ranges, scores of the kept ranges)."""
    with torch.cuda.stream(engine._stream()):
        sm = smooth(engine, engine._h2d(np.ascontiguousarray(logits, np.float32)))
        col_dev = sm[:, phage_index].contiguous()
        preds = breakpoints(engine, col_dev)
        col = col_dev.cpu().numpy()
    engine.ctx.sync()
    bkpts = [b for b in preds if len(b) > 1]
    if not bkpts:
        return [], np.array([])
    lens = [len(b) for b in bkpts]
    knee = knee_point(np.array(lens), list(range(len(bkpts))))
    try:
        idx = lens.index(knee) if knee else int(np.searchsorted(np.array(lens), 1))
        if idx == len(lens):
            return [], np.array([])                      # bkpts[None] raises in the reference -> no regions
        b = bkpts[idx]
        ranges = [b[i:i + 2] for i in range(len(b) - 1)]
        scores = np.array([col[s:e + 1].mean() for s, e in ranges])      # DataFrame.loc[s:e]: end-inclusive
        mask = scores > sensitivity
        return merge_overlapping_ranges(np.array(ranges)[mask]), scores[mask]
    except (ValueError, IndexError):
        return [], np.array([])


def call_regions(engine, data: dict, class_map: dict, fsize: int, stride: int, lc: int = 500_000,
                 sensitivity: float = 1.5, identifier: str = "phage") -> dict[str, dict]:
    """All contigs longer than `lc`: window-index ranges, scores and bp coordinates
    [start*stride, (end-1)*stride + fsize] (prophages.py:765-766)."""
    names = [c.lower() for c in class_map["class"]]
    if identifier not in names:
        return {}
    k = class_map["index"][names.index(identifier)]
    out = {}
    off = data["offsets"]
    for ci, (name, length) in enumerate(zip(data["headers"], data["length"])):
        if length < lc or length <= lc:                  # logits_to_df_v2 keeps >= lc, segment drops <= cutoff
            continue
        ranges, scores = segment_contig(engine, data["predictions"][off[ci]:off[ci + 1]], int(k), sensitivity)
        out[str(name)] = {"ranges": ranges, "scores": scores,
                          "coords": [(s * stride, (e - 1) * stride + fsize) for s, e in ranges]}
    return out
