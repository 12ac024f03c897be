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


def _select_ranges(preds: list[list[int]], col: np.ndarray, sensitivity: float):
    """prophages.py:554-595 after the change-point search: keep the penalties with > 1 segment end, pick one by the knee of
    the breakpoint counts (searchsorted fallback), ranges = consecutive ends, end-inclusive means, sensitivity filter, merge."""
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


def segment_device(engine, logits_dev: torch.Tensor, offsets: np.ndarray, phage_index: int, sensitivity: float = 1.5,
                   n_pen: int = 9, min_size: int = 3):
    """Stage 4b for ALL contigs of a run at once: logits_dev [W, n_cls] fp32 in HBM, contig c = windows
    offsets[c]:offsets[c+1].  One smoothing launch and one change-point launch pair cover every contig (jg_smooth_scores,
    jg_segment_scores_batched: grid = penalties x contigs); one D2H brings the segment ends and the phage column back for the
    scalar knee / filter / merge logic.  Returns [(ranges, scores)] per contig."""
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    nc, w = len(offsets) - 1, int(offsets[-1])
    if nc <= 0 or w <= 0:
        return [([], np.array([]))] * max(nc, 0)
    n_cls = logits_dev.shape[1]
    with torch.cuda.stream(engine._stream()):
        off_dev = engine._h2d(offsets)
        sm = torch.empty((w, n_cls), dtype=torch.float64, device=engine.tdev)
        check(lib.jg_smooth_scores(engine.ctx.handle, logits_dev.data_ptr(), off_dev.data_ptr(), nc, n_cls, 4, sm.data_ptr()))
        col_dev = sm[:, phage_index].contiguous()
        bk = torch.zeros((n_pen, w), dtype=torch.int32, device=engine.tdev)
        nb = torch.zeros((nc, n_pen), dtype=torch.int32, device=engine.tdev)
        check(lib.jg_segment_scores_batched(engine.ctx.handle, col_dev.data_ptr(), off_dev.data_ptr(), nc, w, min_size, n_pen,
                                            bk.data_ptr(), nb.data_ptr()))
        bk_h, nb_h, col_h = bk.cpu(), nb.cpu(), col_dev.cpu()
    engine.ctx.sync()
    bk_h, nb_h, col_h = bk_h.numpy(), nb_h.numpy(), col_h.numpy()
    out = []
    for c in range(nc):
        a, b = int(offsets[c]), int(offsets[c + 1])
        preds = [bk_h[p, a:a + nb_h[c, p]].tolist() for p in range(n_pen)]
        out.append(_select_ranges(preds, col_h[a:b], sensitivity))
    return out


def segment_contig(engine, logits: np.ndarray, phage_index: int, sensitivity: float = 1.5):
    """prophages.py:554-595 for one contig: (merged window-index ranges, scores of the kept ranges)."""
    logits = np.ascontiguousarray(logits, np.float32)
    with torch.cuda.stream(engine._stream()):
        dev = engine._h2d(logits)
    return segment_device(engine, dev, np.array([0, len(logits)], dtype=np.int64), phage_index, sensitivity)[0]


def call_regions(engine, data: dict, class_map: dict, fsize: int, stride: int, lc: int = 500_000,
                 sensitivity: float = 1.5, identifier: str = "phage") -> dict[str, dict]:
    """All contigs longer than `lc`: window-index ranges, scores and bp coordinates
    [start*stride, (end-1)*stride + fsize] (prophages.py:765-766).  The selected contigs' window logits are taken from the
    copy that is still in HBM when `data` comes from `contig_table` of this engine's own result."""
    names = [c.lower() for c in class_map["class"]]
    if identifier not in names:
        return {}
    k = class_map["index"][names.index(identifier)]
    off = np.asarray(data["offsets"], dtype=np.int64)
    sel = [ci for ci, length in enumerate(data["length"]) if not (length < lc or length <= lc)]   # logits_to_df_v2 keeps >= lc, segment drops <= cutoff
    if not sel:
        return {}
    n_win = np.array([off[ci + 1] - off[ci] for ci in sel], dtype=np.int64)
    sub_off = np.concatenate([[0], np.cumsum(n_win)]).astype(np.int64)
    rows = np.concatenate([np.arange(off[ci], off[ci + 1]) for ci in sel])
    with torch.cuda.stream(engine._stream()):
        pred_dev = data.get("_pred_dev")
        if pred_dev is not None:
            logits_dev = pred_dev if len(rows) == pred_dev.shape[0] else pred_dev[engine._h2d(rows)]
        else:
            logits_dev = engine._h2d(np.ascontiguousarray(data["predictions"][rows], np.float32))
    res = segment_device(engine, logits_dev, sub_off, int(k), sensitivity)
    out = {}
    for ci, (ranges, scores) in zip(sel, res):
        out[str(data["headers"][ci])] = {"ranges": ranges, "scores": scores,
                                         "coords": [(s * stride, (e - 1) * stride + fsize) for s, e in ranges]}
    return out
