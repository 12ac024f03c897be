"""Terminal-repeat scan on the device: `scan_for_terminal_repeats` (utils/termini.py:91-189).

Per contig with len >= fsize the reference aligns the first scan_length bases against the last
scan_length bases (direct repeat, DTR) and against their reverse complement (inverted repeat, ITR)
with parasail's Smith-Waterman (match 2, mismatch -100, gap open 100 / extend 5), keeps the better
one if either alignment is longer than 12 columns, and derives the repeat coordinates
(`get_alignment_summary`, termini.py:17-88).  Here both alignments of every contig are one launch
family of `jg_sw_scan` (bucketed by CTA size); only winners scoring >= 104 -- the only ones that can
contain a gap or a mismatch -- go through `jg_sw_trace` for the exact column / gap / identity counts.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from ._cabi import check, lib

JOB = np.dtype([("q0", "<i8"), ("r0", "<i8"), ("n", "<i4"), ("inverted", "<i4")])
TRACE_JOB = np.dtype([("q0", "<i8"), ("r0", "<i8"), ("n", "<i4"), ("inverted", "<i4"), ("end_i", "<i4"), ("end_j", "<i4"),
                      ("dirs_off", "<i8")])
COLUMNS = ["contig_id", "repeat_length", "identities", "identity", "score", "terminal_repeats", "fgaps", "rgaps", "sstart",
           "send", "estart", "eend", "seq_len", "front", "rear"]
TRACE_MIN_SCORE = 104          # below: no gap / mismatch can be on the path (each costs 100 and needs 51 matches either side)
SCRATCH_BYTES = 1 << 30


def scan_lengths(lens: np.ndarray) -> np.ndarray:
    """termini.py:110: min(max(int(len * 0.04), 400), 4000)."""
    return np.minimum(np.maximum((lens.astype(np.float64) * 0.04).astype(np.int64), 400), 4000)


def summary_fields(alig_len: int, f_gaps: int, rc_gaps: int, iden: int, score: int, end_query: int, end_ref: int,
                   seq_len: int, input_length: int, type_: str) -> dict[str, Any]:
    """get_alignment_summary (termini.py:17-88) from the traceback counts instead of the strings
    (identity = safe_divide, utils/misc.py:117-123: rounded to 2 decimals)."""
    s_start, s_end = (end_query - alig_len + f_gaps) + 1, end_query + 1                 # termini.py:49-50
    if type_ == "ITR":
        e_start = (seq_len - input_length) + max(input_length - end_ref, 0)             # termini.py:53-56
        e_end = e_start + (alig_len - rc_gaps)
    else:
        e_start = (seq_len - input_length) + max(end_ref - alig_len, 0)                 # termini.py:59-62
        e_end = (seq_len - input_length) + end_ref
        if (s_end - s_start) >= 250:
            type_ = f"LTR_{type_}"
    return {"repeat_length": alig_len, "identities": iden, "identity": round(iden / alig_len, 2) if alig_len else 0, "score": score,
            "terminal_repeats": type_, "fgaps": f_gaps, "rgaps": rc_gaps, "sstart": s_start, "send": s_end, "estart": e_start,
            "eend": e_end, "seq_len": seq_len}


def _threads_for(n: int) -> int:
    return max(32, ((n + 15) // 16 + 31) // 32 * 32)


def scan_terminal_repeats(engine, codes: torch.Tensor, valid: torch.Tensor, offsets: np.ndarray, names: list[str], fsize: int):
    """codes / valid: the packed contigs on the device (engine.pack); offsets [n+1]; names as the
    FASTA gives them.  Returns a pandas DataFrame with the reference's columns, one row per contig
    with len >= fsize (front / rear alignment strings are not materialised: None)."""
    import pandas as pd
    lens = np.diff(offsets)
    ids = np.flatnonzero(lens >= fsize)
    if len(ids) == 0:
        return pd.DataFrame(columns=COLUMNS)
    L, off = lens[ids], offsets[:-1][ids]
    n = scan_lengths(L)
    jobs = np.zeros(2 * len(ids), dtype=JOB)              # job 2c: DTR, 2c+1: ITR
    jobs["q0"] = np.repeat(off, 2)
    jobs["r0"] = np.repeat(off + L - n, 2)
    jobs["n"] = np.repeat(n, 2)
    jobs["inverted"] = np.tile([0, 1], len(ids))
    res = _run_scan(engine, codes, valid, jobs)           # [2c][4]: score, end_query, end_ref, diagonal run
    dtr, itr = res[0::2], res[1::2]
    # a path scoring < 104 is a pure diagonal run, so its column count is the run length
    cols_d = np.where(dtr[:, 0] >= TRACE_MIN_SCORE, 13, dtr[:, 3])
    cols_i = np.where(itr[:, 0] >= TRACE_MIN_SCORE, 13, itr[:, 3])
    hit = (cols_i > 12) | (cols_d > 12)                                  # termini.py:130
    use_itr = itr[:, 0] > dtr[:, 0]                                      # termini.py:131
    win = np.where(use_itr[:, None], itr, dtr)
    counts = np.stack([win[:, 3], np.zeros(len(ids), np.int64), np.zeros(len(ids), np.int64), win[:, 0] // 2], axis=1)
    need = np.flatnonzero(hit & (win[:, 0] >= TRACE_MIN_SCORE))
    if len(need):
        tj = np.zeros(len(need), dtype=TRACE_JOB)
        sel = 2 * need + use_itr[need].astype(np.int64)
        for k in ("q0", "r0", "n", "inverted"):
            tj[k] = jobs[k][sel]
        tj["end_i"], tj["end_j"] = win[need, 1], win[need, 2]
        counts[need] = _run_trace(engine, codes, valid, tj)
    rows = []
    for c in range(len(ids)):
        header = names[ids[c]].replace(",", "___")
        seq_len, nn = int(L[c]), int(n[c])
        if not hit[c]:
            rows.append({"contig_id": header, **{k: None for k in COLUMNS[1:]}, "seq_len": seq_len})
            continue
        alig_len, f_gaps, rc_gaps, iden = (int(x) for x in counts[c])
        rows.append({"contig_id": header, **summary_fields(alig_len, f_gaps, rc_gaps, iden, int(win[c, 0]), int(win[c, 1]), int(win[c, 2]),
                                                            seq_len, nn, "ITR" if use_itr[c] else "DTR"), "front": None, "rear": None})
    return pd.DataFrame(rows, columns=COLUMNS)


def scan_source(engine, src, fsize: int):
    """The scan for a WindowSource: H2D of the (already loaded, pinned) bases, pack, scan."""
    names, host, offsets = src.load()
    with torch.cuda.stream(engine._stream()):
        codes, valid = engine.pack(host.to(engine.tdev, non_blocking=True))
    return scan_terminal_repeats(engine, codes, valid, offsets, [n.strip() for n in names], fsize)


def _run_scan(engine, codes, valid, jobs: np.ndarray) -> np.ndarray:
    order = np.argsort(jobs["n"], kind="stable")
    sorted_jobs = jobs[order]
    threads = np.array([_threads_for(int(x)) for x in sorted_jobs["n"]])
    out = np.zeros((len(jobs), 4), dtype=np.int64)
    with torch.cuda.stream(engine._stream()):
        d_jobs = engine._h2d(sorted_jobs.view(np.uint8).reshape(len(jobs), JOB.itemsize))
        d_out = engine._empty((len(jobs), 4), torch.int32)
        start = 0
        while start < len(jobs):
            t = int(threads[start])
            end = int(np.searchsorted(threads, t, side="right"))
            check(lib.jg_sw_scan(engine.ctx.handle, codes.data_ptr(), valid.data_ptr(), d_jobs.data_ptr() + start * JOB.itemsize,
                                 end - start, t, int(sorted_jobs["n"][end - 1]), d_out.data_ptr() + start * 16))
            start = end
        host = d_out.cpu().numpy()
    engine.ctx.sync()
    out[order] = host
    return out


def _run_trace(engine, codes, valid, tj: np.ndarray) -> np.ndarray:
    out = np.zeros((len(tj), 4), dtype=np.int64)
    cells = (tj["end_i"].astype(np.int64) + 1) * (tj["end_j"].astype(np.int64) + 1)
    with torch.cuda.stream(engine._stream()):
        scratch = engine._empty((int(min(SCRATCH_BYTES, max(int(cells.max()), int(cells.sum())))),), torch.uint8)
        k = 0
        while k < len(tj):                                   # waves that fit the scratch buffer, one launch per CTA size
            used, e = 0, k
            while e < len(tj) and (used + cells[e] <= scratch.numel() or e == k):
                tj["dirs_off"][e] = used
                used += int(cells[e]); e += 1
            wave = tj[k:e]
            order = np.argsort(wave["end_i"], kind="stable")
            ws = wave[order]
            threads = np.array([_threads_for(int(x) + 1) for x in ws["end_i"]])
            d_jobs = engine._h2d(ws.view(np.uint8).reshape(len(ws), TRACE_JOB.itemsize))
            d_out = engine._empty((len(ws), 4), torch.int32)
            s = 0
            while s < len(ws):
                t = int(threads[s])
                s2 = int(np.searchsorted(threads, t, side="right"))
                check(lib.jg_sw_trace(engine.ctx.handle, codes.data_ptr(), valid.data_ptr(), d_jobs.data_ptr() + s * TRACE_JOB.itemsize,
                                      s2 - s, t, int(ws["end_i"][s2 - 1]) + 1, int(ws["end_j"][s:s2].max()) + 1,
                                      scratch.data_ptr(), d_out.data_ptr() + s * 16))
                s = s2
            res = d_out.cpu().numpy()
            engine.ctx.sync()
            out[k + order] = res
            k = e
    return out
