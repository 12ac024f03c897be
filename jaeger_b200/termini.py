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

# jg_sw_job / jg_sw_trace_job (include/jaeger_b200.h): nq = 0 means a square job (query length = n)
JOB = np.dtype([("q0", "<i8"), ("r0", "<i8"), ("n", "<i4"), ("inverted", "<i4"), ("nq", "<i4"), ("reserved", "<i4")])
TRACE_JOB = np.dtype([("q0", "<i8"), ("r0", "<i8"), ("n", "<i4"), ("inverted", "<i4"), ("nq", "<i4"), ("reserved", "<i4"),
                      ("end_i", "<i4"), ("end_j", "<i4"), ("dirs_off", "<i8"), ("ops_off", "<i8")])
JOB_FIELDS = ("q0", "r0", "n", "inverted", "nq")
COLUMNS = ["contig_id", "repeat_length", "identities", "identity", "score", "terminal_repeats", "fgaps", "rgaps", "sstart",
           "send", "estart", "eend", "seq_len", "front", "rear"]
TRACE_MIN_SCORE = 104          # below: no gap / mismatch can be on the path (each costs 100 and needs 51 matches either side)
SCRATCH_BYTES = 1 << 30


_RC = np.full(256, ord("N"), dtype=np.uint8)          # seqops/transform.py:11-36: unknown letters become N
for _a, _b in {"A": "T", "T": "A", "C": "G", "G": "C", "-": "-", "N": "N", "W": "W", "S": "S", "Y": "R", "R": "Y", "M": "K", "K": "M",
               "B": "V", "V": "B", "H": "D", "D": "H", "a": "T", "t": "A", "g": "C", "c": "G"}.items():
    _RC[ord(_a)] = ord(_b)


def reverse_complement_bytes(a: np.ndarray) -> np.ndarray:
    return _RC[a[::-1]]


def alignment_lines(host: np.ndarray, job, end_i: int, end_j: int, ops: np.ndarray) -> tuple[str, str]:
    """result.traceback.query / .ref of one job from the traceback operations (`ops`: last column first;
    1 pair, 2 gap in the query line, 3 gap in the reference line): the input letters as the FASTA has
    them, '-' for gaps.  The reference line of an inverted job is read off the reverse complement."""
    n, nq = int(job["n"]), int(job["nq"]) or int(job["n"])
    q = host[int(job["q0"]):int(job["q0"]) + nq]
    r = host[int(job["r0"]):int(job["r0"]) + n]
    if int(job["inverted"]):
        r = reverse_complement_bytes(r)
    ops = np.asarray(ops, dtype=np.uint8)
    qi = end_i - np.cumsum(ops != 2) + 1          # row consumed by column c (valid where ops != 2)
    rj = end_j - np.cumsum(ops != 3) + 1
    dash = np.uint8(ord("-"))
    ql = np.where(ops != 2, q[np.clip(qi, 0, nq - 1)], dash)[::-1]
    rl = np.where(ops != 3, r[np.clip(rj, 0, n - 1)], dash)[::-1]
    return ql.tobytes().decode("ascii", "replace"), rl.tobytes().decode("ascii", "replace")


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


def scan_terminal_repeats(engine, codes: torch.Tensor, valid: torch.Tensor, offsets: np.ndarray, names: list[str], fsize: int,
                          host: np.ndarray | None = None):
    """codes / valid: the packed contigs on the device (engine.pack); offsets [n+1]; names as the
    FASTA gives them.  Returns a pandas DataFrame with the reference's columns, one row per contig
    with len >= fsize.  `host`: the ASCII bases the contigs were packed from; with it the `front` /
    `rear` alignment strings (termini.py:83-84) are materialised, without it they stay None."""
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
    ops: dict[int, np.ndarray] = {}
    if len(need):
        tj = np.zeros(len(need), dtype=TRACE_JOB)
        sel = 2 * need + use_itr[need].astype(np.int64)
        for k in JOB_FIELDS:
            tj[k] = jobs[k][sel]
        tj["end_i"], tj["end_j"] = win[need, 1], win[need, 2]
        counts[need], traced = _run_trace(engine, codes, valid, tj, want_ops=host is not None)
        ops = dict(zip(need.tolist(), traced)) if host is not None else {}
    rows = []
    for c in range(len(ids)):
        header = names[ids[c]].replace(",", "___")
        seq_len, nn = int(L[c]), int(n[c])
        if not hit[c]:
            rows.append({"contig_id": header, **{k: None for k in COLUMNS[1:]}, "seq_len": seq_len})
            continue
        alig_len, f_gaps, rc_gaps, iden = (int(x) for x in counts[c])
        front = rear = None
        if host is not None:
            job = jobs[2 * c + int(use_itr[c])]
            front, rear = alignment_lines(host, job, int(win[c, 1]), int(win[c, 2]), ops.get(c, np.ones(alig_len, np.uint8)))
            if use_itr[c]:                                               # termini.py:57: back on the contig's own strand
                rear = reverse_complement_bytes(np.frombuffer(rear.encode(), np.uint8)).tobytes().decode()
        rows.append({"contig_id": header, **summary_fields(alig_len, f_gaps, rc_gaps, iden, int(win[c, 0]), int(win[c, 1]), int(win[c, 2]),
                                                            seq_len, nn, "ITR" if use_itr[c] else "DTR"), "front": front, "rear": rear})
    return pd.DataFrame(rows, columns=COLUMNS)


def scan_source(engine, src, fsize: int):
    """The scan for a WindowSource: H2D of the (already loaded, pinned) bases, pack, scan."""
    names, host, offsets = src.load()
    with torch.cuda.stream(engine._stream()):
        codes, valid = engine.pack(host.to(engine.tdev, non_blocking=True))
    return scan_terminal_repeats(engine, codes, valid, offsets, [n.strip() for n in names], fsize, host=host.numpy())


def _run_scan(engine, codes, valid, jobs: np.ndarray) -> np.ndarray:
    rows_of = np.where(jobs["nq"] > 0, jobs["nq"], jobs["n"])
    order = np.argsort(rows_of, kind="stable")
    sorted_jobs, rows_sorted = jobs[order], rows_of[order]
    threads = np.array([_threads_for(int(x)) for x in rows_sorted])
    out = np.zeros((len(jobs), 4), dtype=np.int64)
    with torch.cuda.stream(engine._stream()):
        d_jobs = engine._h2d(sorted_jobs.view(np.uint8).reshape(len(jobs), JOB.itemsize))
        d_out = engine._empty((len(jobs), 4), torch.int32)
        start = 0
        while start < len(jobs):
            t = int(threads[start])
            end = int(np.searchsorted(threads, t, side="right"))
            check(lib.jg_sw_scan(engine.ctx.handle, codes.data_ptr(), valid.data_ptr(), d_jobs.data_ptr() + start * JOB.itemsize,
                                 end - start, t, int(rows_sorted[end - 1]), int(sorted_jobs["n"][start:end].max()),
                                 d_out.data_ptr() + start * 16))
            start = end
        host = d_out.cpu().numpy()
    engine.ctx.sync()
    out[order] = host
    return out


def _run_trace(engine, codes, valid, tj: np.ndarray, want_ops: bool = False):
    """-> (counts [n, 4], per-job traceback operations (last column first) or None)."""
    out = np.zeros((len(tj), 4), dtype=np.int64)
    ops_out: list[np.ndarray | None] = [None] * len(tj)
    ops_len = (tj["end_i"].astype(np.int64) + tj["end_j"].astype(np.int64) + 2) if want_ops else np.zeros(len(tj), np.int64)
    cells = (tj["end_i"].astype(np.int64) + 1) * (tj["end_j"].astype(np.int64) + 1) + ops_len
    tj["ops_off"] = -1
    with torch.cuda.stream(engine._stream()):
        scratch = engine._empty((int(min(SCRATCH_BYTES, max(int(cells.max()), int(cells.sum())))),), torch.uint8)
        k = 0
        while k < len(tj):                                   # waves that fit the scratch buffer, one launch per CTA size
            used, e = 0, k
            while e < len(tj) and (used + cells[e] <= scratch.numel() or e == k):
                tj["dirs_off"][e] = used
                if want_ops:
                    tj["ops_off"][e] = used + int(cells[e]) - int(ops_len[e])
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
            ops_host = [scratch[int(o):int(o) + int(n)].cpu().numpy() for o, n in zip(wave["ops_off"], ops_len[k:e])] if want_ops else []
            engine.ctx.sync()
            out[k + order] = res
            for w in range(e - k):
                if want_ops:
                    ops_out[k + w] = ops_host[w][:int(out[k + w, 0])]
            k = e
    return out, (ops_out if want_ops else None)


# ---- att sites around prophage regions: prophage_report (postprocess/prophages.py:706-873) -------------------------

PROPHAGE_COLUMNS = ["contig_id", "seq_len", "region_len", "phage_score", "n%", "gc%", "reject", "sstart", "send", "estart", "eend",
                    "att_alignment_length", "att_identities", "att_identity", "att_score", "att_type", "att_fgaps", "att_rgaps",
                    "attL", "attR", "raw_start", "raw_end"]


def _gc_and_n(contig: np.ndarray, s: int, e: int) -> tuple[float, float]:
    """calculate_gc_content / calculate_percentage_of_n (postprocess/helpers.py:710-739) of contig[s:e] (Python slice
    semantics, upper-case letters only; the reference divides by zero on an empty slice)."""
    sub = contig[s:e]
    if len(sub) == 0:
        raise ZeroDivisionError("empty prophage region slice")
    return float(((sub == ord("G")) | (sub == ord("C"))).sum()) / len(sub), float((sub == ord("N")).sum()) / len(sub)


def prophage_report(engine, codes: torch.Tensor, valid: torch.Tensor, host: np.ndarray, offsets: np.ndarray, names: list[str],
                    regions: dict, fsize: int, stride: int | None = None, refined_boundaries: dict | None = None):
    """The att-site search of `prophage_report` for the called regions.  `refined_boundaries` (header ->
    [(raw_start, raw_end, refined_start, refined_end)], prophage_boundaries.refine_regions): the gene-aware ends the search and
    the reported coordinates use instead of the window-grid ones (prophages.py:759-772); None = raw boundaries.  Per region of a contig longer than 500 000 bp, the left flank [start - scan, start + off_set) is
    aligned against the right flank [end - off_set, end + scan) directly and against its reverse complement -- the
    same Smith-Waterman as the terminal-repeat scan, run as rectangular `jg_sw_scan` / `jg_sw_trace` jobs on the packed
    contigs -- and the better one, if either is longer than 12 columns, gives attL / attR and the region coordinates
    (`get_prophage_alignment_summary`, prophages.py:604-703).  `regions`: header -> {"ranges", "scores"} as
    `prophage.call_regions` returns it.  Returns a DataFrame with the reference's columns (contig ids as in the FASTA)."""
    import pandas as pd
    step = stride or fsize
    lens = np.diff(offsets)
    items, jobs = [], []
    for ci in np.flatnonzero(lens > 500_000):                                  # prophages.py:759
        reg = regions.get(names[ci].replace(",", "___")) or regions.get(names[ci])
        if not reg or len(reg["ranges"]) == 0 or len(reg["scores"]) == 0:
            continue
        L, base = int(lens[ci]), int(offsets[ci])
        header = names[ci].replace(",", "___")
        contig_refined = (refined_boundaries.get(header) or refined_boundaries.get(names[ci])) if refined_boundaries else None
        for idx, ((start, end), sc) in enumerate(zip(reg["ranges"], reg["scores"])):
            raw_start, raw_end = int(start * step), int((end - 1) * step + fsize)                    # prophages.py:765-766
            ref_start, ref_end = raw_start, raw_end
            if contig_refined is not None and idx < len(contig_refined):                              # prophages.py:767-770
                ref_start, ref_end = int(contig_refined[idx][2]), int(contig_refined[idx][3])
            region_len = ref_end - ref_start
            scan_length = min(max(int(L * 0.04), 400), 4000)
            off_set = 2000 if region_len // 2 >= 14000 else region_len // 4
            search_start, search_end = max(ref_start - scan_length, 0), min(ref_end + scan_length, L)
            l0, l1, _ = slice(search_start, ref_start + off_set).indices(L)
            r0, r1, _ = slice(ref_end - off_set, search_end).indices(L)
            nl, nr = max(l1 - l0, 0), max(r1 - r0, 0)
            job_id = -1
            if nl > 0 and nr > 0:
                job_id = len(jobs)
                jobs += [(base + l0, base + r0, nr, 0, nl, 0), (base + l0, base + r0, nr, 1, nl, 0)]
            items.append(dict(ci=ci, L=L, base=base, score=sc, raw_start=raw_start, raw_end=raw_end, ref_start=ref_start, ref_end=ref_end, off_set=off_set,
                              search_start=search_start, search_end=search_end, job=job_id))
    if not items:
        return pd.DataFrame(columns=PROPHAGE_COLUMNS)
    jobs = np.array(jobs, dtype=JOB) if jobs else np.zeros(0, dtype=JOB)
    res = _run_scan(engine, codes, valid, jobs) if len(jobs) else np.zeros((0, 4), np.int64)
    picks = []                                          # (item index, winning job index) of regions with an alignment > 12 columns
    for k, it in enumerate(items):
        if it["job"] < 0:
            continue
        dtr, itr = res[it["job"]], res[it["job"] + 1]
        cols_d = 13 if dtr[0] >= TRACE_MIN_SCORE else dtr[3]
        cols_i = 13 if itr[0] >= TRACE_MIN_SCORE else itr[3]
        if cols_i > 12 or cols_d > 12:                                          # prophages.py:802-806
            picks.append((k, it["job"] + int(itr[0] > dtr[0])))
    counts, ops = {}, {}
    need = [(k, j) for k, j in picks if res[j, 0] >= TRACE_MIN_SCORE]
    if need:
        tj = np.zeros(len(need), dtype=TRACE_JOB)
        sel = np.array([j for _, j in need])
        for f in JOB_FIELDS:
            tj[f] = jobs[f][sel]
        tj["end_i"], tj["end_j"] = res[sel, 1], res[sel, 2]
        cnt, traced = _run_trace(engine, codes, valid, tj, want_ops=True)
        for (k, _), c, o in zip(need, cnt, traced):
            counts[k], ops[k] = c, o
    winner = dict(picks)
    rows = []
    for k, it in enumerate(items):
        contig = host[it["base"]:it["base"] + it["L"]]
        name = names[it["ci"]]
        if k not in winner:                                                     # result_object is None, prophages.py:624-650
            s, e = it["ref_start"], it["ref_end"]
            row = {"contig_id": name, "seq_len": it["L"], "region_len": e - s, "phage_score": it["score"], "n%": None,
                   "gc%": _gc_and_n(contig, s, e)[0], "reject": None, "sstart": s, "send": None, "estart": None, "eend": e,
                   **{c: None for c in PROPHAGE_COLUMNS[11:20]}}
        else:
            j = winner[k]
            score, end_q, end_r, run = (int(x) for x in res[j])
            alig_len, f_gaps, rc_gaps, iden = (int(x) for x in counts[k]) if k in counts else (run, 0, 0, score // 2)
            att_l, att_r = alignment_lines(host, jobs[j], end_q, end_r, ops.get(k, np.ones(alig_len, np.uint8)))
            type_ = "ITR" if jobs[j]["inverted"] else "DTR"
            if type_ == "ITR":                                                  # prophages.py:663-667
                s_end = it["search_start"] + end_q + 1
                s_start = s_end - alig_len
                e_start = it["search_end"] - end_r - 1
                e_end = e_start + alig_len
            else:                                                               # prophages.py:668-675
                s_end = it["search_start"] + end_q
                s_start = s_end - alig_len + 1
                e_end = (it["ref_end"] - it["off_set"]) + end_r
                e_start = e_end - alig_len + 1
                if (s_end - s_start) >= 250:
                    type_ = f"LTR_{type_}"
            gc, pn = _gc_and_n(contig, s_start, e_end)
            row = {"contig_id": name, "seq_len": it["L"], "region_len": e_end - s_start, "phage_score": it["score"], "n%": pn,
                   "gc%": gc, "reject": pn > 0.20, "sstart": s_start, "send": s_end, "estart": e_start, "eend": e_end,
                   "att_alignment_length": alig_len, "att_identities": iden, "att_identity": round(iden / alig_len, 2),
                   "att_score": score, "att_type": type_, "att_fgaps": f_gaps, "att_rgaps": rc_gaps, "attL": att_l, "attR": att_r}
        row["raw_start"], row["raw_end"] = it["raw_start"], it["raw_end"]
        rows.append(row)
    return pd.DataFrame(rows, columns=PROPHAGE_COLUMNS)


def prophage_report_loaded(engine, loaded, regions: dict, fsize: int, stride: int | None = None, refined_boundaries: dict | None = None):
    """`prophage_report` for a loaded FASTA (`WindowSource.load()`: names, pinned ASCII bases, offsets): H2D + pack, then
    the att-site scans.  Nothing is copied when no called region lies on a contig longer than 500 000 bp."""
    names, host, offsets = loaded
    names = [n.strip() for n in names]
    lens = np.diff(offsets)
    if not any(lens[i] > 500_000 and len((regions.get(n.replace(",", "___")) or regions.get(n) or {}).get("ranges", []))
               for i, n in enumerate(names)):
        import pandas as pd
        return pd.DataFrame(columns=PROPHAGE_COLUMNS)
    with torch.cuda.stream(engine._stream()):
        codes, valid = engine.pack(host.to(engine.tdev, non_blocking=True))
    return prophage_report(engine, codes, valid, host.numpy(), offsets, names, regions, fsize, stride, refined_boundaries)


def write_prophage_report(df, outdir) -> None:
    """prophages.py:866-873: `prophages_jaeger.tsv` in the <base>_prophages directory, only when there are rows."""
    from pathlib import Path
    if len(df):
        df = df.copy()
        df["contig_id"] = df["contig_id"].apply(lambda x: x.replace("___", ","))
        Path(outdir).mkdir(parents=True, exist_ok=True)
        df.to_csv(Path(outdir) / "prophages_jaeger.tsv", sep="\t", index=False, float_format="%.3f")
