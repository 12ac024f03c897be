"""Oracle (test infrastructure): terminal-repeat scan.

Restates `scan_for_terminal_repeats` / `get_alignment_summary` (utils/termini.py:17-189).  The
alignments themselves come from parasail 1.3.4 (`sw_trace_scan_16`), which is NOT vendored and not
installable here; this file restates the published algorithm it implements -- Smith-Waterman local
alignment with affine gaps (Gotoh): gap of length k costs open + (k-1)*extend, matrix
`matrix_create("ACGT", 2, -100)` (letters outside the alphabet score 0 against everything,
case-insensitive) -- with these conventions where the library leaves a choice:
  * end position = first maximum in column-major order (reference column, then query row), as a
    column-wise scan finds it;
  * traceback priority: H = 0 stops; diagonal > E (gap in the query line) > F (gap in the reference
    line); E and F prefer opening over extending on ties.
PARITY UNPINNED against parasail (the reference's own test mocks it, tests/unit/test_utils_termini.py).
The coordinate arithmetic of `get_alignment_summary` is restated verbatim.
"""
from __future__ import annotations

import numpy as np

NEG = -30000
_COMP = {"A": "T", "T": "A", "C": "G", "G": "C", "-": "-", "N": "N", "W": "W", "S": "S", "Y": "R", "R": "Y", "M": "K", "K": "M",
         "B": "V", "V": "B", "H": "D", "D": "H", "a": "T", "t": "A", "g": "C", "c": "G"}


def reverse_complement(seq: str) -> str:
    """seqops/transform.py:11-36."""
    return "".join(_COMP.get(b, "N") for b in reversed(seq))


def _codes(s: str) -> np.ndarray:
    lut = np.full(256, 4, dtype=np.int64)
    for k, ch in enumerate("ACTG"):
        lut[ord(ch)] = lut[ord(ch.lower())] = k
    return lut[np.frombuffer(s.encode("ascii", "replace"), dtype=np.uint8)]


def sw_align(query: str, ref: str, match=2, mismatch=-100, wild=0, gap_open=100, gap_ext=5) -> dict:
    """Local alignment with traceback: score, end_query, end_ref (0-based), cols (alignment columns),
    qgaps / rgaps (gap characters in the query / reference line), iden (identical pairs) and the two
    alignment lines `qline` / `rline` (result.traceback.query / .ref: the input letters, '-' for gaps)."""
    q, r = _codes(query), _codes(ref)
    m, n = len(q), len(r)
    if m == 0 or n == 0:
        return dict(score=0, end_query=0, end_ref=0, cols=0, qgaps=0, rgaps=0, iden=0, qline="", rline="")
    sub = np.where((q[:, None] > 3) | (r[None, :] > 3), wild, np.where(q[:, None] == r[None, :], match, mismatch))
    H = np.zeros((m + 1, n + 1), dtype=np.int32)
    E = np.full((m + 1, n + 1), NEG, dtype=np.int32)
    F = np.full((m + 1, n + 1), NEG, dtype=np.int32)
    src = np.zeros((m + 1, n + 1), dtype=np.int8)
    e_ext = np.zeros((m + 1, n + 1), dtype=bool)
    f_ext = np.zeros((m + 1, n + 1), dtype=bool)
    for d in range(2, m + n + 1):                 # 1-based cells (i, j) with i + j = d
        i = np.arange(max(1, d - n), min(m, d - 1) + 1)
        j = d - i
        ee, eo = E[i, j - 1] - gap_ext, H[i, j - 1] - gap_open
        fe, fo = F[i - 1, j] - gap_ext, H[i - 1, j] - gap_open
        e = np.maximum(ee, eo)
        f = np.maximum(fe, fo)
        hd = H[i - 1, j - 1] + sub[i - 1, j - 1]
        h = np.maximum(np.maximum(hd, 0), np.maximum(e, f))
        E[i, j], F[i, j], H[i, j] = e, f, h
        e_ext[i, j], f_ext[i, j] = ee > eo, fe > fo
        src[i, j] = np.where(h == 0, 0, np.where(h == hd, 1, np.where(h == e, 2, 3)))
    Hc = H[1:, 1:]
    score = int(Hc.max())
    if score == 0:
        return dict(score=0, end_query=0, end_ref=0, cols=0, qgaps=0, rgaps=0, iden=0, qline="", rline="")
    jj = int(np.flatnonzero((Hc == score).any(axis=0))[0])
    ii = int(np.flatnonzero(Hc[:, jj] == score)[0])
    i, j, state = ii + 1, jj + 1, 0
    cols = qg = rg = iden = 0
    ql: list[str] = []
    rl: list[str] = []
    while i >= 1 and j >= 1:
        if state == 0:
            s = src[i, j]
            if s == 0:
                break
            if s == 1:
                iden += int(q[i - 1] < 4 and q[i - 1] == r[j - 1])
                ql.append(query[i - 1]); rl.append(ref[j - 1])
                cols += 1; i -= 1; j -= 1
            else:
                state = 1 if s == 2 else 2
        elif state == 1:
            cols += 1; qg += 1
            ql.append("-"); rl.append(ref[j - 1])
            state = 1 if e_ext[i, j] else 0
            j -= 1
        else:
            cols += 1; rg += 1
            ql.append(query[i - 1]); rl.append("-")
            state = 2 if f_ext[i, j] else 0
            i -= 1
    return dict(score=score, end_query=ii, end_ref=jj, cols=cols, qgaps=qg, rgaps=rg, iden=iden,
                qline="".join(reversed(ql)), rline="".join(reversed(rl)))


def alignment_summary(res: dict, seq_len: int, record_id: str, input_length: int, type_: str) -> dict:
    """get_alignment_summary (termini.py:17-88) on the result of `sw_align`; `front` = traceback.query, `rear` =
    traceback.ref, reverse-complemented back to the contig's strand for an ITR (termini.py:57, 66, 83-84)."""
    alig_len, f_gaps, rc_gaps = res["cols"], res["qgaps"], res["rgaps"]
    s_start = (res["end_query"] - alig_len + f_gaps) + 1
    s_end = res["end_query"] + 1
    if type_ == "ITR":
        e_start = (seq_len - input_length) + max(input_length - res["end_ref"], 0)
        e_end = e_start + (alig_len - rc_gaps)
        rear = reverse_complement(res.get("rline", ""))
    else:
        rear = res.get("rline", "")
        e_start = (seq_len - input_length) + max(res["end_ref"] - alig_len, 0)
        e_end = (seq_len - input_length) + res["end_ref"]
        if (s_end - s_start) >= 250:
            type_ = f"LTR_{type_}"
    return {"contig_id": record_id, "repeat_length": alig_len, "identities": res["iden"],
            "identity": round(res["iden"] / alig_len, 2) if alig_len else 0, "score": res["score"], "terminal_repeats": type_,
            "fgaps": f_gaps, "rgaps": rc_gaps, "sstart": s_start, "send": s_end, "estart": e_start, "eend": e_end, "seq_len": seq_len,
            "front": res.get("qline", ""), "rear": rear}


EMPTY = {"repeat_length": None, "identities": None, "identity": None, "score": None, "terminal_repeats": None, "fgaps": None,
         "rgaps": None, "sstart": None, "send": None, "estart": None, "eend": None}
EMPTY_STRINGS = {"front": None, "rear": None}


def scan_for_terminal_repeats(records, fsize: int) -> list[dict]:
    """termini.py:91-189: one row per record with len >= fsize, in record order."""
    rows = []
    for name, seq in records:
        seq = seq.decode() if isinstance(seq, bytes) else seq
        seq_len = len(seq)
        if seq_len < fsize:
            continue
        header = name.replace(",", "___")
        n = min(max(int(seq_len * 0.04), 400), 4000)
        itr = sw_align(seq[:n], reverse_complement(seq[-n:]))
        dtr = sw_align(seq[:n], seq[-n:])
        if itr["cols"] > 12 or dtr["cols"] > 12:
            if itr["score"] > dtr["score"]:
                rows.append(alignment_summary(itr, seq_len, header, n, "ITR"))
            else:
                rows.append(alignment_summary(dtr, seq_len, header, n, "DTR"))
        else:
            rows.append({"contig_id": header, **EMPTY, "seq_len": seq_len, **EMPTY_STRINGS})
    return rows


# ---- att sites around prophage regions (postprocess/prophages.py:604-873) ----------------------------------------

def _gc_content(seq: str) -> float:
    """postprocess/helpers.py:710-723 (upper-case G / C only; an empty slice raises in the reference)."""
    return (seq.count("G") + seq.count("C")) / len(seq)


def prophage_alignment_summary(res, seq_len: int, name: str, seq: str, cordinates: dict, phage_score, type_):
    """get_prophage_alignment_summary (prophages.py:604-703); `res` = sw_align result or None."""
    if res is None:
        s, e = cordinates["start"][0], cordinates["end"][0]
        return {"contig_id": name, "seq_len": seq_len, "region_len": e - s, "phage_score": phage_score, "n%": None,
                "gc%": _gc_content(seq[s:e]), "reject": None, "sstart": s, "send": None, "estart": None, "eend": e,
                "att_alignment_length": None, "att_identities": None, "att_identity": None, "att_score": None, "att_type": None,
                "att_fgaps": None, "att_rgaps": None, "attL": None, "attR": None}
    alig_len = res["cols"]
    if type_ == "ITR":
        s_end = cordinates["start"][0] + res["end_query"] + 1
        s_start = s_end - alig_len
        e_start = cordinates["end"][1] - res["end_ref"] - 1
        e_end = e_start + alig_len
    else:
        s_end = cordinates["start"][0] + res["end_query"]
        s_start = s_end - alig_len + 1
        e_end = cordinates["end"][0] + res["end_ref"]
        e_start = e_end - alig_len + 1
        if (s_end - s_start) >= 250:
            type_ = f"LTR_{type_}"
    sub = seq[s_start:e_end]
    pn = sub.count("N") / len(sub)
    return {"contig_id": name, "seq_len": seq_len, "region_len": e_end - s_start, "phage_score": phage_score, "n%": pn,
            "gc%": _gc_content(sub), "reject": pn > 0.20, "sstart": s_start, "send": s_end, "estart": e_start, "eend": e_end,
            "att_alignment_length": alig_len, "att_identities": res["iden"], "att_identity": round(res["iden"] / alig_len, 2),
            "att_score": res["score"], "att_type": type_, "att_fgaps": res["qgaps"], "att_rgaps": res["rgaps"],
            "attL": res["qline"], "attR": res["rline"]}


def prophage_report(records, prophage_cordinates: dict, fsize: int, stride: int | None = None, refined_boundaries: dict | None = None) -> list[dict]:
    """prophage_report (prophages.py:706-873): one row per called region of every contig longer than 500 000 bp; with
    `refined_boundaries` (header -> [(raw_start, raw_end, refined_start, refined_end)]) the search and the coordinates use the
    gene-aware ends (prophages.py:759-772).  `prophage_cordinates`: header -> (window-index ranges,
    scores) as `segment` returns them."""
    step = stride or fsize
    rows = []
    for name, seq in records:
        seq = seq.decode() if isinstance(seq, bytes) else seq
        seq_len = len(seq)
        header = name.replace(",", "___")
        if seq_len <= 500_000:
            continue
        cords, scores = prophage_cordinates.get(header, [[], []])
        if not (len(cords) > 0 and len(scores) > 0):
            continue
        contig_refined = refined_boundaries.get(header) if refined_boundaries else None
        for idx, ((start, end), j) in enumerate(zip(cords, scores)):
            raw_start, raw_end = int(start * step), int((end - 1) * step + fsize)
            r_start, r_end = raw_start, raw_end
            if contig_refined is not None and idx < len(contig_refined):
                _, _, r_start, r_end = contig_refined[idx]
            region_len = r_end - r_start
            scan_length = min(max(int(seq_len * 0.04), 400), 4000)
            off_set = 2000 if region_len // 2 >= 14000 else region_len // 4
            search_start, search_end = max(r_start - scan_length, 0), min(r_end + scan_length, seq_len)
            left, right = seq[search_start:r_start + off_set], seq[r_end - off_set:search_end]
            none_cords = {"start": [r_start, None], "end": [r_end, None]}
            if not left or not right:
                row = prophage_alignment_summary(None, seq_len, name, seq, none_cords, j, None)
            else:
                dtr = sw_align(left, right)
                itr = sw_align(left, reverse_complement(right))
                cords_hit = {"start": [search_start, search_start + off_set], "end": [r_end - off_set, search_end]}
                if itr["cols"] > 12 or dtr["cols"] > 12:
                    if itr["score"] > dtr["score"]:
                        row = prophage_alignment_summary(itr, seq_len, name, seq, cords_hit, j, "ITR")
                    else:
                        row = prophage_alignment_summary(dtr, seq_len, name, seq, cords_hit, j, "DTR")
                else:
                    row = prophage_alignment_summary(None, seq_len, name, seq, none_cords, j, None)
            row["raw_start"], row["raw_end"] = raw_start, raw_end
            rows.append(row)
    return rows
