"""Oracle (test infrastructure): post-hoc refinement of window / contig calls (`jaeger predict --refine`).

NumPy restatement of postprocess/refinement.py:39-247 (`add_score_features`, `refine`, `aggregate_contig`) and of
the driver glue commands/predict.py:115-155 (`_build_refined_contig_df`).  The reference implements these on polars
(>= 1.0, not vendored, not installable here), so the module cannot be imported as it is.
Pinned: the known answers of the reference's own tests/unit/test_refinement.py, and `add_score_features` + `refine` + `aggregate_contig`
run from the reference's source behind a polars stand-in (tests/golden/make_refine_window_goldens.py); both restated
against this file in tests/test_postprocess_cpu.py.  PARITY UNPINNED: polars' own expression evaluation (summation order).  Conventions fixed here where the libraries leave a choice:
  * window logits are widened to float64 before any arithmetic (polars builds the frame from Python row dicts);
  * top class = np.argmax (first maximum); second class = position -2 of a STABLE ascending argsort (NumPy's
    argsort on 6 elements; ties between exactly equal logits are the only case where the sort kind matters);
  * per-contig sums are plain float64 sums in window order (polars does not specify a summation order: the
    comparison tolerance is 1e-9 relative).
Score columns are taken POSITIONALLY from the model's logits (predict.py:140: zip(SCORE_COLS, window_logits)).
"""
from __future__ import annotations

import numpy as np

SCORE_COLS = ["phage_score", "virus_score", "archaea_score", "bacteria_score", "plasmid_score", "eukarya_score"]
CLASSES = [c.replace("_score", "") for c in SCORE_COLS]
MERGE_MAP = {"bacteria_or_plasmid": ("bacteria", "plasmid"), "virus_any": ("phage", "virus")}


def add_score_features(S: np.ndarray) -> dict[str, np.ndarray]:
    """refinement.py:39-73 on a [W, 6] logit matrix."""
    S = np.asarray(S, dtype=np.float64)
    P = np.exp(S - S.max(axis=1, keepdims=True))
    P = P / P.sum(axis=1, keepdims=True)
    top2 = np.sort(S, axis=1)[:, -2:]
    order = np.argsort(S, axis=1, kind="stable")
    return {"top_logit": top2[:, 1], "second_logit": top2[:, 0], "margin": top2[:, 1] - top2[:, 0], "top_prob": P.max(axis=1),
            "entropy": -(P * np.log(P + 1e-12)).sum(axis=1), "top_class": np.array([CLASSES[i] for i in S.argmax(axis=1)], dtype=object),
            "second_class": np.array([CLASSES[i] for i in order[:, -2]], dtype=object)}


def refine(feat: dict[str, np.ndarray], taus: dict, merge_bp: bool = True, merge_pv: bool = True) -> np.ndarray:
    """refinement.py:97-137: merge rules first, then per-class abstain -> refined_prediction [W] (object)."""
    top_class, second_class = feat["top_class"], feat["second_class"]
    top_logit, margin = feat["top_logit"], feat["margin"]
    tau_logit = np.array([float(taus[c]["logit"]) for c in top_class])
    tau_margin = np.array([float(taus[c]["margin"]) for c in top_class])
    refined = top_class.copy()
    if merge_bp:
        m = (((top_class == "bacteria") & (second_class == "plasmid")) | ((top_class == "plasmid") & (second_class == "bacteria"))) \
            & (margin < tau_margin)
        refined[m] = "bacteria_or_plasmid"
    if merge_pv:
        m = (((top_class == "phage") & (second_class == "virus")) | ((top_class == "virus") & (second_class == "phage"))) \
            & (margin < tau_margin)
        refined[m] = "virus_any"
    abstain = ((top_logit < tau_logit) | (margin < tau_margin)) & ~np.isin(refined, list(MERGE_MAP.keys()))
    refined[abstain] = "unknown"
    return refined


def aggregate_contig(contig_ids, S, refined, margin, mode: str = "gated", min_windows: int = 3, merge_split: str = "half",
                     allow_merged_contig_call: bool = False, contig_hedge_margin: float = 1.0) -> dict[str, dict]:
    """refinement.py:140-247 -> {contig_id: row dict}; contigs with fewer than `min_windows` used windows are absent."""
    S = np.asarray(S, dtype=np.float64)
    contig_ids = np.asarray(contig_ids, dtype=object)
    keep = np.ones(len(S), dtype=bool)
    if mode in ("gated", "weighted"):
        keep = refined != "unknown"
    w = np.clip(np.asarray(margin, dtype=np.float64), 0.0, None) if mode == "weighted" else np.ones(len(S))
    share = 0.5 if merge_split == "half" else 1.0
    mult = np.ones((len(S), 6))
    for k, name in enumerate(CLASSES):
        with_class = [lbl for lbl, members in MERGE_MAP.items() if name in members]
        merged = np.isin(refined, list(MERGE_MAP.keys()))
        mult[:, k] = np.where(merged, np.where(np.isin(refined, with_class), share, 0.0), 1.0)
    out: dict[str, dict] = {}
    merge_pairs = {frozenset(members): lbl for lbl, members in MERGE_MAP.items()}
    for cid in dict.fromkeys(contig_ids[keep].tolist()):
        sel = keep & (contig_ids == cid)
        n_used = int(sel.sum())
        if n_used < min_windows:
            continue
        sums = (S[sel] * w[sel, None] * mult[sel]).sum(axis=0)
        order = np.argsort(sums, kind="stable")
        top, second = int(order[-1]), int(order[-2])
        cmargin = float(sums[top] - sums[second])
        call = CLASSES[top]
        if allow_merged_contig_call:
            pair = frozenset((CLASSES[top], CLASSES[second]))
            if cmargin < contig_hedge_margin and pair in merge_pairs:
                call = merge_pairs[pair]
        out[cid] = {"contig_id": cid, **{c: float(v) for c, v in zip(SCORE_COLS, sums)}, "n_windows_used": n_used,
                    "total_weight": float(w[sel].sum()), "n_merged_windows": int(np.isin(refined[sel], list(MERGE_MAP.keys())).sum()),
                    "contig_call": call, "contig_top_class": CLASSES[top], "contig_second_class": CLASSES[second],
                    "contig_top_logit": float(sums[top]), "contig_margin": cmargin}
    return out


def build_refined_contig(headers, predictions, taus: dict, **kw) -> dict[str, dict]:
    """_build_refined_contig_df (commands/predict.py:115-155): per-contig logits [T_i, 6] -> refined contig rows."""
    ids, rows = [], []
    for cid, logits in zip(headers, predictions):
        logits = np.asarray(logits)
        if logits.ndim != 2:
            continue
        ids += [cid] * len(logits)
        rows.append(logits[:, :6])
    if not rows:
        return {}
    S = np.concatenate(rows).astype(np.float64)
    feat = add_score_features(S)
    return aggregate_contig(ids, S, refine(feat, taus), feat["margin"], **kw)
