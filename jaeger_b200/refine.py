"""Post-hoc refinement of the window / contig calls (`jaeger predict --refine`) on the device.

Counterpart of postprocess/refinement.py (`add_score_features`, `refine`, `aggregate_contig`, `load_refinement`) as
the reference driver uses it (commands/predict.py:115-155, 310-335): per-window top-2 features, merge / abstain rules
and the per-contig gated / weighted / unweighted sums are one `jg_refine_contigs` call (two kernels: a per-window
pass and a warp-per-contig segmented reduction); the contig-level top-2 selection over <= 6 sums and the table
assembly stay on the host.  The per-model thresholds come from `<model>_refine.yaml` next to the graph directory.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any

import numpy as np
import torch
import yaml

from ._cabi import check, lib

SCORE_COLS = ["phage_score", "virus_score", "archaea_score", "bacteria_score", "plasmid_score", "eukarya_score"]
CLASSES = [c.replace("_score", "") for c in SCORE_COLS]
MERGE_MAP = {"bacteria_or_plasmid": ("bacteria", "plasmid"), "virus_any": ("phage", "virus")}
LABELS = CLASSES + ["unknown", "bacteria_or_plasmid", "virus_any"]        # codes of jg_refine_contigs' d_label
MODES = {"gated": 0, "weighted": 1, "unweighted": 2}
MERGE_COLUMNS = ["contig_id", "contig_call", "contig_top_logit", "contig_margin", "n_windows_used", "n_merged_windows"]


def load_refinement(path: str | Path, expect_model: str | None = None) -> dict[str, Any]:
    """refinement.py:281-298: schema and model checks of a calibration file."""
    meta = yaml.safe_load(Path(path).read_text())
    if meta.get("schema_version", 0) != 1:
        raise ValueError(f"Unsupported refinement schema version: {meta.get('schema_version')}")
    if expect_model is not None and meta["jaeger_model"] != expect_model:
        raise ValueError(f"Refinement file was calibrated for {meta['jaeger_model']}, but current model is {expect_model}. "
                         "Recalibrate before using.")
    return meta


def tau_vector(taus: dict[str, dict[str, float]]) -> np.ndarray:
    """[12] float64: logit thresholds of the six classes, then their margin thresholds (refinement.py:113-114)."""
    return np.array([float(taus[c]["logit"]) for c in CLASSES] + [float(taus[c]["margin"]) for c in CLASSES], dtype=np.float64)


def refine_contigs(engine, logits: np.ndarray, offsets: np.ndarray, taus: dict, mode: str = "gated", merge_split: str = "half",
                   merge_bp: bool = True, merge_pv: bool = True) -> dict[str, np.ndarray]:
    """Device pass: per-window labels / margins and the per-contig sums.  logits [W, n_cls >= 6] float32, offsets
    [n_contigs + 1] window ranges."""
    if mode not in MODES:
        raise ValueError(f"refine mode {mode!r} (use gated, weighted or unweighted)")
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    if logits.ndim != 2 or logits.shape[1] < 6:
        raise ValueError("the refinement layer needs the six score columns " + ", ".join(SCORE_COLS))
    n_win, n_cls = logits.shape
    n = len(offsets) - 1
    with torch.cuda.stream(engine._stream()):
        d_logits = engine._h2d(logits)
        d_off = engine._h2d(np.ascontiguousarray(offsets, dtype=np.int64))
        d_tau = engine._h2d(tau_vector(taus))
        label = engine._empty((n_win,), torch.uint8)
        margin = engine._empty((n_win,), torch.float64)
        sums = engine._empty((n, 6), torch.float64)
        stats = engine._empty((n, 2), torch.int32)
        weight = engine._empty((n,), torch.float64)
        check(lib.jg_refine_contigs(engine.ctx.handle, d_logits.data_ptr(), d_off.data_ptr(), n, n_win, n_cls, d_tau.data_ptr(),
                                    int(merge_bp), int(merge_pv), MODES[mode], 0.5 if merge_split == "half" else 1.0,
                                    label.data_ptr(), margin.data_ptr(), sums.data_ptr(), stats.data_ptr(), weight.data_ptr()))
        out = {"label": label.cpu().numpy(), "margin": margin.cpu().numpy(), "sums": sums.cpu().numpy(),
               "stats": stats.cpu().numpy(), "total_weight": weight.cpu().numpy()}
    engine.ctx.sync()
    return out


def refined_contig_table(engine, headers, logits: np.ndarray, offsets: np.ndarray, taus: dict, mode: str = "gated",
                         min_windows: int = 3, merge_split: str = "half", allow_merged_contig_call: bool = False,
                         contig_hedge_margin: float = 1.0):
    """_build_refined_contig_df (commands/predict.py:115-155): one row per contig with at least `min_windows` used
    windows -- the six aggregated scores, n_windows_used, total_weight, n_merged_windows, contig_call,
    contig_top_class, contig_second_class, contig_top_logit, contig_margin (refinement.py:203-247)."""
    import pandas as pd
    r = refine_contigs(engine, logits, offsets, taus, mode, merge_split)
    keep = np.flatnonzero(r["stats"][:, 0] >= min_windows)             # refinement.py:214
    S = r["sums"][keep]
    order = np.argsort(S, axis=1, kind="stable")                        # refinement.py:217-219
    rows = np.arange(len(keep))
    top, second = (order[:, -1], order[:, -2]) if len(keep) else (np.zeros(0, int), np.zeros(0, int))
    top_val, second_val = S[rows, top], S[rows, second]
    cmargin = top_val - second_val
    top_class = [CLASSES[i] for i in top]
    second_class = [CLASSES[i] for i in second]
    call = list(top_class)
    if allow_merged_contig_call:                                        # refinement.py:228-237
        pairs = {frozenset(m): lbl for lbl, m in MERGE_MAP.items()}
        call = [pairs[frozenset((t, s))] if m < contig_hedge_margin and frozenset((t, s)) in pairs else t
                for t, s, m in zip(top_class, second_class, cmargin)]
    cols: dict[str, Any] = {"contig_id": [headers[i] for i in keep]}
    for k, name in enumerate(SCORE_COLS):
        cols[name] = S[:, k]
    cols.update({"n_windows_used": r["stats"][keep, 0].astype(np.int64), "total_weight": r["total_weight"][keep],
                 "n_merged_windows": r["stats"][keep, 1].astype(np.int64), "contig_call": call, "contig_top_class": top_class,
                 "contig_second_class": second_class, "contig_top_logit": top_val, "contig_margin": cmargin})
    df = pd.DataFrame(cols)
    df.attrs["window_labels"] = r["label"]
    return df


def merge_into_summary(df, refined):
    """generate_summary's left join of the refined calls (collect.py:534-550); contig ids still carry `___`."""
    import pandas as pd
    return pd.merge(left=df, right=refined[MERGE_COLUMNS], on="contig_id", how="left")
