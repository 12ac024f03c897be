"""Windows -> contigs summary and TSV output on top of the stage-4 device kernels.

Product-side counterpart of the reference's `pred_to_dict` / `generate_summary` /
`write_output` (postprocess/collect.py:247-608): the per-contig reductions run on the GPU
(`jg_aggregate_contigs`), only string assembly and the DataFrame live on the host.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any

import numpy as np
import torch


def _split_points(is_last: np.ndarray) -> np.ndarray:
    """Window offsets of the contigs: [0, end_0, end_1, ...] (collect.py:260-287)."""
    ends = np.flatnonzero(np.asarray(is_last).astype(np.int32) == 1) + 1
    n = len(is_last)
    if len(ends) == 0 or ends[-1] != n:
        ends = np.append(ends, n)           # trailing windows without an is_last flag form a contig
    return np.concatenate([[0], ends]).astype(np.int64)


def window_summaries(frag_pred: np.ndarray, offsets: np.ndarray, class_map: dict[int, str],
                     classes=("virus", "phage")) -> list[str]:
    """helpers.py:73-108 for every contig: "<run length><class initial>" runs, upper-case
    initial for the viral classes."""
    initial = np.array([(class_map[k][0].upper() if class_map[k].lower() in classes else class_map[k][0].lower())
                        for k in sorted(class_map)])
    keys = np.array(sorted(class_map))
    fp = np.asarray(frag_pred)
    n = len(fp)
    if n == 0:
        return []
    start = np.ones(n, dtype=bool)
    start[1:] = fp[1:] != fp[:-1]
    start[offsets[:-1]] = True
    run_starts = np.flatnonzero(start)
    run_len = np.diff(np.append(run_starts, n))
    run_val = fp[run_starts]
    lut = {int(k): str(c) for k, c in zip(keys, initial)}
    pieces = [f"{ln}{lut.get(int(v), '')}" for ln, v in zip(run_len, run_val)]
    owner = np.searchsorted(offsets, run_starts, side="right") - 1
    out = [""] * (len(offsets) - 1)
    bounds = np.flatnonzero(np.diff(owner, prepend=-1))
    for i, b in enumerate(bounds):
        e = bounds[i + 1] if i + 1 < len(bounds) else len(pieces)
        out[owner[b]] = "".join(pieces[b:e])
    return out


# helpers.py:291-315: co-occurrence tiers of the --crf transition prior (lower-cased class names)
_CRF_PRIOR_TIERS = (
    (0.5, (("bacteria", "phage"), ("bacteria", "plasmid"), ("archaea", "phage"), ("archaea", "plasmid"),
           ("phage", "plasmid"), ("eukarya", "virus"))),
    (3.0, (("bacteria", "eukarya"), ("archaea", "eukarya"), ("bacteria", "archaea"), ("eukarya", "phage"),
           ("eukarya", "plasmid"))),
)


def build_transition_costs(class_names, switch_cost: float, prior: str = "biological", user_matrix: dict | None = None) -> np.ndarray:
    """lambda * P of helpers.py:345-390 (P: zero diagonal, 1.0 neutral, tiers or user entries symmetric)."""
    names = [str(n).lower() for n in class_names]
    p = np.ones((len(names), len(names)), dtype=np.float64)
    np.fill_diagonal(p, 0.0)
    if user_matrix:
        for a, row in user_matrix.items():
            a = str(a).lower()
            if a not in names or not isinstance(row, dict):
                continue
            for b, value in row.items():
                b = str(b).lower()
                if b in names:
                    p[names.index(a), names.index(b)] = p[names.index(b), names.index(a)] = float(value)
        np.fill_diagonal(p, 0.0)
    elif prior != "uniform":
        for value, pairs in _CRF_PRIOR_TIERS:
            for a, b in pairs:
                if a in names and b in names:
                    p[names.index(a), names.index(b)] = p[names.index(b), names.index(a)] = value
    return float(switch_cost) * p


def _window_numbers(y_pred) -> dict[str, np.ndarray]:
    """Per-window numbers the tables need.  Taken from the engine's numeric window table when the
    result carries one (no round trip through the meta byte strings), else parsed from meta_0..9."""
    t = getattr(y_pred, "window_table", None)
    if t is not None and t.contig is not None:
        names = np.empty(len(t.headers), dtype=object)      # object array: no copy of a million header strings into a fixed-width buffer
        names[:] = t.headers
        skew = np.where(t.skew100 == (1 << 14), 0, t.skew100).astype(float) / 100.0
        return {"is_last": t.is_last, "headers": names[t.contig], "seqlen": t.seqlen.astype(np.int32),
                "g": t.counts[:, 0].astype(float), "c": t.counts[:, 1].astype(float), "a": t.counts[:, 2].astype(float),
                "t": t.counts[:, 3].astype(float), "gc_skew": skew}
    return {"is_last": np.asarray(y_pred["meta_2"]), "headers": np.array(y_pred["meta_0"], dtype=str),
            "seqlen": np.array(y_pred["meta_4"], dtype=np.int32),
            **{k: np.asarray(y_pred[m]).astype(float) for k, m in (("g", "meta_5"), ("c", "meta_6"), ("a", "meta_7"), ("t", "meta_8"))},
            "gc_skew": np.asarray(y_pred["meta_9"]).astype(float)}


def contig_table(engine, y_pred: dict[str, np.ndarray], fsize: int, term_repeats=None, crf_switch_cost: float | None = None,
                 crf_prior: str = "biological", crf_transition_matrix: dict | None = None) -> dict[str, Any]:
    """The `data` dict of the reference's pred_to_dict (softmax classifier), with the numeric
    columns produced by the device aggregation kernels.  With crf_switch_cost the per-window labels
    and class counts come from the device Viterbi decoder (collect.py:269-287, 343-346)."""
    pred = np.ascontiguousarray(y_pred["prediction"], dtype=np.float32)
    n_cls = pred.shape[1]
    if n_cls < 2:
        raise NotImplementedError("binary (single-logit) classifiers are not on the supported path")
    wn = _window_numbers(y_pred)
    offsets = _split_points(wn["is_last"])
    rel = y_pred.get("reliability")
    on_dev = getattr(y_pred, "device_outputs", None) or {}      # the engine's own result: logits are still in HBM
    with torch.cuda.stream(engine._stream()):
        off_dev = engine._h2d(offsets)
        pred_dev = on_dev["prediction"] if "prediction" in on_dev else engine._h2d(pred)
        rel_dev = None
        if rel is not None:
            rel_dev = on_dev["reliability"] if "reliability" in on_dev else engine._h2d(np.ascontiguousarray(rel, np.float32))
        agg = engine.aggregate(pred_dev, rel_dev, off_dev)
        if crf_switch_cost is not None:
            cm = engine.class_map
            names = [n for _, n in sorted(zip(cm["index"], cm["class"]), key=lambda t: int(t[0]))]
            costs = build_transition_costs(names, crf_switch_cost, crf_prior, crf_transition_matrix)
            agg["frag_pred"], agg["per_class_counts"] = engine.viterbi(pred_dev, off_dev, costs)
        agg = {k: v.cpu().numpy() for k, v in agg.items()}
    engine.ctx.sync()
    first = offsets[:-1]
    n_win = np.diff(offsets)
    headers = wn["headers"][first]
    lengths = wn["seqlen"][first]
    g, c, a, t = wn["g"], wn["c"], wn["a"], wn["t"]
    ns = (fsize - (a + t + g + c)) / fsize                       # collect.py:319-324
    gcs = (g + c) / fsize
    pred_sum, pred_var, consensus = agg["pred_sum"], agg["pred_var"], agg["consensus"].astype(np.int64)
    ood = None
    if rel is not None:
        ood = np.array([f"{k / n:.2f}" for k, n in zip(agg["rel_pos"], n_win)], dtype=np.float16)   # collect.py:233-244, 391-395
    return {
        "headers": headers, "length": lengths, "consensus": consensus,
        "per_class_counts": agg["per_class_counts"], "pred_sum": pred_sum, "pred_var": pred_var,
        "frag_pred": agg["frag_pred"], "offsets": offsets, "ood": ood, "has_reliability": rel is not None,
        "entropy": agg["entropy"], "energy": agg["energy"],
        "host_contam": (pred_sum[:, 1] < pred_var[:, 1]) & (consensus == 1),       # collect.py:357-358
        "prophage_contam": (pred_sum[:, 1] < pred_var[:, 1]) & (consensus == 0),
        "repeats": term_repeats,
        "gc_mean": np.add.reduceat(gcs, first) / n_win, "ns_mean": np.add.reduceat(ns, first) / n_win,
        "predictions": pred, "gc_skews": wn["gc_skew"], "gcs": gcs,
        "_pred_dev": pred_dev,                     # the same logits, still in HBM (stage 4b reads them there)
    }


def generate_summary(data: dict[str, Any], labels, indices, refined_contig=None):
    """Column order of collect.py:472-518 (+ left join of the terminal-repeat columns, 527-532, and of the refined
    contig calls of `--refine`, 534-550)."""
    import pandas as pd
    class_map = {int(k): v for k, v in zip(indices, labels)}
    n = len(data["headers"])
    rel = data["ood"] if data.get("has_reliability", True) else ["unavailable"] * n
    cols = {"contig_id": data["headers"], "length": data["length"],
            "prediction": [class_map[int(x)] for x in data["consensus"]], "entropy": data["entropy"],
            "energy": data["energy"], "reliability_score": rel, "host_contam": data["host_contam"],
            "prophage_contam": data["prophage_contam"], "G+C": data["gc_mean"], "N%": data["ns_mean"]}
    if len(class_map) > 2:
        for i, label in class_map.items():
            cols[f"#_{label}_windows"] = data["per_class_counts"][:, i]
        for i, label in class_map.items():
            cols[f"{label}_score"] = data["pred_sum"][:, i]
            cols[f"{label}_var"] = data["pred_var"][:, i]
    else:
        for i, label in class_map.items():
            cols[f"#_{label}_windows"] = data["per_class_counts"][:, i]
        cols["score"] = list(data["pred_sum"])
        cols["var"] = list(data["pred_var"])
    cols["window_summary"] = window_summaries(data["frag_pred"], data["offsets"], class_map)
    df = pd.DataFrame(cols)
    if data.get("repeats") is not None:
        df = pd.merge(left=df, right=data["repeats"][["contig_id", "terminal_repeats", "repeat_length"]],
                      on="contig_id", how="left")
    if refined_contig is not None:
        from .refine import merge_into_summary
        df = merge_into_summary(df, refined_contig)
    df["contig_id"] = df["contig_id"].str.replace("___", ",")
    return df


def write_output(data: dict[str, Any], labels, indices, output_table_path: str | Path,
                 output_phage_table_path: str | Path, reliability_cutoff: float = 0.5, phage_score: float = 1) -> int:
    """collect.py:561-608: `N% < 0.3` filter, %.3f TSV, and the phage subset."""
    return write_tables(generate_summary(data, labels, indices), labels, data.get("has_reliability", True), output_table_path,
                        output_phage_table_path, reliability_cutoff, phage_score)


def write_tables(df, labels, has_reliability: bool, output_table_path: str | Path, output_phage_table_path: str | Path,
                 reliability_cutoff: float = 0.5, phage_score: float = 1) -> int:
    """The file-writing half of write_output, on an assembled (possibly multi-rank) summary table."""
    df = df.query("`N%` < 0.3")
    df.to_csv(output_table_path, sep="\t", index=False, float_format="%.3f")
    lower = [str(x).lower() for x in labels]
    viral = "phage"
    if "phage" in lower:
        viral = labels[lower.index("phage")]
    elif "virus" in lower:
        viral = labels[lower.index("virus")]
    clause = f" and (reliability_score > {reliability_cutoff})" if has_reliability else ""
    phage_df = df.query(f'(prediction == "{viral}") and ({viral}_score > {phage_score}){clause}')
    if not phage_df.empty:
        phage_df.to_csv(output_phage_table_path, sep="\t", index=False, float_format="%.3f")
    return len(df)


# ---- legacy `default` model tables (collect.py:21-229) --------------------------------------------
def contig_table_legacy(engine, y_pred: dict[str, np.ndarray], fsize: int, ood_params: dict | None, term_repeats=None) -> dict[str, Any]:
    """pred_to_dict_legacy (collect.py:21-96): the modern per-contig reductions plus the
    embedding-based reliability of every window and its per-contig mean, all on the device."""
    pred = np.ascontiguousarray(y_pred["prediction"], dtype=np.float32)
    wn = _window_numbers(y_pred)
    offsets = _split_points(wn["is_last"])
    with torch.cuda.stream(engine._stream()):
        pred_dev, off_dev = engine._h2d(pred), engine._h2d(offsets)
        agg = engine.aggregate(pred_dev, None, off_dev)
        rel = None
        if ood_params is not None:
            p0, cmean = engine.legacy_reliability(engine._h2d(np.ascontiguousarray(y_pred["embedding"], np.float32)), off_dev, ood_params)
            rel = (p0.cpu().numpy(), cmean.cpu().numpy())
        agg = {k: v.cpu().numpy() for k, v in agg.items()}
    engine.ctx.sync()
    first, n_win = offsets[:-1], np.diff(offsets)
    g, c, a, t = wn["g"], wn["c"], wn["a"], wn["t"]
    ns = (fsize - (a + t + g + c)) / fsize
    gcs = (g + c) / fsize
    pred_sum, pred_var, consensus = agg["pred_sum"], agg["pred_var"], agg["consensus"].astype(np.int64)
    return {
        "headers": wn["headers"][first], "length": wn["seqlen"][first],
        "consensus": consensus, "per_class_counts": agg["per_class_counts"], "pred_sum": pred_sum, "pred_var": pred_var,
        "frag_pred": agg["frag_pred"], "offsets": offsets, "ood_windows": rel[0] if rel else None,
        "reliability_score": rel[1] if rel else None, "has_reliability": rel is not None, "entropy": agg["entropy"],
        "host_contam": (pred_sum[:, 1] < pred_var[:, 1]) & (consensus == 1),
        "prophage_contam": (pred_sum[:, 1] < pred_var[:, 1]) & (consensus == 0),
        "repeats": term_repeats, "gc_mean": np.add.reduceat(gcs, first) / n_win, "ns_mean": np.add.reduceat(ns, first) / n_win,
        "predictions": pred, "gc_skews": wn["gc_skew"], "gcs": gcs,
    }


def window_summaries_legacy(frag_pred: np.ndarray, offsets: np.ndarray, phage_pos: int) -> list[str]:
    """get_window_summary_legacy (helpers.py:43-70): runs of phage / non-phage windows as "<n>V" / "<n>n"."""
    return window_summaries((np.asarray(frag_pred) == phage_pos).astype(np.int64), offsets, {0: "n", 1: "V"}, classes=("v",))


def generate_summary_legacy(data: dict[str, Any], labels: list[str], model: str = "default"):
    """generate_summary_legacy (collect.py:99-178) for the bundled `default` model; `labels` is the
    prediction label of each class index (default_labels, or all_labels with --getalllabels)."""
    import pandas as pd
    from . import legacy
    n = len(data["headers"])
    cols: dict[str, Any] = {
        "contig_id": data["headers"], "length": data["length"], "prediction": [labels[x] for x in data["consensus"]],
        "entropy": data["entropy"],
        "reliability_score": data["reliability_score"] if data.get("has_reliability", True) else ["unavailable"] * n,
        "host_contam": data["host_contam"], "prophage_contam": data["prophage_contam"],
    }
    if model == "default":
        cols["G+C"] = data["gc_mean"]
        cols["N%"] = data["ns_mean"]
        order = np.argsort(data["pred_sum"], axis=1)[:, 2:4]                       # collect.py:139-155
        second = (np.prod(order == np.array([2, 1]), axis=1) + 2 * np.prod(order == np.array([3, 1]), axis=1)
                  + 3 * np.prod(order == np.array([0, 1]), axis=1))
        cols["prediction_2"] = [legacy.SECOND_LABELS[int(x)] for x in second]
    for i, label in legacy.ALL_LABELS.items():
        cols[f"#_{label}_windows"] = data["per_class_counts"][:, i]
        cols[f"{label}_score"] = data["pred_sum"][:, i]
        cols[f"{label}_var"] = data["pred_var"][:, i]
    cols["window_summary"] = window_summaries_legacy(data["frag_pred"], data["offsets"], legacy.VINDEX)
    df = pd.DataFrame(cols).set_index("contig_id")
    rep = data.get("repeats")
    if rep is None:
        rep = pd.DataFrame({"contig_id": data["headers"], "terminal_repeats": [None] * n, "repeat_length": [None] * n})
    df = df.join(rep.set_index("contig_id")[["terminal_repeats", "repeat_length"]], how="left").reset_index(names="contig_id")
    df["contig_id"] = df["contig_id"].str.replace("___", ",")
    return df


def write_output_legacy(data: dict[str, Any], labels: list[str], output_table_path, output_phage_table_path,
                        reliability_cutoff: float = 0.5, phage_score: float = 3, model: str = "default") -> int:
    """write_output_legacy (collect.py:181-229)."""
    df = generate_summary_legacy(data, labels, model)
    df.to_csv(output_table_path, sep="\t", index=False, float_format="%.3f")
    clause = f" and (reliability_score > {reliability_cutoff})" if data.get("has_reliability", True) else ""
    df.query(f'(prediction == "phage") and (phage_score > {phage_score}){clause}').to_csv(
        output_phage_table_path, sep="\t", index=False, float_format="%.3f")
    return len(df)


def write_fasta_from_results(loaded, output_tsv, output_fasta, width: int = 70, append: bool = False) -> int:
    """write_fasta_from_results (collect.py:611-639) from the records already in memory (`loaded` =
    WindowSource.load(): names, bases, offsets) instead of a further pass over the input file:
    the records whose name is in the phage table, 70 letters per line.  Returns the record count."""
    import pandas as pd
    try:
        phages = set(pd.read_table(str(output_tsv))["contig_id"].astype(str).to_list())
    except (FileNotFoundError, pd.errors.EmptyDataError):
        phages = set()
    names, host, offsets = loaded
    data = host.numpy()
    n = 0
    with open(str(output_fasta), "ab" if append else "wb") as fh:
        for i, name in enumerate(names):
            if name not in phages:
                continue
            n += 1
            fh.write(b">" + name.encode() + b"\n")
            seq = data[offsets[i]:offsets[i + 1]]
            full = len(seq) // width * width
            if full:
                lines = np.concatenate([seq[:full].reshape(-1, width), np.full((full // width, 1), 10, np.uint8)], axis=1)
                fh.write(lines.tobytes())
            if len(seq) > full:
                fh.write(seq[full:].tobytes() + b"\n")
    return n
