"""Oracle (test infrastructure): windows -> contigs aggregation and the TSV summary.

NumPy restatement of
  * `pred_to_dict`            postprocess/collect.py:247-435
  * `frac_above_threshold`    postprocess/collect.py:233-244
  * `generate_summary`        postprocess/collect.py:438-558
  * `write_output`            postprocess/collect.py:561-608
  * `find_runs`, `get_window_summary`, `update_dict`, `softmax_entropy`, `logsumexp`,
    `energy`, `sigmoid`       postprocess/helpers.py:8-40, 73-127, 175-235
Pinned against the reference functions themselves (importable in the build container with a
pyfastx stub) through tests/golden/make_goldens.py.
"""
from __future__ import annotations

import numpy as np


def find_runs(x):
    """helpers.py:8-40."""
    x = np.asanyarray(x)
    n = x.shape[0]
    if n == 0:
        return np.array([], dtype=x.dtype), np.array([], dtype=int), np.array([], dtype=int)
    start = np.empty(n, dtype=bool)
    start[0] = True
    np.not_equal(x[:-1], x[1:], out=start[1:])
    run_starts = np.nonzero(start)[0]
    return x[run_starts], np.diff(np.append(run_starts, n)), run_starts


def get_window_summary(x, class_map: dict[int, str], classes=("virus", "phage")) -> str:
    """helpers.py:73-108: run-length string, upper-case initial for the viral classes."""
    initial = {k: (v[0].upper() if v.lower() in classes else v[0].lower()) for k, v in class_map.items()}
    items, lengths, _ = find_runs(np.asarray(x).flatten())
    return "".join(f"{n}{initial.get(int(i), '')}" for i, n in zip(items, lengths))


def update_dict(x, num_classes=4):
    """helpers.py:111-127."""
    return {i: 0 for i in range(num_classes)} | dict(zip(x[0], x[1]))


def softmax_entropy(p, axis=-1, eps=1e-12):
    """helpers.py:175-177 (applied by pred_to_dict to the raw logits)."""
    p = np.clip(p, eps, 1.0)
    return -np.sum(p * np.log2(p), axis=axis)


def logsumexp(x, axis=-1):
    xmax = np.max(x, axis=axis, keepdims=True)
    return xmax.squeeze(axis=axis) + np.log(np.sum(np.exp(x - xmax), axis=axis))


def energy(x, axis=-1):
    """helpers.py:189-219 -- note the n_cls != 2 case falls through to the 'binary' branch."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 0:
        return -logsumexp(np.array([x, 0.0]), axis=-1)
    if x.shape[-1] == 2:
        return -logsumexp(x, axis=axis)
    sq = x.squeeze(axis=-1) if x.shape[-1] == 1 else x
    return -logsumexp(np.stack([sq, np.zeros_like(sq)], axis=-1), axis=-1)


def sigmoid(x):
    return 1 / (1 + np.exp(-x))


def frac_above_threshold(pairs, threshold=0.5, fmt="{:.2f}", none_str="-"):
    """collect.py:233-244."""
    if pairs is None:
        return none_str
    arr = np.asarray(pairs, dtype=float)
    if arr.size == 0:
        return fmt.format(0.0)
    return fmt.format((arr > threshold).mean())


def split_points(is_last, n):
    """collect.py:260-287."""
    idx = np.where(np.asarray(is_last, dtype=np.int32) == 1)[0] + 1
    if len(idx) and n == idx[-1]:
        idx = idx[:-1]
    return idx


def aggregate_numeric(prediction, reliability, is_last):
    """The numeric core of pred_to_dict (softmax classifier): per-contig fp16 mean / var,
    consensus, per-window argmax, per-class counts, entropy / energy means, reliability."""
    prediction = np.asarray(prediction)
    sp = split_points(is_last, prediction.shape[0])
    preds = np.split(prediction, sp, axis=0)
    n_cls = prediction.shape[-1]
    pred_sum = np.array([np.squeeze(np.mean(p, axis=0)) for p in preds], dtype=np.float16)
    pred_var = np.array([np.squeeze(np.var(p, axis=0)) for p in preds], dtype=np.float16)
    consensus = np.argmax(pred_sum, axis=1)
    frag_pred = [np.argmax(p, axis=-1) for p in preds]
    counts = np.array([[update_dict(np.unique(fp, return_counts=True), n_cls)[k] for k in range(n_cls)] for fp in frag_pred])
    entropy = np.array([np.squeeze(np.mean(softmax_entropy(p))) for p in preds], dtype=np.float16)
    energy_mean = np.array([np.squeeze(np.mean(energy(p))) for p in preds], dtype=np.float16)
    ood = None
    if reliability is not None:
        ood = np.array([frac_above_threshold(sigmoid(p)) for p in np.split(np.asarray(reliability), sp, axis=0)],
                       dtype=np.float16)
    return dict(pred_sum=pred_sum, pred_var=pred_var, consensus=consensus, frag_pred=frag_pred,
                per_class_counts=counts, entropy=entropy, energy=energy_mean, ood=ood)


def pred_to_dict(y_pred: dict, fsize: int, class_map: dict, term_repeats=None):
    """collect.py:247-435 for the softmax classifier without the opt-in CRF decoding."""
    n = y_pred["prediction"].shape[0]
    sp = split_points(y_pred["meta_2"], n)
    agg = aggregate_numeric(y_pred["prediction"], y_pred.get("reliability"), y_pred["meta_2"])
    headers = np.array([h[0] for h in np.split(np.array(y_pred["meta_0"], dtype=str), sp)])
    lengths = np.array([b[0] for b in np.split(np.array(y_pred["meta_4"], dtype=np.int32), sp)])
    gc_skews = np.split(y_pred["meta_9"].astype(float), sp)
    a, t, g, c = (y_pred[k].astype(float) for k in ("meta_7", "meta_8", "meta_6", "meta_5"))
    ns = np.split((fsize - (a + t + g + c)) / fsize, sp)
    gcs = np.split((g + c) / fsize, sp)
    pred_sum, pred_var, consensus = agg["pred_sum"], agg["pred_var"], agg["consensus"]
    n_cls = class_map["num_classes"]
    data = {
        "headers": headers, "length": lengths, "consensus": consensus,
        "per_class_counts": [dict(zip(range(n_cls), row)) for row in agg["per_class_counts"]],
        "pred_sum": pred_sum, "pred_var": pred_var, "frag_pred": agg["frag_pred"], "ood": agg["ood"],
        "has_reliability": "reliability" in y_pred, "entropy": agg["entropy"], "energy": agg["energy"],
        "host_contam": (pred_sum[:, 1] < pred_var[:, 1]) & (consensus == 1),
        "prophage_contam": (pred_sum[:, 1] < pred_var[:, 1]) & (consensus == 0),
        "repeats": term_repeats, "gc": gcs, "ns": ns,
    }
    data_full = {"predictions": np.split(y_pred["prediction"], sp, axis=0), "headers": headers, "lengths": lengths,
                 "gc_skews": gc_skews, "gcs": gcs}
    return data, data_full


def generate_summary(data, labels, indices):
    """collect.py:438-558 (without the optional refinement merge)."""
    import pandas as pd
    class_map = {int(k): v for k, v in zip(indices, labels)}
    rel = data["ood"] if data.get("has_reliability", True) else ["unavailable"] * len(data["headers"])
    cols = {"contig_id": data["headers"], "length": data["length"],
            "prediction": [class_map[x] for x in data["consensus"]], "entropy": data["entropy"],
            "energy": data["energy"], "reliability_score": rel, "host_contam": data["host_contam"],
            "prophage_contam": data["prophage_contam"]}
    cols["G+C"] = [np.mean(x) for x in data["gc"]]
    cols["N%"] = [np.mean(x) for x in data["ns"]]
    if len(class_map) > 2:
        for i, label in class_map.items():
            cols[f"#_{label}_windows"] = [x[i] for x in data["per_class_counts"]]
        for i, label in class_map.items():
            cols[f"{label}_score"] = [x[i] for x in data["pred_sum"]]
            cols[f"{label}_var"] = [x[i] for x in data["pred_var"]]
    else:
        for i, label in class_map.items():
            cols[f"#_{label}_windows"] = [x[i] for x in data["per_class_counts"]]
        cols["score"] = data["pred_sum"]
        cols["var"] = data["pred_var"]
    cols["window_summary"] = [get_window_summary(x, class_map=class_map) for x in data["frag_pred"]]
    df = pd.DataFrame(cols)
    if data.get("repeats") is not None:
        df = pd.merge(left=df, right=data["repeats"][["contig_id", "terminal_repeats", "repeat_length"]],
                      on="contig_id", how="left")
    df["contig_id"] = df["contig_id"].str.replace("___", ",")
    return df


# ---- --crf window decoding (postprocess/helpers.py:291-449) ------------------------------------
CRF_PRIOR_TIERS = (
    (0.5, (("bacteria", "phage"), ("bacteria", "plasmid"), ("archaea", "phage"), ("archaea", "plasmid"),
           ("phage", "plasmid"), ("eukarya", "virus"))),
    (3.0, (("bacteria", "eukarya"), ("archaea", "eukarya"), ("bacteria", "archaea"), ("eukarya", "phage"),
           ("eukarya", "plasmid"))),
)


def build_transition_costs(class_names, switch_cost, prior="biological", user_matrix=None):
    """helpers.py:345-390."""
    names = [str(n).lower() for n in class_names]
    n = len(names)
    p = np.ones((n, n), dtype=np.float64)
    np.fill_diagonal(p, 0.0)
    if user_matrix:
        for a, row in user_matrix.items():
            a = str(a).lower()
            if a not in names or not isinstance(row, dict):
                continue
            for b, value in row.items():
                b = str(b).lower()
                if b not in names:
                    continue
                i, j = names.index(a), names.index(b)
                p[i, j] = p[j, i] = float(value)
        np.fill_diagonal(p, 0.0)
    elif prior != "uniform":
        for value, pairs in CRF_PRIOR_TIERS:
            for a, b in pairs:
                if a in names and b in names:
                    i, j = names.index(a), names.index(b)
                    p[i, j] = p[j, i] = value
    return float(switch_cost) * p


def viterbi_decode(logits, switch_cost=2.0, transition_costs=None):
    """helpers.py:393-449: float64 log-softmax emissions, delta/back-pointer recursion, first index
    wins each argmax."""
    z = np.asarray(logits, dtype=np.float64)
    if z.ndim == 1:
        z = z.reshape(1, -1)
    t_len, n_classes = z.shape
    em = z - logsumexp(z, axis=-1)[:, None]
    if t_len == 1 or n_classes == 1:
        return np.argmax(em, axis=-1)
    if transition_costs is None:
        costs = np.full((n_classes, n_classes), float(switch_cost))
        np.fill_diagonal(costs, 0.0)
    else:
        costs = np.asarray(transition_costs, dtype=np.float64)
    delta = em[0].copy()
    back = np.zeros((t_len, n_classes), dtype=np.int64)
    for t in range(1, t_len):
        scores = delta[:, None] - costs
        back[t] = np.argmax(scores, axis=0)
        delta = em[t] + scores[back[t], np.arange(n_classes)]
    path = np.empty(t_len, dtype=np.int64)
    path[-1] = int(np.argmax(delta))
    for t in range(t_len - 2, -1, -1):
        path[t] = back[t + 1][path[t + 1]]
    return path
