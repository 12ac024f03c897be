"""Oracle (test infrastructure): forward pass of the bundled legacy `default` model.

Restates `WRes_model_embeddings(input_shape=(None,), dropout_active=False)`
(nnlib/v1/layers.py:399-423) with `ConvolutionalTower(num_res_blocks=5, add_residual=False)`
(:154-207) and `rc_resnet_block` (:90-151) in torch fp32; `tf.nn.gelu` default = exact erf GELU
(:78-79), Keras BatchNormalization epsilon 1e-3, MaxPooling1D(2) floor semantics.
PINNED on the reference's own serialized TensorFlow graph: data/models/test/jaeger_fragment_graph/saved_model.pb
interpreted op by op (oracle/tfgraph.py -> tests/golden/legacy_graph_outputs.npz); this restatement agrees with it to
3e-6 in float64 / 3e-5 in float32 on the 135 health-FASTA windows and on masked random tokens
(tests/test_legacy_graph_pin.py).  TensorFlow's own kernels (float32 rounding order) remain un-run.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _gelu(x):
    return F.gelu(x)            # exact (erf) form


def _conv_same(x, kernel, bias, dilation):
    k = kernel.shape[0]
    total = dilation * (k - 1)
    xt = F.pad(x.transpose(1, 2), (total // 2, total - total // 2))
    return F.conv1d(xt, kernel.permute(2, 1, 0).contiguous(), bias, dilation=dilation).transpose(1, 2)


def _bn(x, bn, eps=1e-3):
    return bn["gamma"] * (x - bn["mean"]) / torch.sqrt(bn["var"] + eps) + bn["beta"]


def forward(w, tokens: np.ndarray, dtype=torch.float32):
    """tokens [B, 6, L] amino-acid ids (0..21) -> {"output": [B,4], "embedding": [B,128]}."""
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)    # noqa: E731
    convs = [dict(kernel=t(c["kernel"]), bias=t(c["bias"])) for c in w["convs"]]
    bns = [{k: t(v) for k, v in b.items()} for b in w["bns"]]
    emb = t(w["embedding"])
    tok = torch.as_tensor(np.asarray(tokens).astype(np.int64))
    b, f, l = tok.shape
    x = emb[tok].reshape(b * f, l, -1)                           # shared weights over the 6 frames
    dil = [1, 2] + [3 + i for i in range(5) for _ in range(2)]
    for i in range(2):
        x = _bn(_gelu(_conv_same(x, convs[i]["kernel"], convs[i]["bias"], dil[i])), bns[i])
        x = F.max_pool1d(x.transpose(1, 2), 2).transpose(1, 2)
    for blk in range(5):
        for j in range(2):
            i = 2 + 2 * blk + j
            x = _bn(_gelu(_conv_same(x, convs[i]["kernel"], convs[i]["bias"], dil[i])), bns[i])
        x = _gelu(x)
    x = x.reshape(b, f, x.shape[1], x.shape[2]).sum(dim=1)        # Add over frames
    x = x.amax(dim=1)                                             # GlobalMaxPool1D
    d = w["dense"]
    h = _gelu(x @ t(d[0]["kernel"]) + t(d[0]["bias"]))
    g = _gelu(h @ t(d[1]["kernel"]) + t(d[1]["bias"]))
    out = g @ t(d[2]["kernel"]) + t(d[2]["bias"])
    return {"output": out.to(torch.float32).numpy(), "embedding": g.to(torch.float32).numpy()}


# ---- legacy post-processing (postprocess/collect.py:21-178, helpers.py:43-70, 495-565) -----------
def ood_predict(emb: np.ndarray, ood: dict) -> np.ndarray:
    """ood_predict_default, "sklearn" variant (helpers.py:558-565), with the bundled model written
    out: CalibratedClassifierCV(prefit, sigmoid) around LogisticRegression ->
    P(class 0) = 1 - expit(-(a * (x . coef + intercept) + b)).  Pinned against the pickled model
    itself (tests/golden/legacy_post.npz)."""
    x = (np.asarray(emb, np.float32) - ood["batch_mean"]) / ood["batch_std"]
    x = x / np.linalg.norm(x, 2, axis=1).reshape(-1, 1)
    dec = x.astype(np.float64) @ np.asarray(ood["coef"], np.float64) + float(ood["intercept"])
    return 1.0 - 1.0 / (1.0 + np.exp(float(ood["cal_a"]) * dec + float(ood["cal_b"])))


def window_summary_legacy(x, phage_pos: int) -> str:
    """get_window_summary_legacy (helpers.py:43-70)."""
    from .postprocess import find_runs
    items, run_length, _ = find_runs(np.asarray(x).flatten() == phage_pos)
    return "".join(f"{n}{'V' if it == phage_pos else 'n'}" for it, n in zip(items, run_length))


def summary_legacy(output, embedding, meta, fsize, ood, labels, all_labels, second, vindex):
    """pred_to_dict_legacy + generate_summary_legacy (collect.py:21-178), `default` model columns,
    without the terminal-repeat join.  Returns a dict of column -> list in the reference's order."""
    from .postprocess import softmax_entropy, split_points, update_dict
    output = np.asarray(output, np.float32)
    split = split_points(np.array(meta[2], dtype=np.int32), output.shape[0])
    preds = np.split(output, split, axis=0)
    oods = [ood_predict(e, ood) for e in np.split(np.asarray(embedding, np.float32), split, axis=0)]
    first = np.concatenate([[0], split]).astype(int)
    headers = np.array(meta[0], dtype=np.str_)[first]
    lengths = np.array(meta[4], dtype=np.int32)[first]
    g, c, a, t = (np.asarray(meta[i]).astype(float) for i in (-4, -5, -3, -2))
    ns = np.split((fsize - (a + t + g + c)) / fsize, split)
    gcs = np.split((g + c) / fsize, split)
    pred_sum = np.array([np.mean(p, axis=0) for p in preds], np.float16)
    pred_var = np.array([np.var(p, axis=0) for p in preds], np.float16)
    consensus = np.argmax(pred_sum, axis=1)
    frag = [np.argmax(p, axis=-1) for p in preds]
    counts = [update_dict(np.unique(f, return_counts=True), len(all_labels)) for f in frag]
    entropy = np.array([np.mean(softmax_entropy(p), axis=0) for p in preds], np.float16)
    cols = {"contig_id": [h.replace("___", ",") for h in headers], "length": lengths, "prediction": [labels[x] for x in consensus],
            "entropy": entropy, "reliability_score": [np.mean(o) for o in oods],
            "host_contam": (pred_sum[:, 1] < pred_var[:, 1]) * (consensus == 1),
            "prophage_contam": (pred_sum[:, 1] < pred_var[:, 1]) * (consensus == 0),
            "G+C": [np.mean(x) for x in gcs], "N%": [np.mean(x) for x in ns]}
    order = np.argsort(pred_sum, axis=1)[:, 2:4]
    sec = (np.prod(order == np.array([2, 1]), axis=1) + np.prod(order == np.array([3, 1]), axis=1) * 2
           + np.prod(order == np.array([0, 1]), axis=1) * 3)
    cols["prediction_2"] = [second[int(x)] for x in sec]
    for i, label in all_labels.items():
        cols[f"#_{label}_windows"] = [d[i] for d in counts]
        cols[f"{label}_score"] = [x[i] for x in pred_sum]
        cols[f"{label}_var"] = [x[i] for x in pred_var]
    cols["window_summary"] = [window_summary_legacy(f, vindex) for f in frag]
    return cols, np.concatenate(oods)
