"""The bundled legacy `default` model (BASELINE config 1) on the B200 engine.

Graph (nnlib/v1/layers.py:399-423 `WRes_model_embeddings`, :154-207 `ConvolutionalTower`,
:90-151 `rc_resnet_block` with add_residual=False), shared over the six frames:
    Embedding(22, 4) -> Conv1D(128, k9, SAME) -> GELU(erf) -> BN(eps 1e-3) -> MaxPool(2)
    -> Conv1D(128, k5, d2) -> GELU -> BN -> MaxPool(2)
    -> 5 x { 2 x [Conv1D(128, k5, d=3+i) -> GELU -> BN] -> GELU }
    -> Add over frames -> GlobalMaxPool1D -> Dense(128, gelu) -> Dense(128, gelu) = embedding -> Dense(4)
No masking anywhere (Keras drops the Embedding mask at the first Conv1D).  The encoder is the
legacy one (preprocess/v1/convert.py:56-125): amino-acid ids, unknown / soft-masked codons -> 0.
"""
from __future__ import annotations

import re
from typing import Any

import numpy as np

from . import codon_tables
from .plan import ConvLaunch, Plan, _np32

BN_EPS = 1e-3      # keras.layers.BatchNormalization default
DILATIONS = [1, 2] + [3 + i for i in range(5) for _ in range(2)]
DEFAULT_LABELS = {0: "non-phage", 1: "phage", 2: "non-phage", 3: "non-phage"}     # data/config.json default_labels
ALL_LABELS = {0: "bacteria", 1: "phage", 2: "eukarya", 3: "archaea"}
SECOND_LABELS = {1: "eukarya", 2: "archaea", 3: "bacteria", 0: ""}                # data/config.json "second"
VINDEX = 1                                                                         # data/config.json "vindex"


def load_ood_params(model_dir) -> dict[str, Any]:
    """The reliability model of the legacy `default` run (commands/predict_legacy.py:99-109):
    LR_ood_4_class_default.pkl (a prefit sigmoid-calibrated logistic regression) + batch_means.npy
    + batch_std.npy, reduced to the arrays the device kernel needs.  scikit-learn is only used to
    unpickle; nothing of it runs on the prediction path."""
    from pathlib import Path
    d = Path(model_dir)
    if d.suffix == ".npz":        # the same arrays exported once (keys as below, optionally prefixed "ood_")
        z = np.load(d)
        get = lambda k: z[k] if k in z.files else z["ood_" + k]
        return {"coef": np.asarray(get("coef"), np.float64).ravel(), "intercept": float(get("intercept")), "cal_a": float(get("cal_a")),
                "cal_b": float(get("cal_b")), "batch_mean": np.asarray(get("batch_mean"), np.float32), "batch_std": np.asarray(get("batch_std"), np.float32)}
    import joblib
    model = joblib.load(d / "LR_ood_4_class_default.pkl")
    cc = model.calibrated_classifiers_[0]
    return {"coef": np.asarray(cc.estimator.coef_, np.float64).ravel(), "intercept": float(cc.estimator.intercept_[0]),
            "cal_a": float(cc.calibrators[0].a_), "cal_b": float(cc.calibrators[0].b_),
            "batch_mean": np.load(d / "batch_means.npy").astype(np.float32), "batch_std": np.load(d / "batch_std.npy").astype(np.float32)}


def weights_from_bundle(tensors: dict[str, np.ndarray]) -> dict[str, Any]:
    """Map the SavedModel bundle of the bundled graph (`_operations/<i>/...`, application order)
    to named weights.  The same bytes live in data/models/default/WRes_1024.h5."""
    ops: dict[int, dict[str, np.ndarray]] = {}
    for k, v in tensors.items():
        m = re.match(r"_operations/(\d+)/(.*?)/\.ATTRIBUTES", k)
        if m:
            ops.setdefault(int(m.group(1)), {})[m.group(2)] = v
    emb, convs, bns, dense = None, [], [], []
    for i in sorted(ops):
        o = ops[i]
        if "_embeddings" in o:
            emb = o["_embeddings"]
        elif "gamma" in o:
            bns.append(dict(gamma=o["gamma"], beta=o["beta"], mean=o["moving_mean"], var=o["moving_variance"]))
        elif "_kernel" in o and o["_kernel"].ndim == 3:
            convs.append(dict(kernel=o["_kernel"], bias=o["bias"]))
        elif "_kernel" in o:
            dense.append(dict(kernel=o["_kernel"], bias=o["bias"]))
    if emb is None or len(convs) != 12 or len(bns) != 12 or len(dense) != 3:
        raise ValueError("not the legacy WRes_model_embeddings bundle")
    return dict(embedding=emb, convs=convs, bns=bns, dense=dense)


def random_weights(seed: int = 0) -> dict[str, Any]:
    rng = np.random.default_rng(seed)

    def conv(k, cin):
        lim = np.sqrt(6.0 / (k * cin))
        return dict(kernel=rng.uniform(-lim, lim, (k, cin, 128)).astype(np.float32), bias=rng.normal(0, 0.05, 128).astype(np.float32))

    def bn():
        return dict(gamma=rng.uniform(0.8, 1.2, 128).astype(np.float32), beta=rng.normal(0, 0.1, 128).astype(np.float32),
                    mean=rng.normal(0, 0.1, 128).astype(np.float32), var=rng.uniform(0.5, 1.5, 128).astype(np.float32))

    def dense(i, o):
        lim = np.sqrt(6.0 / (i + o))
        return dict(kernel=rng.uniform(-lim, lim, (i, o)).astype(np.float32), bias=rng.normal(0, 0.05, o).astype(np.float32))

    return dict(embedding=rng.normal(0, 0.5, (22, 4)).astype(np.float32), convs=[conv(9, 4)] + [conv(5, 128) for _ in range(11)],
                bns=[bn() for _ in range(12)], dense=[dense(128, 128), dense(128, 128), dense(128, 4)])


def compile_legacy_plan(w: dict[str, Any]) -> Plan:
    launches: list[ConvLaunch] = []
    n_masks = 1

    def new_mask():
        nonlocal n_masks
        n_masks += 1
        return n_masks - 1

    def bn_fold(bn):
        s = bn["gamma"].astype(np.float64) / np.sqrt(bn["var"].astype(np.float64) + BN_EPS)
        return _np32(s), _np32(bn["beta"].astype(np.float64) - s * bn["mean"].astype(np.float64))

    def conv(i, in_buf, out_buf, mask_in, halvings, outer_gelu):
        k = w["convs"][i]["kernel"]
        if i == 0:   # fold Embedding(22, 4): one-hot channel = amino-acid id (0..21), zero rows = SAME padding
            table = np.zeros((64, k.shape[1]))
            table[:22] = w["embedding"]
            k = np.einsum("ve,keo->kvo", table, k.astype(np.float64))
        d = DILATIONS[i]
        span = d * (k.shape[0] - 1)
        s2, t2 = bn_fold(w["bns"][i])
        c = ConvLaunch(kernel=_np32(k), bias=_np32(w["convs"][i]["bias"]), dilation=d, pad_left=span // 2, shrink=0,
                       in_buf=in_buf, out_buf=out_buf, mask_in=mask_in, mask_out=new_mask(), masking=0,
                       scale1=_np32(np.ones(128)), shift1=_np32(w["convs"][i]["bias"]), act1="gelu_erf",
                       scale2=s2, shift2=t2, act2="gelu_erf" if outer_gelu else None, halvings=halvings,
                       epi_f32=1)      # BatchNorm FOLLOWS the GELU here with gamma / sigma up to 26 (first layer): everything up to
                                       # the store stays in fp32, a half-precision GELU output would be amplified by that scale
        launches.append(c)
        return c

    def pool(in_buf, out_buf, mask_in, halvings):
        c = ConvLaunch(kernel=np.zeros((1, 128, 128), np.float32), bias=np.zeros(128, np.float32), dilation=1, pad_left=0,
                       shrink=0, in_buf=in_buf, out_buf=out_buf, mask_in=mask_in, mask_out=new_mask(), masking=0,
                       kind=2, halvings=halvings)
        launches.append(c)
        return c

    c = conv(0, 0, 1, 0, 0, False)
    p = pool(1, 2, c.mask_out, 0)
    c = conv(1, 2, 1, p.mask_out, 1, False)
    p = pool(1, 2, c.mask_out, 1)
    buf, mask = 2, p.mask_out
    for i in range(2, 12):
        out = 1 if buf == 2 else 2
        c = conv(i, buf, out, mask, 2, outer_gelu=(i % 2 == 1))      # second conv of a block carries the block's GELU
        buf, mask = out, c.mask_out
    launches.append(ConvLaunch(kernel=np.zeros((1, 128, 128), np.float32), bias=np.zeros(128, np.float32), dilation=1,
                               pad_left=0, shrink=0, in_buf=buf, out_buf=-1, mask_in=mask, mask_out=mask, masking=0,
                               kind=3, halvings=2))
    d = w["dense"]
    return Plan(launches=launches, n_classes=4, feat_dim=128, pool_mode=1, n_taps=0, tap_width=0,
                cls_w=_np32(d[2]["kernel"]), cls_b=_np32(d[2]["bias"]), rel=None, rel_hidden=0, total_shrink=0,
                tok_offset=0, mlp=[_np32(d[0]["kernel"]), _np32(d[0]["bias"]), _np32(d[1]["kernel"]), _np32(d[1]["bias"])],
                mlp_act="gelu_erf")


LEGACY_LUT = codon_tables.device_lut(codon_tables.LEGACY_AA_ID, plus_one=False)
