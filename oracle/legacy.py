"""Oracle (test infrastructure): forward pass of the bundled legacy `default` model.

Restates `WRes_model_embeddings(input_shape=(None,), dropout_active=False)`
(nnlib/v1/layers.py:399-423) with `ConvolutionalTower(num_res_blocks=5, add_residual=False)`
(:154-207) and `rc_resnet_block` (:90-151) in torch fp32; `tf.nn.gelu` default = exact erf GELU
(:78-79), Keras BatchNormalization epsilon 1e-3, MaxPooling1D(2) floor semantics.
PARITY UNPINNED against TensorFlow itself (not installable here); weights are the reference's.
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
