"""Oracle (test infrastructure): the fragment model's forward pass on the CPU (torch).

A layer-by-layer restatement of what the reference's SavedModel graph computes, written
against the un-fused ModelSpec + raw weights so that it shares no code with the product's
plan compiler (BN folding, embedding folding, masked-row constants) or kernels:

  embedding / mask_zero ............. nnlib/builder.py:844-894
  MaskedConv1D ...................... nnlib/v2/layers.py:1217-1280 (x*mask, conv, mask "any")
  MaskedBatchNorm (inference) ....... nnlib/v2/layers.py:918-941  (no re-masking)
  activation gelu (tanh approx.) .... nnlib/v2/layers.py:27-29; Keras 3 `gelu(approximate=True)`
  ResidualBlock / Stack ............. nnlib/v2/layers.py:1882-1915, 2696-2704
  NMDLayer .......................... nnlib/v2/nmd.py:43-77
  MaskedGlobalMax/AvgPooling ........ nnlib/v2/layers.py:517-529, 460-480
  heads ............................. nnlib/builder.py:589-596, 705-713

Mask carried out of a ResidualBlock: Keras 3 `Layer._set_mask_metadata` keeps a mask that an
inner layer already attached to the output tensor, so the block output carries conv2's
(twice-dilated) mask, not the block input's mask.

Pinned per layer: every layer function below equals, to 1e-12, the reference's own `call` body executed on a NumPy stand-in
for TensorFlow (tests/golden/tf_standin.py, make_v2_layer_goldens.py -> v2_layers.npz: MaskedConv1D in all three mask modes,
MaskedBatchNorm at inference incl. return_nmd, MaskedDYT, NMDLayer, GeLU, masked max / average pooling, OODSignalLayer, and
ResidualBlockStack / ResidualBlock.call = `residual_stack` below under the stand-in's Keras 3 `__call__` mask rules; the whole
representation learner = the reference's DynamicModelBuilder._build_block run eagerly on the stand-in, v2_model.npz), and the reference
tests' mask / pooling known answers.  PARITY UNPINNED: TensorFlow / Keras themselves cannot be installed in the
build container, so the stand-in's four Keras `__call__` mask rules (the block comment above depends on them) are a reading of
Keras 3, and TF's float32 kernels are not run; the reference tests' mask / pooling known answers are also restated
(tests/unit/test_mask_mode.py, test_masked_pooling.py, test_nnlib_v2_nmd.py, test_inference_crop.py) -- see
tests/test_oracle_layer_known_answers.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def _act(x, name):
    if name == "gelu":
        return gelu_tanh(x)
    if name == "relu":
        return torch.relu(x)
    return x


def _conv1d_tf(x, kernel, dilation, padding, stride=1):
    """x [N, L, Cin], kernel [k, Cin, Cout] (TF layout) -> [N, L', Cout] with TF padding rules: SAME gives
    L' = ceil(L / stride) with max((L' - 1) * stride + dilation * (k - 1) + 1 - L, 0) zeros, the smaller half in front."""
    k = kernel.shape[0]
    xt = x.transpose(1, 2)
    if padding == "same":
        l_in = x.shape[1]
        l_out = -(-l_in // stride)
        total = max((l_out - 1) * stride + dilation * (k - 1) + 1 - l_in, 0)
        left = total // 2
        xt = F.pad(xt, (left, total - left))
    w = kernel.permute(2, 1, 0).contiguous()      # [Cout, Cin, k]
    return F.conv1d(xt, w, dilation=dilation, stride=stride).transpose(1, 2)


def masked_conv1d(x, mask, kernel, bias, dilation=1, padding="valid", activation=None, mask_mode="any", stride=1):
    """layers.py:1217-1280.  x [B,6,L,C]; mask [B,6,L] float or None."""
    b, f, l, c = x.shape
    out_mask = None
    if mask is not None:
        x = x * mask.unsqueeze(-1)
        k = kernel.shape[0]
        ones = torch.ones(k, 1, 1, dtype=x.dtype)
        mc = _conv1d_tf(mask.reshape(b * f, l, 1), ones, dilation, padding, stride)
        if mask_mode == "any":
            om = mc > 0
        elif mask_mode == "majority":
            om = mc >= (k + 1) // 2
        else:
            om = mc == float(k)
        out_mask = om.squeeze(-1).reshape(b, f, -1).to(x.dtype)
    y = _conv1d_tf(x.reshape(b * f, l, c), kernel, dilation, padding, stride)
    if bias is not None:
        y = y + bias
    y = _act(y, activation)
    return y.reshape(b, f, y.shape[1], y.shape[2]), out_mask


def batchnorm(x, bn, eps=1e-5):
    """layers.py:918-941 inference branch."""
    inv = torch.rsqrt(bn["var"] + eps)
    return bn["gamma"] * ((x - bn["mean"]) * inv) + bn["beta"]


def dyt(x, w, mask):
    """MaskedDYT (layers.py:432-444): gamma * tanh(alpha * x) + beta, multiplied by the incoming mask."""
    out = torch.tanh(w["alpha"] * x) * w["gamma"] + w["beta"]
    return out * mask.unsqueeze(-1) if mask is not None else out


def layernorm(x, w, mask, eps=1e-3):
    """MaskedLayerNormalization.call (layers.py:337-367): the input is multiplied by the mask, the moments run over the channel
    axis (population variance, tf.nn.moments), gamma / beta, and the mask once more."""
    xm = x * mask.unsqueeze(-1) if mask is not None else x
    mean = xm.mean(dim=-1, keepdim=True)
    var = ((xm - mean) ** 2).mean(dim=-1, keepdim=True)
    out = (xm - mean) / torch.sqrt(var + eps) * w["gamma"] + w["beta"]
    return out * mask.unsqueeze(-1) if mask is not None else out


def _norm(x, w, mask, eps=1e-5, ln_eps=1e-3):
    """Dispatch on the weights a norm layer owns: alpha -> MaskedDYT, moving statistics -> BatchNorm, gamma / beta only ->
    MaskedLayerNormalization."""
    if "alpha" in w:
        return dyt(x, w, mask)
    if "mean" in w:
        return batchnorm(x, w, eps)
    return layernorm(x, w, mask, ln_eps)


def nmd_vector(x, mask, moving_mean, eps=1e-5):
    """nmd.py:52-77 inference branch."""
    if mask is not None:
        m = mask.unsqueeze(-1)
        s = (x * m).sum(dim=(1, 2))
        n = m.sum(dim=(1, 2)) + eps
        mean = s / n
    else:
        mean = x.mean(dim=(1, 2))
    return mean - moving_mean


def masked_global_max(x, mask):
    """layers.py:517-529."""
    if mask is None:
        return x.amax(dim=(1, 2))
    m = mask.unsqueeze(-1)
    pooled = torch.where(m > 0, x, torch.tensor(-1.0e9, dtype=x.dtype)).amax(dim=(1, 2))
    has = m.amax(dim=(1, 2))
    return torch.where(has > 0, pooled, torch.zeros_like(pooled))


def masked_global_avg(x, mask):
    """layers.py:460-480."""
    if mask is None:
        return x.mean(dim=(1, 2))
    m = mask.unsqueeze(-1)
    s = (x * m).sum(dim=(1, 2))
    n = torch.clamp(m.sum(dim=(1, 2)), min=1e-7)
    return torch.where(n > 0, s / n, torch.zeros_like(s))


def ood_signals(logits, nmd, signals, eps=1e-10):
    """OODSignalLayer.call (layers.py:1632-1666)."""
    probs = torch.softmax(logits, dim=-1)
    out = []
    for sname in signals:
        if sname == "max_prob":
            out.append(probs.amax(dim=-1, keepdim=True))
        elif sname == "entropy":
            sp = torch.clamp(probs, min=eps)
            out.append(-(sp * torch.log(sp)).sum(dim=-1, keepdim=True))
        elif sname == "energy":
            out.append(torch.logsumexp(logits, dim=-1, keepdim=True))
        elif sname == "margin":
            top2 = torch.topk(probs, 2, dim=-1).values
            out.append(top2[..., 0:1] - top2[..., 1:2])
        elif sname == "nmd_norm":
            out.append(torch.linalg.vector_norm(nmd, dim=-1, keepdim=True))
        else:
            raise ValueError(sname)
    return torch.cat(out, dim=-1)


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def residual_stack(x, mask, blocks, c, dtype=torch.float32):
    """ResidualBlockStack / ResidualBlock.call (layers.py:2696-2704, 1882-1915) on x [B,6,L,C] carrying `mask` (or None):
    returns (output, the mask it carries, NMD vector of the last block's bn2 when c["return_nmd"] else None)."""
    block_nmd = None
    stride = int(c.get("strides", 1))                 # ResidualBlockStack hands `strides` to EVERY block (layers.py:2676-2690)
    for bi, blk in enumerate(blocks):
        m_in = mask if c["use_masking"] else None
        tw = lambda d: {k: _t(v, dtype) for k, v in d.items()}     # noqa: E731
        bias = lambda d: _t(d["bias"], dtype) if c.get("use_bias", True) else None     # noqa: E731
        h, m1 = masked_conv1d(x, m_in, _t(blk["conv1"]["kernel"], dtype), bias(blk["conv1"]), c["dilation"], "same", stride=stride)
        h = _act(_norm(h, tw(blk["bn1"]), m1 if m_in is not None else None, ln_eps=c.get("ln_epsilon", 1e-3)), c["activation"])
        h2, m2 = masked_conv1d(h, m1, _t(blk["conv2"]["kernel"], dtype), bias(blk["conv2"]), c["dilation"], "same")
        if c.get("return_nmd") and bi == len(blocks) - 1:          # layers.py:1897-1898, 2696-2704
            block_nmd = nmd_vector(h2, m2 if m_in is not None else None, _t(blk["bn2"]["mean"], dtype))
        h2 = _norm(h2, tw(blk["bn2"]), m2 if m_in is not None else None, ln_eps=c.get("ln_epsilon", 1e-3))
        shortcut = x
        if "conv3" in blk:                            # layers.py:1855-1864, 1903-1909: 1x1 conv (same stride) + norm on the block input
            sc, m3 = masked_conv1d(x, m_in, _t(blk["conv3"]["kernel"], dtype), bias(blk["conv3"]), c["dilation"], "same", stride=stride)
            shortcut = _norm(sc, tw(blk["bn3"]), m3 if m_in is not None else None, ln_eps=c.get("ln_epsilon", 1e-3))
        x = _act(h2 + shortcut, c["activation"])      # MaskedAdd: no re-masking (layers.py:60-76)
        mask = m2 if m_in is not None else mask
    return x, mask, block_nmd


def forward(spec, weights, tokens: np.ndarray, dtype=torch.float32) -> dict[str, np.ndarray]:
    """tokens [B, 6, L] uint8 (0 = unknown / padding) -> prediction, embedding, nmd, reliability."""
    tok = torch.as_tensor(np.asarray(tokens).astype(np.int64))
    emb = weights.get("embedding")
    if emb is not None:
        e = _t(emb, dtype)
        if spec.uses_token_input:
            x = e[tok]                                        # Embedding(vocab 65, mask_zero)
        else:
            onehot = F.one_hot(tok, 65)[..., 1:].to(dtype)    # all-zero row for token 0
            x = onehot @ e                                    # Dense(E, use_bias=False)
    else:
        x = F.one_hot(tok, 65)[..., 1:].to(dtype)
    mask = (tok != 0).to(dtype) if spec.use_masking else None
    nmds = []
    for layer, lw in zip(spec.layers, weights["layers"]):
        c = layer.cfg
        if layer.kind == "conv":
            m_in = mask if c["use_masking"] else None
            x, mask = masked_conv1d(x, m_in, _t(lw["kernel"], dtype),
                                    _t(lw["bias"], dtype) if c["use_bias"] else None,
                                    c["dilation"], c["padding"], c.get("activation"), mask_mode=c.get("mask_mode", "any"))
        elif layer.kind == "nmd":
            nmds.append(nmd_vector(x, mask, _t(lw["moving_mean"], dtype)))
        elif layer.kind == "norm":
            nw = {k: _t(v, dtype) for k, v in lw.items()}
            if c.get("return_nmd"):               # layers.py:943-954: NMD of the norm's input against its own moving mean
                nmds.append(nmd_vector(x, mask if spec.use_masking else None, nw["mean"], c.get("epsilon", 1e-5)))
            x = _norm(x, nw, mask if spec.use_masking else None, c.get("epsilon", 1e-5), ln_eps=c.get("epsilon", 1e-3))
        elif layer.kind == "act":
            x = _act(x, c.get("activation"))
        elif layer.kind == "resblock":
            x, mask, block_nmd = residual_stack(x, mask, lw["blocks"], c, dtype)
            if block_nmd is not None:
                nmds.append(block_nmd)
        else:
            raise NotImplementedError(layer.kind)
    feat = masked_global_max(x, mask) if spec.pooling == "max" else masked_global_avg(x, mask)
    z = feat                                          # classification head: Dense stack, dropout is the identity (builder.py:589-596)
    for d, dw in zip(spec.classifier, weights["classifier"]):
        z = _act(z @ _t(dw["kernel"], dtype) + _t(dw["bias"], dtype), d.get("activation"))
    out = {"prediction": z, "embedding": feat}
    if nmds:
        out["nmd"] = torch.cat(nmds, dim=-1)
        if spec.reliability is not None and "reliability" in weights:
            r = weights["reliability"]
            rel_in = out["nmd"]
            if getattr(spec, "reliability_signals", None):          # builder.py:716-722: concat(nmd, OODSignalLayer(logits, nmd))
                rel_in = torch.cat([rel_in, ood_signals(out["prediction"], out["nmd"], spec.reliability_signals)], dim=-1)
            h = _act(rel_in @ _t(r[0]["kernel"], dtype) + _t(r[0]["bias"], dtype), spec.reliability[0]["activation"])
            out["reliability"] = h @ _t(r[1]["kernel"], dtype) + _t(r[1]["bias"], dtype)
    return {k: v.to(torch.float32).numpy() for k, v in out.items()}
