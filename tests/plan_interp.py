"""A float64 interpreter of the fused conv-launch plan (jaeger_b200/plan.py) with the SAME
storage semantics as the CUDA kernels: every tensor is stored multiplied by its mask, residual
shortcuts are rebuilt from (stored value, mask, masked-row constant), taps and pooling see only
valid rows.  Comparing it with the un-fused oracle validates the plan compiler on the CPU."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _act(x, a):
    if a == "gelu":
        return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * (x + 0.044715 * x ** 3)))
    if a == "relu":
        return torch.relu(x)
    return x


def run_plan(plan, tokens: np.ndarray, lpad=None):
    tok = torch.as_tensor(tokens.astype(np.int64))
    b, f, lc = tok.shape
    dt = torch.float64
    bufs = {0: F.one_hot(tok, 65)[..., 1:].to(dt)}
    masks = {0: (tok != 0).to(dt)}
    taps = {}
    pooled = None
    final_mask = None
    odd_len = False                                            # parity of the frame length in front of the last stride
    for c in plan.launches:
        x = bufs[c.in_buf]                                     # stored (already masked) input
        if c.kind == 4:                                        # rows -> (even, odd) row planes: twice the channels, ceil(L / 2) rows
            l_in = x.shape[2]
            odd_len = bool(l_in & 1)
            ev, od = x[:, :, 0::2], x[:, :, 1::2]
            if od.shape[2] < ev.shape[2]:
                od = F.pad(od, (0, 0, 0, 1))
            bufs[c.out_buf] = torch.cat([ev, od], dim=-1)
            continue
        kern = c.kernel_odd if (c.kernel_odd is not None and odd_len) else c.kernel
        k = c.kernel.shape[0]
        xt = x.reshape(b * f, x.shape[2], x.shape[3]).transpose(1, 2)
        span = c.dilation * (k - 1)
        xt = F.pad(xt, (c.pad_left, span - c.pad_left - c.shrink)) if c.shrink == 0 else xt
        w = torch.as_tensor(kern, dtype=dt).permute(2, 1, 0).contiguous()
        acc = F.conv1d(xt, w, dilation=c.dilation).transpose(1, 2).reshape(b, f, -1, c.kernel.shape[2])
        m_in = masks[c.mask_in]
        if c.masking:
            mt = m_in.reshape(b * f, 1, -1)
            mt = F.pad(mt, (c.pad_left, span - c.pad_left)) if c.shrink == 0 else mt
            m_out = (F.conv1d(mt, torch.ones(1, 1, k, dtype=dt), dilation=c.dilation) >= c.mask_thr).to(dt).reshape(b, f, -1)
        else:
            m_out = torch.ones(acc.shape[:3], dtype=dt)
        masks[c.mask_out] = m_out
        mo = m_out.unsqueeze(-1)
        if c.tap_mode == 1:
            taps[c.tap_slot] = ((acc + torch.as_tensor(c.bias, dtype=dt)) * mo).sum(dim=(1, 2)) / (m_out.sum(dim=(1, 2)).unsqueeze(-1) + 1e-5) \
                - torch.as_tensor(c.tap_mean, dtype=dt)
        if c.ln1:          # MaskedLayerNormalization over the layer's real channels (padded ones carry acc = bias = gamma = beta = 0)
            u = acc + torch.as_tensor(c.bias, dtype=dt)
            n_real = c.real_cout or u.shape[-1]
            mu = u.sum(-1, keepdim=True) / n_real
            var = (u * u).sum(-1, keepdim=True) / n_real - mu * mu
            v = (u - mu) / torch.sqrt(var.clamp_min(0) + c.ln_eps) * torch.as_tensor(c.scale1, dtype=dt) + torch.as_tensor(c.shift1, dtype=dt)
        else:
            v = acc * torch.as_tensor(c.scale1, dtype=dt) + torch.as_tensor(c.shift1, dtype=dt)
        if c.dyt_g1 is not None:
            v = torch.tanh(v) * torch.as_tensor(c.dyt_g1, dtype=dt) + torch.as_tensor(c.dyt_b1, dtype=dt)
        if c.sc_buf >= 0:
            sc = bufs[c.sc_buf]
            if c.sc_mask >= 0:
                ms = masks[c.sc_mask].unsqueeze(-1)
                sc = torch.where(ms > 0, sc, torch.as_tensor(c.sc_const, dtype=dt).expand_as(sc))
            v = v + sc
        v = _act(v, c.act1)
        if c.tap_mode == 2:
            taps[c.tap_slot] = (v * mo).sum(dim=(1, 2)) / (m_out.sum(dim=(1, 2)).unsqueeze(-1) + 1e-5) \
                - torch.as_tensor(c.tap_mean, dtype=dt)
        if c.scale2 is not None:
            v = v * torch.as_tensor(c.scale2, dtype=dt) + torch.as_tensor(c.shift2, dtype=dt)
            if c.dyt_g2 is not None:
                v = torch.tanh(v) * torch.as_tensor(c.dyt_g2, dtype=dt) + torch.as_tensor(c.dyt_b2, dtype=dt)
            v = _act(v, c.act2)
        if c.tap_mode == 3:
            taps[c.tap_slot] = (v * mo).sum(dim=(1, 2)) / (m_out.sum(dim=(1, 2)).unsqueeze(-1) + 1e-5) \
                - torch.as_tensor(c.tap_mean, dtype=dt)
        if c.pool_mode == 1:
            pm = torch.where(mo > 0, v, torch.tensor(-1e9, dtype=dt)).amax(dim=(1, 2))
            pooled = torch.where(m_out.amax(dim=(1, 2)).unsqueeze(-1) > 0, pm, torch.zeros_like(pm))
        elif c.pool_mode == 2:
            pooled = (v * mo).sum(dim=(1, 2)) / torch.clamp(m_out.sum(dim=(1, 2)).unsqueeze(-1), min=1e-7)
        if c.out_buf >= 0:
            bufs[c.out_buf] = v * mo
    real = plan.real_feat_dim or plan.feat_dim
    z = pooled
    if plan.mlp is not None:                                   # hidden Dense layers, zero-padded to the feature width
        for i in range(plan.mlp_layers):
            z = _act(z @ torch.as_tensor(plan.mlp[2 * i], dtype=dt) + torch.as_tensor(plan.mlp[2 * i + 1], dtype=dt), plan.mlp_act)
    out = {"embedding": pooled[:, :real], "prediction": z @ torch.as_tensor(plan.cls_w, dtype=dt) + torch.as_tensor(plan.cls_b, dtype=dt)}
    if taps:
        out["nmd"] = torch.cat([taps[i] for i in range(len(taps))], dim=-1)
        padded_nmd = out["nmd"]
        if plan.nmd_cols is not None:
            out["nmd"] = out["nmd"][:, torch.as_tensor(plan.nmd_cols)]
        if plan.rel is not None:
            w1, b1, w2, b2 = (torch.as_tensor(a, dtype=dt) for a in plan.rel)
            rel_in = padded_nmd
            if plan.rel_signals:
                z, sig = out["prediction"], []
                pr = torch.softmax(z, dim=-1)
                top2 = torch.topk(pr, 2, dim=-1).values
                vals = {"max_prob": pr.amax(dim=-1), "entropy": -(torch.clamp(pr, min=1e-10) * torch.log(torch.clamp(pr, min=1e-10))).sum(dim=-1),
                        "energy": torch.logsumexp(z, dim=-1), "margin": top2[:, 0] - top2[:, 1], "nmd_norm": rel_in.norm(dim=-1)}
                rel_in = torch.cat([rel_in] + [vals[n].unsqueeze(-1) for n in plan.rel_signals], dim=-1)
            out["reliability"] = _act(rel_in @ w1 + b1, "gelu") @ w2 + b2
    return {k: v.numpy() for k, v in out.items()}
