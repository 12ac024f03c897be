"""ModelSpec + weights -> the fused conv-launch plan the CUDA library executes.

Each `masked_conv1d` / residual-block convolution becomes ONE kernel launch whose epilogue
absorbs the element-wise layers that follow it in the reference graph:

    conv(+bias) -> [NMD tap on the raw output] -> norm (folded to scale/shift) -> [+ shortcut]
    -> activation -> [NMD tap] -> [stand-alone norm -> activation] -> [global pool]

(reference graph: nnlib/builder.py:982-1193; layer math: nnlib/v2/layers.py:918-941 inference
BatchNorm, 1882-1915 ResidualBlock.call, 1217-1280 MaskedConv1D, nnlib/v2/nmd.py:43-77.)

Mask bookkeeping.  A tensor is stored multiplied by the Keras mask it carries; the reference
does NOT re-mask after BatchNorm / activation / add, so where a stored tensor is later used
outside a convolution (residual shortcut) the value at a masked row is reconstructed: with
mask_mode "any" a masked output row has seen only zeroed inputs, hence equals a per-channel
constant (`sc_const`) computed here by pushing a zero accumulator through the same epilogue.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from .modelspec import ModelSpec

ACT = {None: 0, "gelu": 1, "relu": 2, "gelu_erf": 3}
LAYER_INT_FIELDS = 32
LAYER_PTR_FIELDS = 16
# int field indices (mirror enum LayerField in csrc/jaeger_b200.cu)
(LF_KIND, LF_CIN, LF_COUT, LF_K, LF_DIL, LF_PAD_LEFT, LF_SHRINK, LF_IN_BUF, LF_OUT_BUF, LF_SC_BUF,
 LF_ACT1, LF_HAS_AFF2, LF_ACT2, LF_TAP_MODE, LF_TAP_SLOT, LF_POOL_MODE, LF_MASK_IN, LF_MASK_OUT,
 LF_SC_MASK, LF_MASKING, LF_CUM_SHRINK_IN, LF_HALVINGS, LF_DYT1, LF_DYT2, LF_EPI_F32, LF_LEN_CEIL, LF_REAL_CIN,
 LF_REAL_COUT, LF_LN1, LF_LN_EPS, LF_MASK_THR) = range(31)
(LP_KERNEL, LP_BIAS, LP_SCALE1, LP_SHIFT1, LP_SCALE2, LP_SHIFT2, LP_SC_CONST, LP_TAP_MEAN,
 LP_DYT_G1, LP_DYT_B1, LP_DYT_G2, LP_DYT_B2, LP_KERNEL_ODD) = range(13)


class LayerDesc(ctypes.Structure):
    _fields_ = [("i", ctypes.c_int32 * LAYER_INT_FIELDS),
                ("p", ctypes.POINTER(ctypes.c_float) * LAYER_PTR_FIELDS)]


class HeadDesc(ctypes.Structure):
    _fields_ = [("n_classes", ctypes.c_int32), ("feat_dim", ctypes.c_int32), ("pool_mode", ctypes.c_int32),
                ("n_taps", ctypes.c_int32), ("rel_hidden", ctypes.c_int32), ("mlp_hidden", ctypes.c_int32),
                ("mlp_act", ctypes.c_int32), ("reserved", ctypes.c_int32 * 1),
                ("cls_w", ctypes.POINTER(ctypes.c_float)), ("cls_b", ctypes.POINTER(ctypes.c_float)),
                ("rel_w1", ctypes.POINTER(ctypes.c_float)), ("rel_b1", ctypes.POINTER(ctypes.c_float)),
                ("rel_w2", ctypes.POINTER(ctypes.c_float)), ("rel_b2", ctypes.POINTER(ctypes.c_float)),
                ("mlp_w1", ctypes.POINTER(ctypes.c_float)), ("mlp_b1", ctypes.POINTER(ctypes.c_float)),
                ("mlp_w2", ctypes.POINTER(ctypes.c_float)), ("mlp_b2", ctypes.POINTER(ctypes.c_float))]


@dataclass
class ConvLaunch:
    """One fused conv launch (host-side view, also used by tests to read the plan)."""
    kernel: np.ndarray                 # [k, Cin, Cout] fp32 (stem: embedding already folded in)
    bias: np.ndarray
    dilation: int
    pad_left: int
    shrink: int                        # L_out = L_in - shrink
    in_buf: int
    out_buf: int
    sc_buf: int = -1
    scale1: np.ndarray | None = None
    shift1: np.ndarray | None = None
    act1: str | None = None
    scale2: np.ndarray | None = None
    shift2: np.ndarray | None = None
    act2: str | None = None
    tap_mode: int = 0
    tap_slot: int = 0
    tap_mean: np.ndarray | None = None
    pool_mode: int = 0
    mask_in: int = 0
    mask_out: int = 0
    sc_mask: int = -1
    masking: int = 1
    cum_shrink_in: int = 0
    sc_const: np.ndarray | None = None
    # MaskedDYT in place of either norm (layers.py:385-444): y = gamma * tanh(alpha * x) + beta, where alpha
    # is carried by scale1 / scale2 (shift 0) and the per-channel gamma / beta below follow the tanh
    dyt_g1: np.ndarray | None = None
    dyt_b1: np.ndarray | None = None
    dyt_g2: np.ndarray | None = None
    dyt_b2: np.ndarray | None = None
    out_const: np.ndarray | None = None   # value of this launch's output at rows its mask zeroed
    stage: int = 0                        # epilogue fill state while compiling
    kind: int = 1                         # 1 conv, 2 maxpool(2) per frame, 3 frame-sum + global max pool, 4 rows -> (even, odd) planes
    kernel_odd: np.ndarray | None = None  # strided convs: the kernel to use when the layer's INPUT frame length is odd (TF SAME
                                          # padding of a stride-2 conv depends on the parity of the length); `kernel` = even
    epi_f32: int = 0                      # epilogue in fp32 with one rounding at the store (norm after the activation, large scales)
    len_ceil: int = 0                     # the frame length is halved with ceil (SAME stride-2 convs), not floor (MaxPool(2))
    halvings: int = 0                     # MaxPool(2) stages applied to the frame length before this layer
    real_cin: int = 0                     # channel counts of the reference layer (before padding to multiples of 64)
    real_cout: int = 0
    mask_thr: int = 1                     # valid taps an output row needs (layers.py:1245-1252): 1 = "any", (k + 1) // 2 = "majority", k = "strict"
    ln1: int = 0                          # the first norm is a MaskedLayerNormalization: scale1 = gamma, shift1 = beta, bias kept apart
    ln_eps: float = 1e-3


@dataclass
class Plan:
    launches: list[ConvLaunch]
    n_classes: int
    feat_dim: int
    pool_mode: int
    n_taps: int
    tap_width: int
    cls_w: np.ndarray
    cls_b: np.ndarray
    rel: list[np.ndarray] | None
    rel_hidden: int
    total_shrink: int
    real_feat_dim: int | None = None           # feature width before padding to a multiple of 64
    tok_offset: int = 1
    mlp: list[np.ndarray] | None = None        # hidden Dense layers in front of the classifier: [w1, b1(, w2, b2)], width = feat_dim
    mlp_act: str | None = None
    mlp_layers: int = 2                        # how many of them (the legacy head has two)
    emb_pooled: bool = False                   # "embedding" output = the pooled features (v2 models); legacy: the last hidden layer
    rel_signals: list[str] | None = None       # OOD signals appended to the NMD vector (nmd_plus_signals)
    nmd_cols: np.ndarray | None = None         # columns of the padded device NMD vector that are the reference's NMD vector
    flops_per_window_formula: Any = None
    keep: list[Any] = field(default_factory=list)   # keeps ctypes-referenced arrays alive

    def flops_per_window(self, lc: int, frames: int = 6, algorithmic_stem_cin: int | None = None) -> float:
        """2 * frames * L_out * k * Cin * Cout summed over the convolutions.
        `algorithmic_stem_cin` replaces the folded stem's Cin = 64 by the embedding width the
        reference multiplies with."""
        f = 0.0
        for i, c in enumerate(self.launches):
            if c.kind != 1:
                continue
            k, cin, cout = c.kernel.shape
            if i == 0 and algorithmic_stem_cin:
                cin = algorithmic_stem_cin
            f += 2.0 * frames * (((lc - c.cum_shrink_in) >> c.halvings) - c.shrink) * k * cin * cout
        return f


def _np32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _gelu(x):
    return 0.5 * x * (1.0 + np.tanh(0.7978845608028654 * (x + 0.044715 * x ** 3)))


def _act_np(x, a):
    if a == "gelu":
        return _gelu(x)
    if a == "relu":
        return np.maximum(x, 0.0)
    return x


def _bn_fold(bn: dict[str, np.ndarray], eps: float):
    scale = bn["gamma"].astype(np.float64) / np.sqrt(bn["var"].astype(np.float64) + eps)
    shift = bn["beta"].astype(np.float64) - scale * bn["mean"].astype(np.float64)
    return scale, shift


def _is_ln(nw: dict[str, np.ndarray]) -> bool:
    """A MaskedLayerNormalization owns gamma / beta only (no moving statistics, no alpha)."""
    return "alpha" not in nw and "mean" not in nw


def _attach_norm1(c: "ConvLaunch", nw: dict[str, np.ndarray], eps: float, ln_eps: float = 1e-3) -> None:
    """The norm right after a convolution, into the launch's first-norm slot."""
    if _is_ln(nw):
        cout = c.kernel.shape[2]
        c.scale1 = np.asarray(nw.get("gamma", np.ones(cout)), np.float64)
        c.shift1 = np.asarray(nw.get("beta", np.zeros(cout)), np.float64)
        c.ln1, c.ln_eps = 1, float(ln_eps)
        return
    c.scale1, c.shift1, c.dyt_g1, c.dyt_b1 = _norm_fold(nw, eps)


def _norm_fold(nw: dict[str, np.ndarray], eps: float):
    """(scale, shift, dyt_gamma, dyt_beta) of a norm layer's weights: BatchNorm folds to an affine,
    MaskedDYT to the pre-scale alpha plus the post-tanh gamma / beta."""
    if "alpha" in nw:
        c = nw["gamma"].shape[0]
        return np.full(c, float(np.asarray(nw["alpha"]).reshape(-1)[0])), np.zeros(c), nw["gamma"], nw["beta"]
    s, t = _bn_fold(nw, eps)
    return s, t, None, None


def compile_plan(spec: ModelSpec, weights: dict[str, Any]) -> Plan:
    launches: list[ConvLaunch] = []
    emb = weights.get("embedding")
    # buffers: 0 = one-hot stem operand, 1 / 2 = ping-pong activations
    cur_buf, cur_mask, cur_const = 0, 0, None
    cum_shrink = 0
    n_masks = 1
    n_taps = 0
    ch = 64

    def new_mask():
        nonlocal n_masks
        n_masks += 1
        return n_masks - 1

    def other_buf(*used):
        for b in (1, 2, 3, 4, 5):
            if b not in used:
                return b
        raise AssertionError

    halvings = 0          # stride-2 stages passed so far (frame length = ceil(L / 2^halvings))

    def add_conv(kernel, bias, dilation, padding, use_masking, in_buf, in_mask, out_buf, sc_buf=-1, sc_mask=-1, sc_const=None,
                 pad_left=None, kernel_odd=None):
        nonlocal cum_shrink
        k = kernel.shape[0]
        span = dilation * (k - 1)
        if pad_left is not None:
            shrink = 0
        elif padding == "same":
            pad_left, shrink = span // 2, 0
        else:
            pad_left, shrink = 0, span
        if shrink and halvings:
            raise NotImplementedError("a VALID convolution after a strided block")
        c = ConvLaunch(kernel=_np32(kernel), bias=_np32(bias), dilation=dilation, pad_left=pad_left, shrink=shrink,
                       in_buf=in_buf, out_buf=out_buf, sc_buf=sc_buf, mask_in=in_mask, mask_out=new_mask(),
                       sc_mask=sc_mask, masking=int(use_masking), cum_shrink_in=cum_shrink, sc_const=sc_const,
                       halvings=halvings, len_ceil=int(halvings > 0), kernel_odd=None if kernel_odd is None else _np32(kernel_odd))
        cum_shrink += shrink
        launches.append(c)
        return c

    def split_phases(kernel, dilation, cin_pad):
        """A stride-2 SAME conv over rows x[.] as a stride-1 conv over the (even, odd) row planes S[j] = (x[2j], x[2j+1]):
        x[2j + o] = S[j + o // 2][o % 2], so tap t (offset o = t * dilation - pad_left) lands on shift o // 2, phase o % 2.
        TF's SAME padding of the strided conv depends on the parity of the input length L (pad_total = (k-1)d - 1 for even L,
        (k-1)d for odd L): returns (kernel for even L, kernel for odd L, pad_left of the phase conv), both kernels
        [k', 2 * cin_pad, cout] over one common shift range."""
        k, cin, cout = kernel.shape
        span = dilation * (k - 1)
        variants = []
        for total in (max(span - 1, 0), span):                  # even L, odd L
            pl = total // 2
            variants.append([(t, (t * dilation - pl) // 2, (t * dilation - pl) % 2) for t in range(k)])
        lo = min(s for v in variants for _, s, _ in v)
        hi = max(s for v in variants for _, s, _ in v)
        out = []
        for v in variants:
            kk = np.zeros((hi - lo + 1, 2 * cin_pad, cout), np.float64)
            for t, sft, ph in v:
                kk[sft - lo, ph * cin_pad:ph * cin_pad + cin] = kernel[t]
            out.append(kk)
        return out[0], out[1], -lo

    def finish(c: ConvLaunch):
        """Fold bias into shift1 and compute the masked-row constant of the launch output."""
        cout = c.kernel.shape[2]
        s1 = np.ones(cout) if c.scale1 is None else c.scale1
        t1 = np.zeros(cout) if c.shift1 is None else c.shift1
        c.scale1 = _np32(s1)
        if c.ln1:                         # LayerNorm normalises acc + bias before gamma / beta: the bias stays a separate vector;
            c.shift1 = _np32(t1)          # the layer re-masks its output (layers.py:363-365): exactly 0 at masked rows
            c.bias = _np32(c.bias)
        else:
            c.shift1 = _np32(t1 + s1 * c.bias.astype(np.float64))     # acc*s1 + (t1 + s1*bias)
        v = np.zeros(cout) if c.ln1 else c.shift1.astype(np.float64)
        if c.dyt_g1 is not None:          # MaskedDYT multiplies its output by the mask: exactly 0 at masked rows
            c.dyt_g1, c.dyt_b1 = _np32(c.dyt_g1), _np32(c.dyt_b1)
            v = np.zeros(cout)
        if c.sc_buf >= 0:
            v = v + (np.zeros(cout) if c.sc_const is None else c.sc_const.astype(np.float64))
            c.sc_const = _np32(np.zeros(cout) if c.sc_const is None else c.sc_const)
        v = _act_np(v, c.act1)
        if c.scale2 is not None:
            c.scale2, c.shift2 = _np32(c.scale2), _np32(c.shift2)
            v = v * c.scale2 + c.shift2
            if c.dyt_g2 is not None:
                c.dyt_g2, c.dyt_b2 = _np32(c.dyt_g2), _np32(c.dyt_b2)
                v = np.zeros(cout)
            v = _act_np(v, c.act2)
        c.out_const = _np32(v)

    cur: ConvLaunch | None = None
    thresholded = False                   # a convolution with mask_mode strict / majority has been compiled
    for layer, lw in zip(spec.layers, weights["layers"]):
        cfg = layer.cfg
        if layer.kind == "conv":
            kernel = lw["kernel"]
            if not launches:
                # fold the embedding / dense-on-one-hot into the stem: W'[t] = E[1:65] @ W[t]
                extra_bias = None
                if emb is not None:
                    table = (emb[1:65] if spec.uses_token_input else emb[:64]).astype(np.float64)
                    if spec.uses_token_input and not (cfg["use_masking"] and spec.use_masking):
                        # Embedding without mask consumption: token 0 is a real row E[0].  Write
                        # x = onehot @ (E[1:] - E[0]) + E[0]; the constant part is a bias of the
                        # (VALID) stem conv.
                        if cfg["padding"] != "valid":
                            raise NotImplementedError("un-masked Embedding input needs a VALID stem conv")
                        e0 = emb[0].astype(np.float64)
                        table = table - e0
                        extra_bias = np.einsum("e,keo->o", e0, kernel.astype(np.float64))
                    kernel = np.einsum("ve,keo->kvo", table, kernel.astype(np.float64))
                elif kernel.shape[1] != 64:
                    raise NotImplementedError("stem conv without embedding must take the 64-wide one-hot")
            bias = lw.get("bias") if cfg["use_bias"] else None
            bias = np.zeros(kernel.shape[2], np.float32) if bias is None else bias
            if not launches and extra_bias is not None:
                bias = bias.astype(np.float64) + extra_bias
            out_buf = other_buf(cur_buf)
            cur = add_conv(kernel, bias, cfg["dilation"], cfg["padding"], cfg["use_masking"] and spec.use_masking,
                           cur_buf, cur_mask, out_buf)
            if cfg.get("activation"):
                cur.act1, cur.stage = cfg["activation"], 2
            if cur.masking and cfg.get("mask_mode", "any") != "any":
                k_ = kernel.shape[0]
                cur.mask_thr = k_ if cfg["mask_mode"] == "strict" else (k_ + 1) // 2
                thresholded = True
            cur_buf, cur_mask, ch = out_buf, cur.mask_out, kernel.shape[2]
        elif layer.kind == "resblock":
            if thresholded:
                # Under "strict" / "majority" a masked output row may have seen valid taps, so its un-masked value is not a
                # per-channel constant; a residual shortcut (MaskedAdd adds the UN-masked tensor, layers.py:60-76, 1910) would need
                # that value at rows a later `any` convolution re-validates, and the stored tensors keep masked rows at zero.
                raise NotImplementedError("a residual block after a masked_conv1d with mask_mode 'strict' / 'majority' is not supported "
                                          "(the shortcut would need the un-masked tensor); stand-alone convolution stacks are")
            masking = cfg["use_masking"] and spec.use_masking
            stride = int(cfg.get("strides", 1))
            zero_bias = lambda cw: cw["bias"] if cfg.get("use_bias", True) else np.zeros_like(cw["bias"])      # noqa: E731
            for blk in lw["blocks"]:
                x_buf, x_mask = cur_buf, cur_mask
                if cur is not None and cur.out_const is None:
                    finish(cur)
                x_const = launches[-1].out_const if launches else None
                src_buf, cin_pad = x_buf, -(-blk["conv1"]["kernel"].shape[1] // 64) * 64
                if stride == 2:
                    # rows -> (even, odd) row planes with twice the channels and half the length (kind 4), so that the strided
                    # convs of the block run on the stride-1 kernels (layers.py:1840-1864 conv1 / bypass conv of a strided block)
                    src_buf = other_buf(x_buf)
                    launches.append(ConvLaunch(kernel=np.zeros((1, cin_pad, 2 * cin_pad), np.float32), bias=np.zeros(2 * cin_pad, np.float32),
                                               dilation=1, pad_left=0, shrink=0, in_buf=x_buf, out_buf=src_buf, mask_in=x_mask,
                                               mask_out=x_mask, masking=0, kind=4, halvings=halvings, len_ceil=int(halvings > 0),
                                               cum_shrink_in=cum_shrink))
                    halvings += 1
                h_buf = other_buf(x_buf, src_buf)
                if stride == 2:
                    k_even, k_odd, pl = split_phases(blk["conv1"]["kernel"].astype(np.float64), cfg["dilation"], cin_pad)
                    c1 = add_conv(k_even, zero_bias(blk["conv1"]), 1, "same", masking, src_buf, x_mask, h_buf, pad_left=pl, kernel_odd=k_odd)
                else:
                    c1 = add_conv(blk["conv1"]["kernel"], zero_bias(blk["conv1"]), cfg["dilation"], "same", masking, x_buf, x_mask, h_buf)
                _attach_norm1(c1, blk["bn1"], 1e-5, cfg.get("ln_epsilon", 1e-3))
                c1.act1, c1.stage = cfg["activation"], 2
                finish(c1)
                sc_buf, sc_mask, sc_const = x_buf, (x_mask if masking else -1), x_const
                if "conv3" in blk:
                    # 1x1 bypass conv + norm on the block input (layers.py:1855-1864, 1903-1909): its own launch, no activation;
                    # conv2 adds it as the shortcut.  Strided: tap (shift 0, even phase) of the row-plane tensor.
                    b_buf = other_buf(x_buf, src_buf, h_buf)
                    k3 = blk["conv3"]["kernel"].astype(np.float64)
                    if stride == 2:
                        k3p = np.zeros((1, 2 * cin_pad, k3.shape[2]), np.float64)
                        k3p[0, :k3.shape[1]] = k3[0]
                        c3 = add_conv(k3p, zero_bias(blk["conv3"]), 1, "same", masking, src_buf, x_mask, b_buf, pad_left=0)
                    else:
                        c3 = add_conv(k3, zero_bias(blk["conv3"]), 1, "same", masking, x_buf, x_mask, b_buf)
                    _attach_norm1(c3, blk["bn3"], 1e-5, cfg.get("ln_epsilon", 1e-3))
                    c3.stage = 2
                    finish(c3)
                    sc_buf, sc_mask, sc_const = b_buf, (c3.mask_out if masking else -1), c3.out_const
                # conv2 writes the block output in place over the block input when that IS the shortcut (same rows, same
                # thread); with a bypass the block input is dead after conv1 / conv3 and its buffer is reused all the same
                c2 = add_conv(blk["conv2"]["kernel"], zero_bias(blk["conv2"]), cfg["dilation"], "same", masking,
                              h_buf, c1.mask_out, x_buf, sc_buf=sc_buf, sc_mask=sc_mask, sc_const=sc_const)
                _attach_norm1(c2, blk["bn2"], 1e-5, cfg.get("ln_epsilon", 1e-3))
                c2.act1, c2.stage = cfg["activation"], 2
                cur = c2
                cur_buf, cur_mask = x_buf, c2.mask_out
            if cfg.get("return_nmd"):
                # bn2 of the last block also returns mean(conv2 output) - its moving mean (layers.py:943-954): a tap on
                # the raw conv output of that launch
                cur.tap_mode, cur.tap_slot, cur.tap_mean = 1, n_taps, _np32(lw["blocks"][-1]["bn2"]["mean"])
                n_taps += 1
            ch = cfg["filters"]
        elif layer.kind == "nmd":
            if cur is None or cur.tap_mode != 0:
                raise NotImplementedError("nmd layer placement cannot be fused (one tap per convolution)")
            if cur.stage == 0:
                cur.tap_mode = 1
            elif cur.stage == 2:
                cur.tap_mode = 2
            elif cur.stage == 4:
                cur.tap_mode = 3                 # on the launch output: after the stand-alone norm and its activation
            else:
                raise NotImplementedError("nmd layer between a norm and its activation cannot be fused")
            cur.tap_slot, cur.tap_mean = n_taps, _np32(lw["moving_mean"])
            n_taps += 1
        elif layer.kind == "norm":
            if cur is None:
                raise NotImplementedError("norm before the first convolution")
            if _is_ln(lw):
                if cur.stage != 0:
                    raise NotImplementedError("masked_layernorm is fused as the norm right after a convolution only (per-row channel "
                                              "statistics of the accumulator); after an activation it is not supported")
                _attach_norm1(cur, lw, 1e-5, cfg.get("epsilon", 1e-3))
                cur.stage = 1
                continue
            s, t, dg, db = _norm_fold(lw, cfg.get("epsilon", 1e-5))
            if cfg.get("return_nmd"):          # masked_batchnorm(return_nmd=True): an NMD tap on this norm's input
                if cur.tap_mode != 0 or cur.stage not in (0, 2):
                    raise NotImplementedError("masked_batchnorm(return_nmd=True) placement cannot be fused (one tap per convolution)")
                cur.tap_mode = 1 if cur.stage == 0 else 2
                cur.tap_slot, cur.tap_mean = n_taps, _np32(lw["mean"])
                n_taps += 1
            if cur.stage == 0:
                cur.scale1, cur.shift1, cur.dyt_g1, cur.dyt_b1, cur.stage = s, t, dg, db, 1
            elif cur.stage == 2 and cur.scale2 is None:
                cur.scale2, cur.shift2, cur.dyt_g2, cur.dyt_b2, cur.stage = s, t, dg, db, 3
            else:
                raise NotImplementedError("more element-wise layers after a convolution than the fused epilogue holds")
        elif layer.kind == "act":
            a = cfg.get("activation")
            if cur is None:
                raise NotImplementedError("activation before the first convolution")
            if cur.stage in (0, 1):
                cur.act1, cur.stage = a, 2
            elif cur.stage == 3:
                cur.act2, cur.stage = a, 4
            else:
                raise NotImplementedError("more element-wise layers after a convolution than the fused epilogue holds")
    if cur is None:
        raise NotImplementedError("model without convolutions")
    for c in launches:
        if c.out_const is None:
            finish(c)
    # The tensor-core kernel works on channel groups of 64: narrower layers (e.g. the 32-filter
    # 500 bp model) are zero-padded -- padded output channels stay exactly 0 through bias-free
    # affine / GELU / residual / pooling, padded input channels meet zero weights.
    real_feat = ch
    tap_real_widths = {c.tap_slot: c.kernel.shape[2] for c in launches if c.tap_mode}
    for c in launches:
        k, cin, cout = c.kernel.shape
        c.real_cin, c.real_cout = cin, cout
        cin_p, cout_p = -(-cin // 64) * 64, -(-cout // 64) * 64
        if (cin_p, cout_p) == (cin, cout):
            continue
        if c.tap_mode:                   # padded channels tap to exactly 0 (zero sums, zero moving mean)
            tm = np.zeros(cout_p, np.float32)
            tm[:cout] = c.tap_mean
            c.tap_mean = tm
        kp = np.zeros((k, cin_p, cout_p), np.float32)
        kp[:, :cin, :cout] = c.kernel
        c.kernel = kp
        if c.kernel_odd is not None:
            ko = np.zeros((k, cin_p, cout_p), np.float32)
            ko[:, :cin, :cout] = c.kernel_odd
            c.kernel_odd = ko
        if c.kind == 4:
            continue

        def pad(a, fill):
            if a is None:
                return None
            out = np.full(cout_p, fill, np.float32)
            out[:cout] = a
            return out
        c.bias, c.shift1, c.shift2 = pad(c.bias, 0.0), pad(c.shift1, 0.0), pad(c.shift2, 0.0)
        c.scale1, c.scale2 = pad(c.scale1, 0.0 if c.ln1 else 1.0), pad(c.scale2, 1.0)    # LayerNorm: a padded channel normalises to (0 - mean) / sigma and must stay 0
        c.sc_const, c.out_const = pad(c.sc_const, 0.0), pad(c.out_const, 0.0)
        c.dyt_g1, c.dyt_b1, c.dyt_g2, c.dyt_b2 = pad(c.dyt_g1, 0.0), pad(c.dyt_b1, 0.0), pad(c.dyt_g2, 0.0), pad(c.dyt_b2, 0.0)
    ch = -(-ch // 64) * 64
    last = launches[-1]
    last.pool_mode = 1 if spec.pooling == "max" else 2
    last.out_buf = -1                    # the final feature map is only pooled: nothing reads it
    tap_width = 0
    for c in launches:
        if c.tap_mode:
            w = c.kernel.shape[2]
            if tap_width and tap_width != w:
                raise NotImplementedError("NMD taps of different widths")
            tap_width = w
    # NMD columns of the reference (real channels of every tap, tap order) inside the padded device layout
    nmd_cols = np.concatenate([np.arange(real_w) + slot * tap_width for slot, real_w in sorted(tap_real_widths.items())]) \
        if tap_real_widths else np.zeros(0, np.int64)
    rel = None
    rel_hidden = 0
    if spec.reliability is not None and "reliability" in weights:
        r = weights["reliability"]
        if spec.reliability[0]["activation"] != "gelu":
            raise NotImplementedError("reliability hidden activation other than gelu")
        k1 = _np32(r[0]["kernel"])
        n_sig = len(spec.reliability_signals or [])
        if len(nmd_cols) + n_sig != k1.shape[0]:
            raise ValueError(f"reliability head expects {k1.shape[0]} inputs, the layer list provides {len(nmd_cols)} NMD values + {n_sig} signals")
        k1p = np.zeros((n_taps * tap_width + n_sig, k1.shape[1]), np.float32)
        k1p[nmd_cols] = k1[:len(nmd_cols)]
        k1p[n_taps * tap_width:] = k1[len(nmd_cols):]
        rel = [k1p, _np32(r[0]["bias"]), _np32(r[1]["kernel"]), _np32(r[1]["bias"])]
        rel_hidden = r[0]["kernel"].shape[1]
    # classification head (builder.py:589-596): Dense stack; hidden layers are zero-padded to the feature width so that the
    # heads kernel can reuse one buffer (a padded hidden unit is act(0) = 0 for gelu / relu and meets zero weights)
    dense = weights["classifier"]
    if len(dense) != len(spec.classifier):
        raise ValueError(f"classifier head: {len(spec.classifier)} Dense layers in the project, {len(dense)} in the weights")
    mlp, mlp_act, width = None, None, real_feat
    if len(dense) > 1:
        mlp_act = spec.classifier[0]["activation"]
        mlp = []
        for d in dense[:-1]:
            kin, kout = d["kernel"].shape
            if kin != width or kout > ch:
                raise NotImplementedError(f"classifier hidden layer {kin} -> {kout}: input must match, width must not exceed the feature width {ch}")
            wp, bp = np.zeros((ch, ch), np.float32), np.zeros(ch, np.float32)
            wp[:kin, :kout], bp[:kout] = d["kernel"], d["bias"]
            mlp += [wp, bp]
            width = kout
    if dense[-1]["kernel"].shape[0] != width:
        raise ValueError(f"classifier output layer expects {dense[-1]['kernel'].shape[0]} inputs, the head provides {width}")
    cls_w = np.zeros((ch, spec.n_classes), np.float32)
    cls_w[:width] = dense[-1]["kernel"]
    return Plan(launches=launches, n_classes=spec.n_classes, feat_dim=ch, pool_mode=last.pool_mode, n_taps=n_taps,
                tap_width=tap_width, cls_w=cls_w, cls_b=_np32(dense[-1]["bias"]), rel=rel,
                rel_hidden=rel_hidden, total_shrink=cum_shrink, real_feat_dim=real_feat,
                rel_signals=list(spec.reliability_signals) if (rel is not None and spec.reliability_signals) else None,
                nmd_cols=nmd_cols if len(nmd_cols) != n_taps * tap_width else None,
                mlp=mlp, mlp_act=mlp_act, mlp_layers=len(dense) - 1, emb_pooled=True)


def _fptr(a: np.ndarray | None):
    if a is None:
        return ctypes.POINTER(ctypes.c_float)()
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def to_ctypes(plan: Plan):
    """(LayerDesc array, HeadDesc) referencing the plan's arrays (kept alive by `plan.keep`)."""
    arr = (LayerDesc * len(plan.launches))()
    for d, c in zip(arr, plan.launches):
        k, cin, cout = c.kernel.shape
        vals = {LF_KIND: c.kind, LF_HALVINGS: c.halvings, LF_CIN: cin, LF_COUT: cout, LF_K: k, LF_DIL: c.dilation, LF_PAD_LEFT: c.pad_left,
                LF_SHRINK: c.shrink, LF_IN_BUF: c.in_buf, LF_OUT_BUF: c.out_buf, LF_SC_BUF: c.sc_buf,
                LF_ACT1: ACT[c.act1], LF_HAS_AFF2: int(c.scale2 is not None), LF_ACT2: ACT[c.act2],
                LF_TAP_MODE: c.tap_mode, LF_TAP_SLOT: c.tap_slot, LF_POOL_MODE: c.pool_mode, LF_MASK_IN: c.mask_in,
                LF_MASK_OUT: c.mask_out, LF_SC_MASK: c.sc_mask, LF_MASKING: c.masking,
                LF_CUM_SHRINK_IN: c.cum_shrink_in, LF_DYT1: int(c.dyt_g1 is not None), LF_DYT2: int(c.dyt_g2 is not None),
                LF_EPI_F32: c.epi_f32, LF_LEN_CEIL: c.len_ceil, LF_REAL_CIN: c.real_cin, LF_REAL_COUT: c.real_cout,
                LF_LN1: c.ln1, LF_LN_EPS: int(np.float32(c.ln_eps).view(np.int32)), LF_MASK_THR: c.mask_thr}
        for i, v in vals.items():
            d.i[i] = int(v)
        ptrs = {LP_KERNEL: c.kernel, LP_BIAS: c.bias, LP_SCALE1: c.scale1, LP_SHIFT1: c.shift1, LP_SCALE2: c.scale2,
                LP_SHIFT2: c.shift2, LP_SC_CONST: c.sc_const, LP_TAP_MEAN: c.tap_mean,
                LP_DYT_G1: c.dyt_g1, LP_DYT_B1: c.dyt_b1, LP_DYT_G2: c.dyt_g2, LP_DYT_B2: c.dyt_b2, LP_KERNEL_ODD: c.kernel_odd}
        for i, a in ptrs.items():
            d.p[i] = _fptr(a)
    h = HeadDesc()
    h.n_classes, h.feat_dim, h.pool_mode = plan.n_classes, plan.feat_dim, plan.pool_mode
    h.n_taps, h.rel_hidden = plan.n_taps, plan.rel_hidden
    h.cls_w, h.cls_b = _fptr(plan.cls_w), _fptr(plan.cls_b)
    if plan.rel is not None:
        h.rel_w1, h.rel_b1, h.rel_w2, h.rel_b2 = (_fptr(a) for a in plan.rel)
        if plan.rel_signals:
            from .modelspec import OOD_SIGNALS
            code = len(plan.rel_signals)
            for i, name in enumerate(plan.rel_signals):
                code |= (OOD_SIGNALS.index(name) + 1) << (3 * (i + 1))
            h.reserved[0] = code
    if plan.mlp is not None:
        h.mlp_hidden, h.mlp_act = plan.mlp[0].shape[1], ACT[plan.mlp_act]
        ptrs4 = [_fptr(a) for a in plan.mlp] + [_fptr(None)] * (4 - len(plan.mlp))
        h.mlp_w1, h.mlp_b1, h.mlp_w2, h.mlp_b2 = ptrs4
        h.reserved[0] |= (int(plan.mlp_layers) << 20) | (int(bool(plan.emb_pooled)) << 22)
    plan.keep = [arr, h]
    return arr, h
