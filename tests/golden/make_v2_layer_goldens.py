"""Layer-level goldens of the fragment-model (v2) layers, computed by EXECUTING THE REFERENCE'S OWN `call` BODIES
(nnlib/v2/layers.py: GeLU, MaskedConv1D in all three mask modes, MaskedBatchNorm at inference incl. return_nmd, MaskedDYT,
MaskedGlobalMaxPooling, MaskedGlobalAvgPooling, OODSignalLayer; nnlib/v2/nmd.py: NMDLayer) on top of tests/golden/tf_standin.py, a NumPy
stand-in for the ~30 TensorFlow symbols those bodies use (TensorFlow / Keras are not installable here).  The layer math --
what is masked, what the mask becomes, epsilons, which statistics, the NMD definition, pooling sentinels -- is the
reference's code; the primitives (conv1d with TF SAME padding, reductions, tanh-GELU) are the stand-in's, the conv padding
rule being pinned separately on the reference's serialized TF graph (tests/test_legacy_graph_pin.py).  The single layers are called
alone with explicit masks.  The last section runs ResidualBlockStack / ResidualBlock.call from the reference source; there the
mask hand-over between the inner layers follows the stand-in's restatement of Keras 3's `Layer.__call__` rules (see
tf_standin.Layer) -- the block's CODE is the reference's, those four rules are not executed from Keras.
Writes tests/golden/v2_layers.npz (inputs are seeded and stored, outputs float64).

usage:  python tests/golden/make_v2_layer_goldens.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT))
sys.path.insert(0, "/root/reference/src")

import tf_standin          # noqa: E402

tf_standin.install()
from tf_standin import t          # noqa: E402


def main():
    from jaeger.nnlib.v2 import layers as L
    from jaeger.nnlib.v2 import nmd as N
    rng = np.random.default_rng(2024)
    out = {}
    b, f, length, cin, cout = 2, 6, 37, 5, 7
    x = rng.normal(size=(b, f, length, cin))
    mask = rng.random((b, f, length)) < 0.7
    mask[0, :, 10:22] = False                     # a long unknown run (survives an `any` conv)
    mask[1, :, 25:] = False                       # right padding
    out["x"], out["mask"] = x, mask
    # ---- MaskedConv1D -----------------------------------------------------------------------------
    k = 5
    kernel, bias = rng.normal(size=(k, cin, cout)) * 0.3, rng.normal(size=cout) * 0.2
    out["conv_kernel"], out["conv_bias"] = kernel, bias
    for padding in ("valid", "same"):
        for dil in (1, 3):
            for mode in ("any", "majority", "strict"):
                layer = L.MaskedConv1D(filters=cout, kernel_size=k, strides=1, padding=padding, dilation_rate=dil, use_bias=True,
                                       activation=None, mask_mode=mode)
                layer.build((b, f, length, cin))
                layer.kernel, layer.bias = t(kernel), t(bias)
                y = layer.call(t(x), mask=t(mask))
                om = layer.compute_mask(t(x), mask=t(mask))
                tag = f"conv_{padding}_d{dil}_{mode}"
                out[tag + "_y"], out[tag + "_mask"] = np.asarray(y), np.asarray(om).astype(bool)
    layer = L.MaskedConv1D(filters=cout, kernel_size=k, padding="valid", use_bias=False, activation="gelu")
    layer.build((b, f, length, cin))
    layer.kernel = t(kernel)
    out["conv_nomask_gelu_y"] = np.asarray(layer.call(t(x), mask=None))
    # ---- element-wise layers ------------------------------------------------------------------------
    c = cout
    h = rng.normal(size=(b, f, length, c)) * 1.5
    hm = rng.random((b, f, length)) < 0.8
    hm[1] = False; hm[1, :, :9] = True
    out["h"], out["h_mask"] = h, hm
    bn = L.MaskedBatchNorm(return_nmd=True)
    bn.build((b, f, length, c))
    bn.gamma, bn.beta = t(rng.uniform(0.5, 1.5, c)), t(rng.normal(0, 0.2, c))
    bn.moving_mean, bn.moving_variance = t(rng.normal(0, 0.3, c)), t(rng.uniform(0.4, 1.6, c))
    for k_ in ("gamma", "beta", "moving_mean", "moving_variance"):
        out["bn_" + k_] = np.asarray(getattr(bn, k_))
    y, nm = bn.call(t(h), mask=t(hm), training=False)
    out["bn_y"], out["bn_nmd"] = np.asarray(y), np.asarray(nm)
    y, nm = bn.call(t(h), mask=None, training=False)
    out["bn_y_nomask"], out["bn_nmd_nomask"] = np.asarray(y), np.asarray(nm)
    dyt = L.MaskedDYT(alpha_init=0.5)
    dyt.build((b, f, length, c))
    dyt.alpha, dyt.gamma, dyt.beta = t(np.array([0.37])), t(rng.uniform(0.5, 1.5, c)), t(rng.normal(0, 0.2, c))
    out["dyt_alpha"], out["dyt_gamma"], out["dyt_beta"] = (np.asarray(getattr(dyt, a)) for a in ("alpha", "gamma", "beta"))
    out["dyt_y"], out["dyt_y_nomask"] = np.asarray(dyt.call(t(h), mask=t(hm))), np.asarray(dyt.call(t(h), mask=None))
    ln = L.MaskedLayerNormalization(epsilon=1e-3)                   # layers.py:293-367
    ln.build((b, f, length, c))
    ln.gamma, ln.beta = t(rng.uniform(0.5, 1.5, c)), t(rng.normal(0, 0.2, c))
    out["ln_gamma"], out["ln_beta"] = np.asarray(ln.gamma), np.asarray(ln.beta)
    out["ln_y"], out["ln_y_nomask"] = np.asarray(ln.call(t(h), mask=t(hm))), np.asarray(ln.call(t(h), mask=None))
    nl = N.NMDLayer()
    nl.build((b, f, length, c))
    nl.moving_mean = t(rng.normal(0, 0.3, c))
    out["nmd_moving_mean"] = np.asarray(nl.moving_mean)
    out["nmd_y"], out["nmd_y_nomask"] = np.asarray(nl.call(t(h), mask=t(hm), training=False)), np.asarray(nl.call(t(h), mask=None, training=False))
    out["gelu_y"] = np.asarray(L.GeLU().call(t(h)))
    # ---- pooling -------------------------------------------------------------------------------------
    pm = hm.copy(); pm[1] = False                 # a fully masked sample
    out["pool_mask"] = pm
    out["maxpool_y"], out["maxpool_y_nomask"] = np.asarray(L.MaskedGlobalMaxPooling().call(t(h), mask=t(pm))), np.asarray(L.MaskedGlobalMaxPooling().call(t(h), mask=None))
    out["avgpool_y"], out["avgpool_y_nomask"] = np.asarray(L.MaskedGlobalAvgPooling().call(t(h), mask=t(pm))), np.asarray(L.MaskedGlobalAvgPooling().call(t(h), mask=None))
    # ---- OODSignalLayer (reliability_model.mode nmd_plus_signals) -----------------------------------------
    logits = rng.normal(0, 2.5, (9, 6))
    logits[0] = 1.0                                  # uniform
    logits[1, 2] = logits[1, 4] = 7.0                # tie at the top
    logits[2, 0] = 40.0                              # one dominant class
    nmd_vec = rng.normal(0, 0.4, (9, 20))
    out["ood_logits"], out["ood_nmd"] = logits, nmd_vec
    signals = ["max_prob", "entropy", "energy", "margin", "nmd_norm"]
    out["ood_y"] = np.asarray(L.OODSignalLayer(signals=signals).call({"logits": t(logits), "nmd": t(nmd_vec)}))
    # ---- ResidualBlockStack: the reference's block code under the stand-in's Keras-3 __call__ mask rules -------------------
    from tf_standin import get_keras_mask
    cb, kb, db = 8, 5, 3
    xb = rng.normal(size=(2, 6, 41, cb))
    mb = rng.random((2, 6, 41)) < 0.75
    mb[0, :, 8:25] = False
    mb[1, :, 30:] = False
    out["block_x"], out["block_mask"] = xb, mb
    for norm in ("masked_batchnorm", "masked_dyt", "masked_layernorm"):
        for masking in (True, False):
            kw = dict(filters=cb, kernel_size=kb, dilation_rate=db, use_bias=True, norm_type=norm, use_masking=masking, name="resblock_1")
            if norm == "masked_batchnorm":
                kw["return_nmd"] = True
            stack = L.ResidualBlockStack(2, (6, None, cb), **kw)
            xin = t(xb)
            if masking:
                xin._keras_mask = t(mb)
            stack(xin)                                           # builds the sub-layers
            tag = f"block_{ {'masked_batchnorm': 'bn', 'masked_dyt': 'dyt', 'masked_layernorm': 'ln'}[norm] }_{int(masking)}"
            for bi, blk in enumerate(stack.blocks):
                for cname in ("conv1", "conv2"):
                    conv = getattr(blk, cname)
                    conv.kernel, conv.bias = t(rng.normal(size=(kb, cb, cb)) * 0.25), t(rng.normal(size=cb) * 0.2)
                    out[f"{tag}_b{bi}_{cname}_kernel"], out[f"{tag}_b{bi}_{cname}_bias"] = np.asarray(conv.kernel), np.asarray(conv.bias)
                for nname in ("bn1", "bn2"):
                    nl_ = getattr(blk, nname)
                    if norm == "masked_batchnorm":
                        nl_.gamma, nl_.beta = t(rng.uniform(0.5, 1.5, cb)), t(rng.normal(0, 0.2, cb))
                        nl_.moving_mean, nl_.moving_variance = t(rng.normal(0, 0.3, cb)), t(rng.uniform(0.4, 1.6, cb))
                        names = ("gamma", "beta", "moving_mean", "moving_variance")
                    elif norm == "masked_layernorm":
                        nl_.gamma, nl_.beta = t(rng.uniform(0.5, 1.5, cb)), t(rng.normal(0, 0.2, cb))
                        names = ("gamma", "beta")
                    else:
                        nl_.alpha, nl_.gamma, nl_.beta = t(np.array([rng.uniform(0.3, 0.7)])), t(rng.uniform(0.5, 1.5, cb)), t(rng.normal(0, 0.2, cb))
                        names = ("alpha", "gamma", "beta")
                    for a in names:
                        out[f"{tag}_b{bi}_{nname}_{a}"] = np.asarray(getattr(nl_, a))
            xin = t(xb)
            if masking:
                xin._keras_mask = t(mb)
            res = stack(xin)
            y = res[0] if isinstance(res, (list, tuple)) else res
            out[tag + "_y"] = np.asarray(y)
            if isinstance(res, (list, tuple)):
                out[tag + "_nmd"] = np.asarray(res[1])
            m = get_keras_mask(y)
            out[tag + "_outmask"] = np.zeros(0, bool) if m is None else np.asarray(m).astype(bool)
            print(tag, np.asarray(y).shape, "outgoing mask:", None if m is None else int(np.asarray(m).sum()), "of", mb.size, "(input", int(mb.sum()), ")")
    np.savez_compressed(OUT / "v2_layers.npz", **out)
    print("written", OUT / "v2_layers.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
