"""Whole representation learner from the REFERENCE's builder code: `DynamicModelBuilder._build_block`
(nnlib/builder.py:982-1193) is called eagerly -- the "symbolic input" is real data -- on the NumPy stand-in for TensorFlow /
Keras (tests/golden/tf_standin.py), so the layer order, the config keys handed to every layer, the collection and
concatenation of the NMD vectors and the pooling are the reference's code, the layers are the reference's classes
(nnlib/v2/layers.py, nmd.py), and only the TF primitives and Keras' four `__call__` mask rules are the stand-in's.
Input = the Embedding(mask_zero=True) output the builder creates first (builder.py:844-894): E[token], mask = token != 0.
Weights come from a seeded provider and are stored in creation order.  Writes tests/golden/v2_model.npz.

usage:  python tests/golden/make_v2_model_goldens.py
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(OUT.parent.parent))

import tf_standin          # noqa: E402

tf_standin.install()
from tf_standin import t          # noqa: E402

for _m in ("pyfastx", "pydustmasker", "parasail", "polars", "ruptures", "kneed", "h5py"):
    sys.modules.setdefault(_m, types.ModuleType(_m))


def provider_for(seed):
    rng = np.random.default_rng(seed)

    def provider(layer, name, shape):
        if name == "kernel":
            return rng.normal(size=shape) * 0.2
        if name == "bias":
            return rng.normal(size=shape) * 0.1
        if name in ("gamma", "moving_variance"):
            return rng.uniform(0.5, 1.5, shape)
        if name == "alpha":
            return rng.uniform(0.3, 0.7, shape)
        return rng.normal(size=shape) * 0.25               # beta, moving_mean
    return provider


def model_cfg(norm: str, masking: bool, pooling: str):
    """A small member of the nmd_merge family (train_config/nn_config_1500bp_nmd_merge_6_class_*.yaml): stem conv k7 VALID,
    NMD taps, two residual stacks (k5, dilation 3), stand-alone norms + activations."""
    nname = "masked_batchnorm" if norm == "bn" else "masked_dyt"
    tail = [{"name": "nmd"}, {"name": nname, "config": {} if norm == "dyt" else {"return_nmd": False}},
            {"name": "activation", "config": {"activation": "gelu"}}]
    block = {"name": "residual_block", "config": {"use_1x1conv": False, "block_size": 2, "filters": 16, "kernel_size": 5, "dilation_rate": 3,
                                                  "use_bias": True, **({"norm_type": "masked_dyt"} if norm == "dyt" else {})}}
    hidden = [{"name": "masked_conv1d", "config": {"filters": 16, "kernel_size": 7, "strides": 1, "dilation_rate": 1, "use_bias": True,
                                                   "activation": None}}] + tail + [block] + tail + [block] + tail
    return {"hidden_layers": hidden, "pooling": pooling}, masking


def strided_cfg():
    """Strided / bypass residual blocks (nnlib/v2/layers.py:1840-1864, 1903-1909; train_config/nn_config_baseline.yaml family): a
    stride-1 block with use_1x1conv, a stride-2 block that grows the channels, a two-block stack with dilation 2 whose last bn2
    returns its NMD vector.  Masking off (strided blocks are only well-defined without it)."""
    def blk(**kw):
        return {"name": "residual_block", "config": {"use_bias": True, "activation": "gelu", **kw}}
    tail = [{"name": "nmd"}, {"name": "masked_batchnorm", "config": {"return_nmd": False}}, {"name": "activation", "config": {"activation": "gelu"}}]
    hidden = [{"name": "masked_conv1d", "config": {"filters": 16, "kernel_size": 7, "strides": 1, "dilation_rate": 1, "use_bias": True,
                                                   "activation": None}}] + tail + [
        blk(use_1x1conv=True, block_size=1, filters=16, kernel_size=5, strides=1, dilation_rate=1),
        blk(use_1x1conv=False, block_size=1, filters=24, kernel_size=5, strides=2, dilation_rate=1),
        blk(use_1x1conv=False, block_size=2, filters=24, kernel_size=3, strides=1, dilation_rate=2, return_nmd=True)]
    return {"hidden_layers": hidden, "pooling": "max"}, False


def main():
    from jaeger.nnlib import builder as B
    out = {}
    rng = np.random.default_rng(99)
    tok = rng.integers(1, 65, size=(3, 6, 70))
    tok[rng.random(tok.shape) < 0.03] = 0
    tok[0, :, 20:45] = 0
    tok[1, :, 50:] = 0
    emb = rng.normal(size=(65, 12)) * 0.5
    out["tokens"], out["embedding_table"] = tok.astype(np.uint8), emb
    for ci, (norm, masking, pooling) in enumerate([("bn", True, "max"), ("dyt", True, "average"), ("bn", False, "max"), ("strided", False, "max")]):
        cfg, use_masking = strided_cfg() if norm == "strided" else model_cfg(norm, masking, pooling)
        fake = types.SimpleNamespace(use_masking=use_masking, model_cfg={}, input_shape=(6, None), _make_regularizer=lambda *a, **k: None)
        real = B.DynamicModelBuilder.__new__(B.DynamicModelBuilder)
        fake._layers = {"masked_conv1d": B.MaskedConv1D, "masked_batchnorm": B.MaskedBatchNorm, "masked_dyt": B.MaskedDYT, "nmd": B.NMDLayer,
                        "residual_block": B.ResidualBlock_wrapper, "activation": sys.modules["tensorflow"].keras.layers.Activation}
        fake._get_pooler = lambda name, _r=real: B.DynamicModelBuilder._get_pooler(_r, name)
        tf_standin.WEIGHT_LOG.clear()
        tf_standin.WEIGHT_PROVIDER = provider_for(1000 + ci)
        x = t(emb[tok])                                           # Embedding(vocab 65, mask_zero=True), builder.py:844-868
        if use_masking:
            x._keras_mask = t(tok != 0)
        res = B.DynamicModelBuilder._build_block(fake, x, cfg, prefix="rep", nmd_merge=None)
        tf_standin.WEIGHT_PROVIDER = None
        feat, nmd = res
        tag = f"m{ci}"
        out[tag + "_feat"], out[tag + "_nmd"] = np.asarray(feat), np.asarray(nmd)
        out[tag + "_cfg"] = np.array([norm, str(int(masking)), pooling])
        for wi, (lname, wname, arr) in enumerate(tf_standin.WEIGHT_LOG):
            out[f"{tag}_w{wi:03d}_{wname}"] = np.asarray(arr)
        out[tag + "_wnames"] = np.array([f"{l}/{w}" for l, w, _ in tf_standin.WEIGHT_LOG])
        print(tag, norm, masking, pooling, "feat", np.asarray(feat).shape, "nmd", np.asarray(nmd).shape, len(tf_standin.WEIGHT_LOG), "weights")
    np.savez_compressed(OUT / "v2_model.npz", **out)


if __name__ == "__main__":
    main()
