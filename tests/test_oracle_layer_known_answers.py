"""Pins oracle/forward.py's layer restatements on the known answers the reference's own unit tests hold for
them (the only layer-level facts available without TensorFlow): tests/unit/test_mask_mode.py (output-mask rules
of MaskedConv1D in the three mask modes), tests/unit/test_masked_pooling.py (masked global max / average
pooling, the fully-masked sample), tests/unit/test_nnlib_v2_nmd.py (NMD vector = masked channel mean minus the
moving mean) and tests/unit/test_inference_crop.py:42-91 (frame lengths after the VALID stem).  Each test
names the reference test it restates; inputs are rebuilt with the same shapes / seeds where the reference
uses NumPy generators."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import forward as fwd


def _out_mask(mask_1d: np.ndarray, mode: str, k: int = 5, padding: str = "valid", dilation: int = 1) -> np.ndarray:
    """test_mask_mode.py:_run_mask_mode: a MaskedConv1D over ones((1, 1, L, 4)), output mask flattened."""
    length = mask_1d.shape[-1]
    x = torch.ones(1, 1, length, 4)
    kernel = torch.ones(k, 4, 4)
    _, om = fwd.masked_conv1d(x, torch.as_tensor(mask_1d, dtype=torch.float32).reshape(1, 1, length), kernel, None,
                              dilation, padding, None, mask_mode=mode)
    return om.numpy().reshape(-1) > 0


def _mask_with_n(length: int, n_pos: list[int]) -> np.ndarray:
    m = np.ones(length, dtype=bool)
    m[n_pos] = False
    return m


def test_mask_mode_isolated_n():
    """test_isolated_n_any_kills_nothing / _strict_kills_kernel_window / _majority_survives."""
    mask = _mask_with_n(20, [10])
    assert _out_mask(mask, "any").all()
    expected = np.ones(16, dtype=bool)
    expected[6:11] = False
    np.testing.assert_array_equal(_out_mask(mask, "strict"), expected)
    assert _out_mask(mask, "majority").all()


def test_mask_mode_n_runs():
    """test_short_n_run_any_kills_nothing, test_long_n_run_any_kills_only_the_run, _strict_kills_run_plus_halo."""
    assert _out_mask(_mask_with_n(20, [9, 10, 11]), "any").all()
    run5 = _mask_with_n(20, [9, 10, 11, 12, 13])
    expected = np.ones(16, dtype=bool)
    expected[9] = False
    np.testing.assert_array_equal(_out_mask(run5, "any"), expected)
    expected = np.ones(16, dtype=bool)
    expected[5:14] = False
    np.testing.assert_array_equal(_out_mask(run5, "strict"), expected)


def test_mask_mode_right_padding():
    """test_right_padding_any_keeps_real_content: 10 real positions right-padded to 20."""
    mask = _mask_with_n(20, list(range(10, 20)))
    out_any, out_strict = _out_mask(mask, "any"), _out_mask(mask, "strict")
    assert out_any[:10].all() and not out_any[10:].any()
    assert out_strict[:6].all() and not out_strict[6:].any()


def test_conv_output_lengths():
    """MaskedConv1D.compute_output_shape (layers.py:1315-1332) and test_inference_crop.py:42-91: a 665-codon
    frame leaves the VALID k7 stem with 659 positions; SAME keeps the length, dilated or not."""
    x = torch.zeros(1, 6, 665, 4)
    y, _ = fwd.masked_conv1d(x, None, torch.zeros(7, 4, 8), None, 1, "valid")
    assert y.shape == (1, 6, 659, 8)
    y, _ = fwd.masked_conv1d(y, None, torch.zeros(5, 8, 8), None, 3, "same")
    assert y.shape == (1, 6, 659, 8)
    x = torch.zeros(1, 6, 498, 4)
    assert fwd.masked_conv1d(x, None, torch.zeros(7, 4, 8), None, 1, "valid")[0].shape[2] == 492


def test_same_padding_split_is_floor_left_ceil_right():
    """TF SAME with an even total pad (k4: 3 -> left 1, right 2): an impulse response
    shows where tap 0 lands."""
    x = torch.zeros(1, 1, 9, 1)
    x[0, 0, 4, 0] = 1.0
    kernel = torch.arange(1.0, 5.0).reshape(4, 1, 1)
    y, _ = fwd.masked_conv1d(x, None, kernel, None, 1, "same")
    # y[i] = sum_t w[t] * x[i - 1 + t]  ->  the impulse at 4 shows w[t] at i = 5 - t
    np.testing.assert_array_equal(y.numpy().reshape(-1), [0, 0, 4, 3, 2, 1, 0, 0, 0])


def _padded_pair(batch=2, strands=6, length=32, dim=8, valid=20, seed=0):
    """test_masked_pooling.py:_padded_pair."""
    rng = np.random.default_rng(seed)
    x = torch.as_tensor(rng.normal(size=(batch, strands, length, dim)), dtype=torch.float32)
    mask = torch.cat([torch.ones(batch, strands, valid), torch.zeros(batch, strands, length - valid)], dim=-1)
    return x, mask, valid


def test_masked_max_pooling_known_answers():
    """test_masked_max_pooling_matches_truncated_max, _excludes_padded_constants, _no_mask_matches_stock."""
    x, mask, valid = _padded_pair()
    np.testing.assert_allclose(fwd.masked_global_max(x, mask).numpy(), x[:, :, :valid].amax(dim=(1, 2)).numpy(), atol=1e-6)
    x2 = torch.where(mask.unsqueeze(-1) > 0, x * 0.01, torch.tensor(1e6))
    np.testing.assert_allclose(fwd.masked_global_max(x2, mask).numpy(), x2[:, :, :valid].amax(dim=(1, 2)).numpy(), atol=1e-6)
    np.testing.assert_allclose(fwd.masked_global_max(x, None).numpy(), x.amax(dim=(1, 2)).numpy(), atol=1e-6)


def test_fully_masked_sample_pools_to_zero_not_sentinel():
    """test_fully_masked_sample_pools_to_zero_not_sentinel."""
    x = torch.randn(2, 6, 16, 8, generator=torch.Generator().manual_seed(0))
    mask = torch.cat([torch.ones(1, 6, 16), torch.zeros(1, 6, 16)], dim=0)
    pooled = fwd.masked_global_max(x, mask).numpy()
    np.testing.assert_allclose(pooled[0], x[0].amax(dim=(0, 1)).numpy(), atol=1e-6)
    np.testing.assert_array_equal(pooled[1], np.zeros(8, np.float32))


def test_masked_avg_pooling_known_answers():
    """test_masked_avg_pooling_matches_masked_mean, _no_mask_matches_stock."""
    x, mask, valid = _padded_pair()
    np.testing.assert_allclose(fwd.masked_global_avg(x, mask).numpy(), x[:, :, :valid].mean(dim=(1, 2)).numpy(), atol=1e-6)
    np.testing.assert_allclose(fwd.masked_global_avg(x, None).numpy(), x.mean(dim=(1, 2)).numpy(), atol=1e-6)


def test_nmd_vector_is_masked_mean_minus_moving_mean():
    """test_nmd_matches_masked_batch_norm_with_mask: NMDLayer and MaskedBatchNorm(return_nmd=True) both return
    (masked per-example channel mean) - moving_mean (nmd.py:52-77, layers.py:943-966); restated from the formula."""
    g = torch.Generator().manual_seed(42)
    x = torch.randn(4, 6, 32, 8, generator=g)
    mask = torch.cat([torch.ones(2, 6, 32), (torch.rand(2, 6, 32, generator=g) < 0.7).float()], dim=0)
    mm = torch.randn(8, generator=g) * 0.1
    got = fwd.nmd_vector(x, mask, mm).numpy()
    for b in range(4):
        sel = mask[b] > 0
        want = x[b][sel].sum(dim=0) / (sel.sum() + 1e-5) - mm
        np.testing.assert_allclose(got[b], want.numpy(), atol=1e-5)
    np.testing.assert_allclose(fwd.nmd_vector(x, None, mm).numpy(), (x.mean(dim=(1, 2)) - mm).numpy(), atol=1e-6)


@pytest.mark.parametrize("mode,thr", [("any", 1), ("majority", 3), ("strict", 5)])
def test_mask_threshold_rule_on_random_masks(mode, thr):
    """layers.py:1245-1252: any = count > 0, majority = count >= (k + 1) // 2, strict = count == k, with TF SAME
    zero padding counting as invalid; checked by brute force on random masks, dilation 3."""
    rng = np.random.default_rng(3)
    k, d, length = 5, 3, 40
    for _ in range(20):
        m = rng.random(length) < 0.6
        padded = np.concatenate([np.zeros(6, bool), m, np.zeros(6, bool)])     # total pad d*(k-1) = 12, left 6
        cnt = np.array([sum(padded[i + t * d] for t in range(k)) for i in range(length)])
        np.testing.assert_array_equal(_out_mask(m, mode, k, "same", d), cnt >= thr)


def test_oracle_layers_vs_reference_call_bodies():
    """tests/golden/v2_layers.npz = the reference's own `call` bodies (MaskedConv1D in all three mask modes, MaskedBatchNorm at
    inference with return_nmd, MaskedDYT, NMDLayer, GeLU, masked global max / average pooling) executed on a NumPy stand-in
    for TensorFlow (tests/golden/make_v2_layer_goldens.py): every oracle restatement equals them to 1e-12, masks exactly."""
    from pathlib import Path
    z = np.load(Path(__file__).resolve().parent / "golden" / "v2_layers.npz")
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)   # noqa: E731
    x, mask = T(z["x"]), T(z["mask"].astype(np.float64))
    kernel, bias = T(z["conv_kernel"]), T(z["conv_bias"])
    for padding in ("valid", "same"):
        for dil in (1, 3):
            for mode in ("any", "majority", "strict"):
                tag = f"conv_{padding}_d{dil}_{mode}"
                y, om = fwd.masked_conv1d(x, mask, kernel, bias, dil, padding, None, mask_mode=mode)
                assert np.abs(y.numpy() - z[tag + "_y"]).max() < 1e-12, tag
                assert np.array_equal(om.numpy() > 0, z[tag + "_mask"]), tag
    y, om = fwd.masked_conv1d(x, None, kernel, None, 1, "valid", "gelu")
    assert om is None and np.abs(y.numpy() - z["conv_nomask_gelu_y"]).max() < 1e-12
    assert not np.array_equal(z["conv_same_d3_any_mask"], z["conv_same_d3_strict_mask"])
    h, hm = T(z["h"]), T(z["h_mask"].astype(np.float64))
    bn = {"gamma": T(z["bn_gamma"]), "beta": T(z["bn_beta"]), "mean": T(z["bn_moving_mean"]), "var": T(z["bn_moving_variance"])}
    assert np.abs(fwd.batchnorm(h, bn).numpy() - z["bn_y"]).max() < 1e-12            # inference BN does not re-mask
    assert np.abs(fwd.batchnorm(h, bn).numpy() - z["bn_y_nomask"]).max() < 1e-12
    assert np.abs(fwd.nmd_vector(h, hm, bn["mean"]).numpy() - z["bn_nmd"]).max() < 1e-12       # return_nmd: NMD of the norm's input
    assert np.abs(fwd.nmd_vector(h, None, bn["mean"]).numpy() - z["bn_nmd_nomask"]).max() < 1e-12
    lw = {"gamma": T(z["ln_gamma"]), "beta": T(z["ln_beta"])}                       # MaskedLayerNormalization (layers.py:337-367)
    assert np.abs(fwd.layernorm(h, lw, hm, 1e-3).numpy() - z["ln_y"]).max() < 1e-12
    assert np.abs(fwd.layernorm(h, lw, None, 1e-3).numpy() - z["ln_y_nomask"]).max() < 1e-12
    assert np.all(z["ln_y"][~z["h_mask"]] == 0.0)                                   # re-masked: exactly zero at masked rows
    dw = {"alpha": T(z["dyt_alpha"]), "gamma": T(z["dyt_gamma"]), "beta": T(z["dyt_beta"])}
    assert np.abs(fwd.dyt(h, dw, hm).numpy() - z["dyt_y"]).max() < 1e-12
    assert np.abs(fwd.dyt(h, dw, None).numpy() - z["dyt_y_nomask"]).max() < 1e-12
    mm = T(z["nmd_moving_mean"])
    assert np.abs(fwd.nmd_vector(h, hm, mm).numpy() - z["nmd_y"]).max() < 1e-12
    assert np.abs(fwd.nmd_vector(h, None, mm).numpy() - z["nmd_y_nomask"]).max() < 1e-12
    assert np.abs(fwd.gelu_tanh(h).numpy() - z["gelu_y"]).max() < 1e-12
    pm = T(z["pool_mask"].astype(np.float64))
    assert np.abs(fwd.masked_global_max(h, pm).numpy() - z["maxpool_y"]).max() < 1e-12
    assert np.abs(fwd.masked_global_max(h, None).numpy() - z["maxpool_y_nomask"]).max() < 1e-12
    assert np.abs(fwd.masked_global_avg(h, pm).numpy() - z["avgpool_y"]).max() < 1e-12
    assert np.abs(fwd.masked_global_avg(h, None).numpy() - z["avgpool_y_nomask"]).max() < 1e-12
    assert np.all(z["maxpool_y"][1] == 0.0) and np.all(z["avgpool_y"][1] == 0.0)              # the fully masked sample
    sig = fwd.ood_signals(T(z["ood_logits"]), T(z["ood_nmd"]), ["max_prob", "entropy", "energy", "margin", "nmd_norm"])
    assert np.abs(sig.numpy() - z["ood_y"]).max() < 1e-12                                      # OODSignalLayer (layers.py:1632-1666)
    assert z["ood_y"][1, 3] == 0.0 and abs(z["ood_y"][0, 0] - 1 / 6) < 1e-12                   # tie: margin 0; uniform: max_prob 1/6


@pytest.mark.parametrize("norm", ["bn", "dyt", "ln"])
@pytest.mark.parametrize("masking", [1, 0])
def test_oracle_residual_stack_vs_reference_block_code(norm, masking):
    """ResidualBlockStack / ResidualBlock.call executed from the reference's source (two blocks, k5, dilation 3; BatchNorm with
    return_nmd, MaskedDYT; masking on and off) on the NumPy stand-in, whose `Layer.__call__` restates Keras 3's mask rules
    (tests/golden/tf_standin.py): oracle.forward.residual_stack gives the same block output (1e-12), the same NMD vector and
    the same outgoing mask -- conv2's, not the block input's."""
    from pathlib import Path
    z = np.load(Path(__file__).resolve().parent / "golden" / "v2_layers.npz")
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)   # noqa: E731
    tag = f"block_{norm}_{masking}"
    blocks = []
    for bi in range(2):
        blk = {}
        for cname in ("conv1", "conv2"):
            blk[cname] = {"kernel": z[f"{tag}_b{bi}_{cname}_kernel"], "bias": z[f"{tag}_b{bi}_{cname}_bias"]}
        for nname in ("bn1", "bn2"):
            if norm == "bn":
                blk[nname] = {"gamma": z[f"{tag}_b{bi}_{nname}_gamma"], "beta": z[f"{tag}_b{bi}_{nname}_beta"],
                              "mean": z[f"{tag}_b{bi}_{nname}_moving_mean"], "var": z[f"{tag}_b{bi}_{nname}_moving_variance"]}
            elif norm == "ln":
                blk[nname] = {a: z[f"{tag}_b{bi}_{nname}_{a}"] for a in ("gamma", "beta")}
            else:
                blk[nname] = {a: z[f"{tag}_b{bi}_{nname}_{a}"] for a in ("alpha", "gamma", "beta")}
        blocks.append(blk)
    cfg = {"use_masking": bool(masking), "dilation": 3, "activation": "gelu", "return_nmd": norm == "bn"}
    mask = T(z["block_mask"].astype(np.float64)) if masking else None
    y, m_out, nmd = fwd.residual_stack(T(z["block_x"]), mask, blocks, cfg, torch.float64)
    assert np.abs(y.numpy() - z[tag + "_y"]).max() < 1e-12
    if norm == "bn":
        assert np.abs(nmd.numpy() - z[tag + "_nmd"]).max() < 1e-12
    if masking:
        assert np.array_equal(m_out.numpy() > 0, z[tag + "_outmask"])
        assert z[tag + "_outmask"].sum() > z["block_mask"].sum()            # two `any` convolutions validate rows next to valid ones
    else:
        assert m_out is None and z[tag + "_outmask"].size == 0


def v2_model_case(case: int):
    """(spec, weights, tokens, golden features, golden NMD) of case `case` of tests/golden/v2_model.npz: the weights the stand-in's
    seeded provider handed to the reference's layers, re-assembled in creation order into the nested weights dict."""
    from pathlib import Path
    from jaeger_b200.modelspec import parse_project
    z = np.load(Path(__file__).resolve().parent / "golden" / "v2_model.npz")
    tag = f"m{case}"
    norm, masking, pooling = z[tag + "_cfg"].tolist()
    nname = "masked_batchnorm" if norm == "bn" else "masked_dyt"
    tail = [{"name": "nmd"}, {"name": nname, "config": {}}, {"name": "activation", "config": {"activation": "gelu"}}]
    block = {"name": "residual_block", "config": {"block_size": 2, "filters": 16, "kernel_size": 5, "dilation_rate": 3, "use_bias": True,
                                                  **({"norm_type": "masked_dyt"} if norm == "dyt" else {})}}
    hidden = [{"name": "masked_conv1d", "config": {"filters": 16, "kernel_size": 7, "use_bias": True, "activation": None}}] + tail + [block] + tail + [block] + tail
    if norm == "strided":          # make_v2_model_goldens.py:strided_cfg -- bypass / stride-2 / return_nmd blocks, masking off
        def blk(**kw):
            return {"name": "residual_block", "config": {"use_bias": True, "activation": "gelu", **kw}}
        tail = [{"name": "nmd"}, {"name": "masked_batchnorm", "config": {}}, {"name": "activation", "config": {"activation": "gelu"}}]
        hidden = [{"name": "masked_conv1d", "config": {"filters": 16, "kernel_size": 7, "use_bias": True, "activation": None}}] + tail + [
            blk(use_1x1conv=True, block_size=1, filters=16, kernel_size=5, strides=1, dilation_rate=1),
            blk(use_1x1conv=False, block_size=1, filters=24, kernel_size=5, strides=2, dilation_rate=1),
            blk(use_1x1conv=False, block_size=2, filters=24, kernel_size=3, strides=1, dilation_rate=2, return_nmd=True)]
    cfg = {"model": {"name": "m", "use_masking": bool(int(masking)), "class_label_map": [{"class": c, "label": i} for i, c in enumerate("abc")],
                     "embedding": {"use_embedding_layer": True, "input_type": "translated", "input_shape": [6, None], "embedding_size": 12},
                     "string_processor": {"codon": "CODON", "codon_id": "CODON_ID"},
                     "representation_learner": {"hidden_layers": hidden, "pooling": pooling},
                     "classifier": {"hidden_layers": [{"name": "dense", "config": {"units": 3, "activation": None}}]}}}
    spec = parse_project(cfg)
    names = z[tag + "_wnames"].tolist()
    pos = [0]

    def take(*expected):
        out = {}
        for e in expected:
            assert names[pos[0]].endswith("/" + e), (names[pos[0]], e)
            out[e] = z[f"{tag}_w{pos[0]:03d}_{e}"]
            pos[0] += 1
        return out

    def norm_w():
        if norm == "dyt":
            return take("alpha", "gamma", "beta")
        w = take("gamma", "beta", "moving_mean", "moving_variance")
        return {"gamma": w["gamma"], "beta": w["beta"], "mean": w["moving_mean"], "var": w["moving_variance"]}

    layers = []
    for layer in spec.layers:
        if layer.kind == "conv":
            layers.append(take("kernel", "bias"))
        elif layer.kind == "nmd":
            layers.append(take("moving_mean"))
        elif layer.kind == "norm":
            layers.append(norm_w())
        elif layer.kind == "resblock":
            blocks = []
            from jaeger_b200.modelspec import block_has_bypass
            for bi in range(layer.cfg["block_size"]):
                c1 = take("kernel", "bias"); b1 = norm_w(); c2 = take("kernel", "bias"); b2 = norm_w()
                blocks.append({"conv1": c1, "bn1": b1, "conv2": c2, "bn2": b2})
                if block_has_bypass(layer.cfg, bi):            # created (first called) after bn2: layers.py:1903-1909
                    blocks[-1]["conv3"], blocks[-1]["bn3"] = take("kernel", "bias"), norm_w()
            layers.append({"blocks": blocks})
        else:
            layers.append({})
    assert pos[0] == len(names)
    weights = {"embedding": z["embedding_table"], "layers": layers,
               "classifier": [{"kernel": np.zeros((z[tag + "_feat"].shape[1], 3)), "bias": np.zeros(3)}]}
    return spec, weights, z["tokens"], z[tag + "_feat"], z[tag + "_nmd"]


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_oracle_forward_vs_reference_builder_code(case):
    """tests/golden/v2_model.npz: the reference's `DynamicModelBuilder._build_block` (nnlib/builder.py:982-1193) run eagerly on the
    NumPy stand-in -- the reference's layer order, config hand-over, NMD collection / concatenation and pooling over the
    reference's layer classes.  oracle.forward.forward on the same tokens and weights returns the same pooled features and NMD
    vector (BatchNorm + max pooling, MaskedDYT + average pooling, masking off), to float32 rounding."""
    spec, weights, tokens, feat, nmd = v2_model_case(case)
    ref = fwd.forward(spec, weights, tokens, dtype=torch.float64)
    # forward() hands back float32 arrays: agreement to float32 rounding of the float64 computation
    assert np.allclose(ref["embedding"], feat, rtol=3e-7, atol=1e-6), np.abs(ref["embedding"] - feat).max()
    assert np.allclose(ref["nmd"], nmd, rtol=3e-7, atol=1e-6), np.abs(ref["nmd"] - nmd).max()
