"""CPU tests of the host logic: project.yaml parsing, the fused-launch plan compiler (against
the un-fused oracle), the C-ABI surface and the host window planner."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest
import torch
import yaml

from jaeger_b200 import _cabi
from jaeger_b200.engine import B200Engine
from jaeger_b200.modelspec import count_params, init_random, parse_project, standin_1p4m_config, string_processor_config
from jaeger_b200.plan import compile_plan, to_ctypes
from oracle import forward as ofw
from tests.plan_interp import run_plan

ROOT = Path(__file__).resolve().parent.parent
REF_CFG = Path("/root/reference/train_config")


def small_config(pooling="max", use_masking=True, reliability=True):
    cfg = standin_1p4m_config()
    m = cfg["model"]
    m["use_masking"] = use_masking
    m["representation_learner"]["pooling"] = pooling
    m["representation_learner"]["hidden_layers"] = m["representation_learner"]["hidden_layers"][:12]   # stem + 2 stacks
    if reliability:
        m["reliability_model"]["input_shape"] = 384
    else:
        del m["reliability_model"]
    return cfg


def test_standin_matches_survey_parameter_count():
    spec = parse_project(standin_1p4m_config())
    w = init_random(spec, 0)
    assert count_params(spec, w) == 1_447_296                       # SURVEY.md 8d config 2 ("1.4M")
    plan = compile_plan(spec, w)
    assert len(plan.launches) == 17 and plan.n_taps == 5
    assert plan.flops_per_window(665, algorithmic_stem_cin=128) == pytest.approx(11.27e9, rel=1e-3)


@pytest.mark.skipif(not REF_CFG.exists(), reason="reference checkout not mounted")
def test_reference_train_config_parses():
    cfg = yaml.safe_load((REF_CFG / "nn_config_1500bp_nmd_merge_6_class_brain.yaml").read_text())
    spec = parse_project(cfg)
    sp = string_processor_config(spec)
    assert sp["seq_onehot"] is False and sp["codon_depth"] == 1 and sp["vocab_size"] == 65 and sp["ngram_width"] == 3
    w = init_random(spec, 0)
    plan = compile_plan(spec, w)
    assert len(plan.launches) == 13 and plan.n_taps == 4           # stem + 3 stacks of 2 blocks
    assert plan.flops_per_window(665, algorithmic_stem_cin=128) == pytest.approx(8.68e9, rel=2e-3)


@pytest.mark.skipif(not REF_CFG.exists(), reason="reference checkout not mounted")
@pytest.mark.parametrize("name,launches,taps,params", [("nn_config_1500bp_nmd_merge_6_class_brain", 13, 4, 1_116_416),
                                                       ("nn_config_1500bp_nmd_merge_6_class_zeus", 13, 4, 1_177_684),
                                                       ("nn_config_500bp_baseline", 5, 0, 31_712), ("nn_config_500bp_nmd_merge", 5, 2, None)])
def test_every_residual_cnn_train_config_of_the_reference_compiles(name, launches, taps, params):
    """All plain-YAML residual-CNN configs the reference ships (the others are attention / hyena models or Jinja
    templates of the strided pyramid, outside the supported family) parse, initialise and compile to a launch plan."""
    spec = parse_project(yaml.safe_load((REF_CFG / f"{name}.yaml").read_text()))
    w = init_random(spec, 0)
    plan = compile_plan(spec, w)
    assert len(plan.launches) == launches and plan.n_taps == taps
    if params is not None:
        assert count_params(spec, w) == params


def _tokens(seed, b, lc, n_frac=0.02, pad_from=None):
    rng = np.random.default_rng(seed)
    t = rng.integers(1, 65, size=(b, 6, lc)).astype(np.uint8)
    t[rng.random(t.shape) < n_frac] = 0
    t[0, :, 40:75] = 0                     # a long unknown run (survives several 'any' dilations)
    if pad_from is not None:
        t[1, :, pad_from:] = 0             # right padding like a short contig in a padded batch
    return t


@pytest.mark.parametrize("pooling,masking,rel,dyt", [("max", True, True, False), ("average", True, False, False), ("max", False, True, False),
                                                     ("max", True, True, True), ("average", False, False, True)])
def test_plan_equals_unfused_oracle(pooling, masking, rel, dyt):
    from tests.helpers import to_dyt
    cfg = small_config(pooling, masking, rel)
    spec = parse_project(to_dyt(cfg) if dyt else cfg)
    w = init_random(spec, 3)
    # non-trivial biases / BN betas so the masked-row constants matter
    rng = np.random.default_rng(1)
    for lw in w["layers"]:
        for part in ([lw] if "blocks" not in lw else [p for b in lw["blocks"] for p in b.values()]):
            for k in ("bias", "beta"):
                if k in part:
                    part[k] = rng.normal(0, 0.3, part[k].shape).astype(np.float32)
    tok = _tokens(0, 3, 120, pad_from=70)
    ref = ofw.forward(spec, w, tok, dtype=torch.float64)
    got = run_plan(compile_plan(spec, w), tok)
    for k in ref:
        assert np.allclose(ref[k], got[k], rtol=1e-5, atol=2e-6), (k, np.abs(ref[k] - got[k]).max())


def return_nmd_config(masking=True):
    """The NMD vectors taken through `return_nmd: true` (train_config/nn_config_baseline.yaml:205 style) instead of
    stand-alone nmd layers: on the stem's masked_batchnorm, on a residual stack, and on a norm after an activation."""
    cfg = small_config("max", masking, True)
    hl = [l for l in cfg["model"]["representation_learner"]["hidden_layers"] if l["name"] != "nmd"]
    hl = [dict(l, config=dict(l.get("config") or {})) for l in hl]
    assert [l["name"] for l in hl] == ["masked_conv1d", "masked_batchnorm", "activation", "residual_block", "masked_batchnorm",
                                       "activation", "residual_block", "masked_batchnorm", "activation"]
    hl[1]["config"]["return_nmd"] = True          # NMD of the stem conv output
    hl[3]["config"]["return_nmd"] = True          # NMD of the last block's conv2 output (before bn2)
    hl[7]["config"]["return_nmd"] = True          # NMD of a block output (input of the stand-alone norm)
    cfg["model"]["representation_learner"]["hidden_layers"] = hl
    return cfg


@pytest.mark.parametrize("masking", [True, False])
def test_return_nmd_plan_equals_unfused_oracle(masking):
    spec = parse_project(return_nmd_config(masking))
    w = init_random(spec, 5)
    rng = np.random.default_rng(2)
    for lw in w["layers"]:
        for part in ([lw] if "blocks" not in lw else [p for b in lw["blocks"] for p in b.values()]):
            for k in ("bias", "beta"):
                if k in part:
                    part[k] = rng.normal(0, 0.3, part[k].shape).astype(np.float32)
    plan = compile_plan(spec, w)
    assert plan.n_taps == 3 and [c.tap_mode for c in plan.launches] == [1, 0, 0, 0, 1, 0, 0, 0, 2]
    tok = _tokens(4, 3, 120, pad_from=70)
    ref = ofw.forward(spec, w, tok, dtype=torch.float64)
    got = run_plan(plan, tok)
    assert ref["nmd"].shape == (3, 384)
    for k in ref:
        assert np.allclose(ref[k], got[k], rtol=1e-5, atol=2e-6), (k, np.abs(ref[k] - got[k]).max())


def signals_config(signals=None):
    """reliability_model.mode: nmd_plus_signals (the optional block of nn_config_1500bp_nmd_merge_6_class_zeus.yaml:158-165)."""
    cfg = small_config()
    rm = cfg["model"]["reliability_model"]
    rm["mode"] = "nmd_plus_signals"
    rm.pop("input_shape", None)
    if signals is not None:
        rm["signals"] = signals
    return cfg


@pytest.mark.parametrize("signals", [None, ["margin", "energy"], ["nmd_norm"]])
def test_nmd_plus_signals_plan_equals_unfused_oracle(signals):
    from jaeger_b200.plan import to_ctypes
    spec = parse_project(signals_config(signals))
    names = signals or ["max_prob", "entropy", "energy", "margin", "nmd_norm"]
    assert spec.reliability_signals == names
    w = init_random(spec, 9)
    assert w["reliability"][0]["kernel"].shape == (384 + len(names), 8)
    w["classifier"][0]["kernel"] *= 8.0                      # logits far enough apart for the softmax signals to matter
    plan = compile_plan(spec, w)
    _, head = to_ctypes(plan)
    assert head.reserved[0] & 7 == len(names)
    ids = ["max_prob", "entropy", "energy", "margin", "nmd_norm"]
    assert [ids[((head.reserved[0] >> (3 * (i + 1))) & 7) - 1] for i in range(len(names))] == names
    tok = _tokens(2, 3, 120, pad_from=70)
    ref = ofw.forward(spec, w, tok, dtype=torch.float64)
    got = run_plan(plan, tok)
    for k in ref:
        assert np.allclose(ref[k], got[k], rtol=1e-5, atol=2e-6), (k, np.abs(ref[k] - got[k]).max())
    with pytest.raises(ValueError, match="Unsupported signal"):
        parse_project(signals_config(["max_prob", "bogus"]))


def test_ood_signal_known_answers():
    """OODSignalLayer (layers.py:1632-1666) on hand-checkable logits: uniform logits give max_prob 1/n, entropy ln n,
    energy z + ln n, margin 0; a dominant logit gives max_prob ~ 1, entropy ~ 0, margin ~ 1; nmd_norm is the L2 norm."""
    z = torch.tensor([[2.0] * 6, [30.0, 0, 0, 0, 0, 0]], dtype=torch.float64)
    nmd = torch.tensor([[3.0, 4.0], [0.0, 0.0]], dtype=torch.float64)
    s = ofw.ood_signals(z, nmd, ["max_prob", "entropy", "energy", "margin", "nmd_norm"]).numpy()
    assert np.allclose(s[0], [1 / 6, np.log(6), 2 + np.log(6), 0.0, 5.0], atol=1e-12)
    assert np.allclose(s[1], [1.0, 0.0, 30.0, 1.0, 0.0], atol=1e-7)          # entropy: 5 classes clamped at eps = 1e-10


def nmd_merge_500bp_config():
    """train_config/nn_config_500bp_nmd_merge.yaml restated (the reference checkout is not on the GPU box): E64 ->
    conv(32, k7) -> BN -> GELU -> nmd -> 2 residual blocks (32, k3) -> BN -> GELU -> nmd, average pooling, 3 classes,
    reliability head on the 2 x 32 NMD values.  Exercises taps on 32-channel layers (padded to 64 on the device) and a tap
    on a launch's final output (after the stand-alone norm and its activation)."""
    def conv(f, k):
        return {"name": "masked_conv1d", "config": {"filters": f, "kernel_size": k, "strides": 1, "dilation_rate": 1,
                                                    "use_bias": True, "activation": None}}
    tail = [{"name": "masked_batchnorm", "config": {"return_nmd": False}}, {"name": "activation", "config": {"activation": "gelu"}},
            {"name": "nmd", "config": {}}]
    hidden = [conv(32, 7)] + tail + [{"name": "residual_block", "config": {"block_size": 2, "filters": 32, "kernel_size": 3,
                                                                             "use_bias": True}}] + tail
    classes = ["chromosome", "virus", "plasmid"]
    return {"model": {
        "name": "jaeger_500bp_nmd_merge", "activation": "gelu", "classifier_out_dim": 3,
        "class_label_map": [{"class": c, "label": i} for i, c in enumerate(classes)],
        "embedding": {"use_embedding_layer": True, "input_type": "translated", "frames": 6, "input_shape": [6, None], "embedding_size": 64},
        "string_processor": {"seq_onehot": False, "codon": "CODON", "codon_id": "CODON_ID", "masking": False},
        "representation_learner": {"hidden_layers": hidden, "pooling": "average"},
        "classifier": {"input_shape": 32, "hidden_layers": [{"name": "dense", "config": {"units": 3, "activation": None, "use_bias": True}}]},
        "reliability_model": {"input_shape": 64, "hidden_layers": [
            {"name": "dense", "config": {"units": 8, "activation": "gelu", "use_bias": True}}, {"name": "dropout", "config": {"rate": 0.5}},
            {"name": "dense", "config": {"units": 1, "activation": None, "use_bias": True}}]}}}


def test_500bp_nmd_merge_config_taps_on_narrow_layers():
    spec = parse_project(nmd_merge_500bp_config())
    if REF_CFG.exists():                                     # the restated config has the reference file's layer list
        ref_spec = parse_project(yaml.safe_load((REF_CFG / "nn_config_500bp_nmd_merge.yaml").read_text()))
        assert [(l.kind, l.cfg) for l in ref_spec.layers] == [(l.kind, l.cfg) for l in spec.layers]
        assert ref_spec.pooling == spec.pooling and ref_spec.n_classes == spec.n_classes
    w = init_random(spec, 0)
    rng = np.random.default_rng(1)
    for lw in w["layers"]:
        for part in ([lw] if "blocks" not in lw else [p for b in lw["blocks"] for p in b.values()]):
            for k in ("bias", "beta"):
                if k in part:
                    part[k] = rng.normal(0, 0.3, part[k].shape).astype(np.float32)
    plan = compile_plan(spec, w)
    assert [c.tap_mode for c in plan.launches] == [2, 0, 0, 0, 3] and plan.tap_width == 64 and plan.rel[0].shape == (128, 8)
    assert plan.nmd_cols.tolist() == list(range(32)) + list(range(64, 96))
    tok = _tokens(5, 3, 165, pad_from=100)
    ref = ofw.forward(spec, w, tok, dtype=torch.float64)
    got = run_plan(plan, tok)
    assert ref["nmd"].shape == (3, 64) and ref["embedding"].shape == (3, 32)
    for k in ref:
        assert np.allclose(ref[k], got[k], rtol=1e-5, atol=2e-6), (k, np.abs(ref[k] - got[k]).max())


def test_unsupported_layers_fail_loudly():
    cfg = small_config()
    cfg["model"]["representation_learner"]["hidden_layers"].insert(1, {"name": "masked_bilstm", "config": {"units": 8}})
    with pytest.raises(NotImplementedError):
        parse_project(cfg)
    cfg = small_config()
    cfg["model"]["representation_learner"]["hidden_layers"][0]["config"]["mask_mode"] = "strict"
    spec = parse_project(cfg)             # the mode itself is supported; residual blocks behind it are not (plan.py)
    with pytest.raises(NotImplementedError, match="un-masked tensor"):
        compile_plan(spec, init_random(spec, 0))


def test_cabi_exports_every_declared_symbol():
    header = (ROOT / "include" / "jaeger_b200.h").read_text()
    declared = set(re.findall(r"\b(jg_[a-z0-9_]+)\s*\(", header))
    declared -= {"jg_ctx", "jg_model"}
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(str(_cabi.LIB_PATH))
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_cabi.EXPORTED), declared ^ set(_cabi.EXPORTED)
    assert lib.jg_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(_cabi.JaegerB200Error):
        _cabi.Context(0)
    with pytest.raises(_cabi.JaegerB200Error):
        B200Engine(spec=parse_project(small_config()))


def test_host_window_planner_matches_reference_goldens():
    cases = json.loads((ROOT / "tests" / "golden" / "window_indices.json").read_text())
    for c in cases:
        contig, start, nb, ordinal, last = B200Engine.plan_windows(
            np.array([c["seqlen"]]), c["fsize"], c["stride"], c["dyn"], c["thr"])
        assert start.tolist() == c["idx"], c
        assert ordinal.tolist() == list(range(len(c["idx"]))) and last.tolist() == [0] * (len(c["idx"]) - 1) + [1]
        assert (nb == c["fsize"]).all()


def test_host_window_planner_two_pass_and_ragged():
    lens = np.array([5000, 1999, 0, 650, 2000, 137, 100000], dtype=np.int64)
    contig, start, nb, ordinal, last = B200Engine.plan_windows(lens, 2000, 1500)
    assert contig.tolist() == [0, 0, 0, 4] + [6] * 66                   # range(0, L-1999, 1500)
    assert start[:4].tolist() == [0, 1500, 3000, 0]
    c2, s2, nb2, o2, l2 = B200Engine.plan_windows(lens, 2000, 1500, min_len=500, max_len=1999, short_pass=True)
    assert c2.tolist() == [1, 3] and nb2.tolist() == [1999, 650] and l2.tolist() == [1, 1] and s2.tolist() == [0, 0]
    # empty input
    c3, *_ = B200Engine.plan_windows(np.zeros(0, dtype=np.int64), 2000, 1500)
    assert len(c3) == 0


@pytest.mark.skipif(not REF_CFG.exists(), reason="reference checkout not mounted")
def test_500bp_baseline_config_with_narrow_layers_is_padded_correctly():
    """BASELINE config 3 (nn_config_500bp_baseline.yaml: 32 filters): channels are zero-padded to
    64 for the tensor-core kernel; the padded plan must still equal the un-fused oracle."""
    cfg = yaml.safe_load((REF_CFG / "nn_config_500bp_baseline.yaml").read_text())
    spec = parse_project(cfg)
    w = init_random(spec, 2)
    plan = compile_plan(spec, w)
    assert [c.kernel.shape for c in plan.launches] == [(7, 64, 64)] + [(3, 64, 64)] * 4
    assert plan.feat_dim == 64 and plan.real_feat_dim == 32 and plan.pool_mode == 2 and plan.rel is None
    tok = _tokens(4, 3, 165)
    ref = ofw.forward(spec, w, tok, dtype=torch.float64)
    got = run_plan(plan, tok)
    for k in ref:
        assert np.allclose(ref[k], got[k], rtol=1e-5, atol=2e-6), k


def _assert_same_weights(a, b, path=""):
    if isinstance(a, dict):
        assert set(a) == set(b), (path, set(a), set(b))
        for k in a:
            _assert_same_weights(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, list):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _assert_same_weights(x, y, f"{path}[{i}]")
    elif a is None:
        assert b is None, path
    else:
        assert np.array_equal(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)), path


@pytest.mark.parametrize("variant", ["standin", "dyt", "onehot_dense", "return_nmd"])
def test_saved_model_bundle_maps_onto_the_layer_list(tmp_path, variant):
    """`B200Engine(path_dict)` without an exported npz: the SavedModel `variables/` bundle is read by the pure-Python
    reader and mapped onto the project's layer list by graph order, attribute name and shape (a bundle written by
    tests/tf_bundle_writer.py with the key naming a Keras 3 export uses -- the naming verified on the reference's own
    data/models/test bundle; a modern checkpoint is not vendored).  Misaligned bundles are refused."""
    from jaeger_b200.weights import load_saved_model_weights, read_tf_bundle, weights_from_bundle
    from tests.helpers import to_dyt
    from tests.tf_bundle_writer import keras3_export_names, write_bundle
    cfg = {"standin": standin_1p4m_config, "dyt": lambda: to_dyt(standin_1p4m_config()), "onehot_dense": standin_1p4m_config,
           "return_nmd": return_nmd_config}[variant]()
    if variant == "onehot_dense":
        cfg["model"]["embedding"].update(use_embedding_layer=False, input_shape=[6, None, 64])
        cfg["model"]["representation_learner"]["hidden_layers"][0]["config"]["use_bias"] = False
    spec = parse_project(cfg)
    w = init_random(spec, 11)
    graph = tmp_path / "model" / "jaeger_x_1M_fragment_graph"
    write_bundle(graph / "variables", keras3_export_names(spec, w))
    tensors = read_tf_bundle(graph / "variables")
    assert len(tensors) == len(keras3_export_names(spec, w))
    project = tmp_path / "model" / "jaeger_x_1M_fragment_project.yaml"
    project.write_text(yaml.safe_dump(cfg))
    got = load_saved_model_weights({"graph": graph, "project": project}, spec)
    _assert_same_weights(w, got)
    ref = ofw.forward(spec, w, _tokens(1, 2, 90))
    again = ofw.forward(spec, got, _tokens(1, 2, 90))
    assert all(np.array_equal(ref[k], again[k]) for k in ref)
    # a bundle of a different architecture is refused, not guessed
    other = parse_project(small_config())
    with pytest.raises(ValueError, match="does not match"):
        weights_from_bundle(other, tensors)


def test_crop_resolution_and_fsize_warning_match_the_reference():
    """seqops/crop.py (resolve_crop, the 3 * codons + 5 rule) and commands/predict.py:36-64 (_crop_length_warning):
    the known answers of the reference's tests/unit/test_crop.py / test_predict_crop_warning.py, and -- when the checkout
    is mounted -- the reference's own pure-Python functions over a sweep."""
    import sys
    from jaeger_b200.modelspec import crop_length_warning, resolve_crop
    assert resolve_crop({"crop_size": 665}) == (665, 2000) and resolve_crop({"crop_size": 2000, "crop_units": "nucleotide"}) == (665, 2000)
    assert resolve_crop({"crop_size": 498}) == (498, 1499) and resolve_crop({"crop_size": 1500, "crop_units": "nucleotide"}) == (498, 1500)
    for bad in ({}, {"crop_size": 0}, {"crop_size": "665"}, {"crop_size": 10, "crop_units": "bp"}):
        with pytest.raises(ValueError):
            resolve_crop(bad)
    assert crop_length_warning(665, 2000, 2000) is None and crop_length_warning(None, None, 1234) is None
    assert "498 codon frames" in crop_length_warning(665, 2000, 1500) and "prefer --fsize 2000" in crop_length_warning(665, 2000, 1500)
    assert crop_length_warning(None, 2000, 2000) is None and "differs from the model's trained fragment length (2000 nt)" in crop_length_warning(None, 2000, 1500)
    cfg = standin_1p4m_config()
    cfg["model"]["string_processor"]["crop_size"] = 665
    sp = string_processor_config(parse_project(cfg))
    assert (sp["crop_size_codons"], sp["crop_size_nt"], sp["crop_units"]) == (665, 2000, "codon")
    ref_src = Path("/root/reference/src")
    if not ref_src.exists():
        return
    sys.path.insert(0, str(ref_src))
    try:
        from jaeger.seqops import crop as rcrop
    finally:
        sys.path.remove(str(ref_src))
    for size in (1, 5, 6, 165, 498, 665, 681, 2000, 2048):
        for units in ("codon", "nucleotide"):
            if units == "nucleotide" and size < 6:
                continue
            assert resolve_crop({"crop_size": size, "crop_units": units}) == rcrop.resolve_crop({"crop_size": size, "crop_units": units})


def test_host_window_planner_fuzz_vs_reference_window_indices():
    """Property test (hypothesis): the native planner equals `_window_indices` (seqops/io.py:38-71) -- the reference's own
    function when the checkout is mounted, else the oracle restatement -- for arbitrary lengths, window sizes, strides
    and dynamic-stride thresholds, including the banker's-rounding ties of `round(i * (L - fsize) / (n - 1))`."""
    import sys
    import types
    from hypothesis import given, settings, strategies as st
    from oracle import seqwin
    ref_fn = None
    if Path("/root/reference/src").exists():
        sys.path.insert(0, "/root/reference/src")
        saved = sys.modules.get("pyfastx")
        sys.modules.setdefault("pyfastx", types.ModuleType("pyfastx"))
        sys.modules.setdefault("pydustmasker", types.ModuleType("pydustmasker"))
        try:
            from jaeger.seqops import io as rio
            ref_fn = rio._window_indices
        except Exception:
            ref_fn = None
        finally:
            sys.path.remove("/root/reference/src")
            if saved is None:
                sys.modules.pop("pyfastx", None)
        assert ref_fn is not None, "the reference's seqops.io did not import with the pyfastx / pydustmasker stubs"

    @settings(max_examples=400, deadline=None)
    @given(st.integers(1, 4096), st.integers(0, 60000), st.integers(1, 5000), st.booleans(), st.sampled_from([1.0, 1.5, 3.0, 10.0, 25.0]))
    def check(fsize, extra, stride, dyn, thr):
        seqlen = fsize + extra
        want = (ref_fn or seqwin.window_indices)(seqlen, fsize, stride, dyn, thr)
        assert seqwin.window_indices(seqlen, fsize, stride, dyn, thr) == list(want)
        _, start, nb, ordinal, last = B200Engine.plan_windows(np.array([seqlen]), fsize, stride, dyn, thr)
        assert start.tolist() == list(want), (seqlen, fsize, stride, dyn, thr)
        assert (nb == fsize).all() and last.tolist() == [0] * (len(want) - 1) + [1]

    check()


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_plan_compiler_vs_reference_builder_code(case):
    """The PRODUCT's plan compiler against the reference's own builder + layer code: tests/golden/v2_model.npz is
    `DynamicModelBuilder._build_block` run eagerly on a NumPy stand-in for TensorFlow (tests/golden/make_v2_model_goldens.py);
    the fused launch plan compiled from the same weights, executed by the float64 plan interpreter with the kernels' storage
    semantics (masked storage, masked-row constants, folded norms, embedding folded into the stem), gives the same pooled
    features and NMD vector."""
    from tests.test_oracle_layer_known_answers import v2_model_case
    spec, weights, tokens, feat, nmd = v2_model_case(case)
    got = run_plan(compile_plan(spec, weights), tokens)
    assert np.allclose(got["embedding"], feat, rtol=2e-5, atol=2e-5), np.abs(got["embedding"] - feat).max()
    assert np.allclose(got["nmd"], nmd, rtol=2e-5, atol=2e-5), np.abs(got["nmd"] - nmd).max()


def _baseline_3p4m_config():
    """train_config/nn_config_baseline.yaml as a dict (the broken quote of its data_dir line repaired), masking off: strided
    architectures are only well-defined without it."""
    import re
    import yaml
    txt = (Path("/root/reference/train_config/nn_config_baseline.yaml")).read_text()
    cfg = yaml.safe_load(re.sub(r'"/path/to/data directory""', '"x"', txt))
    cfg["model"]["use_masking"] = False
    return cfg


def small_strided_config(filters=(64, 128), bypass_first=True):
    """A small network with every new block variant: a stride-1 block with a 1x1 bypass, strided blocks (channel growth,
    k5 and k3), a block after the stride with dilation 2, return_nmd on the last block, and a two-hidden-layer classifier."""
    def blk(f, k, s=1, d=1, n=1, one=False, nmd=False):
        return {"name": "residual_block", "config": {"block_size": n, "filters": f, "kernel_size": k, "strides": s, "dilation_rate": d,
                                                      "use_1x1conv": one, "use_bias": True, "activation": "gelu", "return_nmd": nmd}}
    hidden = [{"name": "masked_conv1d", "config": {"filters": filters[0], "kernel_size": 7, "use_bias": True, "activation": None}},
              {"name": "masked_batchnorm", "config": {}}, {"name": "activation", "config": {"activation": "gelu"}},
              blk(filters[0], 5, one=bypass_first), blk(filters[1], 5, s=2), blk(filters[1], 3, d=2, n=2), blk(filters[1], 3, s=2, nmd=True)]
    return {"model": {"name": "strided_small", "activation": "gelu", "use_masking": False,
                      "class_label_map": [{"class": c, "label": i} for i, c in enumerate("abcd")],
                      "embedding": {"use_embedding_layer": True, "input_type": "translated", "input_shape": [6, None], "embedding_size": 16},
                      "string_processor": {"seq_onehot": False, "codon": "CODON", "codon_id": "CODON_ID", "masking": False},
                      "representation_learner": {"hidden_layers": hidden, "pooling": "max"},
                      "classifier": {"input_shape": filters[1], "hidden_layers": [
                          {"name": "dense", "config": {"units": 48, "activation": "gelu", "use_bias": True}},
                          {"name": "dropout", "config": {"rate": 0.5}},
                          {"name": "dense", "config": {"units": 32, "activation": "gelu", "use_bias": True}},
                          {"name": "dense", "config": {"units": 4, "activation": None, "use_bias": True}}]},
                      "reliability_model": {"hidden_layers": [{"name": "dense", "config": {"units": 8, "activation": "gelu", "use_bias": True}},
                                                              {"name": "dense", "config": {"units": 1, "activation": None, "use_bias": True}}]}}}


def flat_g1_config():
    """A first-generation (flat schema) project: the keys of commands/configs/nn_config.yaml:36-66 with the shape of the
    `jaeger_1.5M` template -- one-hot input, conv k7, three single-block stacks 64 / 128 / 256 with stride 2, final conv k5,
    max pooling, Dense(128) -> Dense(5) classifier."""
    return {"model": {
        "name": "jaeger_flat_g1", "activation": "gelu",
        "class_label_map": [{"class": c, "label": i} for i, c in enumerate(["bacteria", "phage", "archaea", "virus", "eukarya"])],
        "embedding": {"type": "translated", "strands": 2, "frames": 6, "length": None, "input_shape": [6, None, 64], "embedding_size": 4},
        "string_processor": {"codon": "CODON", "codon_id": "CODON_ID", "crop_size": 1024},
        "representation_learner": {
            "masked_conv1d_1_filters": 64, "masked_conv1d_1_kernel_size": 7, "masked_conv1d_1_strides": 1, "masked_conv1d_1_dilation_rate": 1,
            "block_sizes": [1, 1, 1], "block_filters": [64, 128, 256], "block_kernel_size": [5, 5, 5], "block_kernel_dilation": [1, 1, 1],
            "block_kernel_strides": [2, 2, 2],
            "masked_conv1d_final_kernel_size": 5, "masked_conv1d_final_strides": 1, "masked_conv1d_final_dilation_rate": 1, "pooling": "max"},
        "classifier": {"dense_1_units": 128, "classes": 5},
        "reliability_model": {"dense_1_units": 128}}}


@pytest.mark.parametrize("lc", [665, 498, 166, 165])
def test_strided_and_bypass_blocks_plan_equals_unfused_oracle(lc):
    """ResidualBlock with strides = 2 and / or a 1x1 bypass conv + norm (nnlib/v2/layers.py:1840-1864, 1903-1909), compiled to
    row-plane launches + stride-1 convs (plan.py:split_phases), against the un-fused oracle whose strided convs use TF's SAME
    padding rule -- for even and odd frame lengths (lc - 6 = 659, 492, 160, 159: the padding differs with the parity) -- and a
    classifier head with two hidden Dense layers."""
    spec = parse_project(small_strided_config())
    w = init_random(spec, 5)
    rng = np.random.default_rng(lc)
    tok = rng.integers(0, 65, (3, 6, lc)).astype(np.uint8)
    plan = compile_plan(spec, w)
    assert sum(c.kind == 4 for c in plan.launches) == 2 and any(c.kernel_odd is not None for c in plan.launches)
    got, want = run_plan(plan, tok), ofw.forward(spec, w, tok, dtype=torch.float64)
    for key in ("prediction", "embedding", "nmd", "reliability"):
        assert np.abs(got[key] - want[key]).max() < 1e-6, (key, np.abs(got[key] - want[key]).max())     # the oracle returns float32


def test_reference_baseline_3p4m_and_flat_g1_configs_compile_and_match_the_oracle():
    """The reference's own strided configurations: train_config/nn_config_baseline.yaml (3.4 M parameters: stride-2 blocks with
    bypass, dilations up to 8, 256 channels, MLP head) and the flat first-generation schema of commands/configs/nn_config.yaml
    (the `jaeger_1.5M` template of the G1 models, translated to a layer list by modelspec.flat_schema_to_layer_list)."""
    import yaml
    ref_flat = Path("/root/reference/src/jaeger/commands/configs/nn_config.yaml")
    cases = [(flat_g1_config(), 3)]
    if ref_flat.exists():
        cases += [(_baseline_3p4m_config(), 3), (yaml.safe_load(ref_flat.read_text()), 3)]
    for cfg, n_kind4 in cases:
        spec = parse_project(cfg)
        w = init_random(spec, 1)
        plan = compile_plan(spec, w)
        assert sum(c.kind == 4 for c in plan.launches) == n_kind4
        tok = np.random.default_rng(3).integers(0, 65, (2, 6, 165)).astype(np.uint8)
        got, want = run_plan(plan, tok), ofw.forward(spec, w, tok, dtype=torch.float64)
        for key in want:
            assert np.abs(got[key] - want[key]).max() < 1e-6, (spec.name, key)
    if not ref_flat.exists():
        return
    spec = parse_project(_baseline_3p4m_config())
    n_bias_free = sum(v.size for lw in init_random(spec, 0)["layers"] if "blocks" in lw for b in lw["blocks"]
                      for name in ("conv1", "conv2", "conv3") if name in b and not np.any(b[name]["bias"]) and b[name]["kernel"].shape[2] == 256
                      for v in [b[name]["bias"]])
    assert count_params(spec, init_random(spec, 0)) - n_bias_free == 3_384_352          # "3.4M"


def test_strided_blocks_need_masking_off():
    cfg = small_strided_config()
    cfg["model"]["use_masking"] = True
    with pytest.raises(ValueError, match="use_masking: false"):
        parse_project(cfg)


def layernorm_config(standalone_ln: bool, masking: bool = True, filters: int = 24, bypass: bool = False):
    """A residual CNN whose norms are MaskedLayerNormalization (nnlib/v2/layers.py:293-367): every norm of the residual blocks
    (`norm_type: masked_layernorm`, ResidualBlock._make_norm, layers.py:1826-1834) and, with `standalone_ln`, the stand-alone norm
    after the stem conv; the trailing norm after the stack stays a BatchNorm (a LayerNorm there follows an activation)."""
    conv = {"name": "masked_conv1d", "config": {"filters": filters, "kernel_size": 7, "use_bias": True, "activation": None}}
    n1 = {"name": "masked_layernorm", "config": {"epsilon": 2e-3}} if standalone_ln else {"name": "masked_batchnorm", "config": {}}
    act = {"name": "activation", "config": {"activation": "gelu"}}
    blk = {"name": "residual_block", "config": {"block_size": 2, "filters": filters, "kernel_size": 5, "dilation_rate": 2, "use_bias": True,
                                                "norm_type": "masked_layernorm", "use_1x1conv": bypass}}
    return {"model": {"name": "ln_model", "activation": "gelu", "use_masking": masking,
                      "class_label_map": [{"class": c, "label": i} for i, c in enumerate("abc")],
                      "embedding": {"use_embedding_layer": True, "input_type": "translated", "input_shape": [6, None], "embedding_size": 16},
                      "string_processor": {"codon": "CODON", "codon_id": "CODON_ID", "crop_size": 300},
                      "representation_learner": {"hidden_layers": [conv, n1, act, blk, {"name": "nmd"}, {"name": "masked_batchnorm", "config": {}}, act],
                                                 "pooling": "max"},
                      "classifier": {"hidden_layers": [{"name": "dense", "config": {"units": 3, "activation": None}}]},
                      "reliability_model": {"hidden_layers": [{"name": "dense", "config": {"units": 8, "activation": "gelu", "use_bias": True}},
                                                              {"name": "dense", "config": {"units": 1, "activation": None, "use_bias": True}}]}}}


@pytest.mark.parametrize("standalone_ln,masking,bypass", [(True, True, False), (False, True, True), (True, False, False)])
def test_masked_layernorm_plan_equals_unfused_oracle(standalone_ln, masking, bypass):
    """MaskedLayerNormalization compiled into the conv launch's first-norm slot (row statistics of acc + bias over the REAL channels,
    gamma / beta, exactly 0 at masked rows) against the un-fused oracle, whose LayerNorm is pinned on the reference's own
    `MaskedLayerNormalization.call` and `ResidualBlock.call` (tests/golden/v2_layers.npz: ln_*, block_ln_*)."""
    spec = parse_project(layernorm_config(standalone_ln, masking, bypass=bypass))
    w = init_random(spec, 4)
    plan = compile_plan(spec, w)
    assert sum(c.ln1 for c in plan.launches) == (5 if bypass else 4) + int(standalone_ln)
    rng = np.random.default_rng(8)
    tok = rng.integers(1, 65, (3, 6, 100)).astype(np.uint8)
    tok[0, :, 20:31] = 0
    tok[1, :, ::9] = 0
    got, want = run_plan(plan, tok), ofw.forward(spec, w, tok, dtype=torch.float64)
    for key in ("prediction", "embedding", "nmd", "reliability"):
        assert np.abs(got[key] - want[key]).max() < 1e-6, (key, np.abs(got[key] - want[key]).max())


def test_masked_layernorm_placements_that_are_refused():
    cfg = layernorm_config(True)
    hl = cfg["model"]["representation_learner"]["hidden_layers"]
    hl[-2] = {"name": "masked_layernorm", "config": {}}           # after the residual stack's activation: not a first norm
    spec = parse_project(cfg)
    with pytest.raises(NotImplementedError, match="right after a convolution"):
        compile_plan(spec, init_random(spec, 0))
    cfg = layernorm_config(True)
    cfg["model"]["representation_learner"]["hidden_layers"][1]["config"]["return_nmd"] = True
    with pytest.raises(ValueError, match="return_nmd"):
        parse_project(cfg)


def conv_stack_config(modes=("strict", "majority", "any"), masking=True):
    """A plain MaskedConv1D stack (no residual blocks) with a mask_mode per convolution (nnlib/v2/layers.py:1134-1146, 1245-1252):
    k7 VALID, k5 SAME dilation 2, k3 SAME, each followed by BatchNorm + GELU + an NMD tap."""
    def conv(f, k, mode, padding="valid", d=1):
        return {"name": "masked_conv1d", "config": {"filters": f, "kernel_size": k, "dilation_rate": d, "padding": padding, "use_bias": True,
                                                    "activation": None, "mask_mode": mode}}
    tail = [{"name": "masked_batchnorm", "config": {}}, {"name": "activation", "config": {"activation": "gelu"}}, {"name": "nmd"}]
    hidden = [conv(24, 7, modes[0])] + tail + [conv(24, 5, modes[1], "same", 2)] + tail + [conv(24, 3, modes[2], "same")] + tail
    return {"model": {"name": "conv_stack", "activation": "gelu", "use_masking": masking,
                      "class_label_map": [{"class": c, "label": i} for i, c in enumerate("abc")],
                      "embedding": {"use_embedding_layer": True, "input_type": "translated", "input_shape": [6, None], "embedding_size": 16},
                      "string_processor": {"codon": "CODON", "codon_id": "CODON_ID", "crop_size": 300},
                      "representation_learner": {"hidden_layers": hidden, "pooling": "average"},
                      "classifier": {"hidden_layers": [{"name": "dense", "config": {"units": 3, "activation": None}}]},
                      "reliability_model": {"hidden_layers": [{"name": "dense", "config": {"units": 8, "activation": "gelu", "use_bias": True}},
                                                              {"name": "dense", "config": {"units": 1, "activation": None, "use_bias": True}}]}}}


@pytest.mark.parametrize("modes", [("strict", "strict", "strict"), ("majority", "majority", "majority"), ("strict", "majority", "any")])
def test_mask_modes_strict_and_majority_plan_equals_unfused_oracle(modes):
    """mask_mode strict / majority on stand-alone MaskedConv1D layers: the launch plan (a tap-count threshold per launch) against the
    un-fused oracle, whose three modes are pinned on the reference's own MaskedConv1D.call (v2_layers.npz: conv_*_{any,majority,
    strict}_mask) and on tests/unit/test_mask_mode.py's known answers."""
    spec = parse_project(conv_stack_config(modes))
    w = init_random(spec, 6)
    plan = compile_plan(spec, w)
    thr = {"any": lambda k: 1, "majority": lambda k: (k + 1) // 2, "strict": lambda k: k}
    assert [c.mask_thr for c in plan.launches] == [thr[m](k) for m, k in zip(modes, (7, 5, 3))]
    rng = np.random.default_rng(5)
    tok = rng.integers(1, 65, (4, 6, 100)).astype(np.uint8)
    tok[0, :, 20:31] = 0
    tok[1, :, ::9] = 0
    tok[2, :, 97:] = 0
    got, want = run_plan(plan, tok), ofw.forward(spec, w, tok, dtype=torch.float64)
    for key in ("prediction", "embedding", "nmd", "reliability"):
        assert np.abs(got[key] - want[key]).max() < 1e-6, (key, np.abs(got[key] - want[key]).max())
    base = ofw.forward(parse_project(conv_stack_config(("any", "any", "any"))), w, tok, dtype=torch.float64)
    assert np.abs(base["embedding"] - want["embedding"]).max() > 1e-4 or modes == ("any",) * 3     # the modes do change the result


def test_residual_block_after_a_thresholded_conv_is_refused():
    cfg = conv_stack_config(("strict", "any", "any"))
    cfg["model"]["representation_learner"]["hidden_layers"].append(
        {"name": "residual_block", "config": {"block_size": 1, "filters": 24, "kernel_size": 3, "use_bias": True}})
    spec = parse_project(cfg)
    with pytest.raises(NotImplementedError, match="un-masked tensor"):
        compile_plan(spec, init_random(spec, 0))
    with pytest.raises(ValueError, match="Invalid mask_mode"):
        parse_project(conv_stack_config(("most", "any", "any")))


@pytest.mark.parametrize("variant", ["strided_bypass_mlp", "layernorm"])
def test_saved_model_bundle_with_bypass_blocks_mlp_head_and_layernorm(tmp_path, variant):
    """The bundle -> weights mapping for the round-2 layer variants: residual blocks with the conv3 / bn3 bypass (strided or
    use_1x1conv), a classifier with two hidden Dense layers (picked by chaining the widths), MaskedLayerNormalization groups
    (gamma / beta only).  Also what the reader refuses: several data shards, compressed index blocks, ambiguous Dense shapes."""
    from jaeger_b200 import weights as W
    from tests.tf_bundle_writer import keras3_export_names, write_bundle
    cfg = small_strided_config() if variant == "strided_bypass_mlp" else layernorm_config(True, bypass=True)
    spec = parse_project(cfg)
    w = init_random(spec, 13)
    graph = tmp_path / "model" / "jaeger_y_1M_fragment_graph"
    names = keras3_export_names(spec, w)
    write_bundle(graph / "variables", names)
    tensors = W.read_tf_bundle(graph / "variables")
    assert len(tensors) == len(names)
    got = W.weights_from_bundle(spec, tensors)
    _assert_same_weights({k: w[k] for k in ("layers", "classifier", "reliability") if k in w}, {k: got[k] for k in ("layers", "classifier", "reliability") if k in w})
    tok = _tokens(2, 2, 120)
    ref, again = ofw.forward(spec, w, tok), ofw.forward(spec, {**got, "embedding": w["embedding"]}, tok)
    assert all(np.array_equal(ref[k], again[k]) for k in ref)
    # two unused Dense kernels of the same shape cannot be told apart by shape: refused, not guessed
    dup = dict(tensors)
    k_cls = [k for k, v in tensors.items() if v.ndim == 2 and v.shape == np.asarray(w["classifier"][0]["kernel"]).shape][0]
    dup[k_cls.replace("_operations/", "_operations/9")] = tensors[k_cls]
    with pytest.raises(ValueError, match="ambiguous"):
        W.weights_from_bundle(spec, dup)
    # compressed index block / several data shards
    idx = bytearray((graph / "variables" / "variables.index").read_bytes())
    bad = tmp_path / "bad" / "variables"
    bad.mkdir(parents=True)
    (bad / "variables.data-00000-of-00001").write_bytes((graph / "variables" / "variables.data-00000-of-00001").read_bytes())
    footer = bytes(idx[-48:])
    pos = 0
    for _ in range(2):
        _, pos = W._varint(footer, pos)
    idx_off, pos = W._varint(footer, pos)
    idx_size, pos = W._varint(footer, pos)
    idx[idx_off + idx_size] = 1                      # compression type byte of the index block: snappy
    (bad / "variables.index").write_bytes(bytes(idx))
    with pytest.raises(ValueError, match="compressed"):
        W.read_tf_bundle(bad)
