"""project.yaml -> model specification + weights container.

Mirrors what the reference's `DynamicModelBuilder` reads from a model's `*_project.yaml`
(nnlib/builder.py:844-894 embedding, 982-1193 `_build_block`, 589-596 / 705-713 heads) for
the residual-CNN family the fragment models use, and what `InferModel` reads from it
(nnlib/inference.py:423-483).  Anything outside that family raises NotImplementedError
instead of being approximated.

The weights container is a plain nested dict of float32 NumPy arrays in TensorFlow layout
(conv kernels [k, Cin, Cout], dense kernels [in, out]):

    {"embedding": [vocab, E] | None,
     "layers": [per hidden layer: {} | {"kernel","bias"} | {"gamma","beta","mean","var"} |
                {"moving_mean"} | {"blocks": [{"conv1","bn1","conv2","bn2"}]}],
     "classifier": [{"kernel","bias"}...], "reliability": [{"kernel","bias"}...]}
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Any

import numpy as np
import yaml

CODON_TABLE_NAMES = ("CODON_ID", "AA_ID", "MURPHY10_ID", "PC5_ID")
OOD_SIGNALS = ("max_prob", "entropy", "energy", "margin", "nmd_norm")     # builder.py:645-651 default order; ids 1..5 on the device


@dataclass
class LayerSpec:
    kind: str                      # conv | nmd | norm | act | resblock
    cfg: dict[str, Any] = field(default_factory=dict)


@dataclass
class ModelSpec:
    name: str
    classes: list[dict[str, Any]]
    embedding: dict[str, Any]
    string_processor: dict[str, Any]
    layers: list[LayerSpec]
    pooling: str
    classifier: list[dict[str, Any]]      # dense layer cfgs (units, activation)
    reliability: list[dict[str, Any]] | None
    use_masking: bool = True
    reliability_signals: list[str] | None = None      # reliability_model.mode == "nmd_plus_signals" (builder.py:644-657)

    @property
    def n_classes(self) -> int:
        return self.classifier[-1]["units"]

    @property
    def embedding_size(self) -> int:
        return int(self.embedding.get("embedding_size", 4))

    @property
    def uses_token_input(self) -> bool:
        """True when the SavedModel input is [B,6,L] token ids (Embedding, mask_zero)."""
        return bool(self.embedding.get("use_embedding_layer", False))


def _act_name(cfg: dict[str, Any], default: str | None = None) -> str | None:
    a = cfg.get("activation", default)
    if a in (None, "linear"):
        return None
    a = str(a).lower()
    if a not in ("gelu", "relu"):
        raise NotImplementedError(f"activation {a!r} is not on the supported hot path")
    return a


def parse_project(cfg: dict[str, Any]) -> ModelSpec:
    """Compile the `model:` section of a project.yaml (layer-list schema)."""
    model = cfg["model"] if "model" in cfg else cfg
    emb = dict(model.get("embedding", {}))
    sp = dict(model.get("string_processor", {}))
    if emb.get("input_type", emb.get("type", "translated")) != "translated":
        raise NotImplementedError("only the translated (six-frame codon) input is on the hot path")
    rep = model.get("representation_learner")
    if rep is not None and "hidden_layers" not in rep and "block_sizes" in rep:
        model = flat_schema_to_layer_list(model)              # first-generation project files
        emb = dict(model.get("embedding", {}))
        rep = model["representation_learner"]
    if rep is None or "hidden_layers" not in rep:
        raise NotImplementedError("project.yaml has no representation_learner.hidden_layers (and is not the flat first-generation schema)")
    use_masking = bool(model.get("use_masking", True))
    layers: list[LayerSpec] = []
    for lc in rep["hidden_layers"]:
        name = str(lc.get("name", "")).lower()
        c = dict(lc.get("config") or {})
        if name == "masked_conv1d":
            if int(c.get("strides", 1)) != 1:
                raise NotImplementedError("strided convolutions are not supported")
            mode = str(c.get("mask_mode", "any")).lower()
            if mode not in ("any", "majority", "strict"):
                raise ValueError(f"Invalid mask_mode: {mode!r} (use 'any', 'majority' or 'strict')")       # layers.py:1170-1173
            layers.append(LayerSpec("conv", dict(
                mask_mode=mode,
                filters=int(c["filters"]), kernel_size=int(c["kernel_size"]),
                dilation=int(c.get("dilation_rate", 1)), padding=str(c.get("padding", "valid")).lower(),
                use_bias=bool(c.get("use_bias", True)), activation=_act_name(c),
                use_masking=bool(c.get("use_masking", use_masking)))))
        elif name == "nmd":
            layers.append(LayerSpec("nmd"))
        elif name == "masked_batchnorm":
            # return_nmd (layers.py:943-954): besides the normalised tensor the layer returns the masked per-example
            # channel mean of its INPUT minus its own moving mean -- an NMD tap in front of the norm
            eps = float(c.get("epsilon", 1e-5))
            if c.get("return_nmd") and eps != 1e-5:
                raise NotImplementedError("masked_batchnorm(return_nmd=True) with epsilon != 1e-5 (the tap's count epsilon)")
            layers.append(LayerSpec("norm", dict(type="bn", epsilon=eps, return_nmd=bool(c.get("return_nmd", False)))))
        elif name == "masked_dyt":          # nnlib/v2/layers.py:385-444: gamma * tanh(alpha * x) + beta, re-masked
            if c.get("return_nmd"):
                raise NotImplementedError("masked_dyt(return_nmd=True) is rejected by the reference as well")
            layers.append(LayerSpec("norm", dict(type="dyt", alpha_init=float(c.get("alpha_init", 0.5)))))
        elif name == "masked_layernorm":    # nnlib/v2/layers.py:293-367: per-position normalisation over the channels, re-masked
            if c.get("return_nmd"):
                raise ValueError("return_nmd=True is not supported by MaskedLayerNormalization (the reference raises too, layers.py:307-311)")
            layers.append(LayerSpec("norm", dict(type="ln", epsilon=float(c.get("epsilon", 1e-3)), center=bool(c.get("center", True)),
                                                 scale=bool(c.get("scale", True)))))
        elif name in ("activation", "gelu", "relu"):
            layers.append(LayerSpec("act", dict(activation=_act_name(c, name if name != "activation" else None))))
        elif name == "residual_block":
            strides = int(c.get("strides", 1))
            if strides not in (1, 2):
                raise NotImplementedError(f"residual_block strides={strides}: only 1 and 2 are supported")
            if strides > 1 and bool(c.get("use_masking", use_masking)):
                # layers.py:1884-1888: the block forwards the pre-stride mask, which has the wrong length after a strided
                # conv -- strided architectures are only well-defined without masking
                raise ValueError("residual_block with strides > 1 needs model.use_masking: false (the reference forwards a mask "
                                 "of the pre-stride length past a strided block)")
            if strides > 1 and int(c.get("dilation_rate", 1)) != 1:
                raise NotImplementedError("a strided residual block with dilation_rate > 1 (tf.nn.conv1d refuses the combination too)")
            norm_type = str(c.get("norm_type", "masked_batchnorm")).lower()
            if norm_type not in ("masked_batchnorm", "masked_dyt", "masked_layernorm"):
                raise NotImplementedError(f"residual blocks with norm_type={norm_type!r} are not supported")
            if c.get("return_nmd") and norm_type != "masked_batchnorm":
                raise NotImplementedError("residual_block(return_nmd=True) needs masked_batchnorm (MaskedDYT rejects it, layers.py:396-400)")
            layers.append(LayerSpec("resblock", dict(
                return_nmd=bool(c.get("return_nmd", False)),      # NMD of the LAST block's bn2 (layers.py:1897-1898, 2696-2704)
                block_size=int(c.get("block_size", 1)), filters=int(c["filters"]),
                kernel_size=int(c.get("kernel_size", 3)), dilation=int(c.get("dilation_rate", 1)),
                use_bias=bool(c.get("use_bias", True)), activation=_act_name(c, model.get("activation", "gelu")) or "gelu",
                use_masking=bool(c.get("use_masking", use_masking)),
                strides=strides, use_1x1conv=bool(c.get("use_1x1conv", False)),      # layers.py:1855-1864: bypass conv when asked for or strided
                norm={"masked_dyt": "dyt", "masked_layernorm": "ln"}.get(norm_type, "bn"), alpha_init=float(c.get("alpha_init", 0.5)),
                ln_epsilon=1e-3)))                                # ResidualBlock._make_norm: MaskedLayerNormalization(name=...), default epsilon
        elif name == "dropout":
            continue
        else:
            raise NotImplementedError(f"layer {name!r} is outside the supported residual-CNN family")
    pooling = str(rep.get("pooling", "max")).lower()
    pooling = {"masked_max": "max", "masked_average": "average"}.get(pooling, pooling)       # builder.py:1703-1713 aliases
    if pooling not in ("max", "average"):
        raise NotImplementedError(f"pooling {pooling!r} is not supported")

    def dense_stack(section):
        out = []
        for lc in section.get("hidden_layers", []):
            if str(lc.get("name", "")).lower() == "dense":
                c = lc.get("config") or {}
                out.append(dict(units=int(c["units"]), activation=_act_name(c), use_bias=bool(c.get("use_bias", True))))
        return out

    signals = None
    classifier = dense_stack(model["classifier"])
    if not classifier or classifier[-1]["activation"] is not None:
        raise NotImplementedError("classifier head must end in a linear Dense layer (logits)")
    if len(classifier) > 3 or len({d["activation"] for d in classifier[:-1]}) > 1:
        raise NotImplementedError("classifier head: at most two hidden Dense layers with one activation")
    reliability = None
    if "reliability_model" in model:
        rm = model["reliability_model"]
        mode = rm.get("mode", "nmd")
        if mode not in ("nmd", "nmd_plus_signals"):
            raise ValueError(f"Unsupported reliability_model.mode: {mode!r}. Use 'nmd' or 'nmd_plus_signals'.")   # builder.py:628-632
        if mode == "nmd_plus_signals":
            signals = list(rm.get("signals", OOD_SIGNALS))
            bad = sorted(set(signals) - set(OOD_SIGNALS))
            if bad:
                raise ValueError(f"Unsupported signal(s): {bad}. Supported: {sorted(OOD_SIGNALS)}")                # layers.py:1626-1631
        reliability = dense_stack(rm)
        if len(reliability) != 2 or reliability[1]["units"] != 1:
            raise NotImplementedError("reliability head must be Dense(h, act) -> Dense(1)")
    return ModelSpec(name=str(model.get("name", "jaeger")), classes=list(model.get("class_label_map", [])),
                     embedding=emb, string_processor=sp, layers=layers, pooling=pooling,
                     classifier=classifier, reliability=reliability, use_masking=use_masking, reliability_signals=signals)


def flat_schema_to_layer_list(model: dict[str, Any]) -> dict[str, Any]:
    """First-generation (G1) project files -- `jaeger_57341_1.5M_fragment`, `jaeger_38341_1.4M_fragment`, template
    commands/configs/nn_config.yaml:36-66 -- describe the network with flat keys instead of a layer list.  This rewrites them
    as the layer list the current builder would need for the same network, following the one place the reference still reads
    those keys (nnlib/inference.py:184-260, DynamicInferenceModelBuilder._build_representation_learner):
        masked_conv1d_1_* (SAME)  -> activation
        per stack i: block_sizes[i] residual blocks (block_filters / block_kernel_size / block_kernel_dilation; the stride of
        block_kernel_strides[i] on the first block only, which also gets the 1x1 bypass)
        masked_conv1d_final_* (SAME, block_filters[-1] filters) -> activation -> global pooling
        classifier: Dense(dense_1_units, activation) -> Dense(classes);  reliability: Dense(dense_1_units, activation) -> Dense(1)
    One-hot input [6, L, 64] through Dense(embedding_size, no bias); no mask propagation (these models were traced without
    it, docs/_source/optimizations.md; scripts/convert_legacy_classifier_checkpoint.py:57-60).  The per-stack MaxPooling2D of
    that vestigial builder is NOT reproduced: it applies the heads to a 4-D map and is no specification of the shipped graphs
     -- the tensor shapes of the model's SavedModel bundle are checked against this layer list at load."""
    rep = model["representation_learner"]
    act = str(model.get("activation", "gelu")).lower()
    n = len(rep.get("block_sizes", []))
    get = lambda key, default: list(rep.get(key, [default] * n))     # noqa: E731
    sizes, filters = get("block_sizes", 2), get("block_filters", 128)
    ksize, dil, strides = get("block_kernel_size", 5), get("block_kernel_dilation", 3), get("block_kernel_strides", 1)

    def conv(prefix, f):
        return {"name": "masked_conv1d", "config": {
            "filters": int(f), "kernel_size": int(rep.get(f"{prefix}_kernel_size", 7 if prefix.endswith("_1") else 5)),
            "strides": int(rep.get(f"{prefix}_strides", 1)), "dilation_rate": int(rep.get(f"{prefix}_dilation_rate", 1)),
            "padding": "same", "use_bias": True, "activation": None}}
    hidden = [conv("masked_conv1d_1", rep.get("masked_conv1d_1_filters", 128)), {"name": "activation", "config": {"activation": act}}]
    for i in range(n):
        for j in range(int(sizes[i])):
            st = int(strides[i]) if j == 0 else 1
            hidden.append({"name": "residual_block", "config": {
                "block_size": 1, "filters": int(filters[i]), "kernel_size": int(ksize[i]), "dilation_rate": int(dil[i]),
                "strides": st, "use_1x1conv": st > 1, "use_bias": True, "activation": act}})
    hidden += [conv("masked_conv1d_final", filters[-1] if n else rep.get("masked_conv1d_1_filters", 128)),
               {"name": "activation", "config": {"activation": act}}]
    pooling = str(rep.get("pooling", "max")).lower()
    out = dict(model)
    emb = dict(model.get("embedding", {}))
    emb.setdefault("input_type", emb.get("type", "translated"))
    emb["use_embedding_layer"] = False                 # one-hot [6, L, 64] input, Dense(embedding_size, use_bias=False)
    out["embedding"] = emb
    out["use_masking"] = False
    out["representation_learner"] = {"hidden_layers": hidden, "pooling": {"avg": "average"}.get(pooling, pooling)}

    def head(section, units_out):
        h = int(section.get("dense_1_units", 128))
        return {"hidden_layers": [{"name": "dense", "config": {"units": h, "activation": act, "use_bias": True}},
                                  {"name": "dense", "config": {"units": int(units_out), "activation": None, "use_bias": True}}]}
    cls = model.get("classifier", {})
    n_cls = int(cls.get("classes", len(model.get("class_label_map", [])) or 6))
    out["classifier"] = head(cls, n_cls)
    # the G1 reliability head reads pooled features, not NMD vectors: that graph is not built here, the head is left out
    out.pop("reliability_model", None)
    return out


def load_project(path: str | Path) -> ModelSpec:
    """nnlib/inference.py:438 reads the file with plain yaml.safe_load."""
    return parse_project(yaml.safe_load(Path(path).read_text()) or {})


def string_processor_config(spec: ModelSpec) -> dict[str, Any]:
    """What InferModel._load_string_processor_config derives (nnlib/inference.py:423-483)."""
    cfg = dict(spec.embedding)
    cfg.update(spec.string_processor)
    cfg["input_type"] = cfg.get("type", "translated")
    codon_id_name = cfg.get("codon_id", "CODON_ID")
    if codon_id_name not in CODON_TABLE_NAMES:
        raise NotImplementedError(f"codon_id {codon_id_name!r} is not supported (dicodons are out of scope)")
    from . import codon_tables
    ids = codon_tables.TABLES[codon_id_name]
    cfg["codon_id_name"] = codon_id_name
    cfg["codon_id"] = ids
    cfg["codon_depth"] = max(ids) + 1
    cfg["vocab_size"] = len(ids) + 1
    cfg["ngram_width"] = 3
    shape = spec.embedding.get("input_shape")
    if cfg.get("seq_onehot") is None and shape is not None:
        cfg["seq_onehot"] = len(shape) == 3 and shape[-1] is not None and shape[-1] > 1
    cfg["seq_onehot"] = bool(cfg.get("seq_onehot", False))
    if not cfg["seq_onehot"]:
        cfg["codon_depth"] = 1
    cfg["masking"] = bool(cfg.get("masking", False))
    if "crop_size" in cfg:                          # nnlib/inference.py:466-482: the trained fragment length
        cfg.setdefault("crop_units", "codon")
        cfg["crop_size_codons"], cfg["crop_size_nt"] = resolve_crop(cfg)
    return cfg


def resolve_crop(sp: dict[str, Any]) -> tuple[int, int]:
    """seqops/crop.py:74-93: (codons, nucleotides) of a string_processor config; nt = 3 * codons + 5."""
    if "crop_size" not in sp:
        raise ValueError("string_processor config must define 'crop_size'")
    size = sp["crop_size"]
    if not isinstance(size, int) or isinstance(size, bool) or size <= 0:
        raise ValueError(f"crop_size must be a positive integer, got {size!r}")
    units = sp.get("crop_units", "codon")
    if units == "codon":
        return size, 3 * size + 5
    if units == "nucleotide":
        return (size - 5) // 3, size
    raise ValueError(f"crop_units must be 'codon' or 'nucleotide', got {units!r}")


def crop_length_warning(trained_codons: int | None, trained_nt: int | None, fsize: int) -> str | None:
    """commands/predict.py:36-64: the warning `jaeger predict` logs when --fsize does not match the model's trained
    fragment length (None when it matches or nothing is known)."""
    if trained_codons is not None:
        runtime_codons = (int(fsize) - 5) // 3
        if runtime_codons == trained_codons:
            return None
        nt_hint = f" ({trained_nt} nt)" if trained_nt is not None else ""
        prefer = trained_nt if trained_nt is not None else "used at training"
        return (f"runtime --fsize {fsize} maps to {runtime_codons} codon frames, but the model was trained on {trained_codons} "
                f"codons{nt_hint}. Fixed-length architectures (e.g. hyena) may degrade or collapse to a single class at a "
                f"different length; prefer --fsize {prefer} for this model.")
    if trained_nt is not None and int(fsize) != int(trained_nt):
        return (f"runtime --fsize {fsize} differs from the model's trained fragment length ({trained_nt} nt). Fixed-length "
                f"architectures (e.g. hyena) may degrade at a different length; prefer --fsize {trained_nt} for this model.")
    return None


# ---- weights ---------------------------------------------------------------------------------

def _glorot(rng, shape, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def _bn(rng, c):
    return dict(gamma=np.ones(c, np.float32), beta=np.zeros(c, np.float32),
                mean=rng.normal(0.0, 0.1, c).astype(np.float32),
                var=rng.uniform(0.5, 1.5, c).astype(np.float32))


def _dyt(rng, c, alpha_init=0.5):
    """MaskedDYT variables (layers.py:407-427); randomised away from the initialisers so tests see them."""
    return dict(alpha=np.array([alpha_init * rng.uniform(0.6, 1.4)], np.float32), gamma=rng.uniform(0.5, 1.5, c).astype(np.float32),
                beta=rng.normal(0.0, 0.1, c).astype(np.float32))


def _ln(rng, c):
    """MaskedLayerNormalization variables (layers.py:318-335); randomised away from the initialisers so tests see them."""
    return dict(gamma=rng.uniform(0.5, 1.5, c).astype(np.float32), beta=rng.normal(0.0, 0.1, c).astype(np.float32))


def _conv(rng, k, cin, cout):
    return dict(kernel=_glorot(rng, (k, cin, cout), k * cin, k * cout), bias=np.zeros(cout, np.float32))


def init_random(spec: ModelSpec, seed: int = 0) -> dict[str, Any]:
    """Random-init weights of the architecture (SURVEY.md 8d config 2 recipe: Glorot-uniform
    kernels, zero bias, BN gamma=1 beta=0 mean~N(0,0.1) var~U(0.5,1.5))."""
    rng = np.random.default_rng(seed)
    e = spec.embedding_size
    w: dict[str, Any] = {"layers": []}
    if e > 0:
        if spec.uses_token_input:
            q, _ = np.linalg.qr(rng.normal(size=(max(65, e), max(65, e))))
            w["embedding"] = q[:65, :e].astype(np.float32)
        else:
            q, _ = np.linalg.qr(rng.normal(size=(max(64, e), max(64, e))))
            w["embedding"] = q[:64, :e].astype(np.float32)
        ch = e
    else:
        w["embedding"] = None
        ch = 64
    for layer in spec.layers:
        c = layer.cfg
        if layer.kind == "conv":
            w["layers"].append(_conv(rng, c["kernel_size"], ch, c["filters"]))
            ch = c["filters"]
        elif layer.kind == "norm":
            w["layers"].append(_dyt(rng, ch, c.get("alpha_init", 0.5)) if c.get("type") == "dyt" else _ln(rng, ch) if c.get("type") == "ln" else _bn(rng, ch))
        elif layer.kind == "nmd":
            w["layers"].append(dict(moving_mean=rng.normal(0.0, 0.1, ch).astype(np.float32)))
        elif layer.kind == "resblock":
            blocks = []
            for _ in range(c["block_size"]):
                norm = (lambda n: _dyt(rng, n, c.get("alpha_init", 0.5))) if c.get("norm") == "dyt" else (lambda n: _ln(rng, n)) if c.get("norm") == "ln" else (lambda n: _bn(rng, n))
                blk = dict(conv1=_conv(rng, c["kernel_size"], ch, c["filters"]), bn1=norm(c["filters"]),
                           conv2=_conv(rng, c["kernel_size"], c["filters"], c["filters"]),
                           bn2=norm(c["filters"]))
                if block_has_bypass(c, len(blocks)):
                    blk["conv3"], blk["bn3"] = _conv(rng, 1, ch, c["filters"]), norm(c["filters"])
                if not c.get("use_bias", True):
                    for name in ("conv1", "conv2", "conv3"):
                        if name in blk:
                            blk[name]["bias"] = np.zeros_like(blk[name]["bias"])
                blocks.append(blk)
                ch = c["filters"]
            w["layers"].append(dict(blocks=blocks))
        else:
            w["layers"].append({})
    feat = ch
    w["classifier"], width = [], feat
    for d in spec.classifier:
        w["classifier"].append(dict(kernel=_glorot(rng, (width, d["units"]), width, d["units"]),
                                    bias=(rng.normal(0.0, 0.05, d["units"]).astype(np.float32) if len(spec.classifier) > 1
                                          else np.zeros(d["units"], np.float32))))
        width = d["units"]
    if spec.reliability is not None:
        n_nmd = sum(1 for layer in spec.layers if layer.kind == "nmd" or layer.cfg.get("return_nmd"))
        nmd_dim = 0
        chn = e if e > 0 else 64
        for layer in spec.layers:
            if layer.kind in ("conv", "resblock"):
                chn = layer.cfg["filters"]
            if layer.kind == "nmd" or layer.cfg.get("return_nmd"):
                nmd_dim += chn
        h = spec.reliability[0]["units"]
        nmd_dim += len(spec.reliability_signals or [])
        w["reliability"] = [dict(kernel=_glorot(rng, (nmd_dim, h), nmd_dim, h), bias=np.zeros(h, np.float32)),
                            dict(kernel=_glorot(rng, (h, 1), h, 1), bias=np.zeros(1, np.float32))]
        assert n_nmd > 0, "reliability head needs at least one nmd layer"
    return w


def block_has_bypass(cfg: dict[str, Any], block_index: int) -> bool:
    """layers.py:1855-1864 + 2680-2682: a block gets the 1x1 bypass conv when it is strided (every block of a strided stack
    is) or when `use_1x1conv` is set, which ResidualBlockStack hands to the FIRST block of the stack only."""
    return int(cfg.get("strides", 1)) > 1 or (bool(cfg.get("use_1x1conv", False)) and block_index == 0)


def count_params(spec: ModelSpec, weights: dict[str, Any], representation_only: bool = True) -> int:
    """Parameter count as Keras' rep_model.count_params() sees it (all variables incl. the
    BN moving statistics and NMD moving means)."""
    n = 0
    if weights.get("embedding") is not None:
        n += weights["embedding"].size
    for lw in weights["layers"]:
        if "blocks" in lw:
            for b in lw["blocks"]:
                for part in b.values():
                    n += sum(v.size for v in part.values())
        else:
            n += sum(v.size for v in lw.values())
    if not representation_only:
        for sec in ("classifier", "reliability"):
            for d in weights.get(sec) or []:
                n += d["kernel"].size + d["bias"].size
    return int(n)


def standin_1p4m_config() -> dict[str, Any]:
    """The declared stand-in for `jaeger_38341_1.4M_fragment`: the
    nmd_merge family of train_config/nn_config_1500bp_nmd_merge_6_class_brain.yaml with four
    residual stacks instead of three."""
    def conv(f, k):
        return {"name": "masked_conv1d", "config": {"filters": f, "kernel_size": k, "strides": 1,
                                                    "dilation_rate": 1, "use_bias": True, "activation": None}}
    tail = [{"name": "nmd"}, {"name": "masked_batchnorm", "config": {"return_nmd": False}},
            {"name": "activation", "config": {"activation": "gelu"}}]
    block = {"name": "residual_block", "config": {"use_1x1conv": False, "block_size": 2, "filters": 128,
                                                  "kernel_size": 5, "dilation_rate": 3, "use_bias": True}}
    hidden = [conv(128, 7)] + tail
    for _ in range(4):
        hidden += [block] + tail
    classes = ["bacteria", "phage", "eukarya", "archaea", "plasmid", "virus"]
    return {"model": {
        "name": "jaeger_standin_1p4M", "activation": "gelu", "classifier_out_dim": 6,
        "class_label_map": [{"class": c, "label": i} for i, c in enumerate(classes)],
        "embedding": {"use_embedding_layer": True, "input_type": "translated", "frames": 6,
                      "input_shape": [6, None], "embedding_size": 128},
        "string_processor": {"seq_onehot": False, "codon": "CODON", "codon_id": "CODON_ID", "masking": False},
        "representation_learner": {"hidden_layers": hidden, "pooling": "max"},
        "classifier": {"input_shape": 128, "hidden_layers": [
            {"name": "dropout", "config": {"rate": 0.1}},
            {"name": "dense", "config": {"units": 6, "activation": None, "use_bias": True}}]},
        "reliability_model": {"merge": {"mode": "concat", "axis": -1}, "input_shape": 640, "hidden_layers": [
            {"name": "dropout", "config": {"rate": 0.3}},
            {"name": "dense", "config": {"units": 8, "activation": "gelu", "use_bias": True}},
            {"name": "dropout", "config": {"rate": 0.1}},
            {"name": "dense", "config": {"units": 1, "activation": None, "use_bias": True}}]},
    }}


def baseline_500bp_config() -> dict[str, Any]:
    """train_config/nn_config_500bp_baseline.yaml:17-100 (BASELINE config 3): E64 -> conv(32, k7, VALID) -> BN -> GELU ->
    2 residual stacks (block_size 2, 32 filters, k3, SAME) -> BN -> GELU -> average pool -> Dense 3; no reliability head."""
    def conv(f, k):
        return {"name": "masked_conv1d", "config": {"filters": f, "kernel_size": k, "strides": 1, "dilation_rate": 1,
                                                    "use_bias": True, "activation": None}}
    bn_act = [{"name": "masked_batchnorm", "config": {"return_nmd": False}}, {"name": "activation", "config": {"activation": "gelu"}}]
    block = {"name": "residual_block", "config": {"use_1x1conv": False, "block_size": 2, "filters": 32, "kernel_size": 3, "use_bias": True}}
    return {"model": {
        "name": "jaeger_500bp_baseline", "activation": "gelu",
        "class_label_map": [{"class": c, "label": i} for i, c in enumerate(["chromosome", "virus", "plasmid"])],
        "embedding": {"use_embedding_layer": True, "input_type": "translated", "input_shape": [6, None], "embedding_size": 64},
        "string_processor": {"seq_onehot": False, "codon": "CODON", "codon_id": "CODON_ID", "crop_size": 500, "masking": False},
        "representation_learner": {"hidden_layers": [conv(32, 7)] + bn_act + [block, block] + bn_act, "pooling": "average"},
        "classifier": {"input_shape": 32, "hidden_layers": [{"name": "dense", "config": {"units": 3, "activation": None, "use_bias": True}}]}}}
