"""B200Engine -- the fifth inference engine behind the reference's engine contract.

The reference's `commands/predict.py:687-747` programs against an object with
`class_map`, `string_processor_config` and `predict(dataset) -> dict[str, np.ndarray]`
(`InferModel`, nnlib/inference.py:300-421).  `B200Engine` keeps that contract and owns the
whole device pipeline: 2-bit pack -> window/encode kernel -> tcgen05 conv stack -> heads.

    engine = B200Engine(path_dict)                  # AvailableModels.info[name], or spec=/weights=
    y = engine.predict(WindowSource(fasta, fsize=2000, stride=1500, ...))
    y["prediction"] [W, n_cls] f32, y["reliability"] [W, 1], y["embedding"], y["nmd"],
    y["meta_0".."meta_9"] byte-string arrays in window order (long pass then short pass).

`predict` also accepts the reference's dataset protocol -- an iterable of
`(inputs_dict, meta_0, ..., meta_9)` batches with `inputs_dict["translated"]` holding tokens
[B, 6, L] (or one-hot [B, 6, L, depth]) -- in which case only stage 3 runs on the device.

PyTorch is used for device / pinned memory and streams only.  No CPU fallback exists: the
constructor raises without the CUDA library or a B200.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Iterable

import numpy as np
import torch
import yaml

from . import _cabi, codon_tables
from ._cabi import check, lib
from .modelspec import ModelSpec, init_random, load_project, string_processor_config
from .plan import Plan, compile_plan, to_ctypes

_SKEW_STR = np.array([f"{v / 100: .3f}" for v in range(-100, 101)] + ["-0.000"], dtype="S6")


def read_fasta(path: str | Path):
    """(name, sequence) pairs: name = header up to the first whitespace, sequence = the
    record's lines joined (pyfastx semantics used at seqops/io.py:98-104)."""
    name, chunks = None, []
    with open(path, "rb") as fh:
        for line in fh:
            if line.startswith(b">"):
                if name is not None:
                    yield name, b"".join(chunks)
                fields = line[1:].split()
                name = fields[0].decode() if fields else ""
                chunks = []
            elif name is not None:
                chunks.append(line.strip())
    if name is not None:
        yield name, b"".join(chunks)


def load_fasta(path: str | Path) -> tuple[list[str], torch.Tensor, np.ndarray]:
    """Native one-pass FASTA ingest (jg_fasta_scan / jg_fasta_load): record names, all bases back
    to back in one PINNED host buffer (the H2D source) and the n+1 record offsets."""
    cpath = str(path).encode()
    n, nb, nn = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    check(lib.jg_fasta_scan(cpath, ctypes.byref(n), ctypes.byref(nb), ctypes.byref(nn)))
    host = torch.empty(max(nb.value, 1), dtype=torch.uint8)
    if torch.cuda.is_available():
        host = host.pin_memory()
    offsets = np.zeros(n.value + 1, dtype=np.int64)
    names = ctypes.create_string_buffer(max(nn.value, 1))
    check(lib.jg_fasta_load(cpath, ctypes.c_void_p(host.data_ptr()), offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), names))
    name_list = [x.decode() for x in names.raw[:nn.value].split(b"\0")[:n.value]]
    return name_list, host[:nb.value], offsets


@dataclass
class WindowSource:
    """What `fragment_generator` + `process_string_inference` are parameterised with
    (commands/predict.py:186-245): where the contigs come from and how to window them."""
    fasta: str | Path | None = None
    records: list[tuple[str, bytes | str]] | None = None
    fsize: int = 2000
    stride: int = 1500
    min_len: int | None = None           # < fsize enables the two-pass short-contig mode
    dynamic_stride: bool = False
    dynamic_stride_threshold: float = 10.0
    batch: int = 96                      # only shapes the short pass' padded batches
    dustmask: bool = False               # symmetric DUST soft-masking on the device (reference default: on)
    softmasks: dict[str, np.ndarray] | None = None   # explicit per-contig bool arrays (overrides dustmask)
    outputs: tuple[str, ...] | None = None   # model outputs to bring back (None = all: prediction, reliability, embedding, nmd)
    lazy_meta: bool = False              # build the meta_0..9 byte-string arrays only when they are read

    @classmethod
    def from_host(cls, names: list[str], bases, offsets: np.ndarray, **kw) -> "WindowSource":
        """Contigs that already sit in ONE host buffer (`bases`: uint8 NumPy array or CPU torch tensor, all records
        back to back; `offsets`: n + 1 record starts).  The buffer is used as the H2D source as it is -- pin it
        (torch `pin_memory()`) for an asynchronous copy."""
        src = cls(**kw)
        host = bases if isinstance(bases, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(bases, dtype=np.uint8))
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        src._loaded = (list(names), host[:int(off[-1])], off)
        return src

    def load(self) -> tuple[list[str], torch.Tensor, np.ndarray]:
        """(names, bases in one pinned host buffer, record offsets); read once and kept."""
        if getattr(self, "_loaded", None) is not None:
            return self._loaded
        if self.records is None:
            self._loaded = load_fasta(self.fasta)
            return self._loaded
        recs = [(n, s.encode() if isinstance(s, str) else bytes(s)) for n, s in self.records]
        offsets = np.zeros(len(recs) + 1, dtype=np.int64)
        np.cumsum([len(s) for _, s in recs], out=offsets[1:])
        host = torch.empty(max(int(offsets[-1]), 1), dtype=torch.uint8)
        if torch.cuda.is_available():
            host = host.pin_memory()
        hv = host.numpy()
        for o, (_, s) in zip(offsets[:-1], recs):
            hv[o:o + len(s)] = np.frombuffer(s, dtype=np.uint8)
        self._loaded = [n for n, _ in recs], host[:int(offsets[-1])], offsets
        return self._loaded


class PredictResult(dict):
    """The dict `engine.predict` returns.  With lazy meta the ten `meta_i` byte-string arrays of the
    reference protocol (encode.py:304-316) are built from the numeric window table on first access."""
    window_table = None
    device_outputs = None        # the model outputs as they still sit in HBM (stage 4 reads them there, no re-upload)

    def _build(self, key):
        t = self.window_table
        i = int(key[5:])
        if i == 0:
            return np.array([n.encode() for n in t.headers], dtype="S")[t.contig]
        if i == 9:
            idx = np.where(t.skew100 == (1 << 14), 201, t.skew100.astype(np.int32) + 100)
            return _SKEW_STR[idx]
        src = {1: t.start, 2: t.is_last.astype(np.int32), 3: t.ordinal, 4: t.seqlen}.get(i)
        if src is None:
            src = t.counts[:, i - 5]
        return src.astype("S")

    def __missing__(self, key):
        if self.window_table is not None and isinstance(key, str) and key.startswith("meta_") and key[5:].isdigit() and int(key[5:]) < 10:
            self[key] = v = self._build(key)
            return v
        raise KeyError(key)


@dataclass
class WindowTable:
    """Numeric per-window metadata of the last `predict` call (what meta_0..9 encode)."""
    headers: list[str] = field(default_factory=list)        # per contig, in window order
    contig: np.ndarray | None = None     # window -> index into headers
    start: np.ndarray | None = None
    nbases: np.ndarray | None = None
    ordinal: np.ndarray | None = None
    is_last: np.ndarray | None = None
    seqlen: np.ndarray | None = None
    counts: np.ndarray | None = None     # [W, 4] G, C, A, T
    skew100: np.ndarray | None = None


class B200Engine:
    def __init__(self, path_dict: dict[str, Any] | None = None, *, spec: ModelSpec | None = None,
                 weights: dict[str, Any] | None = None, device: int = 0, workspace_gb: float = 16.0,
                 seed: int = 0, use_ref_kernels: bool = False, legacy_weights: dict[str, Any] | None = None,
                 all_labels: bool = False):
        if not torch.cuda.is_available():
            raise _cabi.JaegerB200Error("no CUDA device: jaeger_b200 has no CPU fallback")
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        # a torch pool stream: it outlives the engine, so pinned buffers that were copied from on it
        # can be released by torch's host allocator at any later time
        self._tstream = torch.cuda.Stream(device=self.tdev)
        self.ctx = _cabi.Context(self.device, self._tstream.cuda_stream)
        self.use_ref_kernels = bool(use_ref_kernels)
        self.class_map = None
        if legacy_weights is not None:
            # the bundled `default` model (predict_legacy.py:188-221): fixed graph, legacy encoder
            from . import legacy
            self.spec, self.weights = None, legacy_weights
            labels = legacy.ALL_LABELS if all_labels else legacy.DEFAULT_LABELS
            self.class_map = {"num_classes": 4, "class": [labels[i] for i in range(4)], "index": [0, 1, 2, 3]}
            self.string_processor_config = {"input_type": "translated", "legacy": True, "masking": True}
            self.plan = legacy.compile_legacy_plan(legacy_weights)
            self.lut, self.case_sensitive = legacy.LEGACY_LUT, 1
            self._create_model(workspace_gb)
            return
        if path_dict is not None:
            project = path_dict.get("project")
            if project is None:
                raise ValueError("model has no *_project.yaml; only layer-list fragment models are supported")
            spec = load_project(project)
            self.class_map = self._load_class_map(path_dict.get("classes"))
            if weights is None:
                from .weights import load_saved_model_weights
                weights = load_saved_model_weights(path_dict, spec)
        if spec is None:
            raise ValueError("B200Engine needs a path_dict or a ModelSpec")
        self.spec = spec
        self.weights = weights if weights is not None else init_random(spec, seed)
        if self.class_map is None and spec.classes:
            self.class_map = {"num_classes": len(spec.classes), "class": [c["class"] for c in spec.classes],
                              "index": [c["label"] for c in spec.classes]}
        self.string_processor_config = string_processor_config(spec)
        self.plan: Plan = compile_plan(spec, self.weights)
        self.lut = codon_tables.device_lut(self.string_processor_config["codon_id"], plus_one=True)
        self.case_sensitive = int(self.string_processor_config["masking"])
        self._create_model(workspace_gb)

    def _create_model(self, workspace_gb: float) -> None:
        layers, head = to_ctypes(self.plan)
        h = ctypes.c_void_p()
        check(lib.jg_model_create(self.ctx.handle, layers, len(self.plan.launches), ctypes.byref(head), 6,
                                  int(self.plan.tok_offset), ctypes.byref(h)))
        self.model = h
        self.workspace_bytes = int(workspace_gb * (1 << 30))
        self.windows = WindowTable()
        self.timings: dict[str, float] = {}

    # ---- reference-compatible helpers ---------------------------------------------------------
    @staticmethod
    def _load_class_map(path):
        """nnlib/inference.py:411-421."""
        if path is None:
            return None
        cm = yaml.safe_load(Path(path).read_text())["classes"]
        return {"num_classes": len(cm), "class": [i["class"] for i in cm], "index": [i["label"] for i in cm]}

    def close(self):
        if getattr(self, "model", None):
            lib.jg_model_destroy(self.model)
            self.model = None
        self.ctx.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- torch plumbing -----------------------------------------------------------------------
    def _stream(self):
        return self._tstream

    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.tdev)

    def _h2d(self, arr: np.ndarray, pinned: bool = True) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if pinned:
            t = t.pin_memory()
        return t.to(self.tdev, non_blocking=True)

    def codons_per_frame(self, n_bases: int, crop: int) -> int:
        off = (-2, -1, 0)[crop % 3]
        v = n_bases - 5 + off
        return 0 if v <= 0 else (v + 2) // 3

    # ---- stages -------------------------------------------------------------------------------
    def pack(self, ascii_dev: torch.Tensor):
        n = ascii_dev.numel()
        codes = torch.zeros((n + 15) // 16 + 4, dtype=torch.int32, device=self.tdev)
        valid = torch.zeros((n + 31) // 32 + 4, dtype=torch.int32, device=self.tdev)
        check(lib.jg_pack_bases(self.ctx.handle, ascii_dev.data_ptr(), n, codes.data_ptr(), valid.data_ptr()))
        return codes, valid

    def dust(self, codes, valid, offsets: np.ndarray, threshold: int = 20, chunk: int = 1024) -> torch.Tensor:
        """Soft-mask bitmap (layout of `valid`) of the low-complexity bases of every contig:
        pydustmasker.DustMasker(seq, window_size=64, score_threshold=20) at seqops/io.py:105-108."""
        lens = np.diff(offsets)
        n_chunks = (lens + chunk - 1) // chunk
        owner = np.repeat(np.arange(len(lens)), n_chunks)
        first = np.concatenate([[0], np.cumsum(n_chunks)])[:-1]
        k = np.arange(int(n_chunks.sum())) - np.repeat(first, n_chunks)
        cbeg = offsets[owner] + k * chunk
        cend = np.minimum(cbeg + chunk, offsets[owner + 1])
        soft = torch.zeros_like(valid)
        if len(cbeg):
            # keep the device copies referenced until the launch is enqueued (a temporary's block
            # would be recycled by the next _h2d of the same expression)
            d_cb, d_ce, d_gb, d_ge = (self._h2d(a) for a in (cbeg, cend, offsets[owner], offsets[owner + 1]))
            check(lib.jg_dust_mask(self.ctx.handle, codes.data_ptr(), valid.data_ptr(), d_cb.data_ptr(), d_ce.data_ptr(),
                                   d_gb.data_ptr(), d_ge.data_ptr(), len(cbeg), int(threshold), soft.data_ptr()))
        return soft

    @staticmethod
    def plan_windows(lens: np.ndarray, fsize: int, stride: int, dynamic_stride=False, threshold=10.0,
                     min_len: int | None = None, max_len: int = 0, short_pass: bool = False):
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        n = ctypes.c_int64(0)
        lp = lens.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
        args = (lp, len(lens), int(fsize), int(stride), int(bool(dynamic_stride)), float(threshold),
                int(fsize if min_len is None else min_len), int(max_len), int(bool(short_pass)))
        null = [ctypes.POINTER(t)() for t in (ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint8)]
        check(lib.jg_plan_windows(*args, ctypes.byref(n), *null))
        w = n.value
        contig = np.empty(w, np.int32); start = np.empty(w, np.int64); nb = np.empty(w, np.int32)
        ordinal = np.empty(w, np.int32); last = np.empty(w, np.uint8)
        if w:
            check(lib.jg_plan_windows(*args, ctypes.byref(n),
                                      contig.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                      start.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                      nb.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                      ordinal.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                      last.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
        return contig, start, nb, ordinal, last

    def encode(self, codes, valid, win_base: torch.Tensor, win_nbases: torch.Tensor, crop: int, lc: int,
               soft: torch.Tensor | None = None, lut: np.ndarray | None = None, case_sensitive: int | None = None):
        w = win_base.numel()
        pitch = (lc + 3) // 4 * 4
        tokens = self._empty((w, 6, pitch), torch.uint8)
        counts = self._empty((w, 4), torch.int32)
        skew = self._empty((w,), torch.int16)
        lut = self.lut if lut is None else np.ascontiguousarray(lut, np.uint8)
        check(lib.jg_encode_windows(self.ctx.handle, codes.data_ptr(), valid.data_ptr(),
                                    soft.data_ptr() if soft is not None else None, win_base.data_ptr(),
                                    win_nbases.data_ptr(), w, int(crop), int(lc), pitch,
                                    lut.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                    self.case_sensitive if case_sensitive is None else int(case_sensitive),
                                    tokens.data_ptr(), counts.data_ptr(), skew.data_ptr()))
        return tokens, counts, skew

    def forward(self, tokens: torch.Tensor, lpad: torch.Tensor, lc: int):
        """tokens [W, 6, pitch] uint8 on device -> dict of device tensors."""
        w, _, pitch = tokens.shape
        p = self.plan
        out = {"prediction": self._empty((w, p.n_classes), torch.float32),
               "embedding": self._empty((w, p.feat_dim), torch.float32)}
        if p.n_taps:
            out["nmd"] = self._empty((w, p.n_taps * p.tap_width), torch.float32)
        if p.rel is not None:
            out["reliability"] = self._empty((w, 1), torch.float32)
        chunk = int(lib.jg_model_max_windows(self.model, int(lc), self.workspace_bytes))
        if chunk < 1:
            raise _cabi.JaegerB200Error("workspace budget too small for one window")
        for b in range(0, w, chunk):
            e = min(w, b + chunk)
            check(lib.jg_model_forward(
                self.ctx.handle, self.model, tokens[b:e].data_ptr(), lpad[b:e].data_ptr(), e - b, int(lc), int(pitch),
                out["prediction"][b:e].data_ptr(),
                out["reliability"][b:e].data_ptr() if "reliability" in out else None,
                out["embedding"][b:e].data_ptr(),
                out["nmd"][b:e].data_ptr() if "nmd" in out else None, int(self.use_ref_kernels)))
        if p.real_feat_dim is not None and p.real_feat_dim != p.feat_dim:
            out["embedding"] = out["embedding"][:, :p.real_feat_dim]     # drop the zero padding channels
        if p.nmd_cols is not None and "nmd" in out:
            out["nmd"] = out["nmd"][:, torch.as_tensor(p.nmd_cols, device=self.tdev)]     # taps of layers narrower than 64 channels
        return out

    def aggregate(self, logits: torch.Tensor, rel: torch.Tensor | None, offsets: torch.Tensor) -> dict[str, torch.Tensor]:
        """Stage 4 on the device: per-contig reductions of window logits (collect.py:293-403).
        offsets [n_contigs + 1] int64: windows of contig c are rows offsets[c]:offsets[c+1]."""
        nc = offsets.numel() - 1
        w, ncls = logits.shape
        out = {"pred_sum": self._empty((nc, ncls), torch.float16), "pred_var": self._empty((nc, ncls), torch.float16),
               "consensus": self._empty((nc,), torch.int32), "per_class_counts": self._empty((nc, ncls), torch.int32),
               "entropy": self._empty((nc,), torch.float16), "energy": self._empty((nc,), torch.float16),
               "rel_pos": self._empty((nc,), torch.int32), "frag_pred": self._empty((w,), torch.int32)}
        check(lib.jg_aggregate_contigs(
            self.ctx.handle, logits.data_ptr(), rel.data_ptr() if rel is not None else None, offsets.data_ptr(), nc,
            w, ncls, out["pred_sum"].data_ptr(), out["pred_var"].data_ptr(), out["consensus"].data_ptr(),
            out["per_class_counts"].data_ptr(), out["entropy"].data_ptr(), out["energy"].data_ptr(),
            out["rel_pos"].data_ptr(), out["frag_pred"].data_ptr()))
        return out

    def viterbi(self, logits: torch.Tensor, offsets: torch.Tensor, costs: np.ndarray) -> tuple[torch.Tensor, torch.Tensor]:
        """--crf window decoding on the device (helpers.py:393-449): (path [W] int32, counts [n_contigs, n_cls] int32)."""
        nc = offsets.numel() - 1
        w, ncls = logits.shape
        costs_dev = self._h2d(np.ascontiguousarray(costs, dtype=np.float64))
        path, counts = self._empty((w,), torch.int32), self._empty((nc, ncls), torch.int32)
        check(lib.jg_viterbi_decode(self.ctx.handle, logits.data_ptr(), offsets.data_ptr(), nc, w, ncls, costs_dev.data_ptr(),
                                    path.data_ptr(), counts.data_ptr()))
        return path, counts

    def legacy_reliability(self, embedding: torch.Tensor, offsets: torch.Tensor, ood: dict[str, Any]) -> tuple[torch.Tensor, torch.Tensor]:
        """Legacy `default` reliability (helpers.py:558-565): per-window P(class 0) and its per-contig mean, float64."""
        w, dim = embedding.shape
        nc = offsets.numel() - 1
        mean, sd = self._h2d(np.ascontiguousarray(ood["batch_mean"], np.float32)), self._h2d(np.ascontiguousarray(ood["batch_std"], np.float32))
        coef = self._h2d(np.ascontiguousarray(ood["coef"], np.float64))
        p0, cmean = self._empty((w,), torch.float64), self._empty((nc,), torch.float64)
        check(lib.jg_legacy_reliability(self.ctx.handle, embedding.data_ptr(), w, dim, mean.data_ptr(), sd.data_ptr(), coef.data_ptr(),
                                        float(ood["intercept"]), float(ood["cal_a"]), float(ood["cal_b"]), offsets.data_ptr(), nc,
                                        p0.data_ptr(), cmean.data_ptr()))
        return p0, cmean

    def set_profiling(self, on: bool) -> None:
        check(lib.jg_model_set_profiling(self.model, int(bool(on))))

    def get_profile(self):
        """Per conv launch of the plan: (summed kernel ms, launches, windows covered)."""
        n = len(self.plan.launches)
        ms = (ctypes.c_double * n)(); cnt = (ctypes.c_int64 * n)(); win = (ctypes.c_double * n)()
        check(lib.jg_model_get_profile(self.model, n, ms, cnt, win))
        return [(ms[i], cnt[i], win[i]) for i in range(n)]

    def classify_long(self, ascii_dev: torch.Tensor, lens: np.ndarray, fsize: int, stride: int):
        """pack -> plan -> encode -> forward -> aggregate for contigs already in device memory
        (long pass only).  Returns (per-contig dict of device tensors, n_windows, n_contigs_with_windows)."""
        offsets = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        codes, valid = self.pack(ascii_dev)
        contig, start, nb, ordinal, last = self.plan_windows(lens, fsize, stride)
        w = len(contig)
        if w == 0:
            return {}, 0, 0
        lc = self.codons_per_frame(fsize, fsize)
        tokens, counts, skew = self.encode(codes, valid, self._h2d(offsets[contig] + start), self._h2d(nb), fsize, lc)
        out = self.forward(tokens, torch.full((w,), lc, dtype=torch.int32, device=self.tdev), lc)
        ends = np.flatnonzero(last) + 1
        woff = np.concatenate([[0], ends]).astype(np.int64)
        agg = self.aggregate(out["prediction"], out.get("reliability"), self._h2d(woff))
        agg["gc_counts"] = counts
        agg["_logits"], agg["_window_offsets"] = out["prediction"], woff       # for stage 4b (region calling) on the device
        return agg, w, len(ends)

    def conv_kernel_names(self) -> str:
        """Which conv kernel the library picked for the launches of the plan (jg_model_kernel_names)."""
        buf = ctypes.create_string_buffer(4096)
        check(lib.jg_model_kernel_names(self.model, buf, len(buf)))
        return buf.value.decode()

    # ---- the engine contract --------------------------------------------------------------------
    def predict(self, dataset, no_progress: bool = False) -> dict[str, np.ndarray]:
        if isinstance(dataset, WindowSource):
            return self._predict_source(dataset)
        return self._predict_batches(dataset)

    def evaluate(self, dataset, no_progress: bool = False) -> dict[str, float]:
        """InferModel.evaluate (nnlib/inference.py:375-408): dataset yields (inputs_dict, y_true_onehot) batches;
        returns the mean categorical cross-entropy of the logits (from_logits=True) and the accuracy."""
        logits_acc, true_acc = [], []
        for inputs, y_true in dataset:
            logits_acc.append(self._predict_batches([(inputs,)])["prediction"])
            true_acc.append(np.asarray(y_true, dtype=np.float32))
        logits = np.concatenate(logits_acc, axis=0).astype(np.float32)
        y_true = np.concatenate(true_acc, axis=0)
        # keras.losses.categorical_crossentropy(from_logits=True): -sum(y * log_softmax(z)), float32
        z = logits - logits.max(axis=1, keepdims=True)
        log_sm = z - np.log(np.exp(z).sum(axis=1, keepdims=True))
        loss = float(np.mean(-(y_true * log_sm).sum(axis=1)))
        accuracy = float(np.mean(np.argmax(logits, axis=1) == np.argmax(y_true, axis=1)))
        return {"loss": loss, "accuracy": accuracy}

    def _predict_batches(self, dataset: Iterable) -> dict[str, np.ndarray]:
        """Reference dataset protocol: (inputs_dict, meta_0..meta_9) batches (inference.py:341-373)."""
        acc: dict[str, list] = {}
        with torch.cuda.stream(self._stream()):
            for inputs, *meta in dataset:
                x = np.asarray(inputs["translated"] if isinstance(inputs, dict) else inputs)
                if x.ndim == 4:      # one-hot [B,6,L,depth] -> tokens (all-zero row = 0)
                    tok = np.where(x.sum(-1) > 0, x.argmax(-1) + 1, 0).astype(np.uint8)
                else:
                    tok = x.astype(np.uint8)
                b, _, lc = tok.shape
                pitch = (lc + 3) // 4 * 4
                padded = np.zeros((b, 6, pitch), np.uint8)
                padded[:, :, :lc] = tok
                lpad = torch.full((b,), lc, dtype=torch.int32, device=self.tdev)
                out = self.forward(self._h2d(padded), lpad, lc)
                for k, v in out.items():
                    acc.setdefault(k, []).append(v.cpu().numpy())
                for i, mt in enumerate(meta):
                    acc.setdefault(f"meta_{i}", []).append(np.asarray(mt))
        self.ctx.sync()
        return {k: np.concatenate(v, axis=0) for k, v in acc.items()}

    def _predict_source(self, src: WindowSource) -> dict[str, np.ndarray]:
        raw_names, host, offsets = src.load()
        fsize, stride = int(src.fsize), int(src.stride)
        lens = np.diff(offsets)
        total = int(offsets[-1])
        two_pass = src.min_len is not None and src.min_len < fsize      # commands/predict.py:771-810
        if two_pass and any(c.kind == 4 for c in self.plan.launches):
            # TF's SAME padding of a strided conv depends on the parity of the (padded) batch length, which varies between the
            # padded batches of the short pass; the row-plane kernels are picked once per forward call
            raise NotImplementedError("--min-len < --fsize (the padded short-contig pass) is not available for models with strided residual blocks")
        passes = [self.plan_windows(lens, fsize, stride, src.dynamic_stride, src.dynamic_stride_threshold,
                                    min_len=fsize, short_pass=False)]
        if two_pass:
            passes.append(self.plan_windows(lens, fsize, stride, src.dynamic_stride, src.dynamic_stride_threshold,
                                            min_len=src.min_len, max_len=fsize - 1, short_pass=True))
        lc_full = self.codons_per_frame(fsize, fsize)
        results: list[dict[str, torch.Tensor]] = []
        tables = []
        with torch.cuda.stream(self._stream()):
            ascii_dev = host.to(self.tdev, non_blocking=True)
            codes, valid = self.pack(ascii_dev)
            soft = None
            if src.dustmask and not src.softmasks:
                soft = self.dust(codes, valid, offsets)
            if src.softmasks:
                bits = np.zeros(total, dtype=bool)
                for o, n in zip(offsets[:-1], raw_names):
                    if n in src.softmasks:
                        m = np.asarray(src.softmasks[n], dtype=bool)
                        bits[o:o + len(m)] = m
                packed = np.packbits(bits, bitorder="little")
                words = np.zeros((total + 31) // 32 * 4 + 16, np.uint8)
                words[:len(packed)] = packed
                soft = self._h2d(words.view(np.int32))
            for pi, (contig, start, nb, ordinal, last) in enumerate(passes):
                w = len(contig)
                if w == 0:
                    continue
                base = self._h2d(offsets[contig] + start)
                nbd = self._h2d(nb)
                if pi == 0:
                    lc = lc_full
                    lpad_h = np.full(w, lc, np.int32)
                else:
                    # padded_batch pads every batch of `batch` windows to its longest member
                    per = np.array([self.codons_per_frame(int(n), fsize) for n in nb], np.int32)
                    lpad_h = per.copy()
                    for b in range(0, w, src.batch):
                        lpad_h[b:b + src.batch] = per[b:b + src.batch].max()
                    lc = int(per.max())
                tokens, counts, skew = self.encode(codes, valid, base, nbd, fsize, lc, soft)
                out = self.forward(tokens, self._h2d(lpad_h), lc)
                out["_counts"], out["_skew"] = counts, skew
                results.append(out)
                tables.append((contig, start, nb, ordinal, last))
            # header clean-up (seqops/io.py:109) on the host while the device works through the launches queued above
            names = [n.strip().replace(",", "___") for n in raw_names]
            keep = None if src.outputs is None else set(src.outputs)
            host_out = [{k: v.cpu() for k, v in r.items() if keep is None or k in keep or k.startswith("_")} for r in results]
        self.ctx.sync()
        if not host_out:
            return {}
        y: dict[str, np.ndarray] = {}
        for k in host_out[0]:
            y[k] = np.concatenate([r[k].numpy() for r in host_out], axis=0)
        contig = np.concatenate([t[0] for t in tables]); start = np.concatenate([t[1] for t in tables])
        nb = np.concatenate([t[2] for t in tables]); ordinal = np.concatenate([t[3] for t in tables])
        last = np.concatenate([t[4] for t in tables])
        counts, skew = y.pop("_counts"), y.pop("_skew")
        self.windows = WindowTable(headers=names, contig=contig, start=start, nbases=nb, ordinal=ordinal,
                                   is_last=last, seqlen=lens[contig], counts=counts, skew100=skew)
        # meta_0..meta_9 exactly as process_string_inference forwards them (encode.py:304-316)
        res = PredictResult(y)
        res.window_table = self.windows
        with torch.cuda.stream(self._stream()):
            res.device_outputs = {k: (results[0][k] if len(results) == 1 else torch.cat([r[k] for r in results], dim=0))
                                  for k in ("prediction", "reliability") if k in results[0]}
        if not src.lazy_meta:
            for i in range(10):
                res[f"meta_{i}"]
        return res
