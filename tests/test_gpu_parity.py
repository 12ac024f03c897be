"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI,
against the CPU oracle on the same seeded inputs and against the committed golden fixtures.

Tolerances.  Integer / byte / index outputs (window starts, flags, tokens, base counts,
gc_skew strings, per-window argmax, class counts, fp16 means / variances): bit-exact.
Conv-stack outputs: the device keeps activations and weights in fp16 (fp32 accumulation, fp32
BatchNorm affine, packed-fp16 GELU / residual add), the oracle is fp32 throughout; with the
random-init stand-in (|logit| <~ 0.3) the bound asserted is  max|logit - oracle| <= 4e-3  and
max|embedding - oracle| <= 1e-2, and tensor-core vs CUDA-core kernels agree to 2e-3.
"""
import json
import zlib
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def standin():
    from jaeger_b200 import B200Engine, init_random, parse_project, standin_1p4m_config
    spec = parse_project(standin_1p4m_config())
    weights = init_random(spec, 0)
    eng = B200Engine(spec=spec, weights=weights)
    yield spec, weights, eng
    eng.close()


def _meta_rows(y):
    n = len(y["meta_0"])
    return [[y["meta_0"][i].decode()] + [y[f"meta_{k}"][i].decode() for k in range(1, 10)] for i in range(n)]


def test_library_is_loaded_and_launches_kernels(standin):
    from jaeger_b200 import _cabi
    assert _cabi.LIB_PATH.exists()
    maps = Path("/proc/self/maps").read_text()
    assert "libjaeger_b200.so" in maps
    _, _, eng = standin
    n0 = eng.ctx.launch_count
    from jaeger_b200 import WindowSource
    eng.predict(WindowSource(fasta=G / "synthetic_contigs.fasta"))
    assert eng.ctx.launch_count - n0 >= 2 + 2 + 17 + 1           # pack, encode, pool fill + expand, 17 convs (mask fused), heads


@pytest.mark.parametrize("key", ["2000_1500_None_None_0", "2048_2048_None_None_0", "500_500_None_None_0", "2000_2000_None_None_1"])
def test_window_metadata_bit_exact_vs_reference_goldens(standin, key):
    """meta_0..meta_9 of every window equal what the reference's fragment_generator emits."""
    from jaeger_b200 import WindowSource
    _, _, eng = standin
    fs, st, _, _, dyn = key.split("_")
    gold = json.loads((G / "fragments_synthetic.json").read_text())[key]
    y = eng.predict(WindowSource(fasta=G / "synthetic_contigs.fasta", fsize=int(fs), stride=int(st), dynamic_stride=bool(int(dyn))))
    got = _meta_rows(y)
    assert len(got) == len(gold)
    for g_row, r_row in zip(got, gold):
        assert g_row == r_row[2:], (g_row, r_row)


def test_two_pass_short_contigs_order_and_metadata(standin):
    from jaeger_b200 import WindowSource
    _, _, eng = standin
    gold = json.loads((G / "fragments_synthetic.json").read_text())
    y = eng.predict(WindowSource(fasta=G / "synthetic_contigs.fasta", fsize=2000, stride=1500, min_len=500))
    want = gold["2000_1500_None_None_0"] + gold["2000_1500_500_1999_0"]          # long pass first (predict.py:806-810)
    assert [r[2:] for r in want] == _meta_rows(y)


def _encode_seqs(eng, seqs, crop, lut=None, case_sensitive=None, soft=None):
    buf = "".join(seqs).encode()
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)])
    with torch.cuda.stream(eng._stream()):
        dev = torch.from_numpy(np.frombuffer(buf, dtype=np.uint8).copy()).to(eng.tdev)
        codes, valid = eng.pack(dev)
        lc = max(eng.codons_per_frame(int(n), crop) for n in lens)
        tok, counts, skew = eng.encode(codes, valid, eng._h2d(off[:-1]), eng._h2d(lens.astype(np.int32)), crop, lc,
                                       soft=soft, lut=lut, case_sensitive=case_sensitive)
        out = tok.cpu().numpy()[:, :, :lc], counts.cpu().numpy(), skew.cpu().numpy()
    eng.ctx.sync()
    return out


def test_tokens_bit_exact_vs_reference_encoder_golden(standin):
    _, _, eng = standin
    z = np.load(G / "tokens_2000.npz")
    tok, _, _ = _encode_seqs(eng, [str(s) for s in z["seqs"]], 2000)
    assert np.array_equal(tok, z["tokens"])
    z = np.load(G / "tokens_more_crops.npz")                       # crops 2048 / 500 from the same reference encoder
    for crop in (2048, 500):
        tok, _, _ = _encode_seqs(eng, [str(s) for s in z[f"seqs_{crop}"]], crop)
        assert np.array_equal(tok[:, :, :z[f"tokens_{crop}"].shape[2]], z[f"tokens_{crop}"]), crop


@pytest.mark.parametrize("crop", [2000, 2048, 1500, 500, 301])
def test_tokens_counts_skew_bit_exact_vs_oracle(standin, crop):
    from oracle import encode as oenc
    from oracle import seqwin
    from tests.helpers import random_contigs
    _, _, eng = standin
    recs = random_contigs(crop, [crop] * 24 + [crop - 7, crop - 100, max(6, crop // 3), 5, 3], n_run_every=3, lower_every=4)
    seqs = [s for _, s in recs]
    tok, counts, skew = _encode_seqs(eng, seqs, crop)
    ref = oenc.encode_windows([s.upper() for s in seqs], crop)
    assert tok.shape == ref.shape and np.array_equal(tok, ref)
    for i, s in enumerate(seqs):
        u = s.upper()
        g, c, a, t = (u.count(x) for x in "GCAT")
        assert counts[i].tolist() == [g, c, a, t]
        want = f"{seqwin.safe_divide(g - c, g + c): .3f}"
        got = "-0.000" if skew[i] == (1 << 14) else f"{skew[i] / 100: .3f}"
        assert got == want, (i, g, c, got, want)


def test_gc_skew_rounding_exhaustive_small_counts(standin):
    """round((g-c)/(g+c), 2) with Python semantics for every (g, c) up to 60 -- all the half-way
    cases (x.xx5) decide by the binary value of the quotient."""
    from oracle import seqwin
    _, _, eng = standin
    seqs, want = [], []
    for g in range(0, 61, 1):
        for c in range(0, 61, 3):
            seqs.append("G" * g + "C" * c + "A" * 7)
            want.append(f"{seqwin.safe_divide(g - c, g + c): .3f}")
    _, counts, skew = _encode_seqs(eng, seqs, 200)
    got = ["-0.000" if v == (1 << 14) else f"{v / 100: .3f}" for v in skew]
    assert got == want


def test_reduced_alphabet_and_legacy_luts(standin):
    from jaeger_b200 import codon_tables as ct
    from oracle import encode as oenc
    from tests.helpers import random_contigs
    _, _, eng = standin
    seqs = [s for _, s in random_contigs(5, [2000] * 6, lower_every=2)]
    for name in ("AA_ID", "MURPHY10_ID", "PC5_ID"):
        tok, _, _ = _encode_seqs(eng, seqs, 2000, lut=ct.device_lut(ct.TABLES[name]))
        assert np.array_equal(tok, oenc.encode_windows(seqs, 2000, codon_id=oenc.CODON_MAPS[name]))
    # legacy `default` encoder: amino-acid ids, unknown -> 0, soft-masked (lower-case) bases unknown
    masks = [np.array([ch.islower() for ch in s]) for s in seqs]
    bits = np.concatenate(masks)
    packed = np.packbits(bits, bitorder="little")
    words = np.zeros((len(bits) + 31) // 32 * 4 + 16, np.uint8)
    words[:len(packed)] = packed
    soft = eng._h2d(words.view(np.int32))
    tok, counts, _ = _encode_seqs(eng, seqs, 2000, lut=ct.device_lut(ct.LEGACY_AA_ID, plus_one=False), case_sensitive=1, soft=soft)
    table = dict(zip(oenc.CODONS, ct.LEGACY_AA_ID))
    soft_seqs = ["".join(ch.lower() if m else ch.upper() for ch, m in zip(s, mk)) for s, mk in zip(seqs, masks)]
    ref = np.stack([oenc.encode_window_legacy(s, 2000, table) for s in soft_seqs]).astype(np.uint8)
    assert np.array_equal(tok, ref)
    for i, s in enumerate(soft_seqs):
        assert counts[i].tolist() == [s.count(x) for x in "GCAT"]          # upper-case only (io.py:124-127)


def test_forward_vs_oracle_long_and_short_pass(standin):
    from jaeger_b200 import B200Engine, WindowSource
    from oracle import encode as oenc
    from oracle import forward as ofw
    from oracle import seqwin
    from tests.helpers import random_contigs
    spec, weights, eng = standin
    recs = random_contigs(1, [2000, 3500, 5000, 9000, 2100, 12000, 800, 1500, 700, 1999])
    src = WindowSource(records=recs, fsize=2000, stride=1500, min_len=500, batch=3)
    y = eng.predict(src)
    long_w = list(seqwin.fragment_windows(recs, 2000, 1500, min_len=2000))
    short_w = list(seqwin.fragment_windows(recs, 2000, 1500, min_len=500, max_len=1999))
    n = len(long_w)
    ref = ofw.forward(spec, weights, oenc.encode_windows([w.seq for w in long_w], 2000))
    assert np.abs(ref["prediction"] - y["prediction"][:n]).max() <= 4e-3
    assert np.abs(ref["embedding"] - y["embedding"][:n]).max() <= 1e-2
    assert np.abs(ref["nmd"] - y["nmd"][:n]).max() <= 4e-3
    assert np.abs(ref["reliability"] - y["reliability"][:n]).max() <= 4e-3
    assert (ref["prediction"].argmax(1) == y["prediction"][:n].argmax(1)).mean() >= 0.999
    # short pass: padded batches of 3 (the reference's padded_batch), each padded to its longest member
    for b in range(0, len(short_w), 3):
        tok = oenc.encode_windows([w.seq for w in short_w[b:b + 3]], 2000)
        r = ofw.forward(spec, weights, tok)
        assert np.abs(r["prediction"] - y["prediction"][n + b:n + b + 3]).max() <= 4e-3
        assert np.abs(r["embedding"] - y["embedding"][n + b:n + b + 3]).max() <= 1e-2
    # tensor-core kernels vs the CUDA-core restatement of the same layer contract
    eng_ref = B200Engine(spec=spec, weights=weights, use_ref_kernels=True)
    y2 = eng_ref.predict(src)
    eng_ref.close()
    for k in ("prediction", "embedding", "nmd", "reliability"):
        assert np.abs(y[k] - y2[k]).max() <= 2e-3, k


def test_masked_windows_all_unknown_and_long_n_runs(standin):
    """Edge cases of the mask semantics: a window of only N (all-masked sample pools to zeros,
    layers.py:517-529), a window with a 600 bp N run, IUPAC codes."""
    from jaeger_b200 import WindowSource
    from oracle import encode as oenc
    from oracle import forward as ofw
    spec, weights, eng = standin
    rng = np.random.default_rng(3)
    base = "".join(rng.choice(list("ACGT"), 2000))
    s1 = "N" * 2000
    s2 = base[:700] + "N" * 600 + base[1300:]
    s3 = base[:10] + "RYKM" + base[14:]
    y = eng.predict(WindowSource(records=[("a", s1), ("b", s2), ("c", s3)], fsize=2000, stride=2000))
    ref = ofw.forward(spec, weights, oenc.encode_windows([s1, s2, s3], 2000))
    assert np.abs(ref["prediction"] - y["prediction"]).max() <= 4e-3
    assert np.abs(ref["embedding"] - y["embedding"]).max() <= 1e-2
    assert np.all(y["embedding"][0] == 0.0)                      # all-masked -> zero features
    assert np.array_equal(y["prediction"][0], np.asarray(weights["classifier"][0]["bias"], dtype=np.float32))


def test_other_architectures_avg_pool_no_masking_relu(standin):
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project, standin_1p4m_config
    from oracle import encode as oenc
    from oracle import forward as ofw
    from tests.helpers import random_contigs
    recs = random_contigs(9, [2000, 2000, 4000])
    for pooling, masking, act in [("average", True, "relu"), ("max", False, "gelu")]:
        cfg = standin_1p4m_config()
        m = cfg["model"]
        m["use_masking"] = masking
        m["representation_learner"]["pooling"] = pooling
        hl = m["representation_learner"]["hidden_layers"][:8]
        for layer in hl:
            if layer["name"] == "activation":
                layer["config"]["activation"] = act
        hl[4] = {"name": "residual_block", "config": {"block_size": 1, "filters": 128, "kernel_size": 3, "dilation_rate": 8, "activation": act}}
        m["representation_learner"]["hidden_layers"] = hl
        m["reliability_model"]["input_shape"] = 256
        spec = parse_project(cfg)
        w = init_random(spec, 4)
        eng = B200Engine(spec=spec, weights=w)
        y = eng.predict(WindowSource(records=recs, fsize=2000, stride=2000))
        from oracle import seqwin
        wins = list(seqwin.fragment_windows(recs, 2000, 2000))
        ref = ofw.forward(spec, w, oenc.encode_windows([x.seq for x in wins], 2000))
        eng.close()
        assert np.abs(ref["prediction"] - y["prediction"]).max() <= 4e-3, (pooling, masking)
        assert np.abs(ref["embedding"] - y["embedding"]).max() <= 1e-2, (pooling, masking)


@pytest.mark.parametrize("masking", [True, False])
def test_masked_dyt_model_family_vs_oracle(masking):
    """The MaskedDYT variant of the architecture (gamma * tanh(alpha x) + beta in place of every BatchNorm,
    re-masked; train_config/nn_config_1500bp_nmd_merge_6_class_zeus.yaml): logits / embeddings / NMD vectors
    against the fp32 oracle, long pass and padded short pass, with unknown-codon runs."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project, standin_1p4m_config
    from oracle import encode as oenc
    from oracle import forward as ofw
    from oracle import seqwin
    from tests.helpers import random_contigs, to_dyt
    cfg = to_dyt(standin_1p4m_config())
    cfg["model"]["use_masking"] = masking
    spec = parse_project(cfg)
    w = init_random(spec, 6)
    recs = random_contigs(31, [2000, 6500, 9000, 2300, 1400, 800])
    eng = B200Engine(spec=spec, weights=w)
    y = eng.predict(WindowSource(records=recs, fsize=2000, stride=1500, min_len=500, batch=2))
    eng_ref = B200Engine(spec=spec, weights=w, use_ref_kernels=True)
    y2 = eng_ref.predict(WindowSource(records=recs, fsize=2000, stride=1500, min_len=500, batch=2))
    eng.close(); eng_ref.close()
    long_w = list(seqwin.fragment_windows(recs, 2000, 1500, min_len=2000))
    n = len(long_w)
    ref = ofw.forward(spec, w, oenc.encode_windows([x.seq for x in long_w], 2000))
    for k, tol in (("prediction", 4e-3), ("embedding", 1e-2), ("nmd", 4e-3), ("reliability", 4e-3)):
        assert np.abs(ref[k] - y[k][:n]).max() <= tol, (k, np.abs(ref[k] - y[k][:n]).max())
        assert np.abs(y[k] - y2[k]).max() <= 4e-3, k                     # tensor-core path vs CUDA-core restatement
    short_w = list(seqwin.fragment_windows(recs, 2000, 1500, min_len=500, max_len=1999))
    r = ofw.forward(spec, w, oenc.encode_windows([x.seq for x in short_w], 2000))
    assert np.abs(r["prediction"] - y["prediction"][n:]).max() <= 4e-3


@pytest.mark.parametrize("ksize", [2, 3])
def test_wide_256_channel_layer_with_accumulator_reuse(ksize):
    """A 128 -> 256 channel convolution has only two TMEM accumulators; with several tiles per SM
    they are reused while three epilogue groups rotate (the case a phase-parity wait can get wrong).
    k = 2 fits the tensor-core kernel's shared memory; k = 3 does not and takes the CUDA-core kernel."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project, standin_1p4m_config
    from oracle import encode as oenc
    from oracle import forward as ofw
    from oracle import seqwin
    from tests.helpers import random_contigs
    cfg = standin_1p4m_config()
    m = cfg["model"]
    hl = m["representation_learner"]["hidden_layers"][:4]
    hl += [{"name": "masked_conv1d", "config": {"filters": 256, "kernel_size": ksize, "strides": 1, "dilation_rate": 2,
                                                  "use_bias": True, "activation": None}},
           {"name": "masked_batchnorm", "config": {"return_nmd": False}}, {"name": "activation", "config": {"activation": "gelu"}}]
    m["representation_learner"]["hidden_layers"] = hl
    m["classifier"]["input_shape"] = 256
    m["reliability_model"]["input_shape"] = 128
    spec = parse_project(cfg)
    w = init_random(spec, 4)
    recs = random_contigs(21, [2000] * 4 + [30000, 50000, 41000])      # 85 windows = 2720 tiles, ~18 per SM
    eng = B200Engine(spec=spec, weights=w)
    y = eng.predict(WindowSource(records=recs, fsize=2000, stride=1500))
    eng.close()
    wins = list(seqwin.fragment_windows(recs, 2000, 1500))
    assert y["prediction"].shape[0] == len(wins)
    ref = ofw.forward(spec, w, oenc.encode_windows([x.seq for x in wins], 2000))
    assert np.abs(ref["prediction"] - y["prediction"]).max() <= 4e-3
    assert np.abs(ref["embedding"] - y["embedding"]).max() <= 1e-2
    assert np.abs(ref["nmd"] - y["nmd"]).max() <= 4e-3


def test_window_independence_and_chunking_at_scale(standin):
    """Size-independent property at a realistic size: a window's logits do not depend on which
    forward chunk it lands in (4 000 windows through 3 different workspace budgets)."""
    from jaeger_b200 import WindowSource
    from tests.helpers import random_contigs
    spec, weights, eng = standin
    recs = random_contigs(21, [50000] * 120, n_run_every=5)            # 120 x 33 windows = 3 960
    outs = []
    for gb in (16.0, 3.0, 1.1):
        eng.workspace_bytes = int(gb * (1 << 30))
        outs.append(eng.predict(WindowSource(records=recs, fsize=2000, stride=1500))["prediction"])
    eng.workspace_bytes = int(16 * (1 << 30))
    assert outs[0].shape == (3960, 6)
    # chunk boundaries move tiles between CTAs; atomics in the taps/pool are order-dependent only in fp32 sums
    assert np.abs(outs[0] - outs[1]).max() <= 1e-5 and np.abs(outs[0] - outs[2]).max() <= 1e-5
    # reverse the contig order: per-window results must follow their contigs
    y_rev = eng.predict(WindowSource(records=recs[::-1], fsize=2000, stride=1500))["prediction"]
    per = 33
    rev_back = np.concatenate([y_rev[(119 - i) * per:(120 - i) * per] for i in range(120)])
    assert np.abs(outs[0] - rev_back).max() <= 1e-5


def test_aggregate_bit_exact_vs_reference_pred_to_dict_golden(standin):
    _, _, eng = standin
    z = np.load(G / "pred_to_dict.npz")
    last = z["meta_2"].astype(np.int32)
    ends = np.flatnonzero(last == 1) + 1
    off = np.concatenate([[0], ends]).astype(np.int64)
    with torch.cuda.stream(eng._stream()):
        agg = eng.aggregate(eng._h2d(z["prediction"]), eng._h2d(z["reliability"]), eng._h2d(off))
        agg = {k: v.cpu().numpy() for k, v in agg.items()}
    eng.ctx.sync()
    assert np.array_equal(agg["pred_sum"].view(np.uint16), z["pred_sum"].view(np.uint16))        # fp16 bit-exact
    assert np.array_equal(agg["pred_var"].view(np.uint16), z["pred_var"].view(np.uint16))
    assert np.array_equal(agg["consensus"], z["consensus"])
    assert np.array_equal(agg["frag_pred"], z["frag_pred"])
    assert np.array_equal(agg["per_class_counts"], z["per_class_counts"])
    # transcendental-based columns: fp16 after log2f / exp / log -- allow one fp16 ulp
    for k in ("entropy", "energy"):
        d = np.abs(agg[k].view(np.int16).astype(np.int32) - z[k].view(np.int16).astype(np.int32))
        assert d.max() <= 1, (k, d.max())
    n_win = np.diff(off)
    ood = np.array([f"{k / n:.2f}" for k, n in zip(agg["rel_pos"], n_win)], dtype=np.float16)   # collect.py:233-244
    assert np.array_equal(ood.view(np.uint16), z["ood"].view(np.uint16))


def test_aggregate_random_large_vs_oracle(standin):
    from oracle import postprocess as opp
    _, _, eng = standin
    rng = np.random.default_rng(8)
    n_win = rng.integers(1, 60, size=3000)
    n_win[::500] = 3333                                             # genome-sized contigs
    W = int(n_win.sum())
    pred = rng.normal(0, 3, (W, 6)).astype(np.float32)
    rel = rng.normal(0, 2, (W, 1)).astype(np.float32)
    last = np.zeros(W, np.int32)
    last[np.cumsum(n_win) - 1] = 1
    off = np.concatenate([[0], np.cumsum(n_win)]).astype(np.int64)
    with torch.cuda.stream(eng._stream()):
        agg = {k: v.cpu().numpy() for k, v in eng.aggregate(eng._h2d(pred), eng._h2d(rel), eng._h2d(off)).items()}
    eng.ctx.sync()
    ref = opp.aggregate_numeric(pred, rel, last)
    assert np.array_equal(agg["pred_sum"].view(np.uint16), ref["pred_sum"].view(np.uint16))
    assert np.array_equal(agg["pred_var"].view(np.uint16), ref["pred_var"].view(np.uint16))
    assert np.array_equal(agg["consensus"], ref["consensus"])
    assert np.array_equal(agg["per_class_counts"], ref["per_class_counts"])
    assert np.array_equal(agg["frag_pred"], np.concatenate(ref["frag_pred"]))
    assert int(agg["per_class_counts"].sum()) == W                   # checksum property


def test_smoothing_vs_reference_golden_and_segmentation_vs_oracle(standin):
    import ctypes
    from jaeger_b200._cabi import check, lib
    from oracle import prophage as opro
    _, _, eng = standin
    z = np.load(G / "smooth.npz")
    T = z["logits"].shape[0]
    with torch.cuda.stream(eng._stream()):
        lg = eng._h2d(z["logits"])
        off = eng._h2d(np.array([0, T], dtype=np.int64))
        out = torch.empty((T, 6), dtype=torch.float64, device=eng.tdev)
        check(lib.jg_smooth_scores(eng.ctx.handle, lg.data_ptr(), off.data_ptr(), 1, 6, 4, out.data_ptr()))
        sm = out.cpu().numpy()
        assert np.abs(sm - z["smoothed"]).max() <= 2e-6                       # fp32 softmax in the reference
        sig = out[:, 1].contiguous()
        bk = torch.zeros((9, T), dtype=torch.int32, device=eng.tdev)
        nb = torch.zeros((9,), dtype=torch.int32, device=eng.tdev)
        check(lib.jg_segment_scores(eng.ctx.handle, sig.data_ptr(), T, 3, 9, bk.data_ptr(), nb.data_ptr()))
        bk, nb = bk.cpu().numpy(), nb.cpu().numpy()
    eng.ctx.sync()
    col = sm[:, 1]
    for p in range(9):
        want = opro.optimal_partition(col, float(p + 1), 3)
        assert bk[p, :nb[p]].tolist() == want, p
        assert want[-1] == T and all(b - a >= 3 for a, b in zip([0] + want[:-1], want))
    # exact ties (integer-valued and constant tracks): the kernel keeps the earliest last change point, like the oracle
    rng = np.random.default_rng(3)
    for track in (np.zeros(40), np.round(rng.normal(size=60) * 2.0), np.repeat([0.0, 3.0, 0.0, 3.0], 6)):
        n = len(track)
        with torch.cuda.stream(eng._stream()):
            sig = eng._h2d(track.astype(np.float64))
            bk = torch.zeros((9, n), dtype=torch.int32, device=eng.tdev)
            nb = torch.zeros((9,), dtype=torch.int32, device=eng.tdev)
            check(lib.jg_segment_scores(eng.ctx.handle, sig.data_ptr(), n, 3, 9, bk.data_ptr(), nb.data_ptr()))
            bk, nb = bk.cpu().numpy(), nb.cpu().numpy()
        eng.ctx.sync()
        for p in range(9):
            assert bk[p, :nb[p]].tolist() == opro.optimal_partition(track, float(p + 1), 3), (n, p)


def test_predict_driver_tsv_equals_oracle_pipeline(standin, tmp_path):
    """End to end through the reference-style driver: the TSV written from the device aggregates
    equals what the oracle's pred_to_dict + generate_summary produce from the same window logits."""
    import io
    import pandas as pd
    from jaeger_b200 import WindowSource
    from jaeger_b200.postprocess import contig_table, generate_summary
    from jaeger_b200.predict import run_core
    from oracle import postprocess as opp
    spec, weights, eng = standin
    res = run_core(input=str(G / "synthetic_contigs.fasta"), output=str(tmp_path), model="standin", allow_random_weights=True, fsize=2000, stride=1500,
                   min_len=500, overwrite=True)
    gold = json.loads((G / "fragments_synthetic.json").read_text())
    assert res["table"].exists()
    assert res["windows"] == len(gold["2000_1500_None_None_0"]) + len(gold["2000_1500_500_1999_0"])   # long + short pass
    with pytest.raises(FileExistsError):
        run_core(input=str(G / "synthetic_contigs.fasta"), output=str(tmp_path), model="standin", allow_random_weights=True)
    y = eng.predict(WindowSource(fasta=G / "synthetic_contigs.fasta", fsize=2000, stride=1500, min_len=500))
    data = contig_table(eng, y, 2000)
    got = generate_summary(data, eng.class_map["class"], eng.class_map["index"])
    odata, _ = opp.pred_to_dict(y, 2000, eng.class_map)
    want = opp.generate_summary(odata, eng.class_map["class"], eng.class_map["index"])
    assert list(got.columns) == list(want.columns)
    for col in want.columns:
        if want[col].dtype.kind in "fc" or str(want[col].dtype) == "float16":
            a, b = got[col].to_numpy(dtype=np.float64), want[col].to_numpy(dtype=np.float64)
            assert np.allclose(a, b, rtol=0, atol=1e-3 if col in ("entropy", "energy") else 1e-12), col
        else:
            assert got[col].tolist() == want[col].tolist(), col
    assert "syn7,with,commas" in got["contig_id"].tolist()                 # "___" restored (collect.py:556)
    tsv = pd.read_csv(res["table"], sep="\t")
    assert len(tsv) == (got["N%"] < 0.3).sum()


def test_prophage_region_calling_vs_oracle(standin):
    from jaeger_b200 import prophage as ppro
    from oracle import prophage as opro
    _, _, eng = standin
    rng = np.random.default_rng(17)
    for T, islands in [(3333, [(400, 430), (1500, 1530), (2800, 2825)]), (700, [(100, 140)]), (350, [])]:
        logits = rng.normal(0, 1.0, (T, 6)).astype(np.float32)
        for a, b in islands:
            logits[a:b, 1] += 7.0
        ranges, scores = ppro.segment_contig(eng, logits, 1, 1.5)
        want_r, want_s = opro.segment(opro.smooth_scores(logits)[:, 1], 1.5)
        assert ranges == want_r, (T, ranges, want_r)
        assert np.allclose(scores, want_s, atol=1e-6)
        if islands:
            # every planted island is called (noise may add short ranges, depending on the penalty the knee selects)
            assert all(any(abs(r[0] - a) <= 12 and abs(r[1] - b) <= 12 for r in ranges) for a, b in islands), (ranges, islands)


def test_config4_genome_with_prophage_option_end_to_end(standin, tmp_path):
    """BASELINE config 4: one 5 Mbp record through the driver with -p.  3333 windows, one TSV row,
    the prophage table equal to the oracle's segmentation of the same window logits."""
    from jaeger_b200.predict import run_core
    from oracle import prophage as opro
    _, _, eng = standin
    rng = np.random.default_rng(404)
    n = 5_000_000
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    seq[1_234_000:1_236_500] = ord("N")
    fa = tmp_path / "genome.fna"
    with open(fa, "wb") as fh:
        fh.write(b">chr1 synthetic genome\n")
        for i in range(0, n, 80):
            fh.write(seq[i:i + 80].tobytes() + b"\n")
    res = run_core(input=str(fa), output=str(tmp_path / "out"), model="standin", allow_random_weights=True, fsize=2000, stride=1500,
                   prophage=True, lc=500_000, sensitivity=1.5, overwrite=True, window_scores=True)
    assert res["num"] == 1 and res["num_written"] == 1
    assert res["windows"] == (n - 2000) // 1500 + 1
    z = np.load(tmp_path / "out" / "standin" / "genome_window_scores.npz", allow_pickle=True)
    logits = np.asarray(z["predictions"][0], dtype=np.float32)
    assert logits.shape[0] == res["windows"]
    cm = eng.class_map
    k = cm["index"][[c.lower() for c in cm["class"]].index("phage")]
    want_r, want_s = opro.segment(opro.smooth_scores(logits)[:, k], 1.5)
    got = res["prophage_regions"].get("chr1", {"ranges": [], "scores": np.array([])})
    assert got["ranges"] == want_r
    assert np.allclose(got["scores"], want_s, atol=1e-6)
    assert (tmp_path / "out" / "standin" / "genome_prophage_regions.tsv").exists()
    # the att-site report: one row per called region, raw coordinates = [start*stride, (end-1)*stride + fsize]
    rep = res["prophage_report"]
    assert len(rep) == len(want_r)
    assert rep["raw_start"].tolist() == [s * 1500 for s, _ in want_r] and rep["raw_end"].tolist() == [(e - 1) * 1500 + 2000 for _, e in want_r]
    assert (tmp_path / "out" / "standin" / "genome_prophages" / "prophages_jaeger.tsv").exists() == (len(want_r) > 0)
    # --genes: a gene across every region end moves the reported ends out of it (prophage_boundaries.py), raw_* columns stay
    if want_r:
        genes = tmp_path / "genes.bed"
        with open(genes, "w") as fh:
            for s, e in want_r:
                fh.write(f"chr1\t{s * 1500 - 700}\t{s * 1500 + 300}\n")
                fh.write(f"chr1\t{(e - 1) * 1500 + 2000 - 100}\t{(e - 1) * 1500 + 2000 + 900}\n")
        res2 = run_core(input=str(fa), output=str(tmp_path / "out2"), model="standin", allow_random_weights=True, fsize=2000, stride=1500,
                        prophage=True, lc=500_000, sensitivity=1.5, overwrite=True, genes=str(genes))
        ref_b = res2["refined_boundaries"]["chr1"]
        assert [(r[2], r[3]) for r in ref_b] == [(max(r[0] - 700, 0), min(r[1] + 900, n)) for r in ref_b]
        assert res2["prophage_report"]["raw_start"].tolist() == rep["raw_start"].tolist()


def test_crf_viterbi_decoding_vs_reference_golden(standin):
    """--crf: device Viterbi paths and per-contig class counts are index-exact against the
    reference's viterbi_decode / pred_to_dict(crf_switch_cost=2.0) goldens, on every cost variant."""
    import torch
    from jaeger_b200.postprocess import build_transition_costs, contig_table
    _, _, eng = standin
    z = np.load(G / "viterbi.npz")
    classes = ["bacteria", "phage", "eukarya", "archaea", "plasmid", "virus"]
    pred, n_win = z["prediction"], z["n_win"]
    off = np.concatenate([[0], np.cumsum(n_win)]).astype(np.int64)
    potts = lambda lam: lam * (1.0 - np.eye(6))
    with torch.cuda.stream(eng._stream()):
        pd_, od = eng._h2d(pred), eng._h2d(off)
        for name, costs in [("potts2", potts(2.0)), ("potts0", potts(0.0)), ("bio2", build_transition_costs(classes, 2.0)),
                            ("uni05", build_transition_costs(classes, 0.5, "uniform")), ("user3", z["costs_user3"])]:
            path, counts = eng.viterbi(pd_, od, costs)
            path, counts = path.cpu().numpy(), counts.cpu().numpy()
            assert np.array_equal(path, z[f"path_{name}"]), name
            assert np.array_equal(counts, np.stack([np.bincount(path[a:b], minlength=6) for a, b in zip(off[:-1], off[1:])])), name
        b = z["binary_logit"]
        two = np.concatenate([np.zeros_like(b), b], -1)
        path, _ = eng.viterbi(eng._h2d(two), eng._h2d(np.array([0, len(b)], np.int64)), 2.0 * (1.0 - np.eye(2)))
        assert np.array_equal(path.cpu().numpy(), z["binary_path"])
    assert np.array_equal(build_transition_costs(classes, 2.0), z["costs_bio2"])
    # through the driver-level table
    meta = {f"meta_{i}": [] for i in range(10)}
    for ci, n in enumerate(n_win):
        for j in range(n):
            for i, v in enumerate([f"c{ci}", j * 1500, int(j == n - 1), j, 2000 + 1500 * (n - 1), 500, 500, 500, 500, " 0.000"]):
                meta[f"meta_{i}"].append(str(v).encode())
    y = {"prediction": pred, **{k: np.array(v) for k, v in meta.items()}}
    data = contig_table(eng, y, 2000, crf_switch_cost=2.0)
    assert np.array_equal(data["per_class_counts"], z["crf_counts"])
    assert np.array_equal(data["frag_pred"], z["crf_frag_pred"])


def test_legacy_postprocess_tables_vs_reference_golden(tmp_path):
    """Legacy `default` post-processing on the device (aggregation + reliability kernel): the TSV
    equals the reference's generate_summary_legacy output for both label sets; per-window reliability
    within 1e-6 of the pickled scikit-learn model (float32 feature normalisation, float64 logit)."""
    import pandas as pd
    from jaeger_b200 import B200Engine, legacy
    from jaeger_b200.postprocess import contig_table_legacy, write_output_legacy
    z = np.load(G / "legacy_post.npz")
    ood = {k[4:]: z[k] for k in z.files if k.startswith("ood_") and k != "ood_windows"}
    y = {"prediction": z["output"], "embedding": z["embedding"], **{f"meta_{i}": z[f"meta_{i}"] for i in range(10)}}
    eng = B200Engine(legacy_weights=legacy.random_weights(0))
    data = contig_table_legacy(eng, y, 2000, ood)
    assert np.abs(data["ood_windows"] - z["ood_windows"]).max() <= 1e-6
    assert np.array_equal(data["pred_sum"], z["pred_sum"]) and np.array_equal(data["pred_var"], z["pred_var"])
    assert np.array_equal(data["consensus"], z["consensus"]) and np.array_equal(data["entropy"], z["entropy"])
    for tag, labels in (("default", legacy.DEFAULT_LABELS), ("all", legacy.ALL_LABELS)):
        n = write_output_legacy(data, [labels[i] for i in range(4)], tmp_path / f"{tag}.tsv", tmp_path / f"{tag}_phages.tsv")
        got = pd.read_csv(tmp_path / f"{tag}.tsv", sep="\t", keep_default_na=False)
        want = pd.read_csv(G / f"summary_legacy_{tag}.tsv", sep="\t", keep_default_na=False)
        assert n == len(want) and list(got.columns) == list(want.columns)
        for col in want.columns:
            assert got[col].tolist() == want[col].tolist(), (tag, col)
        phages = pd.read_csv(tmp_path / f"{tag}_phages.tsv", sep="\t", keep_default_na=False)
        sel = want[(want["prediction"] == "phage") & (want["phage_score"] > 3) & (want["reliability_score"] > 0.5)]
        assert phages["contig_id"].tolist() == sel["contig_id"].tolist()
    eng.close()


def _repeat_contigs():
    from tests.helpers import repeat_contigs
    return repeat_contigs()


def test_terminal_repeat_scan_vs_oracle(standin):
    """device Smith-Waterman scans (DTR / ITR / LTR, mismatch, gaps in either line,
    N runs, soft-masked letters, scan lengths 400..2000) against the oracle restatement -- every field
    of the reference's table is index-exact.  parasail itself: parity unpinned (oracle/termini.py)."""
    import torch
    from jaeger_b200 import WindowSource
    from jaeger_b200.termini import scan_source
    from oracle import termini as ot
    _, _, eng = standin
    recs = _repeat_contigs()
    got = scan_source(eng, WindowSource(records=recs, fsize=2000, stride=1500), 2000)
    want = ot.scan_for_terminal_repeats(recs, 2000)
    assert len(got) == len(want) == len(recs) - 1                     # the 900 bp record is skipped
    kinds = []
    for g, w in zip(got.to_dict("records"), want):
        for k, v in w.items():
            gv = g[k]
            if v is None:
                assert gv is None or gv != gv, (w["contig_id"], k, gv)
            elif isinstance(v, float):
                assert abs(gv - v) < 1e-12, (w["contig_id"], k)
            else:
                assert gv == v, (w["contig_id"], k, gv, v)
        kinds.append(w["terminal_repeats"])
    assert kinds[:7] == [None, "DTR", "ITR", "DTR", "DTR", "DTR", "LTR_DTR"]
    # raw scan parity on random pairs, square and rectangular (ties between equal maxima resolved identically)
    from jaeger_b200.termini import JOB, TRACE_JOB, _run_scan, _run_trace, alignment_lines
    rng = np.random.default_rng(5)
    seq = "".join(rng.choice(list("ACGTN"), 6000, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
    # (q0, n, inverted, nq): nq = 0 -> square
    spans = [(0, 400, 0, 0), (0, 400, 1, 0), (100, 517, 0, 0), (100, 517, 1, 0), (0, 1000, 0, 0), (3, 33, 0, 0), (2000, 1999, 1, 0),
             (50, 1, 0, 0), (10, 700, 0, 300), (10, 300, 1, 900), (500, 1200, 1, 17), (0, 40, 0, 1500), (7, 1, 0, 5), (9, 5, 1, 1)]
    jobs = np.zeros(len(spans), dtype=JOB)
    for k, (q0, n, inv, nq) in enumerate(spans):
        jobs[k] = (q0, 6000 - n - q0 // 2, n, inv, nq, 0)
    with torch.cuda.stream(eng._stream()):
        codes, valid = eng.pack(eng._h2d(np.frombuffer(seq.encode(), np.uint8).copy()))
    res = _run_scan(eng, codes, valid, jobs)
    for k, (q0, n, inv, nq) in enumerate(spans):
        r0 = 6000 - n - q0 // 2
        ref = seq[r0:r0 + n]
        a = ot.sw_align(seq[q0:q0 + (nq or n)], ot.reverse_complement(ref) if inv else ref)
        assert (res[k, 0], res[k, 1], res[k, 2]) == (a["score"], a["end_query"], a["end_ref"]), (k, res[k], a)
        if a["score"] < 104:
            assert res[k, 3] == a["cols"], (k, res[k], a)


def test_sw_trace_operations_give_the_alignment_lines(standin):
    """Rectangular traced jobs: counts and the per-column operations of jg_sw_trace rebuild exactly the oracle's
    traceback.query / traceback.ref lines (gaps in either line, mismatches, direct and inverted, lower-case letters)."""
    from jaeger_b200.termini import JOB, TRACE_JOB, JOB_FIELDS, _run_scan, _run_trace, alignment_lines
    from oracle import termini as ot
    _, _, eng = standin
    rng = np.random.default_rng(21)
    rnd = lambda n: "".join(rng.choice(list("ACGT"), n))
    core = rnd(400)
    mut = lambda s, i: s[:i] + ("A" if s[i] != "A" else "C") + s[i + 1:]
    pairs = [(rnd(300) + core + rnd(200), rnd(100) + core[:150] + core[152:] + rnd(350), 0),               # gap in the ref line
             (rnd(50) + core[:200] + core[203:] + rnd(700), rnd(900) + core + rnd(20), 0),                    # gap in the query line
             (rnd(10) + mut(core, 199).lower() + rnd(10), rnd(400) + core + rnd(333), 1),                     # mismatch, inverted, lower-case query
             (rnd(77) + core + rnd(5), rnd(31) + mut(core[:120], 60) + "GT" + core[120:250] + core[254:] + rnd(600), 1)]
    parts, jobs, pos = [], np.zeros(len(pairs), dtype=JOB), 0
    for k, (q, r, inv) in enumerate(pairs):
        stored = ot.reverse_complement(r) if inv else r
        jobs[k] = (pos, pos + len(q) + 5, len(r), inv, len(q), 0)
        parts += [q, "NNNNN", stored, "N" * 7]
        pos += len(q) + 5 + len(r) + 7
    host = np.frombuffer("".join(parts).encode(), np.uint8).copy()
    with torch.cuda.stream(eng._stream()):
        codes, valid = eng.pack(eng._h2d(host))
    res = _run_scan(eng, codes, valid, jobs)
    tj = np.zeros(len(pairs), dtype=TRACE_JOB)
    for f in JOB_FIELDS:
        tj[f] = jobs[f]
    tj["end_i"], tj["end_j"] = res[:, 1], res[:, 2]
    counts, ops = _run_trace(eng, codes, valid, tj, want_ops=True)
    counts2, none = _run_trace(eng, codes, valid, tj.copy())
    assert none is None and np.array_equal(counts, counts2)
    for k, (q, r, inv) in enumerate(pairs):
        a = ot.sw_align(q, r)
        assert a["score"] >= 104 and a["qgaps"] + a["rgaps"] + (a["cols"] - a["iden"]) >= 1
        assert tuple(res[k, :3]) == (a["score"], a["end_query"], a["end_ref"]), (k, res[k])
        assert tuple(counts[k]) == (a["cols"], a["qgaps"], a["rgaps"], a["iden"]), (k, counts[k], a)
        assert alignment_lines(host, jobs[k], int(res[k, 1]), int(res[k, 2]), ops[k]) == (a["qline"], a["rline"]), k


def test_prophage_att_report_vs_reference_golden(standin, tmp_path):
    """SURVEY.md 8f-4 (att sites): the device att-site search + report for synthetic called regions equals, byte for
    byte, the `prophages_jaeger.tsv` the reference's prophage_report wrote for the same genomes and coordinates with
    its aligner stubbed by the oracle (tests/golden/make_termini_goldens.py) -- exact DTR, traced ITR with a mismatch,
    traced DTR with a gap around a 23 %-N region (reject), a spontaneous 13-mer at the contig start."""
    from jaeger_b200 import WindowSource
    from jaeger_b200.termini import prophage_report_loaded, write_prophage_report
    from tests.helpers import prophage_genomes
    _, _, eng = standin
    recs, cords = prophage_genomes()
    regions = {k: {"ranges": v[0], "scores": v[1]} for k, v in cords.items()}
    src = WindowSource(records=recs, fsize=2000, stride=1500)
    rep = prophage_report_loaded(eng, src.load(), regions, 2000, 1500)
    assert len(rep) == 5 and rep["att_type"].tolist() == ["DTR", "ITR", "DTR", "DTR", "DTR"]
    write_prophage_report(rep, tmp_path / "x_prophages")
    assert (tmp_path / "x_prophages" / "prophages_jaeger.tsv").read_text() == (G / "prophages_jaeger.tsv").read_text()
    # gene-aware ends (prophage_boundaries.py): the ends the reference's own refine_prophage_boundaries produced for a fixed gene
    # table, the report the reference wrote on them -- the refined ends bring a planted 40-bp repeat into reach
    from jaeger_b200 import prophage_boundaries as pb
    from tests.helpers import prophage_gene_calls
    calls = prophage_gene_calls()
    names, _, offsets = src.load()
    refined = pb.refine_regions(regions, names, np.diff(offsets), 2000, 1500, lambda header, ci: calls.get(header))
    assert {k: [list(r) for r in v] for k, v in refined.items()} == json.loads((G / "refined_boundaries.json").read_text())
    rep2 = prophage_report_loaded(eng, src.load(), regions, 2000, 1500, refined_boundaries=refined)
    write_prophage_report(rep2, tmp_path / "z_prophages")
    assert (tmp_path / "z_prophages" / "prophages_jaeger.tsv").read_text() == (G / "prophages_jaeger_refined.tsv").read_text()
    assert int(rep2["att_alignment_length"].iloc[4]) > int(rep["att_alignment_length"].iloc[4])
    # nothing on contigs <= 500 kbp, no file without rows (prophages.py:759, 866)
    none = prophage_report_loaded(eng, src.load(), {"plasmid": regions["plasmid"]}, 2000, 1500)
    assert len(none) == 0
    write_prophage_report(none, tmp_path / "y_prophages")
    assert not (tmp_path / "y_prophages").exists()


def test_driver_tsv_carries_terminal_repeat_columns(standin, tmp_path):
    import pandas as pd
    from jaeger_b200.predict import run_core
    recs = _repeat_contigs()
    fa = tmp_path / "rep.fasta"
    fa.write_text("".join(f">{n}\n{s}\n" for n, s in recs))
    res = run_core(input=str(fa), output=str(tmp_path / "o"), model="standin", allow_random_weights=True, fsize=2000, stride=1500, overwrite=True, dustmask=False)
    tsv = pd.read_csv(res["table"], sep="\t", keep_default_na=False).set_index("contig_id")
    assert tsv.loc["dtr", "terminal_repeats"] == "DTR" and int(float(tsv.loc["dtr", "repeat_length"])) == 120
    assert tsv.loc["itr,comma", "terminal_repeats"] == "ITR"
    assert tsv.loc["ltr", "terminal_repeats"] == "LTR_DTR" and int(float(tsv.loc["ltr", "repeat_length"])) == 700
    assert tsv.loc["plain", "terminal_repeats"] == ""


def _legacy_fixture():
    from jaeger_b200.weights import load_npz_weights
    z = np.load(G / "legacy_default.npz")
    flat = {k[2:]: z[k] for k in z.files if k.startswith("w/")}
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".npz") as fh:
        np.savez(fh.name, **flat)
        w = load_npz_weights(fh.name)
    recs = [(str(n), str(s)) for n, s in zip(z["names"], z["seqs"])]
    return w, recs


def test_legacy_default_model_on_health_fasta_vs_oracle():
    """BASELINE config 1: the bundled `default` weights on the reference's health FASTA
    (135 windows at the CLI defaults).  Tolerance: fp16-stored activations (fp32 epilogues: this graph's BatchNorm follows the
    GELU with scales up to 26) vs the fp32 oracle with real weights, |logit| ~ 3-10: max |diff| <= 0.03 (observed 0.009)
    and identical per-window / per-contig labels."""
    from jaeger_b200 import B200Engine, WindowSource
    from jaeger_b200 import codon_tables as ct
    from jaeger_b200.postprocess import contig_table
    from oracle import encode as oenc
    from oracle import legacy as oleg
    from oracle import postprocess as opp
    from oracle import seqwin
    w, recs = _legacy_fixture()
    eng = B200Engine(legacy_weights=w)
    y = eng.predict(WindowSource(records=recs, fsize=2000, stride=1500))
    wins = list(seqwin.fragment_windows(recs, 2000, 1500))
    assert len(wins) == 135 == y["prediction"].shape[0]
    table = dict(zip(oenc.CODONS, ct.LEGACY_AA_ID))
    tok = np.stack([oenc.encode_window_legacy(x.seq, 2000, table) for x in wins]).astype(np.uint8)
    ref = oleg.forward(w, tok)
    d = np.abs(ref["output"] - y["prediction"])
    assert d.max() <= 0.03, d.max()
    assert np.abs(ref["embedding"] - y["embedding"]).max() <= 0.03
    differ = np.flatnonzero(ref["output"].argmax(1) != y["prediction"].argmax(1))
    top2 = np.sort(ref["output"], axis=1)
    margins = top2[:, -1] - top2[:, -2]
    assert len(differ) <= 2 and np.all(margins[differ] < 2 * d.max()), (differ, margins[differ], d.max())   # only near-ties may flip
    print(f"legacy: max|logit diff| {d.max():.4f}, windows with different argmax {len(differ)} (margins {margins[differ]})")
    # ... and against the reference's own serialized TF graph, interpreted op by op (oracle/tfgraph.py -> golden)
    gold = np.load(G / "legacy_graph_outputs.npz")
    assert zlib.crc32(tok.tobytes()) == int(gold["health_token_crc"])
    dg = np.abs(gold["health_output"] - y["prediction"])
    assert dg.max() <= 0.03 and np.abs(gold["health_embedding"] - y["embedding"]).max() <= 0.03
    flips = np.flatnonzero(gold["health_output"].argmax(1) != y["prediction"].argmax(1))
    assert len(flips) <= 2 and np.all(margins[flips] < 2 * dg.max())
    data = contig_table(eng, y, 2000)
    agg = opp.aggregate_numeric(ref["output"], None, np.array([x.is_last for x in wins]))
    assert np.array_equal(data["consensus"], agg["consensus"])           # 9 / 9 contig labels
    eng.close()


def test_legacy_default_driver_end_to_end_tsv(tmp_path):
    """BASELINE config 1 through the driver (`-m default`): FASTA in, `<base>_jaeger.tsv` out.  Parity
    target of SURVEY.md 8d: identical `prediction` column, scores within tolerance -- against the
    oracle's forward pass + the restated legacy tables (fp16 activations: |score diff| <= 0.1,
    reliability within 0.02)."""
    import pandas as pd
    from jaeger_b200 import codon_tables as ct
    from jaeger_b200.predict import run_core
    from jaeger_b200.weights import save_npz_weights
    from oracle import encode as oenc
    from oracle import legacy as oleg
    from oracle import seqwin
    w, recs = _legacy_fixture()
    fa = tmp_path / "health.fasta"
    fa.write_text("".join(f">{n} some description\n" + "\n".join(s[i:i + 70] for i in range(0, len(s), 70)) + "\n" for n, s in recs))
    save_npz_weights(tmp_path / "w.npz", w)
    res = run_core(input=str(fa), output=str(tmp_path / "out"), model="default", legacy_weights=str(tmp_path / "w.npz"),
                   legacy_ood_dir=str(G / "legacy_post.npz"), fsize=2000, stride=1500, dustmask=False, overwrite=True)
    assert res["table"].name == "health_jaeger.tsv" and res["windows"] == 135 and res["num_written"] == 9
    got = pd.read_csv(res["table"], sep="\t", keep_default_na=False)
    wins = list(seqwin.fragment_windows(recs, 2000, 1500))
    table = dict(zip(oenc.CODONS, ct.LEGACY_AA_ID))
    tok = np.stack([oenc.encode_window_legacy(x.seq, 2000, table) for x in wins]).astype(np.uint8)
    ref = oleg.forward(w, tok)
    z = np.load(G / "legacy_post.npz")
    ood = {k[4:]: z[k] for k in z.files if k.startswith("ood_") and k != "ood_windows"}
    meta = [np.array([str(v).encode() for v in col]) for col in zip(*[
        (x.header, x.index, int(x.is_last), x.ordinal, x.seqlen, x.g, x.c, x.a, x.t, x.gc_skew) for x in wins])]
    all_labels = {0: "bacteria", 1: "phage", 2: "eukarya", 3: "archaea"}
    cols, _ = oleg.summary_legacy(ref["output"], ref["embedding"], tuple(meta), 2000, ood,
                                  ["non-phage", "phage", "non-phage", "non-phage"], all_labels,
                                  {1: "eukarya", 2: "archaea", 3: "bacteria", 0: ""}, 1)
    want = pd.DataFrame(cols)
    assert got["contig_id"].tolist() == want["contig_id"].tolist()
    assert got["prediction"].tolist() == want["prediction"].tolist()
    assert got["length"].tolist() == want["length"].tolist()
    for c in ("G+C", "N%"):
        assert np.allclose(got[c].to_numpy(float), want[c].to_numpy(float), atol=5e-4)
    for lab in all_labels.values():
        assert np.abs(got[f"{lab}_score"].to_numpy(float) - want[f"{lab}_score"].to_numpy(float)).max() <= 0.1, lab
        assert np.abs(got[f"#_{lab}_windows"].to_numpy(int) - want[f"#_{lab}_windows"].to_numpy(int)).max() <= 2, lab
    assert np.abs(got["reliability_score"].to_numpy(float) - want["reliability_score"].to_numpy(float)).max() <= 0.02


def test_500bp_baseline_model_config3():
    """BASELINE config 3: the 500 bp / 32-filter / average-pool architecture (random init) on
    500 bp fragments; channels zero-padded to 64 on the device."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project
    from oracle import encode as oenc
    from oracle import forward as ofw
    from tests.helpers import random_contigs
    def conv(f, k):
        return {"name": "masked_conv1d", "config": {"filters": f, "kernel_size": k, "strides": 1, "dilation_rate": 1, "use_bias": True, "activation": None}}
    bn_act = [{"name": "masked_batchnorm", "config": {"return_nmd": False}}, {"name": "activation", "config": {"activation": "gelu"}}]
    cfg = {"model": {"name": "jaeger_500bp_baseline", "activation": "gelu",
                     "class_label_map": [{"class": c, "label": i} for i, c in enumerate(["chromosome", "virus", "plasmid"])],
                     "embedding": {"use_embedding_layer": True, "input_type": "translated", "input_shape": [6, None], "embedding_size": 64},
                     "string_processor": {"seq_onehot": False, "codon": "CODON", "codon_id": "CODON_ID", "crop_size": 500, "masking": False},
                     "representation_learner": {"hidden_layers": [conv(32, 7)] + bn_act + [
                         {"name": "residual_block", "config": {"use_1x1conv": False, "block_size": 2, "filters": 32, "kernel_size": 3, "use_bias": True}}] + bn_act,
                         "pooling": "average"},
                     "classifier": {"input_shape": 32, "hidden_layers": [{"name": "dense", "config": {"units": 3, "activation": None, "use_bias": True}}]}}}
    spec = parse_project(cfg)
    w = init_random(spec, 2)
    eng = B200Engine(spec=spec, weights=w)
    recs = random_contigs(2, [500] * 300 + [499, 1700], n_run_every=7, lower_every=0)
    y = eng.predict(WindowSource(records=recs, fsize=500, stride=500))
    seqs = [s[i:i + 500] for _, s in recs for i in range(0, len(s) - 499, 500)]
    assert y["prediction"].shape == (len(seqs), 3) == (303, 3) and y["embedding"].shape[1] == 32
    ref = ofw.forward(spec, w, oenc.encode_windows(seqs, 500))
    assert np.abs(ref["prediction"] - y["prediction"]).max() <= 4e-3
    assert np.abs(ref["embedding"] - y["embedding"]).max() <= 1e-2
    assert "stack_resident_kernel" in eng.conv_kernel_names()        # narrow stacks run window-resident (csrc/conv_resident.cuh)
    eng.close()


@pytest.mark.parametrize("variant", ["baseline", "max_relu", "no_masking", "k3_dilated"])
def test_window_resident_kernel_vs_per_layer_kernels_and_oracle(variant, monkeypatch):
    """The window-resident kernel (one launch, every window kept in shared memory through all conv layers) against the
    per-layer kernels it replaces (JG_RESIDENT=0) and the fp32 oracle: the full config-3 graph (2 residual stacks), max pooling
    + ReLU, masking off, dilated taps; windows with N runs (masked codons), a short last window and a frame-length mix."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project
    from jaeger_b200.modelspec import baseline_500bp_config
    from oracle import encode as oenc
    from oracle import forward as ofw
    from tests.helpers import random_contigs
    cfg = baseline_500bp_config()
    rl = cfg["model"]["representation_learner"]
    if variant == "max_relu":
        rl["pooling"] = "max"
        cfg["model"]["activation"] = "relu"
        for layer in rl["hidden_layers"]:
            if layer["name"] == "activation":
                layer["config"]["activation"] = "relu"
    elif variant == "no_masking":
        cfg["model"]["use_masking"] = False
    elif variant == "k3_dilated":
        for layer in rl["hidden_layers"]:
            if layer["name"] == "residual_block":
                layer["config"].update(kernel_size=3, dilation_rate=3)
    spec = parse_project(cfg)
    w = init_random(spec, 5)
    recs = random_contigs(4, [500] * 200 + [499, 1700, 512, 640, 2100], n_run_every=5, lower_every=0)
    seqs = [s[i:i + 500] for _, s in recs for i in range(0, len(s) - 499, 500)]
    ref = ofw.forward(spec, w, oenc.encode_windows(seqs, 500))
    got = {}
    for resident in ("1", "0"):
        monkeypatch.setenv("JG_RESIDENT", resident)
        eng = B200Engine(spec=spec, weights=w)
        got[resident] = eng.predict(WindowSource(records=recs, fsize=500, stride=500))
        assert ("stack_resident_kernel" in eng.conv_kernel_names()) == (resident == "1")
        eng.close()
    for resident, y in got.items():
        assert y["prediction"].shape == (len(seqs), 3)
        assert np.abs(ref["prediction"] - y["prediction"]).max() <= 4e-3, resident
        assert np.abs(ref["embedding"] - y["embedding"]).max() <= 1e-2, resident
    # same arithmetic in both paths (fp32 accumulation order inside the MMA aside)
    assert np.abs(got["1"]["prediction"] - got["0"]["prediction"]).max() <= 2e-3
    # window independence: a window's logits do not depend on its neighbours in the batch
    monkeypatch.setenv("JG_RESIDENT", "1")
    eng = B200Engine(spec=spec, weights=w)
    some = recs[::7]
    y2 = eng.predict(WindowSource(records=some, fsize=500, stride=500))
    names = list(got["1"]["meta_0"])
    for i, nm in enumerate(y2["meta_0"]):
        j = names.index(nm) if nm in names else -1
        if j >= 0 and list(names).count(nm) == 1:
            assert np.array_equal(y2["prediction"][i], got["1"]["prediction"][j])
    eng.close()


def test_dustmask_kernel_vs_sdust_oracle(standin):
    """Symmetric DUST soft-masking on the device vs the oracle's restatement of SDUST, including
    chunk boundaries (contigs of several chunks), N-breaks and low-complexity inserts; then the
    counts / N% path: soft-masked bases are not counted (seqops/io.py:124-127)."""
    from jaeger_b200 import WindowSource
    from oracle import dust as odust
    from oracle import seqwin
    _, _, eng = standin
    rng = np.random.default_rng(12)
    recs = []
    for i, n in enumerate([5000, 2300, 1024, 1023, 1025, 9000, 70, 3, 4000]):
        s = rng.choice(list("ACGT"), n)
        for _ in range(max(1, n // 700)):
            a = int(rng.integers(0, max(1, n - 80)))
            kind = int(rng.integers(0, 3))
            rep = (["A", "AC", "GGT"][kind] * 80)
            ln = min(int(rng.integers(15, 70)), n - a)
            s[a:a + ln] = list(rep[:ln])
        if i == 5:
            s[1000:1010] = "N"
            s[1020:1060] = "T"          # poly-T right after an N break, straddling a chunk boundary
        recs.append((f"d{i}", "".join(s)))
    buf = "".join(s for _, s in recs).encode()
    lens = np.array([len(s) for _, s in recs], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)])
    with torch.cuda.stream(eng._stream()):
        dev = torch.from_numpy(np.frombuffer(buf, dtype=np.uint8).copy()).to(eng.tdev)
        codes, valid = eng.pack(dev)
        soft = eng.dust(codes, valid, off).cpu().numpy()
    eng.ctx.sync()
    bits = np.unpackbits(soft.view(np.uint8), bitorder="little")[:off[-1]].astype(bool)
    n_masked = 0
    for (name, s), a, b in zip(recs, off[:-1], off[1:]):
        want = odust.mask_bits(s)
        assert np.array_equal(bits[a:b], want), (name, np.flatnonzero(bits[a:b] != want)[:10])
        n_masked += int(want.sum())
    assert n_masked > 300
    # end to end: metadata with dustmask on == oracle windows over the soft-masked sequences
    y = eng.predict(WindowSource(records=recs, fsize=2000, stride=1500, dustmask=True))
    masks = {n: odust.mask_bits(s) for n, s in recs}
    wins = list(seqwin.fragment_windows(recs, 2000, 1500, softmasks=masks))
    assert len(wins) == len(y["meta_0"])
    for i, w in enumerate(wins):
        assert [int(y[f"meta_{k}"][i]) for k in (5, 6, 7, 8)] == [w.g, w.c, w.a, w.t], i
        assert y["meta_9"][i].decode() == w.gc_skew


def _refine_case(seed=11):
    from tests.helpers import refine_case
    return refine_case(seed)


@pytest.mark.parametrize("mode,split,allow", [("gated", "half", False), ("weighted", "full", True), ("unweighted", "half", True)])
def test_refinement_layer_vs_oracle(standin, mode, split, allow):
    """`--refine` (postprocess/refinement.py:39-247): per-window refined labels identical to the oracle's (ties between
    exactly equal logits included), per-contig aggregated scores / margins within 1e-9 relative (float64 sums in a
    different order), used / merged window counts, contig calls and the min_windows filter identical."""
    from jaeger_b200.refine import LABELS, refine_contigs, refined_contig_table
    from oracle import refine as orf
    _, _, eng = standin
    z, offsets, headers, taus = _refine_case()
    feat = orf.add_score_features(z)
    want_lab = orf.refine(feat, taus)
    r = refine_contigs(eng, z, offsets, taus, mode, split)
    assert [LABELS[i] for i in r["label"]] == want_lab.tolist()
    assert np.array_equal(r["margin"], feat["margin"])                   # float64 difference of widened float32 logits: exact
    assert {"unknown", "bacteria_or_plasmid", "virus_any"} <= set(want_lab.tolist())
    ids = np.repeat(np.array(headers, dtype=object), np.diff(offsets))
    want = orf.aggregate_contig(ids, z, want_lab, feat["margin"], mode=mode, min_windows=3, merge_split=split,
                                allow_merged_contig_call=allow, contig_hedge_margin=5.0)
    got = refined_contig_table(eng, headers, z, offsets, taus, mode=mode, min_windows=3, merge_split=split,
                               allow_merged_contig_call=allow, contig_hedge_margin=5.0)
    assert got["contig_id"].tolist() == list(want)
    for row in got.to_dict("records"):
        w = want[row["contig_id"]]
        for k, v in w.items():
            if isinstance(v, float):
                assert row[k] == pytest.approx(v, rel=1e-9, abs=1e-12), (row["contig_id"], k)
            else:
                assert row[k] == v, (row["contig_id"], k, row[k], v)
    if mode == "unweighted":
        assert set(got["contig_call"]) & {"bacteria_or_plasmid", "virus_any"}


def test_driver_refine_option_adds_the_refined_columns(standin, tmp_path):
    """run_core(refine=True): the summary TSV gains contig_call, contig_top_logit, contig_margin, n_windows_used,
    n_merged_windows (collect.py:534-550) for contigs with >= 3 informative windows; a calibration file for another
    model is rejected with a warning and the default table is written (predict.py:329-330)."""
    import pandas as pd
    import yaml
    from jaeger_b200.predict import run_core
    from oracle import refine as orf
    from tests.helpers import random_contigs
    recs = random_contigs(3, [2000, 9000, 30000, 5200, 2600])
    fa = tmp_path / "c.fasta"
    fa.write_text("".join(f">{n}\n{s}\n" for n, s in recs))
    taus = {c: {"logit": -0.5, "margin": 0.01, "n": 100} for c in orf.CLASSES}
    cal = tmp_path / "standin_refine.yaml"
    cal.write_text(yaml.safe_dump({"schema_version": 1, "jaeger_model": "standin", "quantile": 0.05, "taus": taus}, sort_keys=False))
    res = run_core(input=str(fa), output=str(tmp_path / "o"), model="standin", allow_random_weights=True, fsize=2000, stride=1500, overwrite=True, dustmask=False,
                   refine=True, refine_file=str(cal), window_scores=True, terminal_repeats=False, save_embedding=True, save_nmd=True,
                   mem=2, precision="fp16")
    emb, nmd = np.load(tmp_path / "o" / "standin" / "c_embedding.npz"), np.load(tmp_path / "o" / "standin" / "c_nmd.npz")
    assert emb["embedding"].shape == (res["windows"], 128) and nmd["embedding"].shape == (res["windows"], 640)     # predict.py:66-112
    assert len(emb["headers"]) == res["windows"] and emb["headers"][0] in (b"c0", "c0")
    with pytest.raises(RuntimeError):
        run_core(input=str(fa), output=str(tmp_path / "o3"), model="standin", allow_random_weights=True, cpu=True)
    tsv = pd.read_csv(res["table"], sep="\t")
    for col in ("contig_call", "contig_top_logit", "contig_margin", "n_windows_used", "n_merged_windows"):
        assert col in tsv.columns
    z = np.load(tmp_path / "o" / "standin" / "c_window_scores.npz", allow_pickle=True)
    want = orf.build_refined_contig([str(h) for h in z["headers"]], list(z["predictions"]), taus, mode="gated", min_windows=3)
    tsv = tsv.set_index("contig_id")
    assert len(want) >= 2
    for cid in tsv.index:
        if cid in want:
            assert tsv.loc[cid, "contig_call"] == want[cid]["contig_call"]
            assert int(tsv.loc[cid, "n_windows_used"]) == want[cid]["n_windows_used"]
            assert abs(float(tsv.loc[cid, "contig_margin"]) - want[cid]["contig_margin"]) < 2e-3
        else:
            assert pd.isna(tsv.loc[cid, "contig_call"])
    cal.write_text(yaml.safe_dump({"schema_version": 1, "jaeger_model": "other", "taus": taus}))
    res = run_core(input=str(fa), output=str(tmp_path / "o2"), model="standin", allow_random_weights=True, fsize=2000, stride=1500, overwrite=True, dustmask=False,
                   refine=True, refine_file=str(cal), terminal_repeats=False)
    assert "contig_call" not in pd.read_csv(res["table"], sep="\t").columns


@pytest.mark.parametrize("masking", [True, False])
def test_return_nmd_taps_through_batchnorm_vs_oracle(masking):
    """`return_nmd: true` on masked_batchnorm / residual_block (layers.py:943-954, 1897-1898; the way
    train_config/nn_config_baseline.yaml:205 feeds its reliability head): the NMD vectors are taken on the raw output of
    the stem conv, on the raw conv2 output of a residual stack's last block (a launch with a shortcut) and on a block
    output -- logits, NMD vectors and reliability against the fp32 oracle on both kernel families."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project
    from oracle import encode as oenc
    from oracle import forward as ofw
    from oracle import seqwin
    from tests.helpers import random_contigs
    from tests.test_plan_cpu import return_nmd_config
    spec = parse_project(return_nmd_config(masking))
    w = init_random(spec, 5)
    recs = random_contigs(17, [2000, 6500, 9000, 2300])
    wins = list(seqwin.fragment_windows(recs, 2000, 1500))
    ref = ofw.forward(spec, w, oenc.encode_windows([x.seq for x in wins], 2000))
    assert ref["nmd"].shape[1] == 384
    for use_ref in (False, True):
        eng = B200Engine(spec=spec, weights=w, use_ref_kernels=use_ref)
        y = eng.predict(WindowSource(records=recs, fsize=2000, stride=1500))
        eng.close()
        for k, tol in (("prediction", 4e-3), ("embedding", 1e-2), ("nmd", 4e-3), ("reliability", 4e-3)):
            assert np.abs(ref[k] - y[k]).max() <= tol, (k, use_ref, float(np.abs(ref[k] - y[k]).max()))


def test_drop_in_model_directory_config_json_and_savedmodel_bundle(tmp_path):
    """The drop-in boundary as a user meets it: a model directory laid out like the reference's
    (`<root>/model/<name>_graph/variables/`, `<name>_classes.yaml`, `<name>_project.yaml`, `<name>_refine.yaml`;
    utils/misc.py:346-392), found through config.json's `model_paths` (-m <name> --config) or through --model_path;
    weights come from the SavedModel bundle.  Output lands in <-o>/<model_id>/ and equals the run with the same
    weights handed over in memory."""
    import pandas as pd
    import yaml
    from jaeger_b200 import init_random, parse_project, standin_1p4m_config
    from jaeger_b200.predict import run_core
    from jaeger_b200 import B200Engine, WindowSource
    from jaeger_b200.postprocess import contig_table, generate_summary
    from oracle import refine as orf
    from tests.helpers import random_contigs
    from tests.tf_bundle_writer import keras3_export_names, write_bundle
    cfg = standin_1p4m_config()
    spec = parse_project(cfg)
    w = init_random(spec, 21)
    name = "jaeger_38341_1.4M_fragment"
    mdir = tmp_path / "models" / "nested" / "model"
    write_bundle(mdir / f"{name}_graph" / "variables", keras3_export_names(spec, w))
    (mdir / f"{name}_project.yaml").write_text(yaml.safe_dump(cfg))
    classes = [{"class": c["class"], "label": c["label"]} for c in cfg["model"]["class_label_map"]]
    (mdir / f"{name}_classes.yaml").write_text(yaml.safe_dump({"classes": classes}))
    taus = {c: {"logit": -1.0, "margin": 0.0, "n": 100} for c in orf.CLASSES}
    (mdir / f"{name}_refine.yaml").write_text(yaml.safe_dump({"schema_version": 1, "jaeger_model": name, "taus": taus}))
    (tmp_path / "config.json").write_text(json.dumps({"model_paths": [str(tmp_path / "models")]}))
    recs = random_contigs(5, [2000, 7000, 12000, 3100])
    fa = tmp_path / "in.fasta"
    fa.write_text("".join(f">{n}\n{s}\n" for n, s in recs))
    r1 = run_core(input=str(fa), output=str(tmp_path / "o1"), model=name, config=str(tmp_path / "config.json"), refine=True, dustmask=False)
    r2 = run_core(input=str(fa), output=str(tmp_path / "o2"), model_path=str(tmp_path / "models"), dustmask=False)
    assert r1["table"] == tmp_path / "o1" / "38341_1.4M" / "in.tsv" and r2["table"] == tmp_path / "o2" / "38341_1.4M" / "in.tsv"
    t1, t2 = pd.read_csv(r1["table"], sep="\t"), pd.read_csv(r2["table"], sep="\t")
    assert "contig_call" in t1.columns and "contig_call" not in t2.columns
    assert t1[t2.columns].equals(t2)
    eng = B200Engine(spec=spec, weights=w)
    y = eng.predict(WindowSource(fasta=fa, fsize=2000, stride=1500, dustmask=False))
    want = generate_summary(contig_table(eng, y, 2000), eng.class_map["class"], eng.class_map["index"])
    eng.close()
    for col in ("contig_id", "prediction", "length", "window_summary"):
        assert t2[col].tolist() == want[col].tolist(), col
    assert np.allclose(t2["phage_score"].to_numpy(), want["phage_score"].to_numpy(dtype=np.float64), atol=6e-4)
    with pytest.raises(ValueError, match="not found"):
        run_core(input=str(fa), output=str(tmp_path / "o3"), model="jaeger_0_none_fragment", config=str(tmp_path / "config.json"))


def test_reliability_head_with_ood_signals_vs_oracle():
    """reliability_model.mode nmd_plus_signals (builder.py:644-657, 716-722; OODSignalLayer layers.py:1598-1683): the
    reliability head reads concat(NMD vector, max_prob, entropy, energy, margin, nmd_norm) computed from the window's
    own logits -- fp32 on the device, against the fp32 oracle (tolerance 4e-3 like the other head outputs)."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project
    from oracle import encode as oenc
    from oracle import forward as ofw
    from oracle import seqwin
    from tests.helpers import random_contigs
    from tests.test_plan_cpu import signals_config, small_config
    recs = random_contigs(23, [2000, 6500, 9000, 2300])
    wins = list(seqwin.fragment_windows(recs, 2000, 1500))
    tok = oenc.encode_windows([x.seq for x in wins], 2000)
    for signals in (None, ["margin", "nmd_norm", "energy"]):
        spec = parse_project(signals_config(signals))
        w = init_random(spec, 9)
        w["classifier"][0]["kernel"] *= 8.0
        w["reliability"][0]["kernel"][384:] *= 3.0            # make the signal rows count
        ref = ofw.forward(spec, w, tok)
        plain = ofw.forward(parse_project(small_config()),
                            dict(w, reliability=[dict(w["reliability"][0], kernel=w["reliability"][0]["kernel"][:384]), w["reliability"][1]]), tok)
        assert np.abs(ref["reliability"] - plain["reliability"]).max() > 0.05        # the signals change the output
        eng = B200Engine(spec=spec, weights=w)
        y = eng.predict(WindowSource(records=recs, fsize=2000, stride=1500))
        eng.close()
        for k, tol in (("prediction", 4e-3), ("nmd", 4e-3), ("reliability", 4e-3)):
            assert np.abs(ref[k] - y[k]).max() <= tol, (k, signals, float(np.abs(ref[k] - y[k]).max()))


def test_500bp_nmd_merge_model_vs_oracle():
    """train_config/nn_config_500bp_nmd_merge.yaml (restated in tests/test_plan_cpu.py): NMD taps on 32-channel layers
    (zero-padded to 64 on the device, sliced back on return) and a tap on a launch's final output, average pooling,
    reliability head -- 500-bp windows against the fp32 oracle, tensor-core and CUDA-core kernels."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project
    from oracle import encode as oenc
    from oracle import forward as ofw
    from oracle import seqwin
    from tests.helpers import random_contigs
    from tests.test_plan_cpu import nmd_merge_500bp_config
    spec = parse_project(nmd_merge_500bp_config())
    w = init_random(spec, 2)
    recs = random_contigs(41, [500, 1700, 2600, 900, 500])
    wins = list(seqwin.fragment_windows(recs, 500, 500))
    ref = ofw.forward(spec, w, oenc.encode_windows([x.seq for x in wins], 500))
    assert ref["nmd"].shape == (len(wins), 64)
    for use_ref in (False, True):
        eng = B200Engine(spec=spec, weights=w, use_ref_kernels=use_ref)
        y = eng.predict(WindowSource(records=recs, fsize=500, stride=500))
        if not use_ref:                                   # narrow stack: the window-resident kernel, NMD taps included
            assert "stack_resident_kernel" in eng.conv_kernel_names()
        eng.close()
        assert y["nmd"].shape == ref["nmd"].shape and y["embedding"].shape == ref["embedding"].shape
        for k, tol in (("prediction", 4e-3), ("embedding", 1e-2), ("nmd", 4e-3), ("reliability", 4e-3)):
            assert np.abs(ref[k] - y[k]).max() <= tol, (k, use_ref, float(np.abs(ref[k] - y[k]).max()))


@pytest.mark.parametrize("tap", ["raw_stem", "after_act", "return_nmd"])
def test_window_resident_kernel_nmd_taps_vs_per_layer_kernels(tap, monkeypatch):
    """NMD taps inside the window-resident kernel -- on the raw stem output (taken from token counts by stem_tap_kernel on both
    paths), after an activation, on the launch output, and `masked_batchnorm(return_nmd=True)` in front of a norm -- against the
    per-layer kernels (JG_RESIDENT=0) and the fp32 oracle, with N runs in the windows (masked sums and counts)."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project
    from oracle import encode as oenc
    from oracle import forward as ofw
    from tests.helpers import random_contigs
    from tests.test_plan_cpu import nmd_merge_500bp_config
    import copy
    cfg = copy.deepcopy(nmd_merge_500bp_config())        # the two tails of the config share their layer dicts
    hl = cfg["model"]["representation_learner"]["hidden_layers"]
    if tap == "raw_stem":            # conv -> nmd -> norm -> act: the tap reads the raw conv output
        hl.insert(1, {"name": "nmd", "config": {}})
        del hl[4]
    elif tap == "return_nmd":        # the stem's BatchNorm returns the NMD of its input
        hl[1] = {"name": "masked_batchnorm", "config": {"return_nmd": True}}
        del hl[3]
    spec = parse_project(cfg)
    w = init_random(spec, 3)
    recs = random_contigs(43, [500] * 90 + [1700, 2600, 900], n_run_every=4, lower_every=0)
    seqs = [s[i:i + 500] for _, s in recs for i in range(0, len(s) - 499, 500)]
    ref = ofw.forward(spec, w, oenc.encode_windows(seqs, 500))
    got = {}
    for resident in ("1", "0"):
        monkeypatch.setenv("JG_RESIDENT", resident)
        eng = B200Engine(spec=spec, weights=w)
        got[resident] = eng.predict(WindowSource(records=recs, fsize=500, stride=500))
        assert ("stack_resident_kernel" in eng.conv_kernel_names()) == (resident == "1")
        eng.close()
        for k, tol in (("prediction", 4e-3), ("embedding", 1e-2), ("nmd", 4e-3), ("reliability", 4e-3)):
            assert np.abs(ref[k] - got[resident][k]).max() <= tol, (k, resident, float(np.abs(ref[k] - got[resident][k]).max()))
    for k in ("prediction", "nmd", "reliability"):
        assert np.abs(got["1"][k] - got["0"][k]).max() <= 2e-3, k


def test_reference_dataset_protocol_batches_and_evaluate(standin):
    """The engine behind the reference's own data pipeline: `predict` fed with the
    `(inputs_dict, meta_0..meta_9)` batches `process_string_inference` yields (float32 tokens [B, 6, Lc], or one-hot
    [B, 6, Lc, 64]) returns the same logits as the engine's own FASTA path, meta columns passed through in window order
    (inference.py:341-373); `evaluate` on `(inputs_dict, y_true_onehot)` batches gives Keras' from-logits categorical
    cross-entropy and the accuracy (inference.py:375-408)."""
    from jaeger_b200 import WindowSource
    from oracle import encode as oenc
    from oracle import seqwin
    from tests.helpers import random_contigs
    _, _, eng = standin
    recs = random_contigs(12, [2000, 5200, 8000, 3500])
    y_src = eng.predict(WindowSource(records=recs, fsize=2000, stride=1500))
    wins = list(seqwin.fragment_windows(recs, 2000, 1500))
    tok = oenc.encode_windows([x.seq for x in wins], 2000)
    n = len(wins)

    def batches(onehot):
        for a in range(0, n, 4):
            t = tok[a:a + 4]
            x = (np.eye(65, dtype=np.float32)[t][..., 1:] if onehot else t.astype(np.float32))
            meta = [np.array([f"{k}_{i}".encode() for i in range(a, min(a + 4, n))]) for k in range(10)]
            yield ({"translated": x}, *meta)

    for onehot in (False, True):
        y = eng.predict(batches(onehot), no_progress=True)
        assert y["prediction"].shape == (n, 6) and np.array_equal(y["prediction"], y_src["prediction"])
        assert np.abs(y["reliability"] - y_src["reliability"]).max() <= 1e-6       # NMD sums are fp32 atomics: order-dependent last bit
        assert [m.decode() for m in y["meta_3"]] == [f"3_{i}" for i in range(n)] and set(k for k in y if k.startswith("meta_")) == {f"meta_{i}" for i in range(10)}
    labels = np.random.default_rng(0).integers(0, 6, n)
    onehot_y = np.eye(6, dtype=np.float32)[labels]
    res = eng.evaluate((({"translated": tok[a:a + 5].astype(np.float32)}, onehot_y[a:a + 5]) for a in range(0, n, 5)))
    z = y_src["prediction"].astype(np.float64)
    lse = np.log(np.exp(z - z.max(1, keepdims=True)).sum(1)) + z.max(1)
    assert abs(res["loss"] - float(np.mean(lse - z[np.arange(n), labels]))) < 1e-5
    assert res["accuracy"] == float(np.mean(z.argmax(1) == labels))


@pytest.mark.gpu
def test_label_agreement_at_scale_real_weights_and_scaled_standin():
    """North-star target: >= 99.9 % per-contig label agreement with the reference's path on the same inputs
    (contig label = argmax of the float16 mean logits, postprocess/collect.py:332-342).
    (A) the reference's REAL `default` weights on 5 000 synthetic contigs (about 12 000 windows): device vs oracle/legacy.py.
    (B) the stand-in architecture with the classifier kernel scaled x25 (|logit| ~ 5-10, so the fp16-activation error is
        no longer small against the logits) on 700 contigs: device vs oracle/forward.py (fp32).
    Also bounds the logit error against the top-2 margin: a window may only flip where the oracle's own margin is within
    twice the observed error.  tools/label_agreement.py runs the same at larger sizes and writes profiles/label_agreement_r2.json."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project, standin_1p4m_config
    from jaeger_b200 import codon_tables as ct
    from oracle import encode as oenc
    from oracle import forward as ofw
    from oracle import legacy as oleg
    from oracle import seqwin
    from tests.helpers import agreement_contigs, agreement_stats
    w, _ = _legacy_fixture()
    recs = agreement_contigs(101, 5000)
    eng = B200Engine(legacy_weights=w)
    y = eng.predict(WindowSource(records=recs, fsize=2000, stride=1500, outputs=("prediction",), lazy_meta=True))
    eng.close()
    wins = list(seqwin.fragment_windows(recs, 2000, 1500))
    table = dict(zip(oenc.CODONS, ct.LEGACY_AA_ID))
    tok = np.stack([oenc.encode_window_legacy(x.seq, 2000, table) for x in wins]).astype(np.uint8)
    ref = np.concatenate([oleg.forward(w, tok[b:b + 256])["output"] for b in range(0, len(tok), 256)])
    st = agreement_stats(y["prediction"], ref, np.array([x.is_last for x in wins]))
    print("legacy real weights:", st)
    assert st["contigs"] == 5000 and st["windows"] >= 10000
    assert st["contig_label_agreement"] >= 0.999, st
    assert st["window_label_agreement"] >= 0.995, st
    assert st["max_abs_logit_err"] <= 0.03, st                      # fp16-stored activations, real weights, |logit| up to ~10 (observed 0.013)
    assert st["largest_margin_among_flips"] <= 2 * st["max_abs_logit_err"], st   # only near-ties flip

    spec = parse_project(standin_1p4m_config())
    wt = init_random(spec, 0)
    wt["classifier"][0]["kernel"] = wt["classifier"][0]["kernel"] * 25.0
    recs = agreement_contigs(202, 700)
    eng = B200Engine(spec=spec, weights=wt)
    y = eng.predict(WindowSource(records=recs, fsize=2000, stride=1500, outputs=("prediction",), lazy_meta=True))
    eng.close()
    wins = list(seqwin.fragment_windows(recs, 2000, 1500))
    tok = oenc.encode_windows([x.seq for x in wins], 2000)
    ref = np.concatenate([ofw.forward(spec, wt, tok[b:b + 96])["prediction"] for b in range(0, len(tok), 96)])
    st = agreement_stats(y["prediction"], ref, np.array([x.is_last for x in wins]))
    print("stand-in, classifier x25:", st)
    assert st["mean_abs_logit"] >= 1.0, st                           # the logits are no longer tiny
    assert st["contig_label_agreement"] >= 0.999, st
    assert st["max_abs_logit_err"] <= 0.1, st                        # 25 x the 4e-3 bound of the unscaled stand-in
    assert st["largest_margin_among_flips"] <= 2 * st["max_abs_logit_err"], st


@pytest.mark.parametrize("which,lc", [("small", 665), ("small", 498), ("small_wide", 665), ("baseline_3p4m", 665), ("flat_g1", 341)])
def test_strided_bypass_blocks_and_mlp_heads_vs_oracle(which, lc):
    """Residual blocks with strides = 2 / a 1x1 bypass conv + norm (nnlib/v2/layers.py:1840-1864, 1903-1909), layers of 256
    channels (run as slices of 64 output channels on the CTA-pair kernel) and multi-layer classification heads
    (builder.py:589-596) on the device against the fp32 oracle: a small network with every variant (odd and even frame
    lengths), the reference's train_config/nn_config_baseline.yaml (3.4 M) and the flat first-generation schema of
    commands/configs/nn_config.yaml.  Tolerance: fp16 activations / weights vs fp32, |logit| <~ 0.5: 6e-3 absolute."""
    import yaml
    from jaeger_b200 import B200Engine, init_random, parse_project
    from oracle import forward as ofw
    from tests.test_plan_cpu import _baseline_3p4m_config, flat_g1_config, small_strided_config
    if which == "small":
        cfg = small_strided_config()
    elif which == "small_wide":                               # 128 -> 256 channels: the wide layers run as slices of 64 output channels
        cfg = small_strided_config(filters=(128, 256))
    elif which == "baseline_3p4m":
        if not Path("/root/reference/train_config/nn_config_baseline.yaml").exists():
            pytest.skip("the reference tree is not on this box: the configuration is exercised by the CPU suite")
        cfg = _baseline_3p4m_config()
    else:
        cfg = flat_g1_config()
    spec = parse_project(cfg)
    w = init_random(spec, 7)
    rng = np.random.default_rng(lc)
    tok = rng.integers(0, 65, (5, 6, lc)).astype(np.uint8)
    tok[1, :, lc - 40:] = 0                                   # unknown codons are ordinary zero rows when masking is off
    eng = B200Engine(spec=spec, weights=w)
    got = eng.predict([({"translated": tok},)])
    ref = ofw.forward(spec, w, tok)
    names = eng.conv_kernel_names()
    eng.close()
    for key in ("prediction", "embedding") + (("nmd",) if "nmd" in ref else ()):
        d = np.abs(got[key] - ref[key]).max()
        assert d < 6e-3 * max(1.0, np.abs(ref[key]).max()), (which, key, d, names)
    assert "conv_ref_kernel" not in names, names             # every conv runs on a tensor-core kernel
    print(which, lc, names, "max |logit diff|", np.abs(got["prediction"] - ref["prediction"]).max())


def test_streamed_run_equals_whole_file_run(standin, tmp_path):
    """f3: `run_core` with the input streamed in chunks of whole records (two pinned buffers, the next chunk parsed while the
    current one is on the device) writes byte-identical tables to the whole-file run -- long pass and padded short pass,
    dust-masking, terminal repeats -- and a second streaming pass writes the same phage FASTA."""
    from jaeger_b200.predict import run_core
    from tests.helpers import random_contigs
    recs = random_contigs(11, [2600, 900, 14000, 2000, 700, 31000, 5200, 1200, 8000, 2100, 600, 45000, 3000])
    fa = tmp_path / "meta.fasta"
    fa.write_text("".join(f">{n} len={len(s)}\n" + "\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + "\n" for n, s in recs))
    common = dict(input=str(fa), model="standin", allow_random_weights=True, fsize=2000, stride=1500, min_len=500, overwrite=True,
                  getsequences=True, pc=-100.0, rc=-1.0)
    a = run_core(output=str(tmp_path / "whole"), stream_mbp=0, **common)
    b = run_core(output=str(tmp_path / "streamed"), stream_mbp=0.02, **common)
    assert b["streamed_chunks"] >= 4 and a["windows"] == b["windows"] and a["num_written"] == b["num_written"]
    assert a["table"].read_text() == b["table"].read_text()
    assert a["phage_table"].exists() == b["phage_table"].exists()      # only written when a contig is called phage (collect.py:601-607)
    if a["phage_table"].exists():
        assert a["phage_table"].read_text() == b["phage_table"].read_text()
    fa_a, fa_b = (tmp_path / d / "standin" / "meta_phages_jaeger.fasta" for d in ("whole", "streamed"))
    assert fa_a.read_bytes() == fa_b.read_bytes()


@pytest.mark.parametrize("standalone_ln,masking,bypass", [(True, True, False), (False, True, True), (True, False, False)])
def test_masked_layernorm_vs_oracle(standalone_ln, masking, bypass):
    """MaskedLayerNormalization (nnlib/v2/layers.py:293-367) as the norm of the residual blocks and as a stand-alone layer after the
    stem: row statistics over the real channels in the single-CTA tensor-core kernel's epilogue (and in the CUDA-core kernel), against
    the fp32 oracle whose LayerNorm is pinned on the reference's own layer code (tests/golden/v2_layers.npz)."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project
    from oracle import encode as oenc
    from oracle import forward as ofw
    from tests.helpers import random_contigs
    from tests.test_plan_cpu import layernorm_config
    for filters in (24, 128):
        spec = parse_project(layernorm_config(standalone_ln, masking, filters=filters, bypass=bypass))
        w = init_random(spec, 4)
        recs = random_contigs(6, [300] * 40 + [299, 900, 310], n_run_every=3, lower_every=0)
        seqs = [s[i:i + 300] for _, s in recs for i in range(0, len(s) - 299, 300)]
        ref = ofw.forward(spec, w, oenc.encode_windows(seqs, 300))
        for use_ref in (False, True):
            eng = B200Engine(spec=spec, weights=w, use_ref_kernels=use_ref)
            y = eng.predict(WindowSource(records=recs, fsize=300, stride=300))
            # random-init LayerNorm models have logits of magnitude 10-20: fp16 activations -> a tolerance relative to the scale
            # (the stand-in with a scaled classifier shows the same 1e-3 of the logit scale, profiles/label_agreement_r2.json)
            for k, tol in (("prediction", 3e-3), ("embedding", 4e-3), ("nmd", 4e-3), ("reliability", 4e-3)):
                bound = tol * max(4.0, float(np.abs(ref[k]).max()))
                assert np.abs(ref[k] - y[k]).max() <= bound, (k, filters, use_ref, float(np.abs(ref[k] - y[k]).max()), bound)
            if not use_ref:
                assert "conv_tc_kernel" in eng.conv_kernel_names()
            eng.close()


@pytest.mark.parametrize("modes", [("strict", "strict", "strict"), ("majority", "majority", "majority"), ("strict", "majority", "any")])
def test_mask_modes_strict_and_majority_vs_oracle(modes):
    """mask_mode strict / majority on stand-alone MaskedConv1D layers (nnlib/v2/layers.py:1245-1252): the validity helper of the
    tensor-core kernels counts the valid taps of every output row and thresholds them (and so does the CUDA-core path); windows
    with N runs, isolated Ns and a masked tail, against the fp32 oracle."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project
    from oracle import encode as oenc
    from oracle import forward as ofw
    from tests.helpers import random_contigs
    from tests.test_plan_cpu import conv_stack_config
    spec = parse_project(conv_stack_config(modes))
    w = init_random(spec, 6)
    recs = random_contigs(9, [300] * 60 + [299, 900, 310], n_run_every=2, lower_every=0)
    seqs = [s[i:i + 300] for _, s in recs for i in range(0, len(s) - 299, 300)]
    ref = ofw.forward(spec, w, oenc.encode_windows(seqs, 300))
    for use_ref in (False, True):
        eng = B200Engine(spec=spec, weights=w, use_ref_kernels=use_ref)
        y = eng.predict(WindowSource(records=recs, fsize=300, stride=300))
        for k, tol in (("prediction", 4e-3), ("embedding", 1e-2), ("nmd", 1e-2), ("reliability", 1e-2)):
            assert np.abs(ref[k] - y[k]).max() <= tol, (k, use_ref, float(np.abs(ref[k] - y[k]).max()))
        eng.close()


def test_window_resident_kernel_other_geometries(monkeypatch):
    """The window-resident kernel away from the 500 bp / 8-tile geometry: 300 bp windows (6 tiles: two epilogue groups own one tile
    only), the padded short-contig pass (--min-len < --fsize: per-window frame lengths below the buffer's, batches padded to their
    longest member) and a single-window call (one CTA), against the per-layer kernels and the oracle."""
    from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project
    from jaeger_b200.modelspec import baseline_500bp_config
    from oracle import encode as oenc
    from oracle import forward as ofw
    from tests.helpers import random_contigs
    cfg = baseline_500bp_config()
    cfg["model"]["string_processor"]["crop_size"] = 300
    spec = parse_project(cfg)
    w = init_random(spec, 7)
    recs = random_contigs(12, [300] * 50 + [150, 180, 299, 240, 200, 900, 1000], n_run_every=4, lower_every=0)
    got = {}
    for resident in ("1", "0"):
        monkeypatch.setenv("JG_RESIDENT", resident)
        eng = B200Engine(spec=spec, weights=w)
        got[resident] = eng.predict(WindowSource(records=recs, fsize=300, stride=300, min_len=150, batch=8))
        assert ("stack_resident_kernel" in eng.conv_kernel_names()) == (resident == "1")
        one = eng.predict(WindowSource(records=recs[:1], fsize=300, stride=300))
        assert one["prediction"].shape == (1, 3)
        got[resident + "_one"] = one
        eng.close()
    assert got["1"]["prediction"].shape == got["0"]["prediction"].shape and got["1"]["prediction"].shape[0] > 55
    assert np.abs(got["1"]["prediction"] - got["0"]["prediction"]).max() <= 2e-3
    assert np.abs(got["1_one"]["prediction"] - got["0_one"]["prediction"]).max() <= 2e-3
    # the long pass against the oracle (full 300-bp windows, in the engine's window order: long pass first)
    seqs = [s[i:i + 300] for _, s in recs if len(s) >= 300 for i in range(0, len(s) - 299, 300)]
    ref = ofw.forward(spec, w, oenc.encode_windows(seqs, 300))
    assert np.abs(ref["prediction"] - got["1"]["prediction"][:len(seqs)]).max() <= 4e-3
