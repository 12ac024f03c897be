"""Config 1 pinned on the reference's own TensorFlow graph: tests/golden/legacy_graph_outputs.npz holds what the
reference's serialized `serving_default` function (data/models/test/jaeger_fragment_graph/saved_model.pb, the legacy
`default` model) returns for the 135 health-FASTA windows and for random tokens, computed by interpreting the GraphDef op
by op (oracle/tfgraph.py, generator tests/golden/make_legacy_graph_goldens.py).  Here the NumPy / torch restatement
oracle/legacy.py -- the checker of the CUDA path -- is held against those outputs."""
from __future__ import annotations

import zlib
from pathlib import Path

import numpy as np
import pytest
import torch

G = Path(__file__).resolve().parent / "golden"


def _weights():
    from jaeger_b200.weights import load_npz_weights
    import tempfile
    z = np.load(G / "legacy_default.npz")
    flat = {k[2:]: z[k] for k in z.files if k.startswith("w/")}
    with tempfile.NamedTemporaryFile(suffix=".npz") as fh:
        np.savez(fh.name, **flat)
        return load_npz_weights(fh.name)


def _health_tokens():
    from jaeger_b200 import codon_tables as ct
    from oracle import encode as oenc
    from oracle import seqwin
    z = np.load(G / "legacy_default.npz")
    recs = [(str(n), str(s)) for n, s in zip(z["names"], z["seqs"])]
    wins = list(seqwin.fragment_windows(recs, 2000, 1500))
    table = dict(zip(oenc.CODONS, ct.LEGACY_AA_ID))
    return np.stack([oenc.encode_window_legacy(x.seq, 2000, table) for x in wins]).astype(np.uint8)


def _random_tokens():
    rng = np.random.default_rng(0)
    tok = rng.integers(1, 22, size=(6, 6, 665)).astype(np.uint8)
    tok[0, :, 100:180] = 0
    tok[1, :, 300:] = 0
    tok[2][rng.random((6, 665)) < 0.1] = 0
    tok[3] = 0
    return tok


@pytest.mark.parametrize("tag", ["health", "random"])
def test_legacy_oracle_equals_the_reference_graph(tag):
    """float64 restatement within 1e-5 of the interpreted TF graph (observed 3e-6: float32 constants inside the graph),
    the float32 one within 1e-4, window labels identical."""
    from oracle import legacy as oleg
    gold = np.load(G / "legacy_graph_outputs.npz")
    tok = _health_tokens() if tag == "health" else _random_tokens()
    assert zlib.crc32(tok.tobytes()) == int(gold[f"{tag}_token_crc"])
    w = _weights()
    r64 = oleg.forward(w, tok, dtype=torch.float64)
    r32 = oleg.forward(w, tok)
    assert np.abs(r64["output"] - gold[f"{tag}_output"]).max() <= 1e-5
    assert np.abs(r64["embedding"] - gold[f"{tag}_embedding"]).max() <= 1e-5
    assert np.abs(r32["output"] - gold[f"{tag}_output"]).max() <= 1e-4
    assert np.array_equal(r32["output"].argmax(1), gold[f"{tag}_output"].argmax(1))
    if tag == "health":
        assert tok.shape[0] == 135


def test_tf_op_restatements_known_answers():
    """The interpreter's non-trivial ops on hand-checkable inputs: SpaceToBatchND / BatchToSpaceND round trip and layout
    (the tf.space_to_batch_nd documentation example), SAME padding split, StridedSlice masks."""
    from oracle.tfgraph import SavedFunction as SF
    x = np.arange(1, 17, dtype=np.float64).reshape(1, 4, 4, 1)
    y = SF._space_to_batch(x, [2, 2], [[0, 0], [0, 0]])
    assert y.shape == (4, 2, 2, 1)
    assert y[..., 0].tolist() == [[[1, 3], [9, 11]], [[2, 4], [10, 12]], [[5, 7], [13, 15]], [[6, 8], [14, 16]]]
    assert np.array_equal(SF._batch_to_space(y, [2, 2], [[0, 0], [0, 0]]), x)
    x1 = np.arange(10, dtype=np.float64).reshape(1, 10, 1)
    y1 = SF._space_to_batch(x1, [3], [[3, 2]])                      # dilation 3: pad to 15, three phases of 5
    assert y1.shape == (3, 5, 1) and y1[:, :, 0].tolist() == [[0, 0, 3, 6, 9], [0, 1, 4, 7, 0], [0, 2, 5, 8, 0]]
    assert np.array_equal(SF._batch_to_space(y1, [3], [[3, 2]]), x1)
    assert SF._same_pad(665, 9, 1) == (4, 4) and SF._same_pad(10, 4, 1) == (1, 2) and SF._same_pad(9, 5, 2) == (2, 2)
    d = np.arange(24).reshape(2, 3, 4)
    at = {"begin_mask": 1, "end_mask": 1, "shrink_axis_mask": 2}
    assert np.array_equal(SF._strided_slice([d, [0, 1], [0, 2], [1, 1]], at), d[:, 1])
    assert SF._strided_slice([np.array([7, 8, 9]), [1], [2], [1]], {"shrink_axis_mask": 1}) == 8


def test_v2_oracle_conv_shares_the_graph_pinned_padding_rule():
    """The TF SAME padding of dilated convolutions (k 5 / 9, dilations 1..7: total = d(k-1), left = total // 2) is what
    the interpreted graph computes through SpaceToBatchND; oracle/legacy._conv_same reproduces it (pinned above) and the
    v2 oracle's separate implementation (oracle/forward._conv1d_tf, used for every MaskedConv1D) equals it."""
    from oracle import forward as ofw
    from oracle import legacy as oleg
    g = torch.Generator().manual_seed(0)
    for k, d, length in [(5, 3, 40), (5, 1, 17), (9, 2, 33), (4, 1, 12), (4, 3, 25), (7, 7, 60)]:
        x = torch.randn(3, length, 6, generator=g, dtype=torch.float64)
        w = torch.randn(k, 6, 5, generator=g, dtype=torch.float64)
        a = ofw._conv1d_tf(x, w, d, "same")
        b = oleg._conv_same(x, w, None, d)
        assert a.shape == (3, length, 5) and torch.allclose(a, b, atol=1e-12)


@pytest.mark.parametrize("tag", ["default", "all_labels"])
def test_config1_tables_oracle_pipeline_vs_reference_end_to_end(tag):
    """BASELINE config 1 end to end: tests/golden/config1_*_jaeger.tsv is what the reference's own fragment_generator +
    serialized TF graph (interpreted) + pred_to_dict_legacy + generate_summary_legacy + bundled reliability model write for
    the health FASTA (tests/golden/make_config1_goldens.py).  The oracle pipeline the GPU driver test is held against
    (oracle windows -> legacy encoder -> oracle/legacy.forward -> oracle legacy tables) reproduces it: identical contig ids,
    lengths, labels, second labels, window counts and run-length summaries; scores / entropy / reliability to 1.5e-3."""
    import io
    import pandas as pd
    from jaeger_b200 import codon_tables as ct
    from oracle import encode as oenc
    from oracle import legacy as oleg
    from oracle import seqwin
    z = np.load(G / "legacy_default.npz")
    recs = [(str(n), str(s)) for n, s in zip(z["names"], z["seqs"])]
    wins = list(seqwin.fragment_windows(recs, 2000, 1500))
    ref = oleg.forward(_weights(), _health_tokens())
    zp = np.load(G / "legacy_post.npz")
    ood = {k[4:]: zp[k] for k in zp.files if k.startswith("ood_") and k != "ood_windows"}
    meta = [np.array([str(v).encode() for v in col]) for col in zip(*[
        (x.header, x.index, int(x.is_last), x.ordinal, x.seqlen, x.g, x.c, x.a, x.t, x.gc_skew) for x in wins])]
    all_labels = {0: "bacteria", 1: "phage", 2: "eukarya", 3: "archaea"}
    labels = ["non-phage", "phage", "non-phage", "non-phage"] if tag == "default" else [all_labels[i] for i in range(4)]
    cols, _ = oleg.summary_legacy(ref["output"], ref["embedding"], tuple(meta), 2000, ood, labels, all_labels,
                                  {1: "eukarya", 2: "archaea", 3: "bacteria", 0: ""}, 1)
    got = pd.DataFrame(cols)
    want = pd.read_csv(G / f"config1_{tag}_jaeger.tsv", sep="\t", keep_default_na=False)
    assert len(want) == 9 and want["prediction"].tolist() == ["phage"] * 9        # the health contigs are phages
    shared = [c for c in want.columns if c in got.columns]
    assert {"contig_id", "length", "prediction", "entropy", "reliability_score", "phage_score", "phage_var", "#_phage_windows",
            "window_summary", "G+C", "N%"} <= set(shared)
    for c in shared:
        w = want[c]
        if c in ("terminal_repeats", "repeat_length"):
            continue
        try:
            wf = w.to_numpy(dtype=np.float64)
        except (ValueError, TypeError):
            assert [str(v) for v in got[c]] == [str(v) for v in w], c
            continue
        assert np.allclose(got[c].to_numpy(dtype=np.float64), wf, atol=1.5e-3), (c, got[c].tolist(), w.tolist())
