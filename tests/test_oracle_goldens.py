"""CPU tests: the oracle against the golden vectors produced by the reference's own modules
(tests/golden/make_goldens.py) and against the reference tests' known answers."""
import json
import zlib
from pathlib import Path

import numpy as np
import pytest

from oracle import encode as oenc
from oracle import postprocess as opp
from oracle import seqwin

G = Path(__file__).resolve().parent / "golden"
REF_FASTA = Path("/root/reference/src/jaeger/data/test/test_contigs.fasta")


def test_window_indices_match_reference():
    cases = json.loads((G / "window_indices.json").read_text())
    assert len(cases) > 300
    for c in cases:
        assert seqwin.window_indices(c["seqlen"], c["fsize"], c["stride"], c["dyn"], c["thr"]) == c["idx"], c


def test_window_indices_reference_known_answers():
    # tests/unit/test_seqops_io.py of the reference
    assert seqwin.window_indices(3400, 2000, 2000, False, 10.0) == [0]
    assert seqwin.window_indices(3400, 2000, 2000, True, 10.0) == [0, 1400]
    assert seqwin.window_indices(3999, 2000, 2000, True, 10.0) == [0, 1999]
    assert seqwin.window_indices(6000, 2000, 2000, True, 10.0) == [0, 2000, 4000]


def _rows(path, fs, st, min_len, max_len=None, dyn=False):
    rows = []
    for w in seqwin.fragment_windows(seqwin.read_fasta(str(path)), fs, st, min_len=min_len, max_len=max_len,
                                     dynamic_stride=dyn):
        rows.append([zlib.crc32(w.seq.encode()), len(w.seq), w.header, str(w.index), str(w.is_last), str(w.ordinal),
                     str(w.seqlen), str(w.g), str(w.c), str(w.a), str(w.t), w.gc_skew])
    return rows


def test_fragment_windows_synthetic_fasta():
    gold = json.loads((G / "fragments_synthetic.json").read_text())
    for key, rows in gold.items():
        fs, st, mn, mx, dyn = key.split("_")
        got = _rows(G / "synthetic_contigs.fasta", int(fs), int(st), None if mn == "None" else int(mn),
                    None if mx == "None" else int(mx), bool(int(dyn)))
        assert got == rows, key


@pytest.mark.skipif(not REF_FASTA.exists(), reason="reference checkout not mounted")
def test_fragment_windows_reference_health_fasta():
    gold = json.loads((G / "fragments_test_contigs.json").read_text())
    for key, rows in gold.items():
        fs, st, mn = key.split("_")
        got = _rows(REF_FASTA, int(fs), int(st), None if mn == "None" else int(mn))
        assert got == rows, key
    # 135 windows at the CLI defaults, 100 at the `health` geometry
    assert len(gold["2000_1500_None"]) == 135 and len(gold["2048_2048_None"]) == 100


def test_validate_fasta_entries(tmp_path):
    assert seqwin.validate_fasta_entries(str(G / "synthetic_contigs.fasta"), 2000) == 12
    short = tmp_path / "short.fasta"
    short.write_text(">s\n" + "ACGT" * 30 + "\n")
    with pytest.raises(ValueError):
        seqwin.validate_fasta_entries(str(short), 2048)     # data/test/test_short.fasta behaviour


def test_encoder_matches_reference_numba_encoder():
    z = np.load(G / "tokens_2000.npz")
    tok = oenc.encode_windows([str(s) for s in z["seqs"]], 2000)
    assert tok.shape == z["tokens"].shape == (len(z["seqs"]), 6, 665)
    assert np.array_equal(tok, z["tokens"])
    assert (tok == 0).any()          # the N / IUPAC window produced unknown codons


@pytest.mark.parametrize("crop,lc", [(2048, 681), (500, 165)])
def test_encoder_matches_reference_numba_encoder_other_crops(crop, lc):
    """Crops 2048 / 500 (the `health` geometry and BASELINE config 3): tests/golden/tokens_more_crops.npz from the
    reference's numba encoder (tests/golden/make_token_goldens_more_crops.py)."""
    z = np.load(G / "tokens_more_crops.npz")
    tok = oenc.encode_windows([str(s) for s in z[f"seqs_{crop}"]], crop)
    assert tok.shape == z[f"tokens_{crop}"].shape == (len(z[f"seqs_{crop}"]), 6, lc)
    assert np.array_equal(tok, z[f"tokens_{crop}"])
    assert (tok == 0).any()


def test_frame_lengths_known_answers():
    # reference tests/unit/test_crop.py, test_inference_crop.py: 2000 -> 665, 1500 -> 498; SURVEY: 2048 -> 681, 500 -> 165
    for n, lc in [(2000, 665), (1500, 498), (2048, 681), (500, 165)]:
        assert oenc.codons_per_frame(n, n) == lc
        assert oenc.encode_window("ACGT" * (n // 4), n).shape == (6, lc)


def test_encoder_lookup_known_answers():
    # reference tests/unit/test_seqops_encode.py: unknown codon -> -1 ; complement of non-ACGT is N
    ids = oenc.encode_window("BBB" + "ACG" * 4, 15)
    assert ids[0, 0] == -1
    assert oenc.encode_window("acgtacgtacgt", 12, masking=True).max() == -1      # lower-case is unknown when masking
    assert oenc.encode_window("acgtacgtacgt", 12, masking=False).min() >= 0


def test_pred_to_dict_and_summary_match_reference():
    z = np.load(G / "pred_to_dict.npz")
    y = {k: z[k] for k in ["prediction", "reliability"] + [f"meta_{i}" for i in range(10)]}
    classes = ["bacteria", "phage", "eukarya", "archaea", "plasmid", "virus"]
    cm = {"num_classes": 6, "class": classes, "index": list(range(6))}
    data, _ = opp.pred_to_dict(y, 2000, cm)
    for k in ("pred_sum", "pred_var", "entropy", "energy", "ood"):
        assert np.array_equal(np.asarray(data[k]).view(np.uint16), z[k].view(np.uint16)), k     # bit-exact fp16
    assert np.array_equal(data["consensus"], z["consensus"])
    assert np.array_equal(np.concatenate(data["frag_pred"]), z["frag_pred"])
    assert np.array_equal(np.array([[d[k] for k in range(6)] for d in data["per_class_counts"]]), z["per_class_counts"])
    assert np.array_equal(data["host_contam"], z["host_contam"]) and np.array_equal(data["prophage_contam"], z["prophage_contam"])
    assert np.allclose([np.mean(x) for x in data["gc"]], z["gc_mean"], rtol=0, atol=0)
    df = opp.generate_summary(data, classes, list(range(6)))
    import io
    import pandas as pd
    buf = io.StringIO()
    df.to_csv(buf, sep="\t", index=False, float_format="%.3f")
    gold = (G / "summary.tsv").read_text()
    gold_df = pd.read_csv(io.StringIO(gold), sep="\t")
    got_df = pd.read_csv(io.StringIO(buf.getvalue()), sep="\t")
    common = [c for c in gold_df.columns if c in got_df.columns]
    assert [c for c in gold_df.columns if c not in ("terminal_repeats", "repeat_length")] == list(got_df.columns)
    pd.testing.assert_frame_equal(gold_df[list(got_df.columns)], got_df[common], check_dtype=False)


def test_viterbi_and_transition_costs_vs_reference_golden():
    """--crf decoding (helpers.py:345-449): the restatement equals the reference on every variant."""
    from oracle import postprocess as opp
    z = np.load(G / "viterbi.npz")
    classes = ["bacteria", "phage", "eukarya", "archaea", "plasmid", "virus"]
    user = {"Bacteria": {"phage": 0.25, "virus": 2.0}, "plasmid": {"archaea": 4.0}, "unknown": {"phage": 9.0}}
    assert np.array_equal(opp.build_transition_costs(classes, 2.0, "biological"), z["costs_bio2"])
    assert np.array_equal(opp.build_transition_costs(classes, 0.5, "uniform"), z["costs_uni05"])
    assert np.array_equal(opp.build_transition_costs(classes, 3.0, "biological", user), z["costs_user3"])
    assert np.array_equal(opp.build_transition_costs(classes[:4], 2.0), z["costs_4cls"])
    splits = np.cumsum(z["n_win"])[:-1]
    for name, lam, costs in [("potts2", 2.0, None), ("potts0", 0.0, None), ("bio2", 2.0, z["costs_bio2"]),
                             ("uni05", 0.5, z["costs_uni05"]), ("user3", 3.0, z["costs_user3"])]:
        got = np.concatenate([opp.viterbi_decode(p, lam, costs) for p in np.split(z["prediction"], splits)])
        assert np.array_equal(got, z[f"path_{name}"]), name
    assert np.array_equal(z["path_potts0"], z["prediction"].argmax(1))        # lambda = 0 is the plain argmax
    b = z["binary_logit"]
    assert np.array_equal(opp.viterbi_decode(np.concatenate([np.zeros_like(b), b], -1), 2.0), z["binary_path"])


def _legacy_post_fixture():
    z = np.load(G / "legacy_post.npz")
    ood = {k[4:]: z[k] for k in z.files if k.startswith("ood_") and k != "ood_windows"}
    meta = tuple(z[f"meta_{i}"] for i in range(10))
    return z, ood, meta


def test_legacy_postprocess_vs_reference_golden():
    """pred_to_dict_legacy + generate_summary_legacy + the pickled reliability model: the restatement
    reproduces the reference's TSV (both label sets) and the per-window reliability of the sklearn model."""
    import io
    import pandas as pd
    from oracle import legacy as oleg
    z, ood, meta = _legacy_post_fixture()
    all_labels = {0: "bacteria", 1: "phage", 2: "eukarya", 3: "archaea"}
    second = {1: "eukarya", 2: "archaea", 3: "bacteria", 0: ""}
    for tag, labels in (("default", ["non-phage", "phage", "non-phage", "non-phage"]), ("all", list(all_labels.values()))):
        cols, ood_w = oleg.summary_legacy(z["output"], z["embedding"], meta, 2000, ood, labels, all_labels, second, 1)
        assert np.allclose(ood_w, z["ood_windows"], rtol=0, atol=1e-6)
        got = pd.DataFrame(cols)
        buf = io.StringIO()
        got.to_csv(buf, sep="\t", index=False, float_format="%.3f")
        got = pd.read_csv(io.StringIO(buf.getvalue()), sep="\t", keep_default_na=False)
        want = pd.read_csv(G / f"summary_legacy_{tag}.tsv", sep="\t", keep_default_na=False)
        for col in got.columns:
            assert got[col].tolist() == want[col].tolist(), (tag, col)


def test_fragment_windows_with_dustmask_flow_vs_reference():
    """fragment_generator(dustmask=True) of the reference with pydustmasker stubbed by the oracle's SDUST
    (tests/golden/make_dust_flow_goldens.py): the oracle's window generator with the same soft-masks gives the same rows --
    windows cut from the soft-masked, upper-cased record, case-sensitive base counts, gc_skew strings, the short pass."""
    import zlib
    from oracle import dust as odust
    from tests.helpers import dust_contigs
    gold = json.loads((G / "fragments_dustmask.json").read_text())
    recs = dust_contigs()
    masks = {n: odust.mask_bits(s.strip().upper()) for n, s in recs}
    assert any(np.asarray(m).any() for m in masks.values())
    for key, kw in {"2000_1500": dict(fragsize=2000, stride=1500), "2000_1500_short": dict(fragsize=2000, stride=1500, min_len=500, max_len=1999),
                    "500_500": dict(fragsize=500, stride=500)}.items():
        got = []
        for w in seqwin.fragment_windows(recs, kw["fragsize"], kw["stride"], softmasks=masks, min_len=kw.get("min_len"), max_len=kw.get("max_len")):
            f = w.csv().split(",")
            got.append([zlib.crc32(f[0].encode()), len(f[0])] + f[1:])
        assert got == gold[key], key
    masked = [r for r in gold["500_500"] if int(r[7]) + int(r[8]) + int(r[9]) + int(r[10]) < r[1]]
    assert len(masked) >= 8                                    # soft-masked bases do not count as A / C / G / T
