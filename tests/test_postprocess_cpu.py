"""CPU tests of host-side post-processing logic (no GPU calls)."""
import json
from pathlib import Path

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from jaeger_b200 import postprocess as pp
from jaeger_b200 import prophage as ppro
from jaeger_b200.weights import load_npz_weights, save_npz_weights
from jaeger_b200.modelspec import init_random, parse_project, standin_1p4m_config
from oracle import postprocess as opp
from oracle import prophage as opro

G = Path(__file__).resolve().parent / "golden"
CLASSES = ["bacteria", "phage", "eukarya", "archaea", "plasmid", "virus"]


def test_window_summaries_match_reference_helper():
    z = np.load(G / "pred_to_dict.npz")
    off = pp._split_points(z["meta_2"].astype(np.int32))
    cm = dict(enumerate(CLASSES))
    got = pp.window_summaries(z["frag_pred"], off, cm)
    want = [opp.get_window_summary(z["frag_pred"][off[i]:off[i + 1]], cm) for i in range(len(off) - 1)]
    assert got == want
    import pandas as pd
    gold = pd.read_csv(G / "summary.tsv", sep="\t")          # written by the reference's generate_summary
    assert got == gold["window_summary"].tolist()


def test_merge_ranges_match_reference_golden():
    for case in json.loads((G / "merge_ranges.json").read_text()):
        assert ppro.merge_overlapping_ranges(case["inp"]) == case["out"]
        assert opro.merge_overlapping_ranges(case["inp"]) == case["out"]


def test_knee_point_product_equals_oracle():
    """Penalty selection of `segment` (postprocess/prophages.py:563-573).  x = breakpoint counts at pen 1..9: non-increasing,
    usually TIED.  The oracle makes kneed's own SciPy calls (interp1d(x, y)(x), argrelextrema); the product writes the tie
    rule out.  Both must agree on every tied count vector, and on the known answers of an interp1d-faithful Kneedle."""
    from scipy.interpolate import interp1d
    x0 = np.array([9, 7, 5, 5, 3, 3, 3, 2, 2])
    assert interp1d(x0, np.arange(9))(x0).tolist() == [0, 1, 3, 3, 6, 6, 6, 8, 8]           # not the identity on ties
    assert ppro._interp_at_nodes(x0.astype(float), np.arange(9.0)).tolist() == [0, 1, 3, 3, 6, 6, 6, 8, 8]
    rng = np.random.default_rng(0)
    n_tied = n_knee = 0
    for it in range(4000):
        n = int(rng.integers(2, 10))
        x = np.sort(rng.integers(2, 14 if it % 2 else 40, n))[::-1]
        y = list(range(n))
        got, want = ppro.knee_point(x, y), opro.knee_locator(x, y)
        assert got == want, (x.tolist(), got, want)
        n_tied += len(set(x.tolist())) < n
        n_knee += got is not None
    assert n_tied >= 1000 and n_knee >= 1000, (n_tied, n_knee)
    # counts where treating interp1d as the identity picks another penalty (round-1 bug)
    assert ppro.knee_point([11, 7, 5, 5, 5, 5, 3], list(range(7))) == 7.0
    assert ppro.knee_point([9, 8, 5, 5, 4, 2, 2, 2], list(range(8))) == 4.0
    assert opro.knee_locator([11, 7, 5, 5, 5, 5, 3], list(range(7))) == 7.0
    assert opro.knee_locator([9, 8, 5, 5, 4, 2, 2, 2], list(range(8))) == 4.0
    assert ppro.knee_point([4, 4, 4], [0, 1, 2]) is None and opro.knee_locator([4, 4, 4], [0, 1, 2]) is None
    assert ppro.knee_point([5], [0]) is None and opro.knee_locator([5], [0]) is None


@settings(max_examples=300, deadline=None)
@given(st.lists(st.integers(2, 60), min_size=2, max_size=9))
def test_knee_point_property_tied_counts(counts):
    x = np.sort(np.array(counts))[::-1]
    y = list(range(len(x)))
    assert ppro.knee_point(x, y) == opro.knee_locator(x, y)


def test_x_axis_clamp_reference_known_answer():
    # reference tests/unit/test_prophage_plots.py:44-72
    assert opro.window_x(5, 1500, 4857) == [0, 1500, 3000, 4500, 4857]


def test_weights_npz_roundtrip(tmp_path):
    spec = parse_project(standin_1p4m_config())
    w = init_random(spec, 0)
    save_npz_weights(tmp_path / "m.weights.npz", w)
    r = load_npz_weights(tmp_path / "m.weights.npz")
    assert np.array_equal(r["embedding"], w["embedding"])
    assert np.array_equal(r["classifier"][0]["kernel"], w["classifier"][0]["kernel"])
    assert np.array_equal(r["layers"][4]["blocks"][1]["conv2"]["kernel"], w["layers"][4]["blocks"][1]["conv2"]["kernel"])


def test_sdust_oracle_masks_low_complexity_only():
    from oracle import dust as odust
    rng = np.random.default_rng(0)
    rnd = lambda n: "".join(rng.choice(list("ACGT"), n))          # noqa: E731
    s = rnd(300) + "A" * 40 + rnd(200) + "AC" * 14 + rnd(300) + "NNNNN" + "T" * 21 + rnd(100)
    iv = odust.sdust_intervals(s)
    assert iv[0][0] <= 300 and iv[0][1] >= 340                    # the poly-A run
    assert any(535 <= a <= 545 and 565 <= b <= 575 for a, b in iv)   # the (AC)n repeat
    assert iv[-1][0] >= 873 and iv[-1][1] >= 890                  # poly-T after the N break, not merged across it
    m = odust.mask(s)
    assert m[310] == "a" and m[100] == s[100] and m[868:873] == "NNNNN"
    # random sequence: almost nothing is low-complexity at T = 20
    assert odust.mask_bits(rnd(20000)).mean() < 0.002
    assert odust.sdust_intervals("ACG") == [] and odust.sdust_intervals("") == []


def test_terminal_repeat_summary_arithmetic_vs_reference_golden():
    """get_alignment_summary (utils/termini.py:17-88) on mock alignment results: the oracle's and the
    product's coordinate arithmetic reproduce the reference for DTR / ITR / LTR, with and without gaps."""
    from jaeger_b200.termini import summary_fields
    from oracle import termini as ot
    cases = json.loads((G / "termini_summary.json").read_text())
    assert len(cases) == 80
    for cs in cases:
        i, want = cs["inp"], cs["out"]
        got = summary_fields(i["cols"], i["qgaps"], i["rgaps"], i["iden"], i["score"], i["end_query"], i["end_ref"], i["seq_len"],
                             i["n"], i["type"])
        orc = ot.alignment_summary({"cols": i["cols"], "qgaps": i["qgaps"], "rgaps": i["rgaps"], "iden": i["iden"], "score": i["score"],
                                    "end_query": i["end_query"], "end_ref": i["end_ref"]}, i["seq_len"], "r", i["n"], i["type"])
        for k, v in want.items():
            assert got[k] == v, (k, got[k], v)
            assert orc[k] == v, (k, orc[k], v)
    assert {c["out"]["terminal_repeats"] for c in cases} == {"DTR", "ITR", "LTR_DTR"}


def test_terminal_repeat_oracle_known_answers():
    """Hand-checkable alignments: exact direct / inverted repeats, one mismatch (kept: 2*119 - 100 > 2*60),
    one deleted base (gap in the query line), nothing in random sequence."""
    from oracle import termini as ot
    rng = np.random.default_rng(1)
    rnd = lambda n: "".join(rng.choice(list("ACGT"), n))
    core = rnd(120)
    a = ot.sw_align(core + rnd(300), rnd(300) + core)
    assert (a["score"], a["cols"], a["qgaps"], a["rgaps"], a["iden"], a["end_query"], a["end_ref"]) == (240, 120, 0, 0, 120, 119, 419)
    mut = core[:60] + ("A" if core[60] != "A" else "C") + core[61:]
    b = ot.sw_align(core, mut)
    assert (b["score"], b["cols"], b["iden"]) == (2 * 119 - 100, 120, 119)
    c = ot.sw_align(core[:58] + core[59:], core)
    assert (c["score"], c["cols"], c["qgaps"], c["rgaps"]) == (2 * 119 - 100, 120, 1, 0)
    d = ot.sw_align("ACGTNNNNACGT", "ACGTACGTACGT")          # letters outside ACGT score 0 against everything
    assert d["score"] == 16 and d["cols"] == 12
    rows = ot.scan_for_terminal_repeats([("x", core + rnd(2500) + ot.reverse_complement(core)), ("y", rnd(2600)), ("z", rnd(100))], 2000)
    assert [r["terminal_repeats"] for r in rows] == ["ITR", None] and rows[0]["repeat_length"] == 120


def test_write_fasta_from_results(tmp_path):
    """collect.py:611-639: records named in the phage table, 70 letters per line, input order."""
    from jaeger_b200.engine import WindowSource
    recs = [("a", "ACGT" * 40), ("b,x", "G" * 70), ("c", "T" * 5), ("d", "ACG" * 50)]
    (tmp_path / "p.tsv").write_text("contig_id\tlength\nd\t150\na\t160\nb,x\t70\n")
    n = pp.write_fasta_from_results(WindowSource(records=recs).load(), tmp_path / "p.tsv", tmp_path / "o.fa")
    want = ""
    for name, seq in recs:
        if name in ("a", "b,x", "d"):
            want += f">{name}\n" + "".join(seq[i:i + 70] + "\n" for i in range(0, len(seq), 70))
    assert n == 3 and (tmp_path / "o.fa").read_text() == want
    assert pp.write_fasta_from_results(WindowSource(records=recs).load(), tmp_path / "missing.tsv", tmp_path / "e.fa") == 0


def _same(a, b):
    if b is None:
        return a is None or a != a
    if isinstance(b, float):
        return a is not None and abs(float(a) - b) < 1e-12
    return a == b


def test_termini_table_oracle_vs_reference_golden():
    """The reference's scan_for_terminal_repeats (utils/termini.py:91-189) run with the aligner stubbed by the oracle's
    sw_align (tests/golden/make_termini_goldens.py): the oracle's restatement of the flow around the aligner -- scan
    length, ITR / DTR choice, > 12 column rule, coordinates, front / rear strings (rear reverse-complemented back
    for an ITR) -- gives the same table."""
    from oracle import termini as ot
    from tests.helpers import repeat_contigs
    want = json.loads((G / "termini_table.json").read_text())
    got = sorted(ot.scan_for_terminal_repeats(repeat_contigs(), 2000), key=lambda r: r["contig_id"])
    assert [r["contig_id"] for r in got] == [r["contig_id"] for r in want] and len(want) == 11
    for g, w in zip(got, want):
        for k, v in w.items():
            assert _same(g[k], v), (w["contig_id"], k, g[k], v)
    by = {r["contig_id"]: r for r in want}
    itr = by["itr___comma"]                               # rear is given as the contig holds it: the reverse complement of front
    assert itr["terminal_repeats"] == "ITR" and itr["rear"] == ot.reverse_complement(itr["front"]) and len(itr["front"]) == 120
    assert by["gap"]["rear"].count("-") == 2 and by["qgap"]["front"].count("-") == 1


def test_alignment_lines_from_traceback_operations():
    """Host side of the `front` / `rear` / `attL` / `attR` strings: the letters are read off the loaded FASTA bytes
    along the traceback operations the trace kernel writes (last column first)."""
    from jaeger_b200.termini import JOB, alignment_lines, reverse_complement_bytes
    from oracle import termini as ot
    rng = np.random.default_rng(8)
    rnd = lambda n: "".join(rng.choice(list("ACGT"), n))
    core = rnd(150)
    left = rnd(40) + core + rnd(25)
    right_plain = rnd(10) + core[:70] + core[72:] + rnd(60)               # two bases missing from the reference line
    right_q = rnd(33) + core[:60] + "TG" + core[60:] + rnd(5)             # two extra bases: gap in the query line
    for right, inverted in ((right_plain, 0), (right_q, 0), (right_plain, 1), ("acgtn" + right_q, 1)):
        stored = ot.reverse_complement(right) if inverted else right      # what the contig holds at r0
        host = np.frombuffer(("NN" + left + "NNN" + stored).encode(), np.uint8)
        a = ot.sw_align(left, ot.reverse_complement(stored) if inverted else stored)
        assert a["qgaps"] + a["rgaps"] >= 1 and a["cols"] > 140
        ops = np.array([2 if q == "-" else 3 if r == "-" else 1 for q, r in zip(a["qline"], a["rline"])], np.uint8)[::-1]
        job = np.array([(2, 2 + len(left) + 3, len(stored), inverted, len(left), 0)], dtype=JOB)[0]
        ql, rl = alignment_lines(host, job, a["end_query"], a["end_ref"], ops)
        assert (ql, rl) == (a["qline"], a["rline"])
    assert reverse_complement_bytes(np.frombuffer(b"ACGTNacgtRYx-", np.uint8)).tobytes().decode() == ot.reverse_complement("ACGTNacgtRYx-")


def test_prophage_report_oracle_vs_reference_golden(tmp_path):
    """The reference's prophage_report (postprocess/prophages.py:706-873, refined_boundaries=None) run with the
    aligner stubbed by the oracle's sw_align: the oracle's restatement of the flank coordinates, the ITR / DTR choice,
    get_prophage_alignment_summary's arithmetic, n% / gc% / reject and the TSV formatting gives the same file."""
    import pandas as pd
    from oracle import termini as ot
    from tests.helpers import prophage_genomes
    recs, cords = prophage_genomes()
    rows = ot.prophage_report(recs, cords, fsize=2000, stride=1500)
    df = pd.DataFrame(rows)
    df["contig_id"] = df["contig_id"].apply(lambda x: x.replace("___", ","))
    df.to_csv(tmp_path / "p.tsv", sep="\t", index=False, float_format="%.3f")
    assert (tmp_path / "p.tsv").read_text() == (G / "prophages_jaeger.tsv").read_text()
    assert [r["att_type"] for r in rows] == ["DTR", "ITR", "DTR", "DTR", "DTR"] and rows[2]["reject"] is True


def test_refinement_oracle_on_reference_known_answers(tmp_path):
    """The known answers of the reference's tests/unit/test_refinement.py (the module itself needs polars and cannot
    be imported here), restated against oracle/refine.py; each block names the reference test."""
    from oracle import refine as orf
    C = orf.CLASSES
    # test_add_score_features_basic
    z = np.zeros((1, 6), np.float32); z[0, 0] = 3.0; z[0, 1] = 1.0
    f = orf.add_score_features(z)
    assert f["top_class"][0] == "phage" and f["second_class"][0] == "virus"
    assert f["top_logit"][0] == pytest.approx(3.0) and f["margin"][0] == pytest.approx(2.0)
    assert 0.0 < f["top_prob"][0] <= 1.0 and f["entropy"][0] >= 0.0
    taus = {c: {"logit": 1.0, "margin": 0.5} for c in C}
    # test_refine_merges_before_abstain
    z = np.zeros((1, 6), np.float32); z[0, 3] = 0.6; z[0, 4] = 0.4
    assert orf.refine(orf.add_score_features(z), taus)[0] == "bacteria_or_plasmid"
    # test_refine_abstains_low_confidence
    z = np.zeros((1, 6), np.float32); z[0, 3] = 0.3; z[0, 5] = 0.1
    assert orf.refine(orf.add_score_features(z), taus)[0] == "unknown"
    # test_aggregate_contig_gated_drops_unknown
    S = np.array([[5.0, 0, 0, 0, 0, 0], [0.0, 0, 0, 0, 0, 0]])
    out = orf.aggregate_contig(["c1", "c1"], S, np.array(["phage", "unknown"], dtype=object), np.array([5.0, 0.0]), mode="gated", min_windows=1)
    assert out["c1"]["n_windows_used"] == 1 and out["c1"]["contig_call"] == "phage"
    # test_aggregate_contig_weighted_uses_margin
    S = np.array([[4.0, 0, 0, 0, 0, 0], [0.0, 2.0, 0, 0, 0, 0]])
    out = orf.aggregate_contig(["c1", "c1"], S, np.array(["phage", "virus"], dtype=object), np.array([4.0, 2.0]), mode="weighted", min_windows=1)
    assert out["c1"]["contig_call"] == "phage" and out["c1"]["total_weight"] == pytest.approx(6.0)
    # test_aggregate_contig_merge_split_half / _full
    S = np.array([[0.0, 0.0, 0.0, 2.0, 1.5, 0.0]])
    out = orf.aggregate_contig(["c1"], S, np.array(["bacteria_or_plasmid"], dtype=object), np.array([0.5]), mode="gated", min_windows=1, merge_split="half")
    assert out["c1"]["contig_call"] == "bacteria" and out["c1"]["bacteria_score"] == 1.0 and out["c1"]["plasmid_score"] == 0.75
    S = np.array([[0.0, 0.0, 0.0, 1.0, 1.2, 0.0]])
    out = orf.aggregate_contig(["c1"], S, np.array(["bacteria_or_plasmid"], dtype=object), np.array([0.5]), mode="gated", min_windows=1, merge_split="full")
    assert out["c1"]["contig_call"] == "plasmid"
    # default min_windows = 3 drops the one-window contig (refinement.py:214)
    assert orf.aggregate_contig(["c1"], S, np.array(["plasmid"], dtype=object), np.array([0.5])) == {}
    # test_save_and_load_refinement / test_load_refinement_rejects_wrong_model (the product's loader; file as save_refinement writes it)
    import yaml
    from jaeger_b200.refine import load_refinement, tau_vector
    path = tmp_path / "refine.yaml"
    full = {c: {"logit": 0.25 * i, "margin": 0.1 * i, "n": 40} for i, c in enumerate(C)}
    full["eukarya"] = {"logit": float("-inf"), "margin": float("-inf"), "n": 3}
    path.write_text(yaml.safe_dump({"schema_version": 1, "jaeger_model": "test_model", "quantile": 0.05, "classes": C, "taus": full}, sort_keys=False))
    meta = load_refinement(path, expect_model="test_model")
    assert meta["taus"]["virus"]["logit"] == pytest.approx(0.25)
    tv = tau_vector(meta["taus"])
    assert tv.shape == (12,) and tv[5] == -np.inf and tv[11] == -np.inf and tv[7] == pytest.approx(0.1)
    with pytest.raises(ValueError):
        load_refinement(path, expect_model="model_b")


def _model_tree(root: Path):
    """A model registry like the reference's data/models: nested `model` directories with graphs, class / project files,
    an embedding-only graph, a legacy .h5 and unrelated files."""
    for rel in ["a/model/jaeger_38341_1.4M_fragment_graph/variables", "a/model/jaeger_38341_1.4M_fragment_embedding_graph",
                "b/deep/er/model/jaeger_57341_1.5M_fragment_graph", "c/not_model/jaeger_1_x_fragment_graph", "d/model"]:
        (root / rel).mkdir(parents=True)
    for rel in ["a/model/jaeger_38341_1.4M_fragment_classes.yaml", "a/model/jaeger_38341_1.4M_fragment_project.yaml",
                "a/model/jaeger_38341_1.4M_fragment.weights.h5", "a/model/README.txt", "b/deep/er/model/jaeger_57341_1.5M_fragment_classes.yaml",
                "c/not_model/jaeger_1_x_fragment_classes.yaml", "d/model/WRes_1024.h5"]:
        (root / rel).write_text("x")


def test_available_models_scan_matches_reference(tmp_path):
    """`AvailableModels(path).info` and `get_model_id` (utils/misc.py:334-396) on a synthetic registry: same model names,
    same graph / classes / project / weights paths.  Compared with the reference's own class when the checkout is mounted."""
    import sys
    from jaeger_b200.predict import available_models, get_model_id
    _model_tree(tmp_path)
    got = available_models([str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "c"), str(tmp_path / "d")])
    assert set(got) >= {"jaeger_38341_1.4M_fragment", "jaeger_38341_1.4M_fragment_embedding", "jaeger_57341_1.5M_fragment"}
    assert "jaeger_1_x_fragment" not in got                                  # only directories named `model` are scanned
    m = got["jaeger_38341_1.4M_fragment"]
    assert m["graph"].name == "jaeger_38341_1.4M_fragment_graph" and m["classes"].name.endswith("_classes.yaml")
    assert m["project"].name.endswith("_project.yaml") and m["weights"].name.endswith(".weights.h5")
    assert get_model_id("jaeger_38341_1.4M_fragment") == "38341_1.4M"
    ref_src = Path("/root/reference/src")
    if not ref_src.exists():
        return
    sys.path.insert(0, str(ref_src))
    try:
        from jaeger.utils.misc import AvailableModels, get_model_id as ref_id
    finally:
        sys.path.remove(str(ref_src))
    want = AvailableModels(path=[str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "c"), str(tmp_path / "d")]).info
    assert set(want) == set(got)
    for name, parts in want.items():
        assert {k: Path(v) for k, v in parts.items() if v is not None} == {k: Path(v) for k, v in got[name].items()}, name
    for name in want:
        if name.count("_") >= 2:
            assert ref_id(name) == get_model_id(name)


def test_driver_rejects_cpu_and_bad_precision_before_touching_a_device(tmp_path):
    from jaeger_b200.predict import run_core
    fa = tmp_path / "x.fasta"
    fa.write_text(">a\n" + "ACGT" * 600 + "\n")
    with pytest.raises(RuntimeError, match="no CPU path"):
        run_core(input=str(fa), output=str(tmp_path / "o"), model="standin", allow_random_weights=True, cpu=True)
    for backend in ("onnx", "quantized", "xla"):          # the reference's alternate backends are not this engine's
        with pytest.raises(RuntimeError, match="belong to the reference"):
            run_core(input=str(fa), output=str(tmp_path / "o"), model="standin", allow_random_weights=True, **{backend: True})
    with pytest.raises(ValueError, match="precision"):
        run_core(input=str(fa), output=str(tmp_path / "o"), model="standin", allow_random_weights=True, precision="int8")
    for prec in ("fp32", "bf16"):            # one numeric mode: another precision is refused, not silently ignored
        with pytest.raises(ValueError, match="not available on the B200 engine"):
            run_core(input=str(fa), output=str(tmp_path / "o"), model="standin", allow_random_weights=True, precision=prec)
    with pytest.raises(ValueError, match="RANDOM-INITIALISED"):
        run_core(input=str(fa), output=str(tmp_path / "o"), model="standin")
    with pytest.raises(ValueError, match="--legacy-weights"):
        run_core(input=str(fa), output=str(tmp_path / "o"))            # default model is `default`, which needs its weights


def test_segment_flow_vs_reference_golden():
    """The reference's own `logits_to_df_v2` + `segment` (postprocess/prophages.py:99-153, 524-602) run with ruptures / kneed
    replaced by stubs that return the oracle's restatements (tests/golden/make_segment_goldens.py): the oracle's `segment`
    reproduces every range and score -- including the reference's quirks: the stretch before the first breakpoint is never a
    range, scores are those of the ranges BEFORE the interval merge, contigs not longer than the cutoff are skipped."""
    cases = json.loads((G / "segment_cases.json").read_text())
    assert len(cases) == 12
    assert sum(len(set(c["counts"])) < len(c["counts"]) for c in cases) >= 10      # tied breakpoint counts are the rule
    for c in cases:
        rng = np.random.default_rng(c["seed"])
        z = rng.normal(0.0, 1.2, (c["t"], 6)).astype(np.float32)
        z[:, 0] += 2.0
        for a, b, lift in c["islands"]:
            z[a:b, 1] += lift
        if c["skipped"]:
            assert c["ranges"] == [] and c.get("short")
            continue
        ranges, scores = opro.segment(opro.smooth_scores(z)[:, 1], c["sens"])
        assert [list(map(int, r)) for r in ranges] == c["ranges"], (c["seed"], ranges, c["ranges"])
        assert np.allclose(np.asarray(scores, dtype=np.float64), np.asarray(c["scores"]), atol=1e-9), c["seed"]
    by = {c["seed"]: c for c in cases}
    assert by[4]["ranges"] == [[3, 32], [469, 500]]              # the first range starts at the first breakpoint, never at 0
    assert len(by[5]["ranges"]) == 2 and len(by[5]["scores"]) == 5


def test_refined_columns_join_vs_reference_generate_summary(tmp_path):
    """collect.py:534-550: the reference's own generate_summary(refined_contig=...) (tests/golden/make_refine_summary_golden.py)
    against the product's join -- same six columns in the same order, NaN cells for contigs without a refined call (so the
    integer counts print as floats), refined rows that match no contig dropped."""
    import pandas as pd
    from jaeger_b200.refine import MERGE_COLUMNS, merge_into_summary
    base = pd.read_csv(G / "summary.tsv", sep="\t")
    base["contig_id"] = base["contig_id"].str.replace(",", "___")            # the join runs before the ids are restored
    refined = pd.DataFrame(json.loads((G / "refined_contig.json").read_text()))
    out = merge_into_summary(base, refined)
    out["contig_id"] = out["contig_id"].str.replace("___", ",")
    out.to_csv(tmp_path / "s.tsv", sep="\t", index=False, float_format="%.3f")
    want = (G / "summary_refined.tsv").read_text()
    assert (tmp_path / "s.tsv").read_text() == want
    assert want.splitlines()[0].split("\t")[-5:] == MERGE_COLUMNS[1:]


def test_refine_window_labels_vs_reference_source():
    """add_score_features + refine executed from the reference's own source (postprocess/refinement.py:39-137) behind a
    four-method stand-in for polars (tests/golden/make_refine_window_goldens.py): top / second class -- exact ties between
    logits included --, margins and refined labels of every window equal the oracle's, with and without the merge rules."""
    from oracle import refine as orf
    from tests.helpers import refine_case
    gold = json.loads((G / "refine_windows.json").read_text())
    z, _, _, taus = refine_case()
    feat = orf.add_score_features(z)
    for tag, tt, kw in (("case", taus, {}), ("no_merge", {c: {"logit": 0.0, "margin": 0.3} for c in orf.CLASSES}, dict(merge_bp=False, merge_pv=False))):
        g = gold[tag]
        assert feat["top_class"].tolist() == g["top_class"] and feat["second_class"].tolist() == g["second_class"]
        assert np.array_equal(feat["margin"], np.asarray(g["margin"]))
        assert orf.refine(feat, tt, **kw).tolist() == g["refined_prediction"], tag
    assert {"unknown", "bacteria_or_plasmid", "virus_any"} <= set(gold["case"]["refined_prediction"])
    assert gold["case"]["top_class"][21] == "phage" and gold["case"]["second_class"][21] == "plasmid"     # all-equal row: argmax first, argsort[-2]


def test_refine_contig_aggregation_vs_reference_source():
    """aggregate_contig executed from the reference's own source (postprocess/refinement.py:140-247) behind a small evaluator
    of the polars expressions it builds (tests/golden/make_refine_window_goldens.py): gating, margin weights, merged-label
    multipliers, the min_windows filter, the contig-level top-2 and hedged calls equal the oracle's in all three modes."""
    from oracle import refine as orf
    from tests.helpers import refine_case
    gold = json.loads((G / "refine_windows.json").read_text())["contigs"]
    z, offsets, headers, taus = refine_case()
    feat = orf.add_score_features(z)
    lab = orf.refine(feat, taus)
    ids = np.repeat(np.array(headers, dtype=object), np.diff(offsets))
    for mode, split, allow in (("gated", "half", False), ("weighted", "full", True), ("unweighted", "half", True)):
        want = gold[f"{mode}_{split}_{int(allow)}"]
        got = orf.aggregate_contig(ids, z, lab, feat["margin"], mode=mode, min_windows=3, merge_split=split,
                                   allow_merged_contig_call=allow, contig_hedge_margin=5.0)
        assert list(got) == list(want), mode
        for cid, w in want.items():
            for k, v in w.items():
                if isinstance(v, float):
                    assert got[cid][k] == pytest.approx(v, rel=1e-12, abs=1e-12), (mode, cid, k)
                else:
                    assert got[cid][k] == v, (mode, cid, k, got[cid][k], v)
    assert {"virus_any", "bacteria_or_plasmid"} <= {w["contig_call"] for w in gold["unweighted_half_1"].values()}


def _all_partitions(n: int, min_size: int):
    """Every list of segment ends (ascending, last = n) whose segments are all >= min_size long."""
    def rec(start):
        if n - start >= min_size:
            yield [n]
        for end in range(start + min_size, n - min_size + 1):
            for rest in rec(end):
                yield [end] + rest
    return list(rec(0))


def _objective(x, ends, pen):
    tot, a = 0.0, 0
    for b in ends:
        seg = x[a:b]
        tot += float(((seg - seg.mean()) ** 2).sum())
        a = b
    return tot + pen * (len(ends) - 1)


def test_optimal_partition_is_the_exact_optimum_on_short_signals():
    """ruptures' KernelCPD(kernel="linear", min_size=3).predict(pen=p) (prophages.py:540-556) minimises
    sum of within-segment squared deviations + pen per change point; the library is not installable, so the
    restatement is held against a brute force over every admissible partition of short signals.
    Tie-break of the restatement (and of the device kernel, tests/test_gpu_parity.py): the dynamic programme
    keeps, for every prefix, the EARLIEST last change point among equal objectives (first arg-min), i.e. the
    longest last segment; exact ties have measure zero on real score tracks."""
    rng = np.random.default_rng(11)
    for case in range(120):
        n = int(rng.integers(3, 15))
        x = rng.normal(size=n) + (rng.random(n) < 0.3) * rng.normal(scale=4.0)
        if case % 4 == 0:
            x = np.round(x)                                  # integer signals: many exact ties
        pen = float(rng.choice([0.0, 0.25, 1.0, 2.0, 9.0]))
        parts = _all_partitions(n, 3)
        objs = np.array([_objective(x, p, pen) for p in parts])
        got = opro.optimal_partition(x, pen, 3)
        assert got in parts
        assert _objective(x, got, pen) <= objs.min() + 1e-9, (case, got, parts[int(objs.argmin())])
        best = [p for p, o in zip(parts, objs) if o <= objs.min() + 1e-9]
        if len(best) == 1:
            assert got == best[0]
    # the stated tie-break on an all-tie signal: no change point at pen 0 on a constant track
    assert opro.optimal_partition(np.zeros(9), 0.0, 3) == [9]
    assert opro.optimal_partition(np.zeros(9), 1.0, 3) == [9]


def test_prophage_boundary_refinement_vs_reference(tmp_path):
    """Gene-aware prophage boundaries (postprocess/prophage_boundaries.py:52-193): the known answers of the reference's
    tests/unit/test_prophage_boundaries.py, and `refine_regions` against what the reference's own `refine_prophage_boundaries` returned
    for the test genomes with its gene caller replaced by a fixed gene table (tests/golden/refined_boundaries.json); gene tables in the
    three accepted formats give the same intervals."""
    from jaeger_b200 import prophage_boundaries as pb
    from tests.helpers import prophage_gene_calls, prophage_genomes
    genes = [(100, 200), (300, 400)]
    assert pb.refine_boundary(50, genes, "left") == 50 and pb.refine_boundary(250, genes, "right") == 250
    assert pb.refine_boundary(150, [(100, 200)], "left") == 100 and pb.refine_boundary(150, [(100, 200)], "right") == 200
    assert pb.refine_boundary(900, [(0, 1000)], "left", max_extension=50) == 850
    assert pb.refine_boundary(100, [(0, 1000)], "right", max_extension=50) == 150
    assert pb.refine_region(150, 550, [(100, 200), (500, 600)]) == (100, 600)
    assert pb.refine_region(250, 700, [(100, 200), (500, 600)]) == (250, 700)
    with pytest.raises(ValueError, match="side must be 'left' or 'right'"):
        pb.refine_boundary(50, [(0, 100)], "upstream")
    recs, cords = prophage_genomes()
    calls = prophage_gene_calls()
    want = {k: [tuple(r) for r in v] for k, v in json.loads((G / "refined_boundaries.json").read_text()).items()}
    got = pb.refine_regions(cords, [n for n, _ in recs], [len(s) for _, s in recs], 2000, 1500, lambda header, ci: calls.get(header))
    assert got == want
    assert any((r[0], r[1]) != (r[2], r[3]) for r in got["genome1___with___commas"])
    assert got["genome1___with___commas"][2] == (450000, 455000, 446000, 459000)          # both extensions capped at 2 * fsize
    # the same calls as GFF3 (1-based closed), BED (0-based half-open), TSV (1-based closed)
    gff, bed, tsv = tmp_path / "g.gff", tmp_path / "g.bed", tmp_path / "g.tsv"
    with open(gff, "w") as a, open(bed, "w") as b, open(tsv, "w") as c:
        a.write("##gff-version 3\n")
        c.write("contig\tbegin\tend\n")
        for contig, iv in calls.items():
            name = contig.replace("___", ",")
            for s, e in iv:
                a.write(f"{name} a description\tprodigal\tCDS\t{s + 1}\t{e}\t.\t+\t0\tID=x\n")
                a.write(f"{name}\tprodigal\tregion\t1\t9\t.\t+\t0\tID=y\n")
                b.write(f"{name}\t{s}\t{e}\n")
                c.write(f"{name}\t{s + 1}\t{e}\n")
    for path in (gff, bed, tsv):
        table = pb.load_gene_table(path)
        assert table == {k: sorted(v) for k, v in calls.items()}, path
        src = pb.gene_source(table, None)
        assert pb.refine_regions(cords, [n for n, _ in recs], [len(s) for _, s in recs], 2000, 1500, src) == want


def test_prophage_report_with_refined_boundaries_oracle_vs_reference_golden(tmp_path):
    """The reference's prophage_report run on the ends its own refine_prophage_boundaries produced (aligner and gene caller stubbed,
    tests/golden/make_termini_goldens.py): the oracle writes the same file; the refined ends bring a planted repeat into reach
    that the raw window-grid ends miss."""
    import pandas as pd
    from oracle import termini as ot
    from tests.helpers import prophage_genomes
    recs, cords = prophage_genomes()
    refined = {k: [tuple(r) for r in v] for k, v in json.loads((G / "refined_boundaries.json").read_text()).items()}
    rows = ot.prophage_report(recs, cords, fsize=2000, stride=1500, refined_boundaries=refined)
    df = pd.DataFrame(rows)
    df["contig_id"] = df["contig_id"].apply(lambda x: x.replace("___", ","))
    df.to_csv(tmp_path / "p.tsv", sep="\t", index=False, float_format="%.3f")
    assert (tmp_path / "p.tsv").read_text() == (G / "prophages_jaeger_refined.tsv").read_text()
    assert (G / "prophages_jaeger_refined.tsv").read_text() != (G / "prophages_jaeger.tsv").read_text()
