"""CPU tests of host-side post-processing logic (no GPU calls)."""
import json
from pathlib import Path

import numpy as np

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
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(2, 10))
        x = np.sort(rng.integers(2, 40, n))[::-1]
        y = list(range(n))
        assert ppro.knee_point(x, y) == opro.knee_locator(x, y)
    assert ppro.knee_point([9, 7, 5, 5, 3, 3, 3, 2, 2], list(range(9))) == 3.0
    assert ppro.knee_point([4, 4, 4], [0, 1, 2]) is None


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
