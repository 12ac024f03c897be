"""`generate_summary(..., refined_contig=...)` of the REFERENCE (postprocess/collect.py:438-558): the left join of the refined
contig calls (`--refine`) into the summary table, written like `write_output` does (`%.3f`).  The per-contig inputs are the
ones of tests/golden/make_goldens.py's pred_to_dict section (same seed), the refined frame is synthetic (two contigs absent:
fewer than min_windows informative windows -> NaN cells).  Writes tests/golden/summary_refined.tsv and refined_contig.json.

usage:  python tests/golden/make_refine_summary_golden.py
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT))
import make_goldens as mg          # noqa: E402  (installs the pyfastx / parasail / ... stubs, adds the reference to sys.path)


def main():
    import pandas as pd
    from jaeger.postprocess import collect as rcollect
    rng = np.random.default_rng(5)
    n_win = [1, 2, 7, 29, 3, 140, 1, 12]
    W = sum(n_win)
    pred = (rng.normal(0, 2.0, (W, 6))).astype(np.float32)
    pred[5:9] = pred[4]
    rel = rng.normal(0, 1.5, (W, 1)).astype(np.float32)
    meta = {f"meta_{i}": [] for i in range(10)}
    for ci, n in enumerate(n_win):
        for j in range(n):
            g, c, a, t = (int(x) for x in rng.integers(300, 600, 4))
            skew = round((g - c) / (g + c), 2)
            for i, v in enumerate([f"contig___{ci}", j * 1500, int(j == n - 1), j, 2000 + 1500 * (n - 1), g, c, a, t, f"{skew: .3f}"]):
                meta[f"meta_{i}"].append(str(v).encode())
    y = {"prediction": pred, "reliability": rel, **{k: np.array(v) for k, v in meta.items()}}
    classes = ["bacteria", "phage", "eukarya", "archaea", "plasmid", "virus"]
    class_map = {"num_classes": 6, "class": classes, "index": list(range(6))}
    rep = pd.DataFrame({"contig_id": [f"contig___{i}" for i in range(len(n_win))], "terminal_repeats": [None] * len(n_win),
                        "repeat_length": [None] * len(n_win)})
    data, _ = rcollect.pred_to_dict(y, fsize=2000, class_map=class_map, term_repeats=rep)
    base = rcollect.generate_summary(data, labels=classes, indices=list(range(6)))
    want_base = pd.read_csv(OUT / "summary.tsv", sep="\t")
    assert base["contig_id"].tolist() == want_base["contig_id"].tolist()          # same inputs as make_goldens.py
    refined = pd.DataFrame({
        "contig_id": [f"contig___{i}" for i in (2, 3, 4, 5, 7, 9)],               # 0, 1, 6 absent; 9 does not exist in the summary
        "phage_score": [1.5, 20.25, 0.5, 300.0, 2.0, 1.0], "contig_call": ["phage", "bacteria_or_plasmid", "virus", "phage", "virus_any", "phage"],
        "contig_top_class": ["phage", "bacteria", "virus", "phage", "phage", "phage"], "contig_top_logit": [3.14159, 20.25, 0.5, 300.0, 2.0, 1.0],
        "contig_margin": [0.001, 0.4996, 1.0, 250.125, 0.0, 1.0], "n_windows_used": [5, 29, 3, 131, 12, 3], "n_merged_windows": [0, 11, 0, 2, 12, 0]})
    df = rcollect.generate_summary(data, labels=classes, indices=list(range(6)), refined_contig=refined)
    df.to_csv(OUT / "summary_refined.tsv", sep="\t", index=False, float_format="%.3f")
    (OUT / "refined_contig.json").write_text(json.dumps(refined.to_dict(orient="list")))
    print(df[["contig_id", "contig_call", "n_windows_used"]])


if __name__ == "__main__":
    main()
