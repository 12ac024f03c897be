"""Window-level refinement from the REFERENCE's own functions: `add_score_features` + `refine`
(postprocess/refinement.py:39-137) executed from the reference source with a minimal stand-in for the four polars calls they
make (DataFrame(dict) / .select(cols).to_numpy() / .with_columns([Series]) / df[col].to_numpy(); polars >= 1.0 is not
installable here) -- everything else in those two functions is NumPy and runs as written, including np.sort / np.argsort on
exactly tied logits.  `aggregate_contig` (polars expressions) is not covered by this stand-in.
Writes tests/golden/refine_windows.json: per window top / second class, margin, refined label for the seeded case of
tests/helpers.refine_case() under two threshold sets.

usage:  python tests/golden/make_refine_window_goldens.py
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(OUT.parent.parent))


def _arr(values):
    """polars returns String columns as object arrays from .to_numpy() (no fixed-width truncation on assignment)."""
    a = np.asarray(values)
    return a.astype(object) if a.dtype.kind in "US" else a


class Series:
    def __init__(self, name, values):
        self.name, self.values = name, _arr(values)

    def to_numpy(self):
        return self.values


class DataFrame:
    def __init__(self, data):
        self.cols = {k: _arr(v) for k, v in data.items()}

    def select(self, names):
        return DataFrame({n: self.cols[n] for n in names})

    def to_numpy(self):
        return np.stack([self.cols[k] for k in self.cols], axis=1)

    def with_columns(self, series):
        out = dict(self.cols)
        for s in (series if isinstance(series, (list, tuple)) else [series]):
            out[s.name] = s.values
        return DataFrame(out)

    def __getitem__(self, name):
        return Series(name, self.cols[name])


pl = types.ModuleType("polars")
pl.DataFrame, pl.Series, pl.Expr = DataFrame, Series, object
sys.modules["polars"] = pl

from tests.helpers import refine_case          # noqa: E402


def main():
    from jaeger.postprocess import refinement as rf
    z, offsets, headers, taus = refine_case()
    out = {}
    for tag, tt, kw in (("case", taus, {}), ("no_merge", {c: {"logit": 0.0, "margin": 0.3} for c in rf.CLASSES}, dict(merge_bp=False, merge_pv=False))):
        df = DataFrame({c: z[:, i].astype(np.float64) for i, c in enumerate(rf.SCORE_COLS)})      # float64 columns, as polars builds from row dicts
        df = rf.add_score_features(df)
        df = rf.refine(df, tt, **kw)
        out[tag] = {"top_class": df["top_class"].to_numpy().tolist(), "second_class": df["second_class"].to_numpy().tolist(),
                    "margin": df["margin"].to_numpy().tolist(), "refined_prediction": df["refined_prediction"].to_numpy().tolist()}
        print(tag, {k: out[tag]["refined_prediction"].count(k) for k in sorted(set(out[tag]["refined_prediction"]))})
    (OUT / "refine_windows.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
