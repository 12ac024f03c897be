"""Window-level refinement from the REFERENCE's own functions: `add_score_features` + `refine`
(postprocess/refinement.py:39-137) executed from the reference source with a minimal stand-in for the polars calls they
make (DataFrame / .select(cols).to_numpy() / .with_columns / df[col].to_numpy(); polars >= 1.0 is not installable here) --
everything else in those two functions is NumPy and runs as written, including np.sort / np.argsort on exactly tied logits --
and `aggregate_contig` (refinement.py:140-247) behind a small evaluator for the lazy expressions it builds (col / lit / when-
then-otherwise / is_in / clip / sum / len, filter, group_by.agg): the evaluator is this file's reading of polars, the
expressions, filters, weights, multipliers and the contig-level selection are the reference's own code.
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


class Expr:
    """The slice of polars' lazy expressions aggregate_contig uses: evaluated against a dict of column arrays."""

    def __init__(self, fn, name=None, agg=False):
        self.fn, self.name, self.agg = fn, name, agg

    def eval(self, cols):
        return self.fn(cols)

    def _bin(self, other, op):
        o = other if isinstance(other, Expr) else Expr(lambda c, v=other: v)
        return Expr(lambda c: op(self.eval(c), o.eval(c)), self.name)

    def __mul__(self, other):
        return self._bin(other, lambda a, b: a * b)

    def __ne__(self, other):
        return self._bin(other, lambda a, b: a != b)

    def __ge__(self, other):
        return self._bin(other, lambda a, b: a >= b)

    def clip(self, lo, hi):
        return Expr(lambda c: np.clip(self.eval(c), lo, hi), self.name)

    def is_in(self, values):
        return Expr(lambda c: np.isin(self.eval(c), list(values)), self.name)

    def sum(self):
        return Expr(lambda c: np.sum(self.eval(c)), self.name, agg=True)

    def alias(self, name):
        return Expr(self.fn, name, self.agg)


class _When:
    def __init__(self, cond):
        self.cond = cond

    def then(self, a):
        self.a = a
        return self

    def otherwise(self, b):
        return Expr(lambda c: np.where(self.cond.eval(c), self.a.eval(c), b.eval(c)))


def _n_rows(cols):
    return len(next(iter(cols.values())))


class _GroupBy:
    def __init__(self, df, key):
        self.df, self.key = df, key

    def agg(self, exprs):
        keys = self.df.cols[self.key]
        out = {self.key: []}
        for k in dict.fromkeys(keys.tolist()):
            sel = keys == k
            sub = {n: v[sel] for n, v in self.df.cols.items()}
            out[self.key].append(k)
            for e in exprs:
                out.setdefault(e.name, []).append(e.eval(sub))
        return DataFrame(out)


class DataFrame:
    def __init__(self, data):
        if isinstance(data, list):                       # list of row dicts
            data = {k: [r[k] for r in data] for k in data[0]}
        self.cols = {k: _arr(v) for k, v in data.items()}

    def __len__(self):
        return _n_rows(self.cols) if self.cols else 0

    def select(self, names):
        return DataFrame({n: self.cols[n] for n in names})

    def to_numpy(self):
        return np.stack([self.cols[k].astype(np.float64) for k in self.cols], axis=1) if len(self) else np.zeros((0, len(self.cols)))

    def with_columns(self, items):
        out = dict(self.cols)
        for s in (items if isinstance(items, (list, tuple)) else [items]):
            if isinstance(s, Series):
                out[s.name] = s.values
            else:
                v = s.eval(self.cols)
                out[s.name] = np.full(len(self), v) if np.ndim(v) == 0 else v
        return DataFrame(out)

    def filter(self, expr):
        m = np.asarray(expr.eval(self.cols), dtype=bool)
        return DataFrame({k: v[m] for k, v in self.cols.items()})

    def group_by(self, key):
        return _GroupBy(self, key)

    def __getitem__(self, name):
        return Series(name, self.cols[name])


pl = types.ModuleType("polars")
pl.DataFrame, pl.Series, pl.Expr = DataFrame, Series, Expr
pl.col = lambda name: Expr(lambda c, n=name: c[n], name)
pl.lit = lambda v: Expr(lambda c, v=v: v)
pl.when = lambda cond: _When(cond)
pl.len = lambda: Expr(lambda c: _n_rows(c), "len", agg=True)
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
        if tag == "case":
            ids = np.repeat(np.array(headers, dtype=object), np.diff(offsets))
            wdf = df.with_columns([Series("contig_id", ids)])
            out["contigs"] = {}
            for mode, split, allow in (("gated", "half", False), ("weighted", "full", True), ("unweighted", "half", True)):
                c = rf.aggregate_contig(wdf, mode=mode, min_windows=3, merge_split=split, allow_merged_contig_call=allow, contig_hedge_margin=5.0)
                rows = {}
                for i, cid in enumerate(c["contig_id"].to_numpy().tolist()):
                    rows[cid] = {k: (c[k].to_numpy()[i].item() if hasattr(c[k].to_numpy()[i], "item") else c[k].to_numpy()[i]) for k in c.cols if k != "contig_id"}
                out["contigs"][f"{mode}_{split}_{int(allow)}"] = rows
                print(mode, [(k, v["contig_call"], v["n_windows_used"]) for k, v in rows.items()])
    (OUT / "refine_windows.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
