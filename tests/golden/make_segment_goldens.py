"""Goldens for prophage region calling from the REFERENCE's own `logits_to_df_v2` + `segment`
(postprocess/prophages.py:99-153, 524-602).  The two third-party calls inside `segment` -- ruptures.KernelCPD(kernel="linear",
min_size=3, jump=1).fit(x).predict(pen) and kneed.KneeLocator(...).knee, neither installable here -- are replaced by stubs that
return the oracle's restatements (oracle/prophage.py: optimal_partition; knee_locator, which makes the same SciPy calls kneed
makes -- interp1d(x, y)(x) and argrelextrema -- so tied breakpoint counts behave as in the library), so the file pins everything the
reference does AROUND them: which penalties' breakpoint lists are kept, the knee / searchsorted index choice (a knee of 0 is
falsy), ranges from consecutive breakpoints, end-inclusive `.loc[s:e]` means, the sensitivity filter, the unsorted interval
merge, the length cutoff and the exception path.  ruptures / kneed themselves stay unpinned.
Writes tests/golden/segment_cases.json (inputs are seeded: only seeds, shapes and results are stored).

usage:  python tests/golden/make_segment_goldens.py
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference/src")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(REF))
sys.path.insert(0, str(OUT.parent.parent))

from oracle import prophage as opro          # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


class _KernelCPD:
    def __init__(self, kernel="linear", min_size=3, jump=1):
        assert kernel == "linear" and jump == 1
        self.min_size = min_size

    def fit(self, signal):
        self.signal = np.asarray(signal, dtype=np.float64)
        return self

    def predict(self, pen):
        return opro.optimal_partition(self.signal, float(pen), self.min_size)


class _KneeLocator:
    def __init__(self, x, y, curve="convex", direction="decreasing"):
        assert curve == "convex" and direction == "decreasing"
        self.knee = opro.knee_locator(x, y)


for mod in ("parasail", "pycirclize", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.lines", "pyfastx", "pydustmasker"):
    _stub(mod, Circos=None, Patch=None, Line2D=None, Fasta=None, DustMasker=None)
_stub("ruptures", KernelCPD=_KernelCPD)
_stub("kneed", KneeLocator=_KneeLocator)

CLASSES = ["bacteria", "phage", "eukarya", "archaea", "plasmid", "virus"]


def case_logits(seed: int, t: int, islands) -> np.ndarray:
    """The seeded window logits of a case (shared with tests/test_postprocess_cpu.py)."""
    rng = np.random.default_rng(seed)
    z = rng.normal(0.0, 1.2, (t, 6)).astype(np.float32)
    z[:, 0] += 2.0
    for a, b, lift in islands:
        z[a:b, 1] += lift
    return z


CASES = [dict(seed=1, t=400, islands=[[120, 160, 7.0]], sens=1.5), dict(seed=2, t=700, islands=[[50, 90, 6.0], [300, 420, 8.0], [600, 640, 5.0]], sens=1.5),
         dict(seed=3, t=350, islands=[], sens=1.5), dict(seed=4, t=500, islands=[[0, 30, 9.0], [470, 500, 9.0]], sens=1.5),
         dict(seed=5, t=900, islands=[[100, 130, 4.0], [135, 170, 4.5], [500, 505, 9.0]], sens=0.5),
         dict(seed=6, t=360, islands=[[10, 350, 6.0]], sens=3.0), dict(seed=7, t=340, islands=[[150, 190, 7.0]], sens=1.5, short=True),
         # breakpoint counts with long tie groups at different positions (the interp1d step of kneed decides the penalty)
         dict(seed=8, t=800, islands=[[60, 100, 5.0], [200, 230, 3.0], [400, 460, 6.5], [700, 720, 3.5]], sens=1.5),
         dict(seed=9, t=1200, islands=[[100, 140, 3.0], [300, 330, 2.5], [500, 620, 7.0], [800, 830, 2.8], [1000, 1100, 4.0]], sens=1.0),
         dict(seed=10, t=600, islands=[[50, 70, 2.5], [120, 150, 2.7], [200, 240, 2.9], [300, 350, 3.1], [450, 520, 3.3]], sens=1.5),
         dict(seed=11, t=3333, islands=[[500, 530, 6.0], [1500, 1540, 7.0], [2500, 2600, 5.0]], sens=1.5),
         dict(seed=12, t=450, islands=[[100, 112, 4.0], [200, 206, 5.0], [300, 303, 9.0]], sens=2.0)]


def main():
    from jaeger.postprocess import prophages as rpro
    class_map = {"num_classes": 6, "class": CLASSES, "index": list(range(6))}
    out = []
    for c in CASES:
        z = case_logits(c["seed"], c["t"], c["islands"])
        length = 1500 * (c["t"] - 1) + 2000
        lc = length + 1 if c.get("short") else 1000                       # `short`: the contig is not longer than the cutoff
        rng = np.random.default_rng(c["seed"] + 100)
        df = rpro.logits_to_df_v2(class_map, {"lc": 1000, "stride": 1500, "fsize": 2000}, np.array(["g"]), [z], np.array([length]),
                                  [rng.normal(0, 0.1, c["t"]).round(2)], [np.full(c["t"], 0.5)])
        res = rpro.segment(df, outdir=None, cutoff_length=lc, sensitivity=c["sens"], identifier="phage")
        ranges, scores = res.get("g", [[], []]) if res else [[], []]
        col = opro.smooth_scores(z)[:, 1]
        counts = [len(b) for b in (opro.optimal_partition(col, float(p)) for p in range(1, 10)) if len(b) > 1]
        out.append(dict(c, counts=counts, ranges=[[int(a), int(b)] for a, b in np.asarray(ranges).reshape(-1, 2)], scores=[float(s) for s in scores],
                        skipped="g" not in res))
        print(c["seed"], counts, out[-1]["ranges"], [round(s, 3) for s in out[-1]["scores"]], out[-1]["skipped"])
    (OUT / "segment_cases.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
