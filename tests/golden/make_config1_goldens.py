"""BASELINE config 1 end to end with the reference's own code: `jaeger predict -m default --no-dustmask` on the health FASTA
(src/jaeger/data/test/test_contigs.fasta, CLI defaults fsize 2000 / stride 1500), assembled from

  * the reference's `fragment_generator` (seqops/io.py; pyfastx replaced by a line reader stub),
  * the legacy six-frame amino-acid encoder (preprocess/v1/convert.py is TF string ops: restated in oracle/encode.py),
  * the reference's serialized TensorFlow graph data/models/test/jaeger_fragment_graph (= the `default` model's weights),
    interpreted op by op by oracle/tfgraph.py,
  * the reference's `pred_to_dict_legacy` + `generate_summary_legacy` (postprocess/collect.py) with the bundled
    calibrated logistic regression (data/models/default/LR_ood_4_class_default.pkl) through `ood_predict_default`.

Writes tests/golden/config1_default_jaeger.tsv and config1_all_labels_jaeger.tsv -- the `<base>_jaeger.tsv` the reference
writes for this input with default / --getalllabels labels (terminal-repeat columns empty: parasail is not installable).

usage:  python tests/golden/make_config1_goldens.py
"""
from __future__ import annotations

import json
import sys
import types
import warnings
from pathlib import Path

import numpy as np

REF = Path("/root/reference/src")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(REF))
sys.path.insert(0, str(OUT.parent.parent))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


class _Fasta:
    def __init__(self, path, build_index=False):
        self.path = path

    def __iter__(self):
        name, chunks = None, []
        with open(self.path) as fh:
            for line in fh:
                if line.startswith(">"):
                    if name is not None:
                        yield name, "".join(chunks)
                    name, chunks = line[1:].split()[0], []
                else:
                    chunks.append(line.strip())
        if name is not None:
            yield name, "".join(chunks)


for mod in ("parasail", "ruptures", "pycirclize", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.lines"):
    _stub(mod, Circos=None, Patch=None, Line2D=None)
_stub("pyfastx", Fasta=_Fasta)
_stub("pydustmasker", DustMasker=None)
_stub("kneed", KneeLocator=None)


def main():
    import joblib
    import pandas as pd
    from jaeger.postprocess import collect as rcollect
    from jaeger.seqops import io as rio
    from jaeger_b200 import codon_tables as ct
    from jaeger_b200.weights import read_tf_bundle
    from oracle import encode as oenc
    from oracle.tfgraph import SavedFunction
    warnings.filterwarnings("ignore")
    fasta = REF / "jaeger" / "data" / "test" / "test_contigs.fasta"
    rows = [s.split(",") for s in rio.fragment_generator(str(fasta), fragsize=2000, stride=1500, dustmask=False, min_len=None, max_len=None)]
    assert len(rows) == 135
    table = dict(zip(oenc.CODONS, ct.LEGACY_AA_ID))
    tok = np.stack([oenc.encode_window_legacy(r[0], 2000, table) for r in rows]).astype(np.uint8)
    gdir = REF / "jaeger" / "data" / "models" / "test" / "jaeger_fragment_graph"
    fn = SavedFunction(gdir, read_tf_bundle(gdir / "variables"))
    outs = {"output": [], "embedding": []}
    for a in range(0, len(tok), 32):
        r = fn.run([tok[a:a + 32, i].astype(np.float32) for i in range(6)])
        for v in r.values():
            outs["output" if v.shape[1] == 4 else "embedding"].append(v.astype(np.float32))
    y = {"y_hat": {k: np.concatenate(v) for k, v in outs.items()},
         "meta": tuple(np.array([r[i + 1].encode() for r in rows]) for i in range(10))}
    mdir = REF / "jaeger" / "data" / "models" / "default"
    params = {"type": "sklearn", "model": joblib.load(mdir / "LR_ood_4_class_default.pkl"), "batch_mean": np.load(mdir / "batch_means.npy"),
              "batch_std": np.load(mdir / "batch_std.npy")}
    config = json.loads((REF / "jaeger" / "data" / "config.json").read_text())["default"]
    config["model"] = "default"
    headers = list(dict.fromkeys(r[1] for r in rows))
    rep = pd.DataFrame({"contig_id": headers, "terminal_repeats": [None] * len(headers), "repeat_length": [None] * len(headers)})
    for tag, labels in (("default", "default_labels"), ("all_labels", "all_labels")):
        config["labels"] = [v for k, v in config[labels].items()]
        data, _ = rcollect.pred_to_dict_legacy(config, y, model="default", fsize=2000, ood_params=params, term_repeats=rep)
        df = rcollect.generate_summary_legacy(config, data)
        df.to_csv(OUT / f"config1_{tag}_jaeger.tsv", sep="\t", index=False, float_format="%.3f")
        print(tag, len(df), "contigs:", df["prediction"].tolist())


if __name__ == "__main__":
    main()
