"""Goldens from the reference's own serialized TensorFlow graph (no TensorFlow needed).

`/root/reference/src/jaeger/data/models/test/jaeger_fragment_graph/saved_model.pb` is the legacy `default` model as
Keras exported it; oracle/tfgraph.py parses its `serving_default` FunctionDef and interprets it op by op in float64 with
the variables of the bundle next to it.  Written: tests/golden/legacy_graph_outputs.npz

  health_output [135, 4], health_embedding [135, 128]   the 135 windows of the reference's health FASTA at the CLI
                                                        defaults (BASELINE config 1), tokens = the legacy encoder's
                                                        (tests/golden/legacy_default.npz holds the records), health_token_crc
  random_output [6, 4], random_embedding [6, 128]       tokens from numpy default_rng(0) with unknown runs / padding

usage:  python tests/golden/make_legacy_graph_goldens.py
"""
from __future__ import annotations

import sys
import zlib
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT.parent.parent))
MODEL = Path("/root/reference/src/jaeger/data/models/test/jaeger_fragment_graph")

from jaeger_b200 import codon_tables as ct          # noqa: E402  (LUT constants and the bundle reader only)
from jaeger_b200.weights import read_tf_bundle       # noqa: E402
from oracle import encode as oenc                   # noqa: E402
from oracle import seqwin                           # noqa: E402
from oracle.tfgraph import SavedFunction            # noqa: E402


def health_tokens() -> np.ndarray:
    z = np.load(OUT / "legacy_default.npz")
    recs = [(str(n), str(s)) for n, s in zip(z["names"], z["seqs"])]
    wins = list(seqwin.fragment_windows(recs, 2000, 1500))
    table = dict(zip(oenc.CODONS, ct.LEGACY_AA_ID))
    return np.stack([oenc.encode_window_legacy(x.seq, 2000, table) for x in wins]).astype(np.uint8)


def random_tokens() -> np.ndarray:
    rng = np.random.default_rng(0)
    tok = rng.integers(1, 22, size=(6, 6, 665)).astype(np.uint8)
    tok[0, :, 100:180] = 0            # unknown run
    tok[1, :, 300:] = 0               # right padding of a short contig
    tok[2][rng.random((6, 665)) < 0.1] = 0
    tok[3] = 0                        # nothing but unknown codons
    return tok


def main():
    fn = SavedFunction(MODEL, read_tf_bundle(MODEL / "variables"))
    out = {}
    for tag, tok in (("health", health_tokens()), ("random", random_tokens())):
        res = {k: [] for k in fn.rets}
        for a in range(0, len(tok), 32):
            r = fn.run([tok[a:a + 32, i].astype(np.float32) for i in range(6)])
            for k in res:
                res[k].append(r[k])
        logits = np.concatenate([x for k, x in res.items() if x[0].shape[1] == 4][0])
        emb = np.concatenate([x for k, x in res.items() if x[0].shape[1] == 128][0])
        out[f"{tag}_output"], out[f"{tag}_embedding"] = logits.astype(np.float32), emb.astype(np.float32)
        out[f"{tag}_token_crc"] = np.array(zlib.crc32(tok.tobytes()), dtype=np.int64)
        print(tag, tok.shape, "->", logits.shape, emb.shape)
    np.savez_compressed(OUT / "legacy_graph_outputs.npz", **out)
    print("written:", OUT / "legacy_graph_outputs.npz")


if __name__ == "__main__":
    main()
