"""Generates the golden vectors under tests/golden/ by importing the REFERENCE's own Python
modules from /root/reference/src (build container only; the GPU box never runs this).

Third-party native modules the reference imports at module top but which are not installed
here (pyfastx, pydustmasker, parasail, ruptures, kneed, matplotlib, pycirclize, tensorflow)
are replaced by inert stubs; only reference functions that are pure Python / NumPy / numba
are executed:

  jaeger.seqops.io._window_indices, fragment_generator (dustmask=False; FASTA iteration through
      a stub with pyfastx's (name, sequence) record semantics)
  jaeger.dataops.convert._process_batch_numba   (the reference's TF-free six-frame encoder)
  jaeger.postprocess.collect.pred_to_dict, generate_summary
  jaeger.postprocess.prophages.logits_to_df_v2
  jaeger.postprocess.helpers.merge_overlapping_ranges, viterbi_decode, build_transition_costs
  jaeger.postprocess.collect.pred_to_dict_legacy, generate_summary_legacy; helpers.ood_predict_default
      (the bundled LR_ood_4_class_default.pkl through the installed scikit-learn)
  jaeger.utils.termini.get_alignment_summary (on mock alignment results; parasail stubbed)

usage:  python tests/golden/make_goldens.py
"""
from __future__ import annotations

import json
import sys
import types
import zlib
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
    return m


class _Fasta:
    """pyfastx.Fasta(path, build_index=False) iteration: (name, sequence) tuples, name = header up
    to the first whitespace, sequence = lines joined."""

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
                elif name is not None:
                    chunks.append(line.strip())
        if name is not None:
            yield name, "".join(chunks)


_stub("pyfastx", Fasta=_Fasta)
_stub("pydustmasker", DustMasker=None)
_stub("parasail")
_stub("ruptures")
_stub("kneed", KneeLocator=None)
_stub("pycirclize", Circos=None)
mpl = _stub("matplotlib")
_stub("matplotlib.pyplot")
_stub("matplotlib.patches", Patch=None)
_stub("matplotlib.lines", Line2D=None)


def viterbi_goldens(rhelp, rcollect):
    """--crf decoding: reference viterbi_decode / build_transition_costs / pred_to_dict(crf_switch_cost=...)."""
    rng = np.random.default_rng(11)
    classes = ["bacteria", "phage", "eukarya", "archaea", "plasmid", "virus"]
    n_win = [1, 2, 3, 50, 400, 7, 3000, 1, 25]
    W = sum(n_win)
    pred = rng.normal(0, 1.2, (W, 6)).astype(np.float32)
    seg = rng.integers(0, 6, W // 20 + 1)
    pred[np.arange(W), seg[np.arange(W) // 20]] += 1.5          # piecewise-constant signal under the noise
    pred[10:14] = pred[9]                                        # exact ties between consecutive windows
    out = {"prediction": pred, "n_win": np.array(n_win)}
    splits = np.cumsum(n_win)[:-1]
    user = {"Bacteria": {"phage": 0.25, "virus": 2.0}, "plasmid": {"archaea": 4.0}, "unknown": {"phage": 9.0}}
    variants = {"potts2": (2.0, None), "potts0": (0.0, None),
                "bio2": (2.0, rhelp.build_transition_costs(classes, 2.0, "biological")),
                "uni05": (0.5, rhelp.build_transition_costs(classes, 0.5, "uniform")),
                "user3": (3.0, rhelp.build_transition_costs(classes, 3.0, "biological", user))}
    for name, (lam, costs) in variants.items():
        out[f"path_{name}"] = np.concatenate([rhelp.viterbi_decode(p, lam, costs) for p in np.split(pred, splits)])
        if costs is not None:
            out[f"costs_{name}"] = costs
    out["costs_4cls"] = rhelp.build_transition_costs(["bacteria", "phage", "eukarya", "archaea"], 2.0)
    z = rng.normal(0, 2.0, (500, 1)).astype(np.float32)         # binary head: stacked [0, z] (collect.py:367-375)
    out["binary_logit"] = z
    out["binary_path"] = rhelp.viterbi_decode(np.concatenate([np.zeros_like(z), z], axis=-1), 2.0)
    # through pred_to_dict: per-class counts with CRF labels
    meta = {f"meta_{i}": [] for i in range(10)}
    for ci, n in enumerate(n_win):
        for j in range(n):
            for i, v in enumerate([f"c{ci}", j * 1500, int(j == n - 1), j, 2000 + 1500 * (n - 1), 500, 500, 500, 500, " 0.000"]):
                meta[f"meta_{i}"].append(str(v).encode())
    y = {"prediction": pred, **{k: np.array(v) for k, v in meta.items()}}
    class_map = {"num_classes": 6, "class": classes, "index": list(range(6))}
    data, _ = rcollect.pred_to_dict(y, fsize=2000, class_map=class_map, term_repeats=None, crf_switch_cost=2.0,
                                    crf_prior="biological")
    out["crf_counts"] = np.array([[d[k] for k in range(6)] for d in data["per_class_counts"]])
    out["crf_frag_pred"] = np.concatenate(data["frag_pred"])
    np.savez_compressed(OUT / "viterbi.npz", **out)


def termini_goldens():
    """get_alignment_summary (utils/termini.py:17-88) on mock parasail results: pins the coordinate
    arithmetic of the terminal-repeat table (parasail itself is not installable here)."""
    _stub("parasail")
    from jaeger.utils import termini as rterm
    rng = np.random.default_rng(31)
    cases = []
    for k in range(40):
        n = int(rng.integers(400, 4001))
        seq_len = int(n / 0.04) + int(rng.integers(0, 5000))
        cols = int(rng.integers(13, min(n, 900)))
        qg, rg = (int(x) for x in rng.integers(0, 4, 2)) if k % 3 == 0 else (0, 0)
        query = "".join(rng.choice(list("ACGT"), cols - qg)) + "-" * qg
        ref = "".join(rng.choice(list("ACGT"), cols - rg)) + "-" * rg
        iden = int(rng.integers(cols // 2, cols - qg - rg + 1))
        comp = "|" * iden + "." * (cols - iden)
        end_query = int(rng.integers(cols - qg - 1, n))
        end_ref = int(rng.integers(cols - rg - 1, n))
        tb = type("TB", (), {"query": query, "ref": ref, "comp": comp})()
        res = type("R", (), {"score": 2 * iden - 100 * (cols - iden), "end_query": end_query, "end_ref": end_ref, "saturated": False,
                             "traceback": tb})()
        for type_ in ("DTR", "ITR"):
            out = rterm.get_alignment_summary(res, seq_len=seq_len, record_id=f"r{k}", input_length=n, type_=type_)
            cases.append({"inp": {"cols": cols, "qgaps": qg, "rgaps": rg, "iden": iden, "score": res.score, "end_query": end_query,
                                  "end_ref": end_ref, "seq_len": seq_len, "n": n, "type": type_},
                          "out": {k2: out[k2] for k2 in ("repeat_length", "identities", "identity", "score", "terminal_repeats", "fgaps",
                                                        "rgaps", "sstart", "send", "estart", "eend", "seq_len")}})
    (OUT / "termini_summary.json").write_text(json.dumps(cases))


def legacy_post_goldens(rhelp, rcollect):
    """Legacy (`default` model) post-processing: ood_predict_default (sklearn variant) on the bundled
    calibrated logistic regression, pred_to_dict_legacy, generate_summary_legacy."""
    import warnings
    import joblib
    warnings.filterwarnings("ignore")
    mdir = REF / "jaeger" / "data" / "models" / "default"
    model = joblib.load(mdir / "LR_ood_4_class_default.pkl")
    cc = model.calibrated_classifiers_[0]
    ood_fix = dict(coef=cc.estimator.coef_.astype(np.float64).ravel(), intercept=np.float64(cc.estimator.intercept_[0]),
                   cal_a=np.float64(cc.calibrators[0].a_), cal_b=np.float64(cc.calibrators[0].b_),
                   batch_mean=np.load(mdir / "batch_means.npy"), batch_std=np.load(mdir / "batch_std.npy"))
    params = {"type": "sklearn", "model": model, "batch_mean": ood_fix["batch_mean"], "batch_std": ood_fix["batch_std"]}
    rng = np.random.default_rng(23)
    n_win = [1, 2, 9, 40, 3, 135, 1, 6]
    W = sum(n_win)
    out = rng.normal(0, 2.5, (W, 4)).astype(np.float32)
    out[20:24] = out[19]
    emb = (rng.normal(0, 1.0, (W, 128)) * rng.uniform(0.2, 3.0, (W, 1)) + ood_fix["batch_mean"]).astype(np.float32)
    meta = [[] for _ in range(10)]
    for ci, n in enumerate(n_win):
        for j in range(n):
            g, c, a, t = (int(x) for x in rng.integers(300, 600, 4))
            skew = round((g - c) / (g + c), 2)
            for i, v in enumerate([f"contig___{ci}", j * 1500, int(j == n - 1), j, 2000 + 1500 * (n - 1), g, c, a, t, f"{skew: .3f}"]):
                meta[i].append(str(v).encode())
    meta = tuple(np.array(m) for m in meta)
    y = {"y_hat": {"output": out, "embedding": emb}, "meta": meta}
    config = json.loads((REF / "jaeger" / "data" / "config.json").read_text())["default"]
    config["model"] = "default"
    import pandas as pd
    rep = pd.DataFrame({"contig_id": [f"contig___{i}" for i in range(len(n_win))], "terminal_repeats": [None] * len(n_win),
                        "repeat_length": [None] * len(n_win)})
    res = {}
    for tag, labels in (("default", "default_labels"), ("all", "all_labels")):
        config["labels"] = [v for k, v in config[labels].items()]
        data, _ = rcollect.pred_to_dict_legacy(config, y, model="default", fsize=2000, ood_params=params, term_repeats=rep)
        df = rcollect.generate_summary_legacy(config, data)
        df.to_csv(OUT / f"summary_legacy_{tag}.tsv", sep="\t", index=False, float_format="%.3f")
        res = data
    np.savez_compressed(OUT / "legacy_post.npz", output=out, embedding=emb, **{f"meta_{i}": m for i, m in enumerate(meta)},
                        ood_windows=np.concatenate(res["ood"]), pred_sum=res["pred_sum"], pred_var=res["pred_var"],
                        consensus=res["consensus"], entropy=res["entropy"],
                        **{f"ood_{k}": v for k, v in ood_fix.items()})


def main():
    import pandas as pd
    from jaeger.seqops import io as rio
    from jaeger.dataops import convert as rconv
    from jaeger.postprocess import collect as rcollect
    from jaeger.postprocess import helpers as rhelp
    from jaeger.postprocess import prophages as rpro

    # ---- window indices ---------------------------------------------------------------------
    cases = []
    rng = np.random.default_rng(0)
    for _ in range(300):
        fs = int(rng.choice([500, 1500, 2000, 2048]))
        L = int(rng.integers(fs, fs * 14))
        st = int(rng.choice([fs, fs // 2, 1500, 700]))
        dyn = bool(rng.integers(0, 2))
        thr = float(rng.choice([10.0, 3.0, 1.5]))
        cases.append(dict(seqlen=L, fsize=fs, stride=st, dyn=dyn, thr=thr,
                          idx=rio._window_indices(L, fs, st, dyn, thr)))
    # the reference tests' own known answers (tests/unit/test_seqops_io.py)
    for L, fs, st, dyn, thr in [(3400, 2000, 2000, False, 10.0), (3400, 2000, 2000, True, 10.0),
                                (3999, 2000, 2000, True, 10.0), (6000, 2000, 2000, True, 10.0)]:
        cases.append(dict(seqlen=L, fsize=fs, stride=st, dyn=dyn, thr=thr, idx=rio._window_indices(L, fs, st, dyn, thr)))
    (OUT / "window_indices.json").write_text(json.dumps(cases))

    # ---- a synthetic FASTA of our own (travels to the GPU box; the reference's test FASTA does not)
    rng = np.random.default_rng(11)
    syn = OUT / "synthetic_contigs.fasta"
    with open(syn, "w") as fh:
        for i, L in enumerate([9275, 12001, 44776, 2000, 1999, 650, 20480, 3400, 15000, 137, 2048, 5000]):
            seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)].copy()
            if i == 2:
                seq[3000:3100] = ord("N")
            if i == 6:
                seq[100:700] |= 0x20                       # lower-case stretch
                seq[[5, 900, 4000]] = [ord("R"), ord("y"), ord("K")]
            name = f"syn{i}" + (",with,commas" if i == 7 else "") + " description text"
            fh.write(f">{name}\n")
            s_ = seq.tobytes().decode()
            for a in range(0, L, 80):
                fh.write(s_[a:a + 80] + "\n")

    def frag_rows(path, fs, st, min_len, max_len=None, dyn=False):
        rows = []
        for s in rio.fragment_generator(str(path), fragsize=fs, stride=st, dustmask=False, min_len=min_len,
                                        max_len=max_len, dynamic_stride=dyn):
            f = s.split(",")
            rows.append([zlib.crc32(f[0].encode()), len(f[0])] + f[1:])
        return rows

    fragsyn = {}
    for fs, st, min_len, max_len, dyn in [(2000, 1500, None, None, False), (2048, 2048, None, None, False),
                                          (2000, 1500, 500, 1999, False), (500, 500, None, None, False),
                                          (2000, 2000, None, None, True)]:
        fragsyn[f"{fs}_{st}_{min_len}_{max_len}_{int(dyn)}"] = frag_rows(syn, fs, st, min_len, max_len, dyn)
    (OUT / "fragments_synthetic.json").write_text(json.dumps(fragsyn))

    # ---- fragment_generator on the reference's health FASTA ------------------------------------
    fasta = REF / "jaeger" / "data" / "test" / "test_contigs.fasta"
    frag = {}
    for fs, st, min_len in [(2000, 1500, None), (2048, 2048, None), (2000, 1500, 500), (12000, 4000, 9000)]:
        rows = []
        for s in rio.fragment_generator(str(fasta), fragsize=fs, stride=st, dustmask=False, min_len=min_len,
                                        max_len=None):
            f = s.split(",")
            rows.append([zlib.crc32(f[0].encode()), len(f[0])] + f[1:])
        frag[f"{fs}_{st}_{min_len}"] = rows
    (OUT / "fragments_test_contigs.json").write_text(json.dumps(frag))

    # ---- six-frame tokens from the reference's numba encoder ------------------------------------
    recs = list(_Fasta(str(syn)))
    seqs = []
    for name, seq in recs:
        seq = seq.upper()
        for st in (0, 1500, 3001):
            if st + 2000 <= len(seq):
                seqs.append(seq[st:st + 2000])
    # add ambiguity codes the way real assemblies carry them
    s = list(seqs[0]); s[100:140] = "N" * 40; s[777] = "R"; seqs.append("".join(s))
    codon_lut, ascii_lut, comp_lut = rconv._build_numba_lookups()
    arr = np.zeros((len(seqs), 2000), dtype=np.uint8)
    for i, q in enumerate(seqs):
        arr[i] = np.frombuffer(q.encode(), dtype=np.uint8)
    tok = rconv._process_batch_numba(arr, np.full(len(seqs), 2000, np.int64), 2000, 665, codon_lut, comp_lut, ascii_lut)
    np.savez_compressed(OUT / "tokens_2000.npz", seqs=np.array(seqs), tokens=tok.astype(np.uint8))

    # ---- pred_to_dict / generate_summary ---------------------------------------------------------
    rng = np.random.default_rng(5)
    n_win = [1, 2, 7, 29, 3, 140, 1, 12]
    W = sum(n_win)
    pred = (rng.normal(0, 2.0, (W, 6))).astype(np.float32)
    pred[5:9] = pred[4]                                   # exact ties between windows
    rel = rng.normal(0, 1.5, (W, 1)).astype(np.float32)
    meta = {f"meta_{i}": [] for i in range(10)}
    for ci, n in enumerate(n_win):
        for j in range(n):
            g, c, a, t = (int(x) for x in rng.integers(300, 600, 4))
            skew = round((g - c) / (g + c), 2)
            for i, v in enumerate([f"contig___{ci}", j * 1500, int(j == n - 1), j, 2000 + 1500 * (n - 1), g, c, a, t,
                                   f"{skew: .3f}"]):
                meta[f"meta_{i}"].append(str(v).encode())
    y = {"prediction": pred, "reliability": rel, **{k: np.array(v) for k, v in meta.items()}}
    classes = ["bacteria", "phage", "eukarya", "archaea", "plasmid", "virus"]
    class_map = {"num_classes": 6, "class": classes, "index": list(range(6))}
    rep = pd.DataFrame({"contig_id": [f"contig___{i}" for i in range(len(n_win))],
                        "terminal_repeats": [None] * len(n_win), "repeat_length": [None] * len(n_win)})
    data, data_full = rcollect.pred_to_dict(y, fsize=2000, class_map=class_map, term_repeats=rep)
    df = rcollect.generate_summary(data, labels=classes, indices=list(range(6)))
    np.savez_compressed(
        OUT / "pred_to_dict.npz", prediction=pred, reliability=rel,
        **{k: np.array(v) for k, v in meta.items()},
        pred_sum=data["pred_sum"], pred_var=data["pred_var"], consensus=data["consensus"],
        per_class_counts=np.array([[d[k] for k in range(6)] for d in data["per_class_counts"]]),
        ood=data["ood"], entropy=data["entropy"], energy=data["energy"], host_contam=data["host_contam"],
        prophage_contam=data["prophage_contam"], frag_pred=np.concatenate(data["frag_pred"]),
        gc_mean=np.array([np.mean(x) for x in data["gc"]]), ns_mean=np.array([np.mean(x) for x in data["ns"]]))
    df.to_csv(OUT / "summary.tsv", sep="\t", index=False, float_format="%.3f")

    # ---- prophage score smoothing ------------------------------------------------------------------
    rng = np.random.default_rng(9)
    T = 400
    logits = rng.normal(0, 1.5, (T, 6)).astype(np.float32)
    logits[120:160, 1] += 6.0
    gsk = rng.normal(0, 0.1, T).round(2)
    out = rpro.logits_to_df_v2(class_map, {"lc": 1000, "stride": 1500, "fsize": 2000}, np.array(["g1"]), [logits],
                               np.array([1500 * (T - 1) + 2000]), [gsk.copy()], [np.full(T, 0.5)])
    t = out["g1"][0]
    np.savez_compressed(OUT / "smooth.npz", logits=logits, gc_skew_in=gsk, smoothed=t[classes].to_numpy(),
                        x=t["length"].to_numpy(), gc_skew=t["gc_skew"].to_numpy())
    merges = []
    for arr_ in ([[0, 3], [2, 5], [8, 9]], [[1, 2]], [[0, 10], [2, 3], [11, 12], [12, 20]]):
        merges.append(dict(inp=arr_, out=[list(map(int, r)) for r in rhelp.merge_overlapping_ranges(np.array(arr_))]))
    (OUT / "merge_ranges.json").write_text(json.dumps(merges))
    viterbi_goldens(rhelp, rcollect)
    legacy_post_goldens(rhelp, rcollect)
    termini_goldens()
    if "--only-post" in sys.argv:
        return
    # ---- BASELINE config 1 fixture: the bundled legacy `default` weights + the health FASTA --------
    from jaeger_b200 import legacy
    from jaeger_b200.weights import read_tf_bundle, _flatten
    w = legacy.weights_from_bundle(read_tf_bundle(REF / "jaeger" / "data" / "models" / "test" / "jaeger_fragment_graph" / "variables"))
    flat = {}
    _flatten(w, "", flat)
    health = list(_Fasta(str(fasta)))
    np.savez_compressed(OUT / "legacy_default.npz", names=np.array([n for n, _ in health]),
                        seqs=np.array([q for _, q in health]), **{f"w/{k}": v for k, v in flat.items()})
    print("goldens written to", OUT)


if __name__ == "__main__":
    main()
