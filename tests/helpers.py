"""Shared synthetic inputs for the parity tests (seeded, no file I/O)."""
from __future__ import annotations

import numpy as np


def random_contigs(seed: int, lengths, n_run_every: int = 3, lower_every: int = 4):
    """Contigs of i.i.d. bases; every n_run_every-th gets a run of N, every lower_every-th
    gets a lower-case stretch and a few IUPAC codes."""
    rng = np.random.default_rng(seed)
    recs = []
    for i, n in enumerate(lengths):
        s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n).copy()
        if n_run_every and i % n_run_every == 1 and n > 400:
            a = int(rng.integers(0, n - 300))
            s[a:a + int(rng.integers(20, 300))] = ord("N")
        if lower_every and i % lower_every == 2 and n > 200:
            a = int(rng.integers(0, n - 150))
            s[a:a + 120] |= 0x20
            s[int(rng.integers(0, n))] = ord("R")
            s[int(rng.integers(0, n))] = ord("y")
        recs.append((f"c{i}", s.tobytes().decode()))
    return recs


def to_dyt(cfg: dict) -> dict:
    """The same architecture with MaskedDYT in place of every MaskedBatchNorm (the reference's
    train_config/nn_config_1500bp_nmd_merge_6_class_zeus.yaml family)."""
    for layer in cfg["model"]["representation_learner"]["hidden_layers"]:
        if layer["name"] == "masked_batchnorm":
            layer["name"] = "masked_dyt"
            layer["config"] = {}
        elif layer["name"] == "residual_block":
            layer["config"]["norm_type"] = "masked_dyt"
    return cfg


_COMP = {"A": "T", "T": "A", "C": "G", "G": "C"}


def _revcomp(s: str) -> str:
    return "".join(_COMP[b] for b in reversed(s))


def repeat_contigs():
    """Contigs with planted terminal repeats: exact direct / inverted, one mismatch, a gap in either line, a long
    (LTR) repeat, N runs, soft-masked letters, a header with a comma, one record below fsize."""
    rng = np.random.default_rng(77)
    rnd = lambda n: "".join(rng.choice(list("ACGT"), n))
    core, long_core = rnd(120), rnd(700)
    mut = core[:60] + ("A" if core[60] != "A" else "C") + core[61:]
    return [("plain", rnd(3000)), ("dtr", core + rnd(3000) + core), ("itr,comma", core + rnd(2500) + _revcomp(core)),
            ("mismatch", core + rnd(2500) + mut), ("gap", core + rnd(2500) + core[:60] + core[62:]),
            ("qgap", core[:58] + core[59:] + rnd(2600) + core), ("ltr", long_core + rnd(30000) + long_core),
            ("short", rnd(900)), ("nrun", "N" * 30 + core + rnd(2200) + core + "n" * 10),
            ("lower", core.lower() + rnd(2400) + core), ("long", rnd(20) + core + rnd(48000) + core + rnd(33)),
            ("both", core + rnd(1200) + _revcomp(core[:80]) + rnd(900) + core[:100])]


def prophage_genomes():
    """(records, prophage coordinates) for the att-site search (fsize 2000, stride 1500): a 520 kbp genome with four
    called regions -- an exact 30-bp direct repeat across the region ends, a 120-bp inverted repeat with one mismatch
    (traced), a 160-bp direct repeat with a deleted base (traced, gap) around a region that is 23 % N (reject), a region at the contig start
    without a planted repeat -- a second > 500 kbp genome without regions, and a 400 kbp contig whose regions the
    500 000 bp rule skips.  Coordinates are window-index ranges + scores, as `segment` returns them."""
    rng = np.random.default_rng(123)
    rnd = lambda n: "".join(rng.choice(list("ACGT"), n))
    g = list(rnd(520_000))

    def put(pos, s):
        g[pos:pos + len(s)] = list(s)
    a30, b120, c160 = rnd(30), rnd(120), rnd(160)
    # region windows [100, 104): raw 150000 .. 156500, off_set 1625 -> left [146000, 151625), right [154875, 160500)
    put(149_000, a30); put(157_000, a30)
    # region windows [200, 221): raw 300000 .. 332000, off_set 2000 -> left [296000, 302000), right [330000, 336000)
    mut = b120[:60] + ("A" if b120[60] != "A" else "C") + b120[61:]
    put(297_500, b120); put(333_000, _revcomp(mut))
    # region windows [300, 303): raw 450000 .. 455000, off_set 1250 -> left [446000, 451250), right [453750, 459000)
    put(447_000, c160); put(457_100, c160[:70] + c160[71:]); put(451_300, "N" * 2400)
    g[448_500] = "n"
    # region windows [340, 341): raw 510000 .. 512000; a 40-bp direct repeat just OUTSIDE the raw search windows
    # ([506000, 510500) / [511500, 516000)) that only the gene-refined ends (508500 .. 512700) bring into reach
    d40 = rnd(40)
    put(505_500, d40); put(516_200, d40)
    recs = [("genome1,with,commas", "".join(g)), ("genome2", rnd(501_000)), ("plasmid", rnd(400_000))]
    cords = {"genome1___with___commas": [[[100, 104], [200, 221], [300, 303], [0, 3], [340, 341]], np.array([4.25, 7.5, 2.125, 1.75, 1.5])],
             "plasmid": [[[10, 20]], np.array([3.0])]}
    return recs, cords


def prophage_gene_calls():
    """Gene intervals (0-based half-open, as the reference's `find_genes` returns them) for `prophage_genomes()`: ends inside genes
    on either side, an intergenic end, extensions beyond 2 * fsize (capped), a gene at the contig start, overlapping genes."""
    return {"genome1___with___commas": [(0, 300), (149_800, 150_400), (150_100, 150_900), (156_200, 156_900), (299_000, 301_000),
                                        (445_000, 452_000), (454_000, 460_000), (508_500, 510_300), (511_900, 512_700)],
            "plasmid": [(14_000, 16_000)]}


def dust_contigs():
    """Contigs with low-complexity stretches (homopolymers, di- / tri-nucleotide repeats, a repeat that straddles a window
    boundary, lower-case input, N runs) between random sequence, plus a short contig for the two-pass mode."""
    rng = np.random.default_rng(64)
    rnd = lambda n: "".join(rng.choice(list("ACGT"), n))
    return [("lc1", rnd(700) + "A" * 90 + rnd(900) + "AC" * 60 + rnd(1500) + "GGT" * 40 + rnd(800)),
            ("lc2,comma", rnd(1950) + "T" * 120 + rnd(2200) + "N" * 40 + rnd(300) + "CA" * 35 + rnd(1000)),
            ("plain", rnd(2600)), ("lower", (rnd(400) + "G" * 70 + rnd(1700)).lower() + rnd(300)),
            ("short", rnd(300) + "AT" * 50 + rnd(400)), ("tiny", rnd(120))]


def refine_case(seed=11):
    """Seeded window logits of 10 contigs with exact ties, near-tied bacteria / plasmid and phage / virus stretches, plus
    per-class thresholds (one class disabled with -inf): the `--refine` test case."""
    rng = np.random.default_rng(seed)
    n_win = [1, 2, 3, 5, 40, 7, 3, 150, 4, 33]
    W = sum(n_win)
    z = rng.normal(0.0, 1.6, (W, 6)).astype(np.float32)
    z[10:14] = z[9]                                    # identical windows
    z[20, 3] = z[20, 4] = z[20].max() + 1.0            # exact tie between bacteria and plasmid at the top
    z[21, :] = 0.0                                     # all equal
    z[22, 0] = z[22, 1] = 2.5; z[22, 2:] = -1.0        # phage / virus tie
    z[60:75, 3] += 3.0; z[60:75, 4] += 2.8             # bacteria ~ plasmid: merged labels
    z[100:130, 0] += 4.0; z[100:130, 1] += 3.9         # phage ~ virus
    offsets = np.concatenate([[0], np.cumsum(n_win)]).astype(np.int64)
    headers = [f"ctg___{i}" for i in range(len(n_win))]
    classes = ["phage", "virus", "archaea", "bacteria", "plasmid", "eukarya"]
    taus = {c: {"logit": 0.4 + 0.15 * i, "margin": 0.2 + 0.1 * i, "n": 50} for i, c in enumerate(classes)}
    taus["archaea"] = {"logit": float("-inf"), "margin": float("-inf"), "n": 2}
    return z, offsets, headers, taus


def agreement_stats(got: np.ndarray, ref: np.ndarray, is_last: np.ndarray) -> dict:
    """Label agreement between device logits `got` and oracle logits `ref` [W, n_cls]:
    per window (argmax) and per contig (argmax of the float16 mean logits, postprocess/collect.py:332-342),
    plus the logit error in absolute terms and relative to the oracle's top-2 margin."""
    got, ref = np.asarray(got, np.float32), np.asarray(ref, np.float32)
    err = np.abs(got - ref).max(axis=1)
    top2 = np.sort(ref, axis=1)
    margin = top2[:, -1] - top2[:, -2]
    win_same = got.argmax(1) == ref.argmax(1)
    ends = np.flatnonzero(np.asarray(is_last).astype(bool)) + 1
    starts = np.concatenate([[0], ends[:-1]])
    cg = np.array([np.float16(got[a:b].mean(axis=0)) for a, b in zip(starts, ends)])
    cr = np.array([np.float16(ref[a:b].mean(axis=0)) for a, b in zip(starts, ends)])
    contig_same = cg.argmax(1) == cr.argmax(1)
    flips = np.flatnonzero(~win_same)
    return {"windows": int(len(got)), "contigs": int(len(ends)),
            "window_label_agreement": float(win_same.mean()), "contig_label_agreement": float(contig_same.mean()),
            "max_abs_logit_err": float(err.max()), "mean_abs_logit_err": float(err.mean()),
            "median_top2_margin": float(np.median(margin)), "mean_abs_logit": float(np.abs(ref).mean()),
            "max_err_over_margin_at_flips": float((err[flips] / np.maximum(margin[flips], 1e-9)).min()) if len(flips) else None,
            "p999_err_over_margin": float(np.quantile(err / np.maximum(margin, 1e-9), 0.999)),
            "flipped_windows": int(len(flips)), "flipped_contigs": int((~contig_same).sum()),
            "largest_margin_among_flips": float(margin[flips].max()) if len(flips) else 0.0}


def agreement_contigs(seed: int, n_contigs: int):
    """Many short contigs (1-4 windows at fsize 2000 / stride 1500) with varied composition: i.i.d. bases with a per-contig
    GC content in [0.25, 0.75], a tenth of them with a low-complexity stretch or an N run."""
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n_contigs):
        n = int(rng.integers(2000, 6600))
        gc = rng.uniform(0.25, 0.75)
        pr = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
        s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n, p=pr).copy()
        if i % 10 == 3:
            a = int(rng.integers(0, n - 200)); s[a:a + int(rng.integers(10, 150))] = ord("N")
        if i % 10 == 7:
            a = int(rng.integers(0, n - 200)); s[a:a + 90] = np.frombuffer(b"ACA", dtype=np.uint8)[np.arange(90) % 3]
        recs.append((f"k{i}", s.tobytes().decode()))
    return recs
