"""Achieved HBM bandwidth of the stage 1 / 2 / 4 kernels (pack, encode, per-contig aggregation) on the bench's
64 Mbp batch: CUDA events on the engine's stream, L2 flushed between repetitions, algorithmic bytes as in DESIGN.md
section 4.  Writes one JSON object (also to profiles/ when --out is given).

    python tools/stage_kernels_bench.py [--reps 20] [--out profiles/stage_kernels_r1.json]
"""
import argparse
import json
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from bench import FSIZE, STRIDE, synth_batch
from jaeger_b200 import B200Engine, parse_project, standin_1p4m_config


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--mbp", type=float, default=64.0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    eng = B200Engine(spec=parse_project(standin_1p4m_config()))
    seq, lens = synth_batch(1, int(a.mbp * 1e6))
    peaks = {}
    try:
        peaks = json.load(open("MEASURED_PEAKS.json"))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs_burst", peaks.get("hbm_gbs", 6551.0)))
    st = eng._stream()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=eng.tdev)
    res = {}
    with torch.cuda.stream(st):
        x = torch.from_numpy(seq).pin_memory().to(eng.tdev, non_blocking=True)
        offsets = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        codes, valid = eng.pack(x)
        contig, start, nb, ordinal, last = eng.plan_windows(lens, FSIZE, STRIDE)
        w = len(contig)
        lc = eng.codons_per_frame(FSIZE, FSIZE)
        d_base, d_nb = eng._h2d(offsets[contig] + start), eng._h2d(nb)
        logits = torch.randn((w, 6), dtype=torch.float32, device=eng.tdev)
        rel = torch.randn((w, 1), dtype=torch.float32, device=eng.tdev)
        woff = eng._h2d(np.concatenate([[0], np.flatnonzero(last) + 1]).astype(np.int64))
        n_contigs = woff.numel() - 1

        def timed(fn):
            ts = []
            for _ in range(a.reps + 3):
                flush.fill_(1)                               # 512 MB > L2: the kernel reads from HBM
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st); fn(); e1.record(st)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            return float(np.median(ts[3:]))

        n = x.numel()
        cases = {
            "pack_bases": (lambda: eng.pack(x), n * 1.0 + n * 0.375),                       # 1 B in, 2 + 1 bit out per base
            "encode_windows": (lambda: eng.encode(codes, valid, d_base, d_nb, FSIZE, lc), w * (750.0 + 4010.0)),
            "aggregate_contigs": (lambda: eng.aggregate(logits, rel, woff), w * 28.0 + n_contigs * (6 * 6 + 24)),
        }
        for name, (fn, nbytes) in cases.items():
            ms = timed(fn)
            res[name] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / hbm_peak}
    out = {"batch_mbp": a.mbp, "windows": int(w), "contigs": int(n_contigs), "hbm_peak_gbs": hbm_peak, "reps": a.reps,
           "note": "pack includes two torch.zeros fills of its outputs; aggregate = 3 kernels", "kernels": res}
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
