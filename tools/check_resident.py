#!/usr/bin/env python
"""GPU check of the window-resident kernel (csrc/conv_resident.cuh) on the config-3 model: resident vs per-layer kernels vs the
fp32 oracle, then windows/s of both on 200 k fragments.   python tools/check_resident.py [n_fragments]"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project          # noqa: E402
from jaeger_b200.modelspec import baseline_500bp_config                                # noqa: E402
from oracle import encode as oenc                                                      # noqa: E402
from oracle import forward as ofw                                                      # noqa: E402
from tests.helpers import random_contigs                                               # noqa: E402


def engine(spec, w, resident):
    os.environ["JG_RESIDENT"] = "1" if resident else "0"
    return B200Engine(spec=spec, weights=w)


def main():
    n_frag = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    spec = parse_project(baseline_500bp_config())
    w = init_random(spec, 2)
    recs = random_contigs(2, [500] * 300 + [499, 1700, 512, 640], n_run_every=7, lower_every=0)
    seqs = [s[i:i + 500] for _, s in recs for i in range(0, len(s) - 499, 500)]
    ref = ofw.forward(spec, w, oenc.encode_windows(seqs, 500))
    out = {}
    for resident in (True, False):
        eng = engine(spec, w, resident)
        y = eng.predict(WindowSource(records=recs, fsize=500, stride=500))
        out[resident] = y
        print("resident" if resident else "per-layer", eng.conv_kernel_names(),
              "max |logit - oracle| =", float(np.abs(ref["prediction"] - y["prediction"]).max()),
              "max |emb - oracle| =", float(np.abs(ref["embedding"] - y["embedding"]).max()), flush=True)
        eng.close()
    print("resident vs per-layer: max |dlogit| =", float(np.abs(out[True]["prediction"] - out[False]["prediction"]).max()))
    # throughput
    rng = np.random.default_rng(0)
    bases = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n_frag * 500)
    offsets = np.arange(n_frag + 1, dtype=np.int64) * 500
    names = [f"f{i}" for i in range(n_frag)]
    for resident in (True, False):
        eng = engine(spec, w, resident)
        src = WindowSource.from_host(names, bases, offsets, fsize=500, stride=500)
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            y = eng.predict(src)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print("resident" if resident else "per-layer", eng.conv_kernel_names(), f"{n_frag / dt / 1e6:.2f} M windows/s end to end (predict)", flush=True)
        eng.close()


if __name__ == "__main__":
    main()
