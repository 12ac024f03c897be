"""cProfile of the end-to-end call on config 3 (1 M fragments of 500 bp from a pinned host buffer): where does the host spend the
time that the device-resident number does not have?  (development aid; run on the GPU box)"""
import cProfile, pstats, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project
from jaeger_b200.modelspec import baseline_500bp_config
from jaeger_b200.postprocess import contig_table

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
spec = parse_project(baseline_500bp_config())
eng = B200Engine(spec=spec, weights=init_random(spec, 0), workspace_gb=24)
rng = np.random.default_rng(0)
bases = torch.from_numpy(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n * 500)).pin_memory()
offsets = np.arange(n + 1, dtype=np.int64) * 500
names = [f"f{i}" for i in range(n)]


def step():
    src = WindowSource.from_host(names, bases, offsets, fsize=500, stride=500, outputs=("prediction", "reliability"), lazy_meta=True)
    y = eng.predict(src)
    return contig_table(eng, y, 500)


for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); torch.cuda.synchronize(); print(f"step: {time.perf_counter() - t0:.3f} s for {n} windows", flush=True)
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
