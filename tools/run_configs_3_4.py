"""BASELINE configs 3 and 4 timed on the device (development aid; numbers quoted in DESIGN.md).
config 3: the 500 bp / 32-filter baseline architecture on N x 500 bp fragments (window-count-bound).
config 4: prophage mode on 8 synthetic 5 Mbp genomes (sliding-window scores + region calling)."""
import sys, time, tempfile
sys.path.insert(0, ".")
from pathlib import Path
import numpy as np, torch
from jaeger_b200 import B200Engine, init_random, parse_project

def conv(f, k):
    return {"name": "masked_conv1d", "config": {"filters": f, "kernel_size": k, "strides": 1, "dilation_rate": 1, "use_bias": True, "activation": None}}
bn_act = [{"name": "masked_batchnorm", "config": {"return_nmd": False}}, {"name": "activation", "config": {"activation": "gelu"}}]
cfg = {"model": {"name": "jaeger_500bp_baseline", "activation": "gelu",
                 "class_label_map": [{"class": c, "label": i} for i, c in enumerate(["chromosome", "virus", "plasmid"])],
                 "embedding": {"use_embedding_layer": True, "input_type": "translated", "input_shape": [6, None], "embedding_size": 64},
                 "string_processor": {"seq_onehot": False, "codon": "CODON", "codon_id": "CODON_ID", "crop_size": 500, "masking": False},
                 "representation_learner": {"hidden_layers": [conv(32, 7)] + bn_act + [
                     {"name": "residual_block", "config": {"use_1x1conv": False, "block_size": 2, "filters": 32, "kernel_size": 3, "use_bias": True}},
                     {"name": "residual_block", "config": {"use_1x1conv": False, "block_size": 2, "filters": 32, "kernel_size": 3, "use_bias": True}}] + bn_act,
                     "pooling": "average"},
                 "classifier": {"input_shape": 32, "hidden_layers": [{"name": "dense", "config": {"units": 3, "activation": None, "use_bias": True}}]}}}
spec = parse_project(cfg)
eng = B200Engine(spec=spec, weights=init_random(spec, 2), workspace_gb=16)
n_frag = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(2)
seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n_frag * 500)]
lens = np.full(n_frag, 500, np.int64)
with torch.cuda.stream(eng._stream()):
    x = torch.from_numpy(seq).to(eng.tdev)
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        agg, w, c = eng.classify_long(x, lens, 500, 500)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"config 3: {w} windows of 500 bp in {dt*1e3:.0f} ms -> {w/dt/1e6:.2f} M windows/s, {w*500/dt/1e6:.0f} Mbp/s", flush=True)
eng.close()

from jaeger_b200.predict import run_core
tmp = Path(tempfile.mkdtemp())
fa = tmp / "genomes.fna"
rng = np.random.default_rng(3)
with open(fa, "wb") as fh:
    for g in range(8):
        p = np.full(5_000_000, 0.5)
        for isl in range(3):
            a = int(rng.integers(200_000, 4_700_000)); p[a:a + int(rng.integers(30_000, 50_000))] = 0.35
        gc = rng.random(5_000_000) < p
        hi = rng.random(5_000_000) < 0.5
        s = np.where(gc, np.where(hi, ord("G"), ord("C")), np.where(hi, ord("A"), ord("T"))).astype(np.uint8)
        fh.write(b">genome%d\n" % g); fh.write(s.tobytes()); fh.write(b"\n")
for rep in range(2):
    t = time.time()
    res = run_core(input=str(fa), output=str(tmp / f"o{rep}"), model="standin", allow_random_weights=True, fsize=2000, stride=1500, prophage=True, lc=500_000,
                   sensitivity=1.5, overwrite=True)
    dt = time.time() - t
    print(f"config 4: 8 x 5 Mbp genomes, {res['windows']} windows, {sum(len(r['ranges']) for r in res['prophage_regions'].values())} regions; "
          f"whole run {dt:.2f} s ({40 / dt:.1f} Mbp/s)", flush=True)
