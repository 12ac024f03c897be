"""Driver-level multi-GPU check: `python -m jaeger_b200.predict` on one GPU vs under torchrun on N GPUs
must write the same tables (rows in the same order; scores within the fp32-atomics noise).
usage: python tools/check_multigpu_driver.py [n_gpus]"""
import subprocess, sys, tempfile
from pathlib import Path
import numpy as np, pandas as pd

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rng = np.random.default_rng(12)
tmp = Path(tempfile.mkdtemp())
fa = tmp / "meta.fasta"
lens = list(rng.integers(2000, 50000, 60)) + [700, 1500, 999, 2000, 650_000]
core = "".join(rng.choice(list("ACGT"), 150))
with open(fa, "w") as fh:
    for i, n in enumerate(lens):
        s = "".join(rng.choice(list("ACGT"), int(n)))
        if i % 9 == 4:
            s = core + s + core
        fh.write(f">ctg{i} len={n}\n")
        for k in range(0, len(s), 80):
            fh.write(s[k:k + 80] + "\n")
import yaml
cal = tmp / "standin_refine.yaml"
cal.write_text(yaml.safe_dump({"schema_version": 1, "jaeger_model": "standin", "taus": {
    c: {"logit": -0.5, "margin": 0.01, "n": 100} for c in ("phage", "virus", "archaea", "bacteria", "plasmid", "eukarya")}}))
common = ["-m", "standin", "--allow-random-weights", "-i", str(fa), "--min-len", "500", "-p", "--lc", "500000", "-s", "0.2", "--overwrite", "--refine", "--refine-file", str(cal)]
subprocess.run([sys.executable, "-m", "jaeger_b200.predict", *common, "-o", str(tmp / "one")], check=True)
subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}", "--master-addr", "127.0.0.1",
                "--master-port", "29541", "-m", "jaeger_b200.predict", *common, "-o", str(tmp / "many")], check=True)
a = pd.read_csv(tmp / "one" / "standin" / "meta.tsv", sep="\t", keep_default_na=False)
b = pd.read_csv(tmp / "many" / "standin" / "meta.tsv", sep="\t", keep_default_na=False)
assert list(a.columns) == list(b.columns) and len(a) == len(b) >= len(lens) - 3, (len(a), len(b))
for col in a.columns:
    if a[col].dtype.kind == "f":
        assert np.allclose(a[col].to_numpy(), b[col].to_numpy(), atol=2e-3), col
    else:
        assert a[col].tolist() == b[col].tolist(), col
pa = (tmp / "one" / "standin" / "meta_prophage_regions.tsv").read_text()
pb = (tmp / "many" / "standin" / "meta_prophage_regions.tsv").read_text()
assert pa == pb, (pa, pb)
assert "contig_call" in a.columns and (a["n_windows_used"] != "").sum() > 10
ra, rb = (tmp / d / "standin" / "meta_prophages" / "prophages_jaeger.tsv" for d in ("one", "many"))
assert ra.exists() == rb.exists() and (not ra.exists() or ra.read_text() == rb.read_text())
print("att-site report rows:", len(ra.read_text().splitlines()) - 1 if ra.exists() else 0)
print(f"multi-GPU driver check ok: {len(a)} rows identical on 1 and {n_gpus} GPUs; DTR rows: {(a['terminal_repeats'] != '').sum()}")
