"""Streaming ingest at scale: wall time and peak host RSS of `python -m jaeger_b200.predict` on synthetic assemblies of
increasing size, streamed in 256 Mbp chunks (bounded memory: the RSS must not grow with the file), and -- with N > 1 GPUs --
the same file under torchrun, every rank streaming its own byte slice; the tables must be identical.
usage: python tools/stream_check.py [sizes in Gbp, comma separated] [n_gpus] [multi-only]   (run on the GPU box; writes
gpurun_out/stream_check.json).  `multi-only` skips the single-process run of every file (large files on many GPUs: only the torchrun
run is timed, the table is checked for its row count); files above 2 Gbp repeat a 1 Gbp block of sequence under fresh record names."""
import json, os, resource, subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from bench import synth_bases, synth_lens

sizes = [float(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,3").split(",")]
n_gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
multi_only = len(sys.argv) > 3 and sys.argv[3] == "multi-only"
tmp = Path(tempfile.mkdtemp())
out = {"chunk_mbp": 256, "runs": []}


def write_fasta(path, gbp, seed):
    if gbp > 2:        # a 1 Gbp block of records, repeated under fresh names (generation time, not content, is the point)
        block = path.with_suffix(".block")
        n1 = write_fasta(block, 1.0, seed)
        body = block.read_bytes()
        recs = body.split(b">")[1:]
        with open(path, "wb") as fh:
            for rep in range(int(round(gbp))):
                fh.write(b"".join(b">r%d_" % rep + r for r in recs))
        block.unlink()
        return n1 * int(round(gbp))
    lens = synth_lens(seed, int(gbp * 1e9))
    with open(path, "wb") as fh:
        for start in range(0, len(lens), 2000):
            part = lens[start:start + 2000]
            bases = synth_bases(7919 * seed + start, part)
            off = np.concatenate([[0], np.cumsum(part)])
            for k in range(len(part)):
                fh.write(b">c%d\n" % (start + k)); fh.write(bases[off[k]:off[k + 1]].tobytes()); fh.write(b"\n")
    return len(lens)


def run(cmd):
    before = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
    t = time.time()
    subprocess.run(cmd, check=True, cwd=ROOT)
    return time.time() - t, resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss


for i, gbp in enumerate(sizes):
    fa = tmp / f"asm{i}.fasta"
    n = write_fasta(fa, gbp, i + 1)
    common = ["-i", str(fa), "-m", "standin", "--allow-random-weights", "--overwrite", "--stream-mbp", "256", "--no-terminal-repeats"]
    rec = {"gbp": gbp, "contigs": n, "file_gb": fa.stat().st_size / 1e9}
    if not multi_only:
        dt, rss_kb = run([sys.executable, "-m", "jaeger_b200.predict", *common, "-o", str(tmp / f"one{i}")])
        rec.update({"wall_s": round(dt, 1), "mbp_per_s_incl_startup": round(gbp * 1e3 / dt, 1), "peak_rss_gb_so_far": round(rss_kb / 1e6, 2)})
    if n_gpus > 1 and multi_only:
        dt2, rss_kb = run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}", "--master-addr", "127.0.0.1",
                           "--master-port", "29551", "-m", "jaeger_b200.predict", *common, "-o", str(tmp / f"many{i}")])
        rows = sum(1 for _ in open(tmp / f"many{i}" / "standin" / f"asm{i}.tsv")) - 1
        rec.update({"n_gpus": n_gpus, "wall_s_multi": round(dt2, 1), "mbp_per_s_multi_incl_startup": round(gbp * 1e3 / dt2, 1),
                    "table_rows": rows, "peak_rss_gb_largest_process": round(rss_kb / 1e6, 2)})
    elif n_gpus > 1:
        dt2, _ = run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}", "--master-addr", "127.0.0.1",
                      "--master-port", "29551", "-m", "jaeger_b200.predict", *common, "-o", str(tmp / f"many{i}")])
        a = (tmp / f"one{i}" / "standin" / f"asm{i}.tsv").read_text()
        b = (tmp / f"many{i}" / "standin" / f"asm{i}.tsv").read_text()
        rec.update({"n_gpus": n_gpus, "wall_s_multi": round(dt2, 1), "tables_identical": a == b})
    out["runs"].append(rec)
    print(json.dumps(rec), flush=True)
    fa.unlink()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "stream_check.json").write_text(json.dumps(out, indent=1))
