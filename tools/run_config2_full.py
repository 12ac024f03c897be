"""BASELINE config 2 at full size through the real driver: write a synthetic 1 Gbp FASTA of 2-50 kbp
contigs, run `jaeger_b200.predict.run_core` on it (FASTA parse, pack, dust, encode, forward,
aggregation, terminal repeats, TSV) and report the wall time of every part.
usage: python tools/run_config2_full.py [Gbp]"""
import sys, time, tempfile, logging
sys.path.insert(0, ".")
from pathlib import Path
import numpy as np
from bench import synth_lens, synth_bases

gbp = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
tmp = Path(tempfile.mkdtemp())
fa = tmp / "assembly.fasta"
t = time.time()
lens = synth_lens(1, int(gbp * 1e9))
with open(fa, "wb") as fh:
    for start in range(0, len(lens), 2000):                     # generate in slabs to bound host memory
        part = lens[start:start + 2000]
        bases = synth_bases(7919 + start, part)
        off = np.concatenate([[0], np.cumsum(part)])
        for k in range(len(part)):
            fh.write(b">c%d\n" % (start + k))
            fh.write(bases[off[k]:off[k + 1]].tobytes())
            fh.write(b"\n")
print(f"wrote {fa.stat().st_size / 1e9:.2f} GB FASTA, {len(lens)} contigs in {time.time() - t:.1f} s", flush=True)
logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s")
from jaeger_b200.predict import run_core
for dust in (False, True):
    t = time.time()
    res = run_core(input=str(fa), output=str(tmp / f"out{int(dust)}"), model="standin", allow_random_weights=True, fsize=2000, stride=1500, overwrite=True, dustmask=dust)
    dt = time.time() - t
    print(f"dustmask={dust}: {res['windows']} windows, {res['num_written']} contigs written; predict stage {res['predict_seconds']:.1f} s, "
          f"whole run {dt:.1f} s -> {gbp * 1e3 / dt:.1f} Mbp/s end to end from the file", flush=True)
