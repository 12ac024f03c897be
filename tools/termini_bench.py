"""Throughput of the terminal-repeat scan on the bench workload (development aid)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from jaeger_b200 import B200Engine, parse_project, standin_1p4m_config
from jaeger_b200.termini import scan_terminal_repeats, scan_lengths
from bench import synth_batch
eng = B200Engine(spec=parse_project(standin_1p4m_config()), workspace_gb=2)
seq, lens = synth_batch(1, int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else int(64e6))
offsets = np.zeros(len(lens) + 1, np.int64); np.cumsum(lens, out=offsets[1:])
names = [f"c{i}" for i in range(len(lens))]
with torch.cuda.stream(eng._stream()):
    codes, valid = eng.pack(torch.from_numpy(seq).to(eng.tdev))
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    df = scan_terminal_repeats(eng, codes, valid, offsets, names, 2000)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    n = scan_lengths(lens[lens >= 2000])
    cells = 2.0 * (n.astype(np.float64) ** 2).sum()
    print(f"{len(lens)} contigs, {lens.sum()/1e6:.1f} Mbp: {dt*1e3:.1f} ms, {cells/dt/1e9:.1f} GCUPS, {lens.sum()/1e6/dt:.1f} Mbp/s, hits {df['terminal_repeats'].notna().sum()}")
