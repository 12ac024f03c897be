"""Segments the SASS of an ncu report's kernel into runs of equal execution count (development aid: who spends the samples).
usage: python tools/ncu_runs.py report.ncu-rep [min_samples]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; min_s = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]; data = rows[2:]
isrc, ins, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
print("total samples", sum(int(r[ins] or 0) for r in data))
runs = []
for i, r in enumerate(data):
    ex = int(r[iex] or 0); sm = int(r[ins] or 0)
    op = r[isrc].split()[0] if r[isrc].split() else ""
    if op.startswith("@") and len(r[isrc].split()) > 1: op = r[isrc].split()[1]
    if runs and runs[-1][2] == ex: runs[-1][1] += 1; runs[-1][3] += sm; runs[-1][5].append(op)
    else: runs.append([i, 1, ex, sm, r[isrc][:50], [op]])
for st, n, ex, sm, first, ops in runs:
    if sm > min_s:
        keys = [o for o in ops if o.startswith(("UTC", "SYNCS", "LDTM", "STS", "LDS", "ATOM", "RED", "FENCE", "MUFU", "LDG", "BAR", "LDL", "STL", "LDC", "MEMBAR"))]
        print(st, n, ex, sm, first.strip(), dict(collections.Counter(k.split(".")[0] for k in keys)))
