"""Top stall samples per SASS instruction of an ncu report (development aid).
usage: python tools/ncu_src_top.py report.ncu-rep [n_top] [filter-substring]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
isrc, ins, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ins] or 0) for r in data)
by_reason = collections.Counter()
by_op = collections.Counter()
for r in data:
    for i in stall:
        by_reason[hdr[i]] += int(r[i] or 0)
    op = r[isrc].split()[0] if r[isrc].split() and not r[isrc].split()[0].startswith("@") else (r[isrc].split()[1] if len(r[isrc].split()) > 1 else "")
    by_op[op.split(".")[0]] += int(r[ins] or 0)
print("total samples", tot)
print("by stall reason:", by_reason.most_common(10))
print("by opcode:", by_op.most_common(25))
for r in sorted(data, key=lambda r: -int(r[ins] or 0))[:ntop]:
    st = sorted(((hdr[i], int(r[i] or 0)) for i in stall if int(r[i] or 0) > 0), key=lambda kv: -kv[1])[:3]
    print(r[ins].rjust(6), r[iex].rjust(9), r[isrc][:80].ljust(80), st)
