"""Print the SASS instructions with the most warp-stall samples from `ncu --page source --csv`.
usage: ncu -i rep.ncu-rep --page source --csv | python tools/ncu_top_stalls.py [N]"""
import csv, sys
n_top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rows = list(csv.reader(sys.stdin))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try:
        n = int(r[ci["# Samples"]])
    except Exception:
        continue
    data.append((n, r))
tot = sum(n for n, _ in data)
print("total samples", tot)
for idx, (n, r) in enumerate(data):
    r.append(idx)
top = sorted(data, key=lambda x: -x[0])[:n_top]
for n, r in top:
    st = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{n:7d} {100*n/tot:5.1f}%  #{r[-1]:5d} {r[ci['Source']].strip()[:70]:70s} {st}")
