"""Top stall-sample SASS lines of an .ncu-rep (read here, no GPU needed): python scripts/ncu_hot.py file.ncu-rep [top-n]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
src, smp, ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if r and r[0] == "Address":      # next kernel of a multi-kernel report: keep the first one only
        break
    if len(r) == len(hdr):
        data.append(r)
tot = sum(int(r[smp]) for r in data)
print(f"total samples {tot}, instructions {sum(int(r[ex]) for r in data)}")
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print("stall totals:", ", ".join(f"{k[6:]} {v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 // max(tot, 1) > 0))
order = sorted(range(len(data)), key=lambda i: -int(data[i][smp]))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[c]), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:6d} {int(r[smp]) * 100.0 / tot:5.1f}%  ex={r[ex]:>9s}  {r[src].strip()[:90]:90s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
