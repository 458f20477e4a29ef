#!/usr/bin/env python3
"""Per-instruction view of an .ncu-rep source page: top instructions by stall samples and by executed count.
usage: ncu_source.py REPORT.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ins = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr):
        continue
    s = int(r[col["# Samples"]] or 0)
    ex = int(r[col["Instructions Executed"]] or 0)
    st = {h[6:]: int(r[col[h]] or 0) for h in stall_cols}
    ins.append((n, r[col["Source"]].strip(), s, ex, st))
tot_s = sum(i[2] for i in ins)
tot_e = sum(i[3] for i in ins)
print(f"total samples {tot_s}, total warp-instructions {tot_e}")
agg = {}
for i in ins:
    for k, v in i[4].items():
        agg[k] = agg.get(k, 0) + v
print("stall totals:", ", ".join(f"{k} {v}" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v))
print(f"--- top {top} by samples")
for n, src, s, ex, st in sorted(ins, key=lambda x: -x[2])[:top]:
    why = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3] if v)
    print(f"{n:6d} {s:7d} {100.0*s/tot_s:5.1f}%  exec {ex:9d}  {src[:70]:70s} {why}")
print(f"--- top {top} by executed count")
for n, src, s, ex, st in sorted(ins, key=lambda x: -x[3])[:top]:
    print(f"{n:6d} exec {ex:9d} {100.0*ex/tot_e:5.1f}%  samples {s:7d}  {src[:80]}")
if len(sys.argv) > 3:
    lo, hi2 = [int(x) for x in sys.argv[3].split(":")]
    print(f"--- instructions {lo}..{hi2}")
    for n, src, s, ex, st in ins[lo:hi2]:
        why = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:2] if v)
        print(f"{n:6d} {s:5d} exec {ex:8d}  {src[:90]:90s} {why}")
