"""Development aid: timeline of CTA 0 of the tensor-core scan from the event trace of an instrumented build
(make TC_INSTRUMENT=1; QG_TC_TRACE=file python tools/tc_timing.py ...). usage: tc_trace.py FILE"""
import sys
from collections import defaultdict
ev = defaultdict(list)
for ln in open(sys.argv[1]):
    r, i, t, c = ln.split()
    ev[int(r)].append((int(t), int(c)))
names = {0: "producer", 1: "mma", 2: "epi_g0", 3: "epi_g1", 4: "helper"}
code_names = {1: {0: "start", 1: "xs_empty_ok", 2: "ring_empty_ok"},
              0: {0: "start", 1: "xs_empty_ok", 2: "ring_empty_ok"}}
t0 = min(e[0][0] for e in ev.values() if e)
for r in sorted(ev):
    e = ev[r]
    print(f"== {names[r]}: {len(e)} events, span {e[-1][0]-e[0][0]} cycles")
    # time spent before each code (delta to previous event), averaged over the steady state (skip first/last 10%)
    lo, hi = len(e) // 10, len(e) - len(e) // 10
    acc = defaultdict(lambda: [0, 0])
    for k in range(max(1, lo), hi):
        d = e[k][0] - e[k - 1][0]
        acc[(e[k - 1][1], e[k][1])][0] += d
        acc[(e[k - 1][1], e[k][1])][1] += 1
    tot = sum(v[0] for v in acc.values())
    for key, (s, n) in sorted(acc.items(), key=lambda x: -x[1][0]):
        print(f"   {key[0]}->{key[1]}: mean {s/n:8.1f} cycles x {n:5d}  = {100.0*s/tot:5.1f}%")
if len(sys.argv) > 2:
    n = int(sys.argv[2])
    allv = sorted((t - t0, r, c) for r in ev for (t, c) in ev[r])
    skip = len(allv) // 2
    for t, r, c in allv[skip:skip + n]:
        print(f"{t:9d} {'        ' * r}{names[r]}:{c}")
