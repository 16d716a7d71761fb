"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals for ONE training step
(the launches between the last two fused all-reduce+update kernels)."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
names = [r[ix["Kernel Name"]] for r in data]
ends = [i for i, n in enumerate(names) if "fused_allreduce" in n]
a, b = ends[-2] + 1, ends[-1] + 1
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data[a:b]:
    n = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
    v = v / 1000 if u in ("ns", "nsecond") else v
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v for _, v in agg.values())
print(f"one step: {b - a} launches, {tot:.0f} us of kernel time (serialised, cold cache under ncu: compare shares)")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{v:8.1f} us {100 * v / tot:5.1f}% {c:4d}x  {n[:100]}")
