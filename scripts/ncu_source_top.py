"""Summarise `ncu -i X.ncu-rep --page source --csv` output: per kernel, the SASS instructions holding most stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; ix = {h: k for k, h in enumerate(hdr)}
        j = i + 2; data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr): data.append(rows[j])
            j += 1
        tot = sum(int(r[ix["# Samples"]]) for r in data)
        print(f"== {name[:110]}\n   {len(data)} SASS instructions, {tot} stall samples")
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[ix[h]]) for r in data) for h in stall_cols}
        print("   by reason:", ", ".join(f"{h[6:]}={v}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
            print(f"   {int(r[ix['# Samples']]):6d} smp  exec {r[ix['Instructions Executed']]:>8}  {r[ix['Source']].strip()[:100]}")
        i = j
    else:
        i += 1
