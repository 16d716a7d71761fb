"""Turn gpurun_out/ncu_hot.ncu-rep into tracked text evidence (runs on the CPU box: `ncu -i`):
  profiles/r2/ncu_hot_kernels_summary.txt   one block per kernel: duration, cycles, tensor-pipe %, SM/L1/L2/DRAM %, bytes, regs,
                                occupancy, plus the top stall reasons and hottest SASS lines (source page)
  profiles/r2/ncu_hot_kernels_details.txt   `--page details` of every captured kernel (all sections)"""
import csv, io, os, subprocess, sys
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ncu_hot.ncu-rep"
labels = []
try:
    sys.path.insert(0, "scripts")
    import re
    src = open("scripts/ncu_hot.py").read()
    labels = eval(re.search(r"LABELS = (\[.*?\])\n", src, re.S).group(1))
except Exception:
    pass
def run(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout
raw = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, units, rows = raw[0], raw[1], raw[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = [("time", "gpu__time_duration.sum"), ("cycles", "sm__cycles_elapsed.max"), ("grid", "launch__grid_size"),
        ("block", "launch__block_size"), ("regs/thread", "launch__registers_per_thread"),
        ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
        ("tensor pipe % of active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L1/TEX %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"),
        ("L2->L1 bytes", "l1tex__m_xbar2l1tex_read_bytes.sum"),
        ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("warp instructions", "smsp__inst_executed.sum")]
src_csv = list(csv.reader(io.StringIO(run("--page", "source", "--csv"))))
blocks, i = [], 0
while i < len(src_csv):
    if src_csv[i] and src_csv[i][0] == "Kernel Name":
        name = src_csv[i][1]; h = src_csv[i + 1]; hi = {x: k for k, x in enumerate(h)}
        j = i + 2; data = []
        while j < len(src_csv) and not (src_csv[j] and src_csv[j][0] == "Kernel Name"):
            if len(src_csv[j]) == len(h): data.append(src_csv[j])
            j += 1
        blocks.append((name, hi, data)); i = j
    else:
        i += 1
by_name = {}
for b in blocks:
    by_name.setdefault(b[0], b)
out = ["ncu --set full --clock-control none --import-source on (scripts/ncu_hot.sh -> scripts/ncu_hot.py; B200, second launch of",
       "every kernel, ResNet-34 batch-96 shapes).  Times under the profiler are NOT benchmark numbers.", ""]
for k, r in enumerate(rows):
    lab = labels[k] if k < len(labels) else ""
    out.append(f"== [{k + 1}] {lab}")
    out.append("   " + r[ix["Kernel Name"]][:150])
    for nm, key in want:
        if key in ix:
            out.append(f"   {nm:26s} {r[ix[key]]:>16s} {units[ix[key]]}")
    blk = by_name.pop(r[ix["Kernel Name"]], None)  # the source page holds ONE entry per kernel function: show it once
    if blk is not None:
        name, hi, data = blk
        tot = sum(int(d[hi["# Samples"]]) for d in data) or 1
        stall = [h for h in hi if h.startswith("stall_") and "Not Issued" not in h]
        agg = sorted(((sum(int(d[hi[h]]) for d in data), h[6:]) for h in stall), reverse=True)[:6]
        out.append("   stall samples: " + ", ".join(f"{n}={100 * v // tot}%" for v, n in agg) + f"  ({tot} samples)")
        for d in sorted(data, key=lambda d: -int(d[hi["# Samples"]]))[:8]:
            out.append(f"      {int(d[hi['# Samples']]):6d} smp  {d[hi['Source']].strip()[:110]}")
    out.append("")
os.makedirs("profiles", exist_ok=True)
open("profiles/r2/ncu_hot_kernels_summary.txt", "w").write("\n".join(out))
open("profiles/r2/ncu_hot_kernels_details.txt", "w").write(run("--page", "details"))
print("\n".join(out[:60]))
