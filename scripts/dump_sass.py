"""Commit-able SASS evidence (no GPU needed): `cuobjdump -sass` of the in-tree extension, split per kernel.

For every kernel of the product extension writes profiles/sass/<short-name>.sass containing the instruction stream
(address + mnemonic + operands; the hex encodings are dropped) of
  * tensor-core kernels: the regions around every UTCHMMA (tcgen05.mma), UTMALDG/UTMASTG (TMA), LDTM/STTM
    (tcgen05.ld/st), UTCBAR (tcgen05.commit) instruction — the main loop of the issuing warp and the epilogue —
    with +-WINDOW lines of context;
  * small kernels: the whole listing.
and profiles/sass_mnemonics_r2.txt with per-kernel counts of the Blackwell-specific mnemonics."""
import collections, os, re, subprocess, sys, glob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = glob.glob(os.path.join(ROOT, "fluxdistributed.jl_b200", "_C.*.so"))[0]
out_dir = os.path.join(ROOT, "profiles", "sass")
os.makedirs(out_dir, exist_ok=True)
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEY = re.compile(r"\b(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UTMAREDG|LDTM|STTM|UTCBAR|UBLKCP|LDGMC|STGMC|REDG|UTCATOMSWS|UCGABAR|SYNCS|ACQBULK|PREEXIT)")
WINDOW = 12
funcs = re.split(r"\n\s*Function : ", txt)[1:]
counts = {}
def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n
def short(d):
    d = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", d)
    d = re.sub(r"\((int|bool)\)", "", d)
    d = re.sub(r"\(.*", "", d)
    d = d.replace("void ", "").replace(" ", "")
    d = d.replace("<", "_").replace(">", "").replace(",", "_")
    return re.sub(r"[^A-Za-z0-9_-]", "_", d)[:110]
index = []
for f in funcs:
    name, _, body = f.partition("\n")
    dm = demangle(name.strip())
    lines = []
    for ln in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append(f"/*{m.group(1)}*/ {m.group(2).strip()};")
    c = collections.Counter()
    hits = []
    for i, ln in enumerate(lines):
        m = KEY.search(ln)
        if m:
            c[re.search(r"\s([A-Z][A-Z0-9_.]*)", " " + ln.split("*/", 1)[1]).group(1)] += 1
            hits.append(i)
    sn = short(dm)
    counts[sn] = c
    keep = set()
    if len(lines) <= 400:
        keep = set(range(len(lines)))
    else:
        for i in hits:
            keep.update(range(max(0, i - WINDOW), min(len(lines), i + WINDOW + 1)))
    with open(os.path.join(out_dir, sn + ".sass"), "w") as fh:
        fh.write(f"// {dm}\n// {len(lines)} SASS instructions; shown: {len(keep)} (regions around tcgen05 / TMA / TMEM / multimem / barrier instructions)\n")
        last = -2
        for i in sorted(keep):
            if i != last + 1:
                fh.write("  ...\n")
            fh.write(lines[i] + "\n")
            last = i
    index.append((sn, len(lines), sum(c.values())))
with open(os.path.join(ROOT, "profiles", "sass_mnemonics_r2.txt"), "w") as fh:
    fh.write("Per-kernel counts of Blackwell-specific SASS mnemonics (scripts/dump_sass.py; listings in profiles/sass/).\n"
             "UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = TMA load (tiled / IM2COL / MULTICAST),\n"
             "LDGMC/STGMC = multimem.ld_reduce / multimem.st, REDG = red.global (vector fp32 atomics), SYNCS = mbarrier ops,\n"
             "ACQBULK / PREEXIT = griddepcontrol.wait / griddepcontrol.launch_dependents (programmatic dependent launch).\n\n")
    for sn, n, k in sorted(index):
        c = counts[sn]
        if not c:
            continue
        fh.write(f"{sn}  ({n} instr)\n    " + ", ".join(f"{k_}:{v}" for k_, v in sorted(c.items(), key=lambda kv: -kv[1])) + "\n")
print(f"{len(index)} kernels -> {out_dir}")
