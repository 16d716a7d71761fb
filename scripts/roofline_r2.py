"""Class-level roofline of one ResNet-34 training step (batch 96, 224x224, bf16) from an ncu launch list.

    python scripts/roofline_r2.py profiles/r2/launches_final.csv [step_ms]

Time per class = sum of `gpu__time_duration.sum` of the kernels of ONE step (launches between the last two fused
all-reduce+update kernels).  Under ncu the kernels run serialised, so the per-class times add up to MORE than the
measured step (weight gradients overlap the main stream in the real step); they are used for shares and for the
achieved-throughput column only.  Work per class is derived from the model, not measured:

  conv FLOPs   2 * MACs of every conv of torchvision-style ResNet-34 at 224x224 (fprop; dgrad = fprop minus the stem;
               wgrad = fprop), times the batch.
  BN bytes     minimal traffic of each pass: forward reads y (+ residual) and writes the activation; backward reads
               dout, y (+ the saved activation for residual layers) and writes dy (the single-pass bound; the two-pass
               kernel re-reads its inputs, mostly from L2).
Denominators: MEASURED_PEAKS.json (copy bandwidth, cuBLAS bf16 burst / sustained).
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 96


def resnet34_layers():
    """(kind, H_out, C_in, C_out, k, residual) for every conv+BN of ResNet-34."""
    out = [("stem", 112, 3, 64, 7, False)]
    cfg = [(64, 3, 56), (128, 4, 28), (256, 6, 14), (512, 3, 7)]
    cin = 64
    for c, nblk, hw in cfg:
        for b in range(nblk):
            out.append(("conv", hw, cin if b == 0 else c, c, 3, False))
            out.append(("conv", hw, c, c, 3, True))
            if b == 0 and cin != c:
                out.append(("proj", hw, cin, c, 1, False))
        cin = c
    return out


def work():
    L = resnet34_layers()
    fl = lambda l: 2.0 * B * l[1] * l[1] * l[2] * l[3] * l[4] * l[4]
    fprop = sum(fl(l) for l in L)
    stem = fl(L[0])
    elems = lambda l: B * l[1] * l[1] * l[3]
    bn_fwd = sum((2 + (1 if l[5] else 0)) * elems(l) * 2 for l in L[1:])          # stem BN is fused with the pool
    bn_bwd = sum((3 + (1 if l[5] else 0)) * elems(l) * 2 for l in L)               # incl. the stem's BN backward
    # stem tail: BN+ReLU+pool forward reads the conv output, writes pooled output (1/4) + arg-max bytes (1/4 x 1 B);
    # the pool backward reads pooled gradient + arg-max bytes and writes the full-resolution gradient
    stem_tail = (elems(L[0]) * 2 + elems(L[0]) // 4 * 3) + (elems(L[0]) // 4 * 3 + elems(L[0]) * 2)
    params = 21_797_672
    upd = params * (4 + 4 + 4 + 4 + 4 + 2)  # grad r, master r/w, momentum r/w, bf16 w
    return dict(fprop=fprop, dgrad=fprop - stem, wgrad=fprop, bn_fwd=bn_fwd, bn_bwd=bn_bwd, stem_tail=stem_tail, upd=upd)


def classify(n):
    m = re.search(r"igemm_kernel<(\d+), (\d+)", n)
    if m:
        return {1: "fprop", 2: "dgrad", 6: "dgrad", 3: "wgrad"}.get(int(m.group(2)), "other (head GEMMs)")
    m = re.search(r"conv_halo_kernel<(\d+), (\d+), (\d+), (\d+)", n)
    if m:
        return "dgrad" if m.group(2) == "1" else "fprop"
    if "wgrad" in n:
        return "wgrad"
    if "bn_apply" in n:
        return "bn_fwd"
    if "bn_bwd" in n:
        return "bn_bwd"
    if "pool" in n:
        return "stem_tail"
    if "fused_allreduce" in n:
        return "upd"
    return "other (head GEMMs)" if "igemm" in n else "other"


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2", "launches_final.csv")
    step_ms = float(sys.argv[2]) if len(sys.argv) > 2 else None
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    bw, burst, sus = peaks["hbm_gbs"] * 1e9, peaks["bf16_tflops"] * 1e12, peaks["bf16_tflops_sustained"] * 1e12
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    ix = {h: i for i, h in enumerate(rows[hi])}
    data = [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    names = [r[ix["Kernel Name"]] for r in data]
    ends = [i for i, n in enumerate(names) if "fused_allreduce" in n]
    a, b = ends[-2] + 1, ends[-1] + 1
    t = collections.defaultdict(float)
    c = collections.Counter()
    for r in data[a:b]:
        k = classify(r[ix["Kernel Name"]])
        t[k] += float(r[ix["Metric Value"]]) * 1e-9
        c[k] += 1
    W = work()
    tot = sum(t.values())
    print(f"one step: {b - a} launches, {tot * 1e3:.2f} ms of kernel time serialised under ncu"
          + (f"; measured step {step_ms:.3f} ms" if step_ms else ""))
    print("| class | launches | time under ncu | work (model-derived) | achieved | of measured peak |")
    print("|---|---|---|---|---|---|")
    label = dict(fprop="conv fprop (incl. stem, BN-statistics epilogues)", dgrad="conv dgrad (incl. residual / dz epilogues)",
                 wgrad="conv wgrad (side stream)", bn_fwd="BN apply (+residual, ReLU)", bn_bwd="BN backward",
                 stem_tail="stem tail (BN+ReLU+pool fwd, pool bwd)", upd="fused update (1 GPU: no exchange)")
    for k in ("fprop", "dgrad", "wgrad", "bn_fwd", "bn_bwd", "stem_tail", "upd"):
        if k in ("fprop", "dgrad", "wgrad"):
            rate = W[k] / t[k]
            print(f"| {label[k]} | {c[k]} | {t[k] * 1e6:.0f} us | {W[k] / 1e12:.3f} TFLOP | {rate / 1e12:.0f} TFLOP/s | "
                  f"{100 * rate / burst:.0f} % of cuBLAS burst, {100 * rate / sus:.0f} % of sustained |")
        else:
            rate = W[k] / t[k]
            print(f"| {label[k]} | {c[k]} | {t[k] * 1e6:.0f} us | {W[k] / 1e6:.0f} MB min. traffic | {rate / 1e12:.2f} TB/s | "
                  f"{100 * rate / bw:.0f} % of copy bandwidth |")
    rest = sum(v for k, v in t.items() if k.startswith("other"))
    print(f"| everything else (head, loss, input normalise, s2d, fills) | {sum(v for k, v in c.items() if k.startswith('other'))} | {rest * 1e6:.0f} us | | | |")
    if step_ms:
        total_flop = W["fprop"] + W["dgrad"] + W["wgrad"]
        r = total_flop / (step_ms * 1e-3)
        print(f"| **whole step (measured, not under ncu)** | | {step_ms:.3f} ms | {total_flop / 1e12:.2f} TFLOP | {r / 1e12:.0f} TFLOP/s | "
              f"{100 * r / burst:.0f} % of burst, {100 * r / sus:.0f} % of sustained |")


if __name__ == "__main__":
    main()
