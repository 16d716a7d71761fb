"""Numerics of the native sm_100a path at the BENCHMARKED shape and over a training run (round-2 VERDICT item 8).

(a) ResNet-34, batch 96, 224x224: one forward/backward with the native kernels (bf16) vs the PyTorch/cuDNN path in bf16
    autocast vs the PyTorch path in fp32, same weights and inputs: logits and per-parameter gradient relative errors.
(b) ResNet-18, batch 32, 64x64: 50 training steps (Momentum 0.01/0.9, fused update) native vs cuDNN path from the same
    seed and data: per-step loss difference.
Writes gpurun_out/numerics_r2.json (copied to profiles/)."""
import copy, json, os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, ".")
import fluxdist as fd
from fluxdist.parallel.engine import Trainer

BF = torch.bfloat16
def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

out = {}
# ---------------------------------------------------------------- (a)
torch.manual_seed(0)
B = int(os.environ.get("NUM_B", "96"))
m = fd.ResNet34(1000).cuda().to(memory_format=torch.channels_last)
x = torch.randn(B, 3, 224, 224, device="cuda").contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 1000, (B,), device="cuda")
def run(mode):
    mm = copy.deepcopy(m)
    if mode == "native":
        fd.nn.set_backend("native")
        lg = mm(x.to(BF)); l = fd.logitcrossentropy(lg, y)
    elif mode == "torch_bf16":
        fd.nn.set_backend("torch")
        with torch.autocast("cuda", dtype=BF):
            lg = mm(x); l = F.cross_entropy(lg.float(), y)
    else:
        fd.nn.set_backend("torch")
        lg = mm(x); l = F.cross_entropy(lg, y)
    l.backward()
    torch.cuda.synchronize()
    fd.nn.set_backend("auto")
    return lg.detach().float(), float(l), {n: p.grad.detach().float().clone() for n, p in mm.named_parameters()}
lg_n, l_n, g_n = run("native")
lg_b, l_b, g_b = run("torch_bf16")
lg_f, l_f, g_f = run("fp32")
per = {n: {"native_vs_fp32": rel(g_n[n], g_f[n]), "cudnn_bf16_vs_fp32": rel(g_b[n], g_f[n]), "native_vs_cudnn_bf16": rel(g_n[n], g_b[n])}
       for n in g_f}
worst = lambda k: max(per.items(), key=lambda kv: kv[1][k])
out["resnet34_b%d_224" % B] = {
    "loss": {"native": l_n, "cudnn_bf16": l_b, "fp32": l_f},
    "logits_rel_err": {"native_vs_fp32": rel(lg_n, lg_f), "cudnn_bf16_vs_fp32": rel(lg_b, lg_f), "native_vs_cudnn_bf16": rel(lg_n, lg_b)},
    "grad_rel_err_max": {k: {"param": worst(k)[0], "err": worst(k)[1][k]} for k in ("native_vs_fp32", "cudnn_bf16_vs_fp32", "native_vs_cudnn_bf16")},
    "grad_rel_err_median": {k: sorted(v[k] for v in per.values())[len(per) // 2] for k in ("native_vs_fp32", "cudnn_bf16_vs_fp32", "native_vs_cudnn_bf16")},
    "conv_weight_grads": {n: v for n, v in per.items() if n.endswith("conv.weight")},
}
print(json.dumps({k: v for k, v in out["resnet34_b%d_224" % B].items() if k != "conv_weight_grads"}, indent=1), flush=True)
del m, g_n, g_b, g_f
torch.cuda.empty_cache()
# ---------------------------------------------------------------- (b)
def curve(backend, steps=50):
    torch.manual_seed(1)
    model = fd.ResNet18(100)
    tr = Trainer(model, fd.Momentum(0.01, 0.9), fd.CudaDevice(0, 0), 32, image_hw=64, input="u8", graph=False, backend=backend, comm="native")
    g = torch.Generator().manual_seed(2)
    xs = torch.randint(0, 256, (8, 32, 64, 64, 3), dtype=torch.uint8, generator=g).cuda()
    ys = torch.randint(0, 100, (8, 32), generator=g).cuda()
    ls = []
    for i in range(steps):
        ls.append(float(tr.step(xs[i % 8], ys[i % 8])))
    tr.group.close()
    return ls
cn, ct = curve("native"), curve("torch")
out["resnet18_b32_64_loss_curve"] = {"native": cn, "cudnn_path": ct,
                                     "max_rel_diff": max(abs(a - b) / max(abs(b), 1e-6) for a, b in zip(cn, ct)),
                                     "max_abs_diff": max(abs(a - b) for a, b in zip(cn, ct))}
print("loss curve max rel diff", out["resnet18_b32_64_loss_curve"]["max_rel_diff"], "abs", out["resnet18_b32_64_loss_curve"]["max_abs_diff"])
print("native", [round(v, 3) for v in cn[::7]]); print("cudnn ", [round(v, 3) for v in ct[::7]])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/numerics_r2.json", "w"), indent=1)
