import sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
e = ops.probe_extension()
torch.manual_seed(0)
rows = 512
A = torch.randn(rows, 16, device="cuda").bfloat16()
B = torch.randn(64, 16, device="cuda").bfloat16()
for mn in (False, True):
    for shift, sbo in [(0, 256), (1, 256), (5, 256), (0, 352), (3, 352), (12, 352), (0, 512)]:
        D = torch.zeros(128, 64, device="cuda")
        e.probe_shift32(A, B, D, shift, sbo, mn); torch.cuda.synchronize()
        pitch = sbo // 32
        if not mn:
            idx = torch.tensor([shift + (m // 8) * pitch + m % 8 for m in range(128)], device="cuda")
            ref = A[idx].float() @ B.float().t(); got = D
        else:
            kidx = torch.tensor([shift + (k // 8) * pitch + k % 8 for k in range(16)], device="cuda")
            ref = A[kidx].float().t() @ B.float().t(); got = D[:16]
        print(f"SW32 mn_major={int(mn)} shift={shift} sbo={sbo}: rel_err={((got-ref).norm()/ref.norm()).item():.3e}")
