import sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
e = ops.probe_extension()
torch.manual_seed(0)
rows = 320
A = torch.randn(rows, 64, device="cuda").bfloat16()
B = torch.randn(64, 64, device="cuda").bfloat16()
def run(shift, sbo, bo, mn, nk16):
    D = torch.zeros(128, 64, device="cuda")
    e.probe_shift(A, B, D, shift, sbo, bo, mn, nk16)
    torch.cuda.synchronize()
    pitch = sbo // 128
    if not mn:
        idx = torch.tensor([shift + (m // 8) * pitch + m % 8 for m in range(128)], device="cuda")
        ref = A[idx].float()[:, :16 * nk16] @ B.float()[:, :16 * nk16].t()
        got = D
    else:
        kidx = torch.tensor([shift + (k // 8) * pitch + k % 8 for k in range(16 * nk16)], device="cuda")
        At = A[kidx].float()          # [k, 64 m]
        ref = At.t() @ B.float()[:, :16 * nk16].t()   # [64 m, 64 n]
        got = D[:64]
    err = ((got - ref).norm() / ref.norm()).item()
    return err
for mn in (False, True):
    for (shift, sbo, bo) in [(0,1024,0),(1,1024,0),(1,1024,1),(3,1024,0),(3,1024,3),(0,1280,0),(1,1280,0),(11,1280,0),(11,1280,3),(0,2048,0),(5,2048,0),(5,2048,5),(0,1152,0),(2,2304,0)]:
        try:
            err = run(shift, sbo, bo, mn, 4)
        except Exception as ex:
            err = "EXC " + str(ex)[:80]
        print(f"mn_major={int(mn)} shift={shift:2d} sbo={sbo} base_off={bo}: rel_err={err}")
