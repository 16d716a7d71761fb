import sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
BF = torch.bfloat16
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
x = cl(torch.randn(96, 64, 56, 56, device="cuda").to(BF)); w = cl((torch.randn(64, 64, 3, 3, device="cuda") * 0.05).to(BF))
for _ in range(3):
    ops.conv3x3_dgrad(x, w, variant=1)
    ops.conv3x3_fprop(x, w, stats=True, variant=1)
    ops.conv3x3_wgrad(x, x)
torch.cuda.synchronize()
