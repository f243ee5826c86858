"""Two igemm launches for ncu: the flagship resblock conv with per-sample stats and the tiny-K D stem conv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcgan_b200 import _lib as L, conv as CV, ops
from pcgan_b200.plan import Geom, OutMap
DEV = "cuda"
N = 64
def run(plans, a, out, bias=None, stats=None, reps=3):
    rs = []
    for sp, wm in plans:
        b = (torch.randn(sp.b_rows * sp.b_k + 64, device=DEV) * 0.02).to(torch.bfloat16)
        rs.append((ops.Igemm(sp), b))
    for _ in range(reps):
        for g, b in rs:
            g.run(a, b, out, bias, stats)
    torch.cuda.synchronize()
xg, rg = Geom(N, 32, 32, 256, 1), Geom(N, 32, 32, 256, 0)
x = torch.randn(xg.numel + 512, device=DEV).to(torch.bfloat16)
out = torch.zeros(rg.numel + 512, dtype=torch.bfloat16, device=DEV)
run(CV.conv_fwd_plans((256, 256, 3, 3), xg, 1, 1, OutMap.nhwc(rg), stats=True, per_sample_stats=True, note="flagship"), x, out,
    bias=torch.zeros(256, device=DEV), stats=torch.zeros(N, 256, 2, device=DEV))
xg0, y0 = Geom(N, 128, 128, 8, 1), Geom(N, 64, 64, 64, 1)
x0 = torch.randn(xg0.numel + 512, device=DEV).to(torch.bfloat16)
o0 = torch.zeros(y0.numel + 512, dtype=torch.bfloat16, device=DEV)
run(CV.conv_fwd_plans((64, 4, 4, 4), xg0, 2, 1, OutMap.nhwc(y0), act=L.ACT_LRELU, act_slope=0.2, note="dstem"), x0, o0, bias=torch.zeros(64, device=DEV))
print("done")
