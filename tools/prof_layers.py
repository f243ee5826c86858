"""A few representative igemm launches for ncu (one launch each after a warm-up launch):
  0 head dgrad (packed 8-ch dY -> 64 ch, padded grid)   1 stem fwd (packed 8 ch -> 64, stats)
  2 resblock wgrad                                     3 E.layer1 conv fwd (56x56x64 -> 64, BN stats)
  4 head fwd (64 -> 3, tanh, NCHW)                      5 up2 phase fwd (ConvT 128 -> 64)
  6 resblock fwd 256 -> 256 3x3 with per-sample statistics (the flagship shape)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcgan_b200 import _lib as L, conv as CV, ops
from pcgan_b200.plan import Geom, OutMap
DEV = "cuda"
N = 64
WHICH = [int(x) for x in os.environ.get("WHICH", "0,1,2,3,4,5").split(",")]


def buf(g):
    return torch.randn(g.numel + 512, device=DEV).to(torch.bfloat16)


def run(plans, a, out, bias=None, stats=None, extra=None, reps=int(os.environ.get("REPS", "1"))):
    rs = []
    for sp, wm in plans:
        b = (torch.randn(sp.b_rows * sp.b_k + 64, device=DEV) * 0.02).to(torch.bfloat16)
        rs.append((ops.Igemm(sp), b))
    for _ in range(reps):
        for g, b in rs:
            if extra is not None:
                g.run(a, extra, out)
            else:
                g.run(a, b, out, bias, stats)
    torch.cuda.synchronize()


S = 128
if 0 in WHICH:
    dyg, xg = Geom(N, S, S, 8, 6), Geom(N, S, S, 64, 3)
    full = Geom(N, S + 6, S + 6, 64, 0)
    run(CV.conv_dgrad_plans((3, 64, 7, 7), dyg, xg, 1, 3, OutMap.nhwc(full), full_padded=True, note="head.dgrad"), buf(dyg), torch.zeros(full.numel + 512, dtype=torch.bfloat16, device=DEV))
if 1 in WHICH:
    xg, rg = Geom(N, S, S, 8, 3), Geom(N, S, S, 64, 0)
    run(CV.conv_fwd_plans((64, 4, 7, 7), xg, 1, 3, OutMap.nhwc(rg), stats=True, per_sample_stats=True, note="stem.fwd"), buf(xg),
        torch.zeros(rg.numel + 512, dtype=torch.bfloat16, device=DEV), bias=torch.zeros(64, device=DEV), stats=torch.zeros(N, 64, 2, device=DEV))
if 2 in WHICH:
    xg = Geom(N, 32, 32, 256, 1)
    sp, wm = CV.conv_wgrad_plan((256, 256, 3, 3), xg, xg, 1, 1, note="res.wgrad")
    run([(sp, wm)], buf(xg), torch.zeros(sp.b_rows * sp.b_k, device=DEV), extra=buf(xg))
if 3 in WHICH:
    xg, rg = Geom(N, 56, 56, 64, 1), Geom(N, 56, 56, 64, 0)
    run(CV.conv_fwd_plans((64, 64, 3, 3), xg, 1, 1, OutMap.nhwc(rg), stats=True, note="E.layer1.fwd"), buf(xg),
        torch.zeros(rg.numel + 512, dtype=torch.bfloat16, device=DEV), stats=torch.zeros(1, 64, 2, device=DEV))
if 4 in WHICH:
    xg = Geom(N, S, S, 64, 3)
    run(CV.conv_fwd_plans((3, 64, 7, 7), xg, 1, 3, OutMap.nchw(N, 3, S, S), act=L.ACT_TANH, note="head.fwd"), buf(xg),
        torch.zeros(N, 3, S, S, device=DEV), bias=torch.zeros(3, device=DEV))
if 5 in WHICH:
    xg, rg = Geom(N, 64, 64, 128, 1), Geom(N, S, S, 64, 0)
    plans = CV.conv_fwd_plans((128, 64, 3, 3), xg, 2, 1, OutMap.nhwc(rg), transposed=True, output_padding=1, stats=True, per_sample_stats=True, note="up2.fwd")
    run(plans[3:4], buf(xg), torch.zeros(rg.numel + 512, dtype=torch.bfloat16, device=DEV), bias=torch.zeros(64, device=DEV), stats=torch.zeros(N, 64, 2, device=DEV))
if 6 in WHICH:
    xg, rg = Geom(N, 32, 32, 256, 1), Geom(N, 32, 32, 256, 0)
    run(CV.conv_fwd_plans((256, 256, 3, 3), xg, 1, 1, OutMap.nhwc(rg), stats=True, per_sample_stats=True, note="res.fwd"), buf(xg),
        torch.zeros(rg.numel + 512, dtype=torch.bfloat16, device=DEV), bias=torch.zeros(256, device=DEV), stats=torch.zeros(N, 256, 2, device=DEV))
print("done")
