"""Micro-benchmark of the igemm kernel on the shapes of the step (CUDA events, L2 flushed between launches)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcgan_b200 import _lib as L, conv as CV, ops
from pcgan_b200.plan import Geom, OutMap

DEV = "cuda"
N = int(os.environ.get("N", "64"))


def time_plans(plans, a, out, extra=None, iters=10, stats=None, bias=None):
    runs = []
    for sp, wm in plans:
        b = torch.randn(sp.b_rows * sp.b_k + 64, device=DEV).to(torch.bfloat16) * 0.02
        runs.append((ops.Igemm(sp), b))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    ts = []
    for it in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for g, b in runs:
            if extra is not None:
                g.run(a, extra, out)     # wgrad: (dY, X)
            else:
                g.run(a, b, out, bias, stats)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, macs):
    tf = 2 * macs / (ms * 1e-3) / 1e12
    print("%-34s %8.3f ms  %7.1f TFLOP/s (algorithmic)  %5.1f%% of 1393.9" % (name, ms, tf, 100 * tf / 1393.9))
    return {"name": name, "ms": ms, "tflops": tf}


def main():
    res = []
    H, C = 32, 256
    xg, rg = Geom(N, H, H, C, 1), Geom(N, H, H, C, 0)
    full = Geom(N, H + 2, H + 2, C, 0)
    x = (torch.randn(xg.numel + 512, device=DEV)).to(torch.bfloat16)
    out = torch.zeros(full.numel + 512, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(N, C, 2, device=DEV)
    bias = torch.zeros(C, device=DEV)
    macs = N * H * H * C * C * 9
    shape = (C, C, 3, 3)
    p = CV.conv_fwd_plans(shape, xg, 1, 1, OutMap.nhwc(rg), stats=True, per_sample_stats=True)
    res.append(report("resblock fwd 256->256 3x3 +stats", time_plans(p, x, out, stats=stats, bias=bias), macs))
    p = CV.conv_fwd_plans(shape, xg, 1, 1, OutMap.nhwc(rg))
    res.append(report("resblock fwd (no stats)", time_plans(p, x, out), macs))
    p = CV.conv_dgrad_plans(shape, xg, xg, 1, 1, OutMap.nhwc(full), full_padded=True)
    res.append(report("resblock dgrad (padded grid)", time_plans(p, x, out), macs))
    sp, wm = CV.conv_wgrad_plan(shape, xg, xg, 1, 1)
    packed = torch.zeros(sp.b_rows * sp.b_k, device=DEV)
    res.append(report("resblock wgrad ksplit=%d" % sp.ksplit, time_plans([(sp, wm)], x, packed, extra=x), macs))
    # stem 7x7 4->64 packed at 128x128
    S = 128
    xg0, r1 = Geom(N, S, S, 8, 3), Geom(N, S, S, 64, 0)
    x0 = torch.randn(xg0.numel + 512, device=DEV).to(torch.bfloat16)
    o1 = torch.zeros(r1.numel + 512, dtype=torch.bfloat16, device=DEV)
    p = CV.conv_fwd_plans((64, 4, 7, 7), xg0, 1, 3, OutMap.nhwc(r1))
    res.append(report("stem fwd 7x7 4->64 (packed)", time_plans(p, x0, o1), N * S * S * 64 * 4 * 49))
    # head 7x7 64->3
    xg2 = Geom(N, S, S, 64, 3)
    x2 = torch.randn(xg2.numel + 512, device=DEV).to(torch.bfloat16)
    o3 = torch.zeros(N * 3 * S * S, device=DEV)
    p = CV.conv_fwd_plans((3, 64, 7, 7), xg2, 1, 3, OutMap.nchw(N, 3, S, S), act=L.ACT_TANH)
    res.append(report("head fwd 7x7 64->3 tanh NCHW", time_plans(p, x2, o3), N * S * S * 64 * 3 * 49))
    # down 3x3 s2 64->128 and 128->256
    xa1, r2 = Geom(N, S, S, 64, 1), Geom(N, S // 2, S // 2, 128, 0)
    xa = torch.randn(xa1.numel + 512, device=DEV).to(torch.bfloat16)
    o = torch.zeros(r2.numel + 512, dtype=torch.bfloat16, device=DEV)
    p = CV.conv_fwd_plans((128, 64, 3, 3), xa1, 2, 1, OutMap.nhwc(r2))
    res.append(report("down1 fwd 3x3 s2 64->128", time_plans(p, xa, o), N * (S // 2) ** 2 * 64 * 128 * 9))
    # convT 256->128 at 32 -> 64
    xb, ru = Geom(N, 32, 32, 256, 1), Geom(N, 64, 64, 128, 0)
    xbb = torch.randn(xb.numel + 512, device=DEV).to(torch.bfloat16)
    o = torch.zeros(ru.numel + 512, dtype=torch.bfloat16, device=DEV)
    p = CV.conv_fwd_plans((256, 128, 3, 3), xb, 2, 1, OutMap.nhwc(ru), transposed=True, output_padding=1)
    res.append(report("up1 fwd convT 256->128 (4 phases)", time_plans(p, xbb, o), N * 32 * 32 * 256 * 128 * 9))
    # 8-channel layers: windowed A operand against the overlapping-stride boxes
    from pcgan_b200 import plan as PL
    st1 = torch.zeros(N, 64, 2, device=DEV)
    b64 = torch.zeros(64, device=DEV)
    dyg, xg64 = Geom(N, S, S, 8, 6), Geom(N, S, S, 64, 3)
    fullg = Geom(N, S + 6, S + 6, 64, 0)
    dy8 = torch.randn(dyg.numel + 512, device=DEV).to(torch.bfloat16)
    ofull = torch.zeros(fullg.numel + 512, dtype=torch.bfloat16, device=DEV)
    for flag in (False, True):
        PL.WINDOW = flag
        tag = "window" if flag else "overlap"
        p = CV.conv_fwd_plans((64, 4, 7, 7), xg0, 1, 3, OutMap.nhwc(r1), stats=True, per_sample_stats=True)
        res.append(report("stem fwd +stats [%s]" % tag, time_plans(p, x0, o1, stats=st1, bias=b64), N * S * S * 64 * 4 * 49))
        p = CV.conv_dgrad_plans((3, 64, 7, 7), dyg, xg64, 1, 3, OutMap.nhwc(fullg), full_padded=True)
        res.append(report("head dgrad 3->64 padded [%s]" % tag, time_plans(p, dy8, ofull), N * S * S * 64 * 3 * 49))
    PL.WINDOW = True
    # N = 128 layers: CTA pairs (one cta_group::2 MMA) against unpaired CTAs with two pipelines
    for flag in (True, False):
        PL.PAIRING = flag
        tag = "pair" if flag else "dual"
        p = CV.conv_fwd_plans((128, 64, 3, 3), xa1, 2, 1, OutMap.nhwc(r2))
        res.append(report("down1 fwd 64->128 s2 [%s]" % tag, time_plans(p, xa, o), N * (S // 2) ** 2 * 64 * 128 * 9))
        o_up = torch.zeros(ru.numel + 512, dtype=torch.bfloat16, device=DEV)
        p = CV.conv_fwd_plans((256, 128, 3, 3), xb, 2, 1, OutMap.nhwc(ru), transposed=True, output_padding=1)
        res.append(report("up1 fwd convT 256->128 [%s]" % tag, time_plans(p, xbb, o_up), N * 32 * 32 * 256 * 128 * 9))
        xd, rd = Geom(N, 32, 32, 128, 1), Geom(N, 16, 16, 256, 0)
        xdb = torch.randn(xd.numel + 512, device=DEV).to(torch.bfloat16)
        od = torch.zeros(rd.numel + 512, dtype=torch.bfloat16, device=DEV)
        p = CV.conv_fwd_plans((256, 128, 4, 4), xd, 2, 1, OutMap.nhwc(rd), stats=True)
        res.append(report("D.5 fwd 4x4 s2 128->256 +BN [%s]" % tag, time_plans(p, xdb, od, stats=torch.zeros(1, 256, 2, device=DEV)), N * 16 * 16 * 128 * 256 * 16))
    PL.PAIRING = True
    # stem weight gradient (dY 64 ch x packed window) and an encoder 64->64 convolution with batch statistics
    dy64 = Geom(N, S, S, 64, 0)
    dyb = torch.randn(dy64.numel + 512, device=DEV).to(torch.bfloat16)
    sp, wm = CV.conv_wgrad_plan((64, 4, 7, 7), dy64, xg0, 1, 3)
    packed = torch.zeros(sp.b_rows * sp.b_k, device=DEV)
    res.append(report("stem wgrad 7x7 4->64 ksplit=%d" % sp.ksplit, time_plans([(sp, wm)], dyb, packed, extra=x0), N * S * S * 64 * 4 * 49))
    xe, re_ = Geom(N, 56, 56, 64, 1), Geom(N, 56, 56, 64, 0)
    xeb = torch.randn(xe.numel + 512, device=DEV).to(torch.bfloat16)
    oe = torch.zeros(re_.numel + 512, dtype=torch.bfloat16, device=DEV)
    ste = torch.zeros(1, 64, 2, device=DEV)
    p = CV.conv_fwd_plans((64, 64, 3, 3), xe, 1, 1, OutMap.nhwc(re_), stats=True)
    res.append(report("E.layer1 fwd 3x3 64->64 @56 +BN stats", time_plans(p, xeb, oe, stats=ste), N * 56 * 56 * 64 * 64 * 9))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_conv.json", "w"), indent=1)


if __name__ == "__main__":
    main()
