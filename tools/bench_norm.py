"""Micro-benchmark of the HBM-bound normalisation kernels on the generator's shapes (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcgan_b200 import _lib as L, ops
from pcgan_b200.plan import Geom
from pcgan_b200.engine import NormState

DEV = "cuda"
N = int(os.environ.get("N", "64"))
ITERS = int(os.environ.get("ITERS", "10"))
PEAK = 6467.4


def timeit(fn, iters=ITERS):
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    ts = []
    for it in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print("%-44s %8.3f ms  %7.1f GB/s  %5.1f%% of %.0f" % (name, ms, gbs, 100 * gbs / PEAK, PEAK))


def main():
    for (H, C, pad, halo) in ((32, 256, 1, L.HALO_REFLECT), (128, 64, 1, L.HALO_ZERO), (64, 128, 1, L.HALO_ZERO)):
        gr, gp = Geom(N, H, H, C, 0), Geom(N, H, H, C, pad)
        gfull = Geom(N, H + 2 * pad, H + 2 * pad, C, 0)
        r = (torch.randn(gr.numel + 512, device=DEV)).to(torch.bfloat16)
        g = (torch.randn(gr.numel + 512, device=DEV)).to(torch.bfloat16)
        gpad = (torch.randn(gfull.numel + 512, device=DEV)).to(torch.bfloat16)
        y = torch.zeros(gp.numel + 512, dtype=torch.bfloat16, device=DEV)
        dx = torch.zeros(gp.numel + 512, dtype=torch.bfloat16, device=DEV)
        ns = NormState(N, C, DEV)
        ns.stats.normal_().abs_()
        ns.stats[..., 1] += H * H
        ops.norm_finalize(ns.stats, N, C, H * H, mean=ns.mean, rstd=ns.rstd, scale=ns.scale, shift=ns.shift)
        el = gr.numel
        tag = "%dx%dx%d" % (H, H, C)
        report("norm_apply " + tag, timeit(lambda: ops.norm_apply(r, gr, y, gp, y_halo=halo, scale=ns.scale, shift=ns.shift, groups=N, act=L.ACT_RELU)), el * 2 + gp.numel * 2)
        kw = dict(mean=ns.mean, rstd=ns.rstd, scale=ns.scale, shift=ns.shift, groups=N, act=L.ACT_RELU, count=H * H, sums=ns.sums, affine=0)
        report("norm_bwd_reduce " + tag, timeit(lambda: ops.norm_bwd_reduce(g, 0, r, gr, **kw)), el * 4)
        report("norm_bwd_apply " + tag, timeit(lambda: ops.norm_bwd_apply(g, 0, r, gr, dx=dx, dx_pad=pad, **kw)), el * 6)
        report("norm_bwd_reduce+fold " + tag, timeit(lambda: ops.norm_bwd_reduce(gpad, pad, r, gr, dy_fold=2, **kw)), el * 2 + gfull.numel * 2)
        report("norm_bwd_apply+fold " + tag, timeit(lambda: ops.norm_bwd_apply(gpad, pad, r, gr, dx=dx, dx_pad=pad, dy_fold=2, **kw)), el * 4 + gfull.numel * 2)
        def both():
            ops.norm_bwd_reduce(g, 0, r, gr, **kw)
            ops.norm_bwd_apply(g, 0, r, gr, dx=dx, dx_pad=pad, **kw)
        report("reduce->apply back to back " + tag, timeit(both), el * 10)
        if H == 32:      # the one-launch cluster kernel (off by default in the step: see pcgan_b200/ops.py)
            def fused():
                ops.NORM_FUSED = True
                ops.norm_bwd(g, 0, r, gr, dx=dx, dx_pad=pad, **kw)
                ops.NORM_FUSED = False
            report("norm_bwd one-pass cluster kernel " + tag, timeit(fused), el * 6)
        report("halo_fold(+add) " + tag, timeit(lambda: ops.halo_fold(gpad, gp, dx, 0, halo=L.HALO_REFLECT, add=g, add_pad=0)), el * 4 + gfull.numel * 2)
        report("norm_finalize [%d][%d]" % (N, C), timeit(lambda: ops.norm_finalize(ns.stats, N, C, H * H, mean=ns.mean, rstd=ns.rstd, scale=ns.scale, shift=ns.shift)), N * C * 24)
    a = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    b = torch.empty_like(a)
    report("torch copy 256 MiB (reference)", timeit(lambda: b.copy_(a)), 2 * a.numel())
    print("done")


if __name__ == "__main__":
    main()
