"""Micro-benchmark of the stream kernels of the normalisation family (norm_apply, norm_bwd_reduce, norm_bwd_apply,
halo_fold) at the generator's shapes, with the arguments the generator passes (pcgan_b200/networks.py _GenProgram).
hot: same buffers back to back (operands L2-resident, as in the step where the producer just wrote them);
cold: a 512 MB write between launches (operands from HBM).  PCGAN_KERNELS_LIB selects another build for A/B."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pcgan_b200 import _lib as L, ops
from pcgan_b200.plan import Geom
from pcgan_b200.engine import zeros_act

DEV = "cuda"


def rnd(g):
    t = zeros_act(g, DEV)
    t[: g.numel] = torch.randn(g.numel, device=DEV).to(torch.bfloat16)
    return t


ONCE = os.environ.get("NB_ONCE") == "1"      # one launch per case (under ncu)


def timeit(name, fn, nbytes, iters=40):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return 0.0, 0.0
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # the launches of one CUDA graph, back to back (the Python call costs more than these kernels run)
    g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    hot = e0.elapsed_time(e1) / iters * 1e3
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)
    cold = []
    for _ in range(8):
        flush.fill_(1)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        cold.append(e0.elapsed_time(e1) * 1e3)
    c = sorted(cold)[len(cold) // 2]
    print("%-44s hot %7.2f us (%6.0f GB/s)   cold %7.2f us (%6.0f GB/s)" % (name, hot, nbytes / hot / 1e3, c, nbytes / c / 1e3), flush=True)
    return hot, c


def main():
    N = int(os.environ.get("NB_N", 64))
    tot = [0.0, 0.0]
    shapes = [tuple(int(v) for v in t.split("x")) for t in os.environ.get("NB_SHAPES", "32x256,64x128,128x64").split(",")]
    for (h, c) in shapes:
        gr, gp = Geom(N, h, h, c, 0), Geom(N, h, h, c, 1)
        el = N * h * h * c * 2
        r, res, y = rnd(gr), rnd(gp), zeros_act(gp, DEV)
        stats = torch.rand(N, c, 2, device=DEV) * h * h
        stats[..., 1] += stats[..., 0] ** 2 / (h * h) + h * h
        mean, rstd, scale, shift = (torch.empty(N, c, device=DEV) for _ in range(4))
        kw = dict(groups=N, stats=stats, count=float(h * h), eps=1e-5, mean_out=mean, rstd_out=rstd, scale_out=scale, shift_out=shift)
        cases = [("apply relu reflect (fused finalize)", lambda: ops.norm_apply(r, gr, y, gp, y_halo=L.HALO_REFLECT, act=L.ACT_RELU, **kw), 2 * el),
                 ("apply relu zero halo", lambda: ops.norm_apply(r, gr, y, gp, y_halo=L.HALO_ZERO, act=L.ACT_RELU, **kw), 2 * el),
                 ("apply + residual, no act, reflect", lambda: ops.norm_apply(r, gr, y, gp, y_halo=L.HALO_REFLECT, act=L.ACT_NONE, res=res, res_pad=1, **kw), 3 * el)]
        ops.norm_apply(r, gr, y, gp, y_halo=L.HALO_REFLECT, act=L.ACT_RELU, **kw)
        sums = torch.zeros(N, c, 2, device=DEV)
        dyp, dy0, dx = rnd(gp), rnd(gr), zeros_act(gp, DEV)
        bk = dict(mean=mean, rstd=rstd, scale=scale, shift=shift, groups=N, count=float(h * h), sums=sums, affine=False)
        cases += [("bwd_reduce relu, folded dy (pad 1)", lambda: ops.norm_bwd_reduce(dyp, 1, r, gr, act=L.ACT_RELU, dy_fold=2, **bk), 2 * el),
                  ("bwd_apply  relu, folded dy (pad 1)", lambda: ops.norm_bwd_apply(dyp, 1, r, gr, act=L.ACT_RELU, dy_fold=2, dx=dx, dx_pad=1, **bk), 3 * el),
                  ("bwd_reduce relu, dy pad 1, halo dropped", lambda: ops.norm_bwd_reduce(dyp, 1, r, gr, act=L.ACT_RELU, dy_fold=1, **bk), 2 * el),
                  ("bwd_apply  relu, dy pad 1, halo dropped", lambda: ops.norm_bwd_apply(dyp, 1, r, gr, act=L.ACT_RELU, dy_fold=1, dx=dx, dx_pad=1, **bk), 3 * el),
                  ("bwd_reduce relu, dy pad 0", lambda: ops.norm_bwd_reduce(dy0, 0, r, gr, act=L.ACT_RELU, **bk), 2 * el),
                  ("bwd_apply  relu, dy pad 0", lambda: ops.norm_bwd_apply(dy0, 0, r, gr, act=L.ACT_RELU, dx=dx, dx_pad=1, **bk), 3 * el),
                  ("bwd_reduce no act, dy pad 0", lambda: ops.norm_bwd_reduce(dy0, 0, r, gr, act=L.ACT_NONE, **bk), 2 * el),
                  ("bwd_apply  no act, dy pad 0", lambda: ops.norm_bwd_apply(dy0, 0, r, gr, act=L.ACT_NONE, dx=dx, dx_pad=1, **bk), 3 * el),
                  ("halo_fold reflect + add", lambda: ops.halo_fold(dyp, gp, dx, 0, halo=L.HALO_REFLECT, add=dy0, add_pad=0), 3 * el)]
        print("--- N=%d %dx%dx%d (%.1f MB per tensor), lib %s" % (N, h, h, c, el / 1e6, os.path.basename(os.path.dirname(L.LIB_PATH)) + "/" + os.path.basename(L.LIB_PATH)))
        for name, fn, nb in cases:
            a, b = timeit(name, fn, nb)
            tot[0] += a
            tot[1] += b
    print("sum of all cases: hot %.1f us, cold %.1f us" % tuple(tot))


if __name__ == "__main__":
    main()
