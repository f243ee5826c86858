"""GPU parity of the tcgen05 implicit-GEMM kernel: the same plans the CPU emulator validates
(tests/test_plan_cpu.py) are launched through the C ABI and compared with fp32 torch convolutions
of the same bf16-rounded operands (the arithmetic the reference dispatches to cuDNN:
models/networks.py:578-605, :747-775).  Tolerance: products of bf16 operands are exact in fp32, so
only the fp32 summation order differs -> rel-L2 <= 2e-5 for fp32 outputs; bf16 outputs add one
rounding (<= 4e-3 relative per element)."""
import pytest
import torch
import torch.nn.functional as F

from oracle.layout import bf16_round, from_padded_nhwc, to_padded_nhwc
from pcgan_b200 import _lib as L
from pcgan_b200 import conv as CV
from pcgan_b200 import ops
from pcgan_b200.plan import Geom, OutMap
from tests.test_plan_cpu import DGRAD_CASES, FWD_CASES, WGRAD_CASES, rel
from tests.test_plan_sweep_cpu import DGRAD_SWEEP, FWD_SWEEP, WGRAD_SWEEP

# the benchmark's shapes plus the seeded sweep of ragged geometries (tests/test_plan_sweep_cpu.py)
FWD_ALL, DGRAD_ALL, WGRAD_ALL = FWD_CASES + FWD_SWEEP, DGRAD_CASES + DGRAD_SWEEP, WGRAD_CASES + WGRAD_SWEEP

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def run_plans(plans, a_flat, w, out, bias=None, stats=None):
    a = a_flat.to(DEV).to(torch.bfloat16)
    wd = w.to(DEV).contiguous()
    keep = []
    for sp, wm in plans:
        b = torch.zeros(sp.b_rows * sp.b_k + 64, dtype=torch.bfloat16, device=DEV)
        ops.gather_cast_bf16(wd, wm.to(DEV), b)
        g = ops.Igemm(sp)
        g.run(a, b, out, bias, stats)
        keep.append((b, g))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("case", FWD_ALL, ids=[c[0] for c in FWD_ALL])
def test_conv_forward(case):
    _, cin, cbuf, cout, k, stride, cp, halo, xpad, H, W, N = case
    torch.manual_seed(0)
    x = torch.randn(N, cin, H, W)
    w = torch.randn(cout, cin, k, k) * 0.1
    bias = torch.randn(cout)
    xg = Geom(N, H, W, cbuf, xpad)
    ho, wo = CV.out_size(H, k, stride, cp), CV.out_size(W, k, stride, cp)
    og = Geom(N, ho, wo, max(8, -(-cout // 8) * 8), 1)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, stride, cp, OutMap.nhwc(og, dtype=L.DT_F32), stats=True)
    out = torch.zeros(og.numel, device=DEV)
    stats = torch.zeros(1, cout, 2, device=DEV)
    run_plans(plans, to_padded_nhwc(x, xpad, halo, cbuf), w, out, bias.to(DEV), stats)
    got = from_padded_nhwc(out.cpu(), N, ho, wo, og.c, 1)[:, :cout]
    xr = F.pad(bf16_round(x), (cp,) * 4, mode="reflect" if halo == "reflect" else "constant")
    ref = F.conv2d(xr, bf16_round(w), bias, stride=stride)
    assert rel(got, ref) < 2e-5
    full = out.cpu()[: og.numel].view(N, ho + 2, wo + 2, og.c)
    assert float(full[:, 0].abs().max()) == 0 and float(full[:, :, 0].abs().max()) == 0
    assert rel(stats.cpu()[0, :, 0], ref.sum((0, 2, 3))) < 1e-4
    assert rel(stats.cpu()[0, :, 1], (ref * ref).sum((0, 2, 3))) < 1e-4


def test_bf16_out_per_sample_stats_nchw_tanh():
    torch.manual_seed(1)
    N, C, H = 3, 64, 16
    x, w = torch.randn(N, C, H, H), torch.randn(64, C, 3, 3) * 0.1
    xg = Geom(N, H, H, C, 1)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, 1, 1, OutMap.nchw(N, 64, H, H), stats=True, per_sample_stats=True, act=L.ACT_TANH)
    stats = torch.zeros(N, 64, 2, device=DEV)
    out = torch.zeros(N * 64 * H * H, device=DEV)
    run_plans(plans, to_padded_nhwc(x, 1, "reflect"), w, out, None, stats)
    pre = F.conv2d(F.pad(bf16_round(x), (1,) * 4, mode="reflect"), bf16_round(w))
    assert rel(out.cpu().view(N, 64, H, H), torch.tanh(pre)) < 2e-5
    assert rel(stats.cpu()[..., 0], pre.sum((2, 3))) < 1e-4
    assert rel(stats.cpu()[..., 1], (pre * pre).sum((2, 3))) < 1e-4
    # bf16 NHWC output of the same conv
    og = Geom(N, H, H, 64, 0)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, 1, 1, OutMap.nhwc(og), act=L.ACT_LRELU, act_slope=0.2)
    outb = torch.zeros(og.numel, dtype=torch.bfloat16, device=DEV)
    run_plans(plans, to_padded_nhwc(x, 1, "reflect"), w, outb)
    assert rel(from_padded_nhwc(outb.float().cpu(), N, H, H, 64, 0), F.leaky_relu(pre, 0.2)) < 4e-3


def test_conv_transpose_forward():
    torch.manual_seed(2)
    N, cin, cout, H = 2, 128, 64, 8
    x, w = torch.randn(N, cin, H, H), torch.randn(cin, cout, 3, 3) * 0.1
    xg, og = Geom(N, H, H, cin, 1), Geom(N, 2 * H, 2 * H, cout, 0)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, 2, 1, OutMap.nhwc(og, dtype=L.DT_F32), transposed=True, output_padding=1,
                              stats=True, per_sample_stats=True)
    stats = torch.zeros(N, cout, 2, device=DEV)
    out = torch.zeros(og.numel, device=DEV)
    run_plans(plans, to_padded_nhwc(x, 1, "zero"), w, out, None, stats)
    ref = F.conv_transpose2d(bf16_round(x), bf16_round(w), stride=2, padding=1, output_padding=1)
    assert rel(from_padded_nhwc(out.cpu(), N, 2 * H, 2 * H, cout, 0), ref) < 2e-5
    assert rel(stats.cpu()[..., 0], ref.sum((2, 3))) < 1e-4


@pytest.mark.parametrize("case", DGRAD_ALL, ids=[c[0] for c in DGRAD_ALL])
def test_conv_dgrad(case):
    _, cin, cout, cobuf, k, stride, cp, xpad, full, H, N, dypad = case
    torch.manual_seed(3)
    w = torch.randn(cout, cin, k, k) * 0.1
    ho = CV.out_size(H, k, stride, cp)
    dy = torch.randn(N, cout, ho, ho)
    cibuf = max(8, -(-cin // 8) * 8)
    xg = Geom(N, H, H, cibuf, xpad)
    flat_same = stride == 1 and ho == H and cobuf >= 64
    dyg = Geom(N, ho, ho, cobuf, xpad if flat_same else dypad)
    og = Geom(N, H + 2 * xpad, H + 2 * xpad, cibuf, 0) if full else Geom(N, H, H, cibuf, 0)
    plans = CV.conv_dgrad_plans(tuple(w.shape), dyg, xg, stride, cp, OutMap.nhwc(og, dtype=L.DT_F32), full_padded=full)
    out = torch.zeros(og.numel, device=DEV)
    run_plans(plans, to_padded_nhwc(dy, dyg.pad, "zero", cobuf), w, out)
    got = from_padded_nhwc(out.cpu(), N, og.h, og.w, cibuf, 0)[:, :cin]
    xp = torch.zeros(N, cin, H + 2 * xpad, H + 2 * xpad, requires_grad=True)
    o = xpad - cp
    xin = xp[:, :, o:H + 2 * xpad - o, o:H + 2 * xpad - o] if o > 0 else xp
    F.conv2d(xin, bf16_round(w), stride=stride).backward(bf16_round(dy))
    ref = xp.grad if full else xp.grad[:, :, xpad:xpad + H, xpad:xpad + H]
    assert rel(got, ref) < 2e-5


def test_conv_transpose_dgrad():
    torch.manual_seed(4)
    N, cin, cout, H = 2, 128, 64, 8
    w = torch.randn(cin, cout, 3, 3) * 0.1
    dy = torch.randn(N, cout, 2 * H, 2 * H)
    dyg, xg, og = Geom(N, 2 * H, 2 * H, cout, 1), Geom(N, H, H, cin, 1), Geom(N, H, H, cin, 0)
    plans = CV.conv_dgrad_plans(tuple(w.shape), dyg, xg, 2, 1, OutMap.nhwc(og, dtype=L.DT_F32), transposed=True)
    out = torch.zeros(og.numel, device=DEV)
    run_plans(plans, to_padded_nhwc(dy, 1, "zero"), w, out)
    x = torch.zeros(N, cin, H, H, requires_grad=True)
    F.conv_transpose2d(x, bf16_round(w), stride=2, padding=1, output_padding=1).backward(bf16_round(dy))
    assert rel(from_padded_nhwc(out.cpu(), N, H, H, cin, 0), x.grad) < 2e-5


def _wgrad(sp, wm, m_flat, n_flat, numel):
    a = m_flat.to(DEV).to(torch.bfloat16)
    b = n_flat.to(DEV).to(torch.bfloat16)
    packed = torch.zeros(sp.b_rows * sp.b_k, device=DEV)
    g = ops.Igemm(sp)
    g.run(a, b, packed)
    dw = torch.zeros(numel, device=DEV)
    ops.scatter_f32(packed, wm.to(DEV), dw)
    torch.cuda.synchronize()
    return dw.cpu()


@pytest.mark.parametrize("case", WGRAD_ALL, ids=[c[0] for c in WGRAD_ALL])
def test_conv_wgrad(case):
    _, cin, cbuf, cout, cobuf, k, stride, cp, halo, xpad, H, N, dypad = case
    torch.manual_seed(5)
    x = torch.randn(N, cin, H, H)
    ho = CV.out_size(H, k, stride, cp)
    dy = torch.randn(N, cout, ho, ho)
    xg, dyg = Geom(N, H, H, cbuf, xpad), Geom(N, ho, ho, cobuf, dypad)
    sp, wm = CV.conv_wgrad_plan((cout, cin, k, k), dyg, xg, stride, cp)
    dyb, xb = to_padded_nhwc(dy, dypad, "zero", cobuf), to_padded_nhwc(x, xpad, halo, cbuf)
    dw = _wgrad(sp, wm, xb, dyb, cout * cin * k * k) if sp.swap_operands else _wgrad(sp, wm, dyb, xb, cout * cin * k * k)
    xr = F.pad(bf16_round(x), (cp,) * 4, mode="reflect" if halo == "reflect" else "constant")
    wref = torch.zeros(cout, cin, k, k, requires_grad=True)
    F.conv2d(xr, wref, stride=stride).backward(bf16_round(dy))
    assert rel(dw.view(cout, cin, k, k), wref.grad) < 2e-5


def test_conv_transpose_wgrad():
    torch.manual_seed(6)
    N, cin, cout, H = 2, 128, 64, 8
    x, dy = torch.randn(N, cin, H, H), torch.randn(N, cout, 2 * H, 2 * H)
    xg, dyg = Geom(N, H, H, cin, 1), Geom(N, 2 * H, 2 * H, cout, 1)
    sp, wm = CV.conv_wgrad_plan((cin, cout, 3, 3), dyg, xg, 2, 1, transposed=True)
    dw = _wgrad(sp, wm, to_padded_nhwc(x, 1, "zero"), to_padded_nhwc(dy, 1, "zero"), cin * cout * 9)
    wref = torch.zeros(cin, cout, 3, 3, requires_grad=True)
    F.conv_transpose2d(bf16_round(x), wref, stride=2, padding=1, output_padding=1).backward(bf16_round(dy))
    assert rel(dw.view(cin, cout, 3, 3), wref.grad) < 2e-5


def test_flagship_shape_resblock_conv():
    """The ResnetBlock convolution at the benchmark shape (networks.py:621-648): 64 x 256 x 32 x 32, bf16 out."""
    torch.manual_seed(7)
    N, C, H = 64, 256, 32
    x = torch.randn(N, C, H, H, device=DEV)
    w = torch.randn(C, C, 3, 3, device=DEV) * 0.02
    xg, og = Geom(N, H, H, C, 1), Geom(N, H, H, C, 0)
    xp = F.pad(x, (1,) * 4, mode="reflect").permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).reshape(-1)
    a = torch.cat([xp, xp.new_zeros(512)])
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, 1, 1, OutMap.nhwc(og), stats=True, per_sample_stats=True)
    (sp, wm), = plans
    b = torch.zeros(sp.b_rows * sp.b_k, dtype=torch.bfloat16, device=DEV)
    ops.gather_cast_bf16(w, wm.to(DEV), b)
    out = torch.zeros(og.numel, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(N, C, 2, device=DEV)
    g = ops.Igemm(sp)
    g.run(a, b, out, None, stats)
    torch.cuda.synchronize()
    ref = F.conv2d(F.pad(x.to(torch.bfloat16).float(), (1,) * 4, mode="reflect"), w.to(torch.bfloat16).float())
    got = out.view(N, H, H, C).permute(0, 3, 1, 2).float()
    assert rel(got, ref) < 4e-3
    assert rel(stats[..., 0], ref.sum((2, 3))) < 1e-3
    assert rel(stats[..., 1], (ref * ref).sum((2, 3))) < 1e-3
