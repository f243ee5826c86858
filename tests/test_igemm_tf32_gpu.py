"""TF32 mode of the implicit-GEMM kernel (tcgen05.mma.kind::tf32 on fp32 NHWC operands, fp32 accumulate and output):
forward, data gradient and weight gradient of the generator's and discriminator's convolution shapes against fp32
torch.nn.functional convolutions of the SAME fp32 operands.  Gate: rel-L2 <= 1e-3 (BASELINE.json's per-layer TF32
tolerance; a 10-bit-mantissa product rounds each operand by 2^-11, measured ~3e-4)."""
import pytest
import torch
import torch.nn.functional as F

from pcgan_b200 import _lib as L
from pcgan_b200 import conv as CV
from pcgan_b200 import ops
from pcgan_b200.plan import Geom, OutMap, SLACK

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-3


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def buf(x, pad, cbuf, halo="zero"):
    """[N, C, H, W] fp32 -> flat padded NHWC fp32 buffer with cbuf channels"""
    n, c, h, w = x.shape
    if pad:
        x = F.pad(x, (pad,) * 4, mode="reflect" if halo == "reflect" else "constant")
    if cbuf > c:
        x = torch.cat([x, x.new_zeros(n, cbuf - c, h + 2 * pad, w + 2 * pad)], 1)
    return torch.cat([x.permute(0, 2, 3, 1).contiguous().reshape(-1), x.new_zeros(SLACK)])


def unbuf(flat, n, h, w, c, pad):
    t = flat[: n * (h + 2 * pad) * (w + 2 * pad) * c].view(n, h + 2 * pad, w + 2 * pad, c)
    return t[:, pad:pad + h, pad:pad + w].permute(0, 3, 1, 2)


def pack(w, wm, numel):
    idx = wm.to(DEV).long()
    flat = w.reshape(-1)
    out = torch.where(idx >= 0, flat[idx.clamp(min=0)], torch.zeros((), device=DEV))
    return torch.cat([out, out.new_zeros(max(numel - out.numel(), 0) + 64)])


def run(plans, a, w, out, bias=None, stats=None):
    for sp, wm in plans:
        assert sp.tf32
        ops.Igemm(sp).run(a, pack(w, wm, sp.b_rows * sp.b_k), out, bias, stats)
    torch.cuda.synchronize()


# name, cin, cbuf, cout, k, stride, cp, halo, xpad, H, N
FWD = [("G.res3x3", 256, 256, 256, 3, 1, 1, "reflect", 1, 32, 8), ("G.down1", 64, 64, 128, 3, 2, 1, "zero", 1, 64, 4),
       ("G.stem7x7", 4, 8, 64, 7, 1, 3, "reflect", 3, 32, 2), ("G.head7x7", 64, 64, 3, 7, 1, 3, "reflect", 3, 32, 2),
       ("D.l0_4x4s2", 4, 8, 64, 4, 2, 1, "zero", 1, 64, 4), ("D.l2_4x4s2", 128, 128, 256, 4, 2, 1, "zero", 1, 32, 4),
       ("D.l3_4x4s1", 256, 256, 512, 4, 1, 1, "zero", 1, 16, 4), ("D.head", 512, 512, 1, 4, 1, 1, "zero", 1, 15, 4)]


@pytest.mark.parametrize("case", FWD, ids=[c[0] for c in FWD])
def test_tf32_forward(case):
    _, cin, cbuf, cout, k, stride, cp, halo, xpad, H, N = case
    torch.manual_seed(0)
    x = torch.randn(N, cin, H, H, device=DEV)
    w = torch.randn(cout, cin, k, k, device=DEV) * 0.05
    bias = torch.randn(cout, device=DEV)
    xg = Geom(N, H, H, cbuf, xpad)
    ho = CV.out_size(H, k, stride, cp)
    og = Geom(N, ho, ho, max(8, -(-cout // 8) * 8), 1)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, stride, cp, OutMap.nhwc(og, dtype=L.DT_F32), stats=True, tf32=True)
    out = torch.zeros(og.numel + SLACK, device=DEV)
    stats = torch.zeros(1, cout, 2, device=DEV)
    run(plans, buf(x, xpad, cbuf, halo), w, out, bias, stats)
    xr = F.pad(x, (cp,) * 4, mode="reflect" if halo == "reflect" else "constant")
    ref = F.conv2d(xr, w, bias, stride=stride)
    e = rel(unbuf(out, N, ho, ho, og.c, 1)[:, :cout], ref)
    print("tf32 fwd %s: %.3e" % (case[0], e))
    assert 1e-6 < e < TOL      # above fp32 round-off: the products really are TF32
    assert rel(stats[0, :, 1], (ref * ref).sum((0, 2, 3))) < 2e-3


def test_tf32_conv_transpose_forward():
    torch.manual_seed(2)
    N, cin, cout, H = 4, 256, 128, 32
    x, w = torch.randn(N, cin, H, H, device=DEV), torch.randn(cin, cout, 3, 3, device=DEV) * 0.05
    xg, og = Geom(N, H, H, cin, 1), Geom(N, 2 * H, 2 * H, cout, 0)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, 2, 1, OutMap.nhwc(og, dtype=L.DT_F32), transposed=True, output_padding=1, tf32=True)
    out = torch.zeros(og.numel + SLACK, device=DEV)
    run(plans, buf(x, 1, cin), w, out)
    ref = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    e = rel(unbuf(out, N, 2 * H, 2 * H, cout, 0), ref)
    print("tf32 convT fwd: %.3e" % e)
    assert e < TOL


# name, cin, cout, cobuf, k, stride, cp, xpad, full, H, N, dypad
DGRAD = [("G.res3x3_full", 256, 256, 256, 3, 1, 1, 1, True, 32, 8, 1), ("G.down1_s2", 64, 128, 128, 3, 2, 1, 1, False, 64, 4, 0),
         ("G.head_to64", 64, 3, 8, 7, 1, 3, 3, True, 32, 2, 6), ("D.l2_s2", 128, 256, 256, 4, 2, 1, 1, False, 32, 4, 1),
         ("D.l3_s1", 256, 512, 512, 4, 1, 1, 1, False, 16, 4, 0), ("G.stem_to4", 4, 64, 64, 7, 1, 3, 3, True, 32, 2, 0)]


@pytest.mark.parametrize("case", DGRAD, ids=[c[0] for c in DGRAD])
def test_tf32_dgrad(case):
    _, cin, cout, cobuf, k, stride, cp, xpad, full, H, N, dypad = case
    torch.manual_seed(3)
    w = torch.randn(cout, cin, k, k, device=DEV) * 0.05
    ho = CV.out_size(H, k, stride, cp)
    dy = torch.randn(N, cout, ho, ho, device=DEV)
    cibuf = max(8, -(-cin // 8) * 8)
    xg = Geom(N, H, H, cibuf, xpad)
    flat_same = stride == 1 and ho == H and cobuf >= 32
    dyg = Geom(N, ho, ho, cobuf, xpad if flat_same else dypad)
    og = Geom(N, H + 2 * xpad, H + 2 * xpad, cibuf, 0) if full else Geom(N, H, H, cibuf, 0)
    plans = CV.conv_dgrad_plans(tuple(w.shape), dyg, xg, stride, cp, OutMap.nhwc(og, dtype=L.DT_F32), full_padded=full, tf32=True)
    out = torch.zeros(og.numel + SLACK, device=DEV)
    run(plans, buf(dy, dyg.pad, cobuf), w, out)
    xp = torch.zeros(N, cin, H + 2 * xpad, H + 2 * xpad, device=DEV, requires_grad=True)
    o = xpad - cp
    xin = xp[:, :, o:H + 2 * xpad - o, o:H + 2 * xpad - o] if o > 0 else xp
    F.conv2d(xin, w, stride=stride).backward(dy)
    ref = xp.grad if full else xp.grad[:, :, xpad:xpad + H, xpad:xpad + H]
    e = rel(unbuf(out, N, og.h, og.w, cibuf, 0)[:, :cin], ref)
    print("tf32 dgrad %s: %.3e" % (case[0], e))
    assert 1e-6 < e < TOL


# name, cin, cbuf, cout, cobuf, k, stride, cp, halo, xpad, H, N, dypad
WGRAD = [("G.res3x3", 256, 256, 256, 256, 3, 1, 1, "reflect", 1, 32, 8, 1), ("G.down1_s2", 64, 64, 128, 128, 3, 2, 1, "zero", 1, 64, 4, 0),
         ("D.l2_s2", 128, 128, 256, 256, 4, 2, 1, "zero", 1, 32, 4, 0), ("G.stem7x7", 4, 8, 64, 64, 7, 1, 3, "reflect", 3, 32, 2, 0),
         ("G.head_cout3", 64, 64, 3, 8, 7, 1, 3, "reflect", 3, 32, 2, 6)]


@pytest.mark.parametrize("case", WGRAD, ids=[c[0] for c in WGRAD])
def test_tf32_wgrad(case):
    _, cin, cbuf, cout, cobuf, k, stride, cp, halo, xpad, H, N, dypad = case
    torch.manual_seed(5)
    x = torch.randn(N, cin, H, H, device=DEV)
    ho = CV.out_size(H, k, stride, cp)
    dy = torch.randn(N, cout, ho, ho, device=DEV)
    xg, dyg = Geom(N, H, H, cbuf, xpad), Geom(N, ho, ho, cobuf, dypad)
    sp, wm = CV.conv_wgrad_plan((cout, cin, k, k), dyg, xg, stride, cp, tf32=True)
    packed = torch.zeros(sp.b_rows * sp.b_k, device=DEV)
    ops.Igemm(sp).run(buf(dy, dypad, cobuf), buf(x, xpad, cbuf, halo), packed)
    dw = torch.zeros(cout * cin * k * k, device=DEV)
    ops.scatter_f32(packed, wm.to(DEV), dw)
    torch.cuda.synchronize()
    xr = F.pad(x, (cp,) * 4, mode="reflect" if halo == "reflect" else "constant")
    wref = torch.zeros(cout, cin, k, k, device=DEV, requires_grad=True)
    F.conv2d(xr.double(), wref.double(), stride=stride).backward(dy.double())
    e = rel(dw.view(cout, cin, k, k), wref.grad)
    print("tf32 wgrad %s: %.3e" % (case[0], e))
    assert 1e-6 < e < TOL
