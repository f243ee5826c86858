"""Planner validated on the CPU: every plan is executed by the igemm emulator (oracle/igemm_emulator.py,
which follows the kernel contract in include/pcgan_kernels.h) and compared with torch.nn.functional
convolutions (the arithmetic the reference dispatches: models/networks.py:578-605, :747-775)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import igemm_emulator as emu
from oracle.layout import bf16_round, from_padded_nhwc, to_padded_nhwc
from pcgan_b200 import _lib as L
from pcgan_b200 import conv as CV
from pcgan_b200 import ops
from pcgan_b200.plan import Geom, OutMap


def pack_weights(w, wmap, rows, k):
    flat = bf16_round(w).reshape(-1)
    idx = wmap.long()
    out = torch.where(idx >= 0, flat[idx.clamp(min=0)], torch.zeros(()))
    return out.reshape(rows * k)


def run_fwd(plans, xflat, w, out_numel, bias=None, stats=None, groups=1):
    out = torch.zeros(out_numel)
    for sp, wm in plans:
        ops.Igemm(sp)  # ctypes conversion + the C side's validation of the descriptor (pcgan_igemm_plan_create, host only)
        b = pack_weights(w, wm, sp.b_rows, sp.b_k)
        view = out[sp.out_elem_offset:]
        emu.run_kmajor(sp, xflat[sp.a_elem_offset:], b, view, bias=bias, stats=stats)
    return out


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


FWD_CASES = [
    # name, cin, cin_buf, cout, k, stride, cp, halo, xpad, H, W, N
    ("res3x3_reflect", 64, 64, 64, 3, 1, 1, "reflect", 1, 32, 32, 2),
    ("res3x3_reflect_small", 128, 128, 32, 3, 1, 1, "reflect", 1, 8, 8, 3),
    ("down3x3_s2", 64, 64, 128, 3, 2, 1, "zero", 1, 16, 16, 2),
    ("d4x4_s2", 64, 64, 128, 4, 2, 1, "zero", 1, 16, 16, 2),
    ("d4x4_s1", 64, 64, 80, 4, 1, 1, "zero", 1, 9, 9, 2),
    ("stem7x7_packed", 4, 8, 64, 7, 1, 3, "reflect", 3, 16, 16, 2),
    ("stem7x7_window", 4, 8, 64, 7, 1, 3, "reflect", 3, 6, 64, 2),          # windowed A operand, one row per tile
    ("stem7x7_window_wide", 4, 8, 64, 7, 1, 3, "reflect", 3, 5, 150, 1),    # two tiles per row, the second ragged
    ("dstem4x4_s2_packed", 4, 8, 64, 4, 2, 1, "zero", 1, 16, 16, 2),
    ("estem7x7_s2_packed", 3, 8, 64, 7, 2, 3, "zero", 3, 32, 32, 2),
    ("head7x7_n3", 64, 64, 3, 7, 1, 3, "reflect", 3, 16, 16, 1),
    ("e3x3_14", 64, 64, 64, 3, 1, 1, "zero", 1, 14, 14, 3),
    ("e3x3_7", 128, 128, 64, 3, 1, 1, "zero", 1, 7, 7, 5),
    ("e1x1_s2", 64, 64, 128, 1, 2, 0, "zero", 1, 8, 8, 2),
    ("tail3x3_c32_packed", 32, 32, 1, 3, 1, 1, "zero", 1, 7, 7, 3),
    ("wide_n512", 64, 64, 512, 3, 1, 1, "zero", 1, 8, 8, 1),
    # AlexNet feature extractor of the identity-preserving loss (networks.py:1218-1240)
    ("alex11x11_s4_packed", 3, 8, 64, 11, 4, 2, "zero", 2, 28, 28, 2),
    ("alex5x5_p2_n192", 64, 64, 192, 5, 1, 2, "zero", 2, 11, 11, 2),
    ("alex3x3_c192_n384", 192, 192, 384, 3, 1, 1, "zero", 1, 13, 13, 2),
]


@pytest.mark.parametrize("case", FWD_CASES, ids=[c[0] for c in FWD_CASES])
def test_conv_forward_plan(case):
    _, cin, cbuf, cout, k, stride, cp, halo, xpad, H, W, N = case
    torch.manual_seed(0)
    x = torch.randn(N, cin, H, W)
    w = torch.randn(cout, cin, k, k) * 0.1
    bias = torch.randn(cout)
    xg = Geom(N, H, W, cbuf, xpad)
    ho, wo = CV.out_size(H, k, stride, cp), CV.out_size(W, k, stride, cp)
    og = Geom(N, ho, wo, max(8, -(-cout // 8) * 8), 1)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, stride, cp, OutMap.nhwc(og, dtype=L.DT_F32), stats=True)
    xflat = to_padded_nhwc(x, xpad, halo, cbuf)
    stats = torch.zeros(1, cout, 2)
    out = run_fwd(plans, xflat, w, og.numel, bias=bias, stats=stats)
    got = from_padded_nhwc(out, N, ho, wo, og.c, 1)[:, :cout]
    xr = bf16_round(x)
    if xpad > cp or halo == "reflect":
        xr = F.pad(xr, (cp,) * 4, mode="reflect" if halo == "reflect" else "constant")
        ref = F.conv2d(xr, bf16_round(w), bias, stride=stride)
    else:
        ref = F.conv2d(xr, bf16_round(w), bias, stride=stride, padding=cp)
    assert rel(got, ref) < 1e-5
    # halo of the output buffer untouched
    full = out[: og.numel].view(N, ho + 2, wo + 2, og.c)
    assert float(full[:, 0].abs().max()) == 0 and float(full[:, :, 0].abs().max()) == 0
    assert rel(stats[0, :, 0], ref.sum((0, 2, 3))) < 1e-4
    assert rel(stats[0, :, 1], (ref * ref).sum((0, 2, 3))) < 1e-4


def test_per_sample_stats_and_nchw_out():
    torch.manual_seed(1)
    N, C, H = 3, 64, 16
    x, w = torch.randn(N, C, H, H), torch.randn(64, C, 3, 3) * 0.1
    xg = Geom(N, H, H, C, 1)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, 1, 1, OutMap.nchw(N, 64, H, H), stats=True, per_sample_stats=True, act=L.ACT_TANH)
    stats = torch.zeros(N, 64, 2)
    out = run_fwd(plans, to_padded_nhwc(x, 1, "reflect"), w, N * 64 * H * H, stats=stats)
    pre = F.conv2d(F.pad(bf16_round(x), (1,) * 4, mode="reflect"), bf16_round(w))
    assert rel(out.view(N, 64, H, H), torch.tanh(pre)) < 1e-5
    assert rel(stats[..., 0], pre.sum((2, 3))) < 1e-4
    assert rel(stats[..., 1], (pre * pre).sum((2, 3))) < 1e-4


def test_conv_transpose_forward_plan():
    torch.manual_seed(2)
    N, cin, cout, H = 2, 128, 64, 8
    x, w = torch.randn(N, cin, H, H), torch.randn(cin, cout, 3, 3) * 0.1
    xg = Geom(N, H, H, cin, 1)
    og = Geom(N, 2 * H, 2 * H, cout, 0)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, 2, 1, OutMap.nhwc(og, dtype=L.DT_F32), transposed=True, output_padding=1,
                              stats=True, per_sample_stats=True)
    assert len(plans) == 4 and sorted(len(p[0].tap_off) for p in plans) == [1, 2, 2, 4]
    stats = torch.zeros(N, cout, 2)
    out = run_fwd(plans, to_padded_nhwc(x, 1, "zero"), w, og.numel, stats=stats)
    ref = F.conv_transpose2d(bf16_round(x), bf16_round(w), stride=2, padding=1, output_padding=1)
    assert rel(from_padded_nhwc(out, N, 2 * H, 2 * H, cout, 0), ref) < 1e-5
    assert rel(stats[..., 0], ref.sum((2, 3))) < 1e-4


DGRAD_CASES = [
    # name, cin, cout, cout_buf, k, stride, cp, xpad, full_padded, H, N, dypad
    ("res3x3_flat_full", 64, 64, 64, 3, 1, 1, 1, True, 16, 2, 1),
    ("res3x3_flat_interior", 64, 64, 64, 3, 1, 1, 1, False, 16, 2, 1),
    ("head7x7_packed_full", 64, 3, 8, 7, 1, 3, 3, True, 16, 2, 6),
    ("head7x7_window_flat_full", 64, 3, 8, 7, 1, 3, 3, True, 64, 2, 6),   # windowed A over the flattened padded grid
    ("head7x7_window_interior", 64, 3, 8, 7, 1, 3, 3, False, 64, 1, 3),
    ("dhead4x4_packed", 64, 1, 8, 4, 1, 1, 1, False, 9, 2, 2),
    ("d4x4_s1_box", 64, 64, 64, 4, 1, 1, 1, False, 9, 2, 0),
    ("down3x3_s2", 64, 128, 128, 3, 2, 1, 1, False, 16, 2, 0),
    ("d4x4_s2", 64, 128, 128, 4, 2, 1, 1, False, 16, 2, 1),
    ("dstem4x4_s2_to4", 4, 64, 64, 4, 2, 1, 1, False, 16, 2, 0),
    ("estem7x7_s2_to3", 3, 64, 64, 7, 2, 3, 3, False, 32, 1, 0),
    ("estem7x7_s2_to3_shift", 3, 64, 64, 7, 2, 3, 3, False, 32, 2, 2),      # zero-haloed dY: shift-sum phases
    ("e1x1_s2", 64, 128, 128, 1, 2, 0, 1, False, 8, 2, 0),
    ("stem7x7_to4_full", 4, 64, 64, 7, 1, 3, 3, True, 16, 1, 0),
    ("alex11x11_s4_to3", 3, 64, 64, 11, 4, 2, 2, False, 28, 2, 2),         # 16 phases of a stride-4 convolution
    ("alex5x5_p2_flat", 64, 192, 192, 5, 1, 2, 2, False, 11, 2, 2),
    ("alex3x3_n384_flat", 192, 384, 384, 3, 1, 1, 1, False, 13, 2, 1),
]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=[c[0] for c in DGRAD_CASES])
def test_conv_dgrad_plan(case):
    _, cin, cout, cobuf, k, stride, cp, xpad, full, H, N, dypad = case
    torch.manual_seed(3)
    w = torch.randn(cout, cin, k, k) * 0.1
    ho = CV.out_size(H, k, stride, cp)
    dy = torch.randn(N, cout, ho, ho)
    cibuf = max(8, -(-cin // 8) * 8)
    xg = Geom(N, H, H, cibuf, xpad)
    flat_same = stride == 1 and ho == H and cobuf >= 64
    dyg = Geom(N, ho, ho, cobuf, xpad if flat_same else dypad)
    if full:
        og = Geom(N, H + 2 * xpad, H + 2 * xpad, cibuf, 0)
    else:
        og = Geom(N, H, H, cibuf, 0)
    plans = CV.conv_dgrad_plans(tuple(w.shape), dyg, xg, stride, cp, OutMap.nhwc(og, dtype=L.DT_F32), full_padded=full)
    out = run_fwd(plans, to_padded_nhwc(dy, dyg.pad, "zero", cobuf), w, og.numel)
    got = from_padded_nhwc(out, N, og.h, og.w, cibuf, 0)[:, :cin]
    # reference: autograd of the padded-input convolution
    xp = torch.zeros(N, cin, H + 2 * xpad, H + 2 * xpad, requires_grad=True)
    o = xpad - cp
    xin = xp[:, :, o:H + 2 * xpad - o, o:H + 2 * xpad - o] if o > 0 else xp
    y = F.conv2d(xin, bf16_round(w), stride=stride)
    y.backward(bf16_round(dy))
    ref = xp.grad if full else xp.grad[:, :, xpad:xpad + H, xpad:xpad + H]
    assert rel(got, ref) < 1e-5


def test_conv_transpose_dgrad_plan():
    torch.manual_seed(4)
    N, cin, cout, H = 2, 128, 64, 8
    w = torch.randn(cin, cout, 3, 3) * 0.1
    dy = torch.randn(N, cout, 2 * H, 2 * H)
    dyg, xg = Geom(N, 2 * H, 2 * H, cout, 1), Geom(N, H, H, cin, 1)
    og = Geom(N, H, H, cin, 0)
    plans = CV.conv_dgrad_plans(tuple(w.shape), dyg, xg, 2, 1, OutMap.nhwc(og, dtype=L.DT_F32), transposed=True)
    out = run_fwd(plans, to_padded_nhwc(dy, 1, "zero"), w, og.numel)
    x = torch.zeros(N, cin, H, H, requires_grad=True)
    F.conv_transpose2d(x, bf16_round(w), stride=2, padding=1, output_padding=1).backward(bf16_round(dy))
    assert rel(from_padded_nhwc(out, N, H, H, cin, 0), x.grad) < 1e-5


WGRAD_CASES = [
    # name, cin, cin_buf, cout, cout_buf, k, stride, cp, halo, xpad, H, N, dypad
    ("res3x3", 64, 64, 64, 64, 3, 1, 1, "reflect", 1, 16, 2, 1),
    ("res3x3_c256", 256, 256, 128, 128, 3, 1, 1, "reflect", 1, 8, 2, 0),
    ("down3x3_s2", 64, 64, 192, 192, 3, 2, 1, "zero", 1, 16, 2, 0),
    ("d4x4_s2", 64, 64, 128, 128, 4, 2, 1, "zero", 1, 16, 3, 1),
    ("d4x4_s1_odd", 64, 64, 64, 64, 4, 1, 1, "zero", 1, 9, 2, 1),
    ("stem7x7_packed", 4, 8, 64, 64, 7, 1, 3, "reflect", 3, 16, 2, 0),
    ("dstem4x4_s2_packed", 4, 8, 64, 64, 4, 2, 1, "zero", 1, 16, 2, 0),
    ("head7x7_cout3", 64, 64, 3, 8, 7, 1, 3, "reflect", 3, 16, 2, 6),
    ("dhead4x4_cout1", 64, 64, 1, 8, 4, 1, 1, "zero", 1, 9, 2, 2),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
def test_conv_wgrad_plan(case):
    _, cin, cbuf, cout, cobuf, k, stride, cp, halo, xpad, H, N, dypad = case
    torch.manual_seed(5)
    x = torch.randn(N, cin, H, H)
    ho = CV.out_size(H, k, stride, cp)
    dy = torch.randn(N, cout, ho, ho)
    xg, dyg = Geom(N, H, H, cbuf, xpad), Geom(N, ho, ho, cobuf, dypad)
    sp, wm = CV.conv_wgrad_plan((cout, cin, k, k), dyg, xg, stride, cp)
    ops.Igemm(sp)   # the C side accepts the descriptor
    packed = torch.zeros(sp.b_rows * sp.b_k)
    dyb, xb = to_padded_nhwc(dy, dypad, "zero", cobuf), to_padded_nhwc(x, xpad, halo, cbuf)
    if sp.swap_operands:   # few output channels: activations on the M side, dY (packed window) on the N side
        emu.run_wgrad(sp, xb[sp.a_elem_offset:], dyb[sp.b_elem_offset:], packed)
    else:
        emu.run_wgrad(sp, dyb[sp.a_elem_offset:], xb[sp.b_elem_offset:], packed)
    dw = torch.zeros(cout * cin * k * k)
    idx = wm.long()
    dw[idx[idx >= 0]] = packed[idx >= 0]
    xr = F.pad(bf16_round(x), (cp,) * 4, mode="reflect" if halo == "reflect" else "constant")
    wref = torch.zeros(cout, cin, k, k, requires_grad=True)
    F.conv2d(xr, wref, stride=stride).backward(bf16_round(dy))
    assert rel(dw.view(cout, cin, k, k), wref.grad) < 1e-5
    # every packed element that maps nowhere must be exactly zero-weight in forward; nothing to check here


def test_conv_transpose_wgrad_plan():
    torch.manual_seed(6)
    N, cin, cout, H = 2, 128, 64, 8
    x, dy = torch.randn(N, cin, H, H), torch.randn(N, cout, 2 * H, 2 * H)
    xg, dyg = Geom(N, H, H, cin, 1), Geom(N, 2 * H, 2 * H, cout, 1)
    sp, wm = CV.conv_wgrad_plan((cin, cout, 3, 3), dyg, xg, 2, 1, transposed=True)
    packed = torch.zeros(sp.b_rows * sp.b_k)
    emu.run_wgrad(sp, to_padded_nhwc(x, 1, "zero"), to_padded_nhwc(dy, 1, "zero"), packed)
    dw = torch.zeros(cin * cout * 9)
    idx = wm.long()
    dw[idx[idx >= 0]] = packed[idx >= 0]
    wref = torch.zeros(cin, cout, 3, 3, requires_grad=True)
    F.conv_transpose2d(bf16_round(x), wref, stride=2, padding=1, output_padding=1).backward(bf16_round(dy))
    assert rel(dw.view(cin, cout, 3, 3), wref.grad) < 1e-5


def test_plan_create_rejects_inconsistent_descriptors():
    """pcgan_igemm_plan_create validates on the host (no GPU): a descriptor the kernel cannot run is refused with a
    message, never launched."""
    from pcgan_b200._lib import PcganError

    def stem():
        xg, og = Geom(2, 6, 64, 8, 3), Geom(2, 6, 64, 64, 1)
        return CV.conv_fwd_plans((64, 4, 7, 7), xg, 1, 3, OutMap.nhwc(og))[0][0]

    def res():
        xg, og = Geom(2, 16, 16, 256, 1), Geom(2, 16, 16, 256, 0)
        return CV.conv_fwd_plans((256, 256, 3, 3), xg, 1, 1, OutMap.nhwc(og))[0][0]

    sp = stem()
    assert sp.a_window == 8 and sp.pair == 0
    ops.Igemm(sp)
    for field, value in (("pair", 1), ("cchunks", 2), ("block_n", 20), ("ksplit", 2), ("a_window", 4), ("wg_box_dim", 2), ("n_valid", 0)):
        bad = stem()
        setattr(bad, field, value)
        with pytest.raises(PcganError):
            ops.Igemm(bad)
    sp = res()
    assert sp.pair == 1 and sp.block_n == 256
    ops.Igemm(sp)
    for field, value in (("pair", 2), ("shift_taps", 3), ("stats_dim", 7)):
        bad = res()
        setattr(bad, field, value)
        if field == "stats_dim":
            bad.stats_mode = L.STATS_ON
        with pytest.raises(PcganError):
            ops.Igemm(bad)
    sp, _ = CV.conv_wgrad_plan((64, 4, 7, 7), Geom(2, 16, 16, 64, 0), Geom(2, 16, 16, 8, 3), 1, 3)
    assert sp.wg_box_dim == 2 and sp.num_taps == 1
    ops.Igemm(sp)
    sp.wg_box_dim = 5
    with pytest.raises(PcganError):
        ops.Igemm(sp)



# ------------------------------------------------------------------------------------------- TF32 plans
# The same planner with tf32=True (fp32 operands, K chunks of 32 elements, 64 x 32 weight-gradient boxes): the emulator
# reads the descriptors with 4-byte elements.  (Arithmetic here is exact fp32 either way: this validates the planning.)
TF32_FWD = [c for c in FWD_CASES if c[0] in ("res3x3_reflect", "down3x3_s2", "d4x4_s2", "d4x4_s1", "stem7x7_packed", "dstem4x4_s2_packed",
                                             "head7x7_n3", "e3x3_7", "e1x1_s2", "wide_n512")]


@pytest.mark.parametrize("case", TF32_FWD, ids=[c[0] for c in TF32_FWD])
def test_conv_forward_plan_tf32(case):
    _, cin, cbuf, cout, k, stride, cp, halo, xpad, H, W, N = case
    torch.manual_seed(0)
    x, w, bias = torch.randn(N, cin, H, W), torch.randn(cout, cin, k, k) * 0.1, torch.randn(cout)
    xg = Geom(N, H, W, cbuf, xpad)
    ho, wo = CV.out_size(H, k, stride, cp), CV.out_size(W, k, stride, cp)
    og = Geom(N, ho, wo, max(8, -(-cout // 8) * 8), 1)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, stride, cp, OutMap.nhwc(og, dtype=L.DT_F32), stats=True, tf32=True)
    assert all(sp.tf32 and sp.a_box[0] == 32 and not sp.pair and not sp.a_window for sp, _ in plans)
    stats = torch.zeros(1, cout, 2)
    out = run_fwd(plans, to_padded_nhwc(x, xpad, halo, cbuf), w, og.numel, bias=bias, stats=stats)
    got = from_padded_nhwc(out, N, ho, wo, og.c, 1)[:, :cout]
    xr = F.pad(bf16_round(x), (cp,) * 4, mode="reflect" if halo == "reflect" else "constant")
    ref = F.conv2d(xr, bf16_round(w), bias, stride=stride)
    assert rel(got, ref) < 1e-5
    assert rel(stats[0, :, 1], (ref * ref).sum((0, 2, 3))) < 1e-4


def test_conv_transpose_forward_plan_tf32():
    torch.manual_seed(2)
    N, cin, cout, H = 2, 128, 64, 8
    x, w = torch.randn(N, cin, H, H), torch.randn(cin, cout, 3, 3) * 0.1
    xg, og = Geom(N, H, H, cin, 1), Geom(N, 2 * H, 2 * H, cout, 0)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, 2, 1, OutMap.nhwc(og, dtype=L.DT_F32), transposed=True, output_padding=1, tf32=True)
    out = run_fwd(plans, to_padded_nhwc(x, 1, "zero"), w, og.numel)
    ref = F.conv_transpose2d(bf16_round(x), bf16_round(w), stride=2, padding=1, output_padding=1)
    assert rel(from_padded_nhwc(out, N, 2 * H, 2 * H, cout, 0), ref) < 1e-5


TF32_DGRAD = [c for c in DGRAD_CASES if c[0] in ("res3x3_flat_full", "res3x3_flat_interior", "head7x7_packed_full", "d4x4_s1_box", "down3x3_s2",
                                                 "d4x4_s2", "dstem4x4_s2_to4", "stem7x7_to4_full")]


@pytest.mark.parametrize("case", TF32_DGRAD, ids=[c[0] for c in TF32_DGRAD])
def test_conv_dgrad_plan_tf32(case):
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
    plans = CV.conv_dgrad_plans(tuple(w.shape), dyg, xg, stride, cp, OutMap.nhwc(og, dtype=L.DT_F32), full_padded=full, tf32=True)
    assert all(sp.tf32 for sp, _ in plans)
    out = run_fwd(plans, to_padded_nhwc(dy, dyg.pad, "zero", cobuf), w, og.numel)
    got = from_padded_nhwc(out, N, og.h, og.w, cibuf, 0)[:, :cin]
    xp = torch.zeros(N, cin, H + 2 * xpad, H + 2 * xpad, requires_grad=True)
    o = xpad - cp
    xin = xp[:, :, o:H + 2 * xpad - o, o:H + 2 * xpad - o] if o > 0 else xp
    F.conv2d(xin, bf16_round(w), stride=stride).backward(bf16_round(dy))
    ref = xp.grad if full else xp.grad[:, :, xpad:xpad + H, xpad:xpad + H]
    assert rel(got, ref) < 1e-5


TF32_WGRAD = [c for c in WGRAD_CASES if c[0] in ("res3x3", "res3x3_c256", "down3x3_s2", "d4x4_s2", "d4x4_s1_odd", "stem7x7_packed", "head7x7_cout3")]


@pytest.mark.parametrize("case", TF32_WGRAD, ids=[c[0] for c in TF32_WGRAD])
def test_conv_wgrad_plan_tf32(case):
    _, cin, cbuf, cout, cobuf, k, stride, cp, halo, xpad, H, N, dypad = case
    torch.manual_seed(5)
    x = torch.randn(N, cin, H, H)
    ho = CV.out_size(H, k, stride, cp)
    dy = torch.randn(N, cout, ho, ho)
    xg, dyg = Geom(N, H, H, cbuf, xpad), Geom(N, ho, ho, cobuf, dypad)
    sp, wm = CV.conv_wgrad_plan((cout, cin, k, k), dyg, xg, stride, cp, tf32=True)
    assert sp.tf32 and not sp.pair and not sp.swap_operands and sp.a_box[0] == 32
    ops.Igemm(sp)
    packed = torch.zeros(sp.b_rows * sp.b_k)
    dyb, xb = to_padded_nhwc(dy, dypad, "zero", cobuf), to_padded_nhwc(x, xpad, halo, cbuf)
    emu.run_wgrad(sp, dyb[sp.a_elem_offset:], xb[sp.b_elem_offset:], packed)
    dw = torch.zeros(cout * cin * k * k)
    idx = wm.long()
    dw[idx[idx >= 0]] = packed[idx >= 0]
    xr = F.pad(bf16_round(x), (cp,) * 4, mode="reflect" if halo == "reflect" else "constant")
    wref = torch.zeros(cout, cin, k, k, requires_grad=True)
    F.conv2d(xr, wref, stride=stride).backward(bf16_round(dy))
    assert rel(dw.view(cout, cin, k, k), wref.grad) < 1e-5


# ------------------------------------------------------------------------------------- UnetGenerator shapes
@pytest.mark.parametrize("cin,cout,H,N", [(128, 3, 8, 2), (512, 512, 1, 3), (256, 64, 4, 2)], ids=["outermost_128to3", "innermost_1x1", "mid"])
def test_unet_conv_transpose_4x4_plans(cin, cout, H, N):
    """nn.ConvTranspose2d(cin, cout, 4, 2, 1) (networks.py:699-713): forward by sub-pixel phases, data gradient (for few
    output channels: packed filter rows of the 8-channel dY) and weight gradient, down to a 1 x 1 input."""
    torch.manual_seed(8)
    x, w = torch.randn(N, cin, H, H), torch.randn(cin, cout, 4, 4) * 0.1
    cob = max(8, -(-cout // 8) * 8)
    xg, og = Geom(N, H, H, cin, 1), Geom(N, 2 * H, 2 * H, cob, 0)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, 2, 1, OutMap.nhwc(og, dtype=L.DT_F32), transposed=True)
    out = run_fwd(plans, to_padded_nhwc(x, 1, "zero"), w, og.numel)
    ref = F.conv_transpose2d(bf16_round(x), bf16_round(w), stride=2, padding=1)
    assert rel(from_padded_nhwc(out, N, 2 * H, 2 * H, cob, 0)[:, :cout], ref) < 1e-5
    dy = torch.randn(N, cout, 2 * H, 2 * H)
    dyg, dxg = Geom(N, 2 * H, 2 * H, cob, 1), Geom(N, H, H, cin, 0)
    plans = CV.conv_dgrad_plans(tuple(w.shape), dyg, xg, 2, 1, OutMap.nhwc(dxg, dtype=L.DT_F32), transposed=True)
    out = run_fwd(plans, to_padded_nhwc(dy, 1, "zero", cob), w, dxg.numel)
    xr = torch.zeros(N, cin, H, H, requires_grad=True)
    F.conv_transpose2d(xr, bf16_round(w), stride=2, padding=1).backward(bf16_round(dy))
    assert rel(from_padded_nhwc(out, N, H, H, cin, 0), xr.grad) < 1e-5
    # weight gradient: M side = the input activations, N side = dY
    sp, wm = CV.conv_wgrad_plan(tuple(w.shape), dyg, xg, 2, 1, transposed=True)
    ops.Igemm(sp)
    packed = torch.zeros(sp.b_rows * sp.b_k)
    emu.run_wgrad(sp, to_padded_nhwc(x, 1, "zero")[sp.a_elem_offset:], to_padded_nhwc(dy, 1, "zero", cob)[sp.b_elem_offset:], packed)
    dw = torch.zeros(w.numel())
    idx = wm.long()
    dw[idx[idx >= 0]] = packed[idx >= 0]
    wr = torch.zeros_like(w, requires_grad=True)
    F.conv_transpose2d(bf16_round(x), wr, stride=2, padding=1).backward(bf16_round(dy))
    assert rel(dw.view_as(w), wr.grad) < 1e-5


def test_unet_down_conv_to_1x1_plan():
    torch.manual_seed(9)
    N, C = 3, 128
    x, w, b = torch.randn(N, C, 2, 2), torch.randn(64, C, 4, 4) * 0.1, torch.randn(64)
    xg, og = Geom(N, 2, 2, C, 1), Geom(N, 1, 1, 64, 1)
    plans = CV.conv_fwd_plans(tuple(w.shape), xg, 2, 1, OutMap.nhwc(og, dtype=L.DT_F32), act=L.ACT_RELU)
    out = run_fwd(plans, to_padded_nhwc(x, 1, "zero"), w, og.numel, bias=b)
    ref = torch.relu(F.conv2d(bf16_round(x), bf16_round(w), b, stride=2, padding=1))
    assert rel(from_padded_nhwc(out, N, 1, 1, 64, 1), ref) < 1e-5


def test_magic_number_divisions_are_exact():
    """The kernel divides tile indices and output coordinates by per-plan constants with (umulhi(x, mul) + x) >> shift
    (csrc/igemm.cu FastDiv); the library's host-only self-test evaluates the same formula against x / d for plan-like,
    power-of-two, neighbouring and random divisors up to 2^31 - 1 and dividends up to 2^31 - 1."""
    lib = L.load()
    for seed in (1, 2, 20260):
        assert lib.pcgan_selftest_fastdiv(seed, 200000) == 0
