"""Convolution -> implicit-GEMM plans (forward, data-gradient, weight-gradient) for
nn.Conv2d and nn.ConvTranspose2d as the reference uses them
(models/networks.py:578-605, :621-648, :747-775, :1014-1027; models/resnet.py:20-28,134).

Pure planning (no GPU): every function returns a list of (IgemmSpec, weight_index_map)
pairs; one pair per launch (strided data-gradients and transposed convolutions run one
launch per sub-pixel phase).  The index map gathers the reference-layout fp32 weight
(OIHW, or IOHW for ConvTranspose2d) into the packed bf16 operand of that launch; for a
weight-gradient plan the same kind of map scatters the packed fp32 result back.

Buffer rules (all NHWC bf16, physically padded, see plan.Geom):
  * the input of a convolution with padding cp lives in a buffer with pad >= cp whose halo
    was written by its producer (zeros for zero padding, reflection for ReflectionPad2d);
  * an output-gradient dY lives in a zero-haloed buffer Geom(n, Ho, Wo, Cout_buf, pad>=0);
    for the "flat" data-gradient of a same-size stride-1 convolution it must have the
    geometry of the convolution's input (pad == input pad);
  * channel counts in buffers are multiples of 8; a buffer with fewer than 64 channels is
    read through the packed-row path.
"""
from typing import List, Tuple

import torch

from . import _lib as L
from .plan import (ANY, ONE, Geom, IgemmSpec, OutMap, plan_box, plan_flat, plan_packed, plan_wgrad_box,
                   plan_shift_flat, plan_wgrad_small_cout, wmap_packed, wmap_shift, wmap_small_cout, wmap_taps, _ceil)


def out_size(h, k, stride, cp, transposed=False, output_padding=0):
    if transposed:
        return (h - 1) * stride - 2 * cp + k + output_padding
    return (h + 2 * cp - k) // stride + 1


# --------------------------------------------------------------------------------- forward
def conv_fwd_plans(w_shape, xg: Geom, stride: int, cp: int, out: OutMap, *, transposed=False, output_padding=0,
                   act=L.ACT_NONE, act_slope=0.0, stats=False, per_sample_stats=False, note="", tf32=False) -> List[Tuple[IgemmSpec, torch.Tensor]]:
    """Forward plans.  w_shape: reference weight shape (OIHW; IOHW when transposed).  tf32: the buffers and the packed
    weights hold fp32 and the MMAs run as kind::tf32 (K chunks of 32 channels; no shift-sum / windowed special forms)."""
    kh, kw = w_shape[2], w_shape[3]
    assert kh == kw
    k = kh
    epi = dict(act=act, act_slope=act_slope, stats=stats, per_sample_stats=per_sample_stats, note=note, tf32=tf32)
    KC = 32 if tf32 else 64
    if not transposed:
        cout, cin = w_shape[0], w_shape[1]
        ho, wo = out_size(xg.h, k, stride, cp), out_size(xg.w, k, stride, cp)
        o = xg.pad - cp
        assert o >= 0, "input buffer pad %d < conv padding %d" % (xg.pad, cp)
        if stride == 1 and xg.c == cin and cin % 64 == 0 and cout <= 4 and 2 <= k <= 8 and not stats and not tf32:
            # few output channels (generator head, networks.py:603-605): horizontal taps in N, shift-sum epilogue
            sp = plan_shift_flat(xg, k, k, cin, cout, [(r + o, o, r) for r in range(k)], out, (o, o + ho), (o, o + wo),
                                 act=act, act_slope=act_slope, note=note)
            return [(sp, wmap_shift(w_shape, k, cin))]
        if xg.c >= 64 or (tf32 and xg.c == cin and cin % KC == 0):
            assert xg.c == cin
            taps = [(r + o, s + o, r * k + s) for r in range(k) for s in range(k)]
            sp = plan_box(xg, taps, cin, cout, ho, wo, stride, out, **epi)
            return [(sp, wmap_taps(w_shape, cout, [(r, s, r * k + s) for r in range(k) for s in range(k)], cin))]
        assert cin <= xg.c
        sp = plan_packed(xg, k, k, stride, o, cout, ho, wo, out, **epi)
        return [(sp, wmap_packed(w_shape, cout, k, k, xg.c, sp.b_k // k))]
    # ConvTranspose2d, stride 2: output phase (py, px) gathers the taps r with (py + cp - r) even,
    # reading input row u + (py + cp - r) / 2 for output row 2u + py.
    assert stride == 2 and xg.c >= 64
    cin, cout = w_shape[0], w_shape[1]
    assert xg.c == cin
    ho, wo = out_size(xg.h, k, 2, cp, True, output_padding), out_size(xg.w, k, 2, cp, True, output_padding)
    plans = []
    for py in range(2):
        for px in range(2):
            taps, wt = [], []
            for r in range(k):
                if (py + cp - r) % 2:
                    continue
                for s in range(k):
                    if (px + cp - s) % 2:
                        continue
                    kidx = len(taps)
                    taps.append(((py + cp - r) // 2 + xg.pad, (px + cp - s) // 2 + xg.pad, kidx))
                    wt.append((r, s, kidx))
            hph, wph = _ceil(ho - py, 2), _ceil(wo - px, 2)
            om = OutMap(base=out.base + py * out.sy + px * out.sx, sn=out.sn, sy=2 * out.sy, sx=2 * out.sx, sc=out.sc, dtype=out.dtype)
            sp = plan_box(xg, taps, cin, cout, hph, wph, 1, om, **epi)
            plans.append((sp, wmap_taps(w_shape, cout, wt, cin, transposed_layout=True, swap=False)))
    return plans


# --------------------------------------------------------------------------- data gradient
def conv_dgrad_plans(w_shape, dyg: Geom, xg: Geom, stride: int, cp: int, out: OutMap, *, transposed=False,
                     full_padded=False, note="", tf32=False) -> List[Tuple[IgemmSpec, torch.Tensor]]:
    """Data-gradient plans.  dyg: geometry of the dY buffer (zero halo).  xg: geometry of the convolution's
    input buffer (used for sizes; channels = buffer channels of dX).  `out` maps pixel (n, y, x) of the
    computed region: the interior of the input (full_padded=False) or its whole padded grid
    (full_padded=True, reflect-padded inputs; fold the halo afterwards with pcgan_halo_fold)."""
    kh, kw = w_shape[2], w_shape[3]
    k = kh
    T = dict(note=note, tf32=tf32)
    if transposed:
        # dX[iy][ix][ci] = sum dY[2iy - cp + r][2ix - cp + s][co] W[ci][co][r][s]: a stride-2 convolution of dY
        cin, cout = w_shape[0], w_shape[1]
        assert stride == 2 and not full_padded
        o = dyg.pad - cp
        assert o >= 0
        if dyg.c < 64:
            # few output channels (the outermost up convolution of UnetGenerator, networks.py:699-701): dY is an 8-channel
            # buffer read as packed filter rows, exactly like the forward of a stride-2 convolution over an image
            assert cout <= dyg.c and not tf32
            sp = plan_packed(dyg, k, k, 2, o, cin, xg.h, xg.w, out, note=note)
            return [(sp, wmap_packed(w_shape, cin, k, k, dyg.c, sp.b_k // k))]
        assert dyg.c == cout and cout % 64 == 0
        taps = [(r + o, s + o, r * k + s) for r in range(k) for s in range(k)]
        sp = plan_box(dyg, taps, cout, cin, xg.h, xg.w, 2, out, **T)
        wt = [(r, s, r * k + s) for r in range(k) for s in range(k)]
        return [(sp, wmap_taps(w_shape, cin, wt, cout, transposed_layout=True, swap=True))]
    cout, cin = w_shape[0], w_shape[1]
    o = xg.pad - cp
    if stride == 1:
        # dXp[u][v] = sum dY[u - r - o][v - s - o] W[co][ci][r][s]   (u, v in padded coordinates of X)
        if full_padded:
            u0, v0, hu, wu = 0, 0, xg.hp, xg.wp
        else:
            u0, v0, hu, wu = xg.pad, xg.pad, xg.h, xg.w
        same = (dyg.h, dyg.w, dyg.pad) == (xg.h, xg.w, xg.pad)
        if dyg.c >= 64 or (tf32 and dyg.c == cout and cout % 32 == 0):
            assert dyg.c == cout
            wt = [(r, s, r * k + s) for r in range(k) for s in range(k)]
            wm = wmap_taps(w_shape, cin, wt, cout, swap=True)
            if same and 2 * cp == k - 1 and cin <= 4 and xg.c == 8 and 2 <= k <= 8 and out.sc == 1 and not tf32:
                # gradient towards a 3/4-channel image (generator stem, networks.py:578-579): shift-sum form over the
                # shared padded grid; row t of the GEMM feeds output position t + (k - 1 - cp)
                sp = plan_shift_flat(dyg, k, k, cout, cin, [(-(r - cp), 0, r) for r in range(k)], out, (u0, u0 + hu), (v0, v0 + wu),
                                     out_shift=k - 1 - cp, note=note)
                return [(sp, wmap_shift(w_shape, k, cout, dgrad=True))]
            if same and 2 * cp == k - 1:
                # flat form over the shared padded grid: dY pixel (y, x) sits at padded (y + pad, x + pad)
                taps = [(-(r - cp), -(s - cp), r * k + s) for r in range(k) for s in range(k)]
                sp = plan_flat(dyg, taps, cout, cin, out, (u0, u0 + hu), (v0, v0 + wu), **T)
                return [(sp, wm)]
            taps = [(u0 - r - o + dyg.pad, v0 - s - o + dyg.pad, r * k + s) for r in range(k) for s in range(k)]
            sp = plan_box(dyg, taps, cout, cin, hu, wu, 1, out, **T)
            return [(sp, wm)]
        # packed: window j <-> s = k-1-j, row tap i <-> r = k-1-i; needs real zero padding around dY
        need = (k - 1 + o) if full_padded else (k - 1 - cp)
        assert dyg.pad >= need, "packed data-gradient needs dY pad >= %d (got %d)" % (need, dyg.pad)
        off = u0 - (k - 1) - o + dyg.pad
        sp = plan_packed(dyg, k, k, 1, off, cin, hu, wu, out, **T)
        return [(sp, wmap_packed(w_shape, cin, k, k, dyg.c, sp.b_k // k, flip=True, swap=True))]
    # stride st: input row iy = st*u + py receives the taps r with (py + cp - r) a multiple of st from dY row u + (py + cp - r)/st
    st = stride
    assert st in (2, 4) and (dyg.c >= 64 or tf32) and dyg.c == cout and not full_padded
    plans = []
    reach = max(abs((p_ + cp - r) // st) for p_ in range(st) for r in range(k) if (p_ + cp - r) % st == 0) if k > 1 else 0
    if st == 2 and cin <= 4 and xg.c == 8 and 2 <= k <= 8 and dyg.pad >= max(reach, 1) and out.sc == 1 and not tf32:
        # gradient towards a 3/4-channel image through a strided convolution (encoder stem, resnet.py:134): one
        # shift-sum launch per sub-pixel phase over the zero-haloed dY grid
        P = dyg.pad
        for py in range(2):
            for px in range(2):
                rows = [r for r in range(k) if (py + cp - r) % 2 == 0]
                cols = sorted((s for s in range(k) if (px + cp - s) % 2 == 0), key=lambda s: (px + cp - s) // 2)
                dxo = [(px + cp - s) // 2 for s in cols]          # ascending horizontal offsets, consecutive integers
                assert dxo == list(range(dxo[0], dxo[0] + len(dxo)))
                hph, wph = _ceil(xg.h - py, 2), _ceil(xg.w - px, 2)
                om = OutMap(base=out.base + py * out.sy + px * out.sx, sn=out.sn, sy=2 * out.sy, sx=2 * out.sx, sc=out.sc, dtype=out.dtype)
                row_taps = [((py + cp - r) // 2, 0, i) for i, r in enumerate(rows)]
                sp = plan_shift_flat(dyg, len(rows), len(cols), cout, cin, row_taps, om, (P, P + hph), (P, P + wph),
                                     out_shift=-dxo[0], note=note)
                plans.append((sp, wmap_shift(w_shape, k, cout, dgrad=True, rows=rows, cols=cols)))
        return plans
    for py in range(st):
        for px in range(st):
            taps, wt = [], []
            for r in range(k):
                if (py + cp - r) % st:
                    continue
                for s in range(k):
                    if (px + cp - s) % st:
                        continue
                    kidx = len(taps)
                    taps.append(((py + cp - r) // st + dyg.pad, (px + cp - s) // st + dyg.pad, kidx))
                    wt.append((r, s, kidx))
            hph, wph = _ceil(xg.h - py, st), _ceil(xg.w - px, st)
            om = OutMap(base=out.base + py * out.sy + px * out.sx, sn=out.sn, sy=st * out.sy, sx=st * out.sx, sc=out.sc, dtype=out.dtype)
            if not taps:  # k == 1: odd phases receive nothing; the caller pre-zeroes dX
                continue
            sp = plan_box(dyg, taps, cout, cin, hph, wph, 1, om, **T)
            plans.append((sp, wmap_taps(w_shape, cin, wt, cout, swap=True)))
    return plans


# ------------------------------------------------------------------------- weight gradient
def conv_wgrad_plan(w_shape, dyg: Geom, xg: Geom, stride: int, cp: int, *, transposed=False, note="", tf32=False) -> Tuple[IgemmSpec, torch.Tensor]:
    """Weight-gradient plan: fp32 [rows][ldo] accumulated with atomics (zero it first), plus the index map that
    scatters it into the reference layout."""
    kh, kw = w_shape[2], w_shape[3]
    k = kh
    if transposed:
        cin, cout = w_shape[0], w_shape[1]
        assert stride == 2 and dyg.c >= cout and xg.c == cin
        o = dyg.pad - cp
        taps = [(r + o, s + o, r * k + s) for r in range(k) for s in range(k)]
        sp = plan_wgrad_box(xg, cin, dyg, cout, taps, xg.h, xg.w, 2, m_origin=(xg.pad, xg.pad), note=note, tf32=tf32)
        wt = [(r, s, r * k + s) for r in range(k) for s in range(k)]
        return sp, wmap_taps(w_shape, cin, wt, cout, transposed_layout=True, swap=True)
    cout, cin = w_shape[0], w_shape[1]
    o = xg.pad - cp
    ho, wo = out_size(xg.h, k, stride, cp), out_size(xg.w, k, stride, cp)
    assert (dyg.h, dyg.w) == (ho, wo) or (dyg.h >= ho and dyg.w >= wo)
    if (not tf32 and stride == 1 and xg.c == 64 and cin == 64 and cout <= 8 and dyg.c == 8 and k <= 8 and dyg.pad >= o + k - 1
            and 2 * xg.pad - o <= dyg.pad and (dyg.h, dyg.w) == (ho, wo)):
        sp = plan_wgrad_small_cout(xg, cin, dyg, k, o, note=note)
        return sp, wmap_small_cout(w_shape, k, sp.box_taps)
    if xg.c >= 64 or (tf32 and xg.c == cin and cin % 32 == 0):
        assert xg.c == cin
        taps = [(r + o, s + o, r * k + s) for r in range(k) for s in range(k)]
        sp = plan_wgrad_box(dyg, cout, xg, cin, taps, ho, wo, stride, m_origin=(dyg.pad, dyg.pad), note=note, tf32=tf32)
        return sp, wmap_taps(w_shape, cout, [(r, s, r * k + s) for r in range(k) for s in range(k)], cin)
    win = _ceil(k * xg.c, 64) * 64
    taps = [(r + o, o, r) for r in range(k)]
    sp = plan_wgrad_box(dyg, cout, xg, cin, taps, ho, wo, stride, m_origin=(dyg.pad, dyg.pad), n_packed_win=win, note=note, tf32=tf32)
    return sp, wmap_packed(w_shape, cout, k, k, xg.c, win)
