"""B200-native drop-ins for the network factory of phymhan/pc-gan (models/networks.py):
define_G / define_D / define_E / GANLoss with the reference's signatures, forward
signatures and state_dict key names, whose forward and backward run entirely on the
hand-written sm_100a kernels behind the C ABI (include/pcgan_kernels.h).

Scope (BASELINE.json north_star): ResnetGenerator (networks.py:565-652),
NLayerDiscriminator (:737-783), SiameseFeature over ResNetFeature/resnet18
(:1008-1083, :1310-1359; models/resnet.py), GANLoss (:386-420).  Other values of
which_model_* raise NotImplementedError.  There is no CPU path: modules must live on a
CUDA device (gpu_ids non-empty), otherwise forward raises.
"""
import functools
import os

import numpy as np
import torch
import torch.nn as nn
from torch.nn import init

from . import _lib as L
from . import custom_ops as CO
from . import ops
from .engine import Arena, ConvRT, NormState, Pool, RunningStats, WeightBank, accumulate_grad, zeros_act
from .plan import Geom, OutMap

EPS = 1e-5
MOMENTUM = 0.1


# ---------------------------------------------------------------------------------------
# small shared pieces
# ---------------------------------------------------------------------------------------
def _unit_forward(conv: ConvRT, xbuf, rbuf, ns: NormState, count, running: RunningStats, *, gamma=None, beta=None, rmean=None,
                  rvar=None, nbt=None, sequential=False):
    """conv (+bias) with statistics in the epilogue.  The statistics are finalized by the norm_apply that consumes them
    (NormState.apply_kw) and the running statistics by the pass's one batched launch (`running`)."""
    if not ns.pooled:
        ns.stats.zero_()
    ns.affine = gamma is not None
    conv.forward(xbuf, rbuf, ns.stats)
    ns.fused = (count, gamma, beta)
    running.add(ns, count, rmean, rvar, nbt, MOMENTUM, sequential)


def _norm_backward(gy, gy_pad, rbuf, rg: Geom, ns: NormState, act, slope, count, dx, dx_pad, *, res=None, res_pad=0,
                   res_scale=None, res_shift=None, res_groups=1, dres=None, dres_pad=0, dy_fold=0, drop_mask=None, post_mask=None):
    """Backward of one norm + activation unit.  dy_fold=2: `gy` is the gradient of the reflect-padded buffer (pad gy_pad)
    straight out of the data-gradient kernel; its halo is folded onto the interior while it is read."""
    kw = dict(res=res, res_pad=res_pad, mean=ns.mean, rstd=ns.rstd, scale=ns.scale, shift=ns.shift, groups=ns.groups,
              res_scale=res_scale, res_shift=res_shift, res_groups=res_groups, act=act, act_slope=slope, count=count,
              sums=ns.sums, dy_fold=dy_fold, affine=ns.affine, drop_mask=drop_mask, post_mask=post_mask)
    if not ns.pooled:
        ns.sums.zero_()
    ops.norm_bwd(gy, gy_pad, rbuf, rg, dx=dx, dx_pad=dx_pad, dres=dres, dres_pad=dres_pad, **kw)


# PCGAN_FOLD_PREPASS=0: fold reflect-padded gradients on the fly inside the streaming consumers (the round-1 scheme)
FOLD_PREPASS = os.environ.get("PCGAN_FOLD_PREPASS", "1") != "0"


def _fold_in_place(gfull, gpad: Geom):
    """The gradient of a reflect-padded buffer (geometry gpad) straight out of a data-gradient kernel: add the mirrored halo
    values onto the interior border pixels with one small launch (ops.halo_accumulate) and tell the consumer to drop the
    halo (dy_fold = 1) — or, without the pre-pass, to fold while it streams (dy_fold = 2)."""
    if not FOLD_PREPASS or 2 * gpad.pad + 2 > min(gpad.h, gpad.w):
        return 2
    ops.halo_accumulate(gfull, gpad)
    return 1


def _add_interior(x, xg: Geom, add, add_pad, out):
    """out (un-padded) = interior(x) + add as one streamed pass: norm_apply with identity scale and `add` as the residual
    (faster than the gather-style halo_fold with a dropped halo: tools/norm_bench.py)."""
    ops.norm_apply(x, xg, out, Geom(xg.n, xg.h, xg.w, xg.c, 0), scale=None, shift=None, groups=1, res=add, res_pad=add_pad, act=L.ACT_NONE)


class _Scratch:
    """Backward-only buffers of a program, created on first use and reused by every call (one backward at a time)."""

    def __init__(self, dev):
        self.dev, self.bufs = dev, {}

    def get(self, g: Geom, tag=""):
        k = (g, tag)
        if k not in self.bufs:
            self.bufs[k] = zeros_act(g, self.dev)
        return self.bufs[k]


def _require_cuda(t, who):
    if not t.is_cuda:
        raise L.PcganError("%s: pcgan_b200 modules run only on a CUDA device (no CPU fallback); input is on %s" % (who, t.device))


# ---------------------------------------------------------------------------------------
# ResnetGenerator
# ---------------------------------------------------------------------------------------
class _GenWorkspace:
    pass


class _GenProgram:
    """ResnetGenerator at fixed (N, S): reflect-pad 7x7 stem, two stride-2 convs, n_blocks ResnetBlocks,
    two transposed convs, reflect-pad 7x7 head with tanh (networks.py:578-605), InstanceNorm + ReLU between."""

    def __init__(self, mod, N, S):
        self.mod, self.N, self.S = mod, N, S
        dev = mod.model[1].weight.device
        self.dev = dev
        ngf, nb = mod.ngf, mod.n_blocks
        C1, C2, C3 = ngf, 2 * ngf, 4 * ngf
        h2, h4 = S // 2, S // 4
        m = mod.model
        G = Geom
        self.g_x0 = G(N, S, S, 8, 3)
        self.g_r1, self.g_a1 = G(N, S, S, C1, 0), G(N, S, S, C1, 1)
        self.g_r2, self.g_a2 = G(N, h2, h2, C2, 0), G(N, h2, h2, C2, 1)
        self.g_r3, self.g_b = G(N, h4, h4, C3, 0), G(N, h4, h4, C3, 1)
        self.g_bfull = G(N, h4 + 2, h4 + 2, C3, 0)
        self.g_u1r, self.g_u1 = G(N, h2, h2, C2, 0), G(N, h2, h2, C2, 1)
        self.g_u2r, self.g_u2 = G(N, S, S, C1, 0), G(N, S, S, C1, 3)
        self.g_u2full = G(N, S + 6, S + 6, C1, 0)
        self.g_x0full = G(N, S + 6, S + 6, 8, 0)
        self.g_dyh = G(N, S, S, 8, 6)
        ps = dict(stats=True, per_sample_stats=True)
        self.stem = ConvRT("G.model.1", m[1].weight, m[1].bias, self.g_x0, 1, 3, OutMap.nhwc(self.g_r1), dyg=G(N, S, S, C1, 3),
                           dx_out=OutMap.nhwc(self.g_x0full), full_padded=True, **ps)
        self.down1 = ConvRT("G.model.4", m[4].weight, m[4].bias, self.g_a1, 2, 1, OutMap.nhwc(self.g_r2), dyg=self.g_r2,
                            dx_out=OutMap.nhwc(self.g_r1), **ps)
        self.down2 = ConvRT("G.model.7", m[7].weight, m[7].bias, self.g_a2, 2, 1, OutMap.nhwc(self.g_r3), dyg=self.g_r3,
                            dx_out=OutMap.nhwc(self.g_r2), **ps)
        self.blocks = []
        for i in range(nb):
            cb = m[10 + i].conv_block
            ca = ConvRT("G.model.%d.conv_block.1" % (10 + i), cb[1].weight, cb[1].bias, self.g_b, 1, 1, OutMap.nhwc(self.g_r3),
                        dyg=self.g_b, dx_out=OutMap.nhwc(self.g_bfull), full_padded=True, **ps)
            cbb = ConvRT("G.model.%d.conv_block.5" % (10 + i), cb[5].weight, cb[5].bias, self.g_b, 1, 1, OutMap.nhwc(self.g_r3),
                         dyg=self.g_b, dx_out=OutMap.nhwc(self.g_bfull), full_padded=True, **ps)
            self.blocks.append((ca, cbb, cb[2], cb[6]))
        b = 10 + nb
        self.i_up1, self.i_up2, self.i_head = b, b + 3, b + 7
        self.up1 = ConvRT("G.model.%d" % b, m[b].weight, m[b].bias, self.g_b, 2, 1, OutMap.nhwc(self.g_u1r), transposed=True,
                          output_padding=1, dyg=self.g_u1, dx_out=OutMap.nhwc(self.g_r3), **ps)
        self.up2 = ConvRT("G.model.%d" % (b + 3), m[b + 3].weight, m[b + 3].bias, self.g_u1, 2, 1, OutMap.nhwc(self.g_u2r),
                          transposed=True, output_padding=1, dyg=self.g_a1, dx_out=OutMap.nhwc(self.g_u1r), **ps)
        self.head = ConvRT("G.model.%d" % (b + 7), m[b + 7].weight, m[b + 7].bias, self.g_u2, 1, 3,
                           OutMap.nchw(N, mod.output_nc, S, S), act=L.ACT_TANH, dyg=self.g_dyh,
                           dx_out=OutMap.nhwc(self.g_u2full), full_padded=True)
        self.scratch = _Scratch(dev)
        self.keep_scratch = False
        self.pool = Pool(lambda key: self._new_ws())
        self.convs = [self.stem, self.down1, self.down2] + [c for blk in self.blocks for c in blk[:2]] + [self.up1, self.up2, self.head]
        self.bank = WeightBank(self.convs, dev)
        if getattr(mod, "defer_wgrad", False):
            self.bank.enable_deferred()

    def _new_ws(self):
        ws, dev, N = _GenWorkspace(), self.dev, self.N
        z = lambda g: zeros_act(g, dev)
        ws.x0, ws.r1, ws.a1, ws.r2, ws.a2, ws.r3 = z(self.g_x0), z(self.g_r1), z(self.g_a1), z(self.g_r2), z(self.g_a2), z(self.g_r3)
        ws.b = [z(self.g_b) for _ in range(len(self.blocks) + 1)]
        ws.ra = [z(self.g_r3) for _ in self.blocks]
        ws.h = [z(self.g_b) for _ in self.blocks]
        ws.rb = [z(self.g_r3) for _ in self.blocks]
        ws.u1r, ws.u1, ws.u2r, ws.u2 = z(self.g_u1r), z(self.g_u1), z(self.g_u2r), z(self.g_u2)
        C1, C2, C3 = self.g_r1.c, self.g_r2.c, self.g_r3.c
        ws.stats_arena, ws.sums_arena = Arena(dev), Arena(dev)
        NS = lambda c: NormState(N, c, dev, ws.stats_arena, ws.sums_arena)
        ws.n1, ws.n2, ws.n3 = NS(C1), NS(C2), NS(C3)
        ws.na = [NS(C3) for _ in self.blocks]
        ws.nb = [NS(C3) for _ in self.blocks]
        ws.nu1, ws.nu2 = NS(C2), NS(C1)
        ws.head_sums = ws.sums_arena.take((1, 8, 2))
        ws.stats_arena.finalize()
        ws.sums_arena.finalize()
        ws.running = RunningStats(dev)
        return ws

    # ---------------------------------------------------------------- forward
    def forward(self, x, z):
        m, S, N = self.mod.model, self.S, self.N
        ws = self.pool.take(0)
        h2, h4 = S // 2, S // 4
        self.bank.ensure_packed()
        ws.stats_arena.zero()
        ws.running.begin()
        ops.pack_nchw(x, ws.x0, self.g_x0, z=z, halo=L.HALO_REFLECT)
        _unit_forward(self.stem, ws.x0, ws.r1, ws.n1, S * S, ws.running, rmean=m[2].running_mean, rvar=m[2].running_var)
        ops.norm_apply(ws.r1, self.g_r1, ws.a1, self.g_a1, y_halo=L.HALO_ZERO, **ws.n1.apply_kw(), act=L.ACT_RELU)
        _unit_forward(self.down1, ws.a1, ws.r2, ws.n2, h2 * h2, ws.running, rmean=m[5].running_mean, rvar=m[5].running_var)
        ops.norm_apply(ws.r2, self.g_r2, ws.a2, self.g_a2, y_halo=L.HALO_ZERO, **ws.n2.apply_kw(), act=L.ACT_RELU)
        _unit_forward(self.down2, ws.a2, ws.r3, ws.n3, h4 * h4, ws.running, rmean=m[8].running_mean, rvar=m[8].running_var)
        nblk = len(self.blocks)
        ops.norm_apply(ws.r3, self.g_r3, ws.b[0], self.g_b, y_halo=L.HALO_REFLECT if nblk else L.HALO_ZERO, **ws.n3.apply_kw(), act=L.ACT_RELU)
        for i, (ca, cb, na_mod, nb_mod) in enumerate(self.blocks):
            _unit_forward(ca, ws.b[i], ws.ra[i], ws.na[i], h4 * h4, ws.running, rmean=na_mod.running_mean, rvar=na_mod.running_var)
            ops.norm_apply(ws.ra[i], self.g_r3, ws.h[i], self.g_b, y_halo=L.HALO_REFLECT, **ws.na[i].apply_kw(), act=L.ACT_RELU)
            _unit_forward(cb, ws.h[i], ws.rb[i], ws.nb[i], h4 * h4, ws.running, rmean=nb_mod.running_mean, rvar=nb_mod.running_var)
            last = i == nblk - 1
            # x + conv_block(x) (networks.py:650-652): residual added after the norm, no activation
            ops.norm_apply(ws.rb[i], self.g_r3, ws.b[i + 1], self.g_b, y_halo=L.HALO_ZERO if last else L.HALO_REFLECT,
                           **ws.nb[i].apply_kw(), res=ws.b[i], res_pad=1, act=L.ACT_NONE)
        _unit_forward(self.up1, ws.b[nblk], ws.u1r, ws.nu1, h2 * h2, ws.running, rmean=m[self.i_up1 + 1].running_mean, rvar=m[self.i_up1 + 1].running_var)
        ops.norm_apply(ws.u1r, self.g_u1r, ws.u1, self.g_u1, y_halo=L.HALO_ZERO, **ws.nu1.apply_kw(), act=L.ACT_RELU)
        _unit_forward(self.up2, ws.u1, ws.u2r, ws.nu2, S * S, ws.running, rmean=m[self.i_up2 + 1].running_mean, rvar=m[self.i_up2 + 1].running_var)
        ops.norm_apply(ws.u2r, self.g_u2r, ws.u2, self.g_u2, y_halo=L.HALO_REFLECT, **ws.nu2.apply_kw(), act=L.ACT_RELU)
        out = torch.empty(N, self.mod.output_nc, S, S, device=self.dev)
        self.head.forward(ws.u2, out)
        ws.running.flush()
        return out, ws

    # --------------------------------------------------------------- backward
    def backward(self, ws, out, dout, need_dx, need_w, need_dz=False):
        S, N, sc = self.S, self.N, self.scratch
        h2, h4 = S // 2, S // 4
        m = self.mod.model
        nblk = len(self.blocks)
        R, Z = L.ACT_RELU, L.ACT_NONE
        self.bank.ensure_packed()
        ws.sums_arena.zero()
        if need_w:
            self.bank.begin_backward()
            # biases in front of an affine-less InstanceNorm have exactly zero gradient
            for c in self.convs[:-1]:
                if c.bias is not None and c.bias.grad is None:
                    c.bias.grad = torch.zeros_like(c.bias)
        # head: d(pre-tanh) = dout * (1 - out^2)
        dyh = sc.get(self.g_dyh)
        ops.pack_nchw(dout, dyh, self.g_dyh, mul_out=out, mul_kind=L.ACT_TANH, halo=L.HALO_ZERO)
        if need_w:
            # every non-convolution gradient of a layer is in place before the layer's weight gradient is launched: that
            # launch may complete a bucket of the gradient all-reduce (WeightBank.layer_done)
            hs = ws.head_sums.t
            ops.norm_bwd_reduce(dyh, 6, dyh, self.g_dyh, sums=hs, count=0.0)
            accumulate_grad(self.head.bias, hs[0, : self.mod.output_nc, 0])
            self.head.backward_weight(dyh, ws.u2)
        dfull = sc.get(self.g_u2full)
        self.head.backward_data(dyh, dfull)
        # up2 unit; the reflect-pad fold of the head's data gradient happens while it is read
        dy = sc.get(self.g_a1, "dy")
        _norm_backward(dfull, 3, ws.u2r, self.g_u2r, ws.nu2, R, 0.0, S * S, dy, 1, dy_fold=_fold_in_place(dfull, self.g_u2))
        if need_w:
            self.up2.backward_weight(dy, ws.u1)
        g = sc.get(self.g_u1r, "g_u1")
        self.up2.backward_data(dy, g)
        # up1 unit
        dy = sc.get(self.g_u1, "dy")
        _norm_backward(g, 0, ws.u1r, self.g_u1r, ws.nu1, R, 0.0, h2 * h2, dy, 1)
        if need_w:
            self.up1.backward_weight(dy, ws.b[nblk])
        gb = sc.get(self.g_r3, "gb0")
        self.up1.backward_data(dy, gb)
        # residual blocks, last to first; gb = gradient of the block output
        for i in range(nblk - 1, -1, -1):
            ca, cb, _, _ = self.blocks[i]
            # the blocks share their backward buffers; keep_scratch (tests) gives every block its own so that the whole
            # chain can be inspected afterwards
            t = str(i) if self.keep_scratch else ""
            dyb = sc.get(self.g_b, "dyb" + t)
            _norm_backward(gb, 0, ws.rb[i], self.g_r3, ws.nb[i], Z, 0.0, h4 * h4, dyb, 1, res=ws.b[i], res_pad=1)
            if need_w:
                cb.backward_weight(dyb, ws.h[i])
            dfull = sc.get(self.g_bfull, "dfull" + t)
            cb.backward_data(dyb, dfull)
            dya = sc.get(self.g_b, "dya" + t)
            _norm_backward(dfull, 1, ws.ra[i], self.g_r3, ws.na[i], R, 0.0, h4 * h4, dya, 1, dy_fold=_fold_in_place(dfull, self.g_b))
            if need_w:
                ca.backward_weight(dya, ws.b[i])
            dfull2 = sc.get(self.g_bfull, "dfull2" + t)
            ca.backward_data(dya, dfull2)
            gprev = sc.get(self.g_r3, ("gbk%d" % i) if self.keep_scratch else "gb%d" % ((nblk - i) % 2))
            if _fold_in_place(dfull2, self.g_b) != 1:
                ops.halo_fold(dfull2, self.g_b, gprev, 0, halo=L.HALO_REFLECT, add=gb, add_pad=0)
            else:
                # 18.8 us at the block shape against 27 us for halo_fold with a dropped halo: 20.76 -> 20.57 ms per step
                _add_interior(dfull2, self.g_b, gb, 0, gprev)
            gb = gprev
        # down2 unit (its output buffer ws.b[0])
        dy = sc.get(self.g_r3, "dy3")
        _norm_backward(gb, 0, ws.r3, self.g_r3, ws.n3, R, 0.0, h4 * h4, dy, 0)
        if need_w:
            self.down2.backward_weight(dy, ws.a2)
        g = sc.get(self.g_r2, "g")
        self.down2.backward_data(dy, g)
        dy = sc.get(self.g_r2, "dy2")
        _norm_backward(g, 0, ws.r2, self.g_r2, ws.n2, R, 0.0, h2 * h2, dy, 0)
        if need_w:
            self.down1.backward_weight(dy, ws.a1)
        g = sc.get(self.g_r1, "g")
        self.down1.backward_data(dy, g)
        # stem unit
        dy = sc.get(Geom(N, S, S, self.g_r1.c, 3), "dy1")
        _norm_backward(g, 0, ws.r1, self.g_r1, ws.n1, R, 0.0, S * S, dy, 3)
        if need_w:
            self.stem.backward_weight(dy, ws.x0)
        dx = dz = None
        if need_dx or need_dz:
            dfull = sc.get(self.g_x0full)
            self.stem.backward_data(dy, dfull)
            gx = sc.get(Geom(N, S, S, 8, 0), "gx")
            ops.halo_fold(dfull, self.g_x0, gx, 0, halo=L.HALO_REFLECT)
            # channels [0, input_nc) are the image, channel input_nc is the constant embedding plane (networks.py:610-611)
            nc = self.mod.input_nc_img
            if need_dz:
                dxz = torch.empty(N, nc + 1, S, S, device=self.dev)
                ops.unpack_resize_bwd(gx, Geom(N, S, S, 8, 0), dxz)
                dx = dxz[:, :nc].contiguous() if need_dx else None
                dz = dxz[:, nc].sum((1, 2))
            else:
                dx = torch.empty(N, nc, S, S, device=self.dev)
                ops.unpack_resize_bwd(gx, Geom(N, S, S, 8, 0), dx)
        return dx, dz


class _ResnetBlockHolder(nn.Module):
    """Parameter holder with the reference's names: conv_block.{1,5} convs, conv_block.{2,6} norms (networks.py:621-648)."""

    def __init__(self, dim, norm_layer, use_bias):
        super().__init__()
        self.conv_block = nn.Sequential(
            nn.Identity(), nn.Conv2d(dim, dim, 3, bias=use_bias), norm_layer(dim), nn.Identity(),
            nn.Identity(), nn.Conv2d(dim, dim, 3, bias=use_bias), norm_layer(dim))


class ResnetGenerator(nn.Module):
    """Same constructor, forward signature and state_dict keys as models/networks.py:565-612; `model` only holds the
    parameters and running statistics (its torch forward is never called) — the math runs in _GenProgram."""

    def __init__(self, input_nc, output_nc, nz=0, ngf=64, norm_layer=nn.BatchNorm2d, dropout=0, n_blocks=6, padding_type="reflect"):
        super().__init__()
        func = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
        if func is not nn.InstanceNorm2d or padding_type != "reflect" or dropout:
            raise NotImplementedError("pcgan_b200 ResnetGenerator: instance norm, reflect padding, no dropout (the wsgan_emb configuration)")
        if ngf % 64 != 0:
            raise NotImplementedError("pcgan_b200 ResnetGenerator needs ngf to be a multiple of 64 (tcgen05 K chunk); got %d" % ngf)
        if input_nc + nz > 8 or output_nc > 8:
            raise NotImplementedError("input_nc + nz and output_nc must be <= 8")
        self.input_nc_img, self.input_nc, self.output_nc, self.ngf, self.n_blocks, self.nz = input_nc, input_nc + nz, output_nc, ngf, n_blocks, nz
        layers = [nn.Identity(), nn.Conv2d(input_nc + nz, ngf, 7), norm_layer(ngf), nn.Identity()]
        for i in range(2):
            c = ngf * 2 ** i
            layers += [nn.Conv2d(c, 2 * c, 3, stride=2, padding=1), norm_layer(2 * c), nn.Identity()]
        for _ in range(n_blocks):
            layers.append(_ResnetBlockHolder(4 * ngf, norm_layer, True))
        for i in range(2):
            c = ngf * 2 ** (2 - i)
            layers += [nn.ConvTranspose2d(c, c // 2, 3, stride=2, padding=1, output_padding=1), norm_layer(c // 2), nn.Identity()]
        layers += [nn.Identity(), nn.Conv2d(ngf, output_nc, 7), nn.Identity()]
        self.model = nn.Sequential(*layers)
        self._programs = {}
        self._key = CO.register_module(self)

    def _program(self, n, s):
        k = (n, s, self.model[1].weight.device)
        if k not in self._programs:
            self._programs[k] = _GenProgram(self, n, s)
        return self._programs[k]

    def zero_wgrad(self):
        """Deferred weight gradients (defer_wgrad = True): clear the packed accumulators (optimizer.zero_grad time)."""
        for prog in self._programs.values():
            prog.bank.zero_wgrad()

    def flush_wgrad(self):
        """... and scatter them into the parameters' .grad after the backward sweeps of the step."""
        for prog in self._programs.values():
            prog.bank.flush_wgrad()

    def forward(self, input, z=None):
        _require_cuda(input, "ResnetGenerator")
        if input.shape[2] != input.shape[3] or input.shape[2] % 4:
            raise NotImplementedError("square inputs with side a multiple of 4")
        if z is None or self.nz != 1:
            raise NotImplementedError("ResnetGenerator needs the 1-channel embedding z (nz=1)")
        out = torch.ops.pcgan.resnet_generator(input, z, list(self.parameters()), self._key)
        CO.finish_forward(self._key, out)
        return out


# ---------------------------------------------------------------------------------------
# UnetGenerator (networks.py:655-733): the default G of wsgan_emb (unet_128)
# ---------------------------------------------------------------------------------------
class _UnetWorkspace:
    pass


class _UnetProgram:
    """UnetGenerator(num_downs = D) at fixed (N, S) with InstanceNorm (wsgan_emb: --norm_G instance).

    Level k = 0 .. D-1 (0 outermost): down convolution k (4x4 s2 p1) takes the image (k = 0) or A[k-1] = LeakyReLU(n_{k-1});
    n_k = InstanceNorm(conv_k) for 0 < k < D-1, the raw convolution for k = 0 and k = D-1.  The reference's LeakyReLU is
    in place, so the skip connection carries LeakyReLU(n_k) and the up path's ReLU turns it into ReLU(n_k)
    (networks.py:692-733).  The concatenation consumed by up convolution k is ONE buffer B[k] of 2 C_k channels: its first
    half ReLU(n_k) and its second half ReLU(InstanceNorm(upconv_{k+1})) are written by two norm_apply launches with a
    channel offset, so no concat pass exists.  Backward: the data gradient of an up convolution is launched once per half
    of its input channels (two weight views), so each half lands in its own dense buffer."""

    def __init__(self, mod, N, S):
        self.mod, self.N, self.S = mod, N, S
        D = mod.num_downs
        if S % (1 << D) or S < (1 << D):
            raise NotImplementedError("UnetGenerator(num_downs=%d) needs a side that is a multiple of %d" % (D, 1 << D))
        self.D = D
        dev = mod.down[0].weight.device
        self.dev = dev
        G = Geom
        C = mod.inner
        sz = [S >> (k + 1) for k in range(D)]
        self.sz, self.C = sz, C
        self.g_x0 = G(N, S, S, 8, 1)
        self.g_r = [G(N, sz[k], sz[k], C[k], 0) for k in range(D)]             # raw conv outputs / dense gradients
        self.g_A = [G(N, sz[k], sz[k], C[k], 1) for k in range(D)]             # activation buffers (A[k], R at k = D-1), dY of up conv k+1
        self.g_B = [G(N, sz[k], sz[k], 2 * C[k], 1) for k in range(D - 1)]     # concat buffers
        ps = dict(stats=True, per_sample_stats=True)
        self.down, self.up, self.up_ds, self.up_du = [], [], [], []
        for k in range(D):
            cv = mod.down[k]
            xg = self.g_x0 if k == 0 else self.g_A[k - 1]
            dx_out = OutMap.nhwc(G(N, S, S, 8, 0)) if k == 0 else OutMap.nhwc(self.g_r[k - 1])
            if k == 0:
                kw = dict(act=L.ACT_LRELU, act_slope=0.2)
                out = OutMap.nhwc(self.g_A[0])
            elif k == D - 1:
                kw = dict(act=L.ACT_RELU)
                out = OutMap.nhwc(self.g_A[k])
            else:
                kw, out = ps, OutMap.nhwc(self.g_r[k])
            self.down.append(ConvRT("Gu.down%d" % k, cv.weight, cv.bias, xg, 2, 1, out, dyg=self.g_r[k], dx_out=dx_out, **kw))
        for k in range(D):
            cv = mod.up[k]
            xg = self.g_A[k] if k == D - 1 else self.g_B[k]
            so = S if k == 0 else sz[k - 1]
            co = mod.output_nc if k == 0 else C[k - 1]
            if k == 0:
                out, kw, dyg = OutMap.nchw(N, co, S, S), dict(act=L.ACT_TANH), G(N, S, S, 8, 1)
            else:
                out, kw, dyg = OutMap.nhwc(self.g_r[k - 1]), ps, self.g_A[k - 1]
            name = "Gu.up%d" % k
            if k == D - 1:
                self.up.append(ConvRT(name, cv.weight, cv.bias, xg, 2, 1, out, transposed=True, dyg=dyg, dx_out=OutMap.nhwc(self.g_r[k]), **kw))
                self.up_ds.append(None); self.up_du.append(None)
            else:
                self.up.append(ConvRT(name, cv.weight, cv.bias, xg, 2, 1, out, transposed=True, dyg=dyg, want_dgrad=False, **kw))
                # the data gradient, one launch set per half of the input channels (views of the IOHW weight)
                half = dict(transposed=True, dyg=dyg, dx_out=OutMap.nhwc(self.g_r[k]), want_wgrad=False, want_fwd=False, trainable=False)
                dummy = OutMap.nhwc(G(N, so, so, max(8, co), 0))
                self.up_ds.append(ConvRT(name + ".dskip", cv.weight[:C[k]], None, self.g_A[k], 2, 1, dummy, **half))
                self.up_du.append(ConvRT(name + ".dup", cv.weight[C[k]:], None, self.g_A[k], 2, 1, dummy, **half))
        self.scratch = _Scratch(dev)
        self.pool = Pool(lambda key: self._new_ws())
        self.convs = self.down + self.up
        extra = [c for c in self.up_ds + self.up_du if c is not None]
        self.bank = WeightBank(self.convs + extra, dev)
        if getattr(mod, "defer_wgrad", False):
            self.bank.enable_deferred()

    def _new_ws(self):
        ws, dev, N, D = _UnetWorkspace(), self.dev, self.N, self.D
        z = lambda g: zeros_act(g, dev)
        ws.x0 = z(self.g_x0)
        ws.A = [z(g) for g in self.g_A]
        ws.B = [z(g) for g in self.g_B]
        ws.r = [None] + [z(self.g_r[k]) for k in range(1, D - 1)] + [None]        # raw down outputs that are normalised
        ws.u = [None] + [z(self.g_r[k - 1]) for k in range(1, D)]                  # raw up outputs u[k] (geometry of level k-1)
        ws.stats_arena, ws.sums_arena = Arena(dev), Arena(dev)
        NS = lambda c: NormState(N, c, dev, ws.stats_arena, ws.sums_arena)
        ws.nd = [None] + [NS(self.C[k]) for k in range(1, D - 1)] + [None]
        ws.nu = [None] + [NS(self.C[k - 1]) for k in range(1, D)]
        ws.b_sums = [ws.sums_arena.take((1, c, 2)) for c in (self.C[0], self.C[D - 1], 8)]
        ws.stats_arena.finalize()
        ws.sums_arena.finalize()
        ws.running = RunningStats(dev)
        return ws

    def forward(self, x, z):
        mod, N, D, sz = self.mod, self.N, self.D, self.sz
        ws = self.pool.take(0)
        self.bank.ensure_packed()
        ws.stats_arena.zero()
        ws.running.begin()
        ops.pack_nchw(x, ws.x0, self.g_x0, z=z, halo=L.HALO_ZERO)
        ident = dict(scale=None, shift=None, groups=1)
        # ---- down path
        self.down[0].forward(ws.x0, ws.A[0])                                   # LeakyReLU(conv + bias) in the epilogue
        ops.norm_apply(ws.A[0], self.g_A[0], ws.B[0], self.g_B[0], act=L.ACT_RELU, y_c0=0, **ident)
        for k in range(1, D - 1):
            nm = mod.down_norm[k]
            _unit_forward(self.down[k], ws.A[k - 1], ws.r[k], ws.nd[k], sz[k] * sz[k], ws.running, rmean=nm.running_mean, rvar=nm.running_var)
            ops.norm_apply(ws.r[k], self.g_r[k], ws.A[k], self.g_A[k], y_halo=L.HALO_ZERO, **ws.nd[k].apply_kw(), act=L.ACT_LRELU, act_slope=0.2)
            ops.norm_apply(ws.r[k], self.g_r[k], ws.B[k], self.g_B[k], scale=ws.nd[k].scale, shift=ws.nd[k].shift, groups=N, act=L.ACT_RELU, y_c0=0)
        self.down[D - 1].forward(ws.A[D - 2], ws.A[D - 1])                     # innermost: ReLU(conv + bias) (uprelu follows downconv directly)
        # ---- up path
        for k in range(D - 1, 0, -1):
            nm = mod.up_norm[k]
            src = ws.A[k] if k == D - 1 else ws.B[k]
            _unit_forward(self.up[k], src, ws.u[k], ws.nu[k], sz[k - 1] * sz[k - 1], ws.running, rmean=nm.running_mean, rvar=nm.running_var)
            ops.norm_apply(ws.u[k], self.g_r[k - 1], ws.B[k - 1], self.g_B[k - 1], **ws.nu[k].apply_kw(), act=L.ACT_RELU, y_c0=self.C[k - 1])
        out = torch.empty(N, mod.output_nc, self.S, self.S, device=self.dev)
        self.up[0].forward(ws.B[0], out)
        ws.running.flush()
        return out, ws

    def _bias_grad(self, conv, dy, dy_pad, g: Geom, sums):
        ops.norm_bwd_reduce(dy, dy_pad, dy, g, sums=sums, count=0.0)
        accumulate_grad(conv.bias, sums[0, : conv.bias.numel(), 0])

    def backward(self, ws, out, dout, need_dx, need_w, need_dz=False):
        mod, N, D, sz, C, sc = self.mod, self.N, self.D, self.sz, self.C, self.scratch
        self.bank.ensure_packed()
        ws.sums_arena.zero()
        if need_w:
            self.bank.begin_backward()
            for c in self.down[1:D - 1] + self.up[1:]:      # biases in front of an affine-less InstanceNorm: zero gradient
                if c.bias is not None and c.bias.grad is None:
                    c.bias.grad = torch.zeros_like(c.bias)
        # ---- outermost up convolution: d(pre-tanh) = dout * (1 - out^2)
        g_dyh = Geom(N, self.S, self.S, 8, 1)
        dyh = sc.get(g_dyh, "dyh")
        ops.pack_nchw(dout, dyh, g_dyh, mul_out=out, mul_kind=L.ACT_TANH, halo=L.HALO_ZERO)
        if need_w:
            self._bias_grad(self.up[0], dyh, 1, g_dyh, ws.b_sums[2].t)
            self.up[0].backward_weight(dyh, ws.B[0])
        gs, gu = sc.get(self.g_r[0], "gs0"), sc.get(self.g_r[0], "gu0")
        self.up_ds[0].backward_data(dyh, gs)
        self.up_du[0].backward_data(dyh, gu)
        skip = [None] * D          # skip[k]: gradient of ReLU(n_k) through the concatenation
        skip[0] = gs
        # ---- up path, outermost to innermost: level k's up convolution produced the second half of B[k-1]
        for k in range(1, D):
            dyu = sc.get(self.g_A[k - 1], "dyu%d" % k)
            _norm_backward(gu, 0, ws.u[k], self.g_r[k - 1], ws.nu[k], L.ACT_RELU, 0.0, sz[k - 1] * sz[k - 1], dyu, 1)
            src = ws.A[k] if k == D - 1 else ws.B[k]
            if need_w:
                self.up[k].backward_weight(dyu, src)
            if k == D - 1:
                gR = sc.get(self.g_r[k], "gR")
                self.up[k].backward_data(dyu, gR)
            else:
                gs, gu = sc.get(self.g_r[k], "gs%d" % k), sc.get(self.g_r[k], "gu%d" % k)
                self.up_ds[k].backward_data(dyu, gs)
                self.up_du[k].backward_data(dyu, gu)
                skip[k] = gs
        # ---- innermost down convolution: ReLU in its epilogue
        k = D - 1
        dy = sc.get(self.g_r[k], "dy%d" % k)
        ops.act_bwd(gR, 0, ws.A[k], 1, dy, 0, self.g_r[k], 0.0)
        if need_w:
            self._bias_grad(self.down[k], dy, 0, self.g_r[k], ws.b_sums[1].t)
            self.down[k].backward_weight(dy, ws.A[k - 1])
        dA = sc.get(self.g_r[k - 1], "dA%d" % (k - 1))
        self.down[k].backward_data(dy, dA)
        # ---- down path, inner to outer: n_k receives the skip gradient through ReLU and dA[k] through LeakyReLU(0.2); both
        # masks are the sign of n_k, so dy_eff = dA + skip * [n_k > 0] followed by the LeakyReLU backward gives their sum
        for k in range(D - 2, -1, -1):
            tmp = sc.get(self.g_r[k], "tmp%d" % k)
            ops.act_bwd(skip[k], 0, ws.A[k], 1, tmp, 0, self.g_r[k], 0.0)
            eff = sc.get(self.g_r[k], "eff%d" % k)
            _add_interior(tmp, self.g_r[k], dA, 0, eff)
            dy = sc.get(self.g_r[k], "dy%d" % k)
            if k > 0:
                _norm_backward(eff, 0, ws.r[k], self.g_r[k], ws.nd[k], L.ACT_LRELU, 0.2, sz[k] * sz[k], dy, 0)
                if need_w:
                    self.down[k].backward_weight(dy, ws.A[k - 1])
                dA = sc.get(self.g_r[k - 1], "dA%d" % (k - 1))
                self.down[k].backward_data(dy, dA)
            else:
                ops.act_bwd(eff, 0, ws.A[0], 1, dy, 0, self.g_r[0], 0.2)
                if need_w:
                    self._bias_grad(self.down[0], dy, 0, self.g_r[0], ws.b_sums[0].t)
                    self.down[0].backward_weight(dy, ws.x0)
        dx = dz = None
        if need_dx or need_dz:
            gx = sc.get(Geom(N, self.S, self.S, 8, 0), "gx")
            self.down[0].backward_data(dy, gx)
            nc = mod.input_nc_img
            if need_dz:
                dxz = torch.empty(N, nc + 1, self.S, self.S, device=self.dev)
                ops.unpack_resize_bwd(gx, Geom(N, self.S, self.S, 8, 0), dxz)
                dx = dxz[:, :nc].contiguous() if need_dx else None
                dz = dxz[:, nc].sum((1, 2))
            else:
                dx = torch.empty(N, nc, self.S, self.S, device=self.dev)
                ops.unpack_resize_bwd(gx, Geom(N, self.S, self.S, 8, 0), dx)
        return dx, dz


class _UnetBlockHolder(nn.Module):
    """Parameter holder with the reference's nesting and names (UnetSkipConnectionBlock.model, networks.py:685-733):
    outermost [downconv, submodule, uprelu, upconv, tanh]; innermost [downrelu, downconv, uprelu, upconv, upnorm]; otherwise
    [downrelu, downconv, downnorm, submodule, uprelu, upconv, upnorm]."""

    def __init__(self, outer_nc, inner_nc, input_nc=None, submodule=None, outermost=False, innermost=False, norm_layer=None):
        super().__init__()
        input_nc = outer_nc if input_nc is None else input_nc
        I = nn.Identity
        self.downconv = nn.Conv2d(input_nc, inner_nc, 4, 2, 1, bias=True)
        self.downnorm = self.upnorm = None
        if outermost:
            self.upconv = nn.ConvTranspose2d(inner_nc * 2, outer_nc, 4, 2, 1)
            seq = [self.downconv, submodule, I(), self.upconv, I()]
        elif innermost:
            self.upconv = nn.ConvTranspose2d(inner_nc, outer_nc, 4, 2, 1, bias=True)
            self.upnorm = norm_layer(outer_nc)
            seq = [I(), self.downconv, I(), self.upconv, self.upnorm]
        else:
            self.upconv = nn.ConvTranspose2d(inner_nc * 2, outer_nc, 4, 2, 1, bias=True)
            self.downnorm, self.upnorm = norm_layer(inner_nc), norm_layer(outer_nc)
            seq = [I(), self.downconv, self.downnorm, submodule, I(), self.upconv, self.upnorm]
        # the convolutions / norms are registered once, under the Sequential (the attributes above are plain references)
        for name in ("downconv", "upconv", "downnorm", "upnorm"):
            object.__setattr__(self, "_" + name, self._modules.pop(name, None) if name in self._modules else getattr(self, name, None))
        self.model = nn.Sequential(*seq)


class UnetGenerator(nn.Module):
    """Same constructor, forward signature and state_dict keys as models/networks.py:659-682 (norm = InstanceNorm2d with
    running statistics, no dropout: the wsgan_emb configuration); the math runs in _UnetProgram."""

    def __init__(self, input_nc, output_nc, nz=0, num_downs=7, ngf=64, norm_layer=nn.BatchNorm2d, dropout=0):
        super().__init__()
        func = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
        if func is not nn.InstanceNorm2d or dropout:
            raise NotImplementedError("pcgan_b200 UnetGenerator: instance norm, no dropout (the wsgan_emb configuration)")
        if ngf % 64 != 0 or input_nc + nz > 8 or output_nc > 8 or nz != 1 or num_downs < 5:
            raise NotImplementedError("ngf must be a multiple of 64, nz == 1, input_nc + nz and output_nc <= 8, num_downs >= 5")
        self.input_nc_img, self.output_nc, self.nz, self.ngf, self.num_downs = input_nc, output_nc, nz, ngf, num_downs
        self.inner = [ngf, ngf * 2, ngf * 4] + [ngf * 8] * (num_downs - 3)
        blk = _UnetBlockHolder(ngf * 8, ngf * 8, norm_layer=norm_layer, innermost=True)
        blocks = [blk]
        for _ in range(num_downs - 5):
            blk = _UnetBlockHolder(ngf * 8, ngf * 8, submodule=blk, norm_layer=norm_layer)
            blocks.append(blk)
        for outer, inner in ((ngf * 4, ngf * 8), (ngf * 2, ngf * 4), (ngf, ngf * 2)):
            blk = _UnetBlockHolder(outer, inner, submodule=blk, norm_layer=norm_layer)
            blocks.append(blk)
        blk = _UnetBlockHolder(output_nc, ngf, input_nc=input_nc + nz, submodule=blk, outermost=True, norm_layer=norm_layer)
        blocks.append(blk)
        self.model = blk
        levels = blocks[::-1]                  # level 0 = outermost
        object.__setattr__(self, "down", [b._downconv for b in levels])
        object.__setattr__(self, "up", [b._upconv for b in levels])
        object.__setattr__(self, "down_norm", [b._downnorm for b in levels])
        object.__setattr__(self, "up_norm", [b._upnorm for b in levels])
        self._programs = {}
        self._key = CO.register_module(self)

    def _program(self, n, s):
        k = (n, s, self.down[0].weight.device)
        if k not in self._programs:
            self._programs[k] = _UnetProgram(self, n, s)
        return self._programs[k]

    def zero_wgrad(self):
        for prog in self._programs.values():
            prog.bank.zero_wgrad()

    def flush_wgrad(self):
        for prog in self._programs.values():
            prog.bank.flush_wgrad()

    def forward(self, input, z=None):
        _require_cuda(input, "UnetGenerator")
        if z is None or input.shape[2] != input.shape[3]:
            raise NotImplementedError("UnetGenerator needs square inputs and the 1-channel embedding z")
        out = torch.ops.pcgan.unet_generator(input, z, list(self.parameters()), self._key)
        CO.finish_forward(self._key, out)
        return out


# ---------------------------------------------------------------------------------------
# NLayerDiscriminator
# ---------------------------------------------------------------------------------------
class _DiscWorkspace:
    pass


class _DiscProgram:
    """NLayerDiscriminator(n_layers=3) at fixed (N, S): conv4x4s2+LReLU, 2x (conv4x4s2 + BN + LReLU),
    conv4x4s1 + BN + LReLU, conv4x4s1 -> 1 (+ sigmoid) (networks.py:745-777)."""

    def __init__(self, mod, N, S, groups=1):
        """groups > 1: the N samples are `groups` independent passes of the network batched into one (each run of
        N / groups samples keeps its own BatchNorm batch: statistics, running-statistics step, backward sums)."""
        if N % groups:
            raise L.PcganError("NLayerDiscriminator: batch %d is not divisible into %d groups" % (N, groups))
        self.mod, self.N, self.S, self.groups = mod, N, S, groups
        m = mod.model
        dev = m[0].weight.device
        self.dev = dev
        ndf = mod.ndf
        G = Geom
        self.g_x0 = G(N, S, S, 8, 1)
        sizes, chans = [S // 2], [ndf]
        self.idx = [0]
        i = 2
        for n in range(1, mod.n_layers):
            sizes.append(sizes[-1] // 2); chans.append(ndf * min(2 ** n, 8)); self.idx.append(i); i += 3
        sizes.append(sizes[-1] - 1); chans.append(ndf * min(2 ** mod.n_layers, 8)); self.idx.append(i); i += 3
        self.i_head = i
        self.sizes, self.chans = sizes, chans
        self.g_y = [G(N, s, s, c, 1) for s, c in zip(sizes, chans)]      # activations (zero halo, pad 1)
        self.g_r = [G(N, s, s, c, 0) for s, c in zip(sizes, chans)]      # raw conv outputs / gradients
        so = sizes[-1] - 1
        self.so = so
        self.convs = []
        # layer 0: bias + LeakyReLU in the epilogue, written straight into the next padded buffer
        self.convs.append(ConvRT("D.model.0", m[0].weight, m[0].bias, self.g_x0, 2, 1, OutMap.nhwc(self.g_y[0]), act=L.ACT_LRELU,
                                 act_slope=0.2, dyg=self.g_r[0], dx_out=OutMap.nhwc(G(N, S, S, 8, 0))))
        for li in range(1, len(sizes)):
            stride = 2 if li < len(sizes) - 1 else 1
            grp = dict(per_sample_stats=True, stats_div=N // groups) if groups > 1 else {}
            self.convs.append(ConvRT("D.model.%d" % self.idx[li], m[self.idx[li]].weight, None, self.g_y[li - 1], stride, 1,
                                     OutMap.nhwc(self.g_r[li]), stats=True, dyg=self.g_r[li], dx_out=OutMap.nhwc(self.g_r[li - 1]), **grp))
        self.g_dyh = G(N, so, so, 8, 2)
        self.head = ConvRT("D.model.%d" % self.i_head, m[self.i_head].weight, m[self.i_head].bias, self.g_y[-1], 1, 1,
                           OutMap.nchw(N, 1, so, so), act=L.ACT_SIGMOID if mod.use_sigmoid else L.ACT_NONE, dyg=self.g_dyh,
                           dx_out=OutMap.nhwc(self.g_r[-1]))
        self.scratch = _Scratch(dev)
        self.pool = Pool(lambda key: self._new_ws())
        self.bank = WeightBank(self.convs + [self.head], dev)
        if getattr(mod, "defer_wgrad", False):
            self.bank.enable_deferred()

    def _new_ws(self):
        ws, dev = _DiscWorkspace(), self.dev
        ws.x0 = zeros_act(self.g_x0, dev)
        ws.y = [zeros_act(g, dev) for g in self.g_y]
        ws.r = [None] + [zeros_act(g, dev) for g in self.g_r[1:]]
        ws.stats_arena, ws.sums_arena = Arena(dev), Arena(dev)
        ws.ns = [None] + [NormState(self.groups, c, dev, ws.stats_arena, ws.sums_arena) for c in self.chans[1:]]
        ws.head_sums = ws.sums_arena.take((1, 8, 2))
        ws.l0_sums = ws.sums_arena.take((1, self.chans[0], 2))
        ws.stats_arena.finalize()
        ws.sums_arena.finalize()
        ws.running = RunningStats(dev)
        return ws

    def forward(self, x, z):
        m, N = self.mod.model, self.N
        ws = self.pool.take(0)
        self.bank.ensure_packed()
        ws.stats_arena.zero()
        ws.running.begin()
        ops.pack_nchw(x, ws.x0, self.g_x0, z=z, halo=L.HALO_ZERO)
        self.convs[0].forward(ws.x0, ws.y[0])
        for li in range(1, len(self.sizes)):
            bn = m[self.idx[li] + 1]
            cnt = (N // self.groups) * self.sizes[li] ** 2
            _unit_forward(self.convs[li], ws.y[li - 1], ws.r[li], ws.ns[li], cnt, ws.running, gamma=bn.weight.detach(),
                          beta=bn.bias.detach(), rmean=bn.running_mean, rvar=bn.running_var, nbt=bn.num_batches_tracked,
                          sequential=self.groups > 1)
            ops.norm_apply(ws.r[li], self.g_r[li], ws.y[li], self.g_y[li], y_halo=L.HALO_ZERO, **ws.ns[li].apply_kw(), act=L.ACT_LRELU, act_slope=0.2)
        out = torch.empty(N, 1, self.so, self.so, device=self.dev)
        self.head.forward(ws.y[-1], out)
        ws.running.flush()
        return out, ws

    def backward(self, ws, out, dout, need_dx, need_w, need_dz=False):
        m, N, sc = self.mod.model, self.N, self.scratch
        self.bank.ensure_packed()
        ws.sums_arena.zero()
        if need_w:
            self.bank.begin_backward()
        dyh = sc.get(self.g_dyh)
        if self.mod.use_sigmoid:
            ops.pack_nchw(dout, dyh, self.g_dyh, mul_out=out, mul_kind=L.ACT_SIGMOID, halo=L.HALO_ZERO)
        else:
            ops.pack_nchw(dout, dyh, self.g_dyh, halo=L.HALO_ZERO)
        if need_w:
            hs = ws.head_sums.t
            ops.norm_bwd_reduce(dyh, 2, dyh, self.g_dyh, sums=hs, count=0.0)
            accumulate_grad(self.head.bias, hs[0, :1, 0])      # before the layer's weight gradient (see _GenProgram.backward)
            self.head.backward_weight(dyh, ws.y[-1])
        g = sc.get(self.g_r[-1], "g%d" % (len(self.sizes) - 1))
        self.head.backward_data(dyh, g)
        for li in range(len(self.sizes) - 1, 0, -1):
            bn = m[self.idx[li] + 1]
            cnt = (N // self.groups) * self.sizes[li] ** 2
            dy = sc.get(self.g_r[li], "dy%d" % li)
            _norm_backward(g, 0, ws.r[li], self.g_r[li], ws.ns[li], L.ACT_LRELU, 0.2, cnt, dy, 0)
            if need_w:
                sums = ws.ns[li].sums[0] if self.groups == 1 else ws.ns[li].sums.sum(0)     # gamma / beta are shared by the groups
                accumulate_grad(bn.weight, sums[:, 1])
                accumulate_grad(bn.bias, sums[:, 0])
                self.convs[li].backward_weight(dy, ws.y[li - 1])
            g = sc.get(self.g_r[li - 1], "g%d" % (li - 1))
            self.convs[li].backward_data(dy, g)
        # layer 0: LeakyReLU backward from the sign of the stored output, bias gradient = sum
        dy = sc.get(self.g_r[0], "dy0")
        s0 = ws.l0_sums.t
        kw = dict(act=L.ACT_LRELU, act_slope=0.2, count=0.0, sums=s0)
        ops.norm_bwd_reduce(g, 0, ws.y[0], self.g_y[0], **kw)
        ops.norm_bwd_apply(g, 0, ws.y[0], self.g_y[0], dx=dy, dx_pad=0, **kw)
        if need_w:
            accumulate_grad(self.convs[0].bias, s0[0, :, 0])
            self.convs[0].backward_weight(dy, ws.x0)
        dx = dz = None
        if need_dx or need_dz:
            gx = sc.get(Geom(N, self.S, self.S, 8, 0), "gx")
            self.convs[0].backward_data(dy, gx)
            nc = self.mod.input_nc_img      # channel nc is the embedding plane (networks.py:780-782)
            if need_dz:
                dxz = torch.empty(N, nc + 1, self.S, self.S, device=self.dev)
                ops.unpack_resize_bwd(gx, Geom(N, self.S, self.S, 8, 0), dxz)
                dx = dxz[:, :nc].contiguous() if need_dx else None
                dz = dxz[:, nc].sum((1, 2))
            else:
                dx = torch.empty(N, nc, self.S, self.S, device=self.dev)
                ops.unpack_resize_bwd(gx, Geom(N, self.S, self.S, 8, 0), dx)
        return dx, dz


class NLayerDiscriminator(nn.Module):
    """Same constructor / forward / state_dict keys as models/networks.py:737-783 (BatchNorm2d, conditional z channel)."""

    def __init__(self, input_nc, nz, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=False):
        super().__init__()
        func = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
        if func is not nn.BatchNorm2d:
            raise NotImplementedError("pcgan_b200 NLayerDiscriminator: batch norm (the wsgan_emb configuration)")
        if ndf % 64 != 0 or input_nc + nz > 8 or nz != 1:
            raise NotImplementedError("ndf must be a multiple of 64, nz == 1, input_nc + nz <= 8")
        self.input_nc_img, self.ndf, self.n_layers, self.use_sigmoid = input_nc, ndf, n_layers, use_sigmoid
        seq = [nn.Conv2d(input_nc + nz, ndf, 4, stride=2, padding=1), nn.Identity()]
        prev = ndf
        for n in range(1, n_layers + 1):
            cur = ndf * min(2 ** n, 8)
            seq += [nn.Conv2d(prev, cur, 4, stride=2 if n < n_layers else 1, padding=1, bias=False), norm_layer(cur), nn.Identity()]
            prev = cur
        seq += [nn.Conv2d(prev, 1, 4, stride=1, padding=1)]
        if use_sigmoid:
            seq += [nn.Identity()]
        self.model = nn.Sequential(*seq)
        self._programs = {}
        self._groups = 1
        self._key = CO.register_module(self)

    def out_size(self, s):
        """side of the patch map for an s x s input: n_layers stride-2 convolutions, then two 4x4 stride-1 pad-1 ones"""
        return s // (2 ** self.n_layers) - 2

    def in_size(self, so):
        return (so + 2) * (2 ** self.n_layers)

    def _program(self, n, s):
        k = (n, s, self.model[0].weight.device, self._groups)
        if k not in self._programs:
            self._programs[k] = _DiscProgram(self, n, s, self._groups)
        return self._programs[k]

    def grouped(self, groups):
        """Context manager: the next forward calls treat their batch as `groups` independent passes batched together
        (cat of the inputs along dim 0): each run of N / groups samples has its own BatchNorm batch, exactly as separate
        calls would, but every layer is one launch."""
        return _Grouped(self, groups)

    def zero_wgrad(self):
        """Deferred weight gradients (defer_wgrad = True): clear the packed accumulators (optimizer.zero_grad time)."""
        for prog in self._programs.values():
            prog.bank.zero_wgrad()

    def flush_wgrad(self):
        """... and scatter them into the parameters' .grad after the backward sweeps of the step."""
        for prog in self._programs.values():
            prog.bank.flush_wgrad()

    def forward(self, input, z=None):
        _require_cuda(input, "NLayerDiscriminator")
        if z is None:
            raise NotImplementedError("NLayerDiscriminator needs the embedding z")
        if input.shape[2] != input.shape[3] or input.shape[2] % (2 ** self.n_layers):
            raise NotImplementedError("square inputs with side a multiple of %d" % 2 ** self.n_layers)
        out = torch.ops.pcgan.nlayer_discriminator(input, z, list(self.parameters()), self._key)
        CO.finish_forward(self._key, out)
        return out


class _Grouped:
    def __init__(self, mod, groups):
        self.mod, self.groups, self.prev = mod, int(groups), 1

    def __enter__(self):
        self.prev, self.mod._groups = self.mod._groups, self.groups
        return self.mod

    def __exit__(self, *exc):
        self.mod._groups = self.prev
        return False


# ---------------------------------------------------------------------------------------
# GANLoss and the scalar losses of the step
# ---------------------------------------------------------------------------------------
def l1_loss(a, b):
    """nn.L1Loss() (wsgan_emb_model.py:142,149): mean |a - b|; gradient flows to `a` only (b is data)."""
    return torch.ops.pcgan.reduce_loss(L.LOSS_L1, a, b.detach().contiguous().float(), 0)


def mse_loss(a, b):
    """nn.MSELoss() (wsgan_emb_model.py:148): mean (a - b)^2; gradient flows to `a` only."""
    return torch.ops.pcgan.reduce_loss(L.LOSS_MSE, a, b.detach().contiguous().float(), 0)


class GANLoss(nn.Module):
    """GANLoss (networks.py:386-420): per-sample targets (bool / int / list of them, one per batch element) expanded
    over the prediction map; nn.BCELoss on the sigmoid output (log clamped at -100) or nn.MSELoss when use_lsgan."""

    def __init__(self, use_lsgan=True, tensor=torch.FloatTensor):
        super().__init__()
        self.kind = L.LOSS_MSE if use_lsgan else L.LOSS_BCE
        self.Tensor = tensor
        self._const = {}

    def get_target_tensor(self, input, target_label):
        """One target value per sample (networks.py:395-405).  bool / int: a cached device constant; list: uploaded;
        a float device tensor (already per sample) is used as is, so a captured step needs no host-to-device copy."""
        if isinstance(target_label, torch.Tensor):
            return target_label.to(device=input.device, dtype=torch.float32).reshape(-1).contiguous()
        if not isinstance(target_label, list):
            key = (float(target_label), input.size(0), input.device)
            t = self._const.get(key)
            if t is None:
                t = self._const[key] = torch.full((input.size(0),), key[0], device=input.device)
            return t
        vals = [float(t) for t in target_label]      # bools become 1 / 0, anything else passes through (:399-403)
        return torch.tensor(np.array(vals, dtype=np.float32), device=input.device)

    def __call__(self, inputs, target_label):
        if not isinstance(inputs, list):
            inputs = [inputs]
        loss = 0.0
        for inp in inputs:
            if inp.dim() < 4:
                inp = inp.view(inp.size(0), -1, 1, 1)
            _require_cuda(inp, "GANLoss")
            t = self.get_target_tensor(inp, target_label)
            n = inp.size(0)
            if t.numel() == 1 and n > 1:
                t = t.expand(n).contiguous()
            loss = loss + torch.ops.pcgan.reduce_loss(self.kind, inp, t, inp.numel() // n)
        return loss


class BinaryNLLLoss(nn.Module):
    """Elo pairwise loss (networks.py:473-482): target LUT[label] in {0, .5, 1}, -mean(t log(p+1e-20) + (1-t) log(1-p+1e-20)).
    from_score fuses the torch.sigmoid the trainer applies first (siamese.py:675) into the same reduction kernel."""

    def __init__(self):
        super().__init__()
        self._lut = {}

    def _target(self, ref, label):
        lut = self._lut.get(ref.device)
        if lut is None:
            lut = self._lut[ref.device] = torch.tensor([0.0, 0.5, 1.0], device=ref.device)
        return lut[label.to(ref.device)].contiguous()

    def __call__(self, prob, label):
        return torch.ops.pcgan.reduce_loss(L.LOSS_ELO_NLL, prob, self._target(prob, label), prob.numel() // prob.size(0))

    def from_score(self, score, label):
        return torch.ops.pcgan.reduce_loss(L.LOSS_ELO_NLL_SCORE, score, self._target(score, label), score.numel() // score.size(0))


# ---------------------------------------------------------------------------------------
# factory (networks.py:22-34, 72-175)
# ---------------------------------------------------------------------------------------
class IdentityMapping(nn.Module):
    def __init__(self, *args):
        super().__init__()

    def forward(self, x):
        return x


def get_norm_layer(norm_type="instance"):
    if norm_type == "batch":
        return functools.partial(nn.BatchNorm2d, affine=True)
    if norm_type == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=True)
    raise NotImplementedError("normalization layer [%s] is outside the wsgan_emb hot path" % norm_type)


def init_weights(net, init_type="normal", gain=0.02):
    """init_weights (networks.py:72-93): Conv/Linear weights by init_type, biases 0, BatchNorm2d weight N(1, gain)."""
    def init_func(m):
        cn = m.__class__.__name__
        if hasattr(m, "weight") and (cn.find("Conv") != -1 or cn.find("Linear") != -1):
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=gain)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=gain)
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if getattr(m, "bias", None) is not None:
                init.constant_(m.bias.data, 0.0)
        elif cn.find("BatchNorm2d") != -1:
            init.normal_(m.weight.data, 1.0, gain)
            init.constant_(m.bias.data, 0.0)

    print("initialize network with %s" % init_type)
    net.apply(init_func)


class LocalDataParallel(torch.nn.DataParallel):
    """What init_net returns for a non-empty gpu_ids (networks.py:96-102): the reference wraps every network in
    nn.DataParallel and tests isinstance(net, DataParallel) / unwraps .module (base_model.py:104,127;
    wsgan_emb_model.py:118,132).  This subclass keeps those call sites working with one process per GPU: it runs the
    module on its single device; gradient averaging across ranks is done by pcgan_b200.dist.GradSync."""

    def __init__(self, module, device_ids):
        super().__init__(module, device_ids=[device_ids[0]])

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)


def init_net(net, init_type="normal", gpu_ids=[]):
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.to(gpu_ids[0])
        net = LocalDataParallel(net, gpu_ids)
    init_weights(net, init_type)
    return net


def define_G(input_nc, output_nc, nz, ngf, which_model_netG="unet_128", norm="batch", nl="relu", dropout=0, init_type="xavier",
             gpu_ids=[], upsample="bilinear", size=512, embed_size=256, n_layers_G=7):
    norm_layer = get_norm_layer(norm_type=norm)
    if which_model_netG == "resnet_9blocks":
        net = ResnetGenerator(input_nc, output_nc, nz, ngf, norm_layer=norm_layer, dropout=dropout, n_blocks=9)
    elif which_model_netG == "resnet_6blocks":
        net = ResnetGenerator(input_nc, output_nc, nz, ngf, norm_layer=norm_layer, dropout=dropout, n_blocks=6)
    elif which_model_netG in ("unet_128", "unet_256", "unet"):
        downs = {"unet_128": 7, "unet_256": 8, "unet": n_layers_G}[which_model_netG]
        net = UnetGenerator(input_nc, output_nc, nz, downs, ngf, norm_layer=norm_layer, dropout=dropout)
    else:
        raise NotImplementedError("Generator [%s] is outside the wsgan_emb hot path (resnet_9blocks / resnet_6blocks / unet_128 / unet_256 / unet)" % which_model_netG)
    return init_net(net, init_type, gpu_ids)


def define_D(input_nc, nz, ndf, which_model_netD, n_layers_D=3, norm="batch", use_sigmoid=False, init_type="normal", num_Ds=1,
             gpu_ids=[], use_projection=True, size=512, embed_size=256, num_classes=1):
    norm_layer = get_norm_layer(norm_type=norm)
    if which_model_netD == "basic":
        net = NLayerDiscriminator(input_nc, nz, ndf, n_layers=3, norm_layer=norm_layer, use_sigmoid=use_sigmoid)
    elif which_model_netD == "n_layers":
        net = NLayerDiscriminator(input_nc, nz, ndf, n_layers_D, norm_layer=norm_layer, use_sigmoid=use_sigmoid)
    else:
        raise NotImplementedError("Discriminator [%s] is outside the wsgan_emb hot path (basic / n_layers)" % which_model_netD)
    return init_net(net, init_type, gpu_ids)


# ---------------------------------------------------------------------------------------
# Elo / siamese encoder: SiameseFeature(ResNetFeature(resnet18)) (networks.py:1008-1083, :1310-1359; resnet.py)
# ---------------------------------------------------------------------------------------
class _EncWorkspace:
    pass


def _bn_forward(conv: ConvRT, xbuf, rbuf, ns: NormState, count, bn, running: RunningStats = None, mask=None, explicit=False):
    """conv -> [Dropout2d mask] -> BatchNorm2d statistics (training mode, running stats updated, resnet.py:57-59).
    Normally the statistics are finalized by the consuming norm_apply and the running statistics by the pass's batched
    launch (`running`).  explicit=True (the shortcut BatchNorm, whose scale / shift the block's last norm_apply needs as
    plain arrays) runs the finalize kernel for scale / shift and still leaves the running statistics to `running`; with a
    dropout mask the convolution emits per-sample statistics that the finalize kernel combines with the mask (and then
    updates the running statistics itself).  ns.groups > 1: each group of samples is its own BatchNorm batch."""
    ns.affine = True
    if not ns.pooled:
        ns.stats.zero_()
    conv.forward(xbuf, rbuf, ns.stats)
    seq = ns.groups > 1
    if mask is None and ns.stats_groups == ns.groups and running is not None:
        running.add(ns, count, bn.running_mean, bn.running_var, bn.num_batches_tracked, MOMENTUM, seq)
        if not explicit:
            ns.fused = (count, bn.weight.detach(), bn.bias.detach())
            return
        ns.fused = None
        ops.norm_finalize(ns.stats, ns.groups, ns.c, count, eps=EPS, momentum=MOMENTUM, gamma=bn.weight.detach(), beta=bn.bias.detach(),
                          mean=ns.mean, rstd=ns.rstd, scale=ns.scale, shift=ns.shift)
        return
    if seq:
        raise L.PcganError("grouped passes with channel dropout are not supported (run the passes one by one)")
    ns.fused = None
    kw = dict(eps=EPS, momentum=MOMENTUM, gamma=bn.weight.detach(), beta=bn.bias.detach(), mean=ns.mean, rstd=ns.rstd,
              scale=ns.scale, shift=ns.shift, running_mean=bn.running_mean, running_var=bn.running_var)
    if ns.stats_groups > ns.groups:
        ops.norm_finalize(ns.stats, 1, ns.c, count, drop_mask=mask, in_groups=ns.stats_groups, **kw)
    else:
        assert mask is None
        ops.norm_finalize(ns.stats, ns.groups, ns.c, count, **kw)
    bn.num_batches_tracked += 1


def _bn_param_grads(bn, ns: NormState):
    """dgamma = sum g*xhat, dbeta = sum g: by-products of the backward reduction (summed over the groups of a grouped pass)."""
    sums = ns.sums[0] if ns.groups == 1 else ns.sums.sum(0)
    if bn.weight.requires_grad:
        accumulate_grad(bn.weight, sums[:, 1])
    if bn.bias.requires_grad:
        accumulate_grad(bn.bias, sums[:, 0])


class _EncBlock:
    """One BasicBlock (resnet.py:31-73) at fixed geometry: conv3x3(stride) [drop] BN ReLU conv3x3 [drop] BN
    (+ 1x1 stride-s conv BN) add ReLU.  dropout=True plans per-sample statistics for the two dropped convolutions."""

    def __init__(self, name, holder, N, hin, cin, c, stride, dropout=False, groups=1):
        G = Geom
        h = hin // stride
        self.name, self.holder, self.N, self.hin, self.cin, self.h, self.c, self.stride = name, holder, N, hin, cin, h, c, stride
        self.dropout, self.groups = dropout, groups
        if dropout and groups > 1:
            raise L.PcganError("grouped passes with channel dropout are not supported")
        grp = dict(per_sample_stats=True, stats_div=N // groups) if groups > 1 else dict(per_sample_stats=dropout)
        self.g_x, self.g_xr = G(N, hin, hin, cin, 1), G(N, hin, hin, cin, 0)
        self.g_r, self.g_y = G(N, h, h, c, 0), G(N, h, h, c, 1)
        dy1 = self.g_y if stride == 1 else self.g_r
        self.c1 = ConvRT(name + ".conv1", holder.conv1.weight, None, self.g_x, stride, 1, OutMap.nhwc(self.g_r), stats=True,
                         dyg=dy1, dx_out=OutMap.nhwc(self.g_xr), want_wgrad=False, **grp)
        self.c2 = ConvRT(name + ".conv2", holder.conv2.weight, None, self.g_y, 1, 1, OutMap.nhwc(self.g_r), stats=True,
                         dyg=self.g_y, dx_out=OutMap.nhwc(self.g_r), want_wgrad=False, **grp)
        self.dy1_pad = dy1.pad
        self.ds = None
        if holder.downsample is not None:
            gd = dict(per_sample_stats=True, stats_div=N // groups) if groups > 1 else {}
            self.ds = ConvRT(name + ".downsample.0", holder.downsample[0].weight, None, self.g_x, stride, 0, OutMap.nhwc(self.g_r),
                             stats=True, dyg=self.g_r, dx_out=OutMap.nhwc(self.g_xr), want_wgrad=False, **gd)

    def new_ws(self, dev, stats_arena=None, sums_arena=None):
        w = _EncWorkspace()
        w.ra, w.h, w.rb, w.y = zeros_act(self.g_r, dev), zeros_act(self.g_y, dev), zeros_act(self.g_r, dev), zeros_act(self.g_y, dev)
        sg = self.N if self.dropout else None
        w.na = NormState(self.groups, self.c, dev, stats_arena, sums_arena, stats_groups=sg)
        w.nb = NormState(self.groups, self.c, dev, stats_arena, sums_arena, stats_groups=sg)
        w.m1 = w.m2 = None
        if self.ds is not None:
            w.rd, w.nd = zeros_act(self.g_r, dev), NormState(self.groups, self.c, dev, stats_arena, sums_arena)
        return w

    def forward(self, xbuf, w, masks=None, running=None):
        hd, cnt = self.holder, (self.N // self.groups) * self.h * self.h
        w.m1, w.m2 = masks if masks is not None else (None, None)
        _bn_forward(self.c1, xbuf, w.ra, w.na, cnt, hd.bn1, running, w.m1)
        ops.norm_apply(w.ra, self.g_r, w.h, self.g_y, y_halo=L.HALO_ZERO, **w.na.apply_kw(),
                       drop_mask=w.m1, act=L.ACT_RELU)
        _bn_forward(self.c2, w.h, w.rb, w.nb, cnt, hd.bn2, running, w.m2)
        if self.ds is not None:
            _bn_forward(self.ds, xbuf, w.rd, w.nd, cnt, hd.downsample[1], running, explicit=True)     # res_scale / res_shift below
            ops.norm_apply(w.rb, self.g_r, w.y, self.g_y, y_halo=L.HALO_ZERO, **w.nb.apply_kw(),
                           drop_mask=w.m2, res=w.rd, res_pad=0, res_scale=w.nd.scale, res_shift=w.nd.shift, res_groups=self.groups, act=L.ACT_RELU)
        else:
            ops.norm_apply(w.rb, self.g_r, w.y, self.g_y, y_halo=L.HALO_ZERO, **w.nb.apply_kw(),
                           drop_mask=w.m2, res=xbuf, res_pad=1, act=L.ACT_RELU)
        return w.y

    def backward(self, xbuf, w, gy, sc, need_w=False):
        """gy: gradient of the block output (unpadded).  Returns the gradient of the block input (unpadded);
        need_w: also accumulate the weight / BatchNorm parameter gradients."""
        hd, cnt = self.holder, (self.N // self.groups) * self.h * self.h
        dyb = sc.get(self.g_y, self.name + "dyb")
        gres = sc.get(self.g_r, self.name + "gres")
        if self.ds is not None:
            res_kw = dict(res=w.rd, res_pad=0, res_scale=w.nd.scale, res_shift=w.nd.shift, res_groups=self.groups)
        else:
            res_kw = dict(res=xbuf, res_pad=1)
        _norm_backward(gy, 0, w.rb, self.g_r, w.nb, L.ACT_RELU, 0.0, cnt, dyb, 1, dres=gres, dres_pad=0, drop_mask=w.m2, **res_kw)
        if need_w:
            _bn_param_grads(hd.bn2, w.nb)
            self.c2.backward_weight(dyb, w.h)
        gh = sc.get(self.g_r, self.name + "gh")
        self.c2.backward_data(dyb, gh)
        dya = sc.get(self.g_y if self.stride == 1 else self.g_r, self.name + "dya")
        _norm_backward(gh, 0, w.ra, self.g_r, w.na, L.ACT_RELU, 0.0, cnt, dya, self.dy1_pad, drop_mask=w.m1)
        if need_w:
            _bn_param_grads(hd.bn1, w.na)
            self.c1.backward_weight(dya, xbuf)
        gx1 = sc.get(self.g_xr, self.name + "gx1")
        self.c1.backward_data(dya, gx1)
        gx = sc.get(self.g_xr, self.name + "gx")
        if self.ds is not None:
            dyd = sc.get(self.g_r, self.name + "dyd")
            _norm_backward(gres, 0, w.rd, self.g_r, w.nd, L.ACT_NONE, 0.0, cnt, dyd, 0)
            if need_w:
                _bn_param_grads(hd.downsample[1], w.nd)
                self.ds.backward_weight(dyd, xbuf)
            gx2 = sc.get(self.g_xr, self.name + "gx2")  # odd phases of a 1x1 stride-2 conv receive nothing: stay zero
            self.ds.backward_data(dyd, gx2)
            _add_interior(gx1, self.g_xr, gx2, 0, gx)
        else:
            _add_interior(gx1, self.g_xr, gres, 0, gx)
        return gx


class _EncHead:
    """conv3x3 512->nf (+bias) BN [drop] LeakyReLU conv3x3 nf->1 (+bias) global average pool (networks.py:1014-1027,
    1056-1057); instantiated for `cnn` and, in noisy mode, for the twin `cnn_logvar` (:1034-1049, 1060-1066)."""

    def __init__(self, name, seq, N, hf, slope, groups=1):
        G = Geom
        self.name, self.seq, self.N, self.hf, self.slope, self.groups = name, seq, N, hf, slope, groups
        grp = dict(per_sample_stats=True, stats_div=N // groups) if groups > 1 else {}
        nf = seq[0].weight.shape[0]
        self.nf = nf
        self.g_f = G(N, hf, hf, 512, 1)
        self.g_rh, self.g_hh = G(N, hf, hf, nf, 0), G(N, hf, hf, nf, 1)
        self.g_fin, self.g_dyfin = G(N, hf, hf, 8, 0), G(N, hf, hf, 8, 1)
        self.c1 = ConvRT(name + ".0", seq[0].weight, seq[0].bias, self.g_f, 1, 1, OutMap.nhwc(self.g_rh), stats=True,
                         dyg=self.g_hh, dx_out=OutMap.nhwc(G(N, hf, hf, 512, 0)), want_wgrad=False, **grp)
        # last conv: the global average pool is its per-sample statistics (sum over the map) / (h*w)
        self.c2 = ConvRT(name + ".4", seq[4].weight, seq[4].bias, self.g_hh, 1, 1, OutMap.nhwc(self.g_fin), stats=True,
                         per_sample_stats=True, dyg=self.g_dyfin, dx_out=OutMap.nhwc(self.g_rh), want_wgrad=False)

    def new_ws(self, dev, stats_arena, sums_arena):
        w = _EncWorkspace()
        w.rh, w.hh, w.fin = zeros_act(self.g_rh, dev), zeros_act(self.g_hh, dev), zeros_act(self.g_fin, dev)
        w.nh = NormState(self.groups, self.nf, dev, stats_arena, sums_arena)
        w.nfin = NormState(self.N, 1, dev, stats_arena, sums_arena)
        w.pm = None
        return w

    def forward(self, fbuf, w, mask=None, running=None):
        N, hf = self.N, self.hf
        w.pm = mask
        _bn_forward(self.c1, fbuf, w.rh, w.nh, (N // self.groups) * hf * hf, self.seq[1], running)
        ops.norm_apply(w.rh, self.g_rh, w.hh, self.g_hh, y_halo=L.HALO_ZERO, **w.nh.apply_kw(),
                       act=L.ACT_LRELU, act_slope=self.slope, post_mask=mask)
        self.c2.forward(w.hh, w.fin, w.nfin.stats)
        return (w.nfin.stats[:, 0, 0] / float(hf * hf)).view(N, 1, 1, 1)

    def backward(self, fbuf, w, gy, sc, gf_out, need_w=False):
        """gy [N,1,1,1] -> gradient of the trunk features written to gf_out (unpadded [N, hf, hf, 512])."""
        N, hf = self.N, self.hf
        gmap = (gy.reshape(N, 1, 1, 1) / float(hf * hf)).expand(N, 1, hf, hf).contiguous()
        dyf = sc.get(self.g_dyfin, self.name + "dyf")
        ops.pack_nchw(gmap, dyf, self.g_dyfin, halo=L.HALO_ZERO)
        if need_w:
            self.c2.backward_weight(dyf, w.hh)
            if self.seq[4].bias.requires_grad:
                accumulate_grad(self.seq[4].bias, gy.reshape(N).sum().reshape(1))
        ghh = sc.get(self.g_rh, self.name + "ghh")
        self.c2.backward_data(dyf, ghh)
        dyh = sc.get(self.g_hh, self.name + "dyh")
        _norm_backward(ghh, 0, w.rh, self.g_rh, w.nh, L.ACT_LRELU, self.slope, (N // self.groups) * hf * hf, dyh, 1, post_mask=w.pm)
        if need_w:
            _bn_param_grads(self.seq[1], w.nh)
            self.c1.backward_weight(dyh, fbuf)
            b = self.seq[0].bias      # feeds a BatchNorm: its gradient is exactly zero
            if b.requires_grad and b.grad is None:
                b.grad = torch.zeros_like(b)
        self.c1.backward_data(dyh, gf_out)


class _EncProgram:
    def __init__(self, mod, N, S, groups=1):
        """groups > 1: N samples = `groups` independent passes batched together (own BatchNorm batch per run of N / groups
        samples), as _DiscProgram; not available with channel dropout (the Monte-Carlo passes draw masks per pass)."""
        if S % 32:
            raise NotImplementedError("encoder input side must be a multiple of 32 (got %d)" % S)
        if N % groups:
            raise L.PcganError("SiameseFeature: batch %d is not divisible into %d groups" % (N, groups))
        self.mod, self.N, self.S, self.groups = mod, N, S, groups
        rn = mod.base.model
        dev = rn.conv1.weight.device
        self.dev = dev
        self.p_drop = float(mod.base.dropout)
        drop = self.p_drop > 0
        G = Geom
        s2, s4 = S // 2, S // 4
        self.g_x0 = G(N, S, S, 8, 3)
        self.g_r0, self.g_a0 = G(N, s2, s2, 64, 0), G(N, s2, s2, 64, 0)
        self.g_dy0 = G(N, s2, s2, 64, 2)     # zero-haloed: the stem's data gradient runs in shift-sum form over this grid
        self.g_p, self.g_pr = G(N, s4, s4, 64, 1), G(N, s4, s4, 64, 0)
        grp = dict(per_sample_stats=True, stats_div=N // groups) if groups > 1 else {}
        self.stem = ConvRT("E.conv1", rn.conv1.weight, None, self.g_x0, 2, 3, OutMap.nhwc(self.g_r0), stats=True, dyg=self.g_dy0,
                           dx_out=OutMap.nhwc(G(N, S, S, 8, 0)), want_wgrad=False, **grp)
        self.blocks = []
        hin, cin = s4, 64
        for li, c in enumerate((64, 128, 256, 512), start=1):
            layer = getattr(rn, "layer%d" % li)
            for bi in range(2):
                stride = 2 if (li > 1 and bi == 0) else 1
                blk = _EncBlock("E.layer%d.%d" % (li, bi), layer[bi], N, hin, cin, c, stride, dropout=drop, groups=groups)
                self.blocks.append(blk)
                hin, cin = hin // stride, c
        self.hf = hin
        self.g_ff = G(N, hin, hin, 512, 0)
        self.heads = [_EncHead("E.cnn", mod.cnn, N, hin, mod.cnn_relu_slope, groups)]
        if mod._noisy:
            self.heads.append(_EncHead("E.cnn_logvar", mod.cnn_logvar, N, hin, mod.cnn_relu_slope, groups))
        if groups > 1 and mod.head_dropout > 0:
            raise L.PcganError("grouped passes with channel dropout are not supported")
        self.head_drop = mod.head_dropout > 0
        self.scratch = _Scratch(dev)
        self.pool = Pool(lambda key: self._new_ws())
        convs = [self.stem]
        for b in self.blocks:
            convs += [b.c1, b.c2] + ([b.ds] if b.ds is not None else [])
        for h in self.heads:
            convs += [h.c1, h.c2]
        self.bank = WeightBank(convs, dev)

    def _new_ws(self):
        ws, dev = _EncWorkspace(), self.dev
        ws.x0, ws.r0, ws.a0, ws.p = zeros_act(self.g_x0, dev), zeros_act(self.g_r0, dev), zeros_act(self.g_a0, dev), zeros_act(self.g_p, dev)
        ws.idx = torch.zeros(self.g_pr.numel, dtype=torch.uint8, device=dev)
        ws.stats_arena, ws.sums_arena = Arena(dev), Arena(dev)
        ws.n0 = NormState(self.groups, 64, dev, ws.stats_arena, ws.sums_arena)
        ws.blk = [b.new_ws(dev, ws.stats_arena, ws.sums_arena) for b in self.blocks]
        ws.heads = [h.new_ws(dev, ws.stats_arena, ws.sums_arena) for h in self.heads]
        ws.stats_arena.finalize()
        ws.sums_arena.finalize()
        ws.running = RunningStats(dev)
        return ws

    def _mask(self, c, p):
        """nn.Dropout2d(p) in training mode: one Bernoulli(1-p) draw per (sample, channel), kept channels scaled by
        1/(1-p).  Tests inject the draws through SiameseFeature.dropout_masks (consumed in module order)."""
        q = self.mod.dropout_masks
        if q:
            m = q.pop(0).to(self.dev, torch.float32).reshape(self.N, c).contiguous()
            return m
        keep = torch.full((self.N, c), 1.0 - p, device=self.dev)
        return torch.bernoulli(keep) / (1.0 - p)

    def forward(self, x):
        mod, N = self.mod, self.N
        rn = mod.base.model
        ws = self.pool.take(0)
        self.bank.ensure_packed()
        ws.stats_arena.zero()
        ws.running.begin()
        ops.pack_nchw(x, ws.x0, self.g_x0, halo=L.HALO_ZERO)
        s2 = self.S // 2
        _bn_forward(self.stem, ws.x0, ws.r0, ws.n0, (N // self.groups) * s2 * s2, rn.bn1, ws.running)
        ops.norm_apply(ws.r0, self.g_r0, ws.a0, self.g_a0, **ws.n0.apply_kw(), act=L.ACT_RELU)
        ops.maxpool_fwd(ws.a0, self.g_a0, ws.p, 1, ws.idx)
        cur = ws.p
        for blk, w in zip(self.blocks, ws.blk):
            masks = (self._mask(blk.c, self.p_drop), self._mask(blk.c, self.p_drop)) if self.p_drop > 0 else None
            cur = blk.forward(cur, w, masks, ws.running)
        outs = []
        for h, w in zip(self.heads, ws.heads):
            outs.append(h.forward(cur, w, self._mask(h.nf, mod.head_dropout) if self.head_drop else None, ws.running))
        ws.running.flush()
        return outs, ws

    def backward(self, ws, gys, need_dx, need_w):
        """gys: one gradient (or None) per head output."""
        N, sc, hf = self.N, self.scratch, self.hf
        rn = self.mod.base.model
        self.bank.ensure_packed()
        ws.sums_arena.zero()
        feats = ws.blk[-1].y
        g = None
        for i, (h, w, gy) in enumerate(zip(self.heads, ws.heads, gys)):
            if gy is None:
                continue
            gf = sc.get(self.g_ff, "gf%d" % i)
            h.backward(feats, w, gy.contiguous().float(), sc, gf, need_w)
            if g is None:
                g = gf
            else:
                gsum = sc.get(self.g_ff, "gfsum")
                ops.halo_fold(g, self.g_ff, gsum, 0, halo=L.HALO_ZERO, add=gf, add_pad=0)
                g = gsum
        if g is None:
            return None
        inputs = [ws.p] + [w.y for w in ws.blk[:-1]]
        for blk, w, xin in zip(reversed(self.blocks), reversed(ws.blk), reversed(inputs)):
            g = blk.backward(xin, w, g, sc, need_w)
        s2 = self.S // 2
        ga0 = sc.get(self.g_a0, "ga0")
        ops.maxpool_bwd(g, 0, ws.idx, ga0, 0, N, s2, s2, 64)
        dy0 = sc.get(self.g_dy0, "dy0")
        _norm_backward(ga0, 0, ws.r0, self.g_r0, ws.n0, L.ACT_RELU, 0.0, (N // self.groups) * s2 * s2, dy0, self.g_dy0.pad)
        if need_w:
            _bn_param_grads(rn.bn1, ws.n0)
            self.stem.backward_weight(dy0, ws.x0)
        if not need_dx:
            return None
        gx = sc.get(Geom(N, self.S, self.S, 8, 0), "gx")
        self.stem.backward_data(dy0, gx)
        dx = torch.empty(N, 3, self.S, self.S, device=self.dev)
        ops.unpack_resize_bwd(gx, Geom(N, self.S, self.S, 8, 0), dx)
        return dx


class _BasicBlockHolder(nn.Module):
    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.drop1 = IdentityMapping()
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.Identity()
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.drop2 = IdentityMapping()
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample


class _ResNet18Holder(nn.Module):
    """Parameter layout of models/resnet.py ResNet(BasicBlock, [2, 2, 2, 2]) without fc (deleted at networks.py:1337)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu, self.maxpool = nn.Identity(), nn.Identity()
        inpl = 64
        for li, planes in enumerate((64, 128, 256, 512), start=1):
            stride = 1 if li == 1 else 2
            ds = None
            if stride != 1 or inpl != planes:
                ds = nn.Sequential(nn.Conv2d(inpl, planes, 1, stride=stride, bias=False), nn.BatchNorm2d(planes))
            setattr(self, "layer%d" % li, nn.Sequential(_BasicBlockHolder(inpl, planes, stride, ds), _BasicBlockHolder(planes, planes)))
            inpl = planes
        self.avgpool = nn.Identity()


class ResNetFeature(nn.Module):
    """networks.py:1310-1359 for resnet18; dropout > 0 = nn.Dropout2d after every BasicBlock convolution (resnet.py:40-49),
    always live (the reference never calls .eval(): that is how its MC-dropout works)."""

    def __init__(self, input_nc=3, which_model="resnet18", dropout=0.0):
        super().__init__()
        if which_model != "resnet18" or input_nc != 3:
            raise NotImplementedError("pcgan_b200 ResNetFeature: resnet18 on RGB input (the wsgan_emb configuration)")
        self.model = _ResNet18Holder()
        self.feature_dim = 512
        self.dropout = float(dropout)

    def load_pretrained(self, state_dict):
        if isinstance(state_dict, str):
            state_dict = torch.load(state_dict)
        self.model.load_state_dict(state_dict, strict=False)


def _head_holder(feature_dim, nf, cnn_pad):
    return nn.Sequential(nn.Conv2d(feature_dim, nf, 3, padding=cnn_pad), nn.BatchNorm2d(nf), IdentityMapping(),
                         nn.Identity(), nn.Conv2d(nf, 1, 3, padding=cnn_pad))


class SiameseFeature(nn.Module):
    """Same constructor / forward / load_pretrained / state_dict keys as models/networks.py:1008-1083 (pooling 'avg',
    cnn_dim = [nf, 1]).  Always runs with batch statistics, as the reference does (no .eval() on the train path).
    noisy=True adds the twin `cnn_logvar` head and forward returns (y, logvar); drop_layer = nn.Dropout2d partial puts a
    channel dropout between the head's BatchNorm and LeakyReLU (networks.py:1021-1023)."""

    def __init__(self, base=None, pooling="avg", cnn_dim=[], cnn_pad=1, cnn_relu_slope=0.2, noisy=False, drop_layer=None):
        super().__init__()
        if pooling != "avg" or len(cnn_dim) != 2 or cnn_dim[1] != 1 or cnn_pad != 1 or cnn_dim[0] % 8 or cnn_dim[0] >= 64:
            raise NotImplementedError("pcgan_b200 SiameseFeature: pooling='avg', cnn_dim=[nf<64 (multiple of 8), 1], cnn_pad=1")
        self.pooling, self.base, self._noisy, self.cnn_relu_slope = pooling, base, bool(noisy), cnn_relu_slope
        self.head_dropout = 0.0
        if isinstance(drop_layer, functools.partial) and drop_layer.func is nn.Dropout2d:
            self.head_dropout = float(drop_layer.keywords.get("p", 0.5))
        nf = cnn_dim[0]
        self.cnn = _head_holder(base.feature_dim, nf, cnn_pad)
        if self._noisy:
            self.cnn_logvar = _head_holder(base.feature_dim, nf, cnn_pad)
        self.feature_dim = 1
        self.dropout_masks = []      # test hook: explicit Dropout2d draws, consumed in module order
        self._programs = {}
        self._groups = 1
        self._key = CO.register_module(self)
        self._last_size = 0

    def _program(self, n, s):
        k = (n, s, self.cnn[0].weight.device, self._groups)
        if k not in self._programs:
            self._programs[k] = _EncProgram(self, n, s, self._groups)
        return self._programs[k]

    def grouped(self, groups):
        """Context manager: see NLayerDiscriminator.grouped."""
        return _Grouped(self, groups)

    def can_group(self):
        """Passes can be batched unless they draw Dropout2d masks (each Monte-Carlo pass has its own)."""
        return float(self.base.dropout) == 0.0 and self.head_dropout == 0.0

    def _run(self, x):
        _require_cuda(x, "SiameseFeature")
        if x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise NotImplementedError("square RGB inputs")
        self._last_size = int(x.shape[2])
        y, lv = torch.ops.pcgan.siamese_feature(x, list(self.parameters()), self._key)
        CO.finish_forward(self._key, y)
        return (y, lv) if self._noisy else (y,)

    def forward(self, x):
        outs = self._run(x)
        return (outs[0], outs[1]) if self._noisy else outs[0]

    def load_pretrained(self, state_dict):
        if isinstance(state_dict, str):
            state_dict = torch.load(state_dict)
        for key in list(state_dict.keys()):
            if key.startswith("cxn") or key.startswith("fc"):
                state_dict.pop(key)
        self.load_state_dict(state_dict, strict=True)

    def load_base(self, state_dict):
        self.base.load_pretrained(state_dict)


class SiameseNetwork(SiameseFeature):
    """The Elo rating trainer's network (networks.py:872-1005) in the configuration siamese.py builds by default
    (:445-449: cnn_dim=[32, 1], fc_dim=[], no cxn): two passes of the shared encoder, each with its own BatchNorm batch,
    score = rating(input1) - rating(input2) (:971-992).  Same state_dict keys as SiameseFeature (base.*, cnn.*)."""

    def __init__(self, base=None, pooling="avg", cnn_dim=[], cnn_pad=1, cnn_relu_slope=0.5, fc_dim=[], fc_relu_slope=0.2,
                 fc_residual=True, dropout=0.5, use_cxn=False, noisy=False, drop_layer=None, rsample=False):
        if fc_dim or use_cxn:
            raise NotImplementedError("pcgan_b200 SiameseNetwork: fc_dim=[] and use_cxn=False (the siamese.py default)")
        super().__init__(base, pooling, cnn_dim, cnn_pad, cnn_relu_slope, noisy, drop_layer)
        self._rsample = rsample

    def forward_once(self, x):
        outs = self._run(x)
        return (outs[0], outs[1]) if self._noisy else (outs[0], None)

    def forward(self, input1, input2):
        feature1, logvar1 = self.forward_once(input1)
        feature2, logvar2 = self.forward_once(input2)
        output = feature1 - feature2
        if self._noisy and self._rsample:
            return feature1, feature2, logvar1, logvar2
        if self._noisy:
            std_ = torch.sqrt(torch.exp(logvar1) + torch.exp(logvar2))
            return feature1, feature2, output, std_
        return feature1, feature2, output

    def load_pretrained(self, state_dict):
        self.base.load_pretrained(state_dict)

    def get_finetune_parameters(self):
        return self.cnn.parameters()


def get_dropout_layer(dropout=0.0):
    if dropout > 0:
        return functools.partial(nn.Dropout2d, p=dropout)
    return IdentityMapping


def define_E(which_model_netE, input_nc=3, init_type="kaiming", pooling="max", cnn_dim=[], cnn_pad=1, cnn_relu_slope=0.2,
             gpu_ids=[], fine_size_E=224, noisy=False, bnn_dropout=0.0):
    if "resnet" not in which_model_netE:
        raise NotImplementedError("Encoder [%s] is outside the wsgan_emb hot path (resnet18)" % which_model_netE)
    base = ResNetFeature(input_nc=input_nc, which_model=which_model_netE, dropout=bnn_dropout)
    net = SiameseFeature(base, pooling=pooling, cnn_dim=cnn_dim, cnn_pad=cnn_pad, cnn_relu_slope=cnn_relu_slope, noisy=noisy,
                         drop_layer=get_dropout_layer(bnn_dropout))
    return init_net(net, init_type, gpu_ids)


# ---------------------------------------------------------------------------------------
# Identity-preserving network: AlexNetFeature (networks.py:1218-1255), define_IP (:179-193)
# ---------------------------------------------------------------------------------------
class _IPWorkspace:
    pass


class _IPProgram:
    """AlexNet feature extractor (pooling 'None') at fixed (N, S): conv11x11 s4 p2 + ReLU, pool, conv5x5 p2 + ReLU, pool,
    three 3x3 convolutions + ReLU, pool (nn.MaxPool2d(3, 2): no padding); frozen: forward and input gradient only.  The
    ReLUs are fused into the convolution epilogues; their backward reads the sign of the stored outputs."""

    def __init__(self, mod, N, S):
        if (S + 4) % 4:
            raise NotImplementedError("AlexNetFeature input side must be a multiple of 4 (got %d)" % S)
        self.mod, self.N, self.S = mod, N, S
        f = mod.features
        dev = f[0].weight.device
        self.dev = dev
        G = Geom
        h1 = (S + 4 - 11) // 4 + 1
        p1, p2 = ops.pool_out(h1, 0), ops.pool_out(ops.pool_out(h1, 0), 0)
        p3 = ops.pool_out(p2, 0)
        self.h1, self.p1, self.p2, self.p3 = h1, p1, p2, p3
        self.g_x0 = G(N, S, S, 8, 2)
        self.g_r1, self.g_d1 = G(N, h1, h1, 64, 0), G(N, h1, h1, 64, 2)
        self.g_p1, self.g_p1r = G(N, p1, p1, 64, 2), G(N, p1, p1, 64, 0)
        self.g_r2, self.g_d2 = G(N, p1, p1, 192, 0), G(N, p1, p1, 192, 2)
        self.g_p2, self.g_p2r = G(N, p2, p2, 192, 1), G(N, p2, p2, 192, 0)
        self.g_y3, self.g_y3r = G(N, p2, p2, 384, 1), G(N, p2, p2, 384, 0)
        self.g_y4, self.g_y4r = G(N, p2, p2, 256, 1), G(N, p2, p2, 256, 0)
        self.g_r5 = G(N, p2, p2, 256, 0)
        self.g_p3 = G(N, p3, p3, 256, 0)
        R = dict(act=L.ACT_RELU, want_wgrad=False)
        self.c1 = ConvRT("IP.features.0", f[0].weight, f[0].bias, self.g_x0, 4, 2, OutMap.nhwc(self.g_r1), dyg=self.g_d1,
                         dx_out=OutMap.nhwc(G(N, S, S, 8, 0)), **R)
        self.c2 = ConvRT("IP.features.3", f[3].weight, f[3].bias, self.g_p1, 1, 2, OutMap.nhwc(self.g_r2), dyg=self.g_d2,
                         dx_out=OutMap.nhwc(self.g_p1r), **R)
        self.c3 = ConvRT("IP.features.6", f[6].weight, f[6].bias, self.g_p2, 1, 1, OutMap.nhwc(self.g_y3), dyg=self.g_y3,
                         dx_out=OutMap.nhwc(self.g_p2r), **R)
        self.c4 = ConvRT("IP.features.8", f[8].weight, f[8].bias, self.g_y3, 1, 1, OutMap.nhwc(self.g_y4), dyg=self.g_y4,
                         dx_out=OutMap.nhwc(self.g_y3r), **R)
        self.c5 = ConvRT("IP.features.10", f[10].weight, f[10].bias, self.g_y4, 1, 1, OutMap.nhwc(self.g_r5), dyg=self.g_y4,
                         dx_out=OutMap.nhwc(self.g_y4r), **R)
        self.scratch = _Scratch(dev)
        self.pool = Pool(lambda key: self._new_ws())
        self.bank = WeightBank([self.c1, self.c2, self.c3, self.c4, self.c5], dev)

    def _new_ws(self):
        ws, dev = _IPWorkspace(), self.dev
        z = lambda g: zeros_act(g, dev)
        ws.x0, ws.r1, ws.p1, ws.r2, ws.p2 = z(self.g_x0), z(self.g_r1), z(self.g_p1), z(self.g_r2), z(self.g_p2)
        ws.y3, ws.y4, ws.r5, ws.p3 = z(self.g_y3), z(self.g_y4), z(self.g_r5), z(self.g_p3)
        ws.i1 = torch.zeros(self.g_p1r.numel, dtype=torch.uint8, device=dev)
        ws.i2 = torch.zeros(self.g_p2r.numel, dtype=torch.uint8, device=dev)
        ws.i3 = torch.zeros(self.g_p3.numel, dtype=torch.uint8, device=dev)
        return ws

    def forward(self, x):
        ws = self.pool.take(0)
        self.bank.ensure_packed()
        ops.pack_nchw(x, ws.x0, self.g_x0, halo=L.HALO_ZERO)
        self.c1.forward(ws.x0, ws.r1)
        ops.maxpool_fwd(ws.r1, self.g_r1, ws.p1, self.g_p1.pad, ws.i1, pool_pad=0)
        self.c2.forward(ws.p1, ws.r2)
        ops.maxpool_fwd(ws.r2, self.g_r2, ws.p2, self.g_p2.pad, ws.i2, pool_pad=0)
        self.c3.forward(ws.p2, ws.y3)
        self.c4.forward(ws.y3, ws.y4)
        self.c5.forward(ws.y4, ws.r5)
        ops.maxpool_fwd(ws.r5, self.g_r5, ws.p3, 0, ws.i3, pool_pad=0)
        feat = torch.empty(self.N, self.p3, self.p3, 256, device=self.dev)
        ops.nhwc_to_f32(ws.p3, self.g_p3, feat)
        return feat.permute(0, 3, 1, 2), ws          # [N, 256, p3, p3] view of the NHWC feature map

    def backward(self, ws, dfeat):
        """dfeat: gradient of the [N, 256, p3, p3] feature map -> gradient of the input image [N, 3, S, S]."""
        N, sc = self.N, self.scratch
        self.bank.ensure_packed()
        g3 = sc.get(self.g_p3, "g3")
        ops.f32_to_nhwc(dfeat.permute(0, 2, 3, 1).contiguous().float(), g3, self.g_p3)
        g = sc.get(self.g_r5, "gr5")
        ops.maxpool_bwd(g3, 0, ws.i3, g, 0, N, self.p2, self.p2, 256, pool_pad=0)
        dy = sc.get(self.g_y4, "dy5")
        ops.act_bwd(g, 0, ws.r5, 0, dy, 1, self.g_r5)
        g = sc.get(self.g_y4r, "g4")
        self.c5.backward_data(dy, g)
        dy = sc.get(self.g_y4, "dy4")
        ops.act_bwd(g, 0, ws.y4, 1, dy, 1, self.g_y4r)
        g = sc.get(self.g_y3r, "g3r")
        self.c4.backward_data(dy, g)
        dy = sc.get(self.g_y3, "dy3")
        ops.act_bwd(g, 0, ws.y3, 1, dy, 1, self.g_y3r)
        g = sc.get(self.g_p2r, "gp2")
        self.c3.backward_data(dy, g)
        gr = sc.get(self.g_r2, "gr2")
        ops.maxpool_bwd(g, 0, ws.i2, gr, 0, N, self.p1, self.p1, 192, pool_pad=0)
        dy = sc.get(self.g_d2, "dy2")
        ops.act_bwd(gr, 0, ws.r2, 0, dy, 2, self.g_r2)
        g = sc.get(self.g_p1r, "gp1")
        self.c2.backward_data(dy, g)
        gr = sc.get(self.g_r1, "gr1")
        ops.maxpool_bwd(g, 0, ws.i1, gr, 0, N, self.h1, self.h1, 64, pool_pad=0)
        dy = sc.get(self.g_d1, "dy1")
        ops.act_bwd(gr, 0, ws.r1, 0, dy, 2, self.g_r1)
        gx = sc.get(Geom(N, self.S, self.S, 8, 0), "gx")
        self.c1.backward_data(dy, gx)
        dx = torch.empty(N, 3, self.S, self.S, device=self.dev)
        ops.unpack_resize_bwd(gx, Geom(N, self.S, self.S, 8, 0), dx)
        return dx


class AlexNetFeature(nn.Module):
    """Same constructor / forward / load_pretrained / state_dict keys (features.{0,3,6,8,10}.*) as
    models/networks.py:1218-1255 with pooling 'None' (what define_IP builds); frozen: no weight gradients."""

    def __init__(self, input_nc=3, pooling="None"):
        super().__init__()
        if input_nc != 3 or pooling not in ("None", None, ""):
            raise NotImplementedError("pcgan_b200 AlexNetFeature: RGB input, pooling 'None' (the wsgan_emb configuration)")
        self.pooling = pooling
        I = nn.Identity
        self.features = nn.Sequential(nn.Conv2d(3, 64, 11, stride=4, padding=2), I(), I(), nn.Conv2d(64, 192, 5, padding=2), I(), I(),
                                      nn.Conv2d(192, 384, 3, padding=1), I(), nn.Conv2d(384, 256, 3, padding=1), I(),
                                      nn.Conv2d(256, 256, 3, padding=1), I(), I())
        self.feature_dim = 256
        self._programs = {}
        self._key = CO.register_module(self)

    def _program(self, n, s):
        k = (n, s, self.features[0].weight.device)
        if k not in self._programs:
            self._programs[k] = _IPProgram(self, n, s)
        return self._programs[k]

    def forward(self, x):
        _require_cuda(x, "AlexNetFeature")
        if x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise NotImplementedError("square RGB inputs")
        out = torch.ops.pcgan.alexnet_feature(x, self._key)
        CO.finish_forward(self._key, out)
        return out

    def load_pretrained(self, state_dict):
        if isinstance(state_dict, str):
            state_dict = torch.load(state_dict)
        for key in list(state_dict.keys()):
            if key.startswith("classifier"):
                state_dict.pop(key)
        self.load_state_dict(state_dict, strict=True)


def define_IP(which_model_netIP, input_nc, gpu_ids=[]):
    """networks.define_IP (networks.py:179-193): no weight initialisation here, the weights are loaded."""
    if which_model_netIP != "alexnet":
        raise NotImplementedError("Identity-preserving model [%s] is outside the wsgan_emb hot path (alexnet)" % which_model_netIP)
    net = AlexNetFeature(input_nc=input_nc, pooling="None")
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.to(gpu_ids[0])
        net = LocalDataParallel(net, gpu_ids)
    return net


class Normalize(nn.Module):
    """networks.py:2421-2439 as called by WSGANEmbModel: `mean or std` is truthy, so it is the identity."""

    def __init__(self, mean=[], std=[]):
        super().__init__()
        self.identity_mapping = mean or std
        if not self.identity_mapping:
            raise NotImplementedError("Normalize with empty mean and std is unreachable in the reference")

    def __call__(self, input):
        return input


def upsample2d(inputTensor, targetSize):
    """util.upsample2d (util/util.py:111-117): bilinear, align_corners=True; identity when sizes match."""
    if targetSize <= 0 or inputTensor.size(2) == targetSize:
        return inputTensor
    _require_cuda(inputTensor, "upsample2d")
    return torch.ops.pcgan.upsample_bilinear_ac(inputTensor, targetSize)


NOISE_QUEUE = []   # test hook: explicit standard-normal draws for resample(), consumed in call order


def resample(mu=0.0, var=0.0):
    """util.resample (util/util.py:136-139): mu + randn * sqrt(var)."""
    std = torch.sqrt(var)
    eps = NOISE_QUEUE.pop(0).to(std.device, std.dtype).view_as(std) if NOISE_QUEUE else torch.randn_like(std)
    return mu + eps * std


def reparameterize(mu, logvar):
    """util.reparameterize (util/util.py:130-133)."""
    std = torch.exp(0.5 * logvar)
    return mu + torch.randn_like(std) * std


def compute_mu_and_var(E, x, T, noisy=False):
    """util.compute_mu_and_var (util/util.py:153-171): Monte-Carlo dropout statistics over T stochastic passes of the
    encoder: mean, E[y^2] - E[y]^2 and (noisy) the mean of exp(logvar)."""
    y_mu = 0.0
    y_sq = 0.0
    if not noisy:
        for _ in range(T):
            y = E(x)
            y_mu = y_mu + 1.0 / T * y
            y_sq = y_sq + 1.0 / T * y ** 2
        return y_mu, y_sq - y_mu ** 2
    s2_mu = 0.0
    for _ in range(T):
        y, logs2 = E(x)
        y_mu = y_mu + 1.0 / T * y
        y_sq = y_sq + 1.0 / T * y ** 2
        s2_mu = s2_mu + 1.0 / T * torch.exp(logs2)
    return y_mu, y_sq - y_mu ** 2, s2_mu
