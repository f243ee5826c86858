"""Teacher-forced, stage-by-stage checks of the three networks' forward and backward chains (the per-layer BF16 gate of
BASELINE.json: rel-L2 <= 2e-2 on activations and gradients), up to the BASELINE geometry (B = 64, 128 x 128).

Whole-network gradient comparisons are dominated by ReLU-mask flips (a relative forward perturbation d flips ~d of
the masks and costs ~sqrt(2d) in the gradient), so here every stage is fed the kernels' OWN stored tensors: the
convolution stages (linear) are re-computed by fp32 autograd from the stored bf16 inputs and the stored upstream
gradient, the normalisation / activation stages from the stored pre-norm tensor, with the kernels' own pre-activation
deciding the ReLU / LeakyReLU mask in the backward (the forward is compared with the unmodified reference op).  A wiring
mistake in a program's backward (a missing fold, residual, phase, transposed weight, shortcut, BatchNorm term) is an O(1)
error in exactly one stage.  Stages: ResnetGenerator (models/networks.py:565-652) with every block, NLayerDiscriminator
(:737-783) with all five convolutions and three BatchNorms, SiameseFeature / ResNet-18 (:1008-1083, models/resnet.py:31-73,
125-196) with the stem, the max pool, all eight BasicBlocks (shortcut convolutions included) and the head."""
import pytest
import torch
import torch.nn.functional as F

from oracle import pcgan_oracle as O
from pcgan_b200 import networks as NW

pytestmark = pytest.mark.gpu
DEV = "cuda"
EPS = 1e-5
# one bf16 rounding of a stage's output is 1.7e-3 rel-L2; stages that chain two stored roundings reach 2.5e-3.
# BASELINE gate for BF16: 2e-2.
TOL = 6e-3


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def bf(t):
    return t.to(torch.bfloat16).float()


def full(buf, g):
    return buf[: g.numel].view(g.n, g.hp, g.wp, g.c).float().permute(0, 3, 1, 2).contiguous()


def inner(buf, g):
    t = full(buf, g)
    return t[:, :, g.pad:g.pad + g.h, g.pad:g.pad + g.w].contiguous() if g.pad else t


class Log:
    def __init__(self):
        self.rows = []

    def check(self, name, got, want, tol=TOL):
        e = rel(got, want)
        self.rows.append((name, e))
        assert e < tol, "%s: rel-L2 %.3e >= %.1e" % (name, e, tol)

    def report(self, title):
        worst = sorted(self.rows, key=lambda t: -t[1])[:6]
        print("%s: %d stages, max %.3e; worst: %s" % (title, len(self.rows), worst[0][1], ", ".join("%s %.2e" % t for t in worst)))


def conv_stage(log, name, x, w, fn, gy, my_y, my_dx, my_dw, tol_y=TOL):
    """A linear stage: y = fn(x, w) by fp32 autograd on the stored bf16 input and the bf16-rounded weights."""
    x = x.clone().requires_grad_(True)
    w = bf(w.detach()).clone().requires_grad_(True)
    y = fn(x, w)
    if my_y is not None:
        log.check(name + ".fwd", my_y, y, tol_y)
    y.backward(gy)
    if my_dx is not None:
        log.check(name + ".dgrad", my_dx, x.grad)
    if my_dw is not None:
        log.check(name + ".wgrad", my_dw, w.grad)
    return x.grad


def norm_stage(log, name, r, gy, my_y, my_dr, ns, *, kind, act="relu", slope=0.0, gamma=None, beta=None, res=None, my_dres=None,
               my_dgamma=None, my_dbeta=None, res_mine=None, tol_bwd=TOL):
    """Instance / batch normalisation (+ residual) (+ ReLU / LeakyReLU), forward and backward, by fp32 autograd from the
    stored bf16 pre-norm tensor `r`.  The kernels normalise with the statistics of the fp32 accumulators, the reference
    here with those of the bf16-rounded tensor: that 1e-4 difference flips the mask of ~1e-3 of the elements (3e-2 in the
    gradient, the sqrt law), so in the backward the reference is given the kernels' own pre-activation for the mask and
    nothing else.  res: residual added before the activation (a leaf whose gradient is compared with my_dres); res_mine:
    the kernels' own value of that residual when they compute it themselves (the BatchNorm of the shortcut convolution)."""
    rr = r.clone().requires_grad_(True)
    if kind == "instance":
        pre = F.instance_norm(rr, eps=EPS)
        groups = r.size(0)
    else:
        gamma = gamma.detach().clone().requires_grad_(True)
        beta = beta.detach().clone().requires_grad_(True)
        pre = F.batch_norm(rr, None, None, gamma, beta, True, 0.0, EPS)
        groups = 1
    mine = ns.scale.view(groups, -1, 1, 1) * r + ns.shift.view(groups, -1, 1, 1)
    if res is not None:
        pre, mine = pre + res, mine + (res.detach() if res_mine is None else res_mine)
    if act == "none":
        y_ref = y = pre
    else:
        neg = slope if act == "lrelu" else 0.0
        y_ref = torch.where(pre > 0, pre, neg * pre)
        y = torch.where(mine > 0, pre, neg * pre)
    log.check(name + ".fwd", my_y, y_ref)
    y.backward(gy)
    want_dr = rr.grad
    if kind == "instance" and r.size(2) * r.size(3) <= 16 and res is None:
        # An instance norm over <= 16 pixels is ill-conditioned in the stored tensor: where a channel's mean is far from 0
        # relative to the spread of its handful of values, the statistics of the bf16-rounded tensor (autograd above) and of
        # the fp32 accumulators (kernels) give visibly different x_hat (up to 3e-2 in the gradient, data dependent).  Like
        # the masks, the statistics are therefore teacher-forced here: the closed-form backward with the kernels' own
        # mean / rstd,  dx = rstd (g - mean(g) - x_hat mean(g x_hat)).
        mk, rk = ns.mean.view(groups, -1, 1, 1), ns.rstd.view(groups, -1, 1, 1)
        xh = (r - mk) * rk
        neg = slope if act == "lrelu" else 0.0
        g = gy if act == "none" else torch.where(mine > 0, gy, neg * gy)
        want_dr = rk * (g - g.mean((2, 3), keepdim=True) - xh * (g * xh).mean((2, 3), keepdim=True))
    log.check(name + ".bwd", my_dr, want_dr, tol_bwd)
    if res is not None and my_dres is not None:
        log.check(name + ".dres", my_dres, res.grad)
    if my_dgamma is not None:
        log.check(name + ".dgamma", my_dgamma, gamma.grad)
        log.check(name + ".dbeta", my_dbeta, beta.grad)


# ------------------------------------------------------------------------------------------------ generator
@pytest.mark.parametrize("N,S,nb", [(2, 32, 1), (3, 64, 9), (64, 128, 9)], ids=["b2_s32_1block", "b3_s64_9blocks", "baseline_b64_s128"])
def test_generator_chain_teacher_forced(N, S, nb):
    sd = O.make_state_dict(O.generator_keys(n_blocks=nb), 41, device=DEV)
    net = NW.init_net(NW.ResnetGenerator(3, 3, 1, 64, norm_layer=NW.get_norm_layer("instance"), n_blocks=nb), "normal", [0])
    mod = net.module
    mod.load_state_dict({k: v.clone() for k, v in sd.items()})
    P = mod._program(N, S)
    P.keep_scratch = True
    a, _, _ = O.synthetic_batch(N, S, 300, device=DEV)
    z = torch.linspace(-1, 1, N, device=DEV).view(N, 1, 1, 1)
    out, ws = P.forward(a.contiguous(), z.view(-1).contiguous())
    dout = torch.randn_like(out)
    dx, _ = P.backward(ws, out, dout, True, True)
    torch.cuda.synchronize()
    sc, log = P.scratch, Log()
    st = mod.state_dict(keep_vars=True)
    Wt = lambda k: st[k]
    grad = lambda k: st[k].grad
    B = lambda k: st[k].detach()
    h4 = S // 4
    b = 10 + nb

    def fold(gfull, c, s, p):
        """adjoint of nn.ReflectionPad2d(p) applied to a padded-grid gradient"""
        xi = torch.zeros(N, c, s, s, device=DEV, requires_grad=True)
        F.pad(xi, (p,) * 4, mode="reflect").backward(gfull)
        return xi.grad.detach()

    PRE = NW.FOLD_PREPASS    # the reflect fold has been applied in place: interior of the padded gradient = folded, halo = raw

    def check_padded_grad(name, mine_full, ref_full, c, s, p):
        """data gradient on the padded grid; returns the folded gradient the next stage receives"""
        if not PRE:
            log.check(name + ".dgrad", mine_full, ref_full)
            return fold(mine_full, c, s, p)
        halo = torch.ones_like(ref_full, dtype=torch.bool)
        halo[:, :, p:p + s, p:p + s] = False
        log.check(name + ".dgrad (halo)", mine_full[halo], ref_full[halo])
        mine = mine_full[:, :, p:p + s, p:p + s].contiguous()
        log.check(name + ".dgrad + fold", mine, fold(ref_full, c, s, p))
        return mine

    # ---- head: conv7x7 over the reflect-padded buffer + tanh
    kh = "model.%d" % (b + 7)
    dyh = inner(sc.get(P.g_dyh), P.g_dyh)[:, :3]
    log.check("head.dtanh", dyh, dout * (1 - out * out))
    u2p = full(ws.u2, P.g_u2)
    ref_full = conv_stage(log, "head", u2p, Wt(kh + ".weight"), lambda x, w: F.conv2d(x, w), dyh, None, None, grad(kh + ".weight"))
    with torch.no_grad():
        log.check("head.fwd", out, torch.tanh(F.conv2d(u2p, bf(Wt(kh + ".weight").detach()), B(kh + ".bias"))), 1e-4)
    log.check("head.bias_grad", grad(kh + ".bias"), dyh.sum((0, 2, 3)), 1e-3)
    # the fold is a small in-place pre-pass (or, PCGAN_FOLD_PREPASS=0, fused into the up2 norm backward: dy_fold=2)
    g_u2 = check_padded_grad("head", full(sc.get(P.g_u2full), P.g_u2full), ref_full, 64, S, 3)
    # ---- up2, up1: ConvTranspose2d + IN + ReLU
    norm_stage(log, "up2.norm", inner(ws.u2r, P.g_u2r), g_u2, inner(ws.u2, P.g_u2), inner(sc.get(P.g_a1, "dy"), P.g_a1), ws.nu2, kind="instance")
    k = "model.%d" % (b + 3)
    g_u1 = inner(sc.get(P.g_u1r, "g_u1"), P.g_u1r)
    conv_stage(log, "up2.conv", inner(ws.u1, P.g_u1), Wt(k + ".weight"),
               lambda x, w: F.conv_transpose2d(x, w, B(k + ".bias"), stride=2, padding=1, output_padding=1),
               inner(sc.get(P.g_a1, "dy"), P.g_a1), inner(ws.u2r, P.g_u2r), g_u1, grad(k + ".weight"))
    norm_stage(log, "up1.norm", inner(ws.u1r, P.g_u1r), g_u1, inner(ws.u1, P.g_u1), inner(sc.get(P.g_u1, "dy"), P.g_u1), ws.nu1, kind="instance")
    k = "model.%d" % b
    gb = inner(sc.get(P.g_r3, "gb0"), P.g_r3)
    conv_stage(log, "up1.conv", inner(ws.b[nb], P.g_b), Wt(k + ".weight"),
               lambda x, w: F.conv_transpose2d(x, w, B(k + ".bias"), stride=2, padding=1, output_padding=1),
               inner(sc.get(P.g_u1, "dy"), P.g_u1), inner(ws.u1r, P.g_u1r), gb, grad(k + ".weight"))
    # ---- the residual blocks, last to first: x + IN(conv(relu(IN(conv(x)))))  (reflect padding in the buffers)
    for i in range(nb - 1, -1, -1):
        p, t = "model.%d.conv_block" % (10 + i), "blk%d" % i
        xres = inner(ws.b[i], P.g_b).clone().requires_grad_(True)
        dyb = inner(sc.get(P.g_b, "dyb%d" % i), P.g_b)
        norm_stage(log, t + ".norm2+res", inner(ws.rb[i], P.g_r3), gb, inner(ws.b[i + 1], P.g_b), dyb, ws.nb[i], kind="instance",
                   act="none", res=xres)
        dfull = full(sc.get(P.g_bfull, "dfull%d" % i), P.g_bfull)
        ref_full = conv_stage(log, t + ".conv2", full(ws.h[i], P.g_b), Wt(p + ".5.weight"), lambda x, w: F.conv2d(x, w, B(p + ".5.bias")),
                              dyb, inner(ws.rb[i], P.g_r3), None, grad(p + ".5.weight"))
        g_ra = check_padded_grad(t + ".conv2", dfull, ref_full, 256, h4, 1)
        dya = inner(sc.get(P.g_b, "dya%d" % i), P.g_b)
        norm_stage(log, t + ".norm1", inner(ws.ra[i], P.g_r3), g_ra, inner(ws.h[i], P.g_b), dya, ws.na[i], kind="instance")
        dfull2 = full(sc.get(P.g_bfull, "dfull2%d" % i), P.g_bfull)
        ref_full = conv_stage(log, t + ".conv1", full(ws.b[i], P.g_b), Wt(p + ".1.weight"), lambda x, w: F.conv2d(x, w, B(p + ".1.bias")),
                              dya, inner(ws.ra[i], P.g_r3), None, grad(p + ".1.weight"))
        g_b = check_padded_grad(t + ".conv1", dfull2, ref_full, 256, h4, 1)
        gprev = inner(sc.get(P.g_r3, "gbk%d" % i), P.g_r3)
        log.check(t + ".fold+skip", gprev, g_b + gb)
        gb = gprev
    # ---- down2, down1 (stride 2, zero padding), stem
    dy3 = inner(sc.get(P.g_r3, "dy3"), P.g_r3)
    norm_stage(log, "down2.norm", inner(ws.r3, P.g_r3), gb, inner(ws.b[0], P.g_b), dy3, ws.n3, kind="instance")
    g2 = inner(sc.get(P.g_r2, "g"), P.g_r2)
    conv_stage(log, "down2.conv", inner(ws.a2, P.g_a2), Wt("model.7.weight"), lambda x, w: F.conv2d(x, w, B("model.7.bias"), stride=2, padding=1),
               dy3, inner(ws.r3, P.g_r3), g2, grad("model.7.weight"))
    dy2 = inner(sc.get(P.g_r2, "dy2"), P.g_r2)
    norm_stage(log, "down1.norm", inner(ws.r2, P.g_r2), g2, inner(ws.a2, P.g_a2), dy2, ws.n2, kind="instance")
    g1 = inner(sc.get(P.g_r1, "g"), P.g_r1)
    conv_stage(log, "down1.conv", inner(ws.a1, P.g_a1), Wt("model.4.weight"), lambda x, w: F.conv2d(x, w, B("model.4.bias"), stride=2, padding=1),
               dy2, inner(ws.r2, P.g_r2), g1, grad("model.4.weight"))
    gdy1 = NW.Geom(N, S, S, 64, 3)
    dy1 = inner(sc.get(gdy1, "dy1"), gdy1)
    norm_stage(log, "stem.norm", inner(ws.r1, P.g_r1), g1, inner(ws.a1, P.g_a1), dy1, ws.n1, kind="instance")
    x0 = full(ws.x0, P.g_x0)[:, :4]                 # (r, g, b, z) reflect-padded
    gx0 = conv_stage(log, "stem.conv", x0, Wt("model.1.weight"), lambda x, w: F.conv2d(x, w, B("model.1.bias")), dy1,
                     inner(ws.r1, P.g_r1), None, grad("model.1.weight"))
    log.check("stem.dgrad+fold -> dx", dx, fold(bf(gx0), 4, S, 3)[:, :3])
    xz = torch.cat([a, z.expand(N, 1, S, S)], 1)
    log.check("input pack", x0, bf(F.pad(xz, (3,) * 4, mode="reflect")), 1e-6)
    log.report("G chain N=%d S=%d blocks=%d" % (N, S, nb))
    assert len(log.rows) >= 25 + 11 * nb


# -------------------------------------------------------------------------------------------- discriminator
@pytest.mark.parametrize("N,S", [(3, 64), (64, 128)], ids=["b3_s64", "baseline_b64_s128"])
def test_discriminator_chain_teacher_forced(N, S):
    sd = O.make_state_dict(O.discriminator_keys(), 42, device=DEV)
    net = NW.define_D(3, 1, 64, "n_layers", 3, "batch", True, "normal", gpu_ids=[0])
    mod = net.module
    mod.load_state_dict({k: v.clone() for k, v in sd.items()})
    P = mod._program(N, S)
    a, _, _ = O.synthetic_batch(N, S, 301, device=DEV)
    z = torch.linspace(-1, 1, N, device=DEV)
    out, ws = P.forward(a.contiguous(), z.contiguous())
    dout = torch.randn_like(out) / out.numel()
    dx, _ = P.backward(ws, out, dout, True, True)
    torch.cuda.synchronize()
    sc, log, m = P.scratch, Log(), mod.model
    nl = len(P.sizes)            # 4 convolutions before the head
    # ---- head: conv4x4 s1 p1 -> 1 channel, bias, sigmoid
    hd = m[P.i_head]
    dyh = inner(sc.get(P.g_dyh), P.g_dyh)[:, :1]
    log.check("head.dsigmoid", dyh, dout * out * (1 - out))
    g = inner(sc.get(P.g_r[-1], "g%d" % (nl - 1)), P.g_r[-1])
    ylast = inner(ws.y[-1], P.g_y[-1])
    conv_stage(log, "head", ylast, hd.weight, lambda x, w: F.conv2d(x, w, padding=1), dyh, None, g, hd.weight.grad)
    with torch.no_grad():
        log.check("head.fwd", out, torch.sigmoid(F.conv2d(ylast, bf(hd.weight.detach()), hd.bias.detach(), padding=1)), 1e-4)
    log.check("head.bias_grad", hd.bias.grad, dyh.sum((0, 2, 3)), 1e-3)
    # ---- conv (no bias) + BatchNorm + LeakyReLU(0.2) layers, last to first
    for li in range(nl - 1, 0, -1):
        conv, bn = m[P.idx[li]], m[P.idx[li] + 1]
        stride = 2 if li < nl - 1 else 1
        dy = inner(sc.get(P.g_r[li], "dy%d" % li), P.g_r[li])
        norm_stage(log, "L%d.bn" % li, inner(ws.r[li], P.g_r[li]), g, inner(ws.y[li], P.g_y[li]), dy, ws.ns[li], kind="batch", act="lrelu",
                   slope=0.2, gamma=bn.weight, beta=bn.bias, my_dgamma=bn.weight.grad, my_dbeta=bn.bias.grad)
        gprev = inner(sc.get(P.g_r[li - 1], "g%d" % (li - 1)), P.g_r[li - 1])
        conv_stage(log, "L%d.conv" % li, inner(ws.y[li - 1], P.g_y[li - 1]), conv.weight, lambda x, w: F.conv2d(x, w, stride=stride, padding=1),
                   dy, inner(ws.r[li], P.g_r[li]), gprev, conv.weight.grad)
        g = gprev
    # ---- layer 0: conv4x4 s2 p1 + bias + LeakyReLU in the epilogue; the backward masks by the sign of the stored output
    c0 = m[0]
    x0 = inner(ws.x0, P.g_x0)[:, :4].clone().requires_grad_(True)
    w0 = bf(c0.weight.detach()).clone().requires_grad_(True)
    b0 = c0.bias.detach().clone().requires_grad_(True)
    pre = F.conv2d(x0, w0, b0, stride=2, padding=1)
    my_y0 = inner(ws.y[0], P.g_y[0])
    log.check("L0.fwd", my_y0, F.leaky_relu(pre, 0.2))
    torch.where(my_y0 > 0, pre, 0.2 * pre).backward(g)
    dy0 = inner(sc.get(P.g_r[0], "dy0"), P.g_r[0])
    log.check("L0.dlrelu", dy0, torch.where(my_y0 > 0, g, 0.2 * g))
    log.check("L0.wgrad", c0.weight.grad, w0.grad)
    log.check("L0.bias_grad", c0.bias.grad, b0.grad)
    log.check("L0.dgrad -> dx", dx, x0.grad[:, :3])
    xz = torch.cat([a, z.view(N, 1, 1, 1).expand(N, 1, S, S)], 1)
    log.check("input pack", full(ws.x0, P.g_x0)[:, :4], bf(F.pad(xz, (1,) * 4)), 1e-6)
    log.report("D chain N=%d S=%d" % (N, S))
    assert len(log.rows) >= 30


# -------------------------------------------------------------------------------------------------- encoder
def _enc_block_chain(log, blk, w, xbuf, gy, sc):
    """One BasicBlock (resnet.py:55-73), stage by stage; gy = gradient of the block output.  Returns the kernels' gradient
    of the block input."""
    hd, t = blk.holder, blk.name
    s = blk.stride
    x = inner(xbuf, blk.g_x)
    get = lambda g, tag: inner(sc.get(g, blk.name + tag), g)
    dyb = get(blk.g_y, "dyb")
    gres = get(blk.g_r, "gres")
    if blk.ds is not None:
        # shortcut: 1x1 stride-s convolution + BatchNorm, added before the final ReLU
        bnd = hd.downsample[1]
        rd = inner(w.rd, blk.g_r).clone().requires_grad_(True)
        gd, bd = bnd.weight.detach().clone().requires_grad_(True), bnd.bias.detach().clone().requires_grad_(True)
        res = F.batch_norm(rd, None, None, gd, bd, True, 0.0, EPS)
        res.retain_grad()
        res_mine = w.nd.scale.view(1, -1, 1, 1) * rd.detach() + w.nd.shift.view(1, -1, 1, 1)
    else:
        res, res_mine = x.clone().requires_grad_(True), None
    norm_stage(log, t + ".bn2+res", inner(w.rb, blk.g_r), gy, inner(w.y, blk.g_y), dyb, w.nb, kind="batch", gamma=hd.bn2.weight, beta=hd.bn2.bias,
               res=res, my_dres=gres, my_dgamma=hd.bn2.weight.grad, my_dbeta=hd.bn2.bias.grad, res_mine=res_mine)
    gh = get(blk.g_r, "gh")
    conv_stage(log, t + ".conv2", inner(w.h, blk.g_y), hd.conv2.weight, lambda a, b: F.conv2d(a, b, padding=1), dyb, inner(w.rb, blk.g_r), gh,
               hd.conv2.weight.grad)
    dya = get(blk.g_y if s == 1 else blk.g_r, "dya")
    norm_stage(log, t + ".bn1", inner(w.ra, blk.g_r), gh, inner(w.h, blk.g_y), dya, w.na, kind="batch", gamma=hd.bn1.weight, beta=hd.bn1.bias,
               my_dgamma=hd.bn1.weight.grad, my_dbeta=hd.bn1.bias.grad)
    gx1 = get(blk.g_xr, "gx1")
    conv_stage(log, t + ".conv1", x, hd.conv1.weight, lambda a, b: F.conv2d(a, b, stride=s, padding=1), dya, inner(w.ra, blk.g_r), gx1,
               hd.conv1.weight.grad)
    gx = get(blk.g_xr, "gx")
    if blk.ds is not None:
        # the shortcut BatchNorm's backward from the kernels' own (bf16-stored) shortcut gradient
        dyd = get(blk.g_r, "dyd")
        rd2 = inner(w.rd, blk.g_r).clone().requires_grad_(True)
        gd2, bd2 = bnd.weight.detach().clone().requires_grad_(True), bnd.bias.detach().clone().requires_grad_(True)
        F.batch_norm(rd2, None, None, gd2, bd2, True, 0.0, EPS).backward(gres)
        log.check(t + ".bn_ds.bwd", dyd, rd2.grad)
        log.check(t + ".bn_ds.dgamma", bnd.weight.grad, gd2.grad)
        log.check(t + ".bn_ds.dbeta", bnd.bias.grad, bd2.grad)
        gx2 = get(blk.g_xr, "gx2")
        conv_stage(log, t + ".ds", x, hd.downsample[0].weight, lambda a, b: F.conv2d(a, b, stride=s), dyd, inner(w.rd, blk.g_r), gx2,
                   hd.downsample[0].weight.grad)
        log.check(t + ".sum", gx, gx1 + gx2)
    else:
        log.check(t + ".sum", gx, gx1 + gres)
    return gx


@pytest.mark.parametrize("N,S", [(8, 64), (16, 224)], ids=["b8_s64", "b16_s224"])
def test_encoder_chain_teacher_forced(N, S):
    sd = O.make_state_dict(O.encoder_keys(), 44, device=DEV)
    net = NW.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, gpu_ids=[0])
    mod = net.module
    mod.load_state_dict({k: v.clone() for k, v in sd.items()})
    P = mod._program(N, S)
    a, _, _ = O.synthetic_batch(N, S, 303, device=DEV)
    outs, ws = P.forward(a.contiguous())
    gy = torch.linspace(-2, 1, N, device=DEV).view(N, 1, 1, 1)
    dx = P.backward(ws, [gy], True, True)
    torch.cuda.synchronize()
    sc, log, rn = P.scratch, Log(), mod.base.model
    hf = P.hf
    # ---- head: conv3x3 512->32 + bias, BN, LeakyReLU(0.7), conv3x3 32->1 + bias, global average pool
    h, hw = P.heads[0], ws.heads[0]
    seq = mod.cnn
    feats = inner(ws.blk[-1].y, P.blocks[-1].g_y)
    hh = inner(hw.hh, h.g_hh)
    with torch.no_grad():
        yref = F.conv2d(hh, bf(seq[4].weight.detach()), seq[4].bias.detach(), padding=1).mean((2, 3), keepdim=True)
    log.check("head.conv2+pool.fwd", outs[0], yref, 1e-4)
    dyf = inner(sc.get(h.g_dyfin, h.name + "dyf"), h.g_dyfin)[:, :1]
    log.check("head.dpool", dyf, (gy / float(hf * hf)).expand(N, 1, hf, hf), 4e-3)
    ghh = inner(sc.get(h.g_rh, h.name + "ghh"), h.g_rh)
    conv_stage(log, "head.conv2", hh, seq[4].weight, lambda x, w: F.conv2d(x, w, padding=1), dyf, None, ghh, seq[4].weight.grad)
    log.check("head.conv2.bias_grad", seq[4].bias.grad, gy.sum().reshape(1), 1e-4)
    dyh = inner(sc.get(h.g_hh, h.name + "dyh"), h.g_hh)
    norm_stage(log, "head.bn", inner(hw.rh, h.g_rh), ghh, hh, dyh, hw.nh, kind="batch", act="lrelu", slope=0.7, gamma=seq[1].weight, beta=seq[1].bias,
               my_dgamma=seq[1].weight.grad, my_dbeta=seq[1].bias.grad)
    g = inner(sc.get(P.g_ff, "gf0"), P.g_ff)
    conv_stage(log, "head.conv1", feats, seq[0].weight, lambda x, w: F.conv2d(x, w, seq[0].bias.detach(), padding=1), dyh, inner(hw.rh, h.g_rh), g,
               seq[0].weight.grad)
    # ---- the eight BasicBlocks, last to first
    inputs = [ws.p] + [w.y for w in ws.blk[:-1]]
    for blk, w, xin in zip(reversed(P.blocks), reversed(ws.blk), reversed(inputs)):
        g = _enc_block_chain(log, blk, w, xin, g, sc)
    # ---- max pool 3x3 s2 p1, BatchNorm + ReLU, stem conv 7x7 s2 p3
    a0 = inner(ws.a0, P.g_a0).clone().requires_grad_(True)
    pool = F.max_pool2d(a0, 3, 2, 1)
    log.check("maxpool.fwd", inner(ws.p, P.g_p), pool, 1e-6)
    pool.backward(g)
    ga0 = inner(sc.get(P.g_a0, "ga0"), P.g_a0)
    log.check("maxpool.bwd", ga0, a0.grad)
    dy0 = inner(sc.get(P.g_dy0, "dy0"), P.g_dy0)
    norm_stage(log, "stem.bn", inner(ws.r0, P.g_r0), ga0, inner(ws.a0, P.g_a0), dy0, ws.n0, kind="batch", gamma=rn.bn1.weight, beta=rn.bn1.bias,
               my_dgamma=rn.bn1.weight.grad, my_dbeta=rn.bn1.bias.grad)
    x0 = inner(ws.x0, P.g_x0)[:, :3]
    gx = conv_stage(log, "stem.conv", x0, rn.conv1.weight, lambda x, w: F.conv2d(x, w, stride=2, padding=3), dy0, inner(ws.r0, P.g_r0), None,
                    rn.conv1.weight.grad)
    log.check("stem.dgrad -> dx", dx, gx)
    log.check("input pack", x0, bf(a), 1e-6)
    log.report("E chain N=%d S=%d" % (N, S))
    assert len(log.rows) >= 90


# ---------------------------------------------------------------------------------- identity-preserving net
@pytest.mark.parametrize("N,S", [(3, 64), (16, 224)], ids=["b3_s64", "b16_s224"])
def test_alexnet_chain_teacher_forced(N, S):
    """AlexNetFeature (models/networks.py:1218-1255; frozen: forward and input gradient), stage by stage: conv11x11 s4 p2,
    the three un-padded 3x3 s2 max pools, conv5x5 p2, the 3x3 convolutions, every fused ReLU and its backward."""
    sd = O.make_state_dict(O.alexnet_keys(), 64, device=DEV)
    net = NW.define_IP("alexnet", 3, [0])
    mod = net.module
    mod.load_state_dict({k: v.clone() for k, v in sd.items()})
    P = mod._program(N, S)
    a, _, _ = O.synthetic_batch(N, S, 305, device=DEV)
    feat, ws = P.forward(a.contiguous())
    dfeat = torch.randn(N, 256, P.p3, P.p3, device=DEV)
    dx = P.backward(ws, dfeat)
    torch.cuda.synchronize()
    sc, log, f = P.scratch, Log(), mod.features
    with torch.no_grad():
        log.check("features vs oracle (end to end)", feat, O.alexnet_forward(sd, a), 3e-2)

    def conv_relu_stage(name, x, conv, stride, pad, my_y, dy_masked, my_dx):
        """y = relu(conv(x) + b): forward from the stored input; the data gradient from the kernels' own masked dy"""
        xx = x.clone().requires_grad_(True)
        pre = F.conv2d(xx, bf(conv.weight.detach()), conv.bias.detach(), stride=stride, padding=pad)
        log.check(name + ".fwd", my_y, torch.relu(pre))
        pre.backward(dy_masked)
        log.check(name + ".dgrad", my_dx, xx.grad)

    def pool_stage(name, x, gy, my_y, my_dx):
        xx = x.clone().requires_grad_(True)
        y = F.max_pool2d(xx, 3, 2)
        log.check(name + ".fwd", my_y, y, 1e-6)
        y.backward(gy)
        log.check(name + ".bwd", my_dx, xx.grad)       # overlapping windows: up to four gradients summed, then one bf16 rounding

    def relu_bwd(name, g, y, my_dy):
        log.check(name + ".drelu", my_dy, torch.where(y > 0, g, torch.zeros_like(g)), 1e-6)

    S0 = NW.Geom(N, S, S, 8, 0)
    g3 = inner(sc.get(P.g_p3, "g3"), P.g_p3)
    log.check("dfeat cast", g3, bf(dfeat))
    r5, gr5 = inner(ws.r5, P.g_r5), inner(sc.get(P.g_r5, "gr5"), P.g_r5)
    pool_stage("pool3", r5, g3, inner(ws.p3, P.g_p3), gr5)
    dy5 = inner(sc.get(P.g_y4, "dy5"), P.g_y4)
    relu_bwd("conv5", gr5, r5, dy5)
    y4, g4 = inner(ws.y4, P.g_y4), inner(sc.get(P.g_y4r, "g4"), P.g_y4r)
    conv_relu_stage("conv5", y4, f[10], 1, 1, r5, dy5, g4)
    dy4 = inner(sc.get(P.g_y4, "dy4"), P.g_y4)
    relu_bwd("conv4", g4, y4, dy4)
    y3, g3r = inner(ws.y3, P.g_y3), inner(sc.get(P.g_y3r, "g3r"), P.g_y3r)
    conv_relu_stage("conv4", y3, f[8], 1, 1, y4, dy4, g3r)
    dy3 = inner(sc.get(P.g_y3, "dy3"), P.g_y3)
    relu_bwd("conv3", g3r, y3, dy3)
    p2, gp2 = inner(ws.p2, P.g_p2), inner(sc.get(P.g_p2r, "gp2"), P.g_p2r)
    conv_relu_stage("conv3", p2, f[6], 1, 1, y3, dy3, gp2)
    r2, gr2 = inner(ws.r2, P.g_r2), inner(sc.get(P.g_r2, "gr2"), P.g_r2)
    pool_stage("pool2", r2, gp2, p2, gr2)
    dy2 = inner(sc.get(P.g_d2, "dy2"), P.g_d2)
    relu_bwd("conv2", gr2, r2, dy2)
    p1, gp1 = inner(ws.p1, P.g_p1), inner(sc.get(P.g_p1r, "gp1"), P.g_p1r)
    conv_relu_stage("conv2", p1, f[3], 1, 2, r2, dy2, gp1)
    r1, gr1 = inner(ws.r1, P.g_r1), inner(sc.get(P.g_r1, "gr1"), P.g_r1)
    pool_stage("pool1", r1, gp1, p1, gr1)
    dy1 = inner(sc.get(P.g_d1, "dy1"), P.g_d1)
    relu_bwd("conv1", gr1, r1, dy1)
    x0 = inner(ws.x0, P.g_x0)[:, :3]
    conv_relu_stage("conv1", x0, f[0], 4, 2, r1, dy1, inner(sc.get(S0, "gx"), S0)[:, :3])
    log.check("conv1.dgrad -> dx", dx, inner(sc.get(S0, "gx"), S0)[:, :3], 1e-6)
    log.check("input pack", x0, bf(a), 1e-6)
    log.report("IP chain N=%d S=%d" % (N, S))
    assert len(log.rows) >= 24


# ------------------------------------------------------------------------------------------ unet generator
@pytest.mark.parametrize("N,S,D", [(2, 128, 7), (16, 128, 7), (2, 64, 5)], ids=["b2_unet128", "b16_unet128", "b2_s64_5downs"])
def test_unet_chain_teacher_forced(N, S, D):
    """UnetGenerator (models/networks.py:659-733; the default G of wsgan_emb), every stage: down convolutions with their
    InstanceNorm + in-place LeakyReLU, the skip halves (ReLU of the LeakyReLU'd tensor) and up halves of the concatenation
    buffers, up convolutions with InstanceNorm + ReLU, tanh; backward: the two halves of every up convolution's data
    gradient, the skip + down-path sum at every level, all weight and bias gradients."""
    sd = O.make_state_dict(O.unet_keys(num_downs=D), 71, device=DEV)
    net = NW.init_net(NW.UnetGenerator(3, 3, 1, D, 64, norm_layer=NW.get_norm_layer("instance")), "normal", [0])
    mod = net.module
    mod.load_state_dict({k: v.clone() for k, v in sd.items()})
    P = mod._program(N, S)
    a, _, _ = O.synthetic_batch(N, S, 306, device=DEV)
    z = torch.linspace(-1, 1, N, device=DEV).view(N, 1, 1, 1)
    out, ws = P.forward(a.contiguous(), z.view(-1).contiguous())
    dout = torch.randn(out.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(307))
    dx, _ = P.backward(ws, out, dout, True, True)
    torch.cuda.synchronize()
    sc, log, C, sz = P.scratch, Log(), P.C, P.sz
    dn, up = mod.down, mod.up
    get = lambda g, tag: inner(sc.get(g, tag), g)
    with torch.no_grad():
        log.check("output vs oracle (end to end)", out, O.unet_forward(sd, a, z, D), 5e-2)
    # ---- outermost up convolution + tanh
    g_dyh = NW.Geom(N, S, S, 8, 1)
    dyh = get(g_dyh, "dyh")[:, :3]
    log.check("up0.dtanh", dyh, dout * (1 - out * out))
    B0 = inner(ws.B[0], P.g_B[0])
    gB0 = conv_stage(log, "up0", B0, up[0].weight, lambda x, w: F.conv_transpose2d(x, w, stride=2, padding=1), dyh, None, None, up[0].weight.grad)
    with torch.no_grad():
        log.check("up0.fwd", out, torch.tanh(F.conv_transpose2d(B0, bf(up[0].weight.detach()), up[0].bias.detach(), stride=2, padding=1)), 1e-4)
    log.check("up0.bias_grad", up[0].bias.grad, dyh.sum((0, 2, 3)), 1e-3)
    log.check("up0.dgrad.skip", get(P.g_r[0], "gs0"), gB0[:, :C[0]])
    log.check("up0.dgrad.up", get(P.g_r[0], "gu0"), gB0[:, C[0]:])
    # ---- up path
    for k in range(1, D):
        gu = get(P.g_r[k - 1], "gu%d" % (k - 1))
        dyu = get(P.g_A[k - 1], "dyu%d" % k)
        norm_stage(log, "up%d.norm" % k, inner(ws.u[k], P.g_r[k - 1]), gu, inner(ws.B[k - 1], P.g_B[k - 1])[:, C[k - 1]:], dyu, ws.nu[k], kind="instance")
        src = inner(ws.A[k], P.g_A[k]) if k == D - 1 else inner(ws.B[k], P.g_B[k])
        gsrc = conv_stage(log, "up%d.conv" % k, src, up[k].weight, lambda x, w: F.conv_transpose2d(x, w, up[k].bias.detach(), stride=2, padding=1),
                          dyu, inner(ws.u[k], P.g_r[k - 1]), None, up[k].weight.grad)
        if k == D - 1:
            log.check("up%d.dgrad" % k, get(P.g_r[k], "gR"), gsrc)
        else:
            log.check("up%d.dgrad.skip" % k, get(P.g_r[k], "gs%d" % k), gsrc[:, :C[k]])
            log.check("up%d.dgrad.up" % k, get(P.g_r[k], "gu%d" % k), gsrc[:, C[k]:])
    # ---- innermost down convolution (ReLU in the epilogue, no norm)
    k = D - 1
    R, gR, dyk = inner(ws.A[k], P.g_A[k]), get(P.g_r[k], "gR"), get(P.g_r[k], "dy%d" % k)
    log.check("down%d.drelu" % k, dyk, torch.where(R > 0, gR, torch.zeros_like(gR)), 1e-6)
    xin = inner(ws.A[k - 1], P.g_A[k - 1])
    xx = xin.clone().requires_grad_(True)
    wq = bf(dn[k].weight.detach()).clone().requires_grad_(True)
    bq = dn[k].bias.detach().clone().requires_grad_(True)
    pre = F.conv2d(xx, wq, bq, stride=2, padding=1)
    log.check("down%d.fwd" % k, R, torch.relu(pre))
    pre.backward(dyk)
    log.check("down%d.wgrad" % k, dn[k].weight.grad, wq.grad)
    log.check("down%d.bias_grad" % k, dn[k].bias.grad, bq.grad, 1e-3)
    log.check("down%d.dgrad" % k, get(P.g_r[k - 1], "dA%d" % (k - 1)), xx.grad)
    # ---- down path, inner to outer
    for k in range(D - 2, -1, -1):
        A_k = inner(ws.A[k], P.g_A[k])
        skip, dA = get(P.g_r[k], "gs%d" % k), get(P.g_r[k], "dA%d" % k)
        eff = get(P.g_r[k], "eff%d" % k)
        log.check("down%d.skip+path" % k, eff, dA + torch.where(A_k > 0, skip, torch.zeros_like(skip)))
        log.check("down%d.skip half" % k, inner(ws.B[k], P.g_B[k])[:, :C[k]], torch.relu(A_k), 1e-6)
        dyk = get(P.g_r[k], "dy%d" % k)
        if k > 0:
            norm_stage(log, "down%d.norm" % k, inner(ws.r[k], P.g_r[k]), eff, A_k, dyk, ws.nd[k], kind="instance", act="lrelu", slope=0.2)
            conv_stage(log, "down%d.conv" % k, inner(ws.A[k - 1], P.g_A[k - 1]), dn[k].weight,
                       lambda x, w: F.conv2d(x, w, dn[k].bias.detach(), stride=2, padding=1), dyk, inner(ws.r[k], P.g_r[k]),
                       get(P.g_r[k - 1], "dA%d" % (k - 1)), dn[k].weight.grad)
        else:
            log.check("down0.dlrelu", dyk, torch.where(A_k > 0, eff, 0.2 * eff))
            x0 = inner(ws.x0, P.g_x0)[:, :4].clone().requires_grad_(True)
            wq = bf(dn[0].weight.detach()).clone().requires_grad_(True)
            bq = dn[0].bias.detach().clone().requires_grad_(True)
            pre = F.conv2d(x0, wq, bq, stride=2, padding=1)
            log.check("down0.fwd", A_k, F.leaky_relu(pre, 0.2))
            pre.backward(dyk)
            log.check("down0.wgrad", dn[0].weight.grad, wq.grad)
            log.check("down0.bias_grad", dn[0].bias.grad, bq.grad, 1e-3)
            log.check("down0.dgrad -> dx", dx, x0.grad[:, :3])
    xz = torch.cat([a, z.expand(N, 1, S, S)], 1)
    log.check("input pack", inner(ws.x0, P.g_x0)[:, :4], bf(xz), 1e-6)
    log.report("Unet chain N=%d S=%d downs=%d" % (N, S, D))
    assert len(log.rows) >= 10 * D
