"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's wsgan_emb hot path.

Plain PyTorch (fp32, runs on CPU or, for speed on the GPU box, on CUDA with TF32 off),
written from the reference's behaviour as functions over reference-named state_dicts;
nothing here is imported by the product (pcgan_b200/), only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

Pinned against the reference itself: tests/golden/make_golden.py imports
/root/reference (phymhan/pc-gan) in the authoring container, runs its modules and its
WSGANEmbModel.optimize_parameters on seeded inputs, and commits the results under
tests/golden/; tests/test_oracle_cpu.py replays them through this file.

Every function cites the reference lines it restates (paths relative to the reference).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

EPS = 1e-5        # nn.InstanceNorm2d / nn.BatchNorm2d default eps
MOMENTUM = 0.1    # default momentum of both
MAGIC_EPS = 1e-20  # models/networks.py:14, models/wsgan_emb_model.py:13


# ------------------------------------------------------------------ parameter layouts ----
def generator_keys(input_nc=3, output_nc=3, nz=1, ngf=64, n_blocks=9):
    """state_dict layout of ResnetGenerator (models/networks.py:565-607) with
    norm = InstanceNorm2d(affine=False, track_running_stats=True) (networks.py:25-26): name -> shape."""
    k = OrderedDict()

    def conv(name, co, ci, ks, transposed=False):
        k[name + ".weight"] = (ci, co, ks, ks) if transposed else (co, ci, ks, ks)
        k[name + ".bias"] = (co,)

    def inorm(name, c):
        k[name + ".running_mean"] = (c,)
        k[name + ".running_var"] = (c,)
        k[name + ".num_batches_tracked"] = ()

    conv("model.1", ngf, input_nc + nz, 7); inorm("model.2", ngf)
    conv("model.4", ngf * 2, ngf, 3); inorm("model.5", ngf * 2)
    conv("model.7", ngf * 4, ngf * 2, 3); inorm("model.8", ngf * 4)
    for i in range(n_blocks):
        p = "model.%d.conv_block" % (10 + i)
        conv(p + ".1", ngf * 4, ngf * 4, 3); inorm(p + ".2", ngf * 4)
        conv(p + ".5", ngf * 4, ngf * 4, 3); inorm(p + ".6", ngf * 4)
    b = 10 + n_blocks
    conv("model.%d" % b, ngf * 2, ngf * 4, 3, transposed=True); inorm("model.%d" % (b + 1), ngf * 2)
    conv("model.%d" % (b + 3), ngf, ngf * 2, 3, transposed=True); inorm("model.%d" % (b + 4), ngf)
    conv("model.%d" % (b + 7), output_nc, ngf, 7)
    return k


def discriminator_keys(input_nc=3, nz=1, ndf=64, n_layers=3):
    """state_dict layout of NLayerDiscriminator with BatchNorm2d (models/networks.py:737-777)."""
    k = OrderedDict()

    def bnorm(name, c):
        k[name + ".weight"] = (c,); k[name + ".bias"] = (c,)
        k[name + ".running_mean"] = (c,); k[name + ".running_var"] = (c,); k[name + ".num_batches_tracked"] = ()

    k["model.0.weight"] = (ndf, input_nc + nz, 4, 4); k["model.0.bias"] = (ndf,)
    idx, prev = 2, ndf
    for n in range(1, n_layers):
        cur = ndf * min(2 ** n, 8)
        k["model.%d.weight" % idx] = (cur, prev, 4, 4); bnorm("model.%d" % (idx + 1), cur)
        idx, prev = idx + 3, cur
    cur = ndf * min(2 ** n_layers, 8)
    k["model.%d.weight" % idx] = (cur, prev, 4, 4); bnorm("model.%d" % (idx + 1), cur)
    idx += 3
    k["model.%d.weight" % idx] = (1, cur, 4, 4); k["model.%d.bias" % idx] = (1,)
    return k


def encoder_keys(cnn_dim=(32, 1), noisy=False):
    """state_dict layout of SiameseFeature(ResNetFeature(resnet18)) (models/networks.py:1008-1049,1310-1343;
    models/resnet.py:31-56,125-179): name -> shape."""
    k = OrderedDict()

    def bnorm(name, c):
        k[name + ".weight"] = (c,); k[name + ".bias"] = (c,)
        k[name + ".running_mean"] = (c,); k[name + ".running_var"] = (c,); k[name + ".num_batches_tracked"] = ()

    k["base.model.conv1.weight"] = (64, 3, 7, 7); bnorm("base.model.bn1", 64)
    inpl = 64
    for li, planes in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            p = "base.model.layer%d.%d" % (li, bi)
            cin = inpl if bi == 0 else planes
            k[p + ".conv1.weight"] = (planes, cin, 3, 3); bnorm(p + ".bn1", planes)
            k[p + ".conv2.weight"] = (planes, planes, 3, 3); bnorm(p + ".bn2", planes)
            if bi == 0 and (li > 1):
                k[p + ".downsample.0.weight"] = (planes, cin, 1, 1); bnorm(p + ".downsample.1", planes)
        inpl = planes
    heads = ["cnn"] + (["cnn_logvar"] if noisy else [])
    for h in heads:
        prev, idx = 512, 0
        for nf in cnn_dim[:-1]:
            k["%s.%d.weight" % (h, idx)] = (nf, prev, 3, 3); k["%s.%d.bias" % (h, idx)] = (nf,)
            bnorm("%s.%d" % (h, idx + 1), nf)
            prev, idx = nf, idx + 4
        k["%s.%d.weight" % (h, idx)] = (cnn_dim[-1], prev, 3, 3); k["%s.%d.bias" % (h, idx)] = (cnn_dim[-1],)
    return k


def unet_keys(input_nc=3, output_nc=3, nz=1, ngf=64, num_downs=7):
    """state_dict layout of UnetGenerator with InstanceNorm2d(track_running_stats=True) (models/networks.py:659-733):
    nested UnetSkipConnectionBlocks, name -> shape, in the reference's order."""
    k = OrderedDict()
    inner = [ngf, ngf * 2, ngf * 4] + [ngf * 8] * (num_downs - 3)

    def inorm(name, c):
        k[name + ".running_mean"] = (c,); k[name + ".running_var"] = (c,); k[name + ".num_batches_tracked"] = ()

    def block(p, level):
        cin = input_nc + nz if level == 0 else inner[level - 1]
        outer = output_nc if level == 0 else inner[level - 1]
        ci = inner[level]
        if level == 0:
            k[p + ".0.weight"] = (ci, cin, 4, 4); k[p + ".0.bias"] = (ci,)
            block(p + ".1.model", 1)
            k[p + ".3.weight"] = (2 * ci, outer, 4, 4); k[p + ".3.bias"] = (outer,)
        elif level == num_downs - 1:
            k[p + ".1.weight"] = (ci, cin, 4, 4); k[p + ".1.bias"] = (ci,)
            k[p + ".3.weight"] = (ci, outer, 4, 4); k[p + ".3.bias"] = (outer,)
            inorm(p + ".4", outer)
        else:
            k[p + ".1.weight"] = (ci, cin, 4, 4); k[p + ".1.bias"] = (ci,)
            inorm(p + ".2", ci)
            block(p + ".3.model", level + 1)
            k[p + ".5.weight"] = (2 * ci, outer, 4, 4); k[p + ".5.bias"] = (outer,)
            inorm(p + ".6", outer)

    block("model.model", 0)
    return k


def unet_forward(sd, x, z, num_downs=7):
    """UnetGenerator.forward (networks.py:677-680) over UnetSkipConnectionBlock.forward (:728-733).  The blocks' LeakyReLU
    is in place (nn.LeakyReLU(0.2, True), :695), so the tensor that is concatenated as the skip connection is
    LeakyReLU(x), not x."""
    def conv(t, name):
        return F.conv2d(t, sd[name + ".weight"], sd[name + ".bias"], stride=2, padding=1)

    def convt(t, name):
        return F.conv_transpose2d(t, sd[name + ".weight"], sd[name + ".bias"], stride=2, padding=1)

    def block(p, t, level):
        a = F.leaky_relu(t, 0.2)
        if level == num_downs - 1:
            u = _inorm(convt(F.relu(conv(a, p + ".1")), p + ".3"), sd, p + ".4")
        else:
            d = _inorm(conv(a, p + ".1"), sd, p + ".2")
            u = _inorm(convt(F.relu(block(p + ".3.model", d, level + 1)), p + ".5"), sd, p + ".6")
        return torch.cat([a, u], 1)

    d = conv(_zcat(x, z), "model.model.0")
    return torch.tanh(convt(F.relu(block("model.model.1.model", d, 1)), "model.model.3"))


def alexnet_keys():
    """state_dict layout of AlexNetFeature (models/networks.py:1218-1240): name -> shape."""
    k = OrderedDict()
    for idx, (co, ci, ks) in ((0, (64, 3, 11)), (3, (192, 64, 5)), (6, (384, 192, 3)), (8, (256, 384, 3)), (10, (256, 256, 3))):
        k["features.%d.weight" % idx] = (co, ci, ks, ks)
        k["features.%d.bias" % idx] = (co,)
    return k


def fill_state_dict_(sd, seed, bias_std=0.02):
    """Deterministic fill in key order, in the spirit of init_weights('normal') (models/networks.py:72-93):
    conv weights ~ N(0, 0.02), norm weights ~ N(1, 0.02); biases get N(0, bias_std) (the reference zeroes them;
    non-zero biases make the parity checks see them). Works on any state_dict-like mapping, so the golden
    script applies the very same fill to the reference's modules."""
    g = torch.Generator().manual_seed(seed)
    for name, t in sd.items():
        with torch.no_grad():
            if name.endswith("num_batches_tracked"):
                t.zero_()
            elif name.endswith("running_mean"):
                t.zero_()
            elif name.endswith("running_var"):
                t.fill_(1.0)
            elif name.endswith(".bias"):
                t.copy_(torch.randn(t.shape, generator=g) * bias_std)
            elif t.dim() == 1:  # norm weight
                t.copy_(1.0 + torch.randn(t.shape, generator=g) * 0.02)
            else:
                t.copy_(torch.randn(t.shape, generator=g) * 0.02)
    return sd


def make_state_dict(keys, seed, device="cpu", requires_grad=False):
    sd = OrderedDict()
    for name, shape in keys.items():
        dt = torch.long if name.endswith("num_batches_tracked") else torch.float32
        sd[name] = torch.zeros(shape, dtype=dt)
    fill_state_dict_(sd, seed)
    out = OrderedDict()
    for name, t in sd.items():
        t = t.to(device)
        if requires_grad and t.dtype == torch.float32 and not name.endswith(("running_mean", "running_var")):
            t.requires_grad_(True)
        out[name] = t
    return out


# ------------------------------------------------------------- bf16 emulation (tests only) ----
class _GradRound(torch.autograd.Function):
    """identity forward; rounds the gradient to bf16 in backward (a gradient stored in a bf16 buffer)"""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class Quant:
    """Rounding model of the bf16 pipeline, applied to the fp32 restatement so that kernel bugs can be told from
    bf16 noise: operands of every convolution (activations, weights) and every stored activation / gradient are
    rounded to bf16 exactly where the kernels store bf16; accumulation, statistics, biases and weight gradients
    stay fp32.  `Quant(False)` is the exact fp32 reference arithmetic."""

    def __init__(self, on):
        self.on = on

    def fwd(self, t):      # value rounded to bf16 as stored; gradient passes unchanged
        return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach() if self.on else t

    def grad(self, t):     # gradient rounded to bf16 as stored
        return _GradRound.apply(t) if self.on else t

    def act(self, t):      # an activation buffer: value and incoming gradient are both bf16
        return self.grad(self.fwd(t)) if self.on else t

    def w(self, t):        # packed bf16 weight operand; the fp32 master weight receives an fp32 gradient
        return self.fwd(t)


EXACT = Quant(False)


# --------------------------------------------------------------------------- layers ----
def _inorm(x, sd, name, q=EXACT):
    """nn.InstanceNorm2d(affine=False, track_running_stats=True) in training mode (networks.py:25-26):
    instance statistics normalise; running stats get the EMA of the batch-mean of the instance stats.
    Under bf16 emulation the statistics come from the fp32 accumulator and are applied to its bf16 copy."""
    if not q.on:
        return F.instance_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], None, None, True, MOMENTUM, EPS)
    x = q.grad(x)
    mean = x.mean((2, 3), keepdim=True)
    var = x.var((2, 3), unbiased=False, keepdim=True)
    with torch.no_grad():
        n = x.size(2) * x.size(3)
        sd[name + ".running_mean"].mul_(1 - MOMENTUM).add_(MOMENTUM * mean.mean(0).flatten())
        sd[name + ".running_var"].mul_(1 - MOMENTUM).add_(MOMENTUM * (var.mean(0).flatten() * n / (n - 1)))
    return (q.fwd(x) - mean) * torch.rsqrt(var + EPS)


def _bnorm(x, sd, name, q=EXACT):
    """nn.BatchNorm2d in training mode (never .eval()'d on the train path: SURVEY appendix A.1)."""
    sd[name + ".num_batches_tracked"] += 1
    if not q.on:
        return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                            True, MOMENTUM, EPS)
    x = q.grad(x)
    mean = x.mean((0, 2, 3), keepdim=True)
    var = x.var((0, 2, 3), unbiased=False, keepdim=True)
    with torch.no_grad():
        n = x.numel() // x.size(1)
        sd[name + ".running_mean"].mul_(1 - MOMENTUM).add_(MOMENTUM * mean.flatten())
        sd[name + ".running_var"].mul_(1 - MOMENTUM).add_(MOMENTUM * (var.flatten() * n / (n - 1)))
    return (q.fwd(x) - mean) * torch.rsqrt(var + EPS) * sd[name + ".weight"].view(1, -1, 1, 1) + sd[name + ".bias"].view(1, -1, 1, 1)


def _rpad(h, p, q):
    return q.grad(F.pad(h, (p,) * 4, mode="reflect"))


def _zcat(x, z):
    """torch.cat((input, z.expand(H, W)), 1)  (networks.py:610-611, :780-782)"""
    zi = z.view(z.size(0), z.size(1), 1, 1).expand(x.size(0), z.size(1), x.size(2), x.size(3))
    return torch.cat((x, zi), 1)


def generator_forward(sd, x, z, n_blocks=9, taps=None, q=EXACT):
    """ResnetGenerator.forward (networks.py:609-612) with the Sequential of :578-605 and ResnetBlock :621-652.
    taps (optional dict) receives the output of every conv and of every block for per-layer checks."""
    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    W = lambda k: q.w(sd[k])
    h = q.act(_zcat(x, z))
    h = rec("model.1", F.conv2d(_rpad(h, 3, q), W("model.1.weight"), sd["model.1.bias"]))
    h = q.act(F.relu(_inorm(h, sd, "model.2", q)))
    h = rec("model.4", F.conv2d(h, W("model.4.weight"), sd["model.4.bias"], stride=2, padding=1))
    h = q.act(F.relu(_inorm(h, sd, "model.5", q)))
    h = rec("model.7", F.conv2d(h, W("model.7.weight"), sd["model.7.bias"], stride=2, padding=1))
    h = rec("act.model.8", q.act(F.relu(_inorm(h, sd, "model.8", q))))
    for i in range(n_blocks):
        p = "model.%d.conv_block" % (10 + i)
        r = rec(p + ".1", F.conv2d(_rpad(h, 1, q), W(p + ".1.weight"), sd[p + ".1.bias"]))
        r = rec("act." + p + ".2", q.act(F.relu(_inorm(r, sd, p + ".2", q))))
        r = rec(p + ".5", F.conv2d(_rpad(r, 1, q), W(p + ".5.weight"), sd[p + ".5.bias"]))
        r = _inorm(r, sd, p + ".6", q)
        h = rec("model.%d" % (10 + i), q.act(h + r))
    b = 10 + n_blocks
    h = rec("model.%d" % b, F.conv_transpose2d(h, W("model.%d.weight" % b), sd["model.%d.bias" % b], stride=2, padding=1, output_padding=1))
    h = q.act(F.relu(_inorm(h, sd, "model.%d" % (b + 1), q)))
    h = rec("model.%d" % (b + 3), F.conv_transpose2d(h, W("model.%d.weight" % (b + 3)), sd["model.%d.bias" % (b + 3)], stride=2, padding=1, output_padding=1))
    h = q.act(F.relu(_inorm(h, sd, "model.%d" % (b + 4), q)))
    h = rec("model.%d" % (b + 7), F.conv2d(_rpad(h, 3, q), W("model.%d.weight" % (b + 7)), sd["model.%d.bias" % (b + 7)]))
    return torch.tanh(q.grad(h))


def discriminator_forward(sd, x, z=None, n_layers=3, use_sigmoid=True, taps=None, q=EXACT):
    """NLayerDiscriminator.forward (networks.py:779-783) over the Sequential of :745-777."""
    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    W = lambda k: q.w(sd[k])
    h = q.act(_zcat(x, z) if z is not None else x)
    h = q.act(F.leaky_relu(q.grad(rec("model.0", F.conv2d(h, W("model.0.weight"), sd["model.0.bias"], stride=2, padding=1))), 0.2))
    idx = 2
    for n in range(1, n_layers + 1):
        stride = 2 if n < n_layers else 1
        h = rec("model.%d" % idx, F.conv2d(h, W("model.%d.weight" % idx), None, stride=stride, padding=1))
        h = q.act(F.leaky_relu(_bnorm(h, sd, "model.%d" % (idx + 1), q), 0.2))
        idx += 3
    h = q.grad(rec("model.%d" % idx, F.conv2d(h, W("model.%d.weight" % idx), sd["model.%d.bias" % idx], stride=1, padding=1)))
    return torch.sigmoid(h) if use_sigmoid else h


def _basic_block(sd, p, x, stride, drop=None, q=EXACT):
    """BasicBlock.forward (models/resnet.py:55-73): conv-[drop]-bn-relu-conv-[drop]-bn (+downsample) add relu."""
    W = lambda k: q.w(sd[k])
    out = F.conv2d(x, W(p + ".conv1.weight"), None, stride=stride, padding=1)
    if drop is not None:
        out = drop(out)
    out = q.act(F.relu(_bnorm(out, sd, p + ".bn1", q)))
    out = F.conv2d(out, W(p + ".conv2.weight"), None, stride=1, padding=1)
    if drop is not None:
        out = drop(out)
    out = _bnorm(out, sd, p + ".bn2", q)
    if (p + ".downsample.0.weight") in sd:
        idn = _bnorm(F.conv2d(x, W(p + ".downsample.0.weight"), None, stride=stride), sd, p + ".downsample.1", q)
    else:
        idn = x
    return q.act(F.relu(out + idn))


def encoder_forward(sd, x, cnn_dim=(32, 1), cnn_relu_slope=0.7, noisy=False, drop=None, taps=None, q=EXACT):
    """SiameseFeature.forward (networks.py:1051-1068) over ResNetFeature.forward (:1344-1354) and the ResNet-18
    trunk (resnet.py:134-138,163-179); pooling='avg' -> nn.AvgPool2d(full size).  `drop`, if given, is a callable
    applied where the reference places nn.Dropout2d (resnet.py:58,63; networks.py:1022)."""
    W = lambda k: q.w(sd[k])
    h = F.conv2d(q.act(x), W("base.model.conv1.weight"), None, stride=2, padding=3)
    h = q.act(F.relu(_bnorm(h, sd, "base.model.bn1", q)))
    h = q.act(F.max_pool2d(h, 3, 2, 1))
    if taps is not None:
        taps["stem"] = h
    for li in range(1, 5):
        for bi in range(2):
            h = _basic_block(sd, "base.model.layer%d.%d" % (li, bi), h, 2 if (li > 1 and bi == 0) else 1, drop, q)
        if taps is not None:
            taps["layer%d" % li] = h

    def head(prefix, t):
        idx = 0
        for _ in cnn_dim[:-1]:
            t = F.conv2d(t, W("%s.%d.weight" % (prefix, idx)), sd["%s.%d.bias" % (prefix, idx)], padding=1)
            t = _bnorm(t, sd, "%s.%d" % (prefix, idx + 1), q)
            if drop is not None:
                t = drop(t)
            t = q.act(F.leaky_relu(t, cnn_relu_slope))
            idx += 4
        t = q.grad(F.conv2d(t, W("%s.%d.weight" % (prefix, idx)), sd["%s.%d.bias" % (prefix, idx)], padding=1))
        return F.avg_pool2d(t, t.size(2))

    y = head("cnn", h)
    if noisy:
        return y, head("cnn_logvar", h)
    return y


def alexnet_forward(sd, x):
    """AlexNetFeature.forward with pooling 'None' (models/networks.py:1218-1248): the torchvision AlexNet feature stack."""
    h = F.relu(F.conv2d(x, sd["features.0.weight"], sd["features.0.bias"], stride=4, padding=2))
    h = F.max_pool2d(h, 3, 2)
    h = F.relu(F.conv2d(h, sd["features.3.weight"], sd["features.3.bias"], padding=2))
    h = F.max_pool2d(h, 3, 2)
    h = F.relu(F.conv2d(h, sd["features.6.weight"], sd["features.6.bias"], padding=1))
    h = F.relu(F.conv2d(h, sd["features.8.weight"], sd["features.8.bias"], padding=1))
    h = F.relu(F.conv2d(h, sd["features.10.weight"], sd["features.10.bias"], padding=1))
    return F.max_pool2d(h, 3, 2)


def upsample2d(x, size):
    """util/util.py:111-117"""
    if size <= 0 or x.size(2) == size:
        return x
    return F.interpolate(x, size=(size, size), mode="bilinear", align_corners=True)


def gan_loss(pred, target_label, use_lsgan=False):
    """GANLoss.__call__ (networks.py:407-420) for one prediction tensor: per-sample targets (bool/int/list)
    expanded to the prediction's shape, then nn.BCELoss (use_lsgan False) or nn.MSELoss."""
    if not isinstance(target_label, (list, tuple)):
        target_label = [target_label]
    vals = [float(int(t)) for t in target_label]
    t = torch.tensor(vals, dtype=pred.dtype, device=pred.device).view(len(vals), 1, 1, 1).expand_as(pred)
    return F.mse_loss(pred, t) if use_lsgan else F.binary_cross_entropy(pred, t)


def elo_nll(prob, label):
    """BinaryNLLLoss (networks.py:473-482): targets LUT[label] in {0, .5, 1}."""
    lut = torch.tensor([0.0, 0.5, 1.0], device=prob.device)
    t = lut[label].view(prob.size(0), 1, 1, 1).expand_as(prob)
    return -(t * torch.log(prob + MAGIC_EPS) + (1 - t) * torch.log(1 - prob + MAGIC_EPS)).mean()


# ------------------------------------------------------------------------ the step ----
class WSGANEmbOracle:
    """WSGANEmbModel with default flags (models/wsgan_emb_model.py): lr_E = 0 (E frozen but in train mode),
    plain encoder, sigmoid + BCE GAN loss, lambda_IP configurable only as 0 (netIP is outside the named path)."""

    def __init__(self, sd_g, sd_d, sd_e, *, lr=2e-4, beta1=0.5, lambda_z=1.0, lambda_a=0.5, lambda_l1=0.0,
                 lambda_a_gan=0.0, fine_size_e=224, relabel_d=(0, 1, 0), emb_mean=0.0, emb_std=1.0, n_blocks=9,
                 n_layers_d=3, detach_fake_b=False, bayesian=False, noisy=False, noisy_var_type="", bnn_T=10,
                 noisy_d=True, noisy_rec=True, dropout=False, drop_masks=None, eps_queue=None, use_real_a=False,
                 sd_ip=None, lambda_ip=0.0, fine_size_ip=224, ip_criterion="mse", generator="resnet", num_downs=7):
        """generator: "resnet" (--which_model_netG resnet_9blocks, n_blocks) or "unet" (unet_128 / unet_256, num_downs).
        bayesian / noisy / noisy_var_type / bnn_T / noisy_D / noisy_rec: the encoder modes of forward() (:218-240)
        and backward_G (:408-430).  Randomness is injected, never drawn: `drop_masks` is a list of Dropout2d masks
        [N, C] consumed in module order (dropout=True places them where the reference has nn.Dropout2d), `eps_queue`
        the standard-normal draws of util.resample (util/util.py:136-139) in call order."""
        self.bayesian, self.noisy, self.nvt, self.T = bayesian, noisy, noisy_var_type, bnn_T
        self.noisy_d, self.noisy_rec, self.dropout = noisy_d, noisy_rec, dropout
        self.drop_masks = drop_masks if drop_masks is not None else []
        self.eps_queue = eps_queue if eps_queue is not None else []
        self.g, self.d, self.e = sd_g, sd_d, sd_e
        self.pg = [t for t in sd_g.values() if t.requires_grad]
        self.pd = [t for t in sd_d.values() if t.requires_grad]
        # wsgan_emb_model.py:153-154
        self.opt_g = torch.optim.Adam(self.pg, lr=lr, betas=(beta1, 0.999))
        self.opt_d = torch.optim.Adam(self.pd, lr=lr, betas=(beta1, 0.999))
        self.lz, self.la, self.l1, self.lag = lambda_z, lambda_a, lambda_l1, lambda_a_gan
        self.fe, self.relabel = fine_size_e, list(relabel_d)
        self.mean, self.std = emb_mean, emb_std
        self.nb, self.nld, self.detach_fake_b = n_blocks, n_layers_d, detach_fake_b
        self.use_real_a = use_real_a     # --use_real_A (:309-322): D's real pairs are built from real_A
        # identity-preserving loss (:130-135, 353-356, 393-396): AlexNet features of fake_B against those of real_A
        self.ip, self.lip, self.fip, self.ip_crit = sd_ip, lambda_ip, fine_size_ip, ip_criterion
        self.generator, self.num_downs = generator, num_downs
        self.losses = {}

    def _G(self, x, z):
        if self.generator == "unet":
            return unet_forward(self.g, x, z, self.num_downs)
        return generator_forward(self.g, x, z, self.nb)

    def _drop(self, t):
        m = self.drop_masks.pop(0).to(t.device, t.dtype)
        return t * m.view(t.size(0), t.size(1), 1, 1)

    def _E(self, x):
        return encoder_forward(self.e, x, noisy=self.noisy, drop=self._drop if self.dropout else None)

    def _mu_var(self, x):
        """util.compute_mu_and_var (util/util.py:153-171)"""
        y_mu, y_sq, s2_mu = 0.0, 0.0, 0.0
        for _ in range(self.T):
            out = self._E(x)
            y, logs2 = out if self.noisy else (out, None)
            y_mu = y_mu + 1.0 / self.T * y
            y_sq = y_sq + 1.0 / self.T * y ** 2
            if self.noisy:
                s2_mu = s2_mu + 1.0 / self.T * torch.exp(logs2)
        return (y_mu, y_sq - y_mu ** 2, s2_mu) if self.noisy else (y_mu, y_sq - y_mu ** 2)

    def _encode(self, x):
        """(y, resampling variance or None): the four branches of forward() (:218-240)"""
        var = None
        if not self.bayesian and not self.noisy:
            y = self._E(x)
        elif not self.bayesian and self.noisy:
            y, logvar = self._E(x)
            if "a" in self.nvt:
                var = torch.exp(logvar)
        elif self.bayesian and not self.noisy:
            y, y_var = self._mu_var(x)
            if "e" in self.nvt:
                var = y_var
        else:
            y, y_var, y_s2 = self._mu_var(x)
            if "a" in self.nvt:
                var = y_s2 + y_var
        return y, var

    def _resample(self, mu, var):
        eps = self.eps_queue.pop(0).to(mu.device, mu.dtype)
        return mu + eps.view_as(mu) * torch.sqrt(var)

    def _norm(self, y):
        return (y - self.mean) / self.std

    def forward(self, real_a, real_b):
        """WSGANEmbModel.forward (:214-259), lr_E <= 0."""
        with torch.no_grad():
            self.real_a_e = upsample2d(real_a, self.fe)
            self.y_a, var_a = self._encode(self.real_a_e)
            self.y_b, var_b = self._encode(upsample2d(real_b, self.fe))
            if var_a is not None:
                self.res_a = self._norm(self._resample(self.y_a, var_a))
                self.res_b = self._norm(self._resample(self.y_b, var_b))
        self.emb_a, self.emb_b = self._norm(self.y_a), self._norm(self.y_b)
        self.cond_b = self.res_b if (self.nvt and self.noisy_d) else self.emb_b
        self.real_a, self.real_b = real_a, real_b
        self.fake_b = self._G(real_a, self.emb_b)
        src = self.fake_b.detach() if self.detach_fake_b else self.fake_b
        self.rec_a = self._G(src, self.emb_a)

    def backward_g(self):
        """WSGANEmbModel.backward_G (:371-437)."""
        for t in self.pd:
            t.requires_grad_(False)   # set_requires_grad(netD, False) (:458)
        self.opt_g.zero_grad()
        L = {}
        pred = discriminator_forward(self.d, self.fake_b, self.cond_b, self.nld)
        L["G_GAN"] = gan_loss(pred, True)
        total = L["G_GAN"]
        if self.lag > 0:
            L["G_GAN_cycle"] = gan_loss(discriminator_forward(self.d, self.rec_a, self.emb_a, self.nld), True) * self.lag
            total = total + L["G_GAN_cycle"]
        if self.lip > 0:
            with torch.no_grad():
                feat_a = alexnet_forward(self.ip, upsample2d(self.real_a, self.fip))
            feat_f = alexnet_forward(self.ip, upsample2d(self.fake_b, self.fip))
            L["G_IP"] = (F.mse_loss if self.ip_crit == "mse" else F.l1_loss)(feat_f, feat_a) * self.lip
            total = total + L["G_IP"]
        if self.l1 > 0:
            L["G_L1"] = F.l1_loss(self.fake_b, self.real_a) * self.l1
            total = total + L["G_L1"]
        if self.la > 0:
            L["G_cycle"] = F.l1_loss(self.rec_a, self.real_a) * self.la
            total = total + L["G_cycle"]
        if self.lz > 0:
            fake_e = upsample2d(self.fake_b, self.fe)
            y_var = y_logvar = None
            if not self.bayesian and not self.noisy:
                pred_y = self._E(fake_e)
            elif not self.bayesian and self.noisy:
                pred_y, y_logvar = self._E(fake_e)
                if "a" in self.nvt:
                    y_var = torch.exp(y_logvar)
            elif self.bayesian and not self.noisy:
                pred_y, y_var = self._mu_var(fake_e)
                if "e" in self.nvt:
                    y_logvar = torch.log(y_var + MAGIC_EPS)
            else:   # (:418-425) the reference feeds real_A_E here
                pred_y, y_var_, y_s2_ = self._mu_var(self.real_a_e)
                y_var = torch.zeros_like(pred_y)
                if "a" in self.nvt:
                    y_var = y_var + y_s2_
                if "e" in self.nvt:
                    y_var = y_var + y_var_
                y_logvar = torch.log(y_var + MAGIC_EPS)
            if self.nvt and self.noisy_rec:
                L["z_rec"] = ((pred_y - self.y_b).pow(2) / y_var.detach() + y_logvar.detach()).sum() / pred_y.size(0) * 0.5 * self.lz
            else:
                L["z_rec"] = F.mse_loss(pred_y, self.y_b) * self.lz
            total = total + L["z_rec"]
        total.backward()
        self.opt_g.step()
        for t in self.pd:
            t.requires_grad_(True)
        return L

    def backward_d(self, label):
        """WSGANEmbModel.backward_D (:300-329)."""
        self.opt_d.zero_grad()
        L = {}
        L["D_fake"] = gan_loss(discriminator_forward(self.d, self.fake_b.detach(), self.cond_b, self.nld), False)
        img, right, wrong = (self.real_a, self.emb_a, self.emb_b) if self.use_real_a else (self.real_b, self.emb_b, self.emb_a)
        L["D_real_right"] = gan_loss(discriminator_forward(self.d, img, right, self.nld), True)
        target = [self.relabel[int(l)] for l in label]
        L["D_real_wrong"] = gan_loss(discriminator_forward(self.d, img, wrong, self.nld), target)
        total = (L["D_fake"] + (L["D_real_right"] + L["D_real_wrong"]) * 0.5) * 0.5
        total.backward()
        self.opt_d.step()
        return L

    def optimize_parameters(self, real_a, real_b, label):
        """WSGANEmbModel.optimize_parameters (:478-484): forward, update_G, update_D."""
        self.forward(real_a, real_b)
        L = self.backward_g()
        L.update(self.backward_d(label))
        self.losses = {k: float(v) for k, v in L.items()}
        return self.losses


class SiameseOracle:
    """The Elo rating trainer's step (siamese.py:590-686, plain branch :673-678) on SiameseNetwork(resnet18, cnn_dim=[32, 1],
    fc_dim=[]) (networks.py:971-992): two encoder passes with separate BatchNorm batches, score = f1 - f2, sigmoid,
    BinaryNLLLoss, Adam(lr) over every parameter (siamese.py:544-551)."""

    def __init__(self, sd, lr=2e-4, cnn_relu_slope=0.7):
        self.sd, self.slope = sd, cnn_relu_slope
        self.params = [t for t in sd.values() if t.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=lr)

    def step(self, img0, img1, label):
        self.opt.zero_grad()
        f1 = encoder_forward(self.sd, img0, cnn_relu_slope=self.slope)
        f2 = encoder_forward(self.sd, img1, cnn_relu_slope=self.slope)
        prob = torch.sigmoid(f1 - f2)
        loss = elo_nll(prob, label)
        loss.backward()
        self.opt.step()
        return float(loss), prob.detach()


def synthetic_batch(batch, size, seed, device="cpu"):
    """UTKFace-shaped synthetic pairs: RGB in [-1, 1], label in {0, 1, 2}."""
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    b = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    label = torch.randint(0, 3, (batch,), generator=g)
    return a.to(device), b.to(device), label
