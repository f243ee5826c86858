"""Every convolution plan the BASELINE step builds (B = 64, 128 x 128, encoder at 224 x 224: configs[2]; and B = 32 at
256 x 256: configs[4]) against torch.nn.functional on the same bf16-rounded operands: forward (+ fused bias, activation,
statistics), data gradient and weight gradient of every ConvRT of the three programs, launched exactly as the step
launches them (ConvRT.forward / backward_data / backward_weight -> C ABI).

Products of bf16 operands are exact in fp32, so only the summation order differs: rel-L2 <= 2e-5 for fp32 outputs (NCHW
network outputs); bf16-stored outputs add one rounding (<= 4e-3).  Weight gradients sum up to 1e6 pixels per element in
fp32 (tensor-core accumulation + split-K atomics): they are compared with an fp64 reference at 4e-5."""
import pytest
import torch
import torch.nn.functional as F

from pcgan_b200 import _lib as L
from pcgan_b200 import networks as NW
from pcgan_b200.conv import out_size
from pcgan_b200.plan import SLACK

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def bf(t):
    return t.to(torch.bfloat16).float()


def tf32_round(t):
    """fp32 -> nearest TF32 (10-bit mantissa), ties away from zero: what cvt.rna.tf32.f32 does.  A TF32-mode producer
    stores its activations this way (the tensor core then truncates nothing), exactly as cuDNN rounds its fp32 inputs."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def rand_buf(g, halo_zero, valid_c, gen, dtype=torch.bfloat16):
    """Padded NHWC buffer of geometry g (bf16, or fp32 for TF32 plans): random interior (channels >= valid_c zero),
    random or zero halo."""
    t = torch.randn(g.n, g.hp, g.wp, g.c, device=DEV, generator=gen)
    t[..., valid_c:] = 0
    if halo_zero and g.pad:
        m = torch.zeros(g.hp, g.wp, device=DEV)
        m[g.pad:g.pad + g.h, g.pad:g.pad + g.w] = 1
        t = t * m.view(1, g.hp, g.wp, 1)
    if dtype == torch.float32:      # TF32 mode: the buffer holds TF32-rounded values, the reference sees the fp32 originals
        flat = torch.cat([tf32_round(t).reshape(-1), torch.zeros(SLACK, dtype=dtype, device=DEV)])
        return flat, t.permute(0, 3, 1, 2)
    flat = torch.cat([t.to(dtype).reshape(-1), torch.zeros(SLACK, dtype=dtype, device=DEV)])
    return flat, t.to(dtype).float().permute(0, 3, 1, 2)      # flat buffer, [N, C, Hp, Wp] fp32 view of it


def alloc_out(om, n, c, h, w):
    numel = om.base + (n - 1) * om.sn + (h - 1) * om.sy + (w - 1) * om.sx + (c - 1) * om.sc + 1 + SLACK
    return torch.zeros(numel, dtype=torch.bfloat16 if om.dtype == L.DT_BF16 else torch.float32, device=DEV)


def read_out(buf, om, n, c, h, w):
    return buf.as_strided((n, c, h, w), (om.sn, om.sc, om.sy, om.sx), om.base).float()


def act_ref(y, act, slope):
    if act == L.ACT_RELU:
        return torch.relu(y)
    if act == L.ACT_LRELU:
        return F.leaky_relu(y, slope)
    if act == L.ACT_TANH:
        return torch.tanh(y)
    if act == L.ACT_SIGMOID:
        return torch.sigmoid(y)
    return y


def tf32_twin(conv):
    """The same convolution as a TF32 runtime: fp32 buffers and outputs, tcgen05.mma.kind::tf32."""
    from dataclasses import replace
    from pcgan_b200.engine import ConvRT
    q = conv.geometry
    f32 = lambda om: replace(om, dtype=L.DT_F32) if om is not None else None
    return ConvRT(conv.name + "[tf32]", conv.weight, conv.bias, q["xg"], q["stride"], q["cp"], f32(q["out"]), transposed=q["transposed"],
                  output_padding=q["output_padding"], act=q["act"], act_slope=q["act_slope"], stats=q["stats"],
                  per_sample_stats=q["per_sample_stats"], dyg=q["dyg"], dx_out=f32(q["dx_out"]), full_padded=q["full_padded"], tf32=True)


def check_conv(conv, gen, log, tf32=False):
    q = conv.geometry
    dt = torch.float32 if tf32 else torch.bfloat16
    # bf16: products are exact, only summation order differs.  TF32: BASELINE.json's per-layer gate, 1e-3.
    T_OUT16, T_OUT32, T_W, T_S = (1e-3, 1e-3, 1e-3, 2e-3) if tf32 else (4e-3, 2e-5, 4e-5, 1e-4)
    xg, stride, cp, tr = q["xg"], q["stride"], q["cp"], q["transposed"]
    w = conv.weight.detach()
    k = w.shape[2]
    cin, cout = (w.shape[0], w.shape[1]) if tr else (w.shape[1], w.shape[0])
    wq = (w if tf32 else bf(w)).clone().requires_grad_(True)
    bias = conv.bias.detach() if conv.bias is not None else None
    o = xg.pad - cp

    def fwd_ref(xfull):
        if tr:
            xi = xfull[:, :cin, xg.pad:xg.pad + xg.h, xg.pad:xg.pad + xg.w]
            return F.conv_transpose2d(xi, wq, bias, stride=stride, padding=cp, output_padding=q["output_padding"])
        xi = xfull[:, :cin, o:xg.hp - o, o:xg.wp - o]
        return F.conv2d(xi, wq, bias, stride=stride)

    # ---- forward
    xbuf, xfull = rand_buf(xg, halo_zero=tr, valid_c=cin, gen=gen, dtype=dt)
    ho = out_size(xg.h, k, stride, cp, tr, q["output_padding"])
    om = q["out"]
    out = alloc_out(om, xg.n, cout, ho, ho)
    groups = xg.n if q["per_sample_stats"] else 1
    stats = torch.zeros(groups, cout, 2, device=DEV) if q["stats"] else None
    conv._wver = None
    conv.forward(xbuf, out, stats)
    with torch.no_grad():
        pre = fwd_ref(xfull)
    e = rel(read_out(out, om, xg.n, cout, ho, ho), act_ref(pre, q["act"], q["act_slope"]))
    log.append((conv.name + ".fwd", e))
    assert e < (T_OUT16 if om.dtype == L.DT_BF16 else T_OUT32), (conv.name, "fwd", e)
    if stats is not None:
        dims = (2, 3) if q["per_sample_stats"] else (0, 2, 3)
        s1, s2 = pre.sum(dims).view(groups, cout), (pre * pre).sum(dims).view(groups, cout)
        e1, e2 = rel(stats[..., 0], s1), rel(stats[..., 1], s2)
        log.append((conv.name + ".stats", max(e1, e2)))
        # sums of ~1e4-1e6 fp32 terms in a different order; sum(x) of a zero-mean channel is a cancellation, so it is
        # bounded against sqrt(count * sum(x^2)) >= |sum(x)|
        cnt = ho * ho * (xg.n // groups)
        assert e2 < T_S and bool(((stats[..., 0] - s1).abs() <= T_S * (cnt * s2).sqrt() + 1e-6).all()), (conv.name, "stats", e1, e2)
    # ---- backward operands
    dyg = q["dyg"]
    if dyg is None:
        return
    dybuf, dyfull = rand_buf(dyg, halo_zero=True, valid_c=cout, gen=gen, dtype=dt)
    assert dyg.h == ho, (conv.name, dyg, ho)
    dy = dyfull[:, :cout, dyg.pad:dyg.pad + dyg.h, dyg.pad:dyg.pad + dyg.w].contiguous()
    # the backward reference runs in fp64 (its own fp32 summation error over ~1e6 pixels would be as large as ours)
    xr = xfull.double().requires_grad_(True)
    wq = wq.detach().double().requires_grad_(True)
    bias = bias.double() if bias is not None else None
    fwd_ref(xr).backward(dy.double())
    # ---- data gradient
    if conv.dgrad:
        dm = q["dx_out"]
        cb = xg.c
        if q["full_padded"]:
            hh, want = xg.hp, xr.grad[:, :cin]
        else:
            hh, want = xg.h, xr.grad[:, :cin, xg.pad:xg.pad + xg.h, xg.pad:xg.pad + xg.w]
        dx = alloc_out(dm, xg.n, cb, hh, hh)
        conv.backward_data(dybuf, dx)
        e = rel(read_out(dx, dm, xg.n, cb, hh, hh)[:, :cin], want)
        log.append((conv.name + ".dgrad", e))
        assert e < (T_OUT16 if dm.dtype == L.DT_BF16 else T_OUT32), (conv.name, "dgrad", e)
    # ---- weight gradient
    conv.weight.grad = None
    conv.weight.requires_grad_(True)
    conv.backward_weight(dybuf, xbuf)
    e = rel(conv.weight.grad, wq.grad)
    log.append((conv.name + ".wgrad", e))
    assert e < T_W, (conv.name, "wgrad", e)
    conv.weight.grad = None


def _programs(B, S, SE, nb=9):
    g = NW.init_net(NW.ResnetGenerator(3, 3, 1, 64, norm_layer=NW.get_norm_layer("instance"), n_blocks=nb), "normal", [0]).module
    d = NW.define_D(3, 1, 64, "n_layers", 3, "batch", True, "normal", gpu_ids=[0]).module
    e = NW.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, gpu_ids=[0]).module
    return g._program(B, S), d._program(B, S), e._program(B, SE)


def _distinct(convs):
    """one representative per distinct set of plans (the nine resblocks share theirs)"""
    seen, out = set(), []
    for c in convs:
        q = c.geometry
        key = (tuple(c.weight.shape), q["xg"], q["stride"], q["cp"], q["transposed"], q["act"], q["stats"], q["per_sample_stats"],
               q["dyg"], q["full_padded"], (q["out"].base, q["out"].sn, q["out"].sy, q["out"].sx, q["out"].sc, q["out"].dtype))
        if key not in seen:
            seen.add(key)
            out.append(c)
    return out


@pytest.mark.parametrize("B,S,SE", [(64, 128, 224), (32, 256, 224), (5, 64, 96)], ids=["c128_b64", "c256_b32", "ragged_b5"])
def test_every_plan_of_the_step(B, S, SE):
    gen = torch.Generator(device=DEV).manual_seed(B * 1000 + S)
    log = []
    progs = _programs(B, S, SE)
    n = 0
    for prog in progs:
        for conv in _distinct(prog.bank.convs):
            check_conv(conv, gen, log)
            n += 1
    torch.cuda.synchronize()
    worst = sorted(log, key=lambda t: -t[1])[:8]
    print("B=%d S=%d: %d convolutions, %d checks; worst:" % (B, S, n, len(log)), ["%s %.2e" % t for t in worst])
    assert n >= 20


@pytest.mark.parametrize("B,S,SE", [(64, 128, 224), (3, 64, 96)], ids=["c128_b64", "ragged_b3"])
def test_every_plan_of_the_step_tf32(B, S, SE):
    """The same sweep in TF32 mode: every convolution of G, D and E re-planned with tf32=True (fp32 buffers, fp32 packed
    weights rounded to TF32, tcgen05.mma.kind::tf32) within BASELINE.json's per-layer TF32 gate of 1e-3."""
    gen = torch.Generator(device=DEV).manual_seed(B * 1000 + S + 1)
    log, n = [], 0
    for prog in _programs(B, S, SE):
        for conv in _distinct(prog.bank.convs):
            check_conv(tf32_twin(conv), gen, log, tf32=True)
            n += 1
    torch.cuda.synchronize()
    worst = sorted(log, key=lambda t: -t[1])[:8]
    print("TF32 B=%d S=%d: %d convolutions, %d checks; worst:" % (B, S, n, len(log)), ["%s %.2e" % t for t in worst])
    assert n >= 20 and max(e for name, e in log if not name.endswith(".stats")) > 1e-5      # really TF32 products, not fp32
