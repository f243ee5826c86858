"""Module-level parity on the GPU: define_G / define_D / define_E / GANLoss of pcgan_b200 against the oracle
restatement of the reference (oracle/pcgan_oracle.py, pinned by tests/test_oracle_cpu.py), same state_dict, same
inputs.  The oracle runs in strict fp32 (TF32 off) on the same device.

Tolerances (BASELINE.json north_star, BF16 mode): per-layer 2e-2 relative L2 teacher-forced — covered conv by conv in
tests/test_igemm_gpu.py — while whole-network outputs compound the bf16 rounding of ~25 stacked layers; SURVEY §4
measured 2.4e-2 on activations end-to-end for exact bf16 emulation, so whole-network gates are 4e-2 on outputs and
1e-1 on gradients (printed, so drifts are visible)."""
import pytest
import torch

from oracle import pcgan_oracle as O
from pcgan_b200 import networks as NW

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def load_into(net, sd):
    mod = net.module if hasattr(net, "module") else net
    mod.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
    return mod


@pytest.mark.parametrize("N,S", [(2, 32), (2, 128)])
def test_generator_forward_backward(N, S):
    sd = O.make_state_dict(O.generator_keys(), 41, device=DEV, requires_grad=True)
    net = NW.define_G(3, 3, 1, 64, "resnet_9blocks", "instance", init_type="normal", gpu_ids=[0])
    mod = load_into(net, sd)
    a, _, _ = O.synthetic_batch(N, S, 300, device=DEV)
    z = torch.linspace(-1, 1, N, device=DEV).view(N, 1, 1, 1)
    w = torch.randn(N, 3, S, S, device=DEV)
    a1 = a.clone().requires_grad_(True)
    out = net(a1, z)
    (out * w).sum().backward()
    a2 = a.clone().requires_grad_(True)
    ref = O.generator_forward(sd, a2, z)
    (ref * w).sum().backward()
    e_out = rel(out, ref)
    e_dx = rel(a1.grad, a2.grad)
    errs = {}
    for k in ("model.1.weight", "model.4.weight", "model.10.conv_block.1.weight", "model.14.conv_block.5.weight",
              "model.18.conv_block.5.weight", "model.19.weight", "model.22.weight", "model.26.weight", "model.26.bias"):
        errs[k] = rel(mod.state_dict(keep_vars=True)[k].grad, sd[k].grad)
    print("G N=%d S=%d out %.3e dx %.3e" % (N, S, e_out, e_dx), {k: "%.2e" % v for k, v in errs.items()})
    assert e_out < 4e-2 and e_dx < 1e-1
    assert max(errs.values()) < 1e-1
    # running statistics of the instance norms follow the reference's EMA
    for k in ("model.2.running_mean", "model.2.running_var", "model.11.conv_block.6.running_var"):
        assert rel(mod.state_dict()[k], sd[k]) < 2e-2, k


@pytest.mark.parametrize("N,S", [(3, 32), (4, 128)])
def test_discriminator_forward_backward_with_ganloss(N, S):
    sd = O.make_state_dict(O.discriminator_keys(), 42, device=DEV, requires_grad=True)
    net = NW.define_D(3, 1, 64, "n_layers", 3, "batch", True, "normal", gpu_ids=[0])
    mod = load_into(net, sd)
    a, _, _ = O.synthetic_batch(N, S, 301, device=DEV)
    z = torch.linspace(-1, 1, N, device=DEV).view(N, 1, 1, 1)
    target = [1, 0, 1, 0][:N]
    a1 = a.clone().requires_grad_(True)
    out = net(a1, z)
    loss = NW.GANLoss(use_lsgan=False)(out, target)
    loss.backward()
    a2 = a.clone().requires_grad_(True)
    ref = O.discriminator_forward(sd, a2, z)
    lref = O.gan_loss(ref, target)
    lref.backward()
    errs = {k: rel(p.grad, sd[k].grad) for k, p in mod.named_parameters()}
    print("D N=%d S=%d out %.3e loss %.6f/%.6f dx %.3e" % (N, S, rel(out, ref), float(loss), float(lref), rel(a1.grad, a2.grad)),
          {k: "%.2e" % v for k, v in errs.items()})
    assert rel(out, ref) < 2e-2
    assert abs(float(loss) - float(lref)) < 2e-2 * abs(float(lref))
    assert rel(a1.grad, a2.grad) < 1.5e-1
    assert max(errs.values()) < 1.5e-1
    for k in ("model.3.running_mean", "model.9.running_var"):
        assert rel(mod.state_dict()[k], sd[k]) < 2e-2, k
    assert int(mod.state_dict()["model.3.num_batches_tracked"]) == 1


def test_discriminator_frozen_and_detached_modes():
    """set_requires_grad(netD, False) during the G update (dgrad only) and fake_B.detach() during the D update (wgrad only)."""
    sd = O.make_state_dict(O.discriminator_keys(), 43, device=DEV, requires_grad=True)
    net = NW.define_D(3, 1, 64, "n_layers", 3, "batch", True, "normal", gpu_ids=[0])
    mod = load_into(net, sd)
    a, _, _ = O.synthetic_batch(2, 64, 302, device=DEV)
    z = torch.zeros(2, 1, 1, 1, device=DEV)
    for p in mod.parameters():
        p.requires_grad_(False)
    a1 = a.clone().requires_grad_(True)
    NW.GANLoss(False)(net(a1, z), True).backward()
    assert a1.grad is not None and all(p.grad is None for p in mod.parameters())
    for p in mod.parameters():
        p.requires_grad_(True)
    NW.GANLoss(False)(net(a, z), False).backward()
    assert all(p.grad is not None for p in mod.parameters())
    with torch.no_grad():
        out = net(a, z)
    assert not out.requires_grad


@pytest.mark.parametrize("N,S", [(2, 64), (2, 224)])
def test_encoder_forward_and_input_gradient(N, S):
    sd = O.make_state_dict(O.encoder_keys(), 44, device=DEV)
    net = NW.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, gpu_ids=[0])
    mod = load_into(net, sd)
    for p in mod.parameters():
        p.requires_grad_(False)
    a, _, _ = O.synthetic_batch(N, S, 303, device=DEV)
    a1 = a.clone().requires_grad_(True)
    y = net(a1)
    gy = torch.tensor([1.0, -2.0], device=DEV).view(N, 1, 1, 1)
    (y * gy).sum().backward()
    a2 = a.clone().requires_grad_(True)
    yr = O.encoder_forward(sd, a2)
    (yr * gy).sum().backward()
    print("E N=%d S=%d y" % (N, S), y.flatten().tolist(), yr.flatten().tolist(), "dx %.3e" % rel(a1.grad, a2.grad))
    assert float((y - yr).abs().max()) < 3e-2 * float(yr.abs().max()) + 1e-3
    assert rel(a1.grad, a2.grad) < 1.5e-1
    for k in ("base.model.bn1.running_mean", "base.model.layer4.1.bn2.running_var", "cnn.1.running_mean"):
        assert rel(mod.state_dict()[k], sd[k]) < 2e-2, k


def test_upsample_and_scalar_losses():
    x = torch.randn(2, 3, 16, 16, device=DEV, requires_grad=True)
    y = NW.upsample2d(x, 28)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    x2 = x.detach().clone().requires_grad_(True)
    y2 = O.upsample2d(x2, 28)
    (y2 * w).sum().backward()
    assert rel(y, y2) < 1e-6 and rel(x.grad, x2.grad) < 1e-5
    a = torch.randn(4, 3, 8, 8, device=DEV, requires_grad=True)
    b = torch.randn(4, 3, 8, 8, device=DEV)
    for mine, ref in ((NW.l1_loss, torch.nn.functional.l1_loss), (NW.mse_loss, torch.nn.functional.mse_loss)):
        a.grad = None
        l = mine(a, b) * 0.5
        l.backward()
        g1 = a.grad.clone()
        a.grad = None
        lr = ref(a, b) * 0.5
        lr.backward()
        assert abs(float(l) - float(lr)) < 1e-5 * abs(float(lr)) and rel(g1, a.grad) < 1e-5
    p = torch.tensor([0.0, 1.0, 1e-30, 0.3, 0.9999999, 0.5, 0.2, 0.8], device=DEV).view(2, 1, 2, 2).requires_grad_(True)
    for tgt in (True, False, [1, 0]):
        p.grad = None
        l = NW.GANLoss(False)(p, tgt)
        l.backward()
        g1 = p.grad.clone()
        p.grad = None
        lr = O.gan_loss(p, tgt)
        lr.backward()
        assert abs(float(l) - float(lr)) < 1e-5 * abs(float(lr)) and rel(g1, p.grad) < 1e-5, tgt
    prob = torch.tensor([0.2, 0.5, 0.9, 0.0, 1.0], device=DEV).view(5, 1, 1, 1)
    label = torch.tensor([0, 1, 2, 2, 0], device=DEV)
    assert abs(float(NW.BinaryNLLLoss()(prob, label)) - float(O.elo_nll(prob, label))) < 1e-5
