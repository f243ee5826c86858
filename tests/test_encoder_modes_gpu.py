"""GPU parity of the encoder's training-time variants against the oracle (same weights, same injected randomness):
  * live nn.Dropout2d + the noisy twin head (networks.py:1008-1068, resnet.py:55-73) forward and input gradient
  * the Elo rating trainer's step: weight / BatchNorm / bias gradients of the whole ResNet-18 + head (siamese.py:590-686)
  * one wsgan_emb step in the Bayesian + noisy mode (BASELINE config 4, wsgan_emb_model.py:218-240, 408-430)
  * one wsgan_emb step with lr_E > 0 (update_G_and_E, :463-476)"""
import pytest
import torch

from oracle import pcgan_oracle as O
from pcgan_b200 import networks as NW
from pcgan_b200 import siamese as SI
from pcgan_b200.wsgan_emb_model import WSGANEmbModel, default_options

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    return float((a.detach().float() - b.detach().float()).norm() / (b.detach().float().norm() + 1e-20))


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _masks(N, p, noisy, calls, seed):
    """Dropout2d draws for `calls` encoder passes in module order: 16 block sites + one per head."""
    g = torch.Generator().manual_seed(seed)
    chans = []
    for c in (64, 128, 256, 512):
        chans += [c] * 4
    chans += [32] * (2 if noisy else 1)
    out = []
    for _ in range(calls):
        for c in chans:
            out.append((torch.rand(N, c, generator=g) >= p).float() / (1 - p))
    return out


def test_encoder_dropout_noisy_forward_and_input_gradient():
    N, S, p = 4, 64, 0.2
    sd = O.make_state_dict(O.encoder_keys(noisy=True), 51, device=DEV)
    net = NW.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, gpu_ids=[0],
                      noisy=True, bnn_dropout=p)
    net.module.load_state_dict({k: v.clone() for k, v in sd.items()})
    for q in net.parameters():
        q.requires_grad_(False)
    masks = _masks(N, p, True, 1, 3)
    a, _, _ = O.synthetic_batch(N, S, 600, device=DEV)
    x1 = a.clone().requires_grad_(True)
    net.module.dropout_masks = [m.clone() for m in masks]
    y, lv = net(x1)
    w1, w2 = torch.randn(N, 1, 1, 1, device=DEV), torch.randn(N, 1, 1, 1, device=DEV)
    (y * w1 + lv * w2).sum().backward()
    assert not net.module.dropout_masks
    mq = [m.to(DEV) for m in masks]
    x2 = a.clone().requires_grad_(True)
    yr, lvr = O.encoder_forward(sd, x2, cnn_relu_slope=0.7, noisy=True, drop=lambda t: t * mq.pop(0).view(t.size(0), t.size(1), 1, 1))
    (yr * w1 + lvr * w2).sum().backward()
    e = (rel(y, yr), rel(lv, lvr), rel(x1.grad, x2.grad))
    print("dropout+noisy encoder: y %.3e logvar %.3e dx %.3e" % e)
    # a 20-layer bf16 network with random-init weights: outputs are small differences of large terms (see test_step_gpu)
    assert e[0] < 8e-2 and e[1] < 8e-2 and e[2] < 4.5e-1
    # only the rating head's gradient requested: the unused log-variance head contributes nothing
    x3 = a.clone().requires_grad_(True)
    net.module.dropout_masks = [m.clone() for m in masks]
    y3, _ = net(x3)
    (y3 * w1).sum().backward()
    assert float(x3.grad.abs().max()) > 0


def test_siamese_training_step_gradients_and_update():
    N, S = 8, 64
    sd = O.make_state_dict(O.encoder_keys(), 52, device=DEV, requires_grad=True)
    net = SI.get_model(gpu_ids=[0])
    net.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
    trainer = SI.EloTrainer(net, lr=2e-4)
    oracle = O.SiameseOracle(sd, lr=2e-4, cnn_relu_slope=0.7)
    a, b, label = O.synthetic_batch(N, S, 610, device=DEV)
    loss, prob = trainer.train_step(a, b, label.to(DEV))
    lref, pref = oracle.step(a, b, label.to(DEV))
    print("siamese step: loss %.5f / %.5f  prob %.3e" % (float(loss), lref, rel(prob, pref)))
    assert abs(float(loss) - lref) < 2e-2 * abs(lref)
    named = dict(net.named_parameters())
    errs = {}
    for k in ("base.model.conv1.weight", "base.model.bn1.weight", "base.model.layer1.0.conv1.weight", "base.model.layer2.0.downsample.0.weight",
              "base.model.layer2.0.downsample.1.bias", "base.model.layer3.1.conv2.weight", "base.model.layer4.1.bn2.weight", "cnn.0.weight",
              "cnn.1.weight", "cnn.1.bias", "cnn.4.weight", "cnn.4.bias"):
        errs[k] = rel(named[k].grad, sd[k].grad)
    print({k: "%.2e" % v for k, v in errs.items()})
    # whole-network bf16 gradients through 20 random-init layers: ReLU-mask flips compound towards the input (sqrt law);
    # printed as a diagnostic.  The gate on these weight / BatchNorm / bias gradients is per stage with identical masks:
    # tests/test_chain_gpu.py::test_encoder_chain_teacher_forced (every convolution and BatchNorm of the network <= 6e-3)
    assert max(errs.values()) < 1.0, errs
    assert errs["cnn.4.bias"] < 1e-3 and errs["cnn.4.weight"] < 8e-2
    assert named["cnn.0.bias"].grad is not None and float(named["cnn.0.bias"].grad.abs().max()) == 0.0
    # Adam moved every trained tensor by about lr
    moved = float((named["base.model.layer1.0.conv1.weight"].detach() - O.make_state_dict(O.encoder_keys(), 52, device=DEV)["base.model.layer1.0.conv1.weight"]).abs().max())
    assert 1e-4 < moved < 3e-4
    # a second step runs on the updated weights (packed operands refreshed)
    loss2, _ = trainer.train_step(a, b, label.to(DEV))
    lref2, _ = oracle.step(a, b, label.to(DEV))
    assert abs(float(loss2) - lref2) < 3e-2 * abs(lref2)


def test_siamese_noisy_branch_matches_oracle():
    """siamese.py:654-659 (--noisy true, no rsample): score / sqrt(exp(logvar1) + exp(logvar2)), second Adam over cnn_logvar."""
    N, S = 8, 64
    sd = O.make_state_dict(O.encoder_keys(noisy=True), 54, device=DEV, requires_grad=True)
    net = SI.get_model(noisy=True, gpu_ids=[0])
    net.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
    trainer = SI.EloTrainer(net, lr=2e-4, lr_sigma=1e-4)
    a, b, label = O.synthetic_batch(N, S, 611, device=DEV)
    w0 = net.cnn_logvar[4].weight.detach().clone()
    loss, prob = trainer.train_step(a, b, label.to(DEV))
    y1, lv1 = O.encoder_forward(sd, a, cnn_relu_slope=0.7, noisy=True)
    y2, lv2 = O.encoder_forward(sd, b, cnn_relu_slope=0.7, noisy=True)
    pref = torch.sigmoid((y1 - y2) / (torch.sqrt(torch.exp(lv1) + torch.exp(lv2)) + 1e-20))
    lref = O.elo_nll(pref, label.to(DEV))
    lref.backward()
    print("noisy siamese step: loss %.5f / %.5f, prob %.3e" % (float(loss), float(lref), rel(prob, pref)))
    assert abs(float(loss) - float(lref)) < 2e-2 * abs(float(lref))
    named = dict(net.named_parameters())
    errs = {k: rel(named[k].grad, sd[k].grad) for k in ("cnn_logvar.4.bias", "cnn_logvar.4.weight", "cnn.4.bias")}
    print({k: "%.2e" % v for k, v in errs.items()})
    assert errs["cnn_logvar.4.bias"] < 2e-2 and errs["cnn.4.bias"] < 2e-2 and errs["cnn_logvar.4.weight"] < 1.5e-1
    moved = float((net.cnn_logvar[4].weight.detach() - w0).abs().max())
    assert 2e-5 < moved < 2.1e-4, moved        # the sigma optimizer stepped with lr_sigma = 1e-4


@pytest.mark.parametrize("mode", ["plain", "rsample_mc"])
def test_siamese_graph_replay_matches_eager(mode):
    """EloTrainer(cuda_graph=True): the captured iteration computes what the per-launch iteration computes."""
    N, S = 8, 64
    kw = dict(noisy=True, rsample=True) if mode == "rsample_mc" else {}
    tr = []
    for graph in (False, True):
        sd = O.make_state_dict(O.encoder_keys(noisy=bool(kw)), 55, device=DEV)
        net = SI.get_model(gpu_ids=[0], **kw)
        net.load_state_dict({k: v.clone() for k, v in sd.items()})
        tr.append(SI.EloTrainer(net, lr=2e-4, cuda_graph=graph, M=2, lb_or_mc="mc"))
    w0 = tr[1].net.base.model.conv1.weight.detach().clone()
    for it in range(5):
        a, b, label = O.synthetic_batch(N, S, 640 + it, device=DEV)
        torch.manual_seed(it)
        le, _ = tr[0].train_step(a, b, label.to(DEV))
        torch.manual_seed(it)
        lg, _ = tr[1].train_step(a, b, label.to(DEV))
        print(mode, it, float(le), float(lg))
        if mode == "plain":      # the reparameterisation draws of the captured graph come from the graph's own Philox offsets
            assert abs(float(le) - float(lg)) <= 0.03 * abs(float(le)) + 1e-4
        assert float(lg) == float(lg)
    assert tr[1]._graph is not None and tr[0]._graph is None
    assert float((tr[1].net.base.model.conv1.weight.detach() - w0).abs().max()) > 1e-4


def _models(B, S, seeds, n_blocks=6, fine_e=64, **flags):
    noisy = bool(flags.get("noisy", False))
    sds = [O.make_state_dict(k, s, device=DEV, requires_grad=rg) for k, s, rg in
           ((O.generator_keys(n_blocks=n_blocks), seeds[0], True), (O.discriminator_keys(), seeds[1], True),
            (O.encoder_keys(noisy=noisy), seeds[2], flags.get("lr_E", 0.0) > 0))]
    opt = default_options(batchSize=B, gpu_ids=[0], fineSize=S, loadSize=S, fineSize_E=fine_e, which_model_netG="resnet_%dblocks" % n_blocks, **flags)
    model = WSGANEmbModel()
    model.initialize(opt)
    model.setup(opt)
    for net, sd in zip((model.netG, model.netD, model.netE), sds):
        net.module.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
    return model, sds


@pytest.mark.parametrize("B,S,T,nb,fe", [(4, 64, 2, 6, 64), (8, 128, 10, 9, 224)], ids=["small_T2", "config4_T10_E224"])
def test_bayesian_noisy_step_matches_oracle(B, S, T, nb, fe):
    """BASELINE configs[3]: --bayesian true --noisy true --noisy_var_type ae --bnn_dropout 0.2; the second case runs it with
    the reference's T = 10 Monte-Carlo passes, the 9-block generator and the encoder at 224 x 224."""
    p = 0.2
    model, sds = _models(B, S, (61, 62, 63), n_blocks=nb, fine_e=fe, bayesian=True, noisy=True, noisy_var_type="ae", bnn_dropout=p, bnn_T=T)
    masks = _masks(B, p, True, 3 * T, 9)
    g = torch.Generator().manual_seed(10)
    eps = [torch.randn(B, 1, 1, 1, generator=g) for _ in range(2)]
    oracle = O.WSGANEmbOracle(*sds, n_blocks=nb, fine_size_e=fe, bayesian=True, noisy=True, noisy_var_type="ae", bnn_T=T, dropout=True,
                              drop_masks=[m.clone() for m in masks], eps_queue=[e.clone() for e in eps])
    model.netE.module.dropout_masks = [m.clone() for m in masks]
    NW.NOISE_QUEUE[:] = [e.clone() for e in eps]
    a, b, label = O.synthetic_batch(B, S, 620, device=DEV)
    model.set_input({"A": a, "B": b, "label": label})
    model.optimize_parameters()
    got = model.get_current_losses()
    want = oracle.optimize_parameters(a, b, label)
    assert not model.netE.module.dropout_masks and not NW.NOISE_QUEUE and not oracle.drop_masks
    print("bayesian step:", {k: "%.5f/%.5f" % (got[k], want[k]) for k in want})
    for k in ("G_GAN", "G_cycle", "D_real_right", "D_real_wrong", "D_fake"):
        assert abs(got[k] - want[k]) <= 0.03 * abs(want[k]) + 1e-5, (k, got[k], want[k])
    # the uncertainty-weighted z term divides by a Monte-Carlo variance over T = 2 passes of a bf16 encoder (a difference
    # of nearly equal numbers): against the oracle it is a diagnostic only
    print("z_rec %.5f / %.5f" % (got["z_rec"], want["z_rec"]))
    assert got["z_rec"] == got["z_rec"] and abs(got["z_rec"]) < 1e3


def test_lr_E_step_trains_the_encoder():
    """update_G_and_E (:463-476): gradients reach E through the embeddings fed to G and D (retain_graph + second
    backward through G's first pass); every network's weights move and the losses stay finite."""
    B, S = 4, 64
    model, sds = _models(B, S, (71, 72, 73), lr_E=1e-4)
    a, b, label = O.synthetic_batch(B, S, 630, device=DEV)
    w0 = {n: getattr(model, "net" + n).module.state_dict()[k].clone() for n, k in (("G", "model.1.weight"), ("D", "model.2.weight"), ("E", "base.model.layer3.0.conv1.weight"))}
    for _ in range(2):
        model.set_input({"A": a, "B": b, "label": label})
        model.optimize_parameters()
    L = model.get_current_losses()
    assert all(v == v and abs(v) < 1e3 for v in L.values()), L
    for n, k in (("G", "model.1.weight"), ("D", "model.2.weight"), ("E", "base.model.layer3.0.conv1.weight")):
        moved = float((getattr(model, "net" + n).module.state_dict()[k] - w0[n]).abs().max())
        assert moved > 1e-5, (n, moved)
    # gradient with respect to the embedding z.  Discriminator: a real gradient (zero-padded conv + LeakyReLU).
    B2 = 4
    sdd = sds[1]
    d = model.netD.module
    d.load_state_dict({k: v.detach().clone() for k, v in sdd.items()})
    x = a[:B2].clone()
    z1 = torch.randn(B2, 1, 1, 1, device=DEV, requires_grad=True)
    out = d(x, z1)
    wgt = torch.randn_like(out)
    (out * wgt).sum().backward()
    z2 = z1.detach().clone().requires_grad_(True)
    (O.discriminator_forward(sdd, x, z2) * wgt).sum().backward()
    print("D dz: %.3e" % rel(z1.grad, z2.grad), z1.grad.flatten().tolist(), z2.grad.flatten().tolist())
    assert rel(z1.grad, z2.grad) < 2.5e-1
    # Generator: the constant z plane only shifts the stem's pre-InstanceNorm output by a per-channel constant, which the
    # norm removes, so the true gradient is a cancellation to ~0 (1e-4 in fp32); ours is the sum of 4096 bf16-rounded
    # pixel gradients: checked against the gradient mass it cancels from
    sd = sds[0]
    g = model.netG.module
    g.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
    z1 = torch.randn(B2, 1, 1, 1, device=DEV, requires_grad=True)
    x1 = x.clone().requires_grad_(True)
    out = g(x1, z1)
    wgt = torch.randn_like(out)
    (out * wgt).sum().backward()
    z2 = z1.detach().clone().requires_grad_(True)
    (O.generator_forward(sd, x, z2, 6) * wgt).sum().backward()
    mass = float(x1.grad.abs().sum((1, 2, 3)).mean()) / 3
    print("G dz: ours %s oracle %s, per-channel |dx| mass %.1f" % (z1.grad.flatten().tolist(), z2.grad.flatten().tolist(), mass))
    assert float((z1.grad - z2.grad).abs().max()) < 2e-3 * mass
