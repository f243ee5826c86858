"""Generates the golden fixtures in tests/golden/ by running the UNMODIFIED reference
(/root/reference, phymhan/pc-gan) on the CPU of the authoring container.  The reference is
not available on the GPU box, so the fixtures (seeds + small result tensors) are committed
and replayed by tests/test_oracle_cpu.py through oracle/pcgan_oracle.py.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.pt
"""
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import pcgan_oracle as O  # noqa: E402  (only fill_state_dict_/synthetic_batch: shared deterministic inputs)

torch.set_num_threads(8)
torch.manual_seed(0)


def sub(t, step):
    return t[..., ::step, ::step].clone()


def golden_generator():
    from models import networks
    net = networks.define_G(3, 3, 1, 8, which_model_netG="resnet_9blocks", norm="instance", init_type="normal", gpu_ids=[])
    O.fill_state_dict_(net.state_dict(), 11)
    a, _, _ = O.synthetic_batch(2, 32, 101)
    z = torch.tensor([0.3, -1.2]).view(2, 1, 1, 1)
    a.requires_grad_(True)
    out = net(a, z)
    w = torch.linspace(-1, 1, out.numel()).view_as(out)
    (out * w).sum().backward()
    sd = net.state_dict()
    fx = {"seed": 11, "ngf": 8, "x_seed": 101, "z": z, "out": out.detach(), "dx": a.grad.clone(),
          "grads": {k: p.grad.clone() for k, p in net.named_parameters()
                    if k in ("model.1.weight", "model.7.weight", "model.10.conv_block.1.weight", "model.18.conv_block.5.weight",
                             "model.19.weight", "model.22.weight", "model.26.weight", "model.26.bias")},
          "running": {k: sd[k].clone() for k in ("model.2.running_mean", "model.2.running_var", "model.11.conv_block.6.running_var", "model.2.num_batches_tracked")}}
    torch.save(fx, os.path.join(HERE, "generator_small.pt"))
    # full-size network, batch 1, subsampled output
    net = networks.define_G(3, 3, 1, 64, which_model_netG="resnet_9blocks", norm="instance", init_type="normal", gpu_ids=[])
    O.fill_state_dict_(net.state_dict(), 21)
    a, _, _ = O.synthetic_batch(1, 128, 102)
    with torch.no_grad():
        out = net(a, torch.tensor([0.7]).view(1, 1, 1, 1))
    torch.save({"seed": 21, "x_seed": 102, "z": 0.7, "out_sub": sub(out, 8), "out_mean": out.mean(), "out_std": out.std()},
               os.path.join(HERE, "generator_full.pt"))


def golden_discriminator():
    from models import networks
    net = networks.define_D(3, 1, 8, "n_layers", 3, "batch", True, "normal", gpu_ids=[])
    O.fill_state_dict_(net.state_dict(), 12)
    a, _, _ = O.synthetic_batch(3, 32, 103)
    z = torch.tensor([0.5, -0.5, 1.5]).view(3, 1, 1, 1)
    a.requires_grad_(True)
    out = net(a, z)
    loss = networks.GANLoss(use_lsgan=False)(out, [1, 0, 1])
    loss.backward()
    sd = net.state_dict()
    fx = {"seed": 12, "ndf": 8, "x_seed": 103, "z": z, "out": out.detach(), "loss": loss.detach(), "dx": a.grad.clone(),
          "grads": {k: p.grad.clone() for k, p in net.named_parameters()},
          "running": {k: sd[k].clone() for k in sd if "running" in k or "tracked" in k}}
    torch.save(fx, os.path.join(HERE, "discriminator_small.pt"))


def golden_encoder():
    from models import networks
    net = networks.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7,
                            gpu_ids=[], noisy=False, bnn_dropout=0.0)
    O.fill_state_dict_(net.state_dict(), 13)
    a, _, _ = O.synthetic_batch(2, 64, 104)
    a.requires_grad_(True)
    y = net(a)
    (y * torch.tensor([1.0, -2.0]).view(2, 1, 1, 1)).sum().backward()
    sd = net.state_dict()
    fx = {"seed": 13, "x_seed": 104, "y": y.detach(), "dx_sub": sub(a.grad, 4), "dx_norm": a.grad.norm(),
          "running": {k: sd[k].clone() for k in ("base.model.bn1.running_mean", "base.model.layer4.1.bn2.running_var", "cnn.1.running_mean")},
          "keys": list(sd.keys())}
    torch.save(fx, os.path.join(HERE, "encoder.pt"))
    # noisy twin head
    net = networks.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7,
                            gpu_ids=[], noisy=True, bnn_dropout=0.0)
    O.fill_state_dict_(net.state_dict(), 14)
    with torch.no_grad():
        y, lv = net(a.detach())
    torch.save({"seed": 14, "x_seed": 104, "y": y, "logvar": lv, "keys": list(net.state_dict().keys())}, os.path.join(HERE, "encoder_noisy.pt"))


def golden_losses():
    from models import networks
    p = torch.tensor([0.0, 1.0, 1e-30, 0.3, 0.9999999, 0.5, 0.2, 0.8]).view(2, 1, 2, 2).requires_grad_(True)
    crit = networks.GANLoss(use_lsgan=False)
    out = {}
    for name, tgt in (("true", True), ("false", False), ("mixed", [1, 0])):
        p.grad = None
        l = crit(p, tgt)
        l.backward()
        out["bce_" + name] = (l.detach(), p.grad.clone())
    p2 = torch.tensor([0.1, 0.7, 0.4, 1.2]).view(2, 1, 1, 2).requires_grad_(True)
    l = networks.GANLoss(use_lsgan=True)(p2, [1, 0])
    l.backward()
    out["mse_mixed"] = (l.detach(), p2.grad.clone())
    out["p"], out["p2"] = p.detach(), p2.detach()
    # BinaryNLLLoss hard-codes .cuda() in __init__ (networks.py:477): construct it with .cuda() neutralised
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        nll = networks.BinaryNLLLoss()
    finally:
        torch.Tensor.cuda = orig
    prob = torch.tensor([0.2, 0.5, 0.9, 0.0, 1.0]).view(5, 1, 1, 1)
    label = torch.tensor([0, 1, 2, 2, 0])
    out["elo_prob"], out["elo_label"], out["elo_loss"] = prob, label, nll(prob, label)
    x = torch.linspace(-1, 1, 2 * 3 * 5 * 5).view(2, 3, 5, 5)
    from util.util import upsample2d
    out["up_in"], out["up_out"] = x, upsample2d(x, 9)
    torch.save(out, os.path.join(HERE, "losses.pt"))


def golden_step():
    """Two full WSGANEmbModel.optimize_parameters() steps (train.py:33-34) at the benchmark architecture,
    batch 2, 128 x 128, fineSize_E 224, lambda_IP 0."""
    from models import networks
    tmp = tempfile.mkdtemp()
    e0 = networks.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, gpu_ids=[])
    pth = os.path.join(tmp, "e.pth")
    torch.save(e0.state_dict(), pth)
    argv = sys.argv
    sys.argv = ["x", "--model", "wsgan_emb", "--gpu_ids", "-1", "--which_model_netG", "resnet_9blocks", "--n_layers_D", "3",
                "--batchSize", "2", "--lambda_IP", "0", "--pretrained_model_path_E", pth, "--sourcefile_A", pth, "--dataroot", tmp,
                "--embedding_bins", "[-2,-1,0,1,2]", "--checkpoints_dir", tmp, "--name", "golden"]
    try:
        from options.train_options import TrainOptions
        from models import create_model
        opt = TrainOptions().parse()
        model = create_model(opt)
        model.setup(opt)
    finally:
        sys.argv = argv
    O.fill_state_dict_(model.netG.state_dict(), 31)
    O.fill_state_dict_(model.netD.state_dict(), 32)
    O.fill_state_dict_(model.netE.state_dict(), 33)
    steps = []
    for it in range(2):
        a, b, label = O.synthetic_batch(2, 128, 200 + it)
        model.set_input({"A": a, "B": b, "label": label, "A_paths": ["a"] * 2, "B_paths": ["b"] * 2})
        model.optimize_parameters()
        steps.append({k: float(v) for k, v in model.get_current_losses().items()})
        if it == 0:
            extra = {"fake_b_sub": sub(model.fake_B.detach(), 8), "y_b": model.y_B.clone(),
                     "g_w_after": model.netG.state_dict()["model.10.conv_block.1.weight"][:4, :4].clone(),
                     "d_w_after": model.netD.state_dict()["model.2.weight"][:4, :4].clone()}
    torch.save({"seeds": (31, 32, 33), "batch_seeds": (200, 201), "steps": steps, "extra": extra}, os.path.join(HERE, "step.pt"))
    print(steps)


def _record_dropout(net, masks):
    """Forward hooks on every nn.Dropout2d: the mask of each call, [N, C], in call order."""
    def hook(mod, inp, out):
        x = inp[0]
        nz = x.abs().amax((2, 3)) > 0
        ratio = torch.where(nz, out.abs().amax((2, 3)) / x.abs().amax((2, 3)).clamp_min(1e-30), torch.full_like(nz, 1.0 / (1 - mod.p), dtype=x.dtype))
        masks.append(ratio.detach().clone())
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.register_forward_hook(hook)


def golden_encoder_dropout():
    """SiameseFeature with the noisy twin head and live nn.Dropout2d (bnn_dropout 0.2: resnet.py:58-65, networks.py:1022),
    masks recorded so the oracle / kernels replay the same draws."""
    from models import networks
    net = networks.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7,
                            gpu_ids=[], noisy=True, bnn_dropout=0.2)
    O.fill_state_dict_(net.state_dict(), 15)
    masks = []
    _record_dropout(net, masks)
    a, _, _ = O.synthetic_batch(2, 64, 105)
    a.requires_grad_(True)
    torch.manual_seed(5)
    y, lv = net(a)
    (y * torch.tensor([1.0, -2.0]).view(2, 1, 1, 1) + lv * torch.tensor([0.5, 0.25]).view(2, 1, 1, 1)).sum().backward()
    torch.save({"seed": 15, "x_seed": 105, "masks": masks, "y": y.detach(), "logvar": lv.detach(), "dx_sub": sub(a.grad, 4),
                "dx_norm": a.grad.norm()}, os.path.join(HERE, "encoder_dropout.pt"))
    print("encoder_dropout: %d masks" % len(masks))


def golden_siamese_step():
    """Two steps of the Elo rating trainer (siamese.py:590-686, plain branch): SiameseNetwork(resnet18, cnn_dim=[32, 1],
    fc_dim=[]), BinaryNLLLoss, Adam(lr 2e-4) over base + cnn (siamese.py:544-551)."""
    from models import networks
    base = networks.ResNetFeature(input_nc=3, which_model="resnet18", dropout=0.0)
    net = networks.SiameseNetwork(base, pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, fc_dim=[],
                                  drop_layer=networks.get_dropout_layer(0.0))
    O.fill_state_dict_(net.state_dict(), 16)
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        crit = networks.BinaryNLLLoss()
    finally:
        torch.Tensor.cuda = orig
    import itertools
    opt = torch.optim.Adam(itertools.chain(net.base.parameters(), net.cnn.parameters()), lr=2e-4)
    steps = []
    for it in range(2):
        a, b, label = O.synthetic_batch(4, 64, 300 + it)
        opt.zero_grad()
        f1, f2, score = net(a, b)
        prob = torch.sigmoid(score)
        loss = crit(prob, label)
        loss.backward()
        if it == 0:
            grads = {k: (p.grad[:6].clone() if p.grad.dim() == 4 else p.grad.clone()) for k, p in net.named_parameters()
                     if k in ("base.model.conv1.weight", "base.model.layer2.0.downsample.0.weight", "base.model.layer4.1.bn2.weight",
                              "cnn.0.weight", "cnn.1.bias", "cnn.4.weight", "cnn.4.bias")}
        opt.step()
        steps.append({"loss": float(loss), "prob": prob.detach().clone()})
    sd = net.state_dict()
    torch.save({"seed": 16, "batch_seeds": (300, 301), "steps": steps, "grads": grads,
                "w_after": sd["base.model.layer1.0.conv1.weight"][:4, :4].clone(), "keys": list(sd.keys())},
               os.path.join(HERE, "siamese_step.pt"))
    print("siamese:", [s_["loss"] for s_ in steps])


def golden_step_bayesian():
    """One WSGANEmbModel.optimize_parameters() in the Bayesian + noisy encoder mode (BASELINE config 4:
    --bayesian true --noisy true --noisy_var_type ae --bnn_dropout 0.2) at a small size (64 x 64, fineSize_E 64, T = 2,
    resnet_6blocks) with every random draw recorded: the Dropout2d masks and the util.resample normals."""
    from models import networks
    tmp = tempfile.mkdtemp()
    e0 = networks.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, gpu_ids=[],
                           noisy=True, bnn_dropout=0.2)
    pth = os.path.join(tmp, "e.pth")
    torch.save(e0.state_dict(), pth)
    argv = sys.argv
    sys.argv = ["x", "--model", "wsgan_emb", "--gpu_ids", "-1", "--which_model_netG", "resnet_6blocks", "--n_layers_D", "3",
                "--batchSize", "2", "--lambda_IP", "0", "--pretrained_model_path_E", pth, "--sourcefile_A", pth, "--dataroot", tmp,
                "--embedding_bins", "[-2,-1,0,1,2]", "--checkpoints_dir", tmp, "--name", "golden_b", "--fineSize", "64", "--loadSize", "64",
                "--fineSize_E", "64", "--bayesian", "true", "--noisy", "true", "--noisy_var_type", "ae", "--bnn_dropout", "0.2", "--bnn_T", "2"]
    try:
        from options.train_options import TrainOptions
        from models import create_model
        opt = TrainOptions().parse()
        model = create_model(opt)
        model.setup(opt)
    finally:
        sys.argv = argv
    O.fill_state_dict_(model.netG.state_dict(), 41)
    O.fill_state_dict_(model.netD.state_dict(), 42)
    O.fill_state_dict_(model.netE.state_dict(), 43)
    masks, eps = [], []
    _record_dropout(model.netE, masks)
    orig_randn_like = torch.randn_like

    def rec_randn_like(t, *a, **k):
        e = orig_randn_like(t, *a, **k)
        eps.append(e.detach().clone())
        return e
    torch.manual_seed(7)
    a, b, label = O.synthetic_batch(2, 64, 400)
    torch.randn_like = rec_randn_like
    try:
        model.set_input({"A": a, "B": b, "label": label, "A_paths": ["a"] * 2, "B_paths": ["b"] * 2})
        model.optimize_parameters()
    finally:
        torch.randn_like = orig_randn_like
    losses = {k: float(v) for k, v in model.get_current_losses().items()}
    torch.save({"seeds": (41, 42, 43), "batch_seed": 400, "masks": masks, "eps": eps, "losses": losses, "y_b": model.y_B.clone(),
                "fake_b_sub": sub(model.fake_B.detach(), 4),
                "g_w_after": model.netG.state_dict()["model.10.conv_block.1.weight"][:4, :4].clone()},
               os.path.join(HERE, "step_bayesian.pt"))
    print("bayesian step:", losses, len(masks), "masks", len(eps), "eps")


def golden_step_variants():
    """One WSGANEmbModel.optimize_parameters() with the non-default branches of the step switched on:
    --lambda_A_GAN 0.5 (:340-342, 380-382) --lambda_L1 0.3 (:343-347, 384-388) --detach_fake_B (:256-259) --use_real_A
    (:309-322), at a small size (64 x 64, fineSize_E 64, resnet_6blocks)."""
    from models import networks
    tmp = tempfile.mkdtemp()
    e0 = networks.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, gpu_ids=[])
    pth = os.path.join(tmp, "e.pth")
    torch.save(e0.state_dict(), pth)
    argv = sys.argv
    sys.argv = ["x", "--model", "wsgan_emb", "--gpu_ids", "-1", "--which_model_netG", "resnet_6blocks", "--n_layers_D", "3",
                "--batchSize", "2", "--lambda_IP", "0", "--pretrained_model_path_E", pth, "--sourcefile_A", pth, "--dataroot", tmp,
                "--embedding_bins", "[-2,-1,0,1,2]", "--checkpoints_dir", tmp, "--name", "golden_v", "--fineSize", "64", "--loadSize", "64",
                "--fineSize_E", "64", "--lambda_A_GAN", "0.5", "--lambda_L1", "0.3", "--detach_fake_B", "--use_real_A"]
    try:
        from options.train_options import TrainOptions
        from models import create_model
        opt = TrainOptions().parse()
        model = create_model(opt)
        model.setup(opt)
    finally:
        sys.argv = argv
    O.fill_state_dict_(model.netG.state_dict(), 51)
    O.fill_state_dict_(model.netD.state_dict(), 52)
    O.fill_state_dict_(model.netE.state_dict(), 53)
    a, b, label = O.synthetic_batch(2, 64, 600)
    model.set_input({"A": a, "B": b, "label": label, "A_paths": ["a"] * 2, "B_paths": ["b"] * 2})
    model.optimize_parameters()
    losses = {k: float(v) for k, v in model.get_current_losses().items()}
    torch.save({"seeds": (51, 52, 53), "batch_seed": 600, "losses": losses,
                "flags": {"lambda_A_GAN": 0.5, "lambda_L1": 0.3, "detach_fake_B": True, "use_real_A": True},
                "fake_b_sub": sub(model.fake_B.detach(), 4),
                "g_w_after": model.netG.state_dict()["model.10.conv_block.1.weight"][:4, :4].clone(),
                "g_stem_w_after": model.netG.state_dict()["model.1.weight"][:4].clone(),
                "d_w_after": model.netD.state_dict()["model.2.weight"][:4, :4].clone()},
               os.path.join(HERE, "step_variants.pt"))
    print("variant step:", losses)


def golden_step_ip():
    """The AlexNet feature extractor alone (networks.py:1218-1255) and one WSGANEmbModel.optimize_parameters() with the
    identity-preserving loss on (--lambda_IP 1, the reference default: :130-135, 353-356, 393-396), fineSize_IP 224, at a
    small image size (64 x 64, fineSize_E 64, resnet_6blocks)."""
    from models import networks
    ip0 = networks.define_IP("alexnet", 3, [])
    sd_ip = ip0.state_dict()
    O.fill_state_dict_(sd_ip, 64)
    x, _, _ = O.synthetic_batch(2, 224, 701)
    with torch.no_grad():
        feat = ip0(x)
    tmp = tempfile.mkdtemp()
    e0 = networks.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, gpu_ids=[])
    pth, pth_ip = os.path.join(tmp, "e.pth"), os.path.join(tmp, "ip.pth")
    torch.save(e0.state_dict(), pth)
    torch.save(sd_ip, pth_ip)
    argv = sys.argv
    sys.argv = ["x", "--model", "wsgan_emb", "--gpu_ids", "-1", "--which_model_netG", "resnet_6blocks", "--n_layers_D", "3",
                "--batchSize", "2", "--lambda_IP", "1.0", "--pretrained_model_path_IP", pth_ip, "--pretrained_model_path_E", pth,
                "--sourcefile_A", pth, "--dataroot", tmp, "--embedding_bins", "[-2,-1,0,1,2]", "--checkpoints_dir", tmp, "--name", "golden_ip",
                "--fineSize", "64", "--loadSize", "64", "--fineSize_E", "64"]
    try:
        from options.train_options import TrainOptions
        from models import create_model
        opt = TrainOptions().parse()
        model = create_model(opt)
        model.setup(opt)
    finally:
        sys.argv = argv
    O.fill_state_dict_(model.netG.state_dict(), 61)
    O.fill_state_dict_(model.netD.state_dict(), 62)
    O.fill_state_dict_(model.netE.state_dict(), 63)
    a, b, label = O.synthetic_batch(2, 64, 700)
    model.set_input({"A": a, "B": b, "label": label, "A_paths": ["a"] * 2, "B_paths": ["b"] * 2})
    model.optimize_parameters()
    losses = {k: float(v) for k, v in model.get_current_losses().items()}
    torch.save({"seeds": (61, 62, 63, 64), "batch_seed": 700, "feat_seed": 701, "feat_sub": feat[:, ::16].clone(), "feat_mean": feat.mean(),
                "keys": list(sd_ip.keys()), "losses": losses, "fake_b_sub": sub(model.fake_B.detach(), 4),
                "g_w_after": model.netG.state_dict()["model.10.conv_block.1.weight"][:4, :4].clone()},
               os.path.join(HERE, "step_ip.pt"))
    print("IP step:", losses)


def golden_unet():
    """UnetGenerator(unet_128, instance norm), the default G of wsgan_emb (networks.py:659-733): output and a few
    gradients on seeded weights and inputs."""
    from models import networks
    net = networks.define_G(3, 3, 1, 64, "unet_128", "instance", "relu", 0, "normal", [])
    sd = net.state_dict()
    O.fill_state_dict_(sd, 71)
    x, _, _ = O.synthetic_batch(2, 128, 801)
    x.requires_grad_(True)
    z = torch.tensor([0.3, -0.8]).view(2, 1, 1, 1)
    out = net(x, z)
    w = torch.linspace(-1, 1, out.numel()).view_as(out)
    (out * w).sum().backward()
    named = dict(net.named_parameters())
    torch.save({"seed": 71, "x_seed": 801, "z": z, "keys": list(sd.keys()), "out_sub": sub(out.detach(), 8), "out_mean": out.mean().detach(),
                "dx_sub": sub(x.grad, 8), "g_down0": named["model.model.0.weight"].grad[:4].clone(),
                "g_up0": named["model.model.3.weight"].grad[:4].clone(),
                "g_inner_down": named["model.model.1.model.3.model.3.model.3.model.3.model.3.model.1.weight"].grad[:2, :2].clone(),
                "g_inner_bias": named["model.model.1.model.3.model.3.model.3.model.3.model.3.model.1.bias"].grad[:8].clone(),
                "rm": sd["model.model.1.model.2.running_mean"][:8].clone()}, os.path.join(HERE, "unet.pt"))
    print("unet:", float(out.mean()), float(out.std()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "unet":
        golden_unet()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ip":
        golden_step_ip()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        golden_step_variants()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "new":     # only the fixtures added later (the others are unchanged)
        golden_encoder_dropout()
        golden_siamese_step()
        golden_step_bayesian()
        sys.exit(0)
    golden_losses()
    golden_generator()
    golden_discriminator()
    golden_encoder()
    golden_step()
    golden_encoder_dropout()
    golden_siamese_step()
    golden_step_bayesian()
    golden_step_variants()
    golden_step_ip()
    golden_unet()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
