"""Pins oracle/pcgan_oracle.py to the reference: replays the golden fixtures produced by
tests/golden/make_golden.py (which ran the unmodified phymhan/pc-gan modules and its
WSGANEmbModel.optimize_parameters on the CPU) and requires fp32 agreement."""
import os

import pytest
import torch

from oracle import pcgan_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(G, name), weights_only=False)


def close(a, b, tol=2e-5):
    a, b = torch.as_tensor(a, dtype=torch.float32), torch.as_tensor(b, dtype=torch.float32)
    return float((a - b).norm()) <= tol * float(b.norm()) + 1e-7


def test_generator_small_forward_backward_and_running_stats():
    fx = load("generator_small.pt")
    sd = O.make_state_dict(O.generator_keys(ngf=fx["ngf"]), fx["seed"], requires_grad=True)
    a, _, _ = O.synthetic_batch(2, 32, fx["x_seed"])
    a.requires_grad_(True)
    out = O.generator_forward(sd, a, fx["z"])
    assert close(out, fx["out"])
    w = torch.linspace(-1, 1, out.numel()).view_as(out)
    (out * w).sum().backward()
    assert close(a.grad, fx["dx"], 1e-4)
    for k, g in fx["grads"].items():
        assert close(sd[k].grad, g, 1e-4), k
    for k, v in fx["running"].items():
        assert close(sd[k], v), k


def test_generator_full_size():
    fx = load("generator_full.pt")
    sd = O.make_state_dict(O.generator_keys(), fx["seed"])
    a, _, _ = O.synthetic_batch(1, 128, fx["x_seed"])
    with torch.no_grad():
        out = O.generator_forward(sd, a, torch.tensor([fx["z"]]).view(1, 1, 1, 1))
    assert close(out[..., ::8, ::8], fx["out_sub"], 1e-4)
    assert abs(float(out.std()) - float(fx["out_std"])) < 1e-5


def test_discriminator_small_with_ganloss():
    fx = load("discriminator_small.pt")
    sd = O.make_state_dict(O.discriminator_keys(ndf=fx["ndf"]), fx["seed"], requires_grad=True)
    a, _, _ = O.synthetic_batch(3, 32, fx["x_seed"])
    a.requires_grad_(True)
    out = O.discriminator_forward(sd, a, fx["z"])
    assert close(out, fx["out"])
    loss = O.gan_loss(out, [1, 0, 1])
    assert close(loss, fx["loss"])
    loss.backward()
    assert close(a.grad, fx["dx"], 1e-4)
    for k, g in fx["grads"].items():
        assert close(sd[k].grad, g, 1e-4), k
    for k, v in fx["running"].items():
        assert close(sd[k].float(), v.float()), k


def test_encoder_keys_forward_backward():
    fx = load("encoder.pt")
    keys = O.encoder_keys()
    assert list(keys.keys()) == fx["keys"]
    sd = O.make_state_dict(keys, fx["seed"])
    a, _, _ = O.synthetic_batch(2, 64, fx["x_seed"])
    a.requires_grad_(True)
    y = O.encoder_forward(sd, a)
    assert close(y, fx["y"], 1e-4)
    (y * torch.tensor([1.0, -2.0]).view(2, 1, 1, 1)).sum().backward()
    assert close(a.grad[..., ::4, ::4], fx["dx_sub"], 2e-4)
    for k, v in fx["running"].items():
        assert close(sd[k], v, 1e-4), k
    fx = load("encoder_noisy.pt")
    keys = O.encoder_keys(noisy=True)
    assert list(keys.keys()) == fx["keys"]
    sd = O.make_state_dict(keys, fx["seed"])
    with torch.no_grad():
        y, lv = O.encoder_forward(sd, a.detach(), noisy=True)
    assert close(y, fx["y"], 1e-4) and close(lv, fx["logvar"], 1e-4)


def test_losses_and_upsample():
    fx = load("losses.pt")
    for name, tgt in (("true", True), ("false", False), ("mixed", [1, 0])):
        p = fx["p"].clone().requires_grad_(True)
        l = O.gan_loss(p, tgt)
        l.backward()
        assert close(l, fx["bce_" + name][0]) and close(p.grad, fx["bce_" + name][1]), name
    p2 = fx["p2"].clone().requires_grad_(True)
    l = O.gan_loss(p2, [1, 0], use_lsgan=True)
    l.backward()
    assert close(l, fx["mse_mixed"][0]) and close(p2.grad, fx["mse_mixed"][1])
    assert close(O.elo_nll(fx["elo_prob"], fx["elo_label"]), fx["elo_loss"])
    assert close(O.upsample2d(fx["up_in"], 9), fx["up_out"])


def test_two_training_steps_match_reference():
    """The whole step (forward, backward_G + Adam, backward_D + Adam) against the reference's
    optimize_parameters: all nine losses of two consecutive steps, and updated weights after step 1."""
    fx = load("step.pt")
    torch.set_num_threads(8)
    sg, sd_, se = fx["seeds"]
    m = O.WSGANEmbOracle(O.make_state_dict(O.generator_keys(), sg, requires_grad=True),
                         O.make_state_dict(O.discriminator_keys(), sd_, requires_grad=True),
                         O.make_state_dict(O.encoder_keys(), se))
    for it, want in enumerate(fx["steps"]):
        a, b, label = O.synthetic_batch(2, 128, fx["batch_seeds"][it])
        got = m.optimize_parameters(a, b, label)
        for k in ("G_GAN", "G_cycle", "z_rec", "D_real_right", "D_real_wrong", "D_fake"):
            assert abs(got[k] - want[k]) <= 2e-4 * abs(want[k]) + 1e-7, (it, k, got[k], want[k])
        if it == 0:
            assert close(m.fake_b.detach()[..., ::8, ::8], fx["extra"]["fake_b_sub"], 1e-4)
            assert close(m.y_b, fx["extra"]["y_b"], 1e-4)
            assert close(m.g["model.10.conv_block.1.weight"][:4, :4], fx["extra"]["g_w_after"], 1e-4)
            assert close(m.d["model.2.weight"][:4, :4], fx["extra"]["d_w_after"], 1e-4)


def test_encoder_with_dropout_and_noisy_head():
    """Live nn.Dropout2d (resnet.py:58-65, networks.py:1022) + cnn_logvar twin head: the reference's recorded masks
    replayed through the oracle."""
    fx = load("encoder_dropout.pt")
    sd = O.make_state_dict(O.encoder_keys(noisy=True), fx["seed"])
    a, _, _ = O.synthetic_batch(2, 64, fx["x_seed"])
    a.requires_grad_(True)
    masks = [m.clone() for m in fx["masks"]]

    def drop(t):
        return t * masks.pop(0).view(t.size(0), t.size(1), 1, 1)

    y, lv = O.encoder_forward(sd, a, cnn_relu_slope=0.7, noisy=True, drop=drop)
    assert not masks, "every recorded mask must be consumed"
    assert close(y, fx["y"], 1e-4) and close(lv, fx["logvar"], 1e-4)
    (y * torch.tensor([1.0, -2.0]).view(2, 1, 1, 1) + lv * torch.tensor([0.5, 0.25]).view(2, 1, 1, 1)).sum().backward()
    assert close(a.grad[..., ::4, ::4], fx["dx_sub"], 1e-3) and close(a.grad.norm(), fx["dx_norm"], 1e-3)


def test_siamese_elo_training_steps_match_reference():
    """siamese.py:590-686 (plain branch): two Adam steps of SiameseNetwork + BinaryNLLLoss."""
    fx = load("siamese_step.pt")
    sd = O.make_state_dict(O.encoder_keys(), fx["seed"], requires_grad=True)
    assert list(sd.keys()) == fx["keys"]
    m = O.SiameseOracle(sd, lr=2e-4, cnn_relu_slope=0.7)
    for it, want in enumerate(fx["steps"]):
        a, b, label = O.synthetic_batch(4, 64, fx["batch_seeds"][it])
        loss, prob = m.step(a, b, label)
        assert abs(loss - want["loss"]) <= 1e-4 * abs(want["loss"]), (it, loss, want["loss"])
        assert close(prob, want["prob"], 1e-3)
        if it == 0:   # .grad still holds the first step's gradients (Adam does not clear them)
            for k, g in fx["grads"].items():
                mine = sd[k].grad[:6] if sd[k].grad.dim() == 4 else sd[k].grad
                assert close(mine, g, 2e-3), k
    assert close(sd["base.model.layer1.0.conv1.weight"][:4, :4], fx["w_after"], 1e-3)


def test_bayesian_noisy_step_matches_reference():
    """BASELINE config 4 (--bayesian true --noisy true --noisy_var_type ae --bnn_dropout 0.2) at 64 x 64, T = 2: every
    Dropout2d mask and resample draw of the reference replayed through the oracle."""
    fx = load("step_bayesian.pt")
    sg, sdd, se = fx["seeds"]
    m = O.WSGANEmbOracle(O.make_state_dict(O.generator_keys(n_blocks=6), sg, requires_grad=True),
                         O.make_state_dict(O.discriminator_keys(), sdd, requires_grad=True),
                         O.make_state_dict(O.encoder_keys(noisy=True), se), n_blocks=6, fine_size_e=64, bayesian=True, noisy=True,
                         noisy_var_type="ae", bnn_T=2, dropout=True, drop_masks=[x.clone() for x in fx["masks"]],
                         eps_queue=[x.clone() for x in fx["eps"]])
    a, b, label = O.synthetic_batch(2, 64, fx["batch_seed"])
    got = m.optimize_parameters(a, b, label)
    assert not m.drop_masks and not m.eps_queue
    for k, v in fx["losses"].items():
        if k in got:
            assert abs(got[k] - v) <= 2e-3 * abs(v) + 1e-5, (k, got[k], v)
    assert close(m.y_b, fx["y_b"], 1e-3)
    assert close(m.fake_b.detach()[..., ::4, ::4], fx["fake_b_sub"], 1e-3)
    assert close(m.g["model.10.conv_block.1.weight"][:4, :4], fx["g_w_after"], 1e-3)


def test_variant_flags_step_matches_reference():
    """The non-default branches of the step — --lambda_A_GAN, --lambda_L1, --detach_fake_B, --use_real_A
    (models/wsgan_emb_model.py:256-259, 309-322, 340-347, 380-388) — against the reference's own optimize_parameters."""
    fx = load("step_variants.pt")
    torch.set_num_threads(8)
    sg, sd_, se = fx["seeds"]
    fl = fx["flags"]
    m = O.WSGANEmbOracle(O.make_state_dict(O.generator_keys(n_blocks=6), sg, requires_grad=True),
                         O.make_state_dict(O.discriminator_keys(), sd_, requires_grad=True),
                         O.make_state_dict(O.encoder_keys(), se), n_blocks=6, fine_size_e=64,
                         lambda_a_gan=fl["lambda_A_GAN"], lambda_l1=fl["lambda_L1"], detach_fake_b=fl["detach_fake_B"],
                         use_real_a=fl["use_real_A"])
    a, b, label = O.synthetic_batch(2, 64, fx["batch_seed"])
    got = m.optimize_parameters(a, b, label)
    for k in ("G_GAN", "G_GAN_cycle", "G_L1", "G_cycle", "z_rec", "D_real_right", "D_real_wrong", "D_fake"):
        assert abs(got[k] - fx["losses"][k]) <= 2e-4 * abs(fx["losses"][k]) + 1e-7, (k, got[k], fx["losses"][k])
    assert close(m.fake_b.detach()[..., ::4, ::4], fx["fake_b_sub"], 1e-4)
    assert close(m.g["model.10.conv_block.1.weight"][:4, :4], fx["g_w_after"], 1e-4)
    assert close(m.g["model.1.weight"][:4], fx["g_stem_w_after"], 1e-4)
    assert close(m.d["model.2.weight"][:4, :4], fx["d_w_after"], 1e-4)



def test_identity_preserving_step_matches_reference():
    """AlexNetFeature (models/networks.py:1218-1255) and the step with --lambda_IP 1 (:130-135, 353-356, 393-396) against
    the reference's own modules (tests/golden/step_ip.pt)."""
    fx = load("step_ip.pt")
    torch.set_num_threads(8)
    sg, sd_, se, sip = fx["seeds"]
    assert list(O.alexnet_keys().keys()) == fx["keys"]
    ip = O.make_state_dict(O.alexnet_keys(), sip)
    x, _, _ = O.synthetic_batch(2, 224, fx["feat_seed"])
    with torch.no_grad():
        feat = O.alexnet_forward(ip, x)
    assert close(feat[:, ::16], fx["feat_sub"], 1e-4) and abs(float(feat.mean()) - float(fx["feat_mean"])) < 1e-5
    m = O.WSGANEmbOracle(O.make_state_dict(O.generator_keys(n_blocks=6), sg, requires_grad=True),
                         O.make_state_dict(O.discriminator_keys(), sd_, requires_grad=True),
                         O.make_state_dict(O.encoder_keys(), se), n_blocks=6, fine_size_e=64, sd_ip=ip, lambda_ip=1.0, fine_size_ip=224)
    a, b, label = O.synthetic_batch(2, 64, fx["batch_seed"])
    got = m.optimize_parameters(a, b, label)
    for k in ("G_GAN", "G_IP", "G_cycle", "z_rec", "D_real_right", "D_real_wrong", "D_fake"):
        assert abs(got[k] - fx["losses"][k]) <= 2e-4 * abs(fx["losses"][k]) + 1e-7, (k, got[k], fx["losses"][k])
    assert close(m.fake_b.detach()[..., ::4, ::4], fx["fake_b_sub"], 1e-4)
    assert close(m.g["model.10.conv_block.1.weight"][:4, :4], fx["g_w_after"], 1e-4)


def test_unet_generator_matches_reference():
    """UnetGenerator (unet_128, InstanceNorm with running statistics; models/networks.py:659-733), incl. the in-place
    LeakyReLU that the skip connections carry, against the reference's module (tests/golden/unet.pt)."""
    fx = load("unet.pt")
    torch.set_num_threads(8)
    keys = O.unet_keys()
    assert list(keys.keys()) == fx["keys"]
    sd = O.make_state_dict(keys, fx["seed"], requires_grad=True)
    x, _, _ = O.synthetic_batch(2, 128, fx["x_seed"])
    x.requires_grad_(True)
    out = O.unet_forward(sd, x, fx["z"])
    assert close(out.detach()[..., ::8, ::8], fx["out_sub"], 1e-4) and abs(float(out.mean()) - float(fx["out_mean"])) < 1e-5
    w = torch.linspace(-1, 1, out.numel()).view_as(out)
    (out * w).sum().backward()
    assert close(x.grad[..., ::8, ::8], fx["dx_sub"], 2e-3)
    assert close(sd["model.model.0.weight"].grad[:4], fx["g_down0"], 2e-3)
    assert close(sd["model.model.3.weight"].grad[:4], fx["g_up0"], 2e-3)
    inner = "model.model.1.model.3.model.3.model.3.model.3.model.3.model.1"
    assert close(sd[inner + ".weight"].grad[:2, :2], fx["g_inner_down"], 2e-3)
    assert close(sd[inner + ".bias"].grad[:8], fx["g_inner_bias"], 2e-3)
    assert close(sd["model.model.1.model.2.running_mean"][:8], fx["rm"], 1e-4)
