"""The whole wsgan_emb step on the GPU against the oracle restatement of WSGANEmbModel.optimize_parameters
(models/wsgan_emb_model.py:478-484), same weights, same synthetic batches.

  * one step, teacher-forced (same weights and Adam state in): the nine losses
  * a 200-step run: D / G loss trajectories compared as run means and 20-step moving averages (gate 2 %, BASELINE.json);
    per-step equality is not meaningful — the reference cannot reproduce its own per-step trajectory under a change
    of summation order (> 2 % apart from step 6) — so the oracle's own fp32-vs-TF32 divergence is printed
    beside ours as the noise floor.
"""
import os

import pytest
import torch

from oracle import pcgan_oracle as O
from pcgan_b200.wsgan_emb_model import WSGANEmbModel, default_options

pytestmark = pytest.mark.gpu
DEV = "cuda"
KEYS = ("G_GAN", "G_cycle", "z_rec", "D_real_right", "D_real_wrong", "D_fake")


def build_pair(B, seeds=(31, 32, 33)):
    sds = [O.make_state_dict(k, s, device=DEV, requires_grad=rg) for k, s, rg in
           ((O.generator_keys(), seeds[0], True), (O.discriminator_keys(), seeds[1], True), (O.encoder_keys(), seeds[2], False))]
    model = WSGANEmbModel()
    opt = default_options(batchSize=B, gpu_ids=[0])
    model.initialize(opt)
    model.setup(opt)
    for net, sd in zip((model.netG, model.netD, model.netE), sds):
        net.module.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
    oracle = O.WSGANEmbOracle(*sds)
    return model, oracle


def test_single_step_losses():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 8
    model, oracle = build_pair(B)
    a, b, label = O.synthetic_batch(B, 128, 500, device=DEV)
    model.set_input({"A": a, "B": b, "label": label})
    model.optimize_parameters()
    got = model.get_current_losses()
    want = oracle.optimize_parameters(a, b, label)
    print("step losses:", {k: "%.5f/%.5f" % (got[k], want[k]) for k in KEYS})
    for k in KEYS:
        if k == "z_rec":
            continue
        assert abs(got[k] - want[k]) <= 0.03 * abs(want[k]) + 1e-5, (k, got[k], want[k])
    # z_rec = MSE of two ~0.05-sized outputs of a random-init 20-layer encoder that differ by ~0.02: against the oracle its
    # relative error is the encoder's end-to-end bf16 error (3e-2 of |y|) amplified by that cancellation, which says
    # nothing about wiring.  Checked instead: (a) the loss arithmetic, teacher-forced on the model's own encoder outputs
    # (exact), (b) every encoder stage against fp32 autograd in tests/test_chain_gpu.py, (c) here only that both are small.
    lz = float(((model.pred_y - model.y_B) ** 2).mean()) * model.opt.lambda_z
    assert abs(got["z_rec"] - lz) <= 1e-5 * abs(lz) + 1e-9, (got["z_rec"], lz)
    assert 0.0 <= got["z_rec"] < 10 * want["z_rec"] + 1e-3
    # after the step both generators moved: compare one updated weight (Adam normalises, so direction matters more than size)
    wg = model.netG.module.model[26].weight.detach()
    wr = oracle.g["model.26.weight"].detach()
    assert float((wg - wr).abs().max()) < 4.1e-4   # at most two Adam steps of lr 2e-4 apart


def test_variant_flags_step_losses():
    """The non-default branches of the step (--lambda_A_GAN, --lambda_L1, --detach_fake_B, --use_real_A:
    models/wsgan_emb_model.py:256-259, 309-322, 340-347, 380-388) against the oracle, whose same branches are pinned to the
    reference by tests/golden/step_variants.pt."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 4
    flags = dict(lambda_A_GAN=0.5, lambda_L1=0.3, detach_fake_B=True, use_real_A=True)
    sds = [O.make_state_dict(k, s, device=DEV, requires_grad=rg) for k, s, rg in
           ((O.generator_keys(), 51, True), (O.discriminator_keys(), 52, True), (O.encoder_keys(), 53, False))]
    model = WSGANEmbModel()
    opt = default_options(batchSize=B, gpu_ids=[0], **flags)
    model.initialize(opt)
    model.setup(opt)
    for net, sd in zip((model.netG, model.netD, model.netE), sds):
        net.module.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
    oracle = O.WSGANEmbOracle(*sds, lambda_a_gan=0.5, lambda_l1=0.3, detach_fake_b=True, use_real_a=True)
    a, b, label = O.synthetic_batch(B, 128, 600, device=DEV)
    model.set_input({"A": a, "B": b, "label": label})
    model.optimize_parameters()
    got = model.get_current_losses()
    want = oracle.optimize_parameters(a, b, label)
    keys = KEYS + ("G_GAN_cycle", "G_L1")
    print("variant step losses:", {k: "%.5f/%.5f" % (got[k], want[k]) for k in keys})
    for k in keys:
        if k == "z_rec":      # see test_single_step_losses
            lz = float(((model.pred_y - model.y_B) ** 2).mean()) * model.opt.lambda_z
            assert abs(got[k] - lz) <= 1e-5 * abs(lz) + 1e-9 and 0.0 <= got[k] < 10 * want[k] + 1e-3, (got[k], lz, want[k])
            continue
        tol = 0.05 if k == "G_GAN_cycle" else 0.03
        assert abs(got[k] - want[k]) <= tol * abs(want[k]) + 1e-5, (k, got[k], want[k])
    # detach_fake_B: the cycle terms reach the generator only through its second pass, yet every layer still moved
    wg = model.netG.module.model[1].weight.detach()
    wr = oracle.g["model.1.weight"].detach()
    assert float((wg - wr).abs().max()) < 4.1e-4


def test_identity_preserving_step_losses():
    """--lambda_IP 1 (the reference's default: models/wsgan_emb_model.py:130-135, 353-356, 393-396): the AlexNet feature
    loss between fake_B and real_A, forward and gradient through the frozen feature extractor into the generator."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 4
    sds = [O.make_state_dict(k, s, device=DEV, requires_grad=rg) for k, s, rg in
           ((O.generator_keys(), 61, True), (O.discriminator_keys(), 62, True), (O.encoder_keys(), 63, False))]
    ip = O.make_state_dict(O.alexnet_keys(), 64, device=DEV)
    model = WSGANEmbModel()
    opt = default_options(batchSize=B, gpu_ids=[0], lambda_IP=1.0)
    model.initialize(opt)
    model.setup(opt)
    for net, sd in zip((model.netG, model.netD, model.netE, model.netIP), sds + [ip]):
        net.module.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
    oracle = O.WSGANEmbOracle(*sds, sd_ip=ip, lambda_ip=1.0, fine_size_ip=224)
    a, b, label = O.synthetic_batch(B, 128, 710, device=DEV)
    model.set_input({"A": a, "B": b, "label": label})
    model.optimize_parameters()
    got = model.get_current_losses()
    want = oracle.optimize_parameters(a, b, label)
    print("IP step losses:", {k: "%.6f/%.6f" % (got[k], want[k]) for k in KEYS + ("G_IP",)})
    for k in ("G_GAN", "G_cycle", "D_real_right", "D_real_wrong", "D_fake", "G_IP"):
        assert abs(got[k] - want[k]) <= 0.03 * abs(want[k]) + 1e-6, (k, got[k], want[k])
    wg = model.netG.module.model[26].weight.detach()
    assert float((wg - oracle.g["model.26.weight"].detach()).abs().max()) < 4.1e-4
    assert all(p.grad is None for p in model.netIP.parameters())      # frozen: no weight gradients are computed


@pytest.mark.skipif(os.environ.get("PCGAN_SKIP_TRAJ") == "1", reason="trajectory test disabled")
def test_loss_trajectories_200_steps():
    steps, B = 200, int(os.environ.get("PCGAN_TRAJ_BATCH", "64"))      # BASELINE configs[2]: 64 pairs per GPU
    model, oracle = build_pair(B)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    _, oracle_tf32 = None, None
    sds = [O.make_state_dict(k, s, device=DEV, requires_grad=rg) for k, s, rg in
           ((O.generator_keys(), 31, True), (O.discriminator_keys(), 32, True), (O.encoder_keys(), 33, False))]
    oracle_tf32 = O.WSGANEmbOracle(*sds)
    hist = {"mine": [], "oracle": [], "oracle_tf32": []}
    for it in range(steps):
        a, b, label = O.synthetic_batch(B, 128, 1000 + it % 50, device=DEV)
        model.set_input({"A": a, "B": b, "label": label})
        model.optimize_parameters()
        hist["mine"].append(model.get_current_losses())
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        hist["oracle"].append(oracle.optimize_parameters(a, b, label))
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        hist["oracle_tf32"].append(oracle_tf32.optimize_parameters(a, b, label))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    def series(h, which):
        if which == "G":
            return torch.tensor([x["G_GAN"] + x["G_cycle"] + x["z_rec"] for x in h])
        return torch.tensor([(x["D_fake"] + (x["D_real_right"] + x["D_real_wrong"]) * 0.5) * 0.5 for x in h])

    def smooth(t, w=20):
        return t.unfold(0, w, 1).mean(1)

    report = {}
    for which in ("G", "D"):
        ref = series(hist["oracle"], which)
        for name in ("mine", "oracle_tf32"):
            s = series(hist[name], which)
            mean_dev = abs(float(s.mean() - ref.mean())) / float(ref.mean())
            sm_dev = float((smooth(s) - smooth(ref)).abs().max()) / float(ref.mean())
            step_dev = float(((s - ref).abs() / ref).max())
            report[(which, name)] = (mean_dev, sm_dev, step_dev)
            print("loss_%s %-12s run-mean dev %.3f%%  20-step-smoothed max dev %.3f%%  per-step max dev %.1f%%  (mean %.4f vs %.4f)" %
                  (which, name, 100 * mean_dev, 100 * sm_dev, 100 * step_dev, float(s.mean()), float(ref.mean())))
    os.makedirs("gpurun_out", exist_ok=True)
    torch.save(hist, "gpurun_out/trajectories.pt")
    for which in ("G", "D"):
        mean_dev, sm_dev, _ = report[(which, "mine")]
        floor = report[(which, "oracle_tf32")]
        assert mean_dev < 0.02, "run-mean loss_%s deviates %.2f%% (noise floor %.2f%%)" % (which, 100 * mean_dev, 100 * floor[0])
        assert sm_dev < max(0.05, 2.5 * floor[1]), "smoothed loss_%s deviates %.2f%% (noise floor %.2f%%)" % (which, 100 * sm_dev, 100 * floor[1])


@pytest.mark.parametrize("segments", [False, True], ids=["one_graph", "three_segments"])
def test_cuda_graph_step_matches_eager(segments):
    """--cuda_graph: the captured-and-replayed step computes what the per-launch step computes (same weights, same
    batches; atomics make both runs non-bit-reproducible, the GAN dynamics amplify that, hence the loose gate), and
    the replay really updates the weights and the running statistics."""
    B, S = 4, 64
    models = []
    for graph in (False, True):
        sds = [O.make_state_dict(k, s, device=DEV, requires_grad=rg) for k, s, rg in
               ((O.generator_keys(), 31, True), (O.discriminator_keys(), 32, True), (O.encoder_keys(), 33, False))]
        m = WSGANEmbModel()
        opt = default_options(batchSize=B, gpu_ids=[0], fineSize=S, loadSize=S, cuda_graph=graph, cuda_graph_warmup=2, cuda_graph_segments=segments)
        m.initialize(opt)
        m.setup(opt)
        for net, sd in zip((m.netG, m.netD, m.netE), sds):
            net.module.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
        models.append(m)
    eager, graphed = models
    w_before = graphed.netG.module.model[26].weight.detach().clone()
    for it in range(6):   # graphed: 2 eager steps, capture + replay at step 2, replays after
        a, b, label = O.synthetic_batch(B, S, 700 + it, device=DEV)
        for m in models:
            m.set_input({"A": a, "B": b, "label": label})
            m.optimize_parameters()
        le, lg = eager.get_current_losses(), graphed.get_current_losses()
        print(it, {k: "%.5f/%.5f" % (le[k], lg[k]) for k in KEYS})
        # steps 0-1 are eager in both models and already differ by ~0.3 % (atomics); that difference grows ~6x per step
        # (SURVEY section 4), so the first replayed step (2) is the tight check and later ones only guard against garbage
        tol = 0.03 if it <= 2 else 0.15
        for k in ("G_GAN", "G_cycle", "D_real_right", "D_real_wrong", "D_fake"):
            assert abs(le[k] - lg[k]) <= tol * abs(le[k]) + 1e-4, (it, k, le[k], lg[k])
    assert len(graphed._graphs) == 1 and not eager._graphs
    wg, we = graphed.netG.module.model[26].weight.detach(), eager.netG.module.model[26].weight.detach()
    assert float((wg - w_before).abs().max()) > 1e-4, "replayed steps must move the weights"
    assert float((wg - we).abs().max()) < 2 * 6 * 2.1e-4
    bn_g, bn_e = graphed.netD.module.model[3], eager.netD.module.model[3]
    assert int(bn_g.num_batches_tracked) == int(bn_e.num_batches_tracked) == 6 * 4
    assert float((bn_g.running_mean - bn_e.running_mean).abs().max()) < 5e-2


def test_step_at_256_and_odd_batch():
    """BASELINE config 5 geometry (256 x 256, larger spatial tiles) and a ragged last batch (DataLoader has no drop_last,
    SURVEY A.11): the programs are re-planned per (batch, size) and the losses match the oracle."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sds = [O.make_state_dict(k, s, device=DEV, requires_grad=rg) for k, s, rg in
           ((O.generator_keys(), 31, True), (O.discriminator_keys(), 32, True), (O.encoder_keys(), 33, False))]
    model = WSGANEmbModel()
    opt = default_options(batchSize=2, gpu_ids=[0], fineSize=256, loadSize=256)
    model.initialize(opt)
    model.setup(opt)
    for net, sd in zip((model.netG, model.netD, model.netE), sds):
        net.module.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
    oracle = O.WSGANEmbOracle(*sds)
    for B, S, seed in ((2, 256, 900), (3, 128, 901), (32, 256, 902)):      # the last one: BASELINE configs[4] as quoted
        a, b, label = O.synthetic_batch(B, S, seed, device=DEV)
        model.set_input({"A": a, "B": b, "label": label})
        model.optimize_parameters()
        got = model.get_current_losses()
        want = oracle.optimize_parameters(a, b, label)
        print("B=%d S=%d:" % (B, S), {k: "%.5f/%.5f" % (got[k], want[k]) for k in KEYS})
        assert tuple(model.fake_B.shape) == (B, 3, S, S)
        for k in KEYS:
            if k == "z_rec" and B > 3:
                assert 0.0 <= got[k] < 10 * want[k] + 1e-3
            elif k == "z_rec":    # ~1e-3 of loss_G: the squared difference of two ~0.05 outputs of a random-init bf16 encoder whose
                # BatchNorms see 2-3 samples; it moves by several 1e-3 from run to run (atomics order): bounded, not matched
                assert 0.0 <= got[k] < 2e-2, (B, S, k, got[k], want[k])
            else:
                assert abs(got[k] - want[k]) <= 0.04 * abs(want[k]) + 1e-5, (B, S, k, got[k], want[k])


def test_checkpoint_round_trip_and_lr_schedule(tmp_path):
    """save_networks / load_networks (base_model.py:96-136) and update_learning_rate (:72-76) around a captured step: the
    learning rate lives on the device in graph mode and the scheduler still drives it."""
    B, S = 2, 64
    opt = default_options(batchSize=B, gpu_ids=[0], fineSize=S, loadSize=S, cuda_graph=True, cuda_graph_warmup=1, niter=2, niter_decay=4,
                          checkpoints_dir=str(tmp_path), name="ckpt", which_model_netG="resnet_6blocks")
    model = WSGANEmbModel()
    model.initialize(opt)
    model.setup(opt)
    a, b, label = O.synthetic_batch(B, S, 950, device=DEV)
    for _ in range(3):      # eager, capture, replay
        model.set_input({"A": a, "B": b, "label": label})
        model.optimize_parameters()
    lr0 = float(model.optimizer_G.param_groups[0]["lr"])
    model.update_learning_rate()
    model.update_learning_rate()
    lr1 = float(model.optimizer_G.param_groups[0]["lr"])
    assert lr0 == pytest.approx(2e-4) and 0 < lr1 < lr0
    w0 = model.netG.module.model[1].weight.detach().clone()
    model.set_input({"A": a, "B": b, "label": label})
    model.optimize_parameters()      # replay with the decayed rate
    step = float((model.netG.module.model[1].weight.detach() - w0).abs().max())
    assert 0 < step < 2 * lr0, (step, lr1)      # Adam: |update| is of the order of the learning rate
    model.save_networks("latest")
    other = WSGANEmbModel()
    other.initialize(default_options(batchSize=B, gpu_ids=[0], fineSize=S, loadSize=S, checkpoints_dir=str(tmp_path), name="ckpt",
                                     which_model_netG="resnet_6blocks"))
    other.load_networks("latest")
    for n in ("G", "D", "E"):
        sa, sb = getattr(model, "net" + n).module.state_dict(), getattr(other, "net" + n).module.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert torch.equal(sa[k], sb[k]), (n, k)
    # the loaded model trains on (packed operands are rebuilt from the loaded master weights)
    other.set_input({"A": a, "B": b, "label": label})
    other.optimize_parameters()
    assert all(v == v for v in other.get_current_losses().values())


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
def test_unet_generator_step_losses(graph):
    """--which_model_netG unet_128 (models/networks.py:659-733) in the whole step: losses against the oracle
    step with the Unet restatement (pinned to the reference by tests/golden/unet.pt), eager and captured."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 4
    sds = [O.make_state_dict(k, s, device=DEV, requires_grad=rg) for k, s, rg in
           ((O.unet_keys(), 61, True), (O.discriminator_keys(), 32, True), (O.encoder_keys(), 33, False))]
    model = WSGANEmbModel()
    opt = default_options(batchSize=B, gpu_ids=[0], which_model_netG="unet_128", cuda_graph=graph, cuda_graph_warmup=1)
    model.initialize(opt)
    model.setup(opt)
    assert type(model.netG.module).__name__ == "UnetGenerator"
    for net, sd in zip((model.netG, model.netD, model.netE), sds):
        net.module.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
    oracle = O.WSGANEmbOracle(*sds, generator="unet")
    w0 = {k: v.detach().clone() for k, v in model.netG.module.state_dict().items() if k.endswith("weight") and v.dim() == 4}
    for it in range(3 if graph else 1):
        a, b, label = O.synthetic_batch(B, 128, 520 + it, device=DEV)
        model.set_input({"A": a, "B": b, "label": label})
        model.optimize_parameters()
        got = model.get_current_losses()
        want = oracle.optimize_parameters(a, b, label)
        print("unet step %d:" % it, {k: "%.5f/%.5f" % (got[k], want[k]) for k in KEYS})
        tol = 0.04 if it == 0 else 0.15       # later steps: GAN dynamics amplify the bf16 difference (see the graph test above)
        for k in KEYS:
            if k == "z_rec":
                assert 0.0 <= got[k] < 10 * want[k] + 2e-2
            else:
                assert abs(got[k] - want[k]) <= tol * abs(want[k]) + 1e-5, (it, k, got[k], want[k])
    if graph:
        assert len(model._graphs) == 1
    # every convolution of the Unet received a gradient and moved by about one Adam step per iteration (Adam's bias-corrected
    # m / sqrt(v) exceeds 1 when successive gradients agree in sign and shrink, so the bound is 3 lr per step)
    for k, v in model.netG.module.state_dict().items():
        if k in w0:
            d = float((v - w0[k]).abs().max())
            assert 1e-5 < d < (3 if graph else 1) * 6e-4, (k, d)
