"""Drop-in surface, checked without a GPU: the modules built by pcgan_b200.networks carry exactly the reference's
state_dict keys and shapes (the key lists of oracle/pcgan_oracle.py are pinned to the reference by tests/golden), the
factories accept the reference's call signatures, unsupported variants fail loudly, and nothing runs on the CPU."""
import inspect

import pytest
import torch

from oracle import pcgan_oracle as O
from pcgan_b200 import networks as NW
from pcgan_b200 import siamese as SI
from pcgan_b200.wsgan_emb_model import WSGANEmbModel, default_options


def _shapes(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


@pytest.mark.parametrize("nb", [6, 9])
def test_generator_state_dict_layout(nb):
    g = NW.define_G(3, 3, 1, 64, "resnet_%dblocks" % nb, "instance", init_type="normal", gpu_ids=[])
    assert _shapes(g) == {k: tuple(s) for k, s in O.generator_keys(n_blocks=nb).items()}
    assert list(g.state_dict().keys()) == list(O.generator_keys(n_blocks=nb).keys())


def test_discriminator_and_encoder_state_dict_layout():
    d = NW.define_D(3, 1, 64, "n_layers", 3, "batch", True, "normal", gpu_ids=[])
    assert _shapes(d) == {k: tuple(s) for k, s in O.discriminator_keys().items()}
    for noisy in (False, True):
        e = NW.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, gpu_ids=[],
                        noisy=noisy, bnn_dropout=0.2 if noisy else 0.0)
        assert list(e.state_dict().keys()) == list(O.encoder_keys(noisy=noisy).keys())
        assert _shapes(e) == {k: tuple(s) for k, s in O.encoder_keys(noisy=noisy).items()}
    base = NW.ResNetFeature(3, "resnet18", 0.0)
    s = NW.SiameseNetwork(base, pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, fc_dim=[], drop_layer=NW.get_dropout_layer(0.0))
    assert list(s.state_dict().keys()) == list(O.encoder_keys().keys())


def test_reference_checkpoint_round_trip():
    """A state_dict in the reference layout loads strictly, and load_pretrained drops cxn / fc keys (networks.py:1070-1078)."""
    sd = O.make_state_dict(O.encoder_keys(), 3)
    e = NW.define_E("resnet18", 3, init_type="normal", pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, gpu_ids=[])
    extra = dict(sd)
    extra["fc.1.weight"] = torch.zeros(1, 2, 1, 1)
    extra["cxn.0.weight"] = torch.zeros(1)
    e.load_pretrained(extra)
    for k, v in e.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_factory_signatures_match_the_reference():
    """Parameter names of the reference's factories (models/networks.py:106-107, 147-148, 232-233)."""
    assert list(inspect.signature(NW.define_G).parameters)[:10] == ["input_nc", "output_nc", "nz", "ngf", "which_model_netG", "norm", "nl", "dropout", "init_type", "gpu_ids"]
    assert list(inspect.signature(NW.define_D).parameters)[:10] == ["input_nc", "nz", "ndf", "which_model_netD", "n_layers_D", "norm", "use_sigmoid", "init_type", "num_Ds", "gpu_ids"]
    assert list(inspect.signature(NW.define_E).parameters) == ["which_model_netE", "input_nc", "init_type", "pooling", "cnn_dim", "cnn_pad", "cnn_relu_slope", "gpu_ids", "fine_size_E", "noisy", "bnn_dropout"]
    for name in ("initialize", "setup", "set_input", "forward", "test", "optimize_parameters", "backward_G", "backward_D", "backward_GE",
                 "backward_G_alone", "update_G", "update_D", "update_G_and_E", "get_current_losses", "get_current_visuals",
                 "save_networks", "load_networks", "update_learning_rate", "set_requires_grad"):
        assert callable(getattr(WSGANEmbModel, name)), name
    assert callable(SI.get_model) and callable(SI.EloTrainer.train_step)


def test_unsupported_variants_and_cpu_fail_loudly():
    with pytest.raises(NotImplementedError):
        NW.define_G(3, 3, 1, 64, "unet_128_input", "instance", gpu_ids=[])
    with pytest.raises(NotImplementedError):
        NW.define_D(3, 1, 64, "pixel", 3, "batch", True, gpu_ids=[])
    with pytest.raises(NotImplementedError):
        NW.define_E("alexnet", 3, gpu_ids=[])
    g = NW.define_G(3, 3, 1, 64, "resnet_6blocks", "instance", init_type="normal", gpu_ids=[])
    with pytest.raises(Exception) as ei:       # no CPU path: a CPU tensor is refused, never computed on
        g(torch.zeros(1, 3, 32, 32), torch.zeros(1, 1, 1, 1))
    assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            m = WSGANEmbModel()
            m.initialize(default_options(gpu_ids=[0]))


def test_fused_adam_is_a_torch_adam_and_has_no_cpu_path():
    """pcgan_b200.optim.FusedAdam keeps torch.optim.Adam's constructor / param_groups / schedulers (the reference builds
    its optimizers at models/wsgan_emb_model.py:153-163) and refuses what it does not implement instead of falling back."""
    import torch
    from pcgan_b200.optim import FusedAdam
    p = [torch.nn.Parameter(torch.zeros(4, 3))]
    o = FusedAdam(p, lr=2e-4, betas=(0.5, 0.999))
    assert isinstance(o, torch.optim.Adam) and o.param_groups[0]["lr"] == 2e-4 and o.param_groups[0]["betas"] == (0.5, 0.999)
    sched = torch.optim.lr_scheduler.LambdaLR(o, lr_lambda=lambda e: 0.5)
    assert o.param_groups[0]["lr"] == pytest.approx(1e-4)
    del sched
    with pytest.raises(NotImplementedError):
        FusedAdam(p, weight_decay=1e-2)
    with pytest.raises(NotImplementedError):
        FusedAdam(p, amsgrad=True)
    p[0].grad = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        o.step()
    assert o.state_dict()["state"] == {}          # nothing was allocated or stepped on the CPU

