"""Drop-in surface against the reference's own front end (authoring container only: /root/reference is absent on the
GPU box, where these tests skip): `--model wsgan_emb_b200` resolves through models.find_model_using_name, the parser
built by our modify_commandline_options yields the namespace the reference's class yields, and the model fails loudly
without a GPU instead of falling back."""
import os
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")

ARGV = ["train.py", "--dataroot", "/tmp/none", "--gpu_ids", "-1", "--which_model_netG", "resnet_9blocks", "--n_layers_D", "3",
        "--batchSize", "4", "--lambda_IP", "0", "--embedding_bins", "[-2,-1,0,1,2]", "--checkpoints_dir", "/tmp/pcgan_dropin_ckpt"]


@pytest.fixture()
def ref_front_end(monkeypatch):
    monkeypatch.syspath_prepend(REF)
    for m in [k for k in sys.modules if k.split(".")[0] in ("models", "options", "util", "data")]:
        monkeypatch.delitem(sys.modules, m)
    import models
    models.__path__.append(os.path.join(ROOT, "integration", "models"))
    yield models
    for m in [k for k in sys.modules if k.split(".")[0] in ("models", "options", "util", "data")]:
        sys.modules.pop(m, None)


def _parse(monkeypatch, model):
    from options.train_options import TrainOptions
    monkeypatch.setattr(sys, "argv", ARGV + ["--model", model, "--name", "dropin_" + model])
    return TrainOptions().parse()


def test_model_resolves_and_options_match_the_reference(ref_front_end, monkeypatch, capsys):
    from pcgan_b200.wsgan_emb_model import WSGANEmbModel
    cls = ref_front_end.find_model_using_name("wsgan_emb_b200")
    assert issubclass(cls, WSGANEmbModel) and issubclass(cls, ref_front_end.BaseModel)
    ours, ref = vars(_parse(monkeypatch, "wsgan_emb_b200")), vars(_parse(monkeypatch, "wsgan_emb"))
    assert ours.pop("cuda_graph") is False and ours.pop("group_passes") is True      # the two flags this package adds
    for k in ("model", "name"):
        ours.pop(k), ref.pop(k)
    assert ours == ref


def test_every_reference_flag_is_read_with_its_default(ref_front_end, monkeypatch):
    """default_options (what bench.py and the GPU tests use in place of the parser) carries the reference's defaults for
    every flag the model reads."""
    from pcgan_b200.wsgan_emb_model import default_options
    ref = vars(_parse(monkeypatch, "wsgan_emb"))
    mine = vars(default_options())
    north_star = dict(which_model_netG="resnet_9blocks", n_layers_D=3, lambda_IP=0.0, gpu_ids=[0], display_visuals=False,
                      embedding_bins="[]", pretrained_model_path_E="", pretrained_model_path_IP="", batchSize=10, upsample="bilinear", attr_bins=[], num_Ds=1,
                      checkpoints_dir="", name="", isTrain=True)
    for k, v in mine.items():
        if k in north_star or k.startswith("cuda_graph") or k == "group_passes" or k not in ref:
            continue
        assert ref[k] == v, (k, ref[k], v)


def test_no_cpu_path(ref_front_end, monkeypatch):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    opt = _parse(monkeypatch, "wsgan_emb_b200")
    with pytest.raises(RuntimeError, match="no CPU path"):
        ref_front_end.create_model(opt)
