"""The Elo rating network trainer of phymhan/pc-gan (siamese.py) on the B200-native encoder: the pairwise training step
of BASELINE config 2 (siamese.py:590-686) and the model factory it uses (siamese.py:422-480).

    net = get_model(gpu_ids=[0])                     # SiameseNetwork(ResNetFeature(resnet18), cnn_dim=[32, 1])
    trainer = EloTrainer(net, lr=2e-4)
    loss = trainer.train_step(img0, img1, label)     # label in {0, 1, 2}: img0 loses / draw / wins (LUT 0, .5, 1)

Forward and backward (data, weight and BatchNorm-parameter gradients of the whole ResNet-18 + head) run on
libpcgan_kernels.so; FusedAdam (torch.optim.Adam's semantics, one launch) updates the fp32 master weights.
"""
import itertools

import torch
from torch.nn import init

from . import networks
from .optim import FusedAdam


def weights_init(m):
    """siamese.py:288-296."""
    cn = m.__class__.__name__
    if cn.find("Conv") != -1 and hasattr(m, "weight"):
        init.normal_(m.weight.data, 0.0, 0.02)
    elif cn.find("BatchNorm2d") != -1:
        init.normal_(m.weight.data, 1.0, 0.02)
        init.constant_(m.bias.data, 0.0)


def get_model(which_model="resnet18", pooling="avg", cnn_dim=(32, 1), cnn_pad=1, cnn_relu_slope=0.7, bnn_dropout=0.0,
              noisy=False, rsample=False, gpu_ids=(0,), pretrained_model_path=""):
    """siamese.py:422-480 in train mode: base network + SiameseNetwork, N(0, 0.02) init, optional ImageNet base."""
    if not gpu_ids or not torch.cuda.is_available():
        raise RuntimeError("pcgan_b200 has no CPU path: the siamese trainer needs a CUDA device")
    base = networks.ResNetFeature(input_nc=3, which_model=which_model, dropout=bnn_dropout)
    net = networks.SiameseNetwork(base, pooling=pooling, cnn_dim=list(cnn_dim), cnn_pad=cnn_pad, cnn_relu_slope=cnn_relu_slope,
                                  fc_dim=[], noisy=noisy, drop_layer=networks.get_dropout_layer(bnn_dropout), rsample=rsample)
    net.apply(weights_init)
    if pretrained_model_path:
        net.load_pretrained(pretrained_model_path)
    if noisy:   # siamese.py:463-467: the log-variance head starts at zero
        last = list(net.cnn_logvar.parameters())
        last[-1].data.fill_(0)
        last[-2].data.fill_(0)
    return net.to(torch.device("cuda", gpu_ids[0]))


class EloTrainer:
    """criterion + optimizer + one training iteration of siamese.py:train (:526-551, :590-686; the non-noisy branches)."""

    def __init__(self, net, lr=2e-4, bayesian=False, T_train=1):
        self.net = net
        self.criterion = networks.BinaryNLLLoss()
        params = itertools.chain(net.base.parameters(), net.cnn.parameters())     # siamese.py:544-549 (no cxn, no fc)
        self.optimizer = FusedAdam(params, lr=lr)
        self.bayesian, self.T_train = bayesian, T_train
        if net._noisy:
            raise NotImplementedError("the noisy (aleatoric) trainer branches of siamese.py:606-660 are not implemented")

    def train_step(self, img0, img1, label):
        """Returns (loss tensor, prob_ of the last pass); siamese.py:598-600, 661-678."""
        self.optimizer.zero_grad()
        passes = self.T_train if self.bayesian else 1
        loss = 0.0
        for _ in range(passes):
            feat1, feat2, score = self.net(img0, img1)
            loss = loss + (1.0 / passes) * self.criterion.from_score(score, label)
        loss.backward()
        self.optimizer.step()
        with torch.no_grad():
            prob = torch.sigmoid(score)
        return loss.detach(), prob
