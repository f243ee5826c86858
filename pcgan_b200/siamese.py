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


MAGIC_EPS = 1e-20      # siamese.py:24


class EloTrainer:
    """criterion + optimizers + one training iteration of siamese.py:train (:526-565, :590-686), every branch:
    plain / bayesian (T_train stochastic passes), noisy (aleatoric twin head, second Adam over cnn_logvar with lr_sigma),
    rsample with the Monte-Carlo ('mc') or lower-bound ('lb') objective over M reparameterised draws.

    cuda_graph=True captures the whole iteration (two encoder passes, loss, backward, Adam) after two eager iterations
    and replays it: the step then costs one graph launch of host time instead of ~500 kernel launches from Python."""

    def __init__(self, net, lr=2e-4, bayesian=False, T_train=1, lr_sigma=None, rsample=None, lb_or_mc="mc", M=1, cuda_graph=False):
        self.net = net
        self.criterion = networks.BinaryNLLLoss()
        params = itertools.chain(net.base.parameters(), net.cnn.parameters())     # siamese.py:544-549 (no cxn, no fc)
        self.use_graph = bool(cuda_graph)
        dev = next(net.parameters()).device
        mk_lr = (lambda v: torch.tensor(float(v), device=dev)) if self.use_graph else float
        self.optimizer = FusedAdam(params, lr=mk_lr(lr))
        self.bayesian, self.T_train = bayesian, T_train
        self.noisy = bool(net._noisy)
        self.rsample = bool(net._rsample if rsample is None else rsample)
        self.lb_or_mc, self.M = lb_or_mc, int(M)
        self.optimizer_sigma = None
        if self.noisy:                                                            # siamese.py:552-553
            self.optimizer_sigma = FusedAdam(net.cnn_logvar.parameters(), lr=mk_lr(lr if lr_sigma is None else lr_sigma))
        self._graph, self._eager, self._static, self._side = None, 0, None, None

    # ---------------------------------------------------------------- objective
    def _loss(self, img0, img1, label):
        """siamese.py:606-677; returns (loss, prob_ of the last pass)"""
        crit, net = self.criterion, self.net
        passes = self.T_train if self.bayesian else 1
        loss, prob = 0.0, None
        for _ in range(passes):
            if not self.noisy:
                _, _, score = net(img0, img1)
                loss = loss + (1.0 / passes) * crit.from_score(score, label)
                prob = score
            elif not self.rsample:
                _, _, score, score_std = net(img0, img1)
                prob = torch.sigmoid(score / (score_std + MAGIC_EPS))
                loss = loss + (1.0 / passes) * crit(prob, label)
            else:
                y1, y2, lv1, lv2 = net(img0, img1)
                if self.lb_or_mc == "mc":
                    prob = 0.0
                    for _m in range(self.M):
                        prob = prob + (1.0 / self.M) * torch.sigmoid(networks.reparameterize(y1, lv1) - networks.reparameterize(y2, lv2))
                    loss = loss + (1.0 / passes) * crit(prob, label)
                else:
                    for _m in range(self.M):
                        prob = torch.sigmoid(networks.reparameterize(y1, lv1) - networks.reparameterize(y2, lv2))
                        loss = loss + (1.0 / (passes * self.M)) * crit(prob, label)
        return loss, prob

    def _iteration(self, img0, img1, label):
        self.optimizer.zero_grad()
        if self.optimizer_sigma is not None:
            self.optimizer_sigma.zero_grad()
        loss, prob = self._loss(img0, img1, label)
        loss.backward()
        self.optimizer.step()
        if self.optimizer_sigma is not None:
            self.optimizer_sigma.step()
        with torch.no_grad():
            prob = torch.sigmoid(prob) if not self.noisy else prob.detach()
        return loss.detach(), prob

    # ------------------------------------------------------------------- driver
    def train_step(self, img0, img1, label):
        """Returns (loss tensor, prob_ of the last pass); siamese.py:598-600, 661-678."""
        if not self.use_graph:
            return self._iteration(img0, img1, label)
        key = (tuple(img0.shape), tuple(label.shape))
        if self._static is None or self._static[0] != key:
            self._static = (key, torch.empty_like(img0), torch.empty_like(img1), torch.empty_like(label))
            self._graph, self._eager = None, 0
        _, s0, s1, sl = self._static
        s0.copy_(img0, non_blocking=True); s1.copy_(img1, non_blocking=True); sl.copy_(label, non_blocking=True)
        if self._graph is not None:
            self._graph.replay()
            return self._out
        # eager warm-up and capture on one side stream (autograd ties gradient accumulators to the stream of first use)
        if self._side is None:
            self._side = torch.cuda.Stream(device=img0.device)
        cur = torch.cuda.current_stream(img0.device)
        if self._eager < 2:
            self._eager += 1
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                out = self._iteration(s0, s1, sl)
            cur.wait_stream(self._side)
            return out
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._side):
            self._out = self._iteration(s0, s1, sl)
        self._graph = g
        g.replay()
        return self._out
