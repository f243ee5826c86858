"""WSGANEmbModel on the B200-native networks: the same public surface train.py drives
(models/base_model.py:43-159, models/wsgan_emb_model.py:82-497 of phymhan/pc-gan) —
initialize / setup / set_input / forward / optimize_parameters / get_current_losses /
get_current_visuals / save_networks / load_networks / update_learning_rate — with the
step's arithmetic running on libpcgan_kernels.so.

What stays PyTorch: parameter storage, torch.optim.Adam, LR schedulers,
checkpoint I/O, control flow.  One process per GPU; gradients are averaged across ranks by
pcgan_b200.dist.GradSync (the reference's nn.DataParallel replaced by NCCL all-reduce).

Supported flags: the wsgan_emb configuration with --which_model_netG resnet_9blocks (or
resnet_6blocks), --which_model_netD n_layers/basic, --lambda_IP (the AlexNet identity-preserving loss), --lambda_L1,
--lambda_A_GAN, --detach_fake_B, --use_real_A, --no_mixed_label_D, the four encoder modes
(--bayesian / --noisy with --noisy_var_type, --bnn_dropout, --bnn_T, --noisy_D, --noisy_rec) and
--lr_E > 0 (update_G_and_E).
"""
import os
from argparse import Namespace
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist
from torch.optim import lr_scheduler

from . import networks
from .dist import GradSync
from .networks import compute_mu_and_var, resample, upsample2d
from .optim import FusedAdam


def default_options(**overrides):
    """The flags WSGANEmbModel reads, with the reference's defaults under `--model wsgan_emb`
    (options/base_options.py:14-62, options/train_options.py:4-29, wsgan_emb_model.py:21-78) and the
    north-star overrides (--which_model_netG resnet_9blocks --n_layers_D 3 --lambda_IP 0)."""
    opt = Namespace(
        isTrain=True, gpu_ids=[0], checkpoints_dir="./checkpoints", name="experiment_name", transforms="resize_and_crop",
        batchSize=10, loadSize=128, fineSize=128, input_nc=3, output_nc=3, ngf=64, ndf=64,
        which_model_netG="resnet_9blocks", which_model_netD="n_layers", n_layers_D=3, n_layers_G=7,
        norm_G="instance", norm_D="batch", nl="relu", dropout=0.0, init_type="normal", upsample="bilinear", num_Ds=1,
        embedding_nc=1, which_model_netE="resnet18", pooling_E="avg", cnn_dim_E=[32, 1], no_cnn_E=False, cnn_pad_E=1,
        cnn_relu_slope_E=0.7, fineSize_E=224, pretrained_model_path_E="", embedding_mean=[0.0], embedding_std=[1.0],
        embedding_bins="[]", display_visuals=False, noisy=False, noisy_D=True, noisy_rec=True, noisy_var_type="",
        bayesian=False, bnn_dropout=0.0, bnn_T=10, attr_bins=[],
        which_model_netIP="alexnet", pretrained_model_path_IP="", fineSize_IP=224, identity_preserving_criterion="mse",
        lambda_L1=0.0, lambda_IP=0.0, lambda_z=1.0, lambda_A=0.5, lambda_A_GAN=0.0, lr_E=0.0, use_real_A=False,
        relabel_D=[0, 1, 0], no_mixed_label_D=False, weight_label_D=[0.5, 0, 0.5], detach_fake_B=False, update_logvar_E=False,
        no_lsgan=True, pool_size=0, lr=2e-4, beta1=0.5, lr_policy="lambda", niter=50, niter_decay=50, epoch_count=1,
        lr_decay_iters=50, continue_train=False, which_epoch="latest", load_model_names=[], verbose=False,
        cuda_graph=False, cuda_graph_warmup=3, cuda_graph_segments=None, group_passes=True)
    for k, v in overrides.items():
        setattr(opt, k, v)
    return opt


def str2bool(v):
    """util.str2bool (util/util.py:94-106)."""
    import argparse
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


def str2list(v):
    """util.str2list (util/util.py): the string form of a list ('[-2,-1,0,1,2]') or an existing list."""
    if isinstance(v, (list, tuple)):
        return list(v)
    import ast
    v = v.strip()
    return list(ast.literal_eval(v)) if v else []


def param_layers(net):
    """The parameters of `net` grouped into layers in forward order: a new layer starts at every convolution, the
    parameters of what follows it (its BatchNorm) belong to it.  Gradient buckets are unions of such layers: the backward
    sweep produces a layer's non-convolution gradients before it launches the layer's weight gradient."""
    layers = []
    for m in net.modules():
        own = list(m.parameters(recurse=False))
        if not own:
            continue
        if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)) or not layers:
            layers.append(own)
        else:
            layers[-1].extend(own)
    return layers


def _dist_ready():
    return dist.is_available() and dist.is_initialized()


def _rank():
    return dist.get_rank() if _dist_ready() else 0


def get_scheduler(optimizer, opt):
    """networks.get_scheduler (networks.py:57-69)."""
    if opt.lr_policy == "lambda":
        def rule(epoch):
            return 1.0 - max(0, epoch + 1 + opt.epoch_count - opt.niter) / float(opt.niter_decay + 1)
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=rule)
    if opt.lr_policy == "step":
        return lr_scheduler.StepLR(optimizer, step_size=opt.lr_decay_iters, gamma=0.1)
    if opt.lr_policy == "plateau":
        return lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=0.2, threshold=0.01, patience=5)
    raise NotImplementedError("learning rate policy [%s] is not implemented" % opt.lr_policy)


class BaseModel:
    """models/base_model.py:7-159."""

    def name(self):
        return "BaseModel"

    def initialize(self, opt):
        self.opt = opt
        self.gpu_ids = opt.gpu_ids
        if not self.gpu_ids or not torch.cuda.is_available():
            raise RuntimeError("pcgan_b200 has no CPU path: run with --gpu_ids <local device> on a CUDA machine")
        self.isTrain = opt.isTrain
        self.device = torch.device("cuda:{}".format(self.gpu_ids[0]))
        # the reference selects the device in BaseOptions.parse (options/base_options.py: torch.cuda.set_device); the
        # kernels are launched on the current device's stream, so it must be the one the tensors live on
        torch.cuda.set_device(self.device)
        # one process per GPU (torchrun ... --gpu_ids $LOCAL_RANK): join the job's process group if the launcher set one up
        if int(os.environ.get("WORLD_SIZE", "1")) > 1 and dist.is_available() and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.device)
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        self.loss_names, self.model_names, self.load_model_names, self.visual_names, self.image_paths = [], [], [], [], []
        self.current_iter = 0
        self.current_batch_size = opt.batchSize

    def setup(self, opt, parser=None):
        if self.isTrain:
            self.schedulers = [get_scheduler(o, opt) for o in self.optimizers]
        if not self.isTrain or opt.continue_train:
            self.load_networks(opt.which_epoch)
        self.broadcast_replicas()
        if _rank() == 0:
            self.print_networks(opt.verbose)

    def update_learning_rate(self):
        for s in self.schedulers:
            s.step()
        print("learning rate = %.7f" % float(self.optimizers[0].param_groups[0]["lr"]))

    def get_image_paths(self):
        return self.image_paths

    def eval(self):
        for n in self.model_names:
            if isinstance(n, str):
                getattr(self, "net" + n).eval()

    def print_networks(self, verbose):
        print("---------- Networks initialized -------------")
        for n in self.model_names:
            if isinstance(n, str):
                net = getattr(self, "net" + n)
                if verbose:
                    print(net)
                print("[Network %s] Total number of parameters : %.3f M" % (n, sum(p.numel() for p in net.parameters()) / 1e6))
        print("-----------------------------------------------")

    def broadcast_replicas(self):
        """Data-parallel replicas start identical (what nn.DataParallel's per-call replicate, networks.py:100, guarantees
        in the reference): every parameter and buffer of every network is broadcast from rank 0."""
        if not _dist_ready() or dist.get_world_size() == 1:
            return
        for n in self.model_names:
            if isinstance(n, str):
                net = getattr(self, "net" + n)
                for t in list(net.parameters()) + list(net.buffers()):
                    dist.broadcast(t.data, 0)

    def get_current_visuals(self):
        return OrderedDict((n, getattr(self, n)) for n in self.visual_names if isinstance(n, str))

    def get_current_losses(self):
        """base_model.py:77-83 (float(getattr(self, 'loss_' + name)) per name); here the device scalars are read back with
        ONE device -> host copy instead of one synchronising copy per loss."""
        names = [n for n in self.loss_names if isinstance(n, str)]
        vals = [getattr(self, "loss_" + n) for n in names]
        dev = [i for i, v in enumerate(vals) if torch.is_tensor(v) and v.is_cuda]
        if len(dev) > 1:
            host = torch.stack([vals[i].detach().float().reshape(()) for i in dev]).tolist()
            for i, h in zip(dev, host):
                vals[i] = h
        return OrderedDict((n, float(v.detach()) if torch.is_tensor(v) else float(v)) for n, v in zip(names, vals))

    def _unwrap(self, net):
        return net.module if isinstance(net, torch.nn.DataParallel) else net

    def save_networks(self, which_epoch):
        if _rank() != 0:      # replicas are identical: one writer
            return
        os.makedirs(self.save_dir, exist_ok=True)
        for n in self.model_names:
            if isinstance(n, str):
                path = os.path.join(self.save_dir, "%s_net_%s.pth" % (which_epoch, n))
                sd = {k: v.detach().cpu() for k, v in self._unwrap(getattr(self, "net" + n)).state_dict().items()}
                torch.save(sd, path)

    def load_networks(self, which_epoch):
        for n in (self.load_model_names or self.model_names):
            if isinstance(n, str):
                path = os.path.join(self.save_dir, "%s_net_%s.pth" % (which_epoch, n))
                sd = torch.load(path, map_location=str(self.device))
                self._unwrap(getattr(self, "net" + n)).load_state_dict(sd)
        self.invalidate_graphs()

    def invalidate_graphs(self):
        """Captured steps read packed bf16 operands that are refreshed by launches recorded at capture time; a weight
        change from outside the optimizers (checkpoint load, manual edit) needs a new capture."""
        if getattr(self, "_graphs", None):
            self._graphs = {}
            self._eager_steps = 0

    def set_requires_grad(self, nets, requires_grad=False):
        if not isinstance(nets, list):
            nets = [nets]
        for net in nets:
            if net is not None:
                for p in net.parameters():
                    p.requires_grad = requires_grad


# The flags `--model wsgan_emb` adds to the parser (wsgan_emb_model.py:21-66): (name, type, default[, extra argparse kw]).
_COMMON_FLAGS = [
    ("norm_G", str, "instance"), ("norm_D", str, "batch"), ("embedding_nc", int, 1), ("which_model_netE", str, "resnet18"),
    ("pooling_E", str, "avg"), ("cnn_dim_E", int, [32, 1], dict(nargs="+")), ("no_cnn_E", None, False), ("cnn_pad_E", int, 1),
    ("cnn_relu_slope_E", float, 0.7), ("fineSize_E", int, 224),
    ("pretrained_model_path_E", str, "pretrained_models/embedding_encoder.pth"),
    ("embedding_mean", float, [0.0], dict(nargs="*")), ("embedding_std", float, [1.0], dict(nargs="*")),
    ("embedding_bins", str, "[]"), ("display_visuals", None, False), ("noisy", str2bool, False), ("noisy_D", str2bool, True),
    ("noisy_rec", str2bool, True), ("noisy_var_type", str, ""), ("bayesian", str2bool, False), ("bnn_dropout", float, 0.0),
    ("bnn_T", int, 10), ("use_projection", str2bool, True), ("sample_embedding_B", str2bool, False),
]
_TRAIN_FLAGS = [
    ("lambda_L1", float, 0.0), ("lambda_IP", float, 1.0), ("lambda_z", float, 1.0), ("lambda_A", float, 0.5),
    ("lambda_A_GAN", float, 0.0), ("lambda_theta_D", float, 0.0), ("lambda_theta_E", float, 0.0),
    ("which_model_netIP", str, "alexnet"), ("pretrained_model_path_IP", str, "pretrained_models/alexnet-owt-4df8aa71.pth"),
    ("fineSize_IP", int, 224), ("lr_E", float, 0.0), ("use_real_A", None, False),
    ("identity_preserving_criterion", str, "mse"), ("relabel_D", int, [0, 1, 0], dict(nargs="*")),
    ("no_mixed_label_D", None, False), ("weight_label_D", float, [0.5, 0, 0.5], dict(nargs="*")),
    ("detach_fake_B", None, False), ("update_logvar_E", str2bool, False),
]
# defaults the model overrides on the base parser (wsgan_emb_model.py:68-80)
_DEFAULT_OVERRIDES = dict(pool_size=0, no_lsgan=True, norm="instance", dataset_mode="wsgan_emb", which_model_netG="unet_128",
                          which_model_netD="n_layers", n_layers_D=4, batchSize=10, loadSize=128, fineSize=128,
                          display_visuals=True, save_epoch_freq=2)


class WSGANEmbModel(BaseModel):
    def name(self):
        return "WSGANEmbModel"

    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        """The same flags and defaults as the reference's model class (wsgan_emb_model.py:20-80), plus --cuda_graph."""
        for spec in _COMMON_FLAGS + (_TRAIN_FLAGS if is_train else []):
            name, typ, default = spec[:3]
            kw = dict(spec[3]) if len(spec) > 3 else {}
            if typ is None:
                parser.add_argument("--" + name, action="store_true")
            else:
                parser.add_argument("--" + name, type=typ, default=default, **kw)
        if is_train:
            parser.add_argument("--cuda_graph", type=str2bool, default=False,
                                help="pcgan_b200: capture optimize_parameters() into a CUDA graph after a few eager steps and replay it")
            parser.add_argument("--group_passes", type=str2bool, default=True,
                                help="pcgan_b200: run the independent passes of a network (the three discriminator passes of backward_D, "
                                     "the two real-image encoder passes of forward) as one batch with one BatchNorm batch per pass")
        parser.set_defaults(**_DEFAULT_OVERRIDES)
        return parser

    def initialize(self, opt):
        BaseModel.initialize(self, opt)
        assert opt.input_nc == opt.output_nc
        self.attr_bins = getattr(opt, "attr_bins", [])
        self.embedding_bins = str2list(getattr(opt, "embedding_bins", "[]"))
        if "a" in opt.noisy_var_type and not opt.noisy:
            raise RuntimeError("Aleatoric only available when noisy is True.")
        if "e" in opt.noisy_var_type and not opt.bayesian:
            raise RuntimeError("Epistemic only available when bayesian is True.")
        if opt.no_cnn_E:
            opt.cnn_dim_E = []
        self.loss_names = ["G_GAN", "G_GAN_cycle", "G_IP", "G_L1", "G_cycle", "z_rec", "D_real_right", "D_real_wrong", "D_fake"]
        self.visual_names = ["real_A", "fake_B", "real_B", "rec_A"] if self.isTrain else ["real_A"]
        self.model_names = ["G", "D", "E"] if self.isTrain else ["G", "E"]
        self.load_model_names = opt.load_model_names
        self.netG = networks.define_G(opt.input_nc, opt.output_nc, opt.embedding_nc, opt.ngf, which_model_netG=opt.which_model_netG,
                                      norm=opt.norm_G, nl=opt.nl, dropout=opt.dropout, init_type=opt.init_type, gpu_ids=self.gpu_ids,
                                      upsample=opt.upsample, n_layers_G=opt.n_layers_G)
        self.netE = networks.define_E(opt.which_model_netE, 3, init_type=opt.init_type, pooling=opt.pooling_E, cnn_dim=opt.cnn_dim_E,
                                      cnn_pad=opt.cnn_pad_E, cnn_relu_slope=opt.cnn_relu_slope_E, gpu_ids=self.gpu_ids,
                                      fine_size_E=opt.fineSize_E, noisy=opt.noisy, bnn_dropout=opt.bnn_dropout)
        if self.isTrain and not opt.continue_train and opt.pretrained_model_path_E:
            self._unwrap(self.netE).load_pretrained(opt.pretrained_model_path_E)
        if self.isTrain:
            self.netD = networks.define_D(opt.output_nc, opt.embedding_nc, opt.ndf, opt.which_model_netD, opt.n_layers_D, opt.norm_D,
                                          opt.no_lsgan, opt.init_type, num_Ds=opt.num_Ds, gpu_ids=self.gpu_ids)
            # the identity-preserving network, which is not saved (:129-135); built only when its loss is on
            self.netIP = None
            if opt.lambda_IP > 0.0:
                self.netIP = networks.define_IP(getattr(opt, "which_model_netIP", "alexnet"), opt.input_nc, self.gpu_ids)
                path = getattr(opt, "pretrained_model_path_IP", "")
                if path:
                    self._unwrap(self.netIP).load_pretrained(path)
                self.set_requires_grad(self.netIP, False)     # never optimised: its weight gradients are not computed
            crit = getattr(opt, "identity_preserving_criterion", "mse").lower()
            if crit not in ("mse", "l1"):
                raise NotImplementedError("Not Implemented")
            self.criterionIP = networks.mse_loss if crit == "mse" else networks.l1_loss
            assert opt.pool_size == 0
            self.criterionGAN = networks.GANLoss(use_lsgan=not opt.no_lsgan)
            self.criterionL1 = networks.l1_loss
            self.criterionRec = networks.mse_loss
            self.criterionCycle = networks.l1_loss
            # --cuda_graph: the whole optimize_parameters() (about 950 kernel launches) is captured once and replayed, so
            # the step costs one graph launch of host time; the learning rate then lives in a device scalar the schedulers fill.
            self.use_graph = bool(getattr(opt, "cuda_graph", False))
            self._graphs, self._eager_steps, self._side = {}, 0, None
            # FusedAdam is torch.optim.Adam (same state, same schedulers) whose step() is one multi-tensor launch
            adam_kw = dict(betas=(opt.beta1, 0.999))
            lr = torch.tensor(float(opt.lr), device=self.device) if self.use_graph else opt.lr
            self.optimizer_G = FusedAdam(self.netG.parameters(), lr=lr, **adam_kw)
            lr = torch.tensor(float(opt.lr), device=self.device) if self.use_graph else opt.lr
            self.optimizer_D = FusedAdam(self.netD.parameters(), lr=lr, **adam_kw)
            self.optimizers = [self.optimizer_G, self.optimizer_D]
            # one process per GPU: flat gradient buffers, averaged over ranks with one NCCL all-reduce per network
            self.sync_G = GradSync(list(self.netG.parameters()), layers=param_layers(self.netG))
            self.sync_D = GradSync(list(self.netD.parameters()), layers=param_layers(self.netD))
            # weight gradients accumulate in packed form over the backward sweeps of an update and reach .grad in one launch
            for net in (self.netG, self.netD):
                self._unwrap(net).defer_wgrad = True
            self.sync_E = None
            if opt.lr_E > 0.0:      # wsgan_emb_model.py:159-163
                if self.use_graph:
                    raise NotImplementedError("--cuda_graph with --lr_E > 0 (two generator updates per step) is not supported")
                if getattr(opt, "update_logvar_E", False):      # :157-158: only the log-variance head is trained
                    assert opt.noisy
                    params_E = list(self._unwrap(self.netE).cnn_logvar.parameters())
                    trained = {id(p) for p in params_E}
                    for p in self.netE.parameters():
                        if id(p) not in trained:
                            p.requires_grad = False
                else:
                    params_E = list(self.netE.parameters())
                self.optimizer_E = FusedAdam(params_E, lr=opt.lr_E, betas=(opt.beta1, 0.999))
                self.optimizers.append(self.optimizer_E)
                self.sync_E = GradSync(params_E)
                # backward_GE keeps the graph (retain_graph=True, :369) and backward_G_alone walks G's first pass and
                # E(real_B) again: their workspaces must outlive the first backward
                for net in (self.netG, self.netE):
                    self._unwrap(net).retain_workspaces = True
            else:
                self.set_requires_grad(self.netE, False)   # :164-165 (E stays in train mode: SURVEY A.1)
            self.relabel_D = opt.relabel_D
            self._relabel_lut = torch.tensor([float(v) for v in opt.relabel_D], device=self.device)
            self._static = {}
            if len(opt.weight_label_D) > 0:
                assert len(opt.weight_label_D) == len(opt.relabel_D)
                self.weight_label_D = [w / sum(opt.weight_label_D) for w in opt.weight_label_D]
            else:
                self.weight_label_D = None
        self.transform_IP = networks.Normalize((0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010))
        self.embedding_normalize = lambda x: (x - opt.embedding_mean[0]) / opt.embedding_std[0]
        if getattr(opt, "display_visuals", False):
            self.pre_generate_embeddings(self.embedding_bins)
        self.transform_E = networks.Normalize((0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010))

    def pre_generate_embeddings(self, embeddings_list):
        """wsgan_emb_model.py:184-191: the normalised embeddings the aging visuals are generated for."""
        self.fixed_embeddings = [self.embedding_normalize(torch.tensor([[[[float(v)]]]], device=self.device))
                                 for v in np.array(embeddings_list).reshape(-1)]

    # ------------------------------------------------------------------ data
    def _stage(self, name, src):
        """Copy a batch tensor into a persistent device buffer (one per name and shape): the captured step reads fixed
        addresses, and the eager step simply uses the same buffers."""
        src = torch.as_tensor(src)
        key = (name, tuple(src.shape), src.dtype)
        buf = self._static.get(key)
        if buf is None:
            buf = self._static[key] = torch.empty(src.shape, dtype=src.dtype, device=self.device)
        buf.copy_(src, non_blocking=True)
        return buf

    def set_input(self, input):
        """wsgan_emb_model.py:193-212.  The pair labels stay on the host in the reference (:199) and are turned into
        discriminator targets with a Python list comprehension (:324); here they are uploaded once with the images so
        that the step itself issues no host-to-device copy."""
        if self.isTrain:
            if not self.opt.no_mixed_label_D:
                self.real_A = self._stage("A", input["A"])
                self.real_B = self._stage("B", input["B"])
                self.image_paths = input.get("B_paths", [])
                self.label_AB = input["label"]
            else:
                self.label_AB = [np.random.choice(range(len(self.relabel_D)), p=self.weight_label_D)]
                k = str(self.label_AB[0])
                self.real_A = self._stage("A", input[k + "_A"])
                self.real_B = self._stage("B", input[k + "_B"])
                self.image_paths = input.get(k + "_B_paths", [])
            lab = torch.as_tensor(self.label_AB).to(torch.int64).reshape(-1)
            if lab.numel() == 1 and self.real_A.size(0) > 1:
                lab = lab.expand(self.real_A.size(0)).contiguous()
            self._label_dev = self._stage("label", lab)
        else:
            self.real_A = input["A"].to(self.device)
            self.image_paths = input.get("A_paths", [])
            if "B" in input:
                self.real_B = input["B"].to(self.device)
                self.image_paths = input.get("B_paths", self.image_paths)     # :208-210
        self.current_iter += 1
        self.current_batch_size = int(self.real_A.size(0))

    # --------------------------------------------------------------- forward
    def _encode(self, x_E):
        """The four encoder modes of forward() (:218-240): returns y and, when --noisy_var_type asks for it, the variance
        the embedding is resampled with."""
        opt = self.opt
        x = self.transform_E(x_E)
        var = None
        if not opt.bayesian and not opt.noisy:
            y = self.netE(x)
        elif not opt.bayesian and opt.noisy:
            y, logvar = self.netE(x)
            if "a" in opt.noisy_var_type:
                var = torch.exp(logvar)
        elif opt.bayesian and not opt.noisy:
            y, y_var = compute_mu_and_var(self.netE, x, opt.bnn_T, False)
            if "e" in opt.noisy_var_type:
                var = y_var
        else:
            y, y_var, y_s2 = compute_mu_and_var(self.netE, x, opt.bnn_T, True)
            if "a" in opt.noisy_var_type:
                var = y_s2 + y_var
        return y, var

    def forward(self):
        """wsgan_emb_model.py:214-259."""
        opt = self.opt
        ip = self.isTrain and opt.lambda_IP > 0.0
        if ip:
            self.real_A_IP = upsample2d(self.real_A, opt.fineSize_IP)      # :215
        netE = self._unwrap(self.netE)
        if getattr(opt, "group_passes", False) and opt.lr_E <= 0.0 and not opt.bayesian and netE.can_group():
            # the two real-image encoder passes as one batch of 2 N samples with one BatchNorm batch per image set
            n = self.real_A.size(0)
            both = upsample2d(torch.cat([self.real_A, self.real_B], 0), opt.fineSize_E)
            self.real_A_E, self.real_B_E = both[:n], both[n:]
            with netE.grouped(2):
                out = self.netE(self.transform_E(both))
            y, logvar = out if opt.noisy else (out, None)
            y_A, y_B = y[:n], y[n:]
            var_A = var_B = None
            if opt.noisy and "a" in opt.noisy_var_type:
                var_A, var_B = torch.exp(logvar[:n]), torch.exp(logvar[n:])
        else:
            self.real_A_E = upsample2d(self.real_A, opt.fineSize_E)
            self.real_B_E = upsample2d(self.real_B, opt.fineSize_E)
            y_A, var_A = self._encode(self.real_A_E)
            y_B, var_B = self._encode(self.real_B_E)
        if var_A is not None:
            self.resample_A = self.embedding_normalize(resample(y_A, var_A))
            self.resample_B = self.embedding_normalize(resample(y_B, var_B))
        self.y_A, self.y_B = y_A, y_B
        self.embedding_A = self.embedding_normalize(self.y_A)
        self.embedding_B = self.embedding_normalize(self.y_B)
        if opt.lr_E <= 0.0:
            self.y_A, self.y_B = self.y_A.detach(), self.y_B.detach()
            self.embedding_A, self.embedding_B = self.embedding_A.detach(), self.embedding_B.detach()
            if opt.noisy_var_type:
                self.resample_A, self.resample_B = self.resample_A.detach(), self.resample_B.detach()
        self.fake_B = self.netG(self.real_A, self.embedding_B)
        self.fake_B_E = upsample2d(self.fake_B, opt.fineSize_E)
        if ip:
            self.fake_B_IP = upsample2d(self.fake_B, opt.fineSize_IP)      # :254
        src = self.fake_B.detach() if opt.detach_fake_B else self.fake_B
        self.rec_A = self.netG(src, self.embedding_A)

    def test(self):
        """wsgan_emb_model.py:261-277."""
        with torch.no_grad():
            if hasattr(self, "real_B"):
                if "real_B" not in self.visual_names:
                    self.visual_names += ["real_B", "fake_B"]
                self.fake_B = self.sample_from_prior()

    def sample_from_prior(self):
        """wsgan_emb_model.py:279-292: A -> B with the embedding the encoder reads off real_B."""
        y_B, _ = self._encode(upsample2d(self.real_B, self.opt.fineSize_E))
        self.embedding_B = self.embedding_normalize(y_B.detach())
        return self.netG(self.real_A, self.embedding_B)

    def sample_from_label(self, label):
        """wsgan_emb_model.py:294-298."""
        emb = torch.tensor([[[[float(self.embedding_bins[label])]]]], device=self.device)
        n = self.real_A.size(0)
        return self.netG(self.real_A, self.embedding_normalize(emb).expand(n, 1, 1, 1).contiguous())

    # -------------------------------------------------------------- backward
    def _cond_B(self):
        """what D is conditioned on for the fake image (:303-306, 373-376)"""
        return self.resample_B if (self.opt.noisy_var_type and self.opt.noisy_D) else self.embedding_B

    def backward_D(self):
        """wsgan_emb_model.py:300-329.  The three discriminator passes are independent (detached inputs, separate
        BatchNorm batches); with --group_passes they run as ONE batch of 3 N samples whose groups of N keep their own
        BatchNorm statistics, running-statistics step and backward sums, in the reference's order (fake, right, wrong)."""
        opt = self.opt
        img = self.real_A if opt.use_real_A else self.real_B
        emb_right, emb_wrong = (self.embedding_A, self.embedding_B) if opt.use_real_A else (self.embedding_B, self.embedding_A)
        target_label = self._relabel_lut[self._label_dev]      # [relabel_D[l] for l in label_AB] (:324), on the device
        if getattr(opt, "group_passes", False):
            n = img.size(0)
            x = torch.cat([self.fake_B.detach(), img, img], 0)
            z = torch.cat([self._cond_B().detach().reshape(n, -1), emb_right.detach().reshape(n, -1), emb_wrong.detach().reshape(n, -1)], 0)
            with self._unwrap(self.netD).grouped(3):
                pred = self.netD(x, z.view(3 * n, -1, 1, 1))
            pred_fake, pred_right, pred_wrong = pred[:n], pred[n:2 * n], pred[2 * n:]
        else:
            pred_fake = self.netD(self.fake_B.detach(), self._cond_B().detach())
            pred_right = self.netD(img, emb_right.detach())
            pred_wrong = self.netD(img, emb_wrong.detach())
        self.loss_D_fake = self.criterionGAN(pred_fake, False)
        self.loss_D_real_right = self.criterionGAN(pred_right, True)
        self.loss_D_real_wrong = self.criterionGAN(pred_wrong, target_label)
        self.loss_D = (self.loss_D_fake + (self.loss_D_real_right + self.loss_D_real_wrong) * 0.5) * 0.5
        self.loss_D.backward()

    def _generator_losses(self):
        """the terms backward_G and backward_GE share (:333-364, 372-404)"""
        opt = self.opt
        self.loss_G_GAN = self.criterionGAN(self.netD(self.fake_B, self._cond_B()), True)
        if opt.lambda_A_GAN > 0.0:
            self.loss_G_GAN_cycle = self.criterionGAN(self.netD(self.rec_A, self.embedding_A), True) * opt.lambda_A_GAN
        else:
            self.loss_G_GAN_cycle = 0.0
        self.loss_G_L1 = self.criterionL1(self.fake_B, self.real_A) * opt.lambda_L1 if opt.lambda_L1 > 0.0 else 0.0
        if opt.lambda_IP > 0.0:      # :353-356, 393-396
            with torch.no_grad():
                feature_A = self.netIP(self.transform_IP(self.real_A_IP))
            self.loss_G_IP = self.criterionIP(self.netIP(self.transform_IP(self.fake_B_IP)), feature_A) * opt.lambda_IP
        else:
            self.loss_G_IP = 0.0
        self.loss_G_cycle = self.criterionCycle(self.rec_A, self.real_A) * opt.lambda_A if opt.lambda_A > 0.0 else 0.0
        return self.loss_G_GAN + self.loss_G_IP + self.loss_G_L1 + self.loss_G_cycle + self.loss_G_GAN_cycle

    def backward_G(self):
        """wsgan_emb_model.py:371-437 with lambda_IP = 0."""
        opt = self.opt
        partial = self._generator_losses()
        if opt.lambda_z > 0.0:
            y_var = y_logvar = None
            if not opt.bayesian and not opt.noisy:
                pred_y = self.netE(self.transform_E(self.fake_B_E))
            elif not opt.bayesian and opt.noisy:
                pred_y, y_logvar = self.netE(self.transform_E(self.fake_B_E))
                if "a" in opt.noisy_var_type:
                    y_var = torch.exp(y_logvar)
            elif opt.bayesian and not opt.noisy:
                pred_y, y_var = compute_mu_and_var(self.netE, self.transform_E(self.fake_B_E), opt.bnn_T, False)
                if "e" in opt.noisy_var_type:
                    y_logvar = torch.log(y_var + 1e-20)
            else:
                # bayesian and noisy: the reference evaluates the encoder on real_A_E here (:419), so this term carries
                # no gradient to G; reproduced as is
                pred_y, y_var_, y_s2_ = compute_mu_and_var(self.netE, self.transform_E(self.real_A_E), opt.bnn_T, True)
                y_var = torch.zeros_like(pred_y)
                if "a" in opt.noisy_var_type:
                    y_var = y_var + y_s2_
                if "e" in opt.noisy_var_type:
                    y_var = y_var + y_var_
                y_logvar = torch.log(y_var + 1e-20)
            self.pred_y = pred_y.detach()      # kept for inspection (tests check the loss arithmetic on it)
            if opt.noisy_var_type and opt.noisy_rec:
                self.loss_z_rec = ((pred_y - self.y_B).pow(2) / y_var.detach() + y_logvar.detach()).sum() / pred_y.size(0) * 0.5 * opt.lambda_z
            else:
                self.loss_z_rec = self.criterionRec(pred_y, self.y_B) * opt.lambda_z
        else:
            self.loss_z_rec = 0.0
        self.loss_G = partial + self.loss_z_rec
        if torch.is_tensor(self.loss_G) and self.loss_G.requires_grad:
            self.loss_G.backward()

    def backward_GE(self):
        """wsgan_emb_model.py:331-369: the generator terms with the graph kept, gradients reach E through the embeddings."""
        self.loss_G = self._generator_losses()
        self.loss_G.backward(retain_graph=True)

    def backward_G_alone(self):
        """wsgan_emb_model.py:439-449."""
        opt = self.opt
        if opt.lambda_z > 0.0:
            out = self.netE(self.transform_E(self.fake_B_E))
            pred_embedding = self.embedding_normalize(out[0] if opt.noisy else out)
            self.loss_z_rec = self.criterionRec(pred_embedding, self.embedding_B.detach()) * opt.lambda_z
            self.loss_z_rec.backward()
        else:
            self.loss_z_rec = 0.0

    def _attach(self, net, sync):
        """hand the network's programs the GradSync whose buckets their backward sweeps complete"""
        for prog in self._unwrap(net)._programs.values():
            prog.bank.attach_sync(sync)

    def _grads_D(self):
        """zero_grad + backward_D; the gradient all-reduce is started bucket by bucket during the sweep, not waited for"""
        self.set_requires_grad(self.netD, True)
        self.sync_D.zero()                 # optimizer_D.zero_grad()
        self._unwrap(self.netD).zero_wgrad()
        self._attach(self.netD, self.sync_D)
        self.backward_D()
        self._unwrap(self.netD).flush_wgrad()
        self.sync_D.start()

    def _grads_G(self):
        self.set_requires_grad(self.netD, False)
        self.sync_G.zero()                 # optimizer_G.zero_grad()
        self._unwrap(self.netG).zero_wgrad()
        self._attach(self.netG, self.sync_G)
        self.backward_G()
        self._unwrap(self.netG).flush_wgrad()
        self.sync_G.start()

    def update_D(self):
        self._grads_D()
        self.sync_D.finish()
        self.optimizer_D.step()

    def update_G(self):
        self._grads_G()
        self.sync_G.finish()
        self.optimizer_G.step()

    def update_G_and_E(self):
        """wsgan_emb_model.py:463-476."""
        self.set_requires_grad(self.netD, False)
        self.sync_G.zero()
        self.sync_E.zero()
        self._unwrap(self.netG).zero_wgrad()
        self.backward_GE()
        self._unwrap(self.netG).flush_wgrad()
        self.sync_G.all_reduce()
        self.sync_E.all_reduce()
        self.optimizer_G.step()
        self.optimizer_E.step()
        if self.opt.lambda_z > 0.0:
            self.sync_G.zero()
            self.sync_E.zero()
            self._unwrap(self.netG).zero_wgrad()
            self.backward_G_alone()
            self._unwrap(self.netG).flush_wgrad()
            self.sync_G.all_reduce()
            self.optimizer_G.step()

    def _step(self):
        """optimize_parameters (:478-484).  update_D reads fake_B (made by forward(), before G's update) and netD only, so
        G's optimizer step commutes with D's backward: the sweeps of both updates are enqueued first, G's gradient
        all-reduce (started bucket by bucket inside its own sweep) runs under D's forward / backward, and D's under G's
        Adam step; each optimizer waits only for its own collective."""
        self.forward()
        if self.opt.lr_E > 0.0:
            self.update_G_and_E()
            self.update_D()
            return
        self._grads_G()
        self._grads_D()
        self.sync_G.finish()
        self.optimizer_G.step()
        self.sync_D.finish()
        self.optimizer_D.step()

    # Fallback (--cuda_graph_segments true): the step as three collective-free graphs with the two gradient all-reduces
    # between them as ordinary NCCL calls, for NCCL builds whose collectives cannot be captured.
    def _seg_forward_backward_G(self):
        self.forward()
        self.set_requires_grad(self.netD, False)
        self.sync_G.zero()
        self._unwrap(self.netG).zero_wgrad()
        self._attach(self.netG, None)
        self.backward_G()
        self._unwrap(self.netG).flush_wgrad()

    def _seg_step_G_backward_D(self):
        self.optimizer_G.step()
        self.set_requires_grad(self.netD, True)
        self.sync_D.zero()
        self._unwrap(self.netD).zero_wgrad()
        self._attach(self.netD, None)
        self.backward_D()
        self._unwrap(self.netD).flush_wgrad()

    def _seg_step_D(self):
        self.optimizer_D.step()

    def _run_segments(self, segs):
        segs[0]()
        self.sync_G.all_reduce()
        segs[1]()
        self.sync_D.all_reduce()
        segs[2]()

    def optimize_parameters(self):
        """wsgan_emb_model.py:478-484.  With --cuda_graph the step is captured after a few eager steps (plans built,
        workspaces pooled, Adam state allocated) and replayed from then on; one capture per batch shape.  On one GPU
        The whole step, including the bucketed NCCL all-reduces of a multi-rank job, is ONE graph (captured in thread-local
        mode: NCCL's watchdog thread polls events while the capture is open)."""
        if not self.use_graph:
            return self._step()
        key = (tuple(self.real_A.shape), tuple(self.real_B.shape))
        g = self._graphs.get(key)
        if g is not None:
            if isinstance(g, list):
                self._run_segments([x.replay for x in g])
            else:
                g.replay()
            return
        # Eager warm-up and capture run on one side stream: autograd ties every parameter's gradient accumulator to the
        # stream of its first use, and a capture may only depend on work of the capturing stream.
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        if self._eager_steps < int(getattr(self.opt, "cuda_graph_warmup", 3)):
            self._eager_steps += 1
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self._step()
            cur.wait_stream(self._side)
            return
        torch.cuda.synchronize()
        segmented = bool(getattr(self.opt, "cuda_graph_segments", None))
        if not segmented:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._side, capture_error_mode="thread_local"):
                self._step()
            self._graphs[key] = g
            g.replay()      # the capture itself executes nothing: run the step that was asked for
            return
        # several ranks: capture each segment, then run it (the next segment's capture needs its side effects in place:
        # the all-reduced gradients, the updated weights), sharing one memory pool so tensors live across segments
        graphs, pool = [], None
        fns = (self._seg_forward_backward_G, self._seg_step_G_backward_D, self._seg_step_D)
        for i, fn in enumerate(fns):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._side, pool=pool):
                fn()
            pool = g.pool()
            g.replay()
            if i == 0:
                self.sync_G.all_reduce()
            elif i == 1:
                self.sync_D.all_reduce()
            graphs.append(g)
        self._graphs[key] = graphs

    def get_current_visuals(self):
        """wsgan_emb_model.py:486-497: the step's images plus, with --display_visuals, real_A[0] aged to every fixed
        embedding (generator passes without gradients)."""
        out = OrderedDict((n, getattr(self, n)) for n in self.visual_names if isinstance(n, str) and hasattr(self, n))
        if getattr(self.opt, "display_visuals", False):
            self.set_requires_grad(self.netG, False)
            for i, emb in enumerate(self.fixed_embeddings):
                out["attr_%d" % i] = self.netG(self.real_A[0:1, ...].contiguous(), emb)
            self.set_requires_grad(self.netG, True)
        return out
