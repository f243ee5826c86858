"""torch.library custom operators of the wsgan_emb hot path (namespace `pcgan`).

Every operator is a thin host-side shim: it resolves the module / plan it was called for, hands raw device pointers to
the C ABI (include/pcgan_kernels.h) and returns torch tensors that only own memory.  Autograd is wired with
torch.library.register_autograd and shape inference with register_fake, so the operators are visible to the dispatcher
(torch.ops.pcgan.*), to FakeTensor tracing and to torch.library.opcheck.

    pcgan::resnet_generator(x, z, params, key)              ResnetGenerator.forward   (models/networks.py:609-612)
    pcgan::nlayer_discriminator(x, z, params, key)          NLayerDiscriminator.forward (:779-783)
    pcgan::siamese_feature(x, params, key)                  SiameseFeature.forward    (:1051-1068)
    pcgan::reduce_loss(kind, pred, target, per_sample)      nn.BCELoss / MSELoss / L1Loss / BinaryNLLLoss (:386-420, 473-482)
    pcgan::upsample_bilinear_ac(x, size)                    util.upsample2d           (util/util.py:111-117)
  and their *_backward companions.

Statefulness, stated plainly: the loss and resize operators are pure functions (torch.library.opcheck passes on them,
tests/test_opcheck_gpu.py).  The three network operators are not: like an nn.Module in training mode they update the
module's running statistics (and num_batches_tracked) in place, lease a pooled workspace that their backward gives back,
and their backward accumulates parameter gradients in packed form that the program scatters into .grad at the end of the
update instead of returning them through autograd.  torch.library accepts autograd formulas only for operators whose
schema declares no mutation, so these side effects are NOT in the schemas (`mutates_args=()`): the operators are meant
for eager execution and CUDA-graph capture, not for functionalising tracers (torch.compile would reorder or drop the
buffer updates).  Leases live in per-thread tables, so concurrent calls from several threads do not cross.

A network operator works on a whole network at once (one static program of kernel launches per input geometry) rather
than on single layers: the per-layer state (padded NHWC buffers, statistics, plans) lives in the program's pooled
workspace, which the forward leases and the backward gives back.  `key` identifies the module instance.
"""
import itertools
import threading
import weakref
from typing import List, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib as L
from . import ops

_MODULES = {}            # key -> weakref(module)
_KEYS = itertools.count(1)


class _PerThread(threading.local):
    """key -> lease tables, one per thread: a forward and the setup_context / finish_forward that claims its lease run on
    the same thread back to back, and a backward sets and reads its lease on the autograd worker it runs on, so calls of
    one module from several threads cannot take each other's workspaces."""

    def __init__(self):
        self.pending, self.bwd = {}, {}


_TLS = _PerThread()


def register_module(mod) -> int:
    key = next(_KEYS)
    _MODULES[key] = weakref.ref(mod)
    return key


def _module(key):
    ref = _MODULES.get(key)
    mod = ref() if ref is not None else None
    if mod is None:
        raise L.PcganError("pcgan custom op: module %d is gone" % key)
    return mod


class Lease:
    """A workspace borrowed by one forward call whose backward is still to come.  It goes back to the pool right after
    that backward, or — if the graph is dropped without a backward, or the module keeps workspaces for a second backward
    through the same graph (retain_workspaces: loss.backward(retain_graph=True), wsgan_emb_model.py:369) — when the
    autograd node that holds the lease dies."""

    def __init__(self, prog, ws):
        self.prog, self.ws = prog, ws

    def release(self):
        if self.ws is not None:
            self.prog.pool.give(0, self.ws)
            self.ws = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def finish_forward(key, out):
    """Called by the module right after its operator: a forward nobody will differentiate returns its workspace now."""
    lease = _TLS.pending.pop(key, None)
    if lease is not None and not (torch.is_grad_enabled() and out.requires_grad):
        lease.release()


def _claim(ctx, key):
    ctx.lease = _TLS.pending.pop(key)


def _after_backward(ctx):
    if not getattr(ctx.lease.prog.mod, "retain_workspaces", False):
        ctx.lease.release()


def _empty(ref):
    return torch.empty(0, device=ref.device)


# ------------------------------------------------------------------------------ generator
@torch.library.custom_op("pcgan::resnet_generator", mutates_args=())
def resnet_generator(x: Tensor, z: Tensor, params: Sequence[Tensor], key: int) -> Tensor:
    mod = _module(key)
    prog = mod._program(x.shape[0], x.shape[2])
    out, ws = prog.forward(x.contiguous().float(), z.contiguous().float().view(-1))
    _TLS.pending[key] = Lease(prog, ws)
    return out


@resnet_generator.register_fake
def _(x, z, params, key):
    return x.new_empty((x.shape[0], _module(key).output_nc, x.shape[2], x.shape[3]), dtype=torch.float32)


@torch.library.custom_op("pcgan::resnet_generator_backward", mutates_args=())
def resnet_generator_backward(dout: Tensor, out: Tensor, key: int, need_dx: bool, need_dz: bool, need_w: bool) -> Tuple[Tensor, Tensor]:
    lease = _TLS.bwd[key]
    dx, dz = lease.prog.backward(lease.ws, out, dout.contiguous(), need_dx, need_w, need_dz)
    return (dx if dx is not None else _empty(dout)), (dz if dz is not None else _empty(dout))


@resnet_generator_backward.register_fake
def _(dout, out, key, need_dx, need_dz, need_w):
    mod = _module(key)
    n, s = out.shape[0], out.shape[2]
    return (out.new_empty((n, mod.input_nc_img, s, s)) if need_dx else out.new_empty(0)), (out.new_empty(n) if need_dz else out.new_empty(0))




def _net_setup(ctx, inputs, output):
    x, z, params, key = inputs
    _claim(ctx, key)
    ctx.key = key
    ctx.need_dx, ctx.need_dz = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    ctx.need_w = any(p.requires_grad for p in params)
    if ctx.need_w:
        ctx.lease.prog.bank.pending += 1      # one more weight-gradient sweep to come (WeightBank.begin_backward)
    ctx.n_params, ctx.z_shape = len(params), z.shape
    ctx.save_for_backward(output)


def _net_backward(op):
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        _TLS.bwd[ctx.key] = ctx.lease
        try:
            dx, dz = op(dout, out, ctx.key, ctx.need_dx, ctx.need_dz, ctx.need_w)
        finally:
            _TLS.bwd.pop(ctx.key, None)
        _after_backward(ctx)
        # parameter gradients do not travel through autograd: the weight-gradient kernels accumulate them in packed form
        # over all backward sweeps of an update and the program scatters them into .grad (WeightBank.flush_wgrad)
        return (dx if ctx.need_dx else None), (dz.view(ctx.z_shape) if ctx.need_dz else None), [None] * ctx.n_params, None
    return backward


torch.library.register_autograd("pcgan::resnet_generator", _net_backward(resnet_generator_backward), setup_context=_net_setup)


# --------------------------------------------------------------------------- unet generator
@torch.library.custom_op("pcgan::unet_generator", mutates_args=())
def unet_generator(x: Tensor, z: Tensor, params: Sequence[Tensor], key: int) -> Tensor:
    """UnetGenerator.forward (models/networks.py:677-680)"""
    mod = _module(key)
    prog = mod._program(x.shape[0], x.shape[2])
    out, ws = prog.forward(x.contiguous().float(), z.contiguous().float().view(-1))
    _TLS.pending[key] = Lease(prog, ws)
    return out


@unet_generator.register_fake
def _(x, z, params, key):
    return x.new_empty((x.shape[0], _module(key).output_nc, x.shape[2], x.shape[3]), dtype=torch.float32)


torch.library.register_autograd("pcgan::unet_generator", _net_backward(resnet_generator_backward), setup_context=_net_setup)


# -------------------------------------------------------------------------- discriminator
@torch.library.custom_op("pcgan::nlayer_discriminator", mutates_args=())
def nlayer_discriminator(x: Tensor, z: Tensor, params: Sequence[Tensor], key: int) -> Tensor:
    mod = _module(key)
    prog = mod._program(x.shape[0], x.shape[2])
    out, ws = prog.forward(x.contiguous().float(), z.contiguous().float().view(-1))
    _TLS.pending[key] = Lease(prog, ws)
    return out


@nlayer_discriminator.register_fake
def _(x, z, params, key):
    so = _module(key).out_size(x.shape[2])
    return x.new_empty((x.shape[0], 1, so, so), dtype=torch.float32)


@torch.library.custom_op("pcgan::nlayer_discriminator_backward", mutates_args=())
def nlayer_discriminator_backward(dout: Tensor, out: Tensor, key: int, need_dx: bool, need_dz: bool, need_w: bool) -> Tuple[Tensor, Tensor]:
    lease = _TLS.bwd[key]
    dx, dz = lease.prog.backward(lease.ws, out, dout.contiguous(), need_dx, need_w, need_dz)
    return (dx if dx is not None else _empty(dout)), (dz if dz is not None else _empty(dout))


@nlayer_discriminator_backward.register_fake
def _(dout, out, key, need_dx, need_dz, need_w):
    mod = _module(key)
    n = out.shape[0]
    s = mod.in_size(out.shape[2])
    return (out.new_empty((n, mod.input_nc_img, s, s)) if need_dx else out.new_empty(0)), (out.new_empty(n) if need_dz else out.new_empty(0))


torch.library.register_autograd("pcgan::nlayer_discriminator", _net_backward(nlayer_discriminator_backward), setup_context=_net_setup)


# -------------------------------------------------------------------------------- encoder
@torch.library.custom_op("pcgan::siamese_feature", mutates_args=())
def siamese_feature(x: Tensor, params: Sequence[Tensor], key: int) -> Tuple[Tensor, Tensor]:
    """(rating, log-variance); the second is empty unless the module has the noisy twin head."""
    mod = _module(key)
    prog = mod._program(x.shape[0], x.shape[2])
    outs, ws = prog.forward(x.contiguous().float())
    _TLS.pending[key] = Lease(prog, ws)
    return outs[0], (outs[1] if len(outs) > 1 else _empty(x))


@siamese_feature.register_fake
def _(x, params, key):
    y = x.new_empty((x.shape[0], 1, 1, 1), dtype=torch.float32)
    return y, (x.new_empty((x.shape[0], 1, 1, 1), dtype=torch.float32) if _module(key)._noisy else x.new_empty(0))


@torch.library.custom_op("pcgan::siamese_feature_backward", mutates_args=())
def siamese_feature_backward(gy: Tensor, glogvar: Tensor, key: int, need_dx: bool, need_w: bool) -> Tensor:
    """empty gy / glogvar = that head received no gradient"""
    lease = _TLS.bwd[key]
    gys = [gy if gy.numel() else None]
    if len(lease.prog.heads) > 1:
        gys.append(glogvar if glogvar.numel() else None)
    dx = lease.prog.backward(lease.ws, gys, need_dx, need_w)
    return dx if dx is not None else _empty(gy if gy.numel() else glogvar)


@siamese_feature_backward.register_fake
def _(gy, glogvar, key, need_dx, need_w):
    ref = gy if gy.numel() else glogvar
    s = _module(key)._last_size
    return ref.new_empty((ref.shape[0], 3, s, s)) if need_dx else ref.new_empty(0)


def _enc_setup(ctx, inputs, output):
    x, params, key = inputs
    _claim(ctx, key)
    ctx.key = key
    ctx.need_dx = ctx.needs_input_grad[0]
    ctx.need_w = any(p.requires_grad for p in params)
    ctx.n_params, ctx.dev = len(params), x.device
    ctx.set_materialize_grads(False)     # an unused head (logvar) then gets None instead of a zero gradient


def _enc_backward(ctx, gy, glv):
    e = torch.empty(0, device=ctx.dev)
    _TLS.bwd[ctx.key] = ctx.lease
    try:
        dx = siamese_feature_backward(gy if gy is not None else e, glv if glv is not None else e, ctx.key, ctx.need_dx, ctx.need_w)
    finally:
        _TLS.bwd.pop(ctx.key, None)
    _after_backward(ctx)
    return (dx if ctx.need_dx else None), [None] * ctx.n_params, None


torch.library.register_autograd("pcgan::siamese_feature", _enc_backward, setup_context=_enc_setup)


# ------------------------------------------------------------------- identity-preserving net
@torch.library.custom_op("pcgan::alexnet_feature", mutates_args=())
def alexnet_feature(x: Tensor, key: int) -> Tensor:
    """AlexNetFeature.forward (models/networks.py:1242-1248), pooling 'None': [N, 3, S, S] -> [N, 256, h, h] (frozen)"""
    mod = _module(key)
    prog = mod._program(x.shape[0], x.shape[2])
    out, ws = prog.forward(x.contiguous().float())
    _TLS.pending[key] = Lease(prog, ws)
    return out.contiguous()


def _alex_side(s):
    h = (s + 4 - 11) // 4 + 1
    for _ in range(3):
        h = (h - 3) // 2 + 1
    return h


@alexnet_feature.register_fake
def _(x, key):
    h = _alex_side(x.shape[2])
    return x.new_empty((x.shape[0], 256, h, h), dtype=torch.float32)


@torch.library.custom_op("pcgan::alexnet_feature_backward", mutates_args=())
def alexnet_feature_backward(dfeat: Tensor, key: int, size: int) -> Tensor:
    lease = _TLS.bwd[key]
    return lease.prog.backward(lease.ws, dfeat)


@alexnet_feature_backward.register_fake
def _(dfeat, key, size):
    return dfeat.new_empty((dfeat.shape[0], 3, size, size), dtype=torch.float32)


def _alex_setup(ctx, inputs, output):
    x, key = inputs
    _claim(ctx, key)
    ctx.key, ctx.size = key, x.shape[2]


def _alex_backward(ctx, dfeat):
    _TLS.bwd[ctx.key] = ctx.lease
    try:
        dx = alexnet_feature_backward(dfeat, ctx.key, ctx.size)
    finally:
        _TLS.bwd.pop(ctx.key, None)
    _after_backward(ctx)
    return dx, None


torch.library.register_autograd("pcgan::alexnet_feature", _alex_backward, setup_context=_alex_setup)


# --------------------------------------------------------------------------------- losses
@torch.library.custom_op("pcgan::reduce_loss", mutates_args=())
def reduce_loss(kind: int, pred: Tensor, target: Tensor, per_sample: int) -> Tensor:
    """mean over all elements of the loss `kind` (pcgan_loss_kind) between pred and target; target holds one value per
    element (per_sample = 0) or one per sample (per_sample = elements per sample)."""
    out = torch.zeros((), device=pred.device)
    ops.loss(kind, pred.contiguous().float(), target, per_sample=per_sample, loss_out=out)
    return out


@reduce_loss.register_fake
def _(kind, pred, target, per_sample):
    return pred.new_empty((), dtype=torch.float32)


@torch.library.custom_op("pcgan::reduce_loss_backward", mutates_args=())
def reduce_loss_backward(gout: Tensor, kind: int, pred: Tensor, target: Tensor, per_sample: int) -> Tensor:
    p = pred.contiguous().float()
    grad = torch.empty_like(p)
    ops.loss(kind, p, target, per_sample=per_sample, weight=1.0, weight_dev=gout.contiguous().float(), grad=grad)
    return grad


@reduce_loss_backward.register_fake
def _(gout, kind, pred, target, per_sample):
    return pred.new_empty(pred.shape, dtype=torch.float32)


def _loss_setup(ctx, inputs, output):
    kind, pred, target, per_sample = inputs
    ctx.kind, ctx.per_sample = kind, per_sample
    ctx.save_for_backward(pred, target)


def _loss_backward(ctx, gout):
    pred, target = ctx.saved_tensors
    return None, reduce_loss_backward(gout, ctx.kind, pred, target, ctx.per_sample), None, None


torch.library.register_autograd("pcgan::reduce_loss", _loss_backward, setup_context=_loss_setup)


# ------------------------------------------------------------------------------- upsample
@torch.library.custom_op("pcgan::upsample_bilinear_ac", mutates_args=())
def upsample_bilinear_ac(x: Tensor, size: int) -> Tensor:
    x = x.contiguous().float()
    out = torch.empty(x.shape[0], x.shape[1], size, size, device=x.device)
    ops.resize_nchw_fwd(x, out)
    return out


@upsample_bilinear_ac.register_fake
def _(x, size):
    return x.new_empty((x.shape[0], x.shape[1], size, size), dtype=torch.float32)


@torch.library.custom_op("pcgan::upsample_bilinear_ac_backward", mutates_args=())
def upsample_bilinear_ac_backward(g: Tensor, h: int, w: int) -> Tensor:
    gi = torch.empty(g.shape[0], g.shape[1], h, w, device=g.device)
    ops.resize_nchw_bwd(g.contiguous().float(), gi)
    return gi


@upsample_bilinear_ac_backward.register_fake
def _(g, h, w):
    return g.new_empty((g.shape[0], g.shape[1], h, w), dtype=torch.float32)


def _up_setup(ctx, inputs, output):
    ctx.hw = (inputs[0].shape[2], inputs[0].shape[3])


def _up_backward(ctx, g):
    return upsample_bilinear_ac_backward(g, ctx.hw[0], ctx.hw[1]), None


torch.library.register_autograd("pcgan::upsample_bilinear_ac", _up_backward, setup_context=_up_setup)
