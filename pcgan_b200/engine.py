"""Execution engine shared by the three networks of the wsgan_emb step.

A network at a fixed input geometry is a static program over padded NHWC bf16 buffers:
  ConvRT   one convolution layer: packed bf16 operands + tcgen05 igemm plans for forward,
           data-gradient and weight-gradient (pcgan_b200/conv.py plans them)
  NormRT   one normalisation layer (instance or batch statistics come out of the conv
           epilogue; finalize + apply + backward are the HBM-bound kernels)
  Workspace all buffers one forward/backward pair needs, pooled per geometry so that a
           training step allocates nothing after warm-up
Only torch tensors (device memory) and the C ABI are used; there is no fallback path.
"""
import os
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import conv as CV
from . import ops
from .plan import Geom, OutMap, SLACK


# PCGAN_BUCKET_OVERLAP=0: no bucket is handed to the collective inside the backward sweep; every bucket is issued (still
# asynchronously) when the sweep has finished (measurement switch)
BUCKET_OVERLAP = os.environ.get("PCGAN_BUCKET_OVERLAP", "1") != "0"
PHASE_STREAMS = os.environ.get("PCGAN_PHASE_STREAMS", "0") != "0"    # launch the sub-pixel phases of a strided data gradient / transposed convolution on forked streams
_SIDE = {}


def _fan_out(launchers):
    """Run independent launches (the phases of one convolution write disjoint outputs) concurrently: the first on the
    current stream, the others on side streams forked from it and joined back.  Under graph capture the fork / join
    events become parallel branches of the graph."""
    if len(launchers) == 1 or not PHASE_STREAMS:
        for fn in launchers:
            fn()
        return
    main = torch.cuda.current_stream()
    dev = main.device
    pool = _SIDE.setdefault(dev, [])
    while len(pool) < len(launchers) - 1:
        pool.append(torch.cuda.Stream(device=dev))
    fork = torch.cuda.Event()
    fork.record(main)
    joins = []
    for i, fn in enumerate(launchers):
        if i == 0:
            fn()
            continue
        side = pool[i - 1]
        side.wait_event(fork)
        with torch.cuda.stream(side):
            fn()
            ev = torch.cuda.Event()
            ev.record(side)
        joins.append(ev)
    for ev in joins:
        main.wait_event(ev)


class ConvRT:
    """Runtime of one convolution at a fixed geometry.

    weight: the reference-layout fp32 master parameter (OIHW; IOHW if transposed).
    xg: geometry of the input buffer; out: OutMap of the forward output.
    dyg: geometry of the output-gradient buffer used by the backward plans."""

    def __init__(self, name, weight, bias, xg: Geom, stride, cp, out: OutMap, *, transposed=False, output_padding=0,
                 act=L.ACT_NONE, act_slope=0.0, stats=False, per_sample_stats=False, dyg: Optional[Geom] = None,
                 dx_out: Optional[OutMap] = None, full_padded=False, want_dgrad=True, want_wgrad=True, tf32=False, stats_div=0,
                 want_fwd=True, trainable=True):
        """trainable=False: a helper runtime over a view of another convolution's weight (never asked for a weight gradient).
        tf32=True: the activation / gradient buffers passed to forward / backward_* hold fp32 (same padded NHWC
        geometry), the packed weights are fp32 rounded to TF32, and the kernels issue tcgen05.mma.kind::tf32 (per-layer
        error 3e-4 .. 8e-4 instead of 2.4e-3 in bf16).  The network programs of this package run their normalisation
        kernels on bf16 buffers, so they build their convolutions with tf32=False; a TF32 runtime is what a caller with
        fp32 activations (or a per-layer precision study) uses."""
        self.name, self.weight, self.bias = name, weight, bias
        # the convolution this runtime was planned for (tests/test_bench_geometry_gpu.py replays every plan against F.conv2d)
        self.geometry = dict(xg=xg, stride=stride, cp=cp, out=out, transposed=transposed, output_padding=output_padding,
                             act=act, act_slope=act_slope, stats=stats, per_sample_stats=per_sample_stats, dyg=dyg,
                             dx_out=dx_out, full_padded=full_padded)
        self.bank = None
        self.trainable = bool(trainable)
        self.tf32 = bool(tf32)
        wdt = torch.float32 if tf32 else torch.bfloat16
        dev = weight.device
        self.dev = dev
        shape = tuple(weight.shape)
        self._wver = None
        self.fwd = []
        for sp, wm in (CV.conv_fwd_plans(shape, xg, stride, cp, out, transposed=transposed, output_padding=output_padding,
                                         act=act, act_slope=act_slope, stats=stats, per_sample_stats=per_sample_stats,
                                         note=name + ".fwd", tf32=tf32) if want_fwd else []):
            sp.stats_div = int(stats_div)      # > 1: per-sample statistics plans emit one statistic per run of stats_div samples
            self.fwd.append((ops.Igemm(sp), wm.to(dev), torch.zeros(sp.b_rows * sp.b_k + 64, dtype=wdt, device=dev)))
        self.dgrad = []
        if want_dgrad and dyg is not None and dx_out is not None:
            for sp, wm in CV.conv_dgrad_plans(shape, dyg, xg, stride, cp, dx_out, transposed=transposed,
                                              full_padded=full_padded, note=name + ".dgrad", tf32=tf32):
                self.dgrad.append((ops.Igemm(sp), wm.to(dev), torch.zeros(sp.b_rows * sp.b_k + 64, dtype=wdt, device=dev)))
        self.wgrad = None
        self._wg_args = (shape, dyg, xg, stride, cp, transposed, name)
        if want_wgrad and dyg is not None:
            self.ensure_wgrad()
        self.transposed = transposed
        self.flops_fwd = sum(g.spec.flops for g, _, _ in self.fwd)

    def ensure_wgrad(self):
        """Weight-gradient plan and its packed fp32 accumulator, built on first need (a frozen network never pays)."""
        if self.wgrad is None:
            shape, dyg, xg, stride, cp, transposed, name = self._wg_args
            if dyg is None:
                raise L.PcganError("%s: no output-gradient geometry was planned, cannot build a weight gradient" % name)
            sp, wm = CV.conv_wgrad_plan(shape, dyg, xg, stride, cp, transposed=transposed, note=name + ".wgrad", tf32=self.tf32)
            self.wgrad = (ops.Igemm(sp), wm.to(self.dev), torch.zeros(sp.b_rows * sp.b_k, dtype=torch.float32, device=self.dev))
        return self.wgrad

    def pack(self):
        """Refresh the packed bf16 operands when the master weight changed (optimizer step / load_state_dict)."""
        v = (self.weight._version, self.weight.data_ptr())
        if v == self._wver:
            return
        w = self.weight.detach()
        gather = ops.gather_tf32 if self.tf32 else ops.gather_cast_bf16
        for _, wm, buf in self.fwd:
            gather(w, wm, buf)
        for _, wm, buf in self.dgrad:
            gather(w, wm, buf)
        self._wver = v

    def forward(self, xbuf, out, stats=None):
        self.pack()
        b = self.bias.detach() if self.bias is not None else None
        _fan_out([(lambda g=g, wbuf=wbuf: g.run(xbuf, wbuf, out, b, stats)) for g, _, wbuf in self.fwd])

    def backward_data(self, dybuf, dxout):
        self.pack()
        _fan_out([(lambda g=g, wbuf=wbuf: g.run(dybuf, wbuf, dxout)) for g, _, wbuf in self.dgrad])

    def backward_weight(self, dybuf, xbuf):
        """Accumulates into weight.grad (allocated on first use).  Under a WeightBank in deferred mode the packed
        gradient keeps accumulating across backward passes and the bank scatters all layers in one launch."""
        g, wm, packed = self.ensure_wgrad()
        deferred = self.bank is not None and self.bank.deferred
        if not deferred:
            packed.zero_()
        if self.transposed or g.spec.swap_operands:
            g.run(xbuf, dybuf, packed)     # M side = input activations, N side = dY
        else:
            g.run(dybuf, xbuf, packed)
        if deferred:
            self.bank.dirty = True
            self.bank.layer_done(self)
            return
        if self.weight.grad is None:
            self.weight.grad = torch.zeros_like(self.weight)
        ops.scatter_f32(packed, wm, self.weight.grad, accumulate=True)


class WeightBank:
    """All convolutions of one program: their packed bf16 operands are refreshed by ONE gather launch when any master
    weight changed (optimizer step, load_state_dict), and — in deferred mode, which the training drivers switch on — their
    packed fp32 weight gradients live in one arena that is cleared by one memset per optimizer step and scattered into the
    .grad tensors by ONE launch per backward sweep (instead of a memset and a scatter per layer and pass)."""

    def __init__(self, convs, dev):
        self.convs, self.dev = list(convs), dev
        if any(c.tf32 for c in self.convs):
            raise L.PcganError("WeightBank batches bf16 operands; TF32 convolutions repack themselves (ConvRT.pack)")
        for c in self.convs:
            c.bank = self
        self.deferred, self.dirty = False, False
        self._gather = None          # (pointer signature, table, count, max_n)
        self._versions = None
        self._scatter = {}           # bucket (or None = all) -> (pointer signature, table, count, max_n)
        self._arena = None
        # overlap of the gradient all-reduce with the backward sweep (pcgan_b200.dist.GradSync): `pending` counts the
        # forward calls of this program whose weight-gradient backward is still to come; during the LAST of them every
        # bucket of the flat gradient buffer is scattered and handed to the collective as soon as its layers are done
        self.sync, self.pending, self.final = None, 0, False
        self._left, self._flushed = {}, set()

    # ------------------------------------------------------------ operands
    def ensure_packed(self):
        ver = tuple((c.weight._version, c.weight.data_ptr()) for c in self.convs)
        if ver == self._versions:
            return
        items = [(c.weight.detach(), wm, buf) for c in self.convs for _, wm, buf in c.fwd + c.dgrad]
        sig = tuple(t.data_ptr() for it in items for t in it)
        if self._gather is None or self._gather[0] != sig:
            table, max_n = ops.batch_table(items, self.dev)
            self._gather = (sig, table, len(items), max_n)
        _, table, count, max_n = self._gather
        ops.gather_cast_bf16_batched(table, count, max_n)
        self._versions = ver
        for c, v in zip(self.convs, ver):
            c._wver = v

    # ----------------------------------------------------- weight gradients
    def enable_deferred(self):
        """Move every convolution's packed gradient into one flat fp32 arena (builds the weight-gradient plans)."""
        if self.deferred:
            return
        trainable = [c for c in self.convs if c._wg_args[1] is not None and c.trainable]
        sizes = []
        for c in trainable:
            g, wm, packed = c.ensure_wgrad()
            sizes.append((packed.numel() + 3) // 4 * 4)
        self._arena = torch.zeros(max(sum(sizes), 4), dtype=torch.float32, device=self.dev)
        off = 0
        for c, n in zip(trainable, sizes):
            g, wm, packed = c.wgrad
            c.wgrad = (g, wm, self._arena[off:off + packed.numel()])
            off += n
        self._trainable = trainable
        self.deferred = True

    def zero_wgrad(self):
        if self.deferred:
            self._arena.zero_()
            self.dirty = False
        # `pending` is NOT reset here: the forward calls of the update may already have run (the generator's do, in
        # WSGANEmbModel.forward()); flush_wgrad() resets it at the end of the update
        self.final = False
        self._flushed = set()

    def attach_sync(self, sync):
        """sync: the GradSync whose flat buffer holds the .grad views of this program's parameters (None: no overlap)."""
        self.sync = sync

    def begin_backward(self):
        """Called by the program at the start of a backward sweep that produces weight gradients."""
        self.pending -= 1
        self.final = bool(BUCKET_OVERLAP and self.deferred and self.sync is not None and self.pending == 0 and self.sync.world_size() > 1)
        if self.final:
            self._left = {}
            for c in self._trainable:
                if c.weight.requires_grad:
                    b = self.sync.bucket_of[id(c.weight)]
                    self._left[b] = self._left.get(b, 0) + 1

    def layer_done(self, conv):
        """The weight gradient of `conv` has been launched; in the final sweep, a bucket whose layers are all done is
        scattered into the flat buffer and its all-reduce starts while the sweep continues."""
        if not self.final:
            return
        b = self.sync.bucket_of[id(conv.weight)]
        self._left[b] -= 1
        if self._left[b] == 0:
            self._flush(b)
            self.sync.bucket_ready(b)

    def _flush(self, bucket):
        live = [c for c in self._trainable if c.weight.requires_grad]
        if bucket is not None:
            convs = [c for c in live if self.sync.bucket_of[id(c.weight)] == bucket]
            self._flushed.add(bucket)
        elif self._flushed:      # what the buckets of the final sweep have not scattered already
            convs = [c for c in live if self.sync.bucket_of[id(c.weight)] not in self._flushed]
        else:
            convs = live
        if not convs:
            return
        for c in convs:
            if c.weight.grad is None:
                c.weight.grad = torch.zeros_like(c.weight)
        items = [(c.wgrad[2], c.wgrad[1], c.weight.grad) for c in convs]
        sig = tuple(t.data_ptr() for it in items for t in it)
        key = bucket if bucket is not None else ("rest", tuple(sorted(self._flushed)))
        cached = self._scatter.get(key)
        if cached is None or cached[0] != sig:
            table, max_n = ops.batch_table(items, self.dev)
            cached = self._scatter[key] = (sig, table, len(items), max_n)
        _, table, count, max_n = cached
        ops.scatter_f32_batched(table, count, max_n, accumulate=True)

    def flush_wgrad(self):
        """Scatter-accumulate the packed gradients of every trainable convolution (that no bucket has flushed yet) into
        its weight.grad."""
        self.pending, self.final = 0, False
        if not (self.deferred and self.dirty):
            return
        self._flush(None)
        self.dirty = False


class Arena:
    """Flat fp32 device buffer handed out in slices: every accumulator of a workspace (forward statistics, backward
    sums) lives in one, so a pass clears all of them with a single memset instead of one per layer."""

    def __init__(self, dev):
        self.dev, self.size, self.flat, self.pending = dev, 0, None, []

    def take(self, shape):
        n = 1
        for d in shape:
            n *= d
        n = (n + 3) // 4 * 4          # keep every slice 16-byte aligned
        holder = _Slice(self, self.size, tuple(shape))
        self.size += n
        self.pending.append(holder)
        return holder

    def finalize(self):
        self.flat = torch.zeros(max(self.size, 4), device=self.dev)
        for h in self.pending:
            n = 1
            for d in h.shape:
                n *= d
            h.t = self.flat[h.off:h.off + n].view(h.shape)
        self.pending = []

    def zero(self):
        self.flat.zero_()


class _Slice:
    def __init__(self, arena, off, shape):
        self.arena, self.off, self.shape, self.t = arena, off, shape, None


class RunningStats:
    """The running-statistics updates (and num_batches_tracked increments) of one network pass, issued as one launch at
    the end of the pass: layers register in forward order; the device table is built once per workspace."""

    def __init__(self, dev):
        self.dev, self.items, self.table = dev, [], None

    def begin(self):
        self.items = []

    def add(self, ns, count, running_mean, running_var, nbt=None, momentum=0.1, sequential=False):
        """sequential: the groups of ns are successive BatchNorm batches (one momentum step each, in order), not the
        samples of one InstanceNorm pass (whose instance statistics are averaged)."""
        self.items.append((ns.stats, running_mean, running_var, nbt, ns.stats.shape[0], ns.c, count, momentum, int(sequential)))

    def flush(self):
        if not self.items:
            return
        sig = tuple((it[0].data_ptr(), 0 if it[1] is None else it[1].data_ptr(), 0 if it[3] is None else it[3].data_ptr(), it[6], it[8])
                    for it in self.items)
        if self.table is None or self.table[0] != sig:
            t, max_c = ops.running_table(self.items, self.dev)
            self.table = (sig, t, len(self.items), max_c)
        _, t, n, max_c = self.table
        ops.norm_running_batched(t, n, max_c)
        self.items = []


class NormState:
    """Per-call statistics of one normalisation layer: [groups][C] each (groups = N for instance norm, 1 for batch norm).
    With arenas, `stats` / `sums` are slices of the workspace's accumulator arenas (cleared once per pass by the program);
    without, they are private tensors the caller clears."""

    def __init__(self, groups, c, dev, stats_arena=None, sums_arena=None, stats_groups=None):
        """stats_groups > groups: the convolution emits per-sample statistics that the finalize kernel folds (with the
        channel-dropout mask) into one batch statistic."""
        self.groups, self.c = groups, c
        self.stats_groups = stats_groups or groups
        self.affine = True
        self._stats = stats_arena.take((self.stats_groups, c, 2)) if stats_arena is not None else None
        self._sums = sums_arena.take((groups, c, 2)) if sums_arena is not None else None
        self._own_stats = torch.zeros(self.stats_groups, c, 2, device=dev) if stats_arena is None else None
        self._own_sums = torch.zeros(groups, c, 2, device=dev) if sums_arena is None else None
        self.mean = torch.empty(groups, c, device=dev)
        self.rstd = torch.empty(groups, c, device=dev)
        self.scale = torch.empty(groups, c, device=dev)
        self.shift = torch.empty(groups, c, device=dev)
        self.fused = None     # (count, gamma, beta): the consuming norm_apply finalizes the statistics itself

    def apply_kw(self, eps=1e-5):
        """Arguments of ops.norm_apply that describe this layer's normalisation."""
        if self.fused is not None:
            count, gamma, beta = self.fused
            return dict(groups=self.groups, stats=self.stats, count=count, eps=eps, gamma=gamma, beta=beta, mean_out=self.mean,
                        rstd_out=self.rstd, scale_out=self.scale, shift_out=self.shift)
        return dict(groups=self.groups, scale=self.scale, shift=self.shift)

    @property
    def stats(self):
        return self._own_stats if self._stats is None else self._stats.t

    @property
    def sums(self):
        return self._own_sums if self._sums is None else self._sums.t

    @property
    def pooled(self):
        return self._stats is not None


def accumulate_grad(p, g):
    """p.grad += g (allocating on first use); parameters that are frozen are skipped by the callers."""
    if p.grad is None:
        p.grad = g.clone()
    else:
        p.grad.add_(g)


class Pool:
    """Workspaces keyed by geometry; a network call takes one and gives it back when its backward has run
    (or right away when no gradient is needed)."""

    def __init__(self, factory):
        self.factory = factory
        self.free: Dict[tuple, list] = {}

    def take(self, key):
        lst = self.free.setdefault(key, [])
        return lst.pop() if lst else self.factory(key)

    def give(self, key, ws):
        self.free.setdefault(key, []).append(ws)


def zeros_act(g: Geom, dev):
    return torch.zeros(g.numel + SLACK, dtype=torch.bfloat16, device=dev)
