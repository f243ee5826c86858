"""Data parallelism for the wsgan_emb step: one process per GPU, gradients averaged over a flat fp32 buffer per network
(NCCL over NVLink / NVSwitch on the box; gloo on CPU in the tests).  Replaces the reference's single-process
nn.DataParallel (models/networks.py:96-102), which re-broadcasts parameters and gathers outputs on every call.

Every parameter's .grad is a view into the flat buffer, so the kernels accumulate weight gradients straight into
communication memory: no packing copy before the collective and no unpacking after it.  The buffer is cut into
reverse-layer buckets (contiguous ranges, last layers first): as soon as the backward sweep has produced the last
gradient of a bucket, its all-reduce is issued asynchronously (NCCL runs it on its own stream, fenced by events to the
kernels that wrote the bucket) while the sweep goes on with the earlier layers; finish() makes the compute stream wait
for the collectives right before the optimizer step.  All of it is capturable into the step's CUDA graph.
BatchNorm statistics stay per rank, as under nn.DataParallel.
"""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, params, group=None, bucket_bytes=12 << 20, layers=None):
        """layers: optional list of parameter lists in FORWARD layer order (their concatenation must be `params` in
        order); buckets are then unions of whole layers, built from the last layer backwards, of about bucket_bytes."""
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("GradSync needs at least one parameter")
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        self.views, self.offset, o = [], {}, 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            self.offset[id(p)] = (o, o + p.numel())
            o += p.numel()
        self.group = group
        self.buckets = self._make_buckets(layers, bucket_bytes)      # [(start, end)] in issue order (last layers first)
        self.bucket_of = {}                                          # id(param) -> bucket index
        for b, (s, e) in enumerate(self.buckets):
            for p in self.params:
                ps, pe = self.offset[id(p)]
                if ps >= s and pe <= e:
                    self.bucket_of[id(p)] = b
        self._issued = [False] * len(self.buckets)
        self._works = []

    def _make_buckets(self, layers, bucket_bytes):
        if not layers:
            return [(0, self.flat.numel())]
        flat_ids = [id(p) for p in self.params]
        if [id(p) for layer in layers for p in layer] != flat_ids:
            raise ValueError("GradSync: `layers` must list every parameter exactly once, in order")
        out, end, size = [], self.flat.numel(), 0
        for layer in reversed(layers):
            if not layer:
                continue
            size += sum(p.numel() for p in layer) * self.flat.element_size()
            start = self.offset[id(layer[0])][0]
            if size >= bucket_bytes:
                out.append((start, end))
                end, size = start, 0
        if end > 0:
            out.append((0, end))
        return out

    # ------------------------------------------------------------------ step protocol
    def zero(self):
        """optimizer.zero_grad(): clears the flat buffer and (re)attaches the views as .grad"""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v
        self._issued = [False] * len(self.buckets)
        self._works = []

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _reduce(self, t, async_op):
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        self._scale = getattr(self, "_scale", []) + [t]
        return w

    def bucket_ready(self, b):
        """The gradients of bucket b are final: start averaging them over the ranks (returns at once)."""
        if self._issued[b]:
            return
        self._issued[b] = True
        if self.world_size() == 1:
            return
        s, e = self.buckets[b]
        self._works.append(self._reduce(self.flat[s:e], True))

    def start(self):
        """Issue every bucket that has not been issued yet (asynchronously)."""
        for b in range(len(self.buckets)):
            self.bucket_ready(b)

    def finish(self):
        """The current stream waits for the outstanding all-reduces (call before the optimizer step)."""
        self.start()
        for w in self._works:
            if w is not None:
                w.wait()
        self._works = []
        scale = getattr(self, "_scale", None)
        if scale:      # backends without an averaging reduction (gloo)
            ws = self.world_size()
            for t in scale:
                t.div_(ws)
            self._scale = []

    def all_reduce(self):
        """Average the gradients over all ranks (mean-reduced losses at the global batch: SURVEY §5.8), blocking form."""
        if self.world_size() == 1:
            return
        self.finish()


def shard_batch(global_batch, rank, world_size):
    """Contiguous per-rank slice of a global batch (nn.DataParallel's scatter along dim 0)."""
    per = global_batch // world_size
    if per * world_size != global_batch:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world_size))
    return slice(rank * per, (rank + 1) * per)
