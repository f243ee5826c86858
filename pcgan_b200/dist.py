"""Data parallelism for the wsgan_emb step: one process per GPU, gradients averaged with one
all-reduce per network over a flat fp32 buffer (NCCL over NVLink / NVSwitch on the box; gloo on
CPU in the tests).  Replaces the reference's single-process nn.DataParallel
(models/networks.py:96-102), which re-broadcasts parameters and gathers outputs on every call.

Every parameter's .grad is a view into the flat buffer, so the kernels accumulate weight
gradients straight into communication memory: no packing copy before the collective and no
unpacking after it.  BatchNorm statistics stay per rank, as under nn.DataParallel.
"""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, params, group=None):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("GradSync needs at least one parameter")
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()
        self.group = group

    def zero(self):
        """optimizer.zero_grad(): clears the flat buffer and (re)attaches the views as .grad"""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def all_reduce(self):
        """Average the gradients over all ranks (mean-reduced losses at the global batch: SURVEY §5.8)."""
        ws = self.world_size()
        if ws == 1:
            return
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(ws)


def shard_batch(global_batch, rank, world_size):
    """Contiguous per-rank slice of a global batch (nn.DataParallel's scatter along dim 0)."""
    per = global_batch // world_size
    if per * world_size != global_batch:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world_size))
    return slice(rank * per, (rank + 1) * per)
