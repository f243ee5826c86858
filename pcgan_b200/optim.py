"""Fused multi-tensor Adam behind torch.optim.Adam's interface.

The reference steps `torch.optim.Adam(net.parameters(), lr, betas=(beta1, 0.999))` once per network per
update (models/wsgan_emb_model.py:153-154, 451-461); with torch's multi-tensor implementation that is about a
dozen launches over 48 (G) or 13 (D) tensors.  `FusedAdam.step()` is two launches of libpcgan_kernels.so
(`pcgan_adam_batched`: the update of every tensor, then the one-thread step counter), reads the learning rate and
the step count from device memory (so it is CUDA-graph capturable and LR schedulers keep working: they write
`param_groups[i]["lr"]`, a float or a device tensor), and keeps the state under torch.optim.Adam's names
(`exp_avg`, `exp_avg_sq`, `step`), so `state_dict()` / `load_state_dict()` interchange with the stock optimizer.

Semantics are torch.optim.Adam's with weight_decay = 0, amsgrad = False, maximize = False (the only form the
reference uses); anything else raises.  Gradients must exist for every parameter at step() (`p.grad is None`
raises): the step walks a pointer table built once, and rebuilt only when a tensor moves.
"""
import torch

from . import ops


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *, maximize=False, **kw):
        if weight_decay != 0 or amsgrad or maximize:
            raise NotImplementedError("FusedAdam: weight_decay / amsgrad / maximize are not implemented (the reference uses none)")
        kw.pop("capturable", None)
        kw.pop("foreach", None)
        kw.pop("fused", None)
        super().__init__(params, lr=lr, betas=betas, eps=eps)
        self._groups = {}     # group index -> bound state

    # ------------------------------------------------------------------ state binding
    def _bind(self, gi, group):
        """Flat first / second moment arenas (each tensor padded to 16 bytes so the kernel's vector path applies), the
        shared device step counter and the pointer table of one parameter group.  Existing state (a loaded checkpoint)
        is carried over."""
        params = [p for p in group["params"] if p.requires_grad]
        if not params:
            return None
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam runs on CUDA tensors only (no CPU fallback)")
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam needs contiguous fp32 parameters")
            if p.grad is None:
                raise RuntimeError("FusedAdam.step(): a parameter has no gradient")
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                raise RuntimeError("FusedAdam needs contiguous fp32 gradients")
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        m = torch.zeros(total, dtype=torch.float32, device=dev)
        v = torch.zeros(total, dtype=torch.float32, device=dev)
        step = torch.zeros((), dtype=torch.float32, device=dev)
        steps_seen = set()
        for p, o in zip(params, offs):
            st = self.state[p]
            mv, vv = m[o:o + p.numel()].view_as(p), v[o:o + p.numel()].view_as(p)
            if "exp_avg" in st:
                mv.copy_(st["exp_avg"])
                vv.copy_(st["exp_avg_sq"])
                steps_seen.add(float(st["step"]))
            st["exp_avg"], st["exp_avg_sq"], st["step"] = mv, vv, step
        if len(steps_seen) > 1:
            raise RuntimeError("FusedAdam: parameters of one group are at different steps: %s" % sorted(steps_seen))
        if steps_seen:
            step.fill_(steps_seen.pop())
        b = dict(params=params, m=m, v=v, step=step, lr_dev=None, lr_val=None, table=None, ptrs=None, count=len(params), max_n=0)
        self._groups[gi] = b
        return b

    @staticmethod
    def _pointers(b):
        return [(p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0) for p in b["params"]]

    def _table(self, b):
        ptrs = self._pointers(b)
        if b["table"] is None or ptrs != b["ptrs"]:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdam: a parameter or gradient moved after the graph's warm-up steps")
            for p in b["params"]:
                if p.grad is None:
                    raise RuntimeError("FusedAdam.step(): a parameter has no gradient")
            st = self.state
            b["table"], b["max_n"] = ops.adam_table([(p, p.grad, st[p]["exp_avg"], st[p]["exp_avg_sq"]) for p in b["params"]],
                                                    b["params"][0].device)
            b["ptrs"] = ptrs
        return b["table"]

    def _lr(self, b, lr):
        """Device scalar holding the group's learning rate: the user's tensor as is, a float through a cached copy."""
        if isinstance(lr, torch.Tensor):
            if lr.device != b["step"].device or lr.dtype != torch.float32:
                raise RuntimeError("FusedAdam: a tensor lr must be an fp32 scalar on the parameters' device")
            return lr
        if b["lr_dev"] is None:
            b["lr_dev"] = torch.zeros((), dtype=torch.float32, device=b["step"].device)
        if b["lr_val"] != float(lr):
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdam: pass lr as a device tensor to change it under CUDA-graph replay")
            b["lr_dev"].fill_(float(lr))
            b["lr_val"] = float(lr)
        return b["lr_dev"]

    # ------------------------------------------------------------------ torch.optim interface
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            b = self._groups.get(gi) or self._bind(gi, group)
            if b is None:
                continue
            table = self._table(b)
            beta1, beta2 = group["betas"]
            ops.adam_batched(table, b["count"], b["max_n"], self._lr(b, group["lr"]), float(beta1), float(beta2), float(group["eps"]), b["step"])
            # the kernel wrote through raw pointers: tell autograd (and the engine's packed-operand cache, which keys on
            # Tensor._version) that the parameters changed
            torch.autograd.graph.increment_version(b["params"])
        return loss

    def state_dict(self):
        """torch.optim.Adam's layout with private copies: the moments are views of this optimizer's arenas and the step
        counter is one tensor shared by a whole group, neither of which another optimizer may alias."""
        sd = super().state_dict()
        sd["state"] = {k: {kk: (vv.clone() if isinstance(vv, torch.Tensor) else vv) for kk, vv in st.items()}
                       for k, st in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        """Loaded moments are copied into fresh arenas at the next step(); a CUDA graph captured before the load still
        points at the old ones and must be captured again (the reference never saves optimizer state: base_model.py:96-107)."""
        super().load_state_dict(state_dict)
        self._groups = {}     # re-bind from the loaded tensors at the next step

    def zero_grad(self, set_to_none=False):
        """Gradients are cleared in place by default: the step's pointer table (and a captured graph) stay valid."""
        super().zero_grad(set_to_none=set_to_none)
