"""torch.library.opcheck on the pure operators of the path (schema, fake kernel, autograd registration, AOT dispatch)."""
import pytest
import torch

from pcgan_b200 import _lib as L
from pcgan_b200 import custom_ops  # noqa: F401  (registers the operators)

pytestmark = pytest.mark.gpu


def test_opcheck_reduce_loss_and_upsample():
    p = torch.rand(4, 1, 6, 6, device="cuda").clamp(0.05, 0.95).requires_grad_(True)
    t = torch.tensor([1.0, 0.0, 1.0, 0.0], device="cuda")
    for kind in (L.LOSS_BCE, L.LOSS_MSE):
        torch.library.opcheck(torch.ops.pcgan.reduce_loss.default, (kind, p, t, 36))
    a = torch.randn(2, 3, 5, 5, device="cuda", requires_grad=True)
    b = torch.randn(2, 3, 5, 5, device="cuda")
    torch.library.opcheck(torch.ops.pcgan.reduce_loss.default, (L.LOSS_L1, a, b, 0))
    x = torch.randn(2, 3, 16, 16, device="cuda", requires_grad=True)
    torch.library.opcheck(torch.ops.pcgan.upsample_bilinear_ac.default, (x, 28))
    torch.library.opcheck(torch.ops.pcgan.reduce_loss_backward.default, (torch.ones((), device="cuda"), L.LOSS_MSE, p.detach(), t, 36))
