"""TEST INFRASTRUCTURE ONLY — plain-torch restatement of the buffer layout the kernels use
(padded NHWC bf16, see include/pcgan_kernels.h) so tests can build inputs and read outputs."""
import torch
import torch.nn.functional as F


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


def to_padded_nhwc(x, pad, halo="zero", c_buf=None, slack=512):
    """x: [N, C, H, W] float -> flat float32 tensor of bf16-rounded values laid out [N][H+2p][W+2p][c_buf] (+slack)."""
    n, c, h, w = x.shape
    c_buf = c_buf or c
    if pad > 0:
        x = F.pad(x, (pad,) * 4, mode="reflect" if halo == "reflect" else "constant")
    if c_buf > c:
        x = torch.cat([x, x.new_zeros(n, c_buf - c, h + 2 * pad, w + 2 * pad)], 1)
    flat = bf16_round(x.permute(0, 2, 3, 1).contiguous()).reshape(-1)
    return torch.cat([flat, flat.new_zeros(slack)])


def from_padded_nhwc(flat, n, h, w, c, pad, interior=True):
    t = flat[: n * (h + 2 * pad) * (w + 2 * pad) * c].view(n, h + 2 * pad, w + 2 * pad, c)
    if interior and pad > 0:
        t = t[:, pad:pad + h, pad:pad + w]
    return t.permute(0, 3, 1, 2).contiguous()
