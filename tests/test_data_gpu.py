"""The augment kernel (pcgan_augment) against the reference's transform stack (data/base_dataset.py:24-64):
transforms.Resize([load, load], BICUBIC) -> crop -> flip -> ToTensor -> Normalize(0.5, 0.5) by PIL / torchvision on the
same decoded images with the same crop origins and flips; and the loader end to end on files."""
import numpy as np
import pytest
import torch

from pcgan_b200 import data as D

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _reference(arr, load, fine, crop, flip):
    from PIL import Image
    im = Image.fromarray(arr).resize((load, load), Image.BICUBIC)
    t = torch.from_numpy(np.asarray(im, dtype=np.uint8).copy()).permute(2, 0, 1).float() / 255.0
    cy, cx = crop
    t = t[:, cy:cy + fine, cx:cx + fine]
    if flip:
        t = t.flip(2)
    return (t - 0.5) / 0.5


@pytest.mark.parametrize("hw,load,fine", [((200, 200), 128, 128), ((300, 180), 143, 128), ((96, 120), 128, 112), ((128, 128), 128, 128)])
def test_augment_matches_pil_transform_stack(hw, load, fine):
    rng = np.random.RandomState(hw[0] + load)
    n = 5
    # smooth + noisy content so that both the antialiasing and the rounding are exercised
    arrs = []
    for i in range(n):
        yy, xx = np.mgrid[0:hw[0], 0:hw[1]]
        base = 127 + 100 * np.sin(yy / (7.0 + i))[..., None] * np.cos(xx / (5.0 + i))[..., None]
        arrs.append(np.clip(base + rng.randint(-40, 40, size=(hw[0], hw[1], 3)), 0, 255).astype(np.uint8))
    crops = [(int(rng.randint(0, load - fine + 1)), int(rng.randint(0, load - fine + 1))) for _ in range(n)]
    flips = [bool(rng.randint(0, 2)) for _ in range(n)]
    out = D.augment([torch.from_numpy(a).to(DEV) for a in arrs], crops, flips, load, fine, DEV)
    torch.cuda.synchronize()
    ref = torch.stack([_reference(a, load, fine, c, f) for a, c, f in zip(arrs, crops, flips)])
    diff = (out.cpu() - ref).abs()
    lsb = 2.0 / 255.0
    frac_exact = float((diff < 1e-6).float().mean())
    print("augment %s -> %d -> %d: max diff %.4f (1 LSB = %.4f), %.2f%% of the pixels bit-exact" % (hw, load, fine, float(diff.max()), lsb, 100 * frac_exact))
    # PIL accumulates with 22-bit fixed-point coefficients, the kernel in fp32: a rounding tie may fall the other way
    assert float(diff.max()) <= lsb + 1e-6 and frac_exact > 0.97


def test_loader_shards_and_prefetches(tmp_path):
    from PIL import Image
    rng = np.random.RandomState(0)
    lines = []
    for i in range(12):
        for tag in "ab":
            Image.fromarray(rng.randint(0, 255, size=(150, 150, 3)).astype(np.uint8)).save(tmp_path / ("%s%d.png" % (tag, i)))
        lines.append("a%d.png b%d.png %d" % (i, i, i % 3))
    (tmp_path / "pairs.txt").write_text("\n".join(lines) + "\n")
    seen = []
    for rank in (0, 1):
        pl = D.PairList(str(tmp_path / "pairs.txt"), str(tmp_path))
        loader = D.GpuPairLoader(pl, 4, 128, 128, DEV, rank=rank, world_size=2, serial_batches=True, seed=3)
        assert len(loader) == 2
        for batch in loader:
            # 12 pairs, global batch 8: one full batch of 4 per rank, then the short last global batch (4 pairs) split 2 + 2
            assert tuple(batch["A"].shape) in ((4, 3, 128, 128), (2, 3, 128, 128)) and batch["A"].is_cuda and batch["B"].dtype == torch.float32
            assert float(batch["A"].abs().max()) <= 1.0 and batch["label"].dtype == torch.int64
            seen += batch["A_paths"]
    assert sorted(seen) == sorted(str(tmp_path / ("a%d.png" % i)) for i in range(12))      # the two ranks cover every pair exactly once


def test_device_prefetcher_copies_ahead_on_a_side_stream():
    """Every batch arrives on the device bit-identical and in order while the consumer keeps the GPU busy; non-tensor
    entries pass through."""
    from pcgan_b200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(5)
    host = [{"A": torch.rand(8, 3, 64, 64, generator=g).pin_memory(), "label": torch.randint(0, 3, (8,), generator=g).pin_memory(),
             "B_paths": ["x%d" % i]} for i in range(6)]
    busy = torch.randn(2048, 2048, device="cuda")
    seen = []
    for i, b in enumerate(DevicePrefetcher(iter(host), "cuda")):
        assert b["A"].is_cuda and b["label"].is_cuda and b["B_paths"] == ["x%d" % i]
        busy = busy @ busy * 1e-3                     # work on the current stream that overlaps the next copy
        seen.append((b["A"].clone(), b["label"].clone()))
    torch.cuda.synchronize()
    assert len(seen) == 6
    for (a, l), h in zip(seen, host):
        assert torch.equal(a.cpu(), h["A"]) and torch.equal(l.cpu(), h["label"])
