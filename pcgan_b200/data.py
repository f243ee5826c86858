"""GPU-side input pipeline of the wsgan_emb trainer: what data/wsgan_emb_dataset.py:14-82,
data/base_dataset.py:24-64 and data/__init__.py:55-74 of phymhan/pc-gan do with PIL transforms in DataLoader workers.

    pairs  = PairList(opt.sourcefile_A, opt.dataroot)            # lines "A_path B_path label", label 0 / 1 / 2
    loader = GpuPairLoader(pairs, batch_size, load_size, fine_size, device, rank, world_size)
    for batch in loader:                                          # {'A', 'B': fp32 [B,3,S,S] on the device, 'label', paths}
        model.set_input(batch); model.optimize_parameters()

Host threads only decode the files (PIL) into pinned uint8 buffers; one kernel launch per image set
(`pcgan_augment`) does Resize(load, BICUBIC, PIL's antialiased two-pass form) -> RandomCrop(fine) ->
RandomHorizontalFlip -> ToTensor -> Normalize(0.5, 0.5) straight into the NCHW batch tensor, on a side stream, one
batch ahead of the step.  Every rank takes its contiguous slice of each global batch (nn.DataParallel's scatter along
dim 0), so the ranks never read the same pair.  As in the reference, the crop origin and the flip are drawn
independently for A and B, and the pair list is reshuffled once per epoch (the reference does it as a side effect of
__len__, wsgan_emb_dataset.py:72-79)."""
import ctypes as C
import os
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib as L


class PairList:
    """wsgan_emb_dataset.py:14-34 with --no_mixed_label_D off: (A_path, B_path, label) triples."""

    def __init__(self, sourcefile, dataroot="", max_dataset_size=float("inf")):
        with open(sourcefile, "r") as f:
            lines = [ln.rstrip("\n") for ln in f.readlines() if ln.strip()]
        self.root = dataroot
        self.items = []
        for ln in lines:
            a, b, lab = ln.split()[:3]
            self.items.append((os.path.join(dataroot, a), os.path.join(dataroot, b), int(lab)))
        self.size = int(min(len(self.items), max_dataset_size))

    def __len__(self):
        return self.size

    def shuffle(self, rng):
        rng.shuffle(self.items)


def rank_indices(num_items, batch_size, rank, world_size, drop_last=False):
    """Indices of this rank, batch by batch: global batch g = items [g*B*W, (g+1)*B*W), rank r takes its r-th run of B
    (nn.DataParallel's scatter).  The last global batch may be short (the reference's DataLoader has no drop_last): it is
    split as evenly as the remaining items allow."""
    out, gb = [], batch_size * world_size
    for start in range(0, num_items, gb):
        n = min(gb, num_items - start)
        if n < gb and drop_last:
            break
        per = -(-n // world_size)
        lo, hi = start + min(rank * per, n), start + min((rank + 1) * per, n)
        if hi > lo:
            out.append(list(range(lo, hi)))
    return out


def _decode(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.ascontiguousarray(np.asarray(im.convert("RGB"), dtype=np.uint8))


def augment(images, crops, flips, load_size, fine_size, device, out=None, stream=None):
    """images: list of uint8 HWC device tensors; crops: [(y, x)]; flips: [bool] -> fp32 [n, 3, fine, fine] in [-1, 1]."""
    n = len(images)
    table = (L.ImageItem * n)()
    for i, (im, (cy, cx), fl) in enumerate(zip(images, crops, flips)):
        if not im.is_cuda or im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or not im.is_contiguous():
            raise L.PcganError("augment: images must be contiguous uint8 [H, W, 3] CUDA tensors (no CPU path)")
        if not (0 <= cy <= load_size - fine_size and 0 <= cx <= load_size - fine_size):
            raise L.PcganError("augment: crop origin outside the resized image")
        table[i].src, table[i].h, table[i].w = im.data_ptr(), im.shape[0], im.shape[1]
        table[i].crop_y, table[i].crop_x, table[i].flip = int(cy), int(cx), int(bool(fl))
    raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(device, non_blocking=False)
    if out is None:
        out = torch.empty(n, 3, fine_size, fine_size, device=device)
    a = L.AugmentArgs(items=raw.data_ptr(), dst=out.data_ptr(), n=n, load=load_size, fine=fine_size)
    st = stream if stream is not None else torch.cuda.current_stream(device)
    L.check(L.load().pcgan_augment(C.byref(a), st.cuda_stream), "augment")
    out._pcgan_keepalive = (raw, images)
    return out


class GpuPairLoader:
    def __init__(self, pairs: PairList, batch_size, load_size, fine_size, device, rank=0, world_size=1, no_flip=False, serial_batches=False,
                 seed=None, workers=8):
        if not torch.cuda.is_available():
            raise L.PcganError("GpuPairLoader needs a CUDA device (no CPU path)")
        self.pairs, self.B, self.load, self.fine = pairs, int(batch_size), int(load_size), int(fine_size)
        self.device, self.rank, self.world = torch.device(device), rank, world_size
        self.no_flip, self.serial = no_flip, serial_batches
        # every rank shuffles the list with the SAME generator (so that the shards stay disjoint) and draws its crops /
        # flips from its own
        self.shared_rng = random.Random(0 if seed is None else seed)
        self.rng = random.Random((1 if seed is None else seed + 1) * 7919 + rank)
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.side = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(rank_indices(len(self.pairs), self.B, self.rank, self.world))

    def _stage(self, idx):
        """decode on host threads, copy to the device and launch the two augment kernels on the side stream"""
        items = [self.pairs.items[i] for i in idx]
        arrs = list(self.pool.map(_decode, [p for it in items for p in it[:2]]))
        n = len(items)
        span = self.load - self.fine
        with torch.cuda.stream(self.side):
            dev = [torch.from_numpy(a).pin_memory().to(self.device, non_blocking=True) for a in arrs]
            outs = []
            for k in range(2):     # A images, then B images: independent crop / flip draws per image, as the reference's transform
                crops = [(self.rng.randint(0, span), self.rng.randint(0, span)) for _ in range(n)]
                flips = [(not self.no_flip) and self.rng.random() < 0.5 for _ in range(n)]
                outs.append(augment(dev[k::2], crops, flips, self.load, self.fine, self.device, stream=self.side))
            ev = torch.cuda.Event()
            ev.record(self.side)
        return {"A": outs[0], "B": outs[1], "label": torch.tensor([it[2] for it in items], dtype=torch.int64),
                "A_paths": [it[0] for it in items], "B_paths": [it[1] for it in items]}, ev

    def __iter__(self):
        if not self.serial:
            self.pairs.shuffle(self.shared_rng)
        batches = rank_indices(len(self.pairs), self.B, self.rank, self.world)
        nxt = self._stage(batches[0]) if batches else None
        for i in range(len(batches)):
            cur, ev = nxt
            nxt = self._stage(batches[i + 1]) if i + 1 < len(batches) else None      # one batch ahead of the step
            torch.cuda.current_stream(self.device).wait_event(ev)
            yield cur


class DevicePrefetcher:
    """Batches that are already tensors on the host (a torch DataLoader with pin_memory=True, as data/__init__.py:64-70 of
    the reference builds): the host -> device copy of batch i + 1 is issued on a side stream while step i runs, so the copy
    is off the critical path of the step that consumes it.

        for batch in DevicePrefetcher(loader, device):        # dict of device tensors (non-tensor values pass through)
            model.set_input(batch); model.optimize_parameters()

    The device tensors live in three rotating sets of persistent buffers (no allocation per step): a yielded batch stays
    valid until two more batches have been requested.  Pinned source tensors copy asynchronously; pageable ones still
    work, the copy is then synchronous (as in torch)."""

    SLOTS = 3

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self._bufs = [dict() for _ in range(self.SLOTS)]
        self._ready = [None] * self.SLOTS
        self._used = [None] * self.SLOTS          # event: the consumer's stream is done with the slot
        self._i = 0
        self._next = None
        self._preload()

    def _preload(self):
        try:
            b = next(self.it)
        except StopIteration:
            self._next = None
            return
        if not self.cuda:
            self._next = (b, -1)
            return
        slot = self._i % self.SLOTS
        self._i += 1
        with torch.cuda.stream(self.stream):
            if self._used[slot] is not None:
                self.stream.wait_event(self._used[slot])
            out = {}
            for k, v in b.items():
                if torch.is_tensor(v):
                    buf = self._bufs[slot].get(k)
                    if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                        buf = self._bufs[slot][k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    buf.copy_(v, non_blocking=True)
                    out[k] = buf
                else:
                    out[k] = v
            if self._ready[slot] is None:
                self._ready[slot] = torch.cuda.Event()
            self._ready[slot].record(self.stream)
        self._next = (out, slot)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        batch, slot = self._next
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._ready[slot])
            # the slot that was handed out two batches ago is refilled next: the consumer's work queued so far covers it
            nxt = self._i % self.SLOTS
            if self._used[nxt] is None:
                self._used[nxt] = torch.cuda.Event()
            self._used[nxt].record(cur)
        self._preload()
        return batch
