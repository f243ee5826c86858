"""Host side of the input pipeline (pcgan_b200/data.py): pair-list parsing and per-rank sharding of the global batches
(data/wsgan_emb_dataset.py:14-34, data/__init__.py:55-74: nn.DataParallel scatters each batch along dim 0)."""
import random

from pcgan_b200.data import PairList, rank_indices


def test_pair_list_and_shuffle(tmp_path):
    src = tmp_path / "pairs.txt"
    src.write_text("".join("a%d.jpg b%d.jpg %d\n" % (i, i, i % 3) for i in range(10)))
    pl = PairList(str(src), "/data", max_dataset_size=8)
    assert len(pl) == 8 and pl.items[3] == ("/data/a3.jpg", "/data/b3.jpg", 0)
    before = list(pl.items)
    pl.shuffle(random.Random(0))
    assert sorted(pl.items) == sorted(before) and pl.items != before


def test_rank_shards_partition_every_global_batch():
    n, B, W = 103, 8, 4
    shards = [rank_indices(n, B, r, W) for r in range(W)]
    seen = sorted(i for s in shards for b in s for i in b)
    assert seen == list(range(n))                                   # every pair exactly once per epoch
    for g in range(len(shards[0])):
        rows = [s[g] for s in shards if g < len(s)]
        flat = [i for b in rows for i in b]
        assert flat == list(range(flat[0], flat[0] + len(flat)))    # contiguous runs, rank order = scatter order
    assert all(len(b) == B for s in shards for b in s[:-1])
    assert rank_indices(n, B, 0, W, drop_last=True)[-1][-1] < (n // (B * W)) * B * W
    assert rank_indices(16, 8, 1, 2) == [[8, 9, 10, 11, 12, 13, 14, 15]]


def test_device_prefetcher_passes_batches_through_in_order():
    """DevicePrefetcher on the CPU device degenerates to the plain iterator (same objects, same order, StopIteration)."""
    import torch
    from pcgan_b200.data import DevicePrefetcher
    batches = [{"A": torch.full((2, 3), float(i)), "label": torch.tensor([i]), "paths": ["p%d" % i]} for i in range(4)]
    got = list(DevicePrefetcher(batches, "cpu"))
    assert len(got) == 4
    for i, b in enumerate(got):
        assert float(b["A"][0, 0]) == i and int(b["label"]) == i and b["paths"] == ["p%d" % i]
    assert list(DevicePrefetcher([], "cpu")) == []
