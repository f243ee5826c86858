"""World-size-2 gloo test (CPU) of the data-parallel host logic: flat gradient buffers, averaging all-reduce,
batch sharding.  The kernels are not involved; NCCL runs the same code path on the box."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pcgan_b200.dist import GradSync, shard_batch
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    sync = GradSync(params)
    sync.zero()
    assert params[0].grad is sync.views[0] and float(sync.flat.abs().sum()) == 0
    data = torch.arange(8, dtype=torch.float32).view(8, 1)
    mine = data[shard_batch(8, rank, world)]
    # a "kernel" accumulating straight into .grad (as ConvRT.backward_weight does)
    params[0].grad.add_(mine.sum())
    params[1].grad.add_(float(rank + 1))
    sync.all_reduce()
    out[rank] = (float(params[0].grad[0, 0]), float(params[1].grad[0]), params[0].grad.data_ptr() == sync.flat.data_ptr())
    sync.zero()
    assert float(params[1].grad.abs().sum()) == 0
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        res = dict(out)
    # rank 0 holds rows 0..3 (sum 6), rank 1 rows 4..7 (sum 22): mean 14; second param: mean(1, 2) = 1.5
    for r in (0, 1):
        assert res[r][0] == pytest.approx(14.0) and res[r][1] == pytest.approx(1.5) and res[r][2]


def _bucket_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pcgan_b200.dist import GradSync
    torch.manual_seed(rank)       # replicas that start from different seeds ...
    layers = [[torch.nn.Parameter(torch.randn(64, 8)), torch.nn.Parameter(torch.randn(64))],
              [torch.nn.Parameter(torch.randn(128, 16))], [torch.nn.Parameter(torch.randn(32, 4)), torch.nn.Parameter(torch.randn(32))]]
    params = [p for l in layers for p in l]
    for p in params:              # ... are made identical the way BaseModel.broadcast_replicas does
        dist.broadcast(p.data, 0)
    sync = GradSync(params, bucket_bytes=2048, layers=layers)
    # buckets are unions of whole layers, last layer first, contiguous and covering the flat buffer
    assert sync.buckets[0][1] == sync.flat.numel() and sync.buckets[-1][0] == 0
    assert all(sync.buckets[i][0] == sync.buckets[i + 1][1] for i in range(len(sync.buckets) - 1)) and len(sync.buckets) >= 2
    for l in layers:
        assert len({sync.bucket_of[id(p)] for p in l}) == 1
    sync.zero()
    for i, p in enumerate(params):
        p.grad.add_(float((rank + 1) * (i + 1)))
    # the sweep hands over buckets as their layers finish (last layers first); the rest goes at finish()
    sync.bucket_ready(0)
    sync.finish()
    out[rank] = ([float(p.grad.flatten()[0]) for p in params], float(params[0].detach().sum()))
    dist.destroy_process_group()


def test_bucketed_async_all_reduce_world2_gloo():
    port = 31500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_bucket_worker, args=(2, port, out), nprocs=2, join=True)
        res = dict(out)
    for r in (0, 1):
        assert res[r][0] == pytest.approx([1.5 * (i + 1) for i in range(5)])
    assert res[0][1] == pytest.approx(res[1][1])      # broadcast made the replicas identical


def test_shard_batch():
    from pcgan_b200.dist import shard_batch
    assert shard_batch(128, 1, 2) == slice(64, 128)
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)
