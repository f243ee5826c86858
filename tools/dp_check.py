"""Multi-GPU check of the data-parallel step (run under torchrun, one rank per GPU): replicas built from different
seeds are broadcast from rank 0, trained for a few steps on DIFFERENT per-rank batches (eager, then the captured graph
with the bucketed NCCL all-reduces inside), and must hold bit-identical weights afterwards (a parameter missed by a
bucket, or averaged before its last gradient arrived, makes the ranks drift apart)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pcgan_b200.wsgan_emb_model import WSGANEmbModel, default_options

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
B, S = int(os.environ.get("B", "8")), int(os.environ.get("S", "64"))
torch.manual_seed(100 + rank)          # different initial weights per rank: setup() must make them identical
opt = default_options(batchSize=B, gpu_ids=[local], fineSize=S, loadSize=S, fineSize_E=64, cuda_graph=True, cuda_graph_warmup=2,
                      which_model_netG=os.environ.get("G", "resnet_9blocks"))
m = WSGANEmbModel()
m.initialize(opt)
m.setup(opt)
assert dist.is_initialized() and dist.get_world_size() == world
print("rank %d: G buckets %s D buckets %s" % (rank, m.sync_G.buckets, m.sync_D.buckets), flush=True)


def checksum():
    tot = torch.zeros(3, dtype=torch.float64, device="cuda")
    for i, net in enumerate((m.netG, m.netD, m.netE)):
        for t in net.parameters():          # running statistics (buffers) are per rank, as under nn.DataParallel
            tot[i] += t.detach().double().abs().sum()
    out = [torch.zeros_like(tot) for _ in range(world)]
    dist.all_gather(out, tot)
    return out


c0 = checksum()
assert all(torch.equal(c, c0[0]) for c in c0), "replicas differ after setup(): %s" % c0
g = torch.Generator().manual_seed(7 + rank)        # different data per rank
losses = []
for it in range(6):                                # 2 eager steps, capture at step 2, replays after
    m.set_input({"A": torch.rand(B, 3, S, S, generator=g) * 2 - 1, "B": torch.rand(B, 3, S, S, generator=g) * 2 - 1,
                 "label": torch.randint(0, 3, (B,), generator=g)})
    m.optimize_parameters()
    losses.append(m.get_current_losses())
torch.cuda.synchronize()
c1 = checksum()
for i, name in enumerate(("G", "D")):
    vals = [float(c[i]) for c in c1]
    assert all(v == vals[0] for v in vals), "net%s weights differ across ranks after 6 steps: %s" % (name, vals)
    assert vals[0] != float(c0[0][i]), "net%s did not train" % name
# BatchNorm running statistics of D / E stay per rank (nn.DataParallel semantics) and the batches differ: E's buffers differ
assert len(m._graphs) == 1 and not isinstance(list(m._graphs.values())[0], list), "the step (with its all-reduces) is ONE captured graph"
ls = torch.tensor([losses[-1]["G_GAN"], losses[-1]["D_fake"]], device="cuda", dtype=torch.float64)
out = [torch.zeros_like(ls) for _ in range(world)]
dist.all_gather(out, ls)
if rank == 0:
    print("OK: %d ranks, identical G / D weights after 6 steps (4 replayed); last losses per rank: %s" % (world, [[round(float(x), 4) for x in o] for o in out]))
m._graphs.clear()          # captured NCCL work must be gone before the communicator is torn down
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
