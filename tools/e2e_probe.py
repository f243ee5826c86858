"""Which input path gives the best end-to-end step: batches copied on the step's stream, or DevicePrefetcher variants."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pcgan_b200.data import DevicePrefetcher

args = argparse.Namespace(workload="c128", batch=64, size=128, no_graph=False)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
run = bench.StepRunner(args, dev, 0)
g = torch.Generator().manual_seed(1)
host = [{"A": (torch.rand(64, 3, 128, 128, generator=g) * 2 - 1).pin_memory(), "B": (torch.rand(64, 3, 128, 128, generator=g) * 2 - 1).pin_memory(),
         "label": torch.randint(0, 3, (64,), generator=g).pin_memory()} for _ in range(4)]
for i in range(12):
    run.step(host[i % 4])
torch.cuda.synchronize()
K = 100


def timed(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    print("%-40s %.3f ms/step" % (name, 1e3 * (time.perf_counter() - t0) / K), flush=True)


def direct():
    for i in range(K):
        run.step(host[i % 4])
        run.losses()


resident = [{"A": h["A"].to(dev), "B": h["B"].to(dev), "label": h["label"]} for h in host]


def resident_sync():
    for i in range(K):
        run.step(resident[i % 4])
        run.losses()


def direct_nosync():
    for i in range(K):
        run.step(host[i % 4])


def pref():
    for b in DevicePrefetcher((host[i % 4] for i in range(K)), dev):
        run.step(b)
        run.losses()


def pref_images_only():
    class P(DevicePrefetcher):
        pass
    it = ({"A": host[i % 4]["A"], "B": host[i % 4]["B"]} for i in range(K))
    for i, b in enumerate(DevicePrefetcher(it, dev)):
        b["label"] = host[i % 4]["label"]
        run.step(b)
        run.losses()


def manual_after():
    st = torch.cuda.Stream()
    bufs = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items() if k != "label"} for _ in range(2)]
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    def issue(i):
        with torch.cuda.stream(st):
            for k in ("A", "B"):
                bufs[i % 2][k].copy_(host[i % 4][k], non_blocking=True)
            evs[i % 2].record(st)
    issue(0)
    for i in range(K):
        torch.cuda.current_stream().wait_event(evs[i % 2])
        b = dict(bufs[i % 2]); b["label"] = host[i % 4]["label"]
        run.step(b)
        if i + 1 < K:
            issue(i + 1)
        run.losses()


K = int(os.environ.get("PROBE_K", 200))
for rep in range(int(os.environ.get("PROBE_REPS", 5))):
    timed("device-resident inputs", resident_sync)
    timed("direct (H2D on the step's stream)", direct)
    timed("DevicePrefetcher", pref)
