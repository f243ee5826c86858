"""Benchmark of the wsgan_emb training step (BASELINE.json metric: PC-GAN train images/sec at 128x128).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port) on host cores

One step = WSGANEmbModel.optimize_parameters() (models/wsgan_emb_model.py:478-484) on one synthetic batch of
64 pairs per GPU: ResNet-9 G forward x2 + backward x2, PatchGAN D forward x4 + backward x4, Elo encoder forward x3
+ data-gradient x1, the GAN / cycle / embedding-reconstruction losses and both Adam steps.  Prints ONE JSON line.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_IMAGE = 180.02   # algorithmic conv FLOPs of one step per image at 128^2, fineSize_E 224, lambda_IP 0
METRIC = "wsgan_emb_train_images_per_sec_128"


WORKLOAD = ("wsgan_emb optimize_parameters, 128x128, ResNet-9 G + 3-layer PatchGAN D + ResNet-18 Elo E@224, lambda_IP 0 "
            "(BASELINE configs[2])")


def config_dict(B, S, world, launch):
    """`config` of the JSON line: the same for both arms (the reference arm times a bounded sample of this workload)."""
    return {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B, "size": S, "parallelism": "dp%d" % world,
            "l2": "4 distinct input batches; ~5 GB of activations per step >> 126 MB L2, no flush needed",
            "launch": launch, "flops_per_image": GFLOP_PER_IMAGE * 1e9}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="pairs per GPU (BASELINE config 3: 64)")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--ref-batch", type=int, default=2, help="pairs per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of the step from Python instead of replaying a CUDA graph")
    ap.add_argument("--dump-igemm", default="", help="write the per-plan igemm timing table of the roofline pass to this file")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1393.9), d.get("hbm_gbs", 6467.4), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def run_reference(args):
    """The reference's own CPU implementation of the step (the pinned oracle port of phymhan/pc-gan's
    optimize_parameters; the Python reference itself cannot travel to the GPU box), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import pcgan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, S = args.ref_batch, args.size
    m = O.WSGANEmbOracle(O.make_state_dict(O.generator_keys(), 31, requires_grad=True),
                         O.make_state_dict(O.discriminator_keys(), 32, requires_grad=True), O.make_state_dict(O.encoder_keys(), 33))
    times = []
    for it in range(args.warmup + args.steps):
        a, b, label = O.synthetic_batch(B, S, 1234 + it)
        t0 = time.perf_counter()
        m.optimize_parameters(a, b, label)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    val = B * len(times) / total
    line = {"metric": METRIC, "value": val, "unit": "images/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.batch, S, args.gpus, "reference arm: torch CPU, %d pairs per step" % B),
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": "%d steps of batch %d (oracle port of the reference step, fp32, torch CPU)" % (len(times), B)},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline(size, seconds_budget=25.0):
    import torch
    from oracle import pcgan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = O.WSGANEmbOracle(O.make_state_dict(O.generator_keys(), 31, requires_grad=True),
                         O.make_state_dict(O.discriminator_keys(), 32, requires_grad=True), O.make_state_dict(O.encoder_keys(), 33))
    B, n, t_total = 2, 0, 0.0
    a, b, label = O.synthetic_batch(B, size, 1234)
    m.optimize_parameters(a, b, label)  # warm-up
    while n < 2 or (t_total < seconds_budget * 0.5 and n < 8):
        a, b, label = O.synthetic_batch(B, size, 1235 + n)
        t0 = time.perf_counter()
        m.optimize_parameters(a, b, label)
        t_total += time.perf_counter() - t0
        n += 1
    return {"value": B * n / t_total, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "%d steps of batch %d after 1 warm-up (oracle port of the reference step, fp32, torch CPU)" % (n, B)}


def main():
    args = parse()
    # a stuck collective or capture must not hang the caller: dump every thread's stack and exit
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_WATCHDOG_S", "900")), exit=True)
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from pcgan_b200 import ops
    from pcgan_b200.wsgan_emb_model import WSGANEmbModel, default_options

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S, K, W = args.batch, args.size, args.steps, max(args.warmup, 3)
    W_eff = W + (4 if not args.no_graph else 0)   # graph mode: 3 eager steps + the capture step come before the W replayed warm-ups

    torch.manual_seed(1234 + rank)
    opt = default_options(batchSize=B, gpu_ids=[local], fineSize=S, loadSize=S, cuda_graph=not args.no_graph)
    model = WSGANEmbModel()
    with contextlib.redirect_stdout(sys.stderr):   # the factories print like the reference's do; stdout carries only the JSON line
        model.initialize(opt)
        model.setup(opt)
    if world > 1:   # identical replicas: broadcast rank 0's random init
        for net in (model.netG, model.netD, model.netE):
            for t in list(net.parameters()) + list(net.buffers()):
                dist.broadcast(t.data, 0)

    # synthetic UTKFace-shaped pool in pinned host memory; distinct batches so nothing is cached
    pool = 4
    host = []
    g = torch.Generator().manual_seed(1234 + rank)
    for _ in range(pool):
        host.append({"A": (torch.rand(B, 3, S, S, generator=g) * 2 - 1).pin_memory(), "B": (torch.rand(B, 3, S, S, generator=g) * 2 - 1).pin_memory(),
                     "label": torch.randint(0, 3, (B,), generator=g)})
    resident = [{"A": h["A"].to(dev), "B": h["B"].to(dev), "label": h["label"]} for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(batch):
        model.set_input(batch)
        model.optimize_parameters()

    launches_per_step = None
    for i in range(W_eff):
        l_before = ops.Stats.launches
        step(resident[i % pool])
        if i == 1:   # an eager step (graph mode captures after 3 of them): the kernels one step launches
            launches_per_step = ops.Stats.launches - l_before
    barrier()

    # ---- timed region 1: device-resident inputs -> `value`
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host = time.perf_counter()
    for i in range(K):
        step(resident[i % pool])
    host_ms = 1e3 * (time.perf_counter() - t_host) / K     # CPU time to enqueue one step (no sync inside)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = float(launches_per_step)
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end through the public API with HOST buffers (H2D of the batch + D2H of the losses)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = None
    for i in range(K):
        step(host[i % pool])
        last = model.get_current_losses()      # float() of the nine losses: device -> host
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline pass: every igemm launch of rank 0 bracketed by CUDA events on its stream.  All ranks run the steps
    # (they contain the gradient all-reduces); only rank 0 records.
    roof = None
    nprof = 2
    if rank == 0:
        ops.Stats.igemm_events = []
        ops.Stats.op_events = [] if args.dump_igemm else None
    was_graph, model.use_graph = model.use_graph, False     # per-launch events need the launches to come from Python
    for i in range(nprof):
        step(resident[i % pool])
    barrier()
    model.use_graph = was_graph
    if rank == 0:
        ev = ops.Stats.igemm_events
        ops.Stats.igemm_events = None
        op_ev, ops.Stats.op_events = ops.Stats.op_events, None
        if args.dump_igemm:
            oagg = {}
            for name, a, b in op_ev:
                d = oagg.setdefault(name, [0, 0.0])
                d[0] += 1; d[1] += a.elapsed_time(b)
            with open(args.dump_igemm + ".ops", "w") as fh:
                tot = sum(v[1] for v in oagg.values()) / nprof
                fh.write("# non-igemm kernels of one step, CUDA events around each launch (includes ~2 us of launch latency each): %.3f ms\n" % tot)
                for name, (n, t) in sorted(oagg.items(), key=lambda kv: -kv[1][1]):
                    fh.write("%-24s n=%4d  %8.3f ms/step\n" % (name, n / nprof, t / nprof))
            agg = {}
            for note, f, a, b in ev:
                d = agg.setdefault(note, [0, 0.0, 0])
                d[0] += 1; d[1] += a.elapsed_time(b); d[2] += f
            with open(args.dump_igemm, "w") as fh:
                for note, (n, t, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    fh.write("%-40s n=%3d  %8.3f ms/step  issued %7.1f TFLOP/s\n" % (note, n / nprof, t / nprof, f / max(t, 1e-9) / 1e9))
        tot_ms = sum(a.elapsed_time(b) for _, _, a, b in ev) / nprof
        n_ig = len(ev) / nprof
        issued = sum(f for _, f, _, _ in ev) / nprof
        peak_tf, _, how = peaks()
        alg = GFLOP_PER_IMAGE * 1e9 * B
        ach = alg / (tot_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "pcgan::igemm_kernel (all %d conv launches of a step)" % round(n_ig), "achieved": ach, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None, "peak_source": how,
                "kernel_ms_per_step": tot_ms, "kernel_share_of_step": tot_ms / (ms / K), "launches_per_step": n_ig,
                "issued_tflops": issued / (tot_ms * 1e-3) / 1e12,
                "algorithmic_flops_per_launch": alg / n_ig, "avg_launch_us": 1e3 * tot_ms / n_ig}

    if rank == 0:
        img = world * B * K
        line = {"metric": METRIC, "value": img / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config_dict(B, S, world, "CUDA graph replay of the captured step" if model.use_graph else "per-kernel launches from Python"),
                "clocks": clocks,
                "e2e": {"value": img / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": 2 * B * 3 * S * S * 4, "d2h_bytes_per_step": 9 * 4,
                        "ms_per_step": ms_e2e / K},
                "gpu_launches": int(round(launches * K)), "gpu_launches_per_step": launches, "host_enqueue_ms_per_step": host_ms,
                "cuda_graph": bool(model.use_graph),
                "roofline": roof, "last_losses": last}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(S)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
