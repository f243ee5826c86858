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

# Workloads = BASELINE.json configs; GFLOP per image (pair) per training step from SURVEY.md §8(d) / BASELINE.md §4:
# algorithmic conv FLOPs (2 x true MACs, no padding).  The default (c128) is the configuration the metric is quoted on.
WORKLOADS = {
    "c128": dict(metric="wsgan_emb_train_images_per_sec_128", batch=64, size=128, gflop=180.02, flags={},
                 text="wsgan_emb optimize_parameters, 128x128, ResNet-9 G + 3-layer PatchGAN D + ResNet-18 Elo E@224, lambda_IP 0 "
                      "(BASELINE configs[2])"),
    "c256": dict(metric="wsgan_emb_train_images_per_sec_256", batch=32, size=256, gflop=679.20, flags={},
                 text="wsgan_emb optimize_parameters, 256x256, batch 32 per GPU (BASELINE configs[4])"),
    "bayesian": dict(metric="wsgan_emb_bayesian_train_images_per_sec_128", batch=64, size=128, gflop=275.13,
                     flags=dict(bayesian=True, noisy=True, noisy_var_type="ae", bnn_dropout=0.2, bnn_T=10),
                     text="wsgan_emb optimize_parameters, --bayesian true --noisy true --noisy_var_type ae --bnn_dropout 0.2 (T = 10 "
                          "Monte-Carlo encoder passes per image), 128x128 (BASELINE configs[3])"),
    "siamese": dict(metric="elo_siamese_train_pairs_per_sec_128", batch=64, size=128, gflop=6.98, flags={},
                    text="siamese.py Elo rating trainer step, ResNet-18 + cnn head, 64 pairs at 128x128, Adam (BASELINE configs[1])"),
}
GFLOP_PER_IMAGE = WORKLOADS["c128"]["gflop"]
METRIC = WORKLOADS["c128"]["metric"]
WORKLOAD = WORKLOADS["c128"]["text"]


def config_dict(B, S, world, launch, wl="c128"):
    """`config` of the JSON line: the same for both arms (the reference arm times a bounded sample of this workload)."""
    w = WORKLOADS[wl]
    return {"workload": w["text"], "workload_id": wl, "batch_per_gpu": B, "global_batch": world * B, "size": S, "parallelism": "dp%d" % world,
            "l2": "4 distinct input batches; GBs of activations per step >> 126 MB L2, no flush needed",
            "launch": launch, "flops_per_image": w["gflop"] * 1e9}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240, help="timed steps (default: about 5 s of timed region at 21 ms per step)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c128", choices=sorted(WORKLOADS), help="BASELINE.json configuration (default: the one the metric is quoted on)")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the cuDNN-eager leg (the oracle port of the reference step on this GPU)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="pairs per GPU (default: the workload's, 64 for BASELINE config 3)")
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--ref-batch", type=int, default=2, help="pairs per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of the step from Python instead of replaying a CUDA graph")
    ap.add_argument("--dump-igemm", default="", help="write the per-plan igemm timing table of the roofline pass to this file")
    args = ap.parse_args()
    args.batch = args.batch or WORKLOADS[args.workload]["batch"]
    args.size = args.size or WORKLOADS[args.workload]["size"]
    return args


def peaks():
    """(sustained bf16 TFLOP/s, burst bf16 TFLOP/s, HBM GB/s, source).  A kernel timed inside a long step is judged
    against the sustained figure, one timed alone against the burst figure; both are reported."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return (d.get("bf16_tflops_sustained", 1393.9), d.get("bf16_tflops", 1689.5), d.get("hbm_gbs", 6467.4),
                "measured (MEASURED_PEAKS.json)")
    return 1400.0, 1700.0, 6650.0, "fallback (B200_PROFILING.md)"


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
        pw = sorted(float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit())
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "power_w": pw[len(pw) // 2] if pw else None}


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


class StepRunner:
    """One workload behind a uniform interface: step(batch) enqueues one training step (batch tensors on the host or on
    the device), losses() reads the step's scalars back to the host."""

    def __init__(self, args, dev, local):
        import torch
        from pcgan_b200.wsgan_emb_model import WSGANEmbModel, default_options
        self.kind, self.dev = args.workload, dev
        B, S = args.batch, args.size
        if self.kind == "siamese":
            from pcgan_b200 import siamese as SI
            self.net = SI.get_model(gpu_ids=[local])
            self.trainer = SI.EloTrainer(self.net, lr=2e-4, cuda_graph=not args.no_graph)
            self.model, self.loss = None, None
            self.n_losses = 1
            return
        opt = default_options(batchSize=B, gpu_ids=[local], fineSize=S, loadSize=S, cuda_graph=not args.no_graph,
                              group_passes=os.environ.get("PCGAN_GROUP", "1") != "0", **WORKLOADS[self.kind]["flags"])
        self.model = WSGANEmbModel()
        with contextlib.redirect_stdout(sys.stderr):   # the factories print like the reference's do; stdout carries only the JSON line
            self.model.initialize(opt)
            self.model.setup(opt)                      # broadcasts rank 0's random init to the other ranks
        self.n_losses = 9

    @property
    def use_graph(self):
        return bool(self.model.use_graph if self.model is not None else self.trainer.use_graph)

    def set_graph(self, on):
        if self.model is not None:
            self.model.use_graph = on
        else:
            self.trainer.use_graph = on

    def step(self, batch):
        if self.model is not None:
            self.model.set_input(batch)
            self.model.optimize_parameters()
        else:
            a, b, l = (batch[k].to(self.dev, non_blocking=True) for k in ("A", "B", "label"))
            self.loss, _ = self.trainer.train_step(a, b, l)

    def losses(self):
        if self.model is not None:
            return self.model.get_current_losses()      # float() of the nine losses: device -> host
        return {"elo_nll": float(self.loss)}


def gpu_reference_baseline(B, S, dev, seconds=12.0):
    """The reference's own GPU path on this B200, as a stated baseline ("the real bar to beat"): the pinned
    oracle port of optimize_parameters run by PyTorch eager + cuDNN on cuda, at the same batch, in the arithmetic modes
    the reference can run in (strict fp32 = what base_model.py:26-27 sets up; TF32 allowed; bf16 autocast)."""
    import torch
    from oracle import pcgan_oracle as O
    out = {}
    modes = (("fp32", False, None), ("tf32", True, None), ("bf16_autocast", True, torch.bfloat16))
    for name, tf32, amp in modes:
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True          # base_model.py:26-27
        m = O.WSGANEmbOracle(O.make_state_dict(O.generator_keys(), 31, device=dev, requires_grad=True),
                             O.make_state_dict(O.discriminator_keys(), 32, device=dev, requires_grad=True),
                             O.make_state_dict(O.encoder_keys(), 33, device=dev))
        a, b, label = O.synthetic_batch(B, S, 1234, device=dev)
        ctx = (lambda: torch.autocast("cuda", dtype=amp)) if amp is not None else contextlib.nullcontext
        try:
            for _ in range(3):
                with ctx():
                    m.optimize_parameters(a, b, label)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n, t0 = 0, time.perf_counter()
            e0.record()
            while n < 5 or (time.perf_counter() - t0 < seconds / len(modes) and n < 40):
                with ctx():
                    m.optimize_parameters(a, b, label)
                n += 1
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            out[name] = {"value": B / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "steps": n}
        except Exception as ex:     # a mode the oracle port cannot run in (e.g. autocast through a custom function) is reported, not fatal
            out[name] = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:120])}
        del m
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return {"what": "oracle port of the reference step, PyTorch eager + cuDNN on this GPU, batch %d, device-resident inputs" % B, "modes": out}


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

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    if args.workload == "siamese" and world > 1:
        raise SystemExit("the siamese workload of this bench is single-GPU")
    B, S, K, W = args.batch, args.size, args.steps, max(args.warmup, 3)

    torch.manual_seed(1234 + rank)
    run = StepRunner(args, dev, local)
    W_eff = W + ((4 if run.model is not None else 3) if run.use_graph else 0)   # graph mode: 3 eager steps + the capture step come before the W replayed warm-ups

    # synthetic UTKFace-shaped pool in pinned host memory; distinct batches so nothing is cached
    pool = 4
    host = []
    g = torch.Generator().manual_seed(1234 + rank)
    for _ in range(pool):
        host.append({"A": (torch.rand(B, 3, S, S, generator=g) * 2 - 1).pin_memory(), "B": (torch.rand(B, 3, S, S, generator=g) * 2 - 1).pin_memory(),
                     "label": torch.randint(0, 3, (B,), generator=g).pin_memory()})
    resident = [{"A": h["A"].to(dev), "B": h["B"].to(dev), "label": h["label"] if run.model is not None else h["label"].to(dev)} for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step = run.step
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
    sampler2 = ClockSampler(local)     # the second region runs on an already warm, power-capped GPU: its clocks are reported beside it
    if rank == 0:
        sampler2.start()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = None
    for i in range(K):
        step(host[i % pool])
        last = run.losses()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clocks_e2e = sampler2.stop() if rank == 0 else None

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline pass: every launch of rank 0 bracketed by CUDA events on its stream.  All ranks run the steps
    # (they contain the gradient all-reduces); only rank 0 records.
    roof = roof_hbm = None
    nprof = 2
    if rank == 0:
        ops.Stats.igemm_events = []
        ops.Stats.op_events = []
    was_graph = run.use_graph
    run.set_graph(False)     # per-launch events need the launches to come from Python
    for i in range(nprof):
        step(resident[i % pool])
    barrier()
    run.set_graph(was_graph)
    if rank == 0:
        ev = ops.Stats.igemm_events
        ops.Stats.igemm_events = None
        op_ev, ops.Stats.op_events = ops.Stats.op_events, None
        oagg = {}
        for name, a, b, nbytes in op_ev:
            d = oagg.setdefault(name, [0, 0.0, 0])
            d[0] += 1; d[1] += a.elapsed_time(b); d[2] += nbytes
        if args.dump_igemm:
            with open(args.dump_igemm + ".ops", "w") as fh:
                tot = sum(v[1] for v in oagg.values()) / nprof
                fh.write("# non-igemm kernels of one step, CUDA events around each launch (includes ~2 us of launch latency each): %.3f ms\n" % tot)
                for name, (n, t_, by) in sorted(oagg.items(), key=lambda kv: -kv[1][1]):
                    fh.write("%-24s n=%4d  %8.3f ms/step  %8.1f MB/step  %7.1f GB/s\n" % (name, n / nprof, t_ / nprof, by / nprof / 1e6, by / max(t_, 1e-9) / 1e6))
            agg = {}
            for note, f, a, b in ev:
                d = agg.setdefault(note, [0, 0.0, 0])
                d[0] += 1; d[1] += a.elapsed_time(b); d[2] += f
            with open(args.dump_igemm, "w") as fh:
                for note, (n, t_, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    fh.write("%-40s n=%3d  %8.3f ms/step  issued %7.1f TFLOP/s\n" % (note, n / nprof, t_ / nprof, f / max(t_, 1e-9) / 1e9))
        tot_ms = sum(a.elapsed_time(b) for _, _, a, b in ev) / nprof
        n_ig = len(ev) / nprof
        issued = sum(f for _, f, _, _ in ev) / nprof
        peak_tf, peak_burst, peak_hbm, how = peaks()
        alg = wl["gflop"] * 1e9 * B
        ach = alg / (tot_ms * 1e-3) / 1e12
        step_tf = alg / (ms / K * 1e-3) / 1e12       # the whole step (every kernel, launch gaps included) against the tensor peak
        roof = {"bound": "tensor", "kernel": "pcgan::igemm_kernel (all %d conv launches of a step)" % round(n_ig), "achieved": ach, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None, "peak_source": how + ": sustained (kernels timed inside a long step)",
                "peak_burst": peak_burst, "frac_burst": ach / peak_burst,
                "step_achieved": step_tf, "step_frac": step_tf / peak_tf, "step_frac_burst": step_tf / peak_burst,
                "timing": "CUDA events around every igemm launch of %d eager steps after the timed region" % nprof,
                "kernel_ms_per_step": tot_ms, "kernel_share_of_step": tot_ms / (ms / K), "launches_per_step": n_ig,
                "issued_tflops": issued / (tot_ms * 1e-3) / 1e12,
                "algorithmic_flops_per_launch": alg / n_ig, "avg_launch_us": 1e3 * tot_ms / n_ig}
        # HBM roofline of the bandwidth-bound family (normalisation / activation / residual / fold kernels)
        fam = ("norm_apply", "norm_bwd_reduce", "norm_bwd_apply", "halo_fold")
        f_ms = sum(oagg[k][1] for k in fam if k in oagg) / nprof
        f_by = sum(oagg[k][2] for k in fam if k in oagg) / nprof
        f_n = sum(oagg[k][0] for k in fam if k in oagg) / nprof
        all_ms = sum(v[1] for v in oagg.values()) / nprof
        if f_ms > 0:
            gbs = f_by / (f_ms * 1e-3) / 1e9
            roof_hbm = {"bound": "hbm", "kernel": "pcgan::norm_apply / norm_bwd_reduce / norm_bwd_apply / halo_fold (%d launches of a step)" % round(f_n),
                        "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                        "traffic": None, "traffic_note": "ncu dram__bytes of these kernels = algorithmic bytes (profiles/r1_norm_ncu_full_metrics.csv: 67.4 MB read for a 2 x 33.6 MB launch)",
                        "algorithmic_bytes_per_step": f_by, "algorithmic_bytes_per_launch": f_by / f_n, "avg_launch_us": 1e3 * f_ms / f_n,
                        "kernel_ms_per_step": f_ms, "all_non_conv_kernels_ms_per_step": all_ms,
                        "timing": "CUDA events around every launch of %d eager steps (each includes ~2 us of launch latency)" % nprof,
                        "per_kernel": {k: {"launches": oagg[k][0] / nprof, "ms": oagg[k][1] / nprof, "GBps": oagg[k][2] / max(oagg[k][1], 1e-9) / 1e6}
                                       for k in fam if k in oagg}}

    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_gpu_reference and run.model is not None and args.workload in ("c128", "c256"):
        gpu_ref = gpu_reference_baseline(B, S, dev)

    if rank == 0:
        img = world * B * K
        unit = "pairs/s" if args.workload == "siamese" else "images/s"
        line = {"metric": wl["metric"], "value": img / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config_dict(B, S, world, "CUDA graph replay of the captured step" if run.use_graph else "per-kernel launches from Python", args.workload),
                "clocks": clocks, "timed_region_s": ms * 1e-3,
                "e2e": {"value": img / (ms_e2e * 1e-3), "unit": unit, "h2d_bytes_per_step": 2 * B * 3 * S * S * 4 + B * 8, "d2h_bytes_per_step": run.n_losses * 4,
                        "ms_per_step": ms_e2e / K, "clocks": clocks_e2e},
                "gpu_launches": int(round(launches * K)), "gpu_launches_per_step": launches, "host_enqueue_ms_per_step": host_ms,
                "cuda_graph": run.use_graph,
                "roofline": roof, "roofline_hbm": roof_hbm, "gpu_reference_baseline": gpu_ref, "last_losses": last}
        if not args.no_cpu_baseline and world == 1 and args.workload == "c128":
            line["cpu_baseline"] = cpu_baseline(S)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # the captured step holds NCCL work: drop the graphs and leave without tearing the communicator down (a
        # destroy_process_group with captured collectives alive can block forever)
        if run.model is not None:
            run.model._graphs.clear()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
