#!/usr/bin/env python
"""bench.py — the Plastic U-Net hot path on B200: train images/sec (fwd + bwd + plastic update + Adam).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels, one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port) on host cores

Workload at N=1 (BASELINE.json configs[1]): Plastic U-Net (UNetp), Oja rule, 128x128, batch 64 per GPU,
synthetic 1-channel images (uniform [0,1) 101x101 zero-padded to 128x128) and Bernoulli masks, random-init
weights (torch.manual_seed(0)).  N>1: same per-GPU batch (weak scaling), gradient + trace-delta all-reduce.

One "step" = one optimisation step over one batch.  `value` is whole-job images/s with the batch already
resident in HBM (a device pool larger than L2 is rotated through); `e2e` is the same step through the
public API with pinned HOST batches (H2D copy of images+masks and D2H read of the loss inside the timed
region, every step).  Timing: CUDA events on the launching stream, barrier + synchronize on both sides,
max over ranks.  Prints ONE JSON line on rank 0.
"""
import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "plastic-unet_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "train images/sec (fwd+bwd+plastic update) @128x128"
UNIT = "images/s"
L2_BYTES = 126 * 1024 * 1024


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def synth_batch(n, size, gen, pad_from=101):
    """Synthetic seismic-shaped data: uniform images pad_from^2 zero-padded to size^2, Bernoulli(0.25) masks."""
    if size == 128 and pad_from == 101:
        img = torch.zeros(n, 1, 128, 128)
        img[:, :, 13:114, 13:114] = torch.rand(n, 1, 101, 101, generator=gen)
        msk = torch.zeros(n, 128, 128)
        msk[:, 13:114, 13:114] = (torch.rand(n, 101, 101, generator=gen) < 0.25).float()
    else:
        img = torch.rand(n, 1, size, size, generator=gen)
        msk = (torch.rand(n, size, size, generator=gen) < 0.25).float()
    return img, msk


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every few ms DURING the timed region (a thread; the
    timed region is tens of ms, too short for an `nvidia-smi -lms` child to report)."""

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons, self.maxclk = [], set(), None
        self._stop = False
        self._thr = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.maxclk = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.0005)

    def start(self):
        if self.nv is None:
            return
        import threading
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def stop(self):
        if self.nv is None or self._thr is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop = True
        self._thr.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.maxclk,
                "samples": len(self.samples), "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's own algorithm (oracle port of train.py:91-112) on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_train_images_per_s(n_images, size, rule, warm=3, seed=0):
    """B=1 sequential steps exactly like train.py:91-112 on the oracle; -> (images/s, n_images, threads)."""
    import plastic_unet_oracle as orc
    from pu_b200 import UNetp
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    net = quiet(UNetp, 1, 1, torch.device("cpu"), rule=rule, nbf=size)  # parameter container only (never run on CPU)
    sd = orc.leaf_state(net.state_dict())
    gen = torch.Generator().manual_seed(1234)
    imgs, msks = synth_batch(n_images + warm, size, gen)
    params = [v for v in sd.values() if v.is_floating_point() and v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)
    crit = torch.nn.BCELoss()
    hebb = torch.zeros(size, size)
    t0 = None
    for i in range(n_images + warm):
        if i == warm:
            t0 = time.perf_counter()
        opt.zero_grad()
        _, y, hebb = orc.forward("unetp", sd, imgs[i:i + 1], hebb.detach(), rule=rule)
        loss = crit(y.view(-1), msks[i].view(-1))
        loss.item()
        loss.backward()
        opt.step()
    dt = time.perf_counter() - t0
    return n_images / dt, n_images, threads


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is pure Python
    + PyTorch CPU ops, there is nothing to compile), all host threads, bounded sample per step."""
    if rank != 0:
        return
    per_step = args.ref_images_per_step
    threads = os.cpu_count() or 1
    # warm-up steps + timed steps, each step = per_step sequential B=1 train iterations
    ips_w, _, _ = cpu_train_images_per_s(max(1, args.warmup) * per_step, args.size, args.rule)
    t0 = time.perf_counter()
    ips, n, threads = cpu_train_images_per_s(args.steps * per_step, args.size, args.rule, warm=0)
    dt = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "UNetp Oja 128x128 (101x101 zero-padded), B=1 sequential steps as train.py:91-112, "
                               "%d images per step" % per_step},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d sequential single-image train steps (fwd+BCE+bwd+Adam+trace) of the oracle port" % n},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# dominant-kernel roofline (measured live with CUDA events on the launching stream)
# --------------------------------------------------------------------------------------------------
def dominant_kernel_roofline(batch, size, math, dev, hbm_peak, peak_src):
    """Times the full-resolution 16->8 conv3x3 (+bias+ReLU, two-source concat: up4.0, the largest single layer of
    UNetp: SURVEY.md §8d) in isolation, rotating over buffers larger than L2.  Algorithmic bytes per launch =
    pixels * (C_in + C_out) * 4 (input read once + output written once)."""
    from pu_b200 import ops
    C0, C1, Cout = 8, 8, 8
    nbuf = max(2, int(2 * L2_BYTES // (batch * size * size * (C0 + C1 + Cout) * 4)) + 1)
    xs0 = [torch.rand(batch, size, size, C0, device=dev) for _ in range(nbuf)]
    xs1 = [torch.rand(batch, size, size, C1, device=dev) for _ in range(nbuf)]
    w = torch.randn(Cout, C0 + C1, 3, 3, device=dev) * 0.1
    b = torch.zeros(Cout, device=dev)
    iters = 20
    with torch.no_grad():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                ops.conv3x3(xs0[i % nbuf], xs1[i % nbuf], w, b, None, True, size, size, 0, 0, 0, 0, math)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # a Python op call costs ~60 us of host time: capture the launches so that the events time the GPU, not the host
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(iters):
                ops.conv3x3(xs0[i % nbuf], xs1[i % nbuf], w, b, None, True, size, size, 0, 0, 0, 0, math)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (iters * reps)
    # one kernel per op call (the weight tiles are built inside the conv kernel)
    alg_bytes = batch * size * size * (C0 + C1 + Cout) * 4
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full` capture
    # (profiles/r1_tc_conv_ncu_summary.md, B = 64 @128x128, TF32 kernel): 67.16 MB read (= the algorithmic input, read
    # once) + 6.59 MB written inside the kernel window; the rest of the 33.55 MB output leaves L2 after the kernel ends.
    traffic, traffic_note = None, None
    if math and batch == 64 and size == 128:
        traffic = 67162368 + 6587904
        traffic_note = "ncu capture of round 1 (profiles/r1_tc_conv_ncu_summary.md); output write-back continues after the kernel window"
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "conv3x3 fwd 16->8 @%dx%d (up4.0, %s)" % (size, size, "tf32 tcgen05" if math else "fp32 ffma"),
            "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
            "traffic_note": traffic_note,
            "peak_source": peak_src, "us_per_launch": ms * 1e3, "algorithmic_bytes_per_launch": alg_bytes}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--rule", default="oja")
    ap.add_argument("--math", default=os.environ.get("PU_CONV_MATH", "tf32"), choices=["fp32", "tf32"])
    ap.add_argument("--model", default="unetp", choices=["unetp", "res", "coord"], help="configs[1] is unetp; the others are the variants of configs[2..4]")
    ap.add_argument("--neurons", type=int, default=16)
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--base", type=int, default=8)
    ap.add_argument("--dropout", type=float, default=0.5)
    ap.add_argument("--infer", action="store_true", help="time the batched forward-only step instead of the train step")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-images-per-step", type=int, default=16)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    from pu_b200 import dp
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return

    import torch.distributed as dist
    from pu_b200 import UNetp, UNetpCoord, UNetpRes, _lib
    from pu_b200.trainer import InferStep, TrainStep

    rank, world, local_rank = dp.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    hbm_peak, _, peak_src = measured_peaks()

    torch.manual_seed(0)
    if args.model == "unetp":
        net = quiet(UNetp, 1, 1, dev, rule=args.rule, nbf=args.size, batched=True, depth=args.depth, base=args.base)
        model_name = "UNetp (Plastic U-Net)" + ("" if (args.depth, args.base) == (4, 8) else " depth %d base %d" % (args.depth, args.base))
    elif args.model == "res":
        net = quiet(UNetpRes, 1, 1, dev, neurons=args.neurons, dropout_ratio=args.dropout, rule=args.rule, nbf=args.size, batched=True,
                    depth=args.depth)
        model_name = "UNetpRes (residual Plastic U-Net) neurons %d dropout %.2f depth %d" % (args.neurons, args.dropout, args.depth)
    else:
        net = quiet(UNetpCoord, 1, 1, dev, rule=args.rule, nbf=args.size, batched=True, depth=args.depth, base=args.base)
        model_name = "UNetpCoord (coord-conv Plastic U-Net)"
    net.conv_math = args.math
    net.train()
    group = dist.group.WORLD if world > 1 else None
    dp.attach(net, group)
    dp.broadcast_parameters(net, 0, group)
    B = args.batch
    if args.infer:
        net.eval()
        ts = InferStep(net, B, args.size, use_graph=not args.no_graph)
        ts.loss = torch.zeros(1, device=dev)
        _step = ts.step
        ts.step = lambda x=None, t=None: (_step(x), ts.loss)[1]
    else:
        ts = TrainStep(net, B, args.size, lr=1e-4, use_graph=not args.no_graph, dp_group=group)

    # ---- data: device pool larger than L2 (rotated) + pinned host pool for the e2e leg
    gen = torch.Generator().manual_seed(1234 + rank)
    per_batch = B * args.size * args.size * 4 * 2
    npool = max(4, int(1.5 * L2_BYTES // per_batch) + 1)
    pool = [synth_batch(B, args.size, gen) for _ in range(npool)]
    dpool = [(x.to(dev), t.to(dev)) for x, t in pool]
    hpool = [(x.pin_memory(), t.pin_memory()) for x, t in pool]
    ts.capture()

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- warm-up
    for i in range(args.warmup):
        ts.step(*dpool[i % npool])
    torch.cuda.synchronize()

    # ---- timed: device-resident inputs
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        ts.step(*dpool[i % npool])
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    eager_launches = _lib.launch_count() - launches0
    loss_end = float(ts.loss)
    if not args.infer and not (loss_end == loss_end and abs(loss_end) < 1e3):
        raise SystemExit("bench.py: training diverged / produced a non-finite loss (%r)" % loss_end)

    # ---- timed: end to end from pinned host memory, loss read back every step
    loss_host = torch.zeros(1).pin_memory()
    for i in range(2):
        ts.step(*hpool[i % npool])
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    if hasattr(ts, "prefetch"):
        # every step's pinned-host batch is copied inside the timed region; the copy of batch i+1 (copy engine, own
        # stream) overlaps step i, the loss of every step is read back with a sync
        ts.prefetch(*hpool[0])
        for i in range(args.steps):
            ts.step_prefetched()
            if i + 1 < args.steps:
                ts.prefetch(*hpool[(i + 1) % npool])
            loss_host.copy_(ts.loss, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            _ = float(loss_host)
    else:
        for i in range(args.steps):
            ts.step(*hpool[i % npool])
            loss_host.copy_(ts.loss, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            _ = float(loss_host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---- max over ranks
    times = torch.tensor([ms, e2e_s * 1000.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(times[0]), float(times[1])

    if rank == 0:
        images = B * world * args.steps
        value = images / (ms_max * 1e-3)
        e2e_value = images / (e2e_ms_max * 1e-3)
        kps = ts.kernels_per_step or 0
        roof = dominant_kernel_roofline(B, args.size, 1 if args.math == "tf32" else 0, dev, hbm_peak, peak_src)
        # whole-step figure against the layer-fused algorithmic bound of SURVEY.md §8d (28.39 MB / image @128)
        alg_mb_per_img = (9.49 if args.infer else 28.39) * (args.size / 128.0) ** 2  # UNetp figures; other models: indicative only
        step_gbs = alg_mb_per_img * 1e6 * B / (ms_max / args.steps * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline and args.model == "unetp" and not args.infer and (args.depth, args.base) == (4, 8):
            ips0, _, thr = cpu_train_images_per_s(4, args.size, args.rule, warm=1)
            n = int(min(256, max(8, ips0 * 12)))  # ~12 s of CPU work
            ips, n, thr = cpu_train_images_per_s(n, args.size, args.rule)
            cpu = {"value": ips, "unit": UNIT, "cores": thr, "kind": "port",
                   "sample": "%d sequential single-image train steps (fwd+BCE+bwd+Adam+trace, train.py:91-112) of the "
                             "oracle port, same synthetic 128x128 data" % n}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32" if args.math == "tf32" else "f32", "data": "synthetic",
            "config": {"workload": "%s %s rule, %s, batch %d per GPU, %s"
                                   % (model_name, args.rule, "1x101x101 zero-padded to 128x128" if args.size == 128 else "1x%dx%d" % (args.size, args.size),
                                      B, "batched inference (forward, zero trace)" if args.infer else "fwd+BCE+bwd+Adam+trace update"),
                       "global_batch": B * world, "parallelism": "dp%d" % world, "conv_math": args.math,
                       "cuda_graph": not args.no_graph,
                       "l2": "inputs rotate over a %d-batch device pool (%.0f MB > 126 MB L2); per-step activation "
                             "working set %.0f MB" % (npool, npool * per_batch / 1e6, 9.49 * B * (args.size / 128.0) ** 2)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": per_batch * world, "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": kps * args.steps,
            "kernels_per_step": kps,
            "roofline": roof,
            "step_roofline": {"bound": "hbm", "achieved": step_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": step_gbs / hbm_peak,
                              "note": "whole step vs the layer-fused algorithmic bytes of SURVEY.md 8d (%.2f MB/image)" % alg_mb_per_img},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "final_loss": loss_end,
            "eager_launches_in_timed_region": eager_launches,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear-down: drop the captured graph (it holds NCCL work) before leaving; a hung destroy_process_group must
        # never keep the GPUs busy, so flush and hard-exit after a final barrier
        del ts.graph
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
