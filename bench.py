#!/usr/bin/env python
"""bench.py — the Plastic U-Net hot path on B200: train images/sec (fwd + bwd + plastic update + Adam).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels, one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU path on the host cores

Workload at N=1 (BASELINE.json configs[1]): Plastic U-Net (UNetp), Oja rule, 128x128, batch 64 per GPU,
synthetic 1-channel images (uniform [0,1) 101x101 zero-padded to 128x128) and Bernoulli masks, random-init
weights (torch.manual_seed(0)).  N>1: same per-GPU batch (weak scaling), gradient + trace-delta all-reduce.

One "step" = one optimisation step over one batch.  `value` is whole-job images/s with the batch already
resident in HBM (a device pool larger than L2 is rotated through); `e2e` is the same step through the
public API with pinned HOST batches (H2D copy of images+masks and D2H read of the loss inside the timed
region, every step).  Timing: CUDA events on the launching stream, barrier + synchronize on both sides,
max over ranks.  A step is ~1 ms, so the K-step region is timed `--regions` times back to back and the
MEDIAN region is reported (`region_ms` lists them all): one scheduler hiccup must not move the number.
Prints ONE JSON line on rank 0.  Extra objects on that line:
  roofline      the WORST of {forward, dgrad, wgrad} of the largest layer (up4.0), timed live
  kernels       the ten slowest layer kernels of the step (live CUDA-event timings, algorithmic bytes, HBM fraction)
  extras        short legs for BASELINE configs 3, 4, 5 (coord-conv, residual n8 @101 Hebb/Oja, 512^2 depth-5 train+infer)
  dp_check      (N > 1) replicas bit-identical after the timed run; N x B data-parallel step == one-process step on N*B
  tc_demo       per-layer TFLOP/s of the tcgen05 conv on the wide configuration (UNetpRes neurons=64 @256x256) vs the TF32 peak
"""
import argparse
import contextlib
import io
import json
import os
import statistics
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
# CPU arm: the reference's own implementation of the path on the host cores.  Nothing from the product
# (pu_b200) is imported here.  oracle/_ref (bytecode of the unmodified reference, oracle/build_ref.py) when it
# travelled with the snapshot -> kind "reference"; else the oracle port -> kind "port".
# --------------------------------------------------------------------------------------------------
def _cpu_reference_net(size, rule, seed):
    """-> (kind, step_fn) where step_fn(img [1,1,H,W], mask [H,W]) runs ONE train.py:91-112 iteration."""
    import build_ref
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    crit = torch.nn.BCELoss()
    if build_ref.available():
        import ref_loader
        ref = ref_loader.load(os.path.join(ROOT, "oracle", "_ref", "src"))
        net = quiet(ref.unet.UNetp, 1, 1, torch.device("cpu"), rule=rule, nbf=size)  # the reference's own module
        net.train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)
        state = {"hebb": net.initialZeroHebb()}

        def step(img, mask):
            opt.zero_grad()
            y, state["hebb"] = net(img, state["hebb"].detach())
            loss = crit(y.view(-1), mask.view(-1))
            loss.item()
            loss.backward()
            opt.step()
        return "reference", step
    import plastic_unet_oracle as orc
    sd = orc.leaf_state(orc.init_state_unetp(size, seed))
    params = [v for v in sd.values() if v.is_floating_point() and v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)
    state = {"hebb": torch.zeros(size, size)}

    def step(img, mask):
        opt.zero_grad()
        _, y, state["hebb"] = orc.forward("unetp", sd, img, state["hebb"].detach(), rule=rule)
        loss = crit(y.view(-1), mask.view(-1))
        loss.item()
        loss.backward()
        opt.step()
    return "port", step


def cpu_train_images_per_s(n_images, size, rule, warm=3, seed=0):
    """B=1 sequential steps exactly like train.py:91-112; -> (images/s, n_images, threads, kind)."""
    kind, step = _cpu_reference_net(size, rule, seed)
    gen = torch.Generator().manual_seed(1234)
    imgs, msks = synth_batch(n_images + warm, size, gen)
    t0 = None
    for i in range(n_images + warm):
        if i == warm:
            t0 = time.perf_counter()
        step(imgs[i:i + 1], msks[i])
    dt = time.perf_counter() - t0
    return n_images / dt, n_images, os.cpu_count() or 1, kind


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    per_step = args.ref_images_per_step
    cpu_train_images_per_s(max(1, args.warmup) * per_step, args.size, args.rule)
    t0 = time.perf_counter()
    ips, n, threads, kind = cpu_train_images_per_s(args.steps * per_step, args.size, args.rule, warm=0)
    dt = time.perf_counter() - t0
    what = "the unmodified reference module (oracle/_ref)" if kind == "reference" else "the oracle port"
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "UNetp Oja 128x128 (101x101 zero-padded), B=1 sequential steps as train.py:91-112, "
                               "%d images per step" % per_step},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "%d sequential single-image train steps (fwd+BCE+bwd+Adam+trace) of %s" % (n, what)},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# per-kernel roofline table (measured live with CUDA events on the launching stream)
# --------------------------------------------------------------------------------------------------
def _time_graph(fn, iters=20, reps=5):
    """us per call of fn(i) from a CUDA graph of `iters` calls (a Python op call costs ~60 us of host time)."""
    with torch.no_grad():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                fn(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(iters):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * reps)


# UNetp conv3x3 layers: name, C0 (first source), C1 (second source of the fused concat), Cout, side / 128
UNETP_LAYERS = [("inc.2", 8, 0, 8, 1), ("down1.0", 8, 0, 16, 2), ("down1.2", 16, 0, 16, 2), ("down2.0", 16, 0, 32, 4),
                ("down2.2", 32, 0, 32, 4), ("down3.0", 32, 0, 64, 8), ("down3.2", 64, 0, 64, 8), ("down4.0", 64, 0, 64, 16),
                ("up1.0", 64, 64, 32, 8), ("up1.2", 32, 0, 32, 8), ("up2.0", 32, 32, 16, 4), ("up2.2", 16, 0, 16, 4),
                ("up3.0", 16, 16, 8, 2), ("up3.2", 8, 0, 8, 2), ("up4.0", 8, 8, 8, 1), ("up4.2", 8, 0, 8, 1)]


def kernel_table(batch, size, math, dev, hbm_peak, peak_src):
    """Times forward / dgrad / wgrad of every conv3x3 layer of UNetp in isolation (buffers rotating over > L2 for the
    large layers).  Algorithmic bytes per launch: fwd = pixels*(Cin+Cout)*4 (input read once + output
    written once); dgrad = the same (dY read, dX written); wgrad = pixels*(Cin+Cout)*4 (X and dY read once)."""
    from pu_b200 import ops
    rows = []
    for name, C0, C1, Cout, div in UNETP_LAYERS:
        s = size // div
        per = batch * s * s * (C0 + C1 + Cout) * 4
        nbuf = min(8, max(2, int(2 * L2_BYTES // per) + 1))
        xs0 = [torch.rand(batch, s, s, C0, device=dev) for _ in range(nbuf)]
        xs1 = [torch.rand(batch, s, s, C1, device=dev) for _ in range(nbuf)] if C1 else [None] * nbuf
        dys = [torch.randn(batch, s, s, Cout, device=dev) for _ in range(nbuf)]
        ys = [torch.rand(batch, s, s, Cout, device=dev) for _ in range(nbuf)]
        w = torch.randn(Cout, C0 + C1, 3, 3, device=dev) * 0.1
        b = torch.zeros(Cout, device=dev)
        pm = bool(math)  # the premasked-gradient protocol of the TF32 mode (modules.UNetp.forward): packed ReLU masks
        ms0 = [torch.randint(0, 256, (batch, s, s, C0 // 8), device=dev, dtype=torch.uint8) for _ in range(nbuf)] if pm else [None] * nbuf

        def fwd(i):
            ops.conv3x3(xs0[i % nbuf], xs1[i % nbuf], w, b, None, True, s, s, 0, 0, 0, 0, math)

        def dgrad(i):
            ops.conv3x3_bwd(dys[i % nbuf], ys[i % nbuf], xs0[i % nbuf], xs1[i % nbuf], w, False, True, s, s, 0, 0, 0, 0, math,
                            True, False, ms0[i % nbuf], None, pm)

        def wgrad(i):
            ops.conv3x3_bwd(dys[i % nbuf], ys[i % nbuf], xs0[i % nbuf], xs1[i % nbuf], w, True, True, s, s, 0, 0, 0, 0, math,
                            False, True, None, None, pm)

        for kind, fn in (("fwd", fwd), ("dgrad", dgrad), ("wgrad", wgrad)):
            us = _time_graph(fn)
            gbs = per / (us * 1e-6) / 1e9
            rows.append({"name": "conv3x3 %s %s %d%s->%d @%dx%d" % (name, kind, C0, "|%d" % C1 if C1 else "", Cout, s, s),
                         "us": round(us, 2), "algorithmic_bytes": per, "achieved_gbs": round(gbs, 1), "frac": round(gbs / hbm_peak, 4)})
        del xs0, xs1, dys, ys, ms0
    top = sorted(rows, key=lambda r: -r["us"])[:10]
    worst = min((r for r in rows if " up4.0 " in r["name"]), key=lambda r: r["frac"])
    roof = {"bound": "hbm", "kernel": worst["name"] + (" (tf32 tcgen05 / mma.sync)" if math else " (fp32 ffma)"),
            "achieved": worst["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": worst["achieved_gbs"] / hbm_peak,
            "traffic": None, "traffic_note": "not measured by this run; ncu dram bytes of the same kernels are in profiles/",
            "peak_source": peak_src, "us_per_launch": worst["us"], "algorithmic_bytes_per_launch": worst["algorithmic_bytes"],
            "selection": "the worst of {fwd, dgrad, wgrad} of up4.0 (8|8->8 @ full resolution), the largest layer of UNetp"}
    return roof, top, rows


# UNetpRes(neurons=64) @256x256: the wide configuration SURVEY.md §8d names for the tensor-core roofline ("TC demo", AI 326
# FLOP/B): its 3x3 conv layers (unet_p_res.py:142-164,223-238,256-272), C_in (first | second source) -> C_out @ side
TC_DEMO_LAYERS = [("conv1 res blocks", 64, 0, 64, 256), ("conv2.0", 64, 0, 128, 128), ("conv2 res blocks", 128, 0, 128, 128),
                  ("conv3.0", 128, 0, 256, 64), ("conv3 res blocks", 256, 0, 256, 64), ("conv4.0", 256, 0, 512, 32),
                  ("conv4 res blocks", 512, 0, 512, 32), ("mid.0", 512, 0, 1024, 16), ("mid res blocks", 1024, 0, 1024, 16),
                  ("uconv4.0 (cat)", 512, 512, 512, 32), ("uconv3.0 (cat)", 256, 256, 256, 64), ("uconv2.0 (cat)", 128, 128, 128, 128),
                  ("uconv1.0 (cat)", 64, 64, 64, 256)]


def tc_demo(dev, bf16_tflops, batch=8):
    """Per-layer TFLOP/s of the tcgen05 conv (forward and dgrad) on the wide configuration, against the TF32 dense peak taken
    as half the measured bf16 cuBLAS throughput (MEASURED_PEAKS.json).  FLOPs = 2 * 9 * C_in * C_out * pixels."""
    from pu_b200 import ops
    peak = bf16_tflops / 2.0
    rows = []
    tot_f = tot_t = 0.0
    for name, C0, C1, Cout, s in TC_DEMO_LAYERS:
        x0 = torch.rand(batch, s, s, C0, device=dev)
        x1 = torch.rand(batch, s, s, C1, device=dev) if C1 else None
        dy = torch.randn(batch, s, s, Cout, device=dev)
        w = torch.randn(Cout, C0 + C1, 3, 3, device=dev) * 0.05
        b = torch.zeros(Cout, device=dev)
        y = torch.rand(batch, s, s, Cout, device=dev)
        flops = 2.0 * 9 * (C0 + C1) * Cout * batch * s * s

        def fwd(i):
            ops.conv3x3(x0, x1, w, b, None, True, s, s, 0, 0, 0, 0, ops.MATH_TF32)

        def dgrad(i):
            ops.conv3x3_bwd(dy, y, x0, x1, w, False, True, s, s, 0, 0, 0, 0, ops.MATH_TF32, True, False, None, None, True)

        # the weights change once per optimisation step, not per launch: pack them once (as TrainStep does at the start of a
        # step, on a side stream) and time the conv kernel alone; the packing passes are timed separately
        ops.PACK_LOG, ops.PACK_CACHE = [], None
        with torch.no_grad():
            fwd(0)
            dgrad(0)
        log, ops.PACK_LOG = ops.PACK_LOG, None
        pack_us = 0.0
        cache = {}
        for (wt, tr, mth, c0) in log:
            cache[(wt.data_ptr(), tr, mth, c0)] = (ops._pack_w_now(wt, tr, mth, c0), None)
            pack_us += _time_graph(lambda i: ops._pack_w_now(wt, tr, mth, c0), iters=3, reps=2)
        ops.PACK_CACHE = cache
        try:
            for kind, fn in (("fwd", fwd), ("dgrad", dgrad)):
                us = _time_graph(fn, iters=5, reps=3)
                tf = flops / (us * 1e-6) / 1e12
                tot_f += flops
                tot_t += us * 1e-6
                rows.append({"layer": "%s %s %d%s->%d @%dx%d" % (name, kind, C0, "|%d" % C1 if C1 else "", Cout, s, s), "us": round(us, 1),
                             "tflops": round(tf, 1), "frac_of_tf32_peak": round(tf / peak, 3),
                             "flat": bool(ops._tc_flat(batch, s, s, C0 if kind == "fwd" else Cout, C1 if kind == "fwd" else 0,
                                                       Cout if kind == "fwd" else C0 + C1))})
            rows[-1]["weight_pack_us_fwd_plus_dgrad"] = round(pack_us, 1)
        finally:
            ops.PACK_CACHE = None
        del x0, x1, dy, w, y
    agg = tot_f / tot_t / 1e12
    return {"config": "UNetpRes(neurons=64) 3x3 conv layers @256x256, batch %d, TF32 (tcgen05 kind::tf32), conv kernel with pre-packed weights" % batch,
            "tf32_peak_tflops": peak, "peak_source": "bf16_tflops / 2 of MEASURED_PEAKS.json", "aggregate_tflops": round(agg, 1),
            "aggregate_frac": round(agg / peak, 3), "best_frac": max(r["frac_of_tf32_peak"] for r in rows), "layers": rows}


# --------------------------------------------------------------------------------------------------
def build_net(model, dev, rule, size, depth=4, base=8, neurons=16, dropout=0.5):
    from pu_b200 import UNetp, UNetpCoord, UNetpRes
    if model == "unetp":
        net = quiet(UNetp, 1, 1, dev, rule=rule, nbf=size, batched=True, depth=depth, base=base)
        name = "UNetp (Plastic U-Net)" + ("" if (depth, base) == (4, 8) else " depth %d base %d" % (depth, base))
    elif model == "res":
        net = quiet(UNetpRes, 1, 1, dev, neurons=neurons, dropout_ratio=dropout, rule=rule, nbf=size, batched=True, depth=depth)
        name = "UNetpRes (residual Plastic U-Net) neurons %d dropout %.2f depth %d" % (neurons, dropout, depth)
    else:
        net = quiet(UNetpCoord, 1, 1, dev, rule=rule, nbf=size, batched=True, depth=depth, base=base)
        name = "UNetpCoord (coord-conv Plastic U-Net)"
    return net, name


def timed_leg(model, dev, group, world, rule, size, B, math, steps, infer=False, pad_from=None, math_override=None, **kw):
    """A short measurement of another BASELINE config: -> dict(images_per_s (whole job), ms_per_step, ...)."""
    import torch.distributed as dist
    from pu_b200 import dp
    from pu_b200.trainer import InferStep, TrainStep
    torch.manual_seed(0)
    net, name = build_net(model, dev, rule, size, **kw)
    math = math_override or math
    net.conv_math = math
    dp.attach(net, group)
    dp.broadcast_parameters(net, 0, group)
    if infer:
        net.eval()
        ts = InferStep(net, B, size).capture()
    else:
        net.train()
        ts = TrainStep(net, B, size, lr=1e-4, dp_group=group).capture()
    gen = torch.Generator().manual_seed(99 + (dist.get_rank() if world > 1 else 0))
    nb = max(2, min(6, int(1.5 * L2_BYTES // (B * size * size * 8)) + 1))
    pool = [tuple(t.to(dev) for t in synth_batch(B, size, gen, pad_from or size)) for _ in range(nb)]

    def one(i):
        if infer:
            ts.step(pool[i % nb][0])
        else:
            ts.step(*pool[i % nb])
    for i in range(3):
        one(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one(i)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / steps
    out = {"model": name, "rule": rule, "size": size, "batch_per_gpu": B, "global_batch": B * world, "mode": "infer" if infer else "train",
           "conv_math": math, "images_per_s": B * world / (ms * 1e-3), "ms_per_step": ms, "steps": steps,
           "kernels_per_step": ts.kernels_per_step}
    if not infer:
        out["final_loss"] = float(ts.loss)
    del ts, net, pool
    torch.cuda.empty_cache()
    return out


def run_extras(dev, group, world, math, steps=10):
    """BASELINE.json configs[2..4], kept out of `value`: coord-conv DP; residual script variant (neurons 8 @101, Hebb and
    Oja, 32 per GPU = 256 on 8 GPUs); scaled 512x512 depth-5 Oja, training + inference."""
    legs = {}
    specs = [
        # the headline workload in the other two math modes: strict fp32 (CUDA-core FFMA kernels: the mode that meets the 1e-3
        # gradient parity) and mixed (strict-fp32 forward = exact ReLU masks, TF32 tensor-core backward)
        ("config2_unetp_oja_128_strict_fp32_mode", dict(model="unetp", rule="oja", size=128, B=64, pad_from=101, math_override="fp32")),
        ("config2_unetp_oja_128_mixed_mode", dict(model="unetp", rule="oja", size=128, B=64, pad_from=101, math_override="mixed")),
        ("config3_coordconv_oja_128", dict(model="coord", rule="oja", size=128, B=64, pad_from=101)),
        ("config4_res_n8_hebb_101", dict(model="res", rule="hebb", size=101, B=32, neurons=8)),
        ("config4_res_n8_oja_101", dict(model="res", rule="oja", size=101, B=32, neurons=8)),
        ("config5_unetp_depth5_oja_512_train", dict(model="unetp", rule="oja", size=512, B=8, depth=5)),
        ("config5_unetp_depth5_oja_512_infer", dict(model="unetp", rule="oja", size=512, B=8, depth=5, infer=True)),
    ]
    for key, kw in specs:
        try:
            legs[key] = timed_leg(dev=dev, group=group, world=world, math=math, steps=steps, **kw)
        except Exception as e:  # an extra must never take the headline down with it
            legs[key] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    return legs


def dp_check(net, ts, dev, group, world, rank, math):
    """(1) replicas bit-identical after the timed run: trace and flat parameters (a checksum of the raw bits);
    (2) one data-parallel step over a fixed global batch (8 images per rank) == one single-process step over the same
    world*8 images on rank 0, from the same weights.  -> dict (rank 0) or None."""
    import torch.distributed as dist
    from pu_b200 import dp
    from pu_b200.trainer import TrainStep

    def bits_sum(t):
        return t.detach().contiguous().view(torch.int32).to(torch.int64).sum().view(1)

    mine = torch.cat([bits_sum(ts.hebb), bits_sum(ts.flat_p)])
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    identical = all(bool(torch.equal(c, allc[0])) for c in allc)
    # (2)
    b_loc, size = 8, net.nbf
    gen = torch.Generator().manual_seed(4242)
    gx, gt = synth_batch(b_loc * world, size, gen)
    torch.manual_seed(7)
    net_dp, _ = build_net("unetp", dev, net.rule, size)
    net_dp.conv_math = math
    dp.attach(net_dp, group)
    dp.broadcast_parameters(net_dp, 0, group)
    sd0 = {k: v.detach().clone() for k, v in net_dp.state_dict().items()}
    ts_dp = TrainStep(net_dp, b_loc, size, lr=1e-3, dp_group=group).capture()
    lo, hi = dp.shard_range(b_loc * world, rank, world)
    loss_dp = ts_dp.step(gx[lo:hi].to(dev), gt[lo:hi].to(dev)).clone()
    dist.all_reduce(loss_dp, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize()
    res = None
    if rank == 0:
        net_1, _ = build_net("unetp", dev, net.rule, size)
        net_1.conv_math = math
        net_1.load_state_dict(sd0)
        ts_1 = TrainStep(net_1, b_loc * world, size, lr=1e-3, dp_group=None).capture()
        loss_1 = float(ts_1.step(gx.to(dev), gt.to(dev)))
        torch.cuda.synchronize()
        num = den = 0.0
        for (k, p), (_, q) in zip(net_dp.named_parameters(), net_1.named_parameters()):
            num += float((p.double() - q.double()).pow(2).sum())
            den += float((q.double() - sd0[k].double()).pow(2).sum())
        trace_err = float((ts_dp.hebb - ts_1.hebb).abs().max() / ts_1.hebb.abs().max().clamp_min(1e-30))
        res = {"replicas_bit_identical_after_run": identical, "global_batch": b_loc * world,
               "dp_vs_single_process": {"loss_dp_mean": float(loss_dp) / world, "loss_single": loss_1,
                                        "update_l2_rel": (num / max(den, 1e-300)) ** 0.5, "trace_max_rel": trace_err},
               "tolerance": "fp32 atomics / all-reduce order: update 2e-3 (fp32) / 5e-2 (tf32, TF32 rounding flips), trace 1e-4"}
        del ts_1, net_1
    del ts_dp, net_dp
    dist.barrier()
    return res


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--regions", type=int, default=11, help="how many times the K-step region is timed (median reported)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--rule", default="oja")
    ap.add_argument("--math", default=os.environ.get("PU_CONV_MATH", "tf32"), choices=["fp32", "tf32", "mixed"])
    ap.add_argument("--model", default="unetp", choices=["unetp", "res", "coord"], help="configs[1] is unetp; the others are the variants of configs[2..4]")
    ap.add_argument("--neurons", type=int, default=16)
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--base", type=int, default=8)
    ap.add_argument("--dropout", type=float, default=0.5)
    ap.add_argument("--infer", action="store_true", help="time the batched forward-only step instead of the train step")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the legs for BASELINE configs 3-5, the kernel table and dp_check")
    ap.add_argument("--dp-check", action="store_true", help="run dp_check (N > 1) even with --no-extras")
    ap.add_argument("--ref-images-per-step", type=int, default=16)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return

    import torch.distributed as dist
    from pu_b200 import _lib, dp
    from pu_b200.trainer import InferStep, TrainStep

    rank, world, local_rank = dp.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    hbm_peak, bf16_peak, peak_src = measured_peaks()
    headline = (args.model == "unetp" and not args.infer and (args.depth, args.base) == (4, 8))

    torch.manual_seed(0)
    net, model_name = build_net(args.model, dev, args.rule, args.size, depth=args.depth, base=args.base, neurons=args.neurons,
                                dropout=args.dropout)
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

    # ---- timed: device-resident inputs; the K-step region is repeated `regions` times
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    region_ms = []
    it = 0
    for r in range(max(1, args.regions)):
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            ts.step(*dpool[it % npool])
            it += 1
        e1.record()
        torch.cuda.synchronize()
        barrier()
        region_ms.append(e0.elapsed_time(e1))
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
    e2e_regions = []
    for r in range(max(1, min(args.regions, 5))):
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
        e2e_regions.append((time.perf_counter() - t0) * 1000.0)
        barrier()

    # ---- the same step fed from a device-resident dataset (pu_b200.data): only the B sample indices cross PCIe per step
    dd_ms = None
    if hasattr(ts, "step_indices") and args.size == 128:
        from pu_b200.data import DeviceDataset, epoch_indices
        gen2 = torch.Generator().manual_seed(77 + rank)
        ds = DeviceDataset(torch.rand(512, 1, 101, 101, generator=gen2), (torch.rand(512, 101, 101, generator=gen2) < 0.25).float(), dev, pad_to=128)
        idx = epoch_indices(512, 0, B, seed=rank).pin_memory()
        for i in range(2):
            ts.step_indices(ds, idx[i % idx.shape[0]])
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            ts.step_indices(ds, idx[i % idx.shape[0]])
            loss_host.copy_(ts.loss, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            _ = float(loss_host)
        dd_ms = (time.perf_counter() - t0) * 1000.0 / args.steps
        del ds

    # ---- max over ranks, region by region; then the median region
    times = torch.tensor(region_ms + e2e_regions, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    region_ms = [float(v) for v in times[:len(region_ms)]]
    e2e_regions = [float(v) for v in times[len(region_ms):]]
    ms_med, e2e_ms_med = statistics.median(region_ms), statistics.median(e2e_regions)

    check = extras = None
    if world > 1 and headline and (not args.no_extras or args.dp_check):
        check = dp_check(net, ts, dev, group, world, rank, args.math)
    if headline and not args.no_extras:
        extras = run_extras(dev, group, world, args.math)

    if rank == 0:
        images = B * world * args.steps
        value = images / (ms_med * 1e-3)
        e2e_value = images / (e2e_ms_med * 1e-3)
        kps = ts.kernels_per_step or 0
        roof = top = demo = None
        if not args.no_extras:
            roof, top, _ = kernel_table(B, args.size, 1 if args.math == "tf32" else 0, dev, hbm_peak, peak_src)
            if headline and args.math == "tf32":
                try:
                    demo = tc_demo(dev, bf16_peak)
                except Exception as e:
                    demo = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        # whole-step figure against the layer-fused algorithmic bound of SURVEY.md §8d (28.39 MB / image @128)
        alg_mb_per_img = (9.49 if args.infer else 28.39) * (args.size / 128.0) ** 2  # UNetp figures; other models: indicative only
        step_gbs = alg_mb_per_img * 1e6 * B / (ms_med / args.steps * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline and headline:
            ips0, _, thr, _ = cpu_train_images_per_s(4, args.size, args.rule, warm=1)
            n = int(min(256, max(8, ips0 * 12)))  # ~12 s of CPU work
            ips, n, thr, kind = cpu_train_images_per_s(n, args.size, args.rule)
            cpu = {"value": ips, "unit": UNIT, "cores": thr, "kind": kind,
                   "sample": "%d sequential single-image train steps (fwd+BCE+bwd+Adam+trace, train.py:91-112) of %s, same synthetic "
                             "128x128 data" % (n, "the unmodified reference module (oracle/_ref)" if kind == "reference" else "the oracle port")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_med / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"tf32": "tf32", "mixed": "f32 forward / tf32 backward", "fp32": "f32"}[args.math], "data": "synthetic",
            "config": {"workload": "%s %s rule, %s, batch %d per GPU, %s"
                                   % (model_name, args.rule, "1x101x101 zero-padded to 128x128" if args.size == 128 else "1x%dx%d" % (args.size, args.size),
                                      B, "batched inference (forward, zero trace)" if args.infer else "fwd+BCE+bwd+Adam+trace update"),
                       "global_batch": B * world, "parallelism": "dp%d" % world, "conv_math": args.math,
                       "cuda_graph": not args.no_graph,
                       "timing": "median of %d back-to-back regions of exactly %d steps each (CUDA events, max over ranks per region)"
                                 % (len(region_ms), args.steps),
                       "l2": "inputs rotate over a %d-batch device pool (%.0f MB > 126 MB L2); per-step activation "
                             "working set %.0f MB" % (npool, npool * per_batch / 1e6, 9.49 * B * (args.size / 128.0) ** 2)},
            "region_ms": [round(v, 3) for v in region_ms],
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": per_batch * world, "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": e2e_ms_med / args.steps, "region_ms": [round(v, 3) for v in e2e_regions]},
            "e2e_device_dataset": None if dd_ms is None else {
                "value": B * world / (dd_ms * 1e-3), "unit": UNIT, "ms_per_step": dd_ms, "h2d_bytes_per_step": 8 * B * world,
                "note": "rank-0 wall clock; dataset resident in HBM (pu_b200.data.DeviceDataset), batch gathered + zero-padded 101->128 on "
                        "the device, loss read back every step"},
            "gpu_launches": kps * args.steps,
            "kernels_per_step": kps,
            "roofline": roof,
            "kernels": top,
            "step_roofline": {"bound": "hbm", "achieved": step_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": step_gbs / hbm_peak,
                              "note": "whole step vs the layer-fused algorithmic bytes of SURVEY.md 8d (%.2f MB/image)" % alg_mb_per_img},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "final_loss": loss_end,
            "eager_launches_in_timed_region": eager_launches,
            "dp_check": check,
            "extras": extras,
            "tc_demo": demo,
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
