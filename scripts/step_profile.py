"""Per-kernel GPU time of one training step in its real context (warm L2, real neighbours) via CUPTI / torch.profiler.

    python scripts/step_profile.py [--math tf32] [--side 0|2] [--graph] [--model unetp] ... > gpurun_out/step_profile.txt

ncu's launch list serialises the kernels and flushes the caches before each one, which inflates every small kernel; this
script traces an ordinary eager (or graph-replayed) TrainStep instead.  Not a benchmark: profiler overhead is in the wall
time, only the per-kernel durations are meaningful."""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "plastic-unet_b200"))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--math", default="tf32")
ap.add_argument("--side", type=int, default=0)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--model", default="unetp")
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--rule", default="oja")
ap.add_argument("--neurons", type=int, default=16)
ap.add_argument("--depth", type=int, default=4)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--order", action="store_true", help="also print the kernels of the last step in launch order")
ap.add_argument("--timeline", action="store_true", help="also print the last step as a per-stream timeline (start, end, stream, kernel)")
args = ap.parse_args()
os.environ["PU_WGRAD_SIDE"] = str(args.side)

from pu_b200.trainer import TrainStep  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
net, name = bench.build_net(args.model, dev, args.rule, args.size, depth=args.depth, neurons=args.neurons)
net.conv_math = args.math
net.train()
ts = TrainStep(net, args.batch, args.size, lr=1e-4, use_graph=args.graph).capture()
gen = torch.Generator().manual_seed(1)
pool = [tuple(t.to(dev) for t in bench.synth_batch(args.batch, args.size, gen)) for _ in range(4)]
for i in range(3):
    ts.step(*pool[i % 4])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(args.steps):
        ts.step(*pool[i % 4])
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.OrderedDict()
for e in evs:
    k = e.name.replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "").replace("pu::", "")
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print("# %s %s rule, %dx%d, B=%d, math=%s, side streams=%d, graph=%s: %d steps, sum of kernel time %.1f us/step"
      % (name, args.rule, args.size, args.size, args.batch, args.math, args.side, args.graph, args.steps, tot / args.steps))
print("%-64s %6s %10s %8s %8s" % ("kernel", "n/step", "us/step", "share", "avg us"))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-64s %6.1f %10.1f %7.1f%% %8.2f" % (k[:64], n / args.steps, t / args.steps, 100 * t / tot, t / n))
if args.order:
    evs.sort(key=lambda e: e.time_range.start)
    per = len(evs) // args.steps
    print("\n# last step in launch order (start offset us, duration us)")
    t0 = evs[-per].time_range.start
    for e in evs[-per:]:
        print("%9.1f %8.2f  %s" % (e.time_range.start - t0, e.device_time if hasattr(e, "device_time") else e.cuda_time,
                                   e.name.split("(")[0].replace("void ", "").replace("pu::", "")[:80]))

if args.timeline:
    # kineto events carry the stream (device_resource_id): one line per kernel, start / end relative to the step's first kernel
    kev = [e for e in prof.profiler.kineto_results.events() if "cuda" in str(e.device_type()).lower()]
    kev.sort(key=lambda e: e.start_ns())
    per = len(kev) // args.steps
    last = kev[-per:]
    t0 = last[0].start_ns()
    streams = {}
    print("\n# last step timeline: start us, end us, stream, kernel   (stream 0 = the step's main stream)")
    for e in last:
        sid = streams.setdefault(e.device_resource_id(), len(streams))
        nm = e.name().replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "").replace("pu::", "")[:60]
        print("%9.1f %9.1f  s%d  %s" % ((e.start_ns() - t0) / 1e3, (e.start_ns() + e.duration_ns() - t0) / 1e3, sid, nm))
