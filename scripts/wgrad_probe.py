"""Isolated conv3x3 wgrad launch loop: python scripts/wgrad_probe.py C0 C1 Cout size batch"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plastic-unet_b200"))
import torch  # noqa: E402

from pu_b200 import _lib  # noqa: E402

C0, C1, Cout, size, B = [int(v) for v in (sys.argv[1:6] + ["8", "0", "8", "128", "64"][len(sys.argv[1:6]):])]
MATH = int(os.environ.get("WG_MATH", "1"))
dev = "cuda"
nbuf = 4
xs0 = [torch.rand(B, size, size, C0, device=dev) for _ in range(nbuf)]
xs1 = [torch.rand(B, size, size, C1, device=dev) for _ in range(nbuf)] if C1 else [None] * nbuf
gs = [torch.randn(B, size, size, Cout, device=dev) for _ in range(nbuf)]
dw = torch.empty(Cout, C0 + C1, 3, 3, device=dev)


def run(i):
    x1 = xs1[i % nbuf]
    _lib.call("pu_conv3x3_wgrad", xs0[i % nbuf].data_ptr(), size, size, C0, 0, 0, x1.data_ptr() if C1 else None, size, size, C1, 0, 0,
              gs[i % nbuf].data_ptr(), dw.data_ptr(), None, B, size, size, Cout, MATH, torch.cuda.current_stream().cuda_stream)


ITERS = 20
if os.environ.get("NO_GRAPH"):  # for ncu: plain launches
    for i in range(4):
        run(i)
    torch.cuda.synchronize()
    sys.exit(0)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for i in range(3):
        run(i)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for i in range(ITERS):
        run(i)
graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    graph.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1000 / (5 * ITERS)
gb = B * size * size * (C0 + C1 + Cout) * 4 / 1e9
gf = 2.0 * B * size * size * 9 * (C0 + C1) * Cout / 1e9
print("wgrad %d|%d->%d @%d B=%d: %.1f us/launch, %.0f GB/s algorithmic, %.1f TFLOP/s" % (C0, C1, Cout, size, B, us, gb / (us * 1e-6), gf / (us * 1e-6) / 1e3))
