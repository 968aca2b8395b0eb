"""Isolated conv3x3 launch loop for ncu captures: python scripts/conv_probe.py [tf32|fp32] [C0] [C1] [Cout] [size] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plastic-unet_b200"))
import torch  # noqa: E402

from pu_b200 import ops  # noqa: E402

math = ops.MATH_TF32 if (len(sys.argv) < 2 or sys.argv[1] == "tf32") else ops.MATH_FP32
C0, C1, Cout, size, B = [int(v) for v in (sys.argv[2:7] + ["8", "8", "8", "128", "64"][len(sys.argv[2:7]):])]
dev = "cuda"
nbuf = 4
xs0 = [torch.rand(B, size, size, C0, device=dev) for _ in range(nbuf)]
xs1 = [torch.rand(B, size, size, C1, device=dev) for _ in range(nbuf)] if C1 else [None] * nbuf
w = torch.randn(Cout, C0 + C1, 3, 3, device=dev) * 0.1
b = torch.zeros(Cout, device=dev)
ITERS = 20
with torch.no_grad():
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            ops.conv3x3(xs0[i % nbuf], xs1[i % nbuf], w, b, None, True, size, size, 0, 0, 0, 0, math)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    # the op call is launch-bound from Python (~60 us of host time per call): capture ITERS calls into one CUDA graph
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(ITERS):
            ops.conv3x3(xs0[i % nbuf], xs1[i % nbuf], w, b, None, True, size, size, 0, 0, 0, 0, math)
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
print("conv3x3 %s %d|%d->%d @%d B=%d: %.1f us/launch (pack+conv, graph replay), %.0f GB/s algorithmic"
      % (sys.argv[1] if len(sys.argv) > 1 else "tf32", C0, C1, Cout, size, B, us, gb / (us * 1e-6)))
