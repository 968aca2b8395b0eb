"""Isolated conv3x3 dgrad launch loop (TF32 mode, premasked protocol): python scripts/dgrad_probe.py C0 C1 Cout size batch
(the layer's forward shape: the dgrad is a Cout -> C0|C1 conv with packed-mask epilogue).  PU_TC_DEBUG / PU_TC_FLAT apply."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plastic-unet_b200"))
import torch  # noqa: E402

from pu_b200 import ops  # noqa: E402

C0, C1, Cout, size, B = [int(v) for v in (sys.argv[1:6] + ["8", "8", "8", "128", "64"][len(sys.argv[1:6]):])]
dev = "cuda"
nbuf = 4
xs0 = [torch.rand(B, size, size, C0, device=dev) for _ in range(nbuf)]
xs1 = [torch.rand(B, size, size, C1, device=dev) for _ in range(nbuf)] if C1 else [None] * nbuf
dys = [torch.randn(B, size, size, Cout, device=dev) for _ in range(nbuf)]
ys = [torch.rand(B, size, size, Cout, device=dev) for _ in range(nbuf)]
ms0 = [torch.randint(0, 256, (B, size, size, C0 // 8), device=dev, dtype=torch.uint8) for _ in range(nbuf)]
ms1 = [torch.randint(0, 256, (B, size, size, C1 // 8), device=dev, dtype=torch.uint8) for _ in range(nbuf)] if C1 else [None] * nbuf
w = torch.randn(Cout, C0 + C1, 3, 3, device=dev) * 0.1


def run(i):
    k = i % nbuf
    ops.conv3x3_bwd(dys[k], ys[k], xs0[k], xs1[k], w, False, True, size, size, 0, 0, 0, 0, ops.MATH_TF32, True, False, ms0[k], ms1[k], True)


ITERS = 20
with torch.no_grad():
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            run(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if os.environ.get("NO_GRAPH"):
        sys.exit(0)
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
print("dgrad of %d|%d->%d @%d B=%d: %.1f us/launch, %.0f GB/s algorithmic" % (C0, C1, Cout, size, B, us, gb / (us * 1e-6)))
