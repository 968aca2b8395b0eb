"""Isolated plastic-head launches for ncu captures: python scripts/head_probe.py [N] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plastic-unet_b200"))
import torch  # noqa: E402

from pu_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = "cuda"
X = torch.randn(B * N, N, device=dev)
w = 0.01 * torch.randn(N, N, device=dev)
alpha = 0.01 * torch.rand(N, N, device=dev)
hebb = 0.05 * torch.randn(N, N, device=dev)
gS = torch.randn(B * N, N, device=dev)
with torch.no_grad():
    for _ in range(3):
        S, weff = ops.plastic_head(X, w, alpha, hebb)
        ops.plastic_head_bwd(gS, X, S, weff, alpha, hebb, True, True, False)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            S, weff = ops.plastic_head(X, w, alpha, hebb)
            ops.plastic_head_bwd(gS, X, S, weff, alpha, hebb, True, True, False)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
print("plastic head fwd + bwd N=%d B=%d: %.1f us per (fwd + bwd) pair" % (N, B, e0.elapsed_time(e1) * 1000 / 50))
