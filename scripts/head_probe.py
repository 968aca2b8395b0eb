"""Isolated plastic-head launches (CUDA events over graphs of 10): python scripts/head_probe.py [N] [B]
Times the separate strict-fp32 form (head GEMM, BCE, sigmoid backward, gX GEMM, gW GEMM) against the fused training-step
form (pu_plastic_head_bce + pu_plastic_head_wgrad_tc)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plastic-unet_b200"))
import torch  # noqa: E402

from pu_b200 import _lib, ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = "cuda"
X = torch.randn(B * N, N, device=dev)
w = 0.01 * torch.randn(N, N, device=dev)
alpha = 0.01 * torch.rand(N, N, device=dev)
hebb = 0.05 * torch.randn(N, N, device=dev)
T = (torch.rand(B * N, N, device=dev) > 0.5).float()
loss = torch.zeros(1, device=dev)
gS = torch.empty_like(X)


def timed(what, fn, reps=10):
    with torch.no_grad():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    print("%-58s N=%d B=%d: %6.1f us" % (what, N, B, e0.elapsed_time(e1) * 1000 / (5 * reps)))


def separate_critical():
    S, weff = ops.plastic_head(X, w, alpha, hebb)
    _lib.call("pu_bce_fwd_bwd", S.data_ptr(), T.data_ptr(), loss.data_ptr(), gS.data_ptr(), S.numel(), torch.cuda.current_stream().cuda_stream)
    return ops.plastic_head_bwd(gS, X, S, weff, alpha, hebb, True, True, False)


S0, weff0 = ops.plastic_head(X, w, alpha, hebb)
gA0 = torch.randn_like(X)
timed("separate: head fwd + bce + (sigmoid bwd, gX, gW, galpha)", separate_critical)
timed("separate: head fwd only", lambda: ops.plastic_head(X, w, alpha, hebb))
timed("fused: pu_plastic_head_bce (fwd + loss + gA + gX)", lambda: ops.plastic_head_bce(X, w, alpha, hebb, T, True))
wf = ops.head_weff(w, alpha, hebb)
timed("fused, Weff computed beforehand", lambda: ops.plastic_head_bce(X, w, alpha, hebb, T, True, wf))
timed("fused: pu_plastic_head_bce without gX", lambda: ops.plastic_head_bce(X, w, alpha, hebb, T, False))
timed("parameter gradients: pu_plastic_head_wgrad_tc (split-K, %d term(s))" % ops.HEAD_WGRAD_TERMS, lambda: ops.plastic_head_wgrad(X, gA0, alpha, hebb, True, False))
