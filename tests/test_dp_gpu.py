"""Data parallelism on real GPUs (needs >= 2 devices: `gpurun --gpus 2`; skipped on a 1-GPU box).

world x B data-parallel TrainStep (NCCL gradient all-reduce + the deferred, side-stream trace-delta all-reduce, both
captured in the step's CUDA graph) == one process stepping on the same world*B global batch: losses, parameters after K
steps, trace; and the replicas stay bit-identical (SURVEY.md §8c recipe 4, §8e)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu
if torch.cuda.device_count() < 2:
    pytest.skip("needs >= 2 CUDA devices", allow_module_level=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, math, use_graph, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "plastic-unet_b200"))
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist
    from conftest import quiet
    import pu_b200
    from pu_b200 import dp
    from pu_b200.trainer import TrainStep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dp.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    B, size, steps, lr = 4, 64, 3, 1e-3
    g = torch.Generator().manual_seed(5)
    xs = [torch.rand(B * world, 1, size, size, generator=g) for _ in range(steps)]
    ts_ = [(torch.rand(B * world, size, size, generator=g) < 0.3).float() for _ in range(steps)]
    torch.manual_seed(0)
    net = quiet(pu_b200.UNetp, 1, 1, dev, rule="oja", nbf=size, batched=True)
    net.conv_math = math
    dp.attach(net, dist.group.WORLD)
    dp.broadcast_parameters(net, 0, dist.group.WORLD)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    step = TrainStep(net, B, size, lr=lr, use_graph=use_graph, dp_group=dist.group.WORLD).capture()
    lo, hi = dp.shard_range(B * world, rank, world)
    losses = []
    for x, t in zip(xs, ts_):
        l = step.step(x[lo:hi].to(dev), t[lo:hi].to(dev)).clone()
        dist.all_reduce(l)
        losses.append(float(l) / world)
    torch.cuda.synchronize()
    # replicas bit-identical
    mine = torch.cat([step.hebb.flatten(), step.flat_p]).view(torch.int32)
    ref = mine.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([int(torch.equal(mine, ref))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    res = {"bit_identical": bool(int(same))}
    if rank == 0:
        torch.manual_seed(0)
        net1 = quiet(pu_b200.UNetp, 1, 1, dev, rule="oja", nbf=size, batched=True)
        net1.conv_math = math
        net1.load_state_dict(sd0)
        one = TrainStep(net1, B * world, size, lr=lr, use_graph=use_graph).capture()
        l1 = [float(one.step(x.to(dev), t.to(dev))) for x, t in zip(xs, ts_)]
        torch.cuda.synchronize()
        num = den = 0.0
        for (k, p), (_, q) in zip(net.named_parameters(), net1.named_parameters()):
            num += float((p.double() - q.double()).pow(2).sum())
            den += float((q.double() - sd0[k].double()).pow(2).sum())
        res.update(fused_active=step._fused is not None, losses_dp=losses, losses_1=l1, upd=(num / den) ** 0.5,
                   trace=float((step.hebb - one.hebb).abs().max() / one.hebb.abs().max()))
        torch.save(res, out)
    dist.barrier()
    del step.graph
    torch.cuda.synchronize()
    dist.destroy_process_group()


@pytest.mark.parametrize("math,use_graph,fused", [("fp32", True, 1), ("tf32", True, 1), ("tf32", False, 1), ("tf32", True, 0), ("fp32", False, 0)])
def test_dp2_trainstep_matches_single_process(tmp_path, monkeypatch, math, use_graph, fused):
    """fused = 1: the gradient exchange fused with Adam over NVLink peer memory (pu_adam_allreduce_step, symmetric memory);
    fused = 0: ncclAllReduce + pu_adam_step."""
    import torch.multiprocessing as mp
    world = 2
    out = str(tmp_path / "res.pt")
    monkeypatch.setenv("PU_DP_FUSED", str(fused))
    mp.spawn(_worker, args=(world, _free_port(), math, use_graph, out), nprocs=world, join=True)
    res = torch.load(out)
    print("\n[dp2 %s graph=%s fused=%d] %s" % (math, use_graph, fused, res))
    assert res["fused_active"] == bool(fused)
    assert res["bit_identical"]
    assert max(abs(a - b) for a, b in zip(res["losses_dp"], res["losses_1"])) < (1e-5 if math == "fp32" else 2e-4)
    assert res["trace"] < (1e-4 if math == "fp32" else 1e-3)
    assert res["upd"] < (2e-3 if math == "fp32" else 5e-2)
