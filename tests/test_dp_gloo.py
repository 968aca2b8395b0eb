"""world_size-2 gloo tests (CPU) of the data-parallel host logic (pu_b200/dp.py): batch sharding, the trace-delta
all-reduce + shared epilogue, and the gradient all-reduce reproduce the single-process batched semantics
(oracle recipe 4: world=1,B == world=2,B/2).  Compute on each rank is done with the oracle."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rule, out_path):
    for p in (os.path.join(ROOT, "plastic-unet_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import plastic_unet_oracle as orc
    from conftest import Case
    from pu_b200 import dp
    torch.set_num_threads(1)
    r, w, _ = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    c = Case("unetp_oja_n32")
    sd = orc.leaf_state(c.state_dict())
    if rank != 0:  # replicas must start identical: perturb, then broadcast from rank 0
        with torch.no_grad():
            for v in sd.values():
                if v.is_floating_point():
                    v.add_(1.0)

    class Holder(torch.nn.Module):
        def __init__(self, tensors):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.detach().clone()) for t in tensors if t.is_floating_point()])

    keys = [k for k, v in sd.items() if v.is_floating_point()]
    h = Holder([sd[k] for k in keys])
    dp.broadcast_parameters(h, 0)
    for k, p in zip(keys, h.ps):
        sd[k] = p.detach().clone().requires_grad_(not k.endswith(("running_mean", "running_var")))
    B = 4
    g = torch.Generator().manual_seed(77)
    x = torch.rand(B, 1, 32, 32, generator=g)
    target = (torch.rand(B, 32, 32, generator=g) > 0.6).float()
    hebb = 0.05 * torch.randn(32, 32, generator=g)
    lo, hi = dp.shard_range(B, rank, world)
    # local forward on this rank's shard (shared trace), local mean loss
    maps = orc.unetp_body(sd, x[lo:hi]).view(hi - lo, 32, 32)
    outs = torch.stack([orc.plastic_head(maps[i], sd["w"], sd["alpha"], hebb)[1] for i in range(hi - lo)])
    loss = orc.bce_mean(outs.reshape(-1), target[lo:hi].reshape(-1))
    loss.backward()
    # gradient exchange: one flat arena, sum then / world  (TrainStep does the same with pu_adam_step's grad_scale)
    flat = torch.cat([sd[k].grad.reshape(-1) if sd[k].grad is not None else torch.zeros(sd[k].numel()) for k in keys])
    dp.all_reduce_sum_(flat)
    flat /= world
    # trace exchange: [N*N + N] payload = (sum_k outer(pre_k, post_k), sum_k post_k^2), then the shared epilogue
    pre, post = maps[:, 0, :].detach(), outs[:, 0, :].detach()
    delta_q = torch.cat([torch.einsum("ki,kj->ij", pre, post).reshape(-1), (post ** 2).sum(0)])
    dp.all_reduce_sum_(delta_q)
    hebb_new = dp.trace_epilogue(hebb, delta_q, sd["eta"].detach(), rule, B)
    torch.save({"flat": flat, "hebb": hebb_new, "outs": outs.detach(), "lo": lo}, out_path % rank)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("rule", ["oja", "hebb"])
def test_dp2_matches_single_process(tmp_path, rule):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import plastic_unet_oracle as orc
    from conftest import Case
    world = 2
    out = str(tmp_path / "rank%d.pt")
    mp.spawn(_worker, args=(world, _free_port(), rule, out), nprocs=world, join=True)
    res = [torch.load(out % r) for r in range(world)]
    # ---- single-process reference: the batched extension on the full batch
    c = Case("unetp_oja_n32")
    sd = orc.leaf_state(c.state_dict())
    keys = [k for k, v in sd.items() if v.is_floating_point()]
    B = 4
    g = torch.Generator().manual_seed(77)
    x = torch.rand(B, 1, 32, 32, generator=g)
    target = (torch.rand(B, 32, 32, generator=g) > 0.6).float()
    hebb = 0.05 * torch.randn(32, 32, generator=g)
    _, outs, hebb_ref = orc.forward("unetp", sd, x, hebb, rule=rule)
    orc.bce_mean(outs.reshape(-1), target.reshape(-1)).backward()
    flat_ref = torch.cat([sd[k].grad.reshape(-1) if sd[k].grad is not None else torch.zeros(sd[k].numel()) for k in keys])
    # every rank ends with the same averaged gradient and the same trace, equal to the single-process result
    assert torch.equal(res[0]["flat"], res[1]["flat"])
    assert torch.equal(res[0]["hebb"], res[1]["hebb"])  # bit-identical trace on all ranks
    assert float((res[0]["flat"] - flat_ref).norm() / flat_ref.norm()) < 1e-5
    assert float((res[0]["hebb"] - hebb_ref).abs().max() / hebb_ref.abs().max()) < 1e-5
    for r in range(world):
        lo = res[r]["lo"]
        assert torch.allclose(res[r]["outs"], outs[lo:lo + B // world], atol=1e-6)


def test_shard_range_and_epilogue_errors():
    sys.path.insert(0, os.path.join(ROOT, "plastic-unet_b200"))
    from pu_b200 import dp
    assert dp.shard_range(256, 3, 8) == (96, 128)
    with pytest.raises(ValueError):
        dp.shard_range(10, 0, 4)
    with pytest.raises(ValueError):
        dp.trace_epilogue(torch.zeros(2, 2), torch.zeros(6), 0.1, "bogus", 1)
