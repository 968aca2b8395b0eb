"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  The path shards by batch: every rank holds a full replica of the weights
and of the plastic trace; per step there is one exchange — the gradient all-reduce (one flat arena, mean)
and the trace-delta all-reduce ([N*N + N] floats: sum_k outer(pre_k, post_k) and sum_k post_k^2), after
which every rank applies the identical epilogue, so the trace stays bit-identical on all ranks.

This module is host logic only (no kernels) so that it can be exercised with gloo on CPU.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's env (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).
    -> (rank, world, local_rank).  Single-process when WORLD_SIZE is absent or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def attach(net, group=None):
    """Switch a Plastic U-Net module to the data-parallel trace (split form + all-reduce of the delta)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        net.dp_group, net.dp_world = None, 1
        return net
    net.dp_group = group if group is not None else dist.group.WORLD
    net.dp_world = dist.get_world_size(group)
    return net


def shard_range(global_batch, rank, world):
    """Contiguous, even batch shard of this rank: [lo, hi).  The global batch must divide evenly — the trace
    epilogue divides by the global sample count and the loss is a mean of equal-sized local means."""
    if global_batch % world != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def broadcast_parameters(net, src=0, group=None):
    """Make every replica start from rank `src`'s weights and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(net.parameters()) + list(net.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def all_reduce_sum_(t, group=None):
    """In-place sum over ranks (no-op single-process).  Used for the trace delta and the flat gradient arena."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def trace_epilogue(hebb, delta_q, eta, rule, k_global):
    """Reference-side statement of what pu_trace_apply computes after the all-reduce (host tensors; used by the
    CPU tests to pin the DP semantics):  hebb: (1-eta)*hebb + eta*delta/K ; oja: hebb*(1 - eta*q/K) + eta*delta/K."""
    n = hebb.shape[0]
    delta = delta_q[: n * n].view(n, n)
    q = delta_q[n * n:]
    if rule == "hebb":
        return (1 - eta) * hebb + eta * delta / k_global
    if rule == "oja":
        return hebb * (1 - eta * q / k_global)[None, :] + eta * delta / k_global
    raise ValueError("Must select one learning rule ('hebb' or 'oja')")
