"""Input pipeline: per-rank epoch sampler (host logic, CPU) and the device-resident dataset's
gather + zero-pad kernel (GPU), against the reference's own per-step input construction (train.py:94-95) and the
padding rule of BASELINE configs (oracle.pad_101_to_128)."""
import numpy as np
import pytest
import torch

import plastic_unet_oracle as orc
from conftest import quiet


def test_epoch_indices_partition_the_dataset():
    from pu_b200.data import epoch_indices
    n, batch, world = 1003, 16, 4
    shards = [epoch_indices(n, 3, batch, r, world, seed=7) for r in range(world)]
    steps = n // (batch * world)
    assert all(tuple(s.shape) == (steps, batch) for s in shards)
    flat = torch.cat([s.reshape(-1) for s in shards])
    assert flat.unique().numel() == flat.numel() == steps * batch * world  # disjoint, no repeats
    assert int(flat.min()) >= 0 and int(flat.max()) < n
    # deterministic per (seed, epoch), different between epochs, identical permutation on every rank
    assert torch.equal(shards[1], epoch_indices(n, 3, batch, 1, world, seed=7))
    assert not torch.equal(shards[1], epoch_indices(n, 4, batch, 1, world, seed=7))
    one = epoch_indices(n, 3, batch * world, 0, 1, seed=7)  # the single-process global batches
    assert torch.equal(one.view(steps, world, batch)[:, 2, :], shards[2])
    with pytest.raises(ValueError):
        epoch_indices(10, 0, 16, 0, 1)
    with pytest.raises(ValueError):
        epoch_indices(100, 0, 8, 4, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("Hs,pad", [(101, 128), (37, None), (21, 32)])
def test_gather_pad_matches_the_reference_input_construction(Hs, pad):
    from pu_b200.data import DeviceDataset
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(Hs)
    X = np.asarray(torch.rand(23, 1, Hs, Hs, generator=g).double())       # float64 like utils/data_set.py:43
    Y = np.asarray((torch.rand(23, 1, Hs, Hs, generator=g) > 0.6).double())
    ds = DeviceDataset(X, Y, dev, pad_to=pad)
    idx = torch.tensor([5, 0, 22, 5, 13], dtype=torch.int64, device=dev)
    side = pad or Hs
    x_out = torch.full((5, 1, side, side), -1.0, device=dev)
    t_out = torch.full((5, side, side), -1.0, device=dev)
    ds.gather(idx, x_out, t_out)
    for b, i in enumerate(idx.tolist()):
        t_img = torch.from_numpy(np.array([X[i].astype(np.float32)]))  # train.py:94
        y_t = torch.from_numpy(Y[i].astype(np.float32))                # train.py:95
        if pad:
            p0 = (pad - Hs) // 2
            t_img = torch.nn.functional.pad(t_img, (p0, pad - Hs - p0, p0, pad - Hs - p0))
            y_t = torch.nn.functional.pad(y_t, (p0, pad - Hs - p0, p0, pad - Hs - p0))
            if (Hs, pad) == (101, 128):
                assert torch.equal(t_img, orc.pad_101_to_128(torch.from_numpy(np.array([X[i].astype(np.float32)]))))
        assert torch.equal(x_out[b].cpu(), t_img[0]) and torch.equal(t_out[b].cpu(), y_t[0])


@pytest.mark.gpu
def test_step_indices_equals_step_on_host_batches():
    """TrainStep.step_indices(dataset, idx) == TrainStep.step(x[idx], target[idx]) (bitwise: same kernels, same inputs)."""
    import pu_b200
    from pu_b200.data import DeviceDataset, epoch_indices
    from pu_b200.trainer import TrainStep
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    X = torch.rand(40, 1, 27, 27, generator=g)
    Y = (torch.rand(40, 27, 27, generator=g) > 0.5).float()
    ds = DeviceDataset(X, Y, dev, pad_to=32)
    idx = epoch_indices(40, 0, 8, seed=1)
    res = []
    for mode in ("indices", "host"):
        torch.manual_seed(0)
        net = quiet(pu_b200.UNetp, 1, 1, dev, rule="oja", nbf=32, batched=True)
        ts = TrainStep(net, 8, 32, lr=1e-3).capture()
        losses = []
        for s in range(idx.shape[0]):
            if mode == "indices":
                losses.append(float(ts.step_indices(ds, idx[s])))
            else:
                xb = torch.nn.functional.pad(X[idx[s]], (2, 3, 2, 3))
                tb = torch.nn.functional.pad(Y[idx[s]], (2, 3, 2, 3))
                losses.append(float(ts.step(xb.to(dev), tb.to(dev))))
        res.append((losses, ts.flat_p.clone(), ts.hebb.clone()))
    assert np.allclose(res[0][0], res[1][0], rtol=0, atol=1e-6)
    assert float((res[0][1] - res[1][1]).abs().max()) < 1e-6 and float((res[0][2] - res[1][2]).abs().max()) < 1e-7
