"""Input pipeline of the train step: the whole dataset lives on the device, a batch is assembled (and
zero-padded, e.g. 101 -> 128) by ONE kernel from a vector of sample indices, and every rank draws a disjoint shard of a
per-epoch permutation.

The reference converts one float64 numpy image per step to float32 and copies it host -> device (train.py:94-95), from
arrays built by utils/data_set.py:43-44,57-63 ([n, 1, 101, 101] images, [n, 1, 101, 101] masks in {0, 1}).  A TGS-salt
sized set (4 000 images x 101 x 101 fp32 x 2) is 326 MB: 0.2 % of the B200's HBM.
"""
import torch

from . import _lib


def epoch_indices(n, epoch, batch, rank=0, world=1, seed=0, drop_last=True):
    """The sample indices of one epoch for this rank: a seeded permutation of range(n) (the same on every rank), cut into
    global batches of batch*world, of which rank r takes rows [r*batch, (r+1)*batch).  -> int64 tensor [steps, batch].
    Host logic only (CPU tensor); every sample is used at most once per epoch and the shards of the ranks are disjoint."""
    if batch <= 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad (batch, rank, world)")
    g = torch.Generator().manual_seed(int(seed) * 1000003 + int(epoch))
    perm = torch.randperm(n, generator=g)
    gb = batch * world
    steps = n // gb
    if steps == 0:
        raise ValueError("dataset of %d samples is smaller than one global batch of %d" % (n, gb))
    if not drop_last and n % gb:
        raise ValueError("drop_last=False needs n divisible by the global batch (equal shards on every rank)")
    return perm[: steps * gb].view(steps, world, batch)[:, rank, :].contiguous()


class DeviceDataset:
    """images [n, C, Hs, Ws], masks [n, Hs, Ws] (or [n, 1, Hs, Ws]) as float32 on the device; batches are gathered + padded
    into caller-provided static buffers (TrainStep.x / TrainStep.target) so that they can feed a captured CUDA graph."""

    def __init__(self, images, masks, device, pad_to=None):
        images = torch.as_tensor(images)
        masks = torch.as_tensor(masks)
        if masks.dim() == 4:
            masks = masks[:, 0]
        if images.dim() != 4 or masks.dim() != 3 or images.shape[0] != masks.shape[0] or images.shape[-2:] != masks.shape[-2:]:
            raise ValueError("expected images [n, C, H, W] and masks [n, H, W]")
        self.images = images.to(device=device, dtype=torch.float32).contiguous()
        self.masks = masks.to(device=device, dtype=torch.float32).contiguous()
        self.n, self.C, self.Hs, self.Ws = self.images.shape
        self.Hd, self.Wd = (pad_to, pad_to) if pad_to else (self.Hs, self.Ws)
        if self.Hd < self.Hs or self.Wd < self.Ws:
            raise ValueError("pad_to is smaller than the images")
        # 101 -> 128: 13 pixels top/left, 14 bottom/right
        self.oy, self.ox = (self.Hd - self.Hs) // 2, (self.Wd - self.Ws) // 2

    def gather(self, idx, x_out, target_out):
        """x_out [B, C, Hd, Wd], target_out [B, Hd, Wd] <- the (padded) samples idx [B] (int64, device)."""
        if not (idx.is_cuda and idx.dtype == torch.int64 and idx.is_contiguous()):
            raise RuntimeError("DeviceDataset.gather: idx must be a contiguous int64 CUDA tensor")
        B = idx.shape[0]
        if tuple(x_out.shape) != (B, self.C, self.Hd, self.Wd) or tuple(target_out.shape) != (B, self.Hd, self.Wd):
            raise RuntimeError("DeviceDataset.gather: output buffers do not match (B, C, %d, %d)" % (self.Hd, self.Wd))
        st = torch.cuda.current_stream().cuda_stream
        _lib.call("pu_gather_pad", self.images.data_ptr(), idx.data_ptr(), x_out.data_ptr(), B, self.C, self.Hs, self.Ws,
                  self.Hd, self.Wd, self.oy, self.ox, st)
        _lib.call("pu_gather_pad", self.masks.data_ptr(), idx.data_ptr(), target_out.data_ptr(), B, 1, self.Hs, self.Ws,
                  self.Hd, self.Wd, self.oy, self.ox, st)
