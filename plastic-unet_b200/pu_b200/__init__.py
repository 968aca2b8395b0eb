"""plastic-unet_b200 — B200-native Plastic U-Net hot path (drop-in for the reference's `unet` package).

    from pu_b200 import UNetp, UNetpRes, UNetpCoord

The modules keep the reference constructors, ``forward(x, hebb) -> (activout, hebb')``,
``initialZeroHebb()`` and ``state_dict`` keys (reference src/unet/unet_p.py, unet_p_res.py); every op
underneath is a torch custom op over hand-written sm_100a kernels (see ops.py, ../csrc).
"""
from .modules import UNetp, UNetpRes, UNetpCoord  # noqa: F401

__all__ = ["UNetp", "UNetpRes", "UNetpCoord"]
