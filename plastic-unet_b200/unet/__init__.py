"""Drop-in replacement for the reference's `unet` package (reference src/unet/__init__.py:1-2).

Put `plastic-unet_b200/` on sys.path ahead of the reference's `src/` and
`from unet import UNetp, UNetpRes` in train.py / eval.py / infer.py resolves to the B200 modules.
"""
from pu_b200 import UNetp, UNetpRes, UNetpCoord  # noqa: F401
