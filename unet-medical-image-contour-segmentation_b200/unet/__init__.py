"""Same import surface as the reference's ``unet/__init__.py`` (``from unet import UNet_S, UNet``, train.py:14)."""
from .unet_model import UNet, UNet_S  # noqa: F401
