"""unetb200 -- host-side Python layer of the B200-native UNet hot path.

``_lib``        ctypes binding of libunetb200.so (C ABI: include/unetb200.h)
``ops``         tensor-level wrappers (pointer / shape / stream marshalling only)
``functional``  torch.autograd.Function per UNet part
``losses``      dice / boundary / fused CE+dice
``ddp``         data-parallel gradient all-reduce (NCCL) overlapped with backward
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops", "functional", "losses", "ddp"]
