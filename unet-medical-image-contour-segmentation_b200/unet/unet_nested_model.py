"""Stub for ``unet.unet_nested_model``: the reference's train.py:16 imports ``UNetPlusPlus_S`` and
``UNetPlusPlus`` from a module that is NOT in the reference repository (``import train`` fails there with
ModuleNotFoundError).  The names exist so that train.py imports unchanged; they are outside the hot path."""


class _Missing:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} is not part of the reference repository nor of the B200 hot path")


class UNetPlusPlus_S(_Missing):
    pass


class UNetPlusPlus(_Missing):
    pass
