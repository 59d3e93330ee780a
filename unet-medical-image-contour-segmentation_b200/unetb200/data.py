"""Device-side input pipeline: what ``BasicDataset.__getitem__`` does to the decoded
uint8 arrays (utils/data_loading.py:100-132, scale == 1) -- the 90/180/270 degree rotation augmentation (:91-98,
:119-121), ``float32(v) / 255`` when the image holds a value > 1 (:86-87), the gray-level -> class-index mapping of
the mask (:73-79), ``.float()`` / ``.long()`` (:130-131) -- as two kernels per batch on the raw bytes.

The host moves 1 byte per pixel and channel (+1 for the mask) through pinned memory instead of 4 + 8; the fp32
channels_last image batch and the int64 mask batch the training step expects (train.py:113-114) are produced in HBM.
Decoding the PNG files and a ``scale != 1`` BICUBIC resize (data_loading.py:69) stay on the host (PIL): out of scope.
Results are bit-identical to the reference's arrays.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _as_u8_batch(arrays, what):
    """list of equally shaped uint8 arrays (or one stacked array / tensor) -> pinned uint8 tensor [B, H, W(, C)]"""
    if isinstance(arrays, torch.Tensor):
        t = arrays
    else:
        if isinstance(arrays, np.ndarray):
            arr = arrays
        else:
            shapes = {tuple(np.shape(a)) for a in arrays}
            if len(shapes) != 1:
                raise ValueError(f"unetb200.data.{what}: every sample of a batch must have the same shape, got {shapes}")
            arr = np.stack([np.asarray(a) for a in arrays])
        t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.dtype != torch.uint8:
        raise ValueError(f"unetb200.data.{what}: expects uint8 samples (decoded image bytes), got {t.dtype}")
    return t.contiguous()


def check_rotations(rots, B, H, W):
    """-> (list of quarter turns in 0..3, transposed flag).  On a non-square image the turns of one batch must all
    have the same parity, otherwise the samples would not stack (the default collate raises as well)."""
    if rots is None:
        return None, False
    rots = [int(k) % 4 for k in rots]
    if len(rots) != B:
        raise ValueError(f"unetb200.data: {len(rots)} rotations for a batch of {B}")
    odd = {k & 1 for k in rots}
    if H != W and len(odd) > 1:
        raise ValueError("unetb200.data: mixing even and odd quarter turns on non-square images gives samples of "
                         "different shapes")
    return rots, (H != W and odd == {1})


def preprocess_batch(images_u8, masks_u8=None, rots=None, device="cuda", lut=None):
    """images_u8: B decoded images, uint8 [H,W] or [H,W,C] each (``numpy.asarray(PIL image)``); masks_u8: B uint8
    [H,W] gray-level masks or None; rots: B quarter turns (``idx % 4`` of data_loading.py:103: 0, 1, 2, 3 = 0, 90, 180,
    270 degrees counter-clockwise) or None.  Returns ``{'image': float32 [B,C,H',W'] channels_last, 'mask': int64
    [B,H',W']}`` on ``device`` -- the batch dict of train.py:111-114 after its ``.to(device)`` calls."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("unetb200.data.preprocess_batch: the input pipeline runs on CUDA only (no CPU fallback)")
    img = _as_u8_batch(images_u8, "preprocess_batch")
    if img.dim() == 3:
        img = img.unsqueeze(-1)
    if img.dim() != 4:
        raise ValueError(f"unetb200.data.preprocess_batch: images must be [B,H,W] or [B,H,W,C], got {tuple(img.shape)}")
    B, H, W, Cc = img.shape
    rots, transposed = check_rotations(rots, B, H, W)
    Ho, Wo = (W, H) if transposed else (H, W)
    L = ops.lib()
    with torch.cuda.device(dev):
        def up(t):
            t = t if t.is_cuda else (t if t.is_pinned() else t.pin_memory()).to(dev, non_blocking=True)
            return t
        d_img = up(img)
        d_rot = up(torch.tensor(rots, dtype=torch.int32)) if rots is not None else None
        out = torch.empty((B, Ho, Wo, Cc), dtype=torch.float32, device=dev)
        flags = torch.empty(B, dtype=torch.int32, device=dev)
        ops._run("preprocess_image_u8", L.unetb200_preprocess_image_u8, ops._p(d_img), B, H, W, Cc, ops._p(d_rot),
                 int(transposed), ops._p(out), ops._p(flags), ops._stream(), kernels=2,
                 nbytes=float(B * H * W * Cc) * (2 + 4))
        result = {"image": out.permute(0, 3, 1, 2)}
        if masks_u8 is not None:
            msk = _as_u8_batch(masks_u8, "preprocess_batch")
            if tuple(msk.shape) != (B, H, W):
                raise ValueError(f"unetb200.data.preprocess_batch: masks {tuple(msk.shape)} do not match images "
                                 f"{(B, H, W)} (data_loading.py:115)")
            d_msk = up(msk)
            d_lut = None
            if lut is not None:
                d_lut = torch.as_tensor(lut, dtype=torch.int64).reshape(256).contiguous()
                d_lut = up(d_lut)
            mout = torch.empty((B, Ho, Wo), dtype=torch.int64, device=dev)
            ops._run("preprocess_mask_u8", L.unetb200_preprocess_mask_u8, ops._p(d_msk), B, H, W, ops._p(d_rot),
                     int(transposed), ops._p(d_lut), ops._p(mout), ops._stream(), nbytes=float(B * H * W) * (1 + 8))
            result["mask"] = mout
    return result
