"""Device-side tails of ``evaluate.py`` and ``predict.py``.

The reference turns the logits into a score / a label map with a chain of small ATen kernels and full-size
temporaries (``argmax`` -> ``==`` -> ``.float()`` twice -> three reductions -> ``where`` ...; ``F.interpolate`` writes
a second full-resolution logits tensor before ``argmax`` reads it back).  Here each tail is ONE pass over the
logits through libunetb200.so:

  * :func:`argmax_class_dice`   evaluate.py:111-117   argmax + class-c counts -> dice_coeff
  * :func:`binary_dice`         evaluate.py:56-66     sigmoid threshold + counts -> dice_coeff
  * :func:`resize_argmax`       predict.py:26-27      bilinear resize (align_corners=False) + argmax
  * :func:`evaluate`            evaluate.py:12-173    the reference's evaluate() with the tail on the device
  * :func:`predict_img`         predict.py:15-29      the reference's predict_img() with the tail on the device

Index / count results are exact; the dice is the reference's fp32 formula on the exact counts.  No CPU fallback.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import BF16, F32, I64, U8

_IDX = {torch.int64: I64, torch.uint8: U8}


def _logits_arg(logits, what):
    ops.require_cuda(logits, what)
    if logits.dim() != 4:
        raise ValueError(f"unetb200.{what}: logits must be [B, C, H, W], got {tuple(logits.shape)}")
    if logits.dtype not in (torch.float32, torch.bfloat16):
        logits = logits.float()              # fp16 autocast output of a foreign model: widen, never fall back
    return logits, (BF16 if logits.dtype == torch.bfloat16 else F32)


def _target_arg(mask_true, shape, what):
    if mask_true is None:
        return None, F32
    ops.require_cuda(mask_true, what)
    if tuple(mask_true.shape) != tuple(shape):
        raise ValueError(f"unetb200.{what}: target shape {tuple(mask_true.shape)} != {tuple(shape)}")
    if mask_true.dtype not in (torch.float32, torch.int64):
        mask_true = mask_true.float() if mask_true.is_floating_point() else mask_true.long()
    return mask_true.contiguous(), (I64 if mask_true.dtype == torch.int64 else F32)


def _eval_counts(logits, mask_true, cls, mode, index_dtype, epsilon, what):
    logits, ldt = _logits_arg(logits, what)
    B, Cc, H, W = logits.shape
    tgt, tdt = _target_arg(mask_true, (B, H, W), what)
    dev = logits.device
    pred = torch.empty((B, H, W), dtype=index_dtype, device=dev) if index_dtype is not None else None
    counts = torch.empty((B, 4), dtype=torch.int64, device=dev)
    dice = torch.empty(1, dtype=torch.float32, device=dev) if tgt is not None else None
    ops._run("eval_counts", ops.lib().unetb200_eval_counts, ops._p(logits), ldt, *logits.stride(), ops._p(tgt), tdt,
             B, Cc, H, W, int(cls), int(mode), ops._p(pred), _IDX.get(index_dtype, I64), ops._p(counts),
             float(epsilon), ops._p(dice), ops._stream(), kernels=2 if dice is not None else 1,
             nbytes=float(B * H * W) * (Cc * logits.element_size() + (tgt.element_size() if tgt is not None else 0)
                                        + (pred.element_size() if pred is not None else 0)))
    return pred, (dice.reshape(()) if dice is not None else None), counts


def argmax_class_dice(mask_pred, mask_true, c=2, index_dtype=torch.int64, epsilon=1e-6):
    """evaluate.py:111-117 in one pass: returns ``(mask_pred_indices, current_dice, counts)`` where
    ``mask_pred_indices = mask_pred.argmax(dim=1)`` (``index_dtype`` int64 like torch, or uint8 -- what the
    post-processing converts it to anyway, evaluate.py:128; ``None`` skips the write), ``current_dice =
    dice_coeff((indices == c).float(), (mask_true == c).float(), reduce_batch_first=False)`` as a 0-d device
    tensor (no host sync) and ``counts`` int64 [B,4] = per image {|pred & true|, |pred|, |true|, 0}."""
    return _eval_counts(mask_pred, mask_true, c, 0, index_dtype, epsilon, "argmax_class_dice")


def binary_dice(mask_pred, mask_true, index_dtype=torch.uint8, epsilon=1e-6):
    """evaluate.py:56-66 (n_classes == 1) in one pass: ``mask_true // 2``, ``sigmoid(mask_pred.squeeze(1)) > 0.5``
    and ``dice_coeff(binary, mask_true, reduce_batch_first=False)``.  The dice is NaN when a target value is outside
    [0, 4) (the reference's AssertionError at :57, reported without a host sync); ``counts[:, 3]`` holds the number
    of such pixels."""
    if mask_pred.dim() != 4 or mask_pred.shape[1] != 1:
        raise ValueError("unetb200.binary_dice: expects [B, 1, H, W] logits")
    return _eval_counts(mask_pred, mask_true, 1, 1, index_dtype, epsilon, "binary_dice")


def resize_argmax(mask_pred, size, index_dtype=torch.int64):
    """predict.py:26-27: ``F.interpolate(mask_pred, size, mode='bilinear').argmax(dim=1)`` without materialising
    the resized logits.  Returns [B, H, W] ``index_dtype`` (int64 or uint8)."""
    logits, ldt = _logits_arg(mask_pred, "resize_argmax")
    B, Cc, h, w = logits.shape
    H, W = int(size[0]), int(size[1])
    if index_dtype not in _IDX:
        raise ValueError("unetb200.resize_argmax: index_dtype must be torch.int64 or torch.uint8")
    out = torch.empty((B, H, W), dtype=index_dtype, device=logits.device)
    ops._run("resize_argmax", ops.lib().unetb200_resize_argmax, ops._p(logits), ldt, *logits.stride(), B, Cc, h, w,
             H, W, ops._p(out), _IDX[index_dtype], ops._stream(),
             nbytes=float(B) * (Cc * h * w * logits.element_size() + H * W * out.element_size()))
    return out


# ------------------------------------------------------------------------------------------------
# reference-shaped entry points
# ------------------------------------------------------------------------------------------------
@torch.inference_mode()
def evaluate(net, dataloader, device, amp, epoch_pred_dir=None, postprocess=False, postprocess_fn=None,
             target_class=2):
    """``evaluate.evaluate`` (evaluate.py:12-173) with the per-batch tail on the device: same arguments, same
    ``(dice_original, dice_postprocessed, min_dice)`` return, ONE host read at the end instead of one per batch.

    The OpenCV post-processing (utils/post_process.py) is outside the hot path: pass the
    reference's ``postprocess_mask`` as ``postprocess_fn`` to keep that leg (it then runs on the host on the uint8
    label map this function already produced, evaluate.py:124-139); with ``postprocess=False`` the post-processed
    score equals the original one (evaluate.py:169-170).  NOTE the default here is ``postprocess=False`` (the
    reference's is True): the post-processing itself is not part of this package.  ``epoch_pred_dir`` saves the
    prediction PNGs like evaluate.py:92-107,146-166 (host IO on the uint8 label maps: one device->host copy per batch)."""
    if postprocess and postprocess_fn is None:
        raise ValueError("unetb200.evaluate: postprocess=True needs postprocess_fn (utils.post_process.postprocess_mask)")
    net.eval()
    num_val_batches = len(dataloader)
    scores, scores_post = [], []
    post_dir = None
    if epoch_pred_dir is not None:
        import os
        os.makedirs(epoch_pred_dir, exist_ok=True)
        if postprocess:                                   # evaluate.py:38-40
            post_dir = os.path.join(epoch_pred_dir, "postprocessed")
            os.makedirs(post_dir, exist_ok=True)
    with torch.autocast(device.type, enabled=amp):
        for batch_index, batch in enumerate(dataloader):
            image, mask_true = batch['image'], batch['mask']
            image = image.to(device=device, dtype=torch.float32, memory_format=torch.channels_last)
            mask_true = mask_true.to(device=device)
            mask_pred = net(image)
            if net.n_classes == 1:
                pred, dice, _ = binary_dice(mask_pred, mask_true)
                scale = 255
            else:
                pred, dice, _ = argmax_class_dice(mask_pred, mask_true, c=target_class, index_dtype=torch.uint8)
                scale = 1
            scores.append(dice)
            host = pred.cpu().numpy() if (postprocess or epoch_pred_dir is not None) else None
            if epoch_pred_dir is not None:
                _save_pngs(host, epoch_pred_dir, batch_index, net.n_classes)
            if postprocess:
                import numpy as np
                post = np.stack([postprocess_fn(m * scale) // scale for m in host]).astype(np.float32)
                if post_dir is not None:
                    _save_pngs(post.astype(np.uint8), post_dir, batch_index, net.n_classes)
                post = torch.from_numpy(post).to(device)
                if net.n_classes == 1:
                    true = torch.div(mask_true.float(), 2, rounding_mode="floor")
                else:
                    true = (mask_true == target_class).float()
                    post = (post == target_class).float()
                from .losses import dice_coeff
                scores_post.append(dice_coeff(post, true, reduce_batch_first=False))
    net.train()
    if not scores:
        return 0, 0, 10
    s = torch.stack(scores)
    sp = torch.stack(scores_post) if postprocess else s
    # evaluate.py:85 takes min(original, post-processed) in the binary branch, :121 the original only
    per_batch_min = (torch.minimum(s, sp) if net.n_classes == 1 else s).min()
    total, total_post, mn = torch.stack([s.sum(), sp.sum(), per_batch_min]).tolist()      # the only host read
    n = max(num_val_batches, 1)
    return total / n, total_post / n, min(mn, 10)


def _save_pngs(labels, directory, batch_index, n_classes):
    """evaluate.py:92-107 (binary: 0 / 255) and :146-166 (classes 0 / 1 / 2 -> 0 / 128 / 255): pred_batch{b}_sample{i}.png"""
    import os

    import numpy as np
    from PIL import Image
    lut = np.zeros(256, dtype=np.uint8)
    if n_classes == 1:
        lut[1] = 255
    else:
        lut[1], lut[2] = 128, 255
    for i, m in enumerate(labels):
        Image.fromarray(lut[np.asarray(m, dtype=np.uint8)]).save(os.path.join(directory, f"pred_batch{batch_index}_sample{i}.png"))


def predict_img(model, img, device, out_size=None, index_dtype=torch.int64):
    """``predict.predict_img`` (predict.py:15-29) for an already pre-processed image tensor ``img`` [C,H,W] or
    [B,C,H,W] (BasicDataset.preprocess / :mod:`unetb200.data` output): eval-mode forward under autocast, then the
    fused resize + argmax.  ``out_size`` = (H, W) of the original image (defaults to the input size, where the
    resize is the identity).  Returns the label map on the device; ``.cpu().numpy()`` is the caller's last step."""
    model.eval()
    if img.dim() == 3:
        img = img.unsqueeze(0)
    img = img.to(device=device, dtype=torch.float32, memory_format=torch.channels_last)
    with torch.no_grad(), torch.autocast(device.type, enabled=True):
        mask_pred = model(img)
        size = tuple(out_size) if out_size is not None else tuple(img.shape[-2:])
        idx = resize_argmax(mask_pred, size, index_dtype=index_dtype)
    return idx.squeeze(0) if idx.shape[0] == 1 else idx
