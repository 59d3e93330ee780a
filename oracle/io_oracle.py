"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the callers either side of the UNet step
(SURVEY.md section 8(f) N1 and N3): the evaluate / predict tails and the uint8 input pipeline.

  * ``utils/data_loading.py:65-89``  BasicDataset.preprocess (scale == 1)  -> :func:`preprocess_image`, :func:`preprocess_mask`
  * ``utils/data_loading.py:91-98``  rotate_image_and_mask                 -> :func:`rotate`
  * ``utils/data_loading.py:100-132`` __getitem__ (rotation index, dtypes)  -> :func:`make_batch`
  * ``evaluate.py:111-117``          multi-class tail                      -> :func:`eval_multiclass`
  * ``evaluate.py:56-66``            binary tail                           -> :func:`eval_binary`
  * ``predict.py:26-27``             resize + argmax                       -> :func:`predict_tail`,
                                     :func:`resize_argmax_exact` (ATen's arithmetic restated op by op)

Byte / index work: numpy integer arithmetic, exact.  The dice of the tails is the reference's own
``dice_coeff`` formula (dice_score.py:5-25) on 0/1 tensors.

PARITY PINNING: the reference has no tests or fixtures; ``tests/golden/make_golden_io.py`` runs the
reference's own ``BasicDataset.preprocess`` / ``rotate_image_and_mask`` (PIL) / ``dice_coeff`` and the
literal evaluate.py / predict.py tail expressions in the build container and commits
``tests/golden/golden_io_v1.pt``; ``tests/test_oracle_io.py`` re-checks this module against it.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# input pipeline
# --------------------------------------------------------------------------------------
def rotate(a: np.ndarray, k: int) -> np.ndarray:
    """PIL ``Image.rotate(90*k, expand=True)`` (data_loading.py:95-97): for multiples of 90 degrees PIL
    dispatches to ``transpose(ROTATE_90/180/270)``, a counter-clockwise quarter-turn permutation of the
    pixels == ``numpy.rot90(a, k)`` on the [H, W(, C)] array."""
    return np.ascontiguousarray(np.rot90(a, k % 4, axes=(0, 1)))


def preprocess_image(img_u8: np.ndarray) -> np.ndarray:
    """data_loading.py:81-89 for scale == 1 (the resize at :69 is the identity): [H,W] -> [1,H,W],
    [H,W,C] -> [C,H,W]; ``/ 255.0`` in float32 only when some value exceeds 1."""
    img = img_u8
    img = img[np.newaxis, ...] if img.ndim == 2 else img.transpose((2, 0, 1))
    if (img > 1).any():
        img = img.astype(np.float32) / 255.0
    return img


def preprocess_mask(mask_u8: np.ndarray) -> np.ndarray:
    """data_loading.py:73-79: zeros, then 255 -> 2, 128 -> 1, 0 -> 0 (any other gray level stays 0)."""
    mask = np.zeros(mask_u8.shape, dtype=np.int8)
    mask[mask_u8 == 255] = 2
    mask[mask_u8 == 128] = 1
    mask[mask_u8 == 0] = 0
    return mask


def make_batch(images_u8, masks_u8, rots):
    """data_loading.py:100-132 for a list of samples: rotate (index k = rotation_idx), preprocess, and
    the dtypes of the returned dict (:130-131); stacked like the default collate."""
    imgs, msks = [], []
    for im, mk, k in zip(images_u8, masks_u8, rots):
        im, mk = rotate(im, k), rotate(mk, k)
        imgs.append(torch.as_tensor(preprocess_image(im).copy()).float().contiguous())
        msks.append(torch.as_tensor(preprocess_mask(mk).copy()).long().contiguous())
    return torch.stack(imgs), torch.stack(msks)


# --------------------------------------------------------------------------------------
# evaluate / predict tails
# --------------------------------------------------------------------------------------
def dice_coeff(input, target, reduce_batch_first=False, epsilon=1e-6):
    """dice_score.py:5-25."""
    assert input.size() == target.size()
    assert input.dim() == 3 or not reduce_batch_first
    sum_dim = (-1, -2) if input.dim() == 2 or not reduce_batch_first else (-1, -2, -3)
    inter = 2 * (input * target).sum(dim=sum_dim)
    sets_sum = input.sum(dim=sum_dim) + target.sum(dim=sum_dim)
    sets_sum = torch.where(sets_sum == 0, inter, sets_sum)
    return ((inter + epsilon) / (sets_sum + epsilon)).mean()


def eval_multiclass(mask_pred: torch.Tensor, mask_true: torch.Tensor, c: int = 2):
    """evaluate.py:111-117 -> (argmax indices int64 [B,H,W], dice of class c, counts int64 [B,3])."""
    idx = mask_pred.argmax(dim=1)
    pred_c = (idx == c).float()
    true_c = (mask_true == c).float()
    counts = torch.stack([(pred_c * true_c).sum((-1, -2)), pred_c.sum((-1, -2)), true_c.sum((-1, -2))], 1).long()
    return idx, dice_coeff(pred_c, true_c, reduce_batch_first=False), counts


def eval_binary(mask_pred: torch.Tensor, mask_true: torch.Tensor):
    """evaluate.py:56-66 (n_classes == 1): mask_true //= 2, sigmoid, threshold, dice."""
    mask_true = mask_true.clone().float()
    mask_true //= 2
    assert mask_true.min() >= 0 and mask_true.max() <= 1, 'True mask indices should be in [0, 1]'
    prob = torch.sigmoid(mask_pred.squeeze(1))
    binary = (prob > 0.5).float()
    counts = torch.stack([(binary * mask_true).sum((-1, -2)), binary.sum((-1, -2)), mask_true.sum((-1, -2))], 1).long()
    return binary, dice_coeff(binary, mask_true, reduce_batch_first=False), counts


def predict_tail(mask_pred: torch.Tensor, size) -> torch.Tensor:
    """predict.py:26-27: F.interpolate(..., mode='bilinear') then argmax over classes -> int64 [B,H,W]."""
    return F.interpolate(mask_pred, size, mode='bilinear').argmax(dim=1)


def resize_argmax_exact(mask_pred: torch.Tensor, size) -> torch.Tensor:
    """The same, restating ATen's upsample_bilinear2d arithmetic (UpSample.h area_pixel_compute_source_index,
    align_corners=False; value = h0*(w0*x00 + w1*x01) + h1*(w0*x10 + w1*x11) in fp32, rounded to the input dtype)
    with numpy float32 operations in a fixed order -- what the CUDA kernel computes bit for bit."""
    B, C, h, w = mask_pred.shape
    H, W = size
    x = mask_pred.float().numpy()
    f32 = np.float32

    def axis(n_in, n_out):
        scale = f32(n_in) / f32(n_out)
        src = scale * (np.arange(n_out, dtype=f32) + f32(0.5)) - f32(0.5)
        src = np.maximum(src, f32(0)).astype(f32)
        i0 = np.minimum(src.astype(np.int64), n_in - 1)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (src - i0.astype(f32)).astype(f32)
        return i0, i1, (f32(1) - l1).astype(f32), l1

    y0, y1, hy0, hy1 = axis(h, H)
    x0, x1, wx0, wx1 = axis(w, W)
    v00 = x[:, :, y0][:, :, :, x0]
    v01 = x[:, :, y0][:, :, :, x1]
    v10 = x[:, :, y1][:, :, :, x0]
    v11 = x[:, :, y1][:, :, :, x1]
    top = (wx0 * v00).astype(f32) + (wx1 * v01).astype(f32)
    bot = (wx0 * v10).astype(f32) + (wx1 * v11).astype(f32)
    val = (hy0[:, None] * top).astype(f32) + (hy1[:, None] * bot).astype(f32)
    val = torch.from_numpy(val.astype(f32)).to(mask_pred.dtype).float()
    return val.argmax(dim=1)
