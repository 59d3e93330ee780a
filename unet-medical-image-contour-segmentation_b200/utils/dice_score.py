"""Drop-in for the reference's ``utils/dice_score.py`` (same three call signatures, dice_score.py:5,28,33);
the reductions run in libunetb200.so (one fused pass, fp64 cross-block accumulation)."""
from unetb200.losses import dice_coeff, dice_loss, multiclass_dice_coeff  # noqa: F401
