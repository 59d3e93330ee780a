"""Worker of tests/test_callers_ref.py: run the reference's UNMODIFIED train.py / evaluate.py / predict.py
(the git-ignored baseline/_ref copy) end to end on a tiny on-disk dataset, either

  --mode dropin : with this repo's drop-in `unet` / `utils.dice_score` / `utils.boundary_loss` first on sys.path
                  (everything else -- utils.data_loading, utils.post_process, evaluate, predict, train -- is the
                  reference's own file), or
  --mode stock  : with the reference's own modules only (stock PyTorch: cuDNN / ATen, fp16 autocast + GradScaler).

Exercises exactly what SURVEY.md section 8(b) lists as the boundary: train.train_model's loop (train.py:105-163: fp16-default
autocast, GradScaler, clip_grad_norm_, RMSprop, channels_last inputs, DataLoader workers), evaluate.evaluate
(inference_mode, post-processing, PNG saving) and predict.predict_img.  Writes one JSON file with the observable
results.  train.py cannot be imported as shipped (it imports unet.unet_nested_model and yolo, which the reference
does not contain, train.py:16,18; utils/utils.py needs matplotlib): the drop-in package ships those stubs, the stock
mode gets them injected here."""
import argparse
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "baseline", "_ref")
PKG = os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200")


def make_dataset(root, n_train=3, n_val=1, size=64):
    """grey PNG images + contour masks coded 0 / 128 / 255 (utils/data_loading.py:70-77), <id>.png / <id>_mask.png"""
    import numpy as np
    from PIL import Image
    rng = np.random.RandomState(0)
    yy, xx = np.mgrid[0:size, 0:size]
    for split, n in (("train", n_train), ("val", n_val)):
        di, dm = os.path.join(root, "imgs", split), os.path.join(root, "masks", split)
        os.makedirs(di, exist_ok=True)
        os.makedirs(dm, exist_ok=True)
        for i in range(n):
            mask = np.full((size, size), 128, np.uint8)
            cy, cx, r = rng.uniform(0.3, 0.7) * size, rng.uniform(0.3, 0.7) * size, rng.uniform(0.18, 0.3) * size
            d2 = (yy - cy) ** 2 + (xx - cx) ** 2
            mask[d2 < r * r] = 255
            mask[d2 < (0.4 * r) ** 2] = 0
            img = (0.25 * 255 + 0.5 * mask + rng.normal(0, 8, mask.shape)).clip(0, 255).astype(np.uint8)
            Image.fromarray(img, "L").save(os.path.join(di, f"{split}{i}.png"))
            Image.fromarray(mask, "L").save(os.path.join(dm, f"{split}{i}_mask.png"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", required=True, choices=["dropin", "stock"])
    ap.add_argument("--work", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--device", default="cuda")
    a = ap.parse_args()
    sys.path[:0] = ([PKG] if a.mode == "dropin" else []) + [REF]
    plt = types.ModuleType("matplotlib.pyplot")
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    if a.mode == "stock":                                   # the modules train.py imports but the reference lacks
        nested = types.ModuleType("unet.unet_nested_model")
        nested.UNetPlusPlus_S = nested.UNetPlusPlus = object
        yolo, yseg = types.ModuleType("yolo"), types.ModuleType("yolo.yolov8_seg_model")
        yseg.YOLOv8_Seg_S = object
        yolo.yolov8_seg_model = yseg
        sys.modules.update({"unet.unet_nested_model": nested, "yolo": yolo, "yolo.yolov8_seg_model": yseg})
    import numpy as np
    import torch
    from PIL import Image
    os.makedirs(a.work, exist_ok=True)
    os.chdir(a.work)
    make_dataset(os.path.join(a.work, "data"))

    import evaluate as ref_evaluate
    import predict as ref_predict
    import train as ref_train
    import unet
    assert os.path.samefile(ref_train.__file__, os.path.join(REF, "train.py"))
    assert os.path.samefile(ref_evaluate.__file__, os.path.join(REF, "evaluate.py"))
    assert os.path.samefile(ref_predict.__file__, os.path.join(REF, "predict.py"))
    where = os.path.dirname(os.path.abspath(unet.__file__))
    assert os.path.samefile(where, os.path.join(PKG if a.mode == "dropin" else REF, "unet")), where
    import utils.dice_score as ds
    assert (os.path.dirname(os.path.abspath(ds.__file__)) == os.path.join(PKG, "utils")) == (a.mode == "dropin")

    from pathlib import Path
    data = Path(a.work) / "data"
    ref_train.dir_img_train, ref_train.dir_img_val = data / "imgs/train", data / "imgs/val"
    ref_train.dir_mask_train, ref_train.dir_mask_val = data / "masks/train", data / "masks/val"
    losses = []

    class Bar:                       # train.py reports its running loss only through tqdm: capture it
        def __init__(self, *args, **kw):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def update(self, n):
            pass

        def set_postfix(self, **kw):
            losses.append(float(kw.get("loss (total)", float("nan"))))

    ref_train.tqdm = Bar
    device = torch.device(a.device)
    torch.manual_seed(0)
    model = unet.UNet(n_channels=1, n_classes=3, bilinear=False)
    model = model.to(memory_format=torch.channels_last).to(device)       # train.py:262,282
    w0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    torch.manual_seed(1)                                                  # DataLoader shuffle order
    ref_train.train_model(model=model, device=device, epochs=1, batch_size=2, learning_rate=1e-5, img_scale=1.0,
                          amp=True, save_checkpoint=False)
    sd = model.state_dict()
    moved = max(float((sd[k].float() - w0[k].float()).abs().max()) for k in sd if k.endswith("weight"))
    pngs = sorted(os.listdir(os.path.join(a.work, "predictions", "epoch_1")))

    from torch.utils.data import DataLoader
    from utils.data_loading import BasicDataset
    val = BasicDataset(str(data / "imgs/val"), str(data / "masks/val"), 1.0, augment=False)
    loader = DataLoader(val, batch_size=1, shuffle=False)
    scores = [float(s) for s in ref_evaluate.evaluate(model, loader, device, True, None, postprocess=True)]
    assert model.training                                                 # evaluate.py:166 puts it back
    img = Image.open(data / "imgs/val/val0.png")
    mask = ref_predict.predict_img(model, img, device)
    assert isinstance(mask, np.ndarray) and mask.shape == (64, 64) and mask.dtype == np.int64
    np.save(a.out + ".mask.npy", mask)
    out = {"mode": a.mode, "losses": losses, "val_scores": scores, "weights_moved": moved, "pngs": len(pngs),
           "saved_model": os.path.exists(os.path.join(a.work, "model_epoch1.pth")),
           "num_batches_tracked": int(sd["inc.double_conv.1.num_batches_tracked"]),
           "running_mean_inc": sd["inc.double_conv.1.running_mean"].float().cpu().tolist(),
           "running_var_up4": sd["up4.conv.double_conv.4.running_var"].float().cpu().tolist(),
           "state_keys": len(sd), "unet_from": where}
    with open(a.out, "w") as f:
        json.dump(out, f)
    print("callers_ref_worker", a.mode, "ok: losses", [round(v, 4) for v in losses], "val", scores)


if __name__ == "__main__":
    main()
