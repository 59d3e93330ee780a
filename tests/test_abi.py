"""CPU: the C-ABI library loads, exports every symbol include/unetb200.h declares, its struct layout
matches the ctypes mirror, and the drop-in modules expose the reference's surface.  No compute calls."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "unetb200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(unetb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from unetb200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libunetb200.so does not export {n}"
        assert n in _lib.PROTOTYPES, f"ctypes binding lacks a prototype for {n}"
    assert set(_lib.PROTOTYPES) == set(names)
    assert lib.unetb200_version() >= 100


def test_gconv_struct_layout(tmp_path):
    from unetb200._lib import GConv
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "unetb200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                 'sizeof(unetb200_gconv_t), offsetof(unetb200_gconv_t, tap_dx), offsetof(unetb200_gconv_t, ld_in),'
                 'offsetof(unetb200_gconv_t, N), offsetof(unetb200_gconv_t, ld_out));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert got == [ctypes.sizeof(GConv), GConv.tap_dx.offset, GConv.ld_in.offset, GConv.N.offset, GConv.ld_out.offset]


def test_sass_is_blackwell_native():
    """tcgen05.mma / tcgen05.ld / TMA must be present in the shipped SASS (B200_PROFILING.md table)."""
    from unetb200 import _lib
    try:
        sass = subprocess.check_output(["cuobjdump", "-sass", _lib.LIB_PATH], text=True, stderr=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in subprocess.check_output(["cuobjdump", "-lelf", _lib.LIB_PATH], text=True)


def test_state_dict_surface():
    import unet
    from oracle import unet_oracle as O
    for args, nkeys in (((1, 2, False), 118), ((1, 2, True), 110), ((3, 4, False), 118)):
        torch.manual_seed(0)
        m = unet.UNet(*args)
        st = O.build_state(*args, seed=0)
        sd = m.state_dict()
        assert len(sd) == nkeys and list(sd) == list(st)
        for k in sd:                      # identical draws to the reference constructor (via the oracle)
            assert torch.equal(sd[k], st[k]), k
        assert (m.n_channels, m.n_classes, m.bilinear) == args
    assert sum(p.numel() for p in unet.UNet(1, 2).parameters()) == 31036546
    assert sum(p.numel() for p in unet.UNet(1, 2, True).parameters()) == 17261890


def test_reference_import_surface():
    """Names train.py / predict.py / evaluate.py import (train.py:13-20)."""
    from unet import UNet, UNet_S  # noqa: F401
    from unet.unet_model import UNet_SA, UNet_T  # noqa: F401
    from unet.unet_nested_model import UNetPlusPlus, UNetPlusPlus_S  # noqa: F401
    from unet.unet_parts import DoubleConv, Down, OutConv, Up  # noqa: F401
    from utils.boundary_loss import boundary_loss  # noqa: F401
    from utils.dice_score import dice_coeff, dice_loss, multiclass_dice_coeff  # noqa: F401
    from yolo.yolov8_seg_model import YOLOv8_Seg_S  # noqa: F401
    m = UNet(1, 3)
    assert hasattr(m, "use_checkpointing")
    for name in ("inc", "down1", "down2", "down3", "down4", "up1", "up2", "up3", "up4", "outc"):
        assert hasattr(m, name)


def test_no_cpu_fallback():
    import unet
    from utils.boundary_loss import boundary_loss
    from utils.dice_score import dice_loss
    m = unet.UNet(1, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dice_loss(torch.zeros(2, 4, 4), torch.zeros(2, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        boundary_loss(torch.zeros(2, 4, 4), torch.zeros(2, 4, 4))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.replace("oracle/", "").lower() or f.endswith(".md"), os.path.join(dirpath, f)
