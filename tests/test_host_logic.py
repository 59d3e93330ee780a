"""CPU: host-side logic -- NHWC stride detection, channel-slice aliasing, weight packing index maps,
precision policy."""
import torch

from unetb200 import functional as UF
from unetb200 import ops


def test_nhwc_ld():
    t = torch.empty(2, 5, 6, 8).permute(0, 3, 1, 2)          # logical [2,8,5,6], NHWC
    assert ops.nhwc_ld(t) == 8
    assert ops.nhwc_ld(t[:, :3]) == 8                         # channel slice keeps the pixel stride
    assert ops.nhwc_ld(torch.empty(2, 8, 5, 6)) is None       # NCHW contiguous
    assert ops.nhwc_ld(torch.empty(2, 1, 5, 6)) == 1          # C == 1: both layouts coincide
    assert ops.nhwc_ld(torch.empty(2, 8, 5, 6).contiguous(memory_format=torch.channels_last)) == 8
    s = ops.channel_slice(t, 2, 4)
    assert s.shape == (2, 4, 5, 6) and s.data_ptr() == t.data_ptr() + 2 * 4 and ops.nhwc_ld(s) == 8
    assert s._base is None                                    # alias without an autograd view relation
    t.zero_()
    s.fill_(1.0)
    assert t[:, 2:6].eq(1).all() and t[:, :2].eq(0).all() and t[:, 6:].eq(0).all()


def _emulate_pack(w, n0, n1, n2, s0, s1, s2, off):
    # dst[i0][i1][i2] = storage[off + i0*s0 + i1*s1 + i2*s2]   (what unetb200_pack_weights does)
    flat = torch.as_strided(w, (w.untyped_storage().nbytes() // w.element_size(),), (1,), 0)
    i0 = torch.arange(n0).view(n0, 1, 1)
    i1 = torch.arange(n1).view(1, n1, 1)
    i2 = torch.arange(n2).view(1, 1, n2)
    return flat[w.storage_offset() + off + i0 * s0 + i1 * s1 + i2 * s2]


def _capture_pack(kind, w):
    """the job functional._JOBS[kind] describes for `w`, executed by the emulation above"""
    src, (n0, n1, n2), (s0, s1, s2), off, shape = UF._JOBS[kind](w)
    assert shape[0] * shape[1] == n0 * n1 * n2
    return _emulate_pack(src, n0, n1, n2, s0, s1, s2, off)


def test_pack_index_maps():
    g = torch.Generator().manual_seed(0)
    for fmt in (torch.contiguous_format, torch.channels_last):
        w = torch.randn(6, 4, 3, 3, generator=g).contiguous(memory_format=fmt)       # OIHW
        p = _capture_pack("f3", w)                                                    # [co][t][ci]
        assert torch.equal(p, w.permute(0, 2, 3, 1).reshape(6, 9, 4))
        p = _capture_pack("d3", w)                                                    # [ci][t'][co], flipped taps
        assert torch.equal(p, w.flip(2, 3).permute(1, 2, 3, 0).reshape(4, 9, 6))
        wt = torch.randn(8, 5, 2, 2, generator=g).contiguous(memory_format=fmt)      # IOHW
        p = _capture_pack("fT", wt)                                                   # [q][co][ci]
        assert torch.equal(p, wt.permute(2, 3, 1, 0).reshape(4, 5, 8))
        p = _capture_pack("dT", wt)                                                   # [ci][q][co]
        assert torch.equal(p, wt.permute(0, 2, 3, 1).reshape(8, 4, 5))


def test_grad_sinks_are_offered_only_when_nothing_accumulates():
    w = torch.nn.Parameter(torch.randn(4, 3, 3, 3).contiguous(memory_format=torch.channels_last))
    buf = torch.zeros(w.numel())
    view = buf.as_strided(w.shape, w.stride())
    UF.set_grad_sinks([w], [view])
    d = UF.grad_dst(w)
    assert d is not None and d is not view and d.data_ptr() == view.data_ptr() and d.stride() == w.stride()
    w.grad = torch.zeros_like(w)
    assert UF.grad_dst(w) is None                      # an existing .grad would be accumulated into: no sink
    UF.set_grad_sinks([w], [view], always=True)
    assert UF.grad_dst(w) is not None                  # the owner of p.grad drives backward through autograd.grad
    UF.clear_grad_sinks([w])
    assert UF.grad_dst(w) is None
    try:
        UF.set_grad_sinks([w], [torch.zeros(4, 3, 3, 3)])   # contiguous view for a channels_last parameter
        raise AssertionError("layout mismatch accepted")
    except ValueError:
        pass


def test_dgrad_packing_is_the_conv_transpose():
    """conv3x3(dy, flipped/transposed W) == autograd dgrad of conv3x3(x, W): validates the tap algebra
    the CUDA dgrad relies on."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 7, 6, generator=g, requires_grad=True)
    w = torch.randn(5, 4, 3, 3, generator=g)
    dy = torch.randn(2, 5, 7, 6, generator=g)
    F.conv2d(x, w, padding=1).backward(dy)
    wd = w.flip(2, 3).permute(1, 0, 2, 3)                     # [ci][co][kh'][kw']
    assert torch.allclose(F.conv2d(dy, wd, padding=1), x.grad, atol=1e-4)


def test_precision_policy(monkeypatch):
    from unetb200._lib import ALGO_AUTO, ALGO_PREFER_TC, ALGO_SIMT
    assert UF.conv_algo(torch.bfloat16) == ALGO_AUTO
    monkeypatch.setenv("UNET_B200_PRECISION", "fp32")
    assert UF.conv_algo(torch.float32) == ALGO_SIMT
    monkeypatch.setenv("UNET_B200_PRECISION", "tf32")
    assert UF.conv_algo(torch.float32) == ALGO_PREFER_TC
    x = torch.zeros(1)
    assert UF.compute_dtype(x) == torch.float32
    assert UF.compute_dtype(x.bfloat16()) == torch.bfloat16      # (autocast -> bf16 is checked on the GPU)


def test_wgrad_split_counts_fill_whole_waves():
    """The tcgen05 wgrad kernels hold one CTA (or CTA pair) per SM, so a split-K grid runs in waves: the plan
    must not launch e.g. 297 CTAs on 148 SMs (2.007 waves cost 3 -- this cost 1.2 ms/step before pick_splits).
    Host-only: unetb200_gconv_wgrad_plan needs no GPU (148 SMs assumed without one)."""
    import ctypes as C
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "unet-medical-image-contour-segmentation_b200"))
    from unetb200 import _lib
    lib = _lib.load()
    B = 16
    # (Cin, Cout, H) of the 3x3 convolutions of UNet(1,2) at 512x512 that run on tcgen05
    layers = [(64, 64, 512), (128, 64, 512), (64, 128, 256), (128, 128, 256), (256, 128, 256), (128, 256, 128),
              (256, 256, 128), (512, 256, 128), (256, 512, 64), (512, 512, 64), (1024, 512, 64), (512, 1024, 32),
              (1024, 1024, 32)]
    for cin, cout, hw in layers:
        d = _lib.GConv()
        d.dtype, d.algo = _lib.BF16, _lib.ALGO_TC
        d.B, d.Hm, d.Wm, d.Cin, d.ntaps = B, hw, hw, cin, 9
        for t in range(9):
            d.tap_dy[t], d.tap_dx[t] = t // 3 - 1, t % 3 - 1
        d.in_scale, d.in_off_y, d.in_off_x, d.Hin, d.Win, d.ld_in = 1, 0, 0, hw, hw, cin
        d.N, d.nquad, d.out_scale, d.out_off_y, d.out_off_x, d.Hout, d.Wout, d.ld_out = cout, 1, 1, 0, 0, hw, hw, cout
        splits, used = C.c_int(0), C.c_int(0)
        assert lib.unetb200_gconv_wgrad_plan(C.byref(d), C.byref(splits), C.byref(used)) == 0
        assert used.value == _lib.ALGO_TC
        chunks = cin // 64
        if cout % 128 == 0 and (3 * chunks) % 4 == 0:          # CTA-pair kernel: slots = SM pairs
            tiles, slots = (3 * chunks // 4) * (cout // 128), 74
        else:                                                   # N-stacked kernel: one CTA per (x chunk, dY group)
            tiles, slots = chunks * (cout // 64), 148
        ctas = tiles * splits.value
        waves = -(-ctas // slots)
        assert ctas / (waves * slots) >= 0.95, (cin, cout, hw, splits.value, ctas, waves)


def test_side_stream_eligibility_rules():
    """A weight gradient may be produced on the side stream only when autograd will keep the tensor untouched
    (functional._grad_kept_as_is): pure host logic, checked on CPU tensors."""
    import os
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "unet-medical-image-contour-segmentation_b200"))
    from unetb200 import functional as UF
    w = torch.nn.Parameter(torch.randn(8, 4, 3, 3).contiguous(memory_format=torch.channels_last))
    with torch.no_grad():
        g_same = torch.empty_like(w)
        g_other = torch.empty(8, 4, 3, 3)
        assert UF._grad_kept_as_is(w, g_same)
        assert not UF._grad_kept_as_is(w, g_other)                      # layout contract: autograd would re-layout it
        w.grad = torch.zeros_like(w)
        assert not UF._grad_kept_as_is(w, g_same)                       # accumulation into an existing .grad
        w.grad = None
        h = w.register_hook(lambda g: g)
        assert not UF._grad_kept_as_is(w, g_same)                       # a tensor hook reads the gradient during backward
        h.remove()
        h2 = w.register_post_accumulate_grad_hook(lambda p: None)
        assert not UF._grad_kept_as_is(w, g_same)
        h2.remove()
        assert UF._grad_kept_as_is(w, g_same)
        assert not UF._grad_kept_as_is(w, g_same.double())
    assert not UF._grad_kept_as_is(w, g_same)                           # grad mode on: double backward would clone
