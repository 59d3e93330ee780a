"""-m gpu: parity of the CUDA path (called through the C ABI) against the CPU oracle / PyTorch-CPU fp32
references on the same seeded inputs.  See tests/gpu_checks.py for the checks and tolerances."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _assert_all(results):
    bad = [(label, err, tol) for label, err, tol in results if not (err <= tol)]
    assert not bad, "parity failures: " + "; ".join(f"{l}: err={e:.3e} > tol={t:.1e}" for l, e, t in bad)


@pytest.fixture(scope="module")
def G():
    import gpu_checks
    return gpu_checks


def test_library_is_the_one_in_tree(G):
    from unetb200 import _lib
    lib = _lib.load()
    import ctypes
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.unetb200_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)) == 0
    assert major.value == 10, f"built for sm_100a, running on sm_{major.value}{minor.value}"
    with torch.autocast("cuda", enabled=True):
        assert G.UF.compute_dtype(torch.zeros(1)) == torch.bfloat16


@pytest.mark.parametrize("group", ["layout", "bn_fwd", "maxpool", "bn_bwd", "upsample"])
def test_memory_bound_kernels(G, golden, group):
    _assert_all(G.all_groups()[group](golden))


@pytest.mark.parametrize("group", ["ce_dice", "dice", "boundary"])
def test_losses(G, golden, group):
    _assert_all(G.all_groups()[group](golden))


def test_fused_optimizer(G, golden):
    _assert_all(G.all_groups()["optim"](golden))


def test_outconv(G, golden):
    _assert_all(G.all_groups()["outconv"](golden))


def test_conv_simt(G, golden):
    _assert_all(G.all_groups()["conv_simt"](golden))


@pytest.mark.parametrize("group", ["conv_tc_first", "conv_tc", "conv_tc_tf32", "conv_tc_x3", "conv_layouts", "conv_bnfold", "dgrad_bnbwd", "conv_narrow", "narrow_bounds"])
def test_conv_tcgen05(G, golden, group):
    _assert_all(G.all_groups()[group](golden))


@pytest.mark.parametrize("group", ["parts_fp32", "parts_bf16"])
def test_parts_against_reference_fixtures(G, golden, group):
    _assert_all(G.all_groups()[group](golden))


@pytest.mark.parametrize("group", ["unet_fp32", "unet_fp32_b", "unet_tf32", "unet_bf16", "unet_bf16_bil", "unet_ragged", "unet_infer", "graph_side_stream", "unet_widths", "unet_sa", "checkpointing",
                                   "north_star_tf32x3", "segments", "prepack", "north_star_bf16", "north_star_bf16_b", "full_c2", "full_c3", "full_c5"])
def test_unet_training_step(G, golden, group):
    _assert_all(G.all_groups()[group](golden))
