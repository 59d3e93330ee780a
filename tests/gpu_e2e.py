"""End-to-end UNet training-step parity on the GPU.

The CPU fp32 oracle is the ground truth.  For a ReLU network at random initialisation (BatchNorm beta = 0
puts every ReLU threshold exactly at the batch mean) any change of rounding -- TF32, bf16, or merely a
different fp32 summation order -- flips the ReLU mask of the activations closest to zero, and every flipped
element changes the gradients by O(1) at that position.  On random labels the weight gradients are
incoherent sums, so those flips do not average out: measured on B200, PyTorch's *own* GPU path (cuDNN)
sits 0.6 % (fp32), 10 % (tf32) and 37 % (bf16) away from the CPU fp32 gradients in relative L2.  The
north_star's absolute gradient tolerances can therefore not be met by any GPU implementation on this
workload, the reference's included.  The gate used here is the meaningful one: our path must be as close
to the oracle as the reference's own GPU path (torch ATen/cuDNN, run in the test only) is:

    err(ours, oracle) <= 1.25 * err(torch_gpu, oracle) + margin

for logits (max-rel), loss, argmax agreement and gradients (relative L2: median and worst over the 64
parameter tensors), plus absolute checks where they are attainable (fp32 forward, loss, running stats).
"""
import os
import statistics

import torch

import gpu_checks as G
from gpu_checks import DEV, O, host, rel


def _errors(logits, loss, grads, r_logits, r_loss, r_grads):
    l2 = [O.rel_l2(grads[k], r_grads[k]) for k in r_grads]
    return {
        "logits_maxrel": rel(logits, r_logits),
        "loss_rel": abs(loss - float(r_loss)) / abs(float(r_loss)),
        "argmax_mismatch": (logits.argmax(1) != r_logits.argmax(1)).float().mean().item(),
        "grad_l2_median": statistics.median(l2),
        "grad_l2_worst": max(l2),
    }


MARGIN = {"logits_maxrel": 2e-4, "loss_rel": 2e-5, "argmax_mismatch": 5e-4, "grad_l2_median": 5e-3,
          "grad_l2_worst": 1e-2}
ABS_FWD = {"fp32": 1e-3, "tf32": 2e-2, "bf16": 1e-1}      # forward logits, max-rel (north_star: 1e-3 / 2e-2)


def gate(nc, ncls, bilinear, B, H, W, mode, boundary_coeff=0.0, fused=True):
    import unet
    tag = f"unet{nc}_{ncls}_{'bil' if bilinear else 'convT'}_{B}x{H}x{W}_{mode}{'' if fused else '_unfused'}"
    st = O.build_state(nc, ncls, bilinear, seed=0)
    img, msk = O.synthetic_batch(B, nc, ncls, H, W)
    ref_st = {k: v.clone() for k, v in st.items()}
    r_logits, r_loss, r_grads = O.training_step(ref_st, img, msk, ncls, bilinear, boundary_coeff=boundary_coeff)
    t_logits, t_loss, t_grads = G.torch_gpu_step(st, img, msk, ncls, bilinear, mode)
    if boundary_coeff:       # the boundary term carries no gradient: only the scalar loss moves
        t_loss = t_loss + boundary_coeff * float(O.boundary_loss(t_logits, msk.float(), edge_width=51, edge_weight=7))
    model = unet.UNet(nc, ncls, bilinear)
    model.load_state_dict(st)
    model = model.to(DEV).to(memory_format=torch.channels_last).train()
    os.environ["UNET_B200_PRECISION"] = "tf32" if mode == "tf32" else "fp32"
    logits, loss, grads = G.unet_step_gpu(model, img, msk, amp=(mode == "bf16"), boundary_coeff=boundary_coeff,
                                          fused=fused)
    ours = _errors(logits, loss, grads, r_logits, r_loss, r_grads)
    theirs = _errors(t_logits, t_loss, t_grads, r_logits, r_loss, r_grads)
    res = []
    for k in ours:
        res.append((f"{tag}_{k} (torch_gpu {theirs[k]:.2e})", ours[k], 1.25 * theirs[k] + MARGIN[k]))
    res.append((f"{tag}_logits_abs", ours["logits_maxrel"], ABS_FWD[mode]))
    res.append((f"{tag}_loss_abs", ours["loss_rel"], 1e-5 if mode == "fp32" else 1e-3))
    sd = model.state_dict()
    rw = max(rel(host(sd[k]), ref_st[k]) for k in sd if "running" in k)
    res.append((f"{tag}_running_stats", rw, {"fp32": 1e-4, "tf32": 5e-3, "bf16": 2e-2}[mode]))
    nbt = all(int(sd[k]) == int(ref_st[k]) for k in sd if "tracked" in k)
    res.append((f"{tag}_num_batches_tracked", 0.0 if nbt else 1.0, 0.0))
    if mode == "fp32":
        # exact mode: only isolated ReLU flips separate us from the oracle
        res.append((f"{tag}_grad_l2_worst_abs", ours["grad_l2_worst"], 5e-2))
    if os.environ.get("UNETB200_TEST_VERBOSE"):
        for k in r_grads:
            print(f"      grad {k:<50s} ours l2 {O.rel_l2(grads[k], r_grads[k]):.3e}  torch_gpu l2 "
                  f"{O.rel_l2(t_grads[k], r_grads[k]):.3e}")
    return res


def infer_gate(nc, ncls, bilinear, B, H, W, mode):
    """predict.py / evaluate.py style forward (BASELINE.json configs[4] shape class): eval-mode BatchNorm (running
    statistics), torch.inference_mode, optional bf16 autocast; logits and argmax mask against the oracle."""
    import unet
    tag = f"infer{nc}_{ncls}_{'bil' if bilinear else 'convT'}_{B}x{H}x{W}_{mode}"
    st = O.build_state(nc, ncls, bilinear, seed=0)
    g = torch.Generator().manual_seed(123)
    for k in st:                                  # non-trivial running statistics
        if k.endswith("running_mean"):
            st[k] = 0.1 * torch.randn(st[k].shape, generator=g)
        elif k.endswith("running_var"):
            st[k] = 0.5 + torch.rand(st[k].shape, generator=g)
    img, _ = O.synthetic_batch(B, nc, ncls, H, W)
    r_logits = O.unet_forward({k: v.clone() for k, v in st.items()}, img, bilinear, training=False)
    model = unet.UNet(nc, ncls, bilinear)
    model.load_state_dict(st)
    model = model.to(DEV).to(memory_format=torch.channels_last).eval()
    os.environ["UNET_B200_PRECISION"] = "tf32" if mode == "tf32" else "fp32"
    x = img.to(DEV).contiguous(memory_format=torch.channels_last)
    with torch.inference_mode(), torch.autocast("cuda", enabled=(mode == "bf16")):
        logits = model(x)
    logits = host(logits.float())
    agree = (logits.argmax(1) == r_logits.argmax(1)).float().mean().item()
    res = [(f"{tag}_logits", rel(logits, r_logits), {"fp32": 1e-3, "tf32": 2e-2, "bf16": 1e-1}[mode]),
           (f"{tag}_argmax_mismatch", 1.0 - agree, {"fp32": 1e-3, "tf32": 5e-3, "bf16": 2e-2}[mode])]
    sd = model.state_dict()
    unchanged = all(torch.equal(host(sd[k]), st[k]) for k in sd if "running" in k or "tracked" in k)
    res.append((f"{tag}_buffers_untouched", 0.0 if unchanged else 1.0, 0.0))
    return res


def graph_gate():
    """The whole-step CUDA graph and the side-stream weight gradients change scheduling, not arithmetic: three
    optimizer steps replayed from a graph must follow the eager, single-stream trajectory."""
    import unet
    from unetb200 import losses as UL
    from unetb200 import ops
    from unetb200.graph import GraphedStep
    st = O.build_state(1, 2, False, seed=0)
    img, msk = O.synthetic_batch(2, 1, 2, 64, 64)
    x = img.to(DEV).contiguous(memory_format=torch.channels_last)
    t = msk.to(DEV)

    def make():
        m = unet.UNet(1, 2, False)
        m.load_state_dict(st)
        m = m.to(DEV).to(memory_format=torch.channels_last).train()
        # plain SGD: RMSprop's first steps are sign-like (g / sqrt((1-alpha) g^2)), which turns rounding-level
        # gradient noise into O(lr) weight differences and makes a trajectory comparison meaningless
        o = torch.optim.SGD(m.parameters(), lr=1e-3, foreach=True)

        def step(xx, tt):
            o.zero_grad(set_to_none=True)
            with torch.autocast("cuda", enabled=True):
                loss = UL.training_criterion(m(xx), tt, boundary_coeff=0.2)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            o.step()
            return loss
        return m, step

    # eager, everything on one stream
    side_was = ops._SIDE_ON
    ops._SIDE_ON = False
    m1, step1 = make()
    l1 = [float(step1(x, t).detach()) for _ in range(5)]
    ops._SIDE_ON = side_was
    # graph (2 eager warm-up steps + 1 captured... the capture itself does not execute), then replays -- with the
    # weight gradients on the side stream, the most intricate schedule the library offers
    wg_was = ops._WGRAD_SIDE
    ops._WGRAD_SIDE = True
    m2, step2 = make()
    g = GraphedStep(step2, (x, t), warmup=2)
    l2 = [float(g.replay().detach()) for _ in range(3)]
    torch.cuda.synchronize()
    ops._WGRAD_SIDE = wg_was
    # bf16 training from random init is chaotic (a ReLU-mask flip changes later steps), so the multi-step comparison
    # only has to catch gross errors (a race gives NaN or errors >> 1); the tight check is the single-step one below
    res = [("graph_loss_step3", abs(l2[0] - l1[2]) / abs(l1[2]), 2e-3),
           ("graph_loss_step5", abs(l2[2] - l1[4]) / abs(l1[4]), 5e-3)]
    w1 = {k: host(v) for k, v in m1.state_dict().items() if v.dtype.is_floating_point}
    w2 = {k: host(v) for k, v in m2.state_dict().items() if v.dtype.is_floating_point}
    # (zero-initialised biases are pure accumulated gradient after 5 steps: a bf16 ReLU flip moves them by tens of
    # per cent in relative terms, so they are compared on the scale of the weights they sit next to)
    worst = max(O.rel_l2(w2[k], w1[k]) for k in w1 if "running" not in k and not k.endswith(".bias"))
    worst_b = max(float((w2[k] - w1[k]).abs().max()) for k in w1 if k.endswith(".bias"))
    res.append(("graph_biases_after_5_steps_abs", worst_b, 1e-3))
    res.append(("graph_weights_after_5_steps_rel_l2", worst, 1e-2))
    # one backward from identical weights: side-stream weight gradients == single-stream weight gradients
    def grads_of(side):
        ops._SIDE_ON = side
        ops._WGRAD_SIDE = side
        m = unet.UNet(1, 2, False)
        m.load_state_dict(st)
        m = m.to(DEV).to(memory_format=torch.channels_last).train()
        with torch.autocast("cuda", enabled=True):
            loss = UL.training_criterion(m(x), t, boundary_coeff=0.2)
        loss.backward()
        torch.cuda.synchronize()
        return {k: host(p.grad) for k, p in m.named_parameters()}
    wgrad_side_was = ops._WGRAD_SIDE
    ga, gb = grads_of(False), grads_of(True)
    ops._SIDE_ON, ops._WGRAD_SIDE = side_was, wgrad_side_was
    res.append(("side_stream_grads_vs_single_stream_rel_l2", max(O.rel_l2(gb[k], ga[k]) for k in ga), 1e-3))
    res.append(("graph_losses_finite", 0.0 if all(v == v and abs(v) < 1e3 for v in l1 + l2) else 1.0, 0.0))
    return res


def width_gate(cls_name, nc, ncls, bilinear, B, H, W, mode):
    """UNet_S / UNet_T (base widths 16 / 8, reference unet_model.py:52-138; UNet_S is what train.py:253 trains by
    default): same modules at narrower widths -- channel counts below 64 run on the CUDA-core engine, the rest on
    tcgen05.  State = the module's own seeded init, shared with the oracle (which is width-agnostic)."""
    import unet.unet_model as UM
    tag = f"{cls_name}_{nc}_{ncls}_{'bil' if bilinear else 'convT'}_{B}x{H}x{W}_{mode}"
    torch.manual_seed(7)
    model = getattr(UM, cls_name)(nc, ncls, bilinear)
    st = {k: v.detach().clone() for k, v in model.state_dict().items()}
    img, msk = O.synthetic_batch(B, nc, ncls, H, W)
    ref_st = {k: v.clone() for k, v in st.items()}
    r_logits, r_loss, r_grads = O.training_step(ref_st, img, msk, ncls, bilinear)
    model = model.to(DEV).to(memory_format=torch.channels_last).train()
    os.environ["UNET_B200_PRECISION"] = mode if mode in ("tf32", "tf32x3") else "fp32"
    logits, loss, grads = G.unet_step_gpu(model, img, msk, amp=(mode == "bf16"))
    os.environ["UNET_B200_PRECISION"] = "fp32"
    e = _errors(logits, loss, grads, r_logits, r_loss, r_grads)
    lim = {"fp32": dict(logits_maxrel=1e-3, loss_rel=1e-5, argmax_mismatch=1e-3, grad_l2_median=2e-2, grad_l2_worst=1e-1),
           "tf32x3": dict(logits_maxrel=1e-3, loss_rel=1e-5, argmax_mismatch=1e-3, grad_l2_median=2e-2, grad_l2_worst=1e-1),
           # no autocast, TF32 allowed: the narrow layers' fprop / dgrad on the TF32 form of the TMA-staged kernel
           "tf32": dict(logits_maxrel=2e-2, loss_rel=1e-3, argmax_mismatch=1e-2, grad_l2_median=2e-1, grad_l2_worst=6e-1),
           "bf16": dict(logits_maxrel=1e-1, loss_rel=5e-3, argmax_mismatch=3e-2, grad_l2_median=6e-1, grad_l2_worst=1.5)}[mode]
    res = [(f"{tag}_{k}", e[k], lim[k]) for k in e]
    sd = model.state_dict()
    res.append((f"{tag}_running_stats", max(rel(host(sd[k]), ref_st[k]) for k in sd if "running" in k),
                1e-4 if mode in ("fp32", "tf32x3") else 2e-2))
    return res


def checkpoint_gate():
    """UNet.use_checkpointing() (reference unet_model.py:40-50, reached from train.py:294-299 after an OOM): the same
    step with every stage under activation re-computation gives the same loss, gradients, BatchNorm running statistics
    and num_batches_tracked, with a lower peak of live memory."""
    import unet
    from unetb200 import losses as UL
    st = O.build_state(1, 2, False, seed=0)
    img, msk = O.synthetic_batch(4, 1, 2, 192, 192)
    x = img.to(DEV).contiguous(memory_format=torch.channels_last)
    t = msk.to(DEV)
    res, out = [], {}
    for mode in ("plain", "recompute"):
        m = unet.UNet(1, 2, False)
        m.load_state_dict(st)
        m = m.to(DEV).to(memory_format=torch.channels_last).train()
        if mode == "recompute":
            m.use_checkpointing()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        with torch.autocast("cuda", enabled=True):
            loss = UL.training_criterion(m(x), t, boundary_coeff=0.2)
        after_fwd = torch.cuda.memory_allocated() - base
        loss.backward()
        torch.cuda.synchronize()
        out[mode] = (float(loss.detach()), {k: host(p.grad) for k, p in m.named_parameters()},
                     {k: host(v) for k, v in m.state_dict().items() if "running" in k or "tracked" in k}, after_fwd)
    (l0, g0, s0, m0), (l1, g1, s1, m1) = out["plain"], out["recompute"]
    res.append(("ckpt_loss_equal", abs(l1 - l0), 0.0))
    res.append(("ckpt_grads_equal_rel_l2", max(O.rel_l2(g1[k], g0[k]) for k in g0), 1e-6))
    res.append(("ckpt_bn_buffers_equal", max(float((s1[k].double() - s0[k].double()).abs().max()) for k in s0), 0.0))
    res.append(("ckpt_live_after_forward_ratio", m1 / max(m0, 1), 0.6))
    return res


def sa_gate_op():
    """SpatialAttention gate (x * sigmoid(conv7x7([mean_c x, max_c x]))), forward and backward, against the fixture the
    UNMODIFIED reference module produced (tests/golden/make_golden_sa.py): fp32 storage to 1e-5, bf16 storage to the
    bf16 tolerance of the other memory-bound ops; also into a channel slice of a wider buffer (the concat buffer)."""
    from unetb200 import functional as UF
    from unetb200 import ops
    g = torch.load(os.path.join(G.ROOT, "tests", "golden", "golden_sa_v1.pt"), weights_only=False)["gate"]
    res = []
    for dt_, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-2)):
        x = G.dev_nhwc(g["x"], dt_).requires_grad_(True)
        w = g["w"].to(DEV).requires_grad_(True)
        B, C, H, W = x.shape
        buf = ops.empty_nhwc(B, 2 * C, H, W, dt_, DEV)
        buf.fill_(3.0)
        out = UF.SpatialGateFn.apply(x, w, UF._Cfg(out=ops.channel_slice(buf, 0, C), save=True))
        out.backward(g["gy"].to(DEV).to(dt_))
        tag = str(dt_)[6:]
        res.append((f"sa_gate_fwd_{tag}", rel(host(out), g["y"]), tol))
        res.append((f"sa_gate_slice_untouched_{tag}", (host(ops.channel_slice(buf, C, C)) - 3.0).abs().max().item(), 0.0))
        res.append((f"sa_gate_gx_{tag}", rel(host(x.grad), g["gx"]), tol if dt_ == torch.float32 else 3e-2))
        res.append((f"sa_gate_gw_{tag}", rel(host(w.grad), g["gw"]), 1e-4 if dt_ == torch.float32 else 3e-2))
    return res


_COND = {}


def _cond_state(nc, ncls, bilinear):
    """oracle.conditioned_state (10 fp32 CPU steps of the reference arithmetic on structured data), cached."""
    key = (nc, ncls, bilinear)
    if key not in _COND:
        _COND[key] = O.conditioned_state(nc, ncls, bilinear)
    return {k: v.clone() for k, v in _COND[key].items()}


def _dice_of(logits, msk, ncls):
    return float(O.multiclass_dice_coeff(torch.softmax(logits.float(), 1),
                                         torch.nn.functional.one_hot(msk, ncls).permute(0, 3, 1, 2).float(),
                                         reduce_batch_first=True))


def north_star_gate(nc, ncls, bilinear, B, H, W, mode="bf16", boundary_coeff=0.2, state="conditioned",
                    vs_torch_gpu=True, storage_tol=2e-2):
    """north_star's OWN tolerances, on a well-conditioned configuration (non-trivial BatchNorm gamma / beta / running
    statistics after ten reference training steps, contour-style masks): bf16 -- logits and gradients within 2e-2
    relative, dice within 1e-3, argmax equal on >= 99.9 % of the pixels; fp32/TF32 -- 1e-3.

    Two oracles: (1) the fp32 CPU restatement of the reference (the bar above); (2) the same restatement with bf16
    *storage* (oracle.Rounding): it differs from a real bf16 implementation by accumulation order only, so every one
    of the 64 gradient tensors has to match it within ``storage_tol`` -- a mis-scaled dgrad / wgrad in a single layer
    cannot hide behind ReLU-mask flips there, not even at random initialisation (state='random')."""
    import unet
    tag = f"ns{nc}_{ncls}_{'bil' if bilinear else 'convT'}_{B}x{H}x{W}_{mode}_{state}"
    if state == "conditioned":
        st = _cond_state(nc, ncls, bilinear)
        img, msk = O.structured_batch(B, nc, ncls, H, W, seed=9)
    else:
        st = O.build_state(nc, ncls, bilinear, seed=0)
        img, msk = O.synthetic_batch(B, nc, ncls, H, W)
    names = O.param_names(st)
    ref_st = {k: v.clone() for k, v in st.items()}
    r_logits, r_loss, r_grads = O.training_step(ref_st, img, msk, ncls, bilinear, boundary_coeff=boundary_coeff)
    model = unet.UNet(nc, ncls, bilinear)
    model.load_state_dict(st)
    model = model.to(DEV).to(memory_format=torch.channels_last).train()
    os.environ["UNET_B200_PRECISION"] = {"tf32": "tf32", "tf32x3": "tf32x3"}.get(mode, "fp32")
    logits, loss, grads = G.unet_step_gpu(model, img, msk, amp=(mode == "bf16"), boundary_coeff=boundary_coeff)
    tol = 2e-2 if mode == "bf16" else 1e-3
    # below ~260 k pixels the deepest layers see 8x8 .. 16x16 maps: a handful of ReLU flips there moves a whole
    # (tiny, ~1e-5) gradient tensor by a few per cent, in bf16 STORAGE of the reference arithmetic itself as well
    # (tests/test_oracle_cond.py: 3.6e-2).  The per-tensor WORST case is therefore gated wider on the small
    # configurations; median, global and every full-size figure are held to north_star's number.
    small = B * H * W < 4 * 256 * 256
    worst_tol = (1e-1 if small else 4e-2) if mode == "bf16" else (2e-2 if small else 5 * tol)
    if storage_tol == 2e-2 and small:
        storage_worst_tol = 4e-2
    else:
        storage_worst_tol = storage_tol
    cat = lambda g: torch.cat([g[k].reshape(-1).double() for k in names])  # noqa: E731
    res = []
    if state == "conditioned":
        l2 = [O.rel_l2(grads[k], r_grads[k]) for k in names]
        mx = [rel(grads[k], r_grads[k]) for k in names]
        res += [(f"{tag}_logits_maxrel", rel(logits, r_logits), tol),
                (f"{tag}_argmax_mismatch", (logits.argmax(1) != r_logits.argmax(1)).float().mean().item(), 1e-3),
                (f"{tag}_dice_abs", abs(_dice_of(logits, msk, ncls) - _dice_of(r_logits, msk, ncls)), 1e-3),
                (f"{tag}_loss_rel", abs(loss - float(r_loss)) / abs(float(r_loss)), 1e-3),
                (f"{tag}_grad_all_rel_l2", O.rel_l2(cat(grads), cat(r_grads)), tol),
                (f"{tag}_grad_l2_median", statistics.median(l2), tol),
                (f"{tag}_grad_maxrel_median", statistics.median(mx), tol),
                (f"{tag}_grad_l2_worst", max(l2), worst_tol)]
        sd = model.state_dict()
        res.append((f"{tag}_running_stats", max(rel(host(sd[k]), ref_st[k]) for k in sd if "running" in k), tol))
    if mode == "bf16":
        e_st = {k: v.clone() for k, v in st.items()}
        e_logits, e_loss, e_grads = O.training_step(e_st, img, msk, ncls, bilinear, boundary_coeff=boundary_coeff,
                                                    q=O.Rounding(torch.bfloat16))
        el2 = {k: O.rel_l2(grads[k], e_grads[k]) for k in names}
        wk = max(el2, key=el2.get)
        res += [(f"{tag}_vs_bf16_storage_oracle_logits", rel(logits, e_logits), storage_tol),
                (f"{tag}_vs_bf16_storage_oracle_loss", abs(loss - float(e_loss)) / abs(float(e_loss)), 1e-3),
                (f"{tag}_vs_bf16_storage_oracle_argmax", (logits.argmax(1) != e_logits.argmax(1)).float().mean().item(), 1e-3),
                (f"{tag}_vs_bf16_storage_oracle_grad_l2_median", statistics.median(el2.values()), storage_tol),
                (f"{tag}_vs_bf16_storage_oracle_grad_l2_worst[{wk}]", el2[wk], storage_worst_tol),
                (f"{tag}_vs_bf16_storage_oracle_grad_all", O.rel_l2(cat(grads), cat(e_grads)), storage_tol)]
        if os.environ.get("UNETB200_TEST_VERBOSE"):
            for k in names:
                print(f"      grad {k:<50s} vs fp32 {O.rel_l2(grads[k], r_grads[k]):.3e}  vs bf16-storage {el2[k]:.3e}  "
                      f"storage-vs-fp32 {O.rel_l2(e_grads[k], r_grads[k]):.3e}")
    if vs_torch_gpu and mode in ("bf16", "tf32"):
        # ours against the reference's own GPU path (torch ATen / cuDNN, test-only) under the same precision
        t_logits, t_loss, t_grads = G.torch_gpu_step(st, img, msk, ncls, bilinear, mode)
        tl2 = [O.rel_l2(grads[k], t_grads[k]) for k in names]
        res += [(f"{tag}_vs_torch_gpu_logits", rel(logits, t_logits), tol if state == "conditioned" else 1e-1),
                (f"{tag}_vs_torch_gpu_grad_l2_median", statistics.median(tl2), tol if state == "conditioned" else 6e-1)]
    return res


def full_size_gate(which):
    """BASELINE.json's configurations at their FULL sizes against the CPU oracle (about a minute of host time each).

    'c2': configs[1]/[3]  UNet(1,2,False) bf16 training step, B=16, 512x512, CE + dice + 0.2*boundary -- against the fp32
          oracle with north_star's bf16 tolerances and against the bf16-storage oracle (every gradient tensor).
    'c3': configs[2]      UNet(1,2,True) bilinear, fp32/TF32 exactness mode, B=16, 512x512 -- 1e-3.
    'c5': configs[4]      UNet(3,4,False).eval() forward, B=8, 1024x1024, bf16 -- logits 2e-2, argmax 99.9 %."""
    import unet
    if which == "c2":
        return north_star_gate(1, 2, False, 16, 512, 512, "bf16", boundary_coeff=0.2, vs_torch_gpu=False)
    if which == "c3":
        mode = os.environ.get("UNETB200_C3_MODE", "tf32x3")
        return north_star_gate(1, 2, True, 16, 512, 512, mode, boundary_coeff=0.2, vs_torch_gpu=False)
    assert which == "c5"
    st = _cond_state(3, 4, False)
    img, _ = O.structured_batch(8, 3, 4, 1024, 1024, seed=9)
    r_logits = O.unet_forward({k: v.clone() for k, v in st.items()}, img, False, training=False)
    e_logits = O.unet_forward({k: v.clone() for k, v in st.items()}, img, False, training=False,
                              q=O.Rounding(torch.bfloat16))
    model = unet.UNet(3, 4, False)
    model.load_state_dict(st)
    model = model.to(DEV).to(memory_format=torch.channels_last).eval()
    x = img.to(DEV).contiguous(memory_format=torch.channels_last)
    with torch.inference_mode(), torch.autocast("cuda", enabled=True):
        logits = host(model(x).float())
    tag = "c5_infer3_4_8x1024x1024_bf16"
    return [(f"{tag}_logits_maxrel", rel(logits, r_logits), 2e-2),
            (f"{tag}_argmax_mismatch", (logits.argmax(1) != r_logits.argmax(1)).float().mean().item(), 1e-3),
            (f"{tag}_vs_bf16_storage_oracle_logits", rel(logits, e_logits), 2e-2),
            (f"{tag}_vs_bf16_storage_oracle_argmax", (logits.argmax(1) != e_logits.argmax(1)).float().mean().item(), 1e-3),
            (f"{tag}_finite", 0.0 if bool(torch.isfinite(logits).all()) else 1.0, 0.0)]


def segments_gate():
    """Scheduling variants of the same backward pass must give the same gradients: (a) loss.backward(); (b) the
    same with data-parallel gradient sinks (kernels write into the bucket views); (c) the four-segment
    torch.autograd.grad form (ddp.segmented_backward); (d) ddp.SegmentedStep -- forward graph, four backward-segment
    graphs, optimizer graph -- replayed.  Single process (world size 1: the all-reduce launches are no-ops; the
    2-GPU NCCL run of the same code is tests/test_ddp_nccl.py)."""
    import unet
    from unetb200 import ddp
    from unetb200 import functional as UF
    from unetb200 import losses as UL
    res = []
    for bilinear in (False, True):
        tag = "seg_" + ("bil" if bilinear else "convT")
        st = O.build_state(1, 2, bilinear, seed=0)
        img, msk = O.synthetic_batch(2, 1, 2, 64, 96)
        x = img.to(DEV).contiguous(memory_format=torch.channels_last)
        t = msk.to(DEV)

        def make():
            m = unet.UNet(1, 2, bilinear)
            m.load_state_dict(st)
            return m.to(DEV).to(memory_format=torch.channels_last).train()

        def fwd_loss(m):
            def f(xx, tt):
                with torch.autocast("cuda", enabled=True):
                    return UL.training_criterion(m(xx), tt, boundary_coeff=0.2)
            return f

        m = make()
        la = fwd_loss(m)(x, t)
        la.backward()
        ga = {k: host(p.grad) for k, p in m.named_parameters()}
        # (b) sinks
        m = make()
        red = ddp.GradAllReducer(m, bucket_bytes=8 << 20)
        fwd_loss(m)(x, t).backward()
        red.finish()
        inside = all(any(p.grad.data_ptr() == v.data_ptr() for v in b["views"]) for b in red.buckets for p in b["params"])
        res.append((f"{tag}_sink_grads_live_in_buckets", 0.0 if inside else 1.0, 0.0))
        res.append((f"{tag}_sink_grads_equal", max(rel(host(p.grad), ga[k]) for k, p in m.named_parameters()), 1e-6))
        red.remove()
        # (c) segmented, eager
        m = make()
        m._taps = {}
        lc = fwd_loss(m)(x, t)
        taps, m._taps = m._taps, None
        segs = ddp.segment_params(m)
        gl = ddp.segmented_backward(m, lc, taps, segs)
        got = {}
        names = {id(p): k for k, p in m.named_parameters()}
        for ps, gs in zip(segs, gl):
            for p, g in zip(ps, gs):
                got[names[id(p)]] = host(g)
        res.append((f"{tag}_segmented_covers_all_params", float(len(set(ga) - set(got))), 0.0))
        res.append((f"{tag}_segmented_grads_equal", max(rel(got[k], ga[k]) for k in ga), 1e-6))
        res.append((f"{tag}_segmented_loss_equal", abs(float(lc) - float(la)), 0.0))
        del taps, gl, lc
        # (d) graphs; the "optimizer" leaves the weights alone so the gradients can be compared after a replay
        m = make()
        ticks = torch.zeros((), device=DEV)
        step = ddp.SegmentedStep(m, fwd_loss(m), lambda: ticks.add_(1), (x, t), warmup=1)
        before = float(ticks)
        nbt0 = int(m.state_dict()["inc.double_conv.1.num_batches_tracked"])   # warm-up steps ran, the captures did not
        for _ in range(2):
            ld = step(x, t)
        torch.cuda.synchronize()
        res.append((f"{tag}_graph_optimizer_replayed", abs(float(ticks) - before - 2.0), 0.0))
        res.append((f"{tag}_graph_loss_equal", abs(float(ld) - float(la)), 1e-6))
        res.append((f"{tag}_graph_grads_equal", max(rel(host(p.grad), ga[k]) for k, p in m.named_parameters()), 1e-6))
        sd = m.state_dict()
        res.append((f"{tag}_graph_bn_steps", float(abs(int(sd["inc.double_conv.1.num_batches_tracked"]) - nbt0 - 2)), 0.0))
        step.release()
        res.append((f"{tag}_sinks_released", float(len(UF._GRAD_SINK)), 0.0))
    return res


def prepack_gate():
    """One multi-tensor launch packs every 3x3 / transposed-conv weight; the packed operands are cached per
    parameter version: FusedRMSprop (raw-pointer updates) must invalidate them, inference must re-use them."""
    import unet
    from unetb200 import functional as UF
    from unetb200 import ops
    from unetb200.optim import FusedRMSprop
    res = []
    torch.manual_seed(3)
    m = unet.UNet(3, 2, False).to(DEV).to(memory_format=torch.channels_last)
    for dt_, tol in ((torch.bfloat16, 0.0), (torch.float32, 0.0)):
        UF._PRE.clear()
        n0 = ops.LAUNCHES
        UF.prepack(m, dt_, need_dgrad=True)
        res.append((f"prepack_one_launch_{dt_}", float(ops.LAUNCHES - n0 - 1), 0.0))
        worst = 0.0
        for mod in m.modules():
            if isinstance(mod, torch.nn.Conv2d) and mod.kernel_size == (3, 3):
                w = mod.weight.detach()
                Co, Ci = w.shape[:2]
                f = UF.pack3x3_fprop(mod.weight, dt_)
                worst = max(worst, float((f.float() - w.permute(0, 2, 3, 1).reshape(Co, 9 * Ci).to(dt_).float()).abs().max()))
                if Ci >= 16:
                    d = UF.pack3x3_dgrad(mod.weight, dt_)
                    ref = w.flip(2, 3).permute(1, 2, 3, 0).reshape(Ci, 9 * Co).to(dt_).float()
                    worst = max(worst, float((d.float() - ref).abs().max()))
            elif isinstance(mod, torch.nn.ConvTranspose2d):
                w = mod.weight.detach()
                Ci, Co = w.shape[:2]
                f = UF.packT_fprop(mod.weight, dt_)
                worst = max(worst, float((f.float() - w.permute(2, 3, 1, 0).reshape(4 * Co, Ci).to(dt_).float()).abs().max()))
                d = UF.packT_dgrad(mod.weight, dt_)
                worst = max(worst, float((d.float() - w.permute(0, 2, 3, 1).reshape(Ci, 4 * Co).to(dt_).float()).abs().max()))
        res.append((f"prepack_values_{dt_}", worst, tol))
        res.append((f"prepack_lookups_launched_nothing_{dt_}", float(ops.LAUNCHES - n0 - 1), 0.0))
    # a contiguous (OIHW) model takes the generic path of the same kernel
    m2 = unet.UNet(1, 2, True).to(DEV)
    UF.prepack(m2, torch.bfloat16, need_dgrad=True)
    w = m2.down1.maxpool_conv[1].double_conv[0].weight
    ref = w.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(64, 9 * 128).bfloat16().float()
    res.append(("prepack_oihw_dgrad", float((UF.pack3x3_dgrad(w, torch.bfloat16).float() - ref).abs().max()), 0.0))
    # inference: second forward re-uses the operands
    m.eval()
    x = torch.rand(1, 3, 64, 64, device=DEV).contiguous(memory_format=torch.channels_last)
    with torch.inference_mode(), torch.autocast("cuda", enabled=True):
        m(x)
        n1 = ops.LAUNCHES
        y1 = m(x)
        per_fwd = ops.LAUNCHES - n1
        UF._PRE.clear()
        n2 = ops.LAUNCHES
        y2 = m(x)
        res.append(("prepack_cached_forward_saves_the_pack_launch", float((ops.LAUNCHES - n2) - per_fwd - 1), 0.0))
        res.append(("prepack_cached_forward_same_logits", float((y1.float() - y2.float()).abs().max()), 0.0))
    # training: the fused optimizer updates through raw pointers and must invalidate the cache
    m.train()
    opt = FusedRMSprop(m.parameters(), lr=1e-2, momentum=0.9)
    with torch.autocast("cuda", enabled=True):
        m(x).float().square().mean().backward()
    w = m.inc.double_conv[3].weight
    v0 = w._version
    before = UF.pack3x3_fprop(w, torch.bfloat16).clone()
    opt.step(clip_max_norm=1.0)
    res.append(("fused_optimizer_bumps_version", 0.0 if w._version > v0 else 1.0, 0.0))
    after = UF.pack3x3_fprop(w, torch.bfloat16)
    ref = w.detach().permute(0, 2, 3, 1).reshape(64, 9 * 64).bfloat16()
    res.append(("repacked_after_optimizer_step", float((after.float() - ref.float()).abs().max()), 0.0))
    res.append(("weights_really_moved", 0.0 if float((after.float() - before.float()).abs().max()) > 0 else 1.0, 0.0))
    return res


GROUPS = {
    "full_c2": lambda gd: full_size_gate("c2"),
    "full_c3": lambda gd: full_size_gate("c3"),
    "full_c5": lambda gd: full_size_gate("c5"),
    "north_star_tf32x3": lambda gd: north_star_gate(1, 2, True, 2, 128, 128, "tf32x3", vs_torch_gpu=False)
                         + north_star_gate(1, 2, False, 2, 64, 96, "tf32x3", vs_torch_gpu=False),
    "segments": lambda gd: segments_gate(),
    "prepack": lambda gd: prepack_gate(),
    "north_star_bf16": lambda gd: north_star_gate(1, 2, False, 2, 128, 128) + north_star_gate(1, 2, False, 4, 256, 256, vs_torch_gpu=False),
    "north_star_bf16_b": lambda gd: north_star_gate(1, 2, True, 2, 128, 128) + north_star_gate(3, 4, False, 2, 128, 160),
    "unet_widths": lambda gd: width_gate("UNet_S", 1, 3, False, 2, 64, 64, "fp32") + width_gate("UNet_S", 1, 3, False, 2, 128, 128, "bf16")
                   + width_gate("UNet_T", 3, 2, True, 1, 64, 96, "fp32") + width_gate("UNet_S", 1, 2, False, 2, 128, 128, "tf32")
                   + width_gate("UNet_T", 1, 2, False, 2, 64, 64, "tf32x3"),
    "unet_sa": lambda gd: sa_gate_op() + width_gate("UNet_SA", 1, 2, False, 2, 64, 64, "fp32")
               + width_gate("UNet_SA", 3, 3, True, 1, 48, 80, "fp32") + width_gate("UNet_SA", 1, 2, False, 2, 128, 128, "bf16"),
    "checkpointing": lambda gd: checkpoint_gate(),
    "graph_side_stream": lambda gd: graph_gate(),
    "unet_infer": lambda gd: infer_gate(3, 4, False, 2, 128, 160, "fp32") + infer_gate(3, 4, False, 2, 128, 160, "bf16")
                  + infer_gate(1, 2, True, 1, 96, 96, "tf32"),
    "unet_fp32": lambda gd: gate(1, 2, False, 2, 64, 64, "fp32") + gate(1, 2, True, 2, 64, 64, "fp32", fused=False),
    "unet_fp32_b": lambda gd: gate(3, 4, False, 1, 96, 80, "fp32") + gate(1, 2, False, 2, 128, 128, "fp32", boundary_coeff=0.2),
    "unet_tf32": lambda gd: gate(1, 2, True, 2, 128, 128, "tf32") + gate(1, 2, False, 2, 128, 128, "tf32"),
    "unet_bf16": lambda gd: gate(1, 2, False, 2, 128, 128, "bf16", boundary_coeff=0.2) + gate(3, 4, False, 1, 160, 96, "bf16"),
    # sizes that are not multiples of 16: the pools floor, Up pads x1 to the skip (unet_parts.py:85-88), odd widths
    "unet_ragged": lambda gd: gate(1, 2, False, 2, 100, 84, "fp32") + gate(1, 2, True, 1, 72, 100, "fp32")
                   + gate(3, 4, False, 2, 200, 136, "bf16") + infer_gate(1, 2, False, 1, 90, 122, "fp32")
                   + width_gate("UNet_S", 1, 2, False, 2, 100, 84, "bf16") + width_gate("UNet_T", 1, 2, False, 1, 72, 100, "fp32"),
    "unet_bf16_bil": lambda gd: gate(1, 2, True, 2, 128, 128, "bf16") + gate(1, 2, False, 4, 256, 256, "bf16", fused=False),
}
