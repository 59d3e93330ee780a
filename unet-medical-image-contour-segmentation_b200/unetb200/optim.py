"""Fused optimizer step for the UNet training loop: clip_grad_norm_ + RMSprop in two multi-tensor kernels.

``FusedRMSprop`` has torch.optim.RMSprop's constructor, ``state_dict`` layout (``step``, ``square_avg``,
``momentum_buffer``) and arithmetic (reference train.py:80-84); ``step(clip_max_norm=...)`` additionally folds
``torch.nn.utils.clip_grad_norm_(params, max_norm)`` (train.py:157) into the same pass.  All 64 parameter
tensors are updated by one launch each of ``unetb200_grad_sqnorm`` and ``unetb200_rmsprop_step`` instead of
~25 foreach launches; capture-safe (no host sync).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import ops


def _dense(t):
    """Non-overlapping and dense: the strides are a permutation of a contiguous layout."""
    expect = 1
    for st, sz in sorted((st, sz) for sz, st in zip(t.shape, t.stride()) if sz > 1):
        if st != expect:
            return False
        expect *= sz
    return True


class FusedRMSprop(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0, momentum=0, centered=False,
                 capturable=True, foreach=None, maximize=False, differentiable=False):
        if centered or maximize or differentiable:
            raise NotImplementedError("FusedRMSprop: centered / maximize / differentiable are not implemented")
        if lr < 0 or eps < 0 or momentum < 0 or weight_decay < 0 or alpha < 0:
            raise ValueError("FusedRMSprop: negative hyper-parameter")
        super().__init__(params, dict(lr=lr, alpha=alpha, eps=eps, weight_decay=weight_decay, momentum=momentum,
                                      centered=False, capturable=capturable, foreach=foreach, maximize=False,
                                      differentiable=False))
        self._norm = None

    def _tables(self, tensors_lists):
        n = len(tensors_lists[0])
        arrs = []
        for lst in tensors_lists:
            if lst is None:
                arrs.append(None)
                continue
            a = (C.c_void_p * n)(*[t.data_ptr() for t in lst])
            arrs.append(a)
        numel = (C.c_int64 * n)(*[t.numel() for t in tensors_lists[0]])
        return arrs, numel, n

    @torch.no_grad()
    def step(self, closure=None, clip_max_norm=None, write_clipped_grad=True):
        """One update.  clip_max_norm: fold clip_grad_norm_(all params of this optimizer, clip_max_norm) in;
        returns the total gradient norm (a 0-dim device tensor) in that case, else the closure's loss.  Like
        clip_grad_norm_ (train.py:157) the clipped gradients are written back to ``p.grad`` (same pass, one more
        store per element); ``write_clipped_grad=False`` leaves ``p.grad`` unclipped."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        groups = []
        all_grads = []
        steps = []
        for group in self.param_groups:
            ps, gs, sqs, moms = [], [], [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                ops.require_cuda(p, "FusedRMSprop")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise ValueError("FusedRMSprop: fp32 parameters and gradients only")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["square_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if group["momentum"] > 0:
                        st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad
                # the kernel walks raw storage: all four tensors must share one dense element order
                if g.stride() != p.stride():          # re-layout into a temporary; p.grad itself is left alone
                    g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
                if st["square_avg"].stride() != p.stride() or not _dense(p):
                    raise ValueError("FusedRMSprop: parameters must be dense and state must share their layout")
                steps.append(st["step"])
                ps.append(p)
                gs.append(g)
                sqs.append(st["square_avg"])
                if group["momentum"] > 0:
                    moms.append(st["momentum_buffer"])
            if ps:
                groups.append((group, ps, gs, sqs, moms if group["momentum"] > 0 else None))
                all_grads += gs
        if not groups:
            return loss
        torch._foreach_add_(steps, 1)                  # one launch for all step counters
        ops.side_stream_sync()
        norm = None
        sumsq = None
        if clip_max_norm is not None:
            sumsq = torch.empty(1, dtype=torch.float64, device=all_grads[0].device)
            (ga,), numel, n = self._tables([all_grads])
            ops._run("grad_sqnorm", ops.lib().unetb200_grad_sqnorm, ga, numel, n, ops._p(sumsq), ops._stream(),
                     nbytes=4.0 * sum(g.numel() for g in all_grads))
            norm = sumsq.sqrt().float().reshape(())
        for group, ps, gs, sqs, moms in groups:
            (wa, ga, sa, ma), numel, n = self._tables([ps, gs, sqs, moms])
            ops._run("rmsprop_step", ops.lib().unetb200_rmsprop_step, wa, ga, sa, ma, numel, n,
                     ops._p(sumsq) if sumsq is not None else C.c_void_p(0),
                     float(clip_max_norm) if clip_max_norm is not None else 0.0, float(group["lr"]),
                     float(group["alpha"]), float(group["eps"]), float(group["weight_decay"]), float(group["momentum"]),
                     1 if (clip_max_norm is not None and write_clipped_grad) else 0, ops._stream(),
                     nbytes=4.0 * sum(p.numel() for p in ps) * (7 if moms else 5))
            # the kernel writes through raw pointers: tell autograd (saved-tensor checks) and the packed-operand
            # cache of unetb200.functional that these tensors changed
            torch._C._increment_version(ps)
        return norm if clip_max_norm is not None else loss
