"""Data-parallel gradient exchange for the UNet step (BASELINE.json configs[3]).

The reference is single-device (train.py:244); this layer has no counterpart there.  Semantics are
standard DDP: replicated weights (broadcast from rank 0), per-replica BatchNorm batch statistics (the
reference uses plain nn.BatchNorm2d), gradients mean-reduced over ranks.  The only exchange step of
the path is that all-reduce, so it is the only collective: NCCL over NVLink 5 / NVSwitch through
torch.distributed, issued per bucket from autograd hooks *while backward is still running* (buckets
fill in reverse layer order: outc, up4 ... inc), on NCCL's own stream, joined once before the
optimizer step.  Works with the gloo backend on CPU tensors too (used by the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _side_stream_sync():
    try:
        from . import ops
    except Exception:  # noqa: BLE001  (CPU-only use of this module in the gloo tests)
        return
    if torch.cuda.is_available():
        ops.side_stream_sync()


def broadcast_module_state(module, src=0):
    """Replicate parameters and buffers from `src` (rank-0 convention for BN running stats)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src)


class GradAllReducer:
    """Bucketed, backward-overlapped mean all-reduce of ``module``'s gradients.

    Usage per step:  ``loss.backward(); reducer.finish()``  (then clip / optimizer.step()).
    After ``finish()`` every ``p.grad`` is a view into its bucket holding the rank-averaged gradient.
    """

    def __init__(self, module, bucket_bytes=32 << 20, group=None, overlap=None):
        # overlap=True launches each bucket's all-reduce from the autograd hook that completes it (while backward
        # is still running); overlap=False launches them all from finish().  The conv kernels are persistent
        # grids of one CTA (pair) per SM, so an NCCL kernel that holds SMs during backward delays whole CTA
        # pairs -- measured on 2 GPUs, see DESIGN.md section 5.  Default: UNETB200_DDP_OVERLAP (0/1), else on.
        if overlap is None:
            import os
            overlap = os.environ.get("UNETB200_DDP_OVERLAP", "1") != "0"
        self.overlap = bool(overlap)
        self.manual = False          # True: hooks are inert, the caller drives pack_all / allreduce_all / point_grads
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.params = [p for p in module.parameters() if p.requires_grad]
        # reverse registration order ~ order in which backward produces the gradients
        order = list(reversed(self.params))
        self.buckets = []          # list of dict(params, offsets, numel, buffer, pending, work)
        cur, cur_bytes = [], 0
        for p in order:
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > bucket_bytes:
                self._close_bucket(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close_bucket(cur)
        self._where = {}
        for bi, b in enumerate(self.buckets):
            for pi, p in enumerate(b["params"]):
                self._where[p] = (bi, pi)
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.launched = 0

    def _close_bucket(self, params):
        offsets, n = [], 0
        for p in params:
            offsets.append(n)
            n += p.numel()
        dev = params[0].device
        self.buckets.append(dict(params=list(params), offsets=offsets, numel=n,
                                 buffer=torch.zeros(n, dtype=torch.float32, device=dev),
                                 pending=len(params), work=None))

    def _views(self, b):
        """Bucket slices shaped AND strided like their parameters (a channels_last parameter gets a channels_last
        gradient view), so that optimizers and AccumulateGrad see the layout they expect and nothing re-copies."""
        out = []
        for p, o in zip(b["params"], b["offsets"]):
            flat = b["buffer"][o:o + p.numel()]
            dense = sorted((st, sz) for sz, st in zip(p.shape, p.stride()) if sz > 1)
            expect, ok = 1, True
            for st, sz in dense:
                ok = ok and st == expect
                expect *= sz
            out.append(flat.as_strided(p.shape, p.stride()) if ok and p.dim() > 0 else flat.view_as(p))
        return out

    # ---- explicit three-phase form (used when the step is replayed as CUDA graphs: the collective stays outside
    # the captured regions) -- pack_all() inside the backward graph, allreduce_all() eagerly, then point_grads()
    def pack_all(self):
        """Copy every gradient into its bucket (no communication)."""
        _side_stream_sync()
        for b in self.buckets:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in b["params"]]
            torch._foreach_copy_(self._views(b), [g.to(torch.float32) for g in grads])

    def allreduce_all(self):
        """Mean all-reduce of every bucket; returns when the results are ordered on the current stream."""
        if self.world == 1:
            return
        works = []
        for b in self.buckets:
            if dist.get_backend(self.group) == "nccl":
                works.append(dist.all_reduce(b["buffer"], op=dist.ReduceOp.AVG, group=self.group, async_op=True))
            else:
                b["buffer"].mul_(1.0 / self.world)
                works.append(dist.all_reduce(b["buffer"], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()

    def point_grads(self):
        """p.grad = its (averaged) bucket view."""
        for b in self.buckets:
            for p, v in zip(b["params"], self._views(b)):
                p.grad = v

    def _on_grad(self, p):
        if self.world == 1 or self.manual:
            return
        bi, _ = self._where[p]
        b = self.buckets[bi]
        b["pending"] -= 1
        if b["pending"] == 0 and self.overlap:
            self._launch(b)

    def _launch(self, b):
        _side_stream_sync()          # weight gradients are produced on the library's side stream
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in b["params"]]
        torch._foreach_copy_(self._views(b), [g.to(torch.float32) for g in grads])
        if dist.get_backend(self.group) == "nccl":       # NCCL averages inside the collective
            b["work"] = dist.all_reduce(b["buffer"], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:
            b["buffer"].mul_(1.0 / self.world)
            b["work"] = dist.all_reduce(b["buffer"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.launched += 1

    def finish(self):
        """Join the outstanding all-reduces and point every p.grad at its averaged bucket view."""
        if self.world == 1:
            return
        for b in self.buckets:
            if b["work"] is None and b["pending"] != len(b["params"]):
                self._launch(b)            # not launched from a hook: overlap off, or parameters without a gradient
        for b in self.buckets:
            if b["work"] is not None:
                b["work"].wait()
                for p, v in zip(b["params"], self._views(b)):
                    if p.grad is not None:
                        p.grad = v.to(p.grad.dtype) if p.grad.dtype != torch.float32 else v
                b["work"] = None
            b["pending"] = len(b["params"])

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
