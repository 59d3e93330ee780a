"""CPU, world_size 2, gloo: the data-parallel gradient exchange (unetb200.ddp) -- bucketing, hook-driven
launch during backward, mean reduction, rank-0 broadcast of parameters/buffers."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "unet-medical-image-contour-segmentation_b200"))
    from unetb200 import ddp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                      # deliberately different initial weights
        model = nn.Sequential(nn.Conv2d(2, 4, 3, padding=1), nn.BatchNorm2d(4), nn.ReLU(), nn.Conv2d(4, 3, 1),
                              nn.Flatten(), nn.Linear(3 * 6 * 6, 5))
        ddp.broadcast_module_state(model)
        w0 = [p.detach().clone() for p in model.parameters()]
        gathered = [[torch.zeros_like(w) for _ in range(world)] for w in w0]
        for w, g in zip(w0, gathered):
            dist.all_gather(g, w)
        same_weights = all(torch.equal(g[0], g[1]) for g in gathered)

        red = ddp.GradAllReducer(model, bucket_bytes=256)   # tiny buckets -> several of them
        nb = len(red.buckets)
        x = torch.randn(3, 2, 6, 6, generator=torch.Generator().manual_seed(7 + rank))
        # local gradients without the reducer's effect: a twin model
        twin = nn.Sequential(nn.Conv2d(2, 4, 3, padding=1), nn.BatchNorm2d(4), nn.ReLU(), nn.Conv2d(4, 3, 1),
                             nn.Flatten(), nn.Linear(3 * 6 * 6, 5))
        twin.load_state_dict(model.state_dict())
        twin(x).square().sum().backward()
        local = [p.grad.clone() for p in twin.parameters()]
        expect = []
        for g in local:
            parts = [torch.zeros_like(g) for _ in range(world)]
            dist.all_gather(parts, g)
            expect.append(sum(parts) / world)

        ok_steps = True
        for _ in range(2):                                  # two steps: bucket state must reset
            model.zero_grad(set_to_none=True)
            model(x).square().sum().backward()
            red.finish()
            for p, e in zip(model.parameters(), expect):
                ok_steps &= torch.allclose(p.grad, e, rtol=1e-5, atol=1e-6)
        launched_hooks = red.launched
        # explicit three-phase form (the CUDA-graph step): channels_last parameters, gradients keep their layout,
        # and p.grad stays a live view of the bucket (a later pack_all + allreduce_all must show through it)
        model = model.to(memory_format=torch.channels_last)
        red.remove()
        red2 = ddp.GradAllReducer(model, bucket_bytes=256)
        red2.manual = True
        ok_manual = True
        for it in range(2):
            model.zero_grad(set_to_none=True)
            model(x).square().sum().backward()
            red2.pack_all()
            red2.allreduce_all()
            if it == 0:
                red2.point_grads()
                views = [p.grad for p in model.parameters()]
            else:                                           # second round: read through the views taken in round one
                for p, v in zip(model.parameters(), views):
                    p.grad = v
            for p, e in zip(model.parameters(), expect):
                ok_manual &= torch.allclose(p.grad, e, rtol=1e-5, atol=1e-6)
                ok_manual &= p.grad.stride() == p.stride()
        out[rank] = (same_weights, nb, bool(ok_steps) and bool(ok_manual), launched_hooks)
    finally:
        dist.destroy_process_group()


def test_grad_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        same_weights, nb, ok, launched = out[rank]
        assert same_weights, "broadcast_module_state did not replicate rank 0's parameters"
        assert nb > 1, "expected several buckets"
        assert ok, "averaged gradients differ from the mean of the per-rank gradients"
        assert launched == 2 * nb


def test_single_process_is_a_noop():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "unet-medical-image-contour-segmentation_b200"))
    from unetb200 import ddp
    m = nn.Linear(4, 3)
    red = ddp.GradAllReducer(m)
    m(torch.ones(2, 4)).sum().backward()
    g = m.weight.grad.clone()
    red.finish()
    assert torch.equal(m.weight.grad, g) and red.launched == 0
