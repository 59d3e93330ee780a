"""Worker of tests/test_ddp_nccl.py: one process per GPU under torchrun, NCCL backend.

Checks, on hardware, the semantics SURVEY.md section 8(e) states for the data-parallel step:
  * after the exchange every p.grad equals the MEAN over ranks of the gradients each rank computes alone on its own
    batch (1e-6 relative to the tensor's max), in both ways of driving it: autograd hooks (GradAllReducer.finish)
    and the graph-replayed segmented step (SegmentedStep);
  * BatchNorm running statistics stay per replica (they equal the single-GPU run on that rank's batch and differ
    between ranks);
  * broadcast_module_state replicates rank 0's weights.
Prints one line 'RANK r: OK ...' or raises.
"""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200"), ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import unet
    from oracle import unet_oracle as O
    from unetb200 import ddp
    from unetb200 import losses as UL
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    bilinear = False
    img, msk = O.synthetic_batch(2, 1, 2, 64, 64, rank=rank)          # every rank its own batch
    x = img.to(dev).contiguous(memory_format=torch.channels_last)
    t = msk.to(dev)

    def make(seed):
        torch.manual_seed(seed)
        return unet.UNet(1, 2, bilinear).to(dev).to(memory_format=torch.channels_last).train()

    def fwd_loss(m):
        def f(xx, tt):
            with torch.autocast("cuda", enabled=True):
                return UL.training_criterion(m(xx), tt, boundary_coeff=0.2)
        return f

    # replicated weights: ranks start from different seeds, rank 0 wins
    m = make(100 + rank)
    ddp.broadcast_module_state(m)
    st = {k: v.detach().clone() for k, v in m.state_dict().items()}
    flat = torch.cat([v.reshape(-1).float() for v in st.values()])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    assert all(torch.equal(both[0], b) for b in both), "broadcast_module_state: weights differ between ranks"

    # what each rank computes alone, and the mean over ranks (gathered, averaged in fp64)
    fwd_loss(m)(x, t).backward()
    local_grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    local_running = {k: v.detach().clone() for k, v in m.state_dict().items() if "running" in k}
    expect = {}
    for k, g in local_grads.items():
        flat_g = g.reshape(-1).contiguous()              # logical order (the parameters are channels_last)
        parts = [torch.empty_like(flat_g) for _ in range(world)]
        dist.all_gather(parts, flat_g)
        expect[k] = (sum(p.double() for p in parts) / world).float()

    def check(model, what):
        worst = 0.0
        for k, p in model.named_parameters():
            e = expect[k]
            worst = max(worst, float((p.grad.float().reshape(-1) - e).abs().max() / e.abs().max().clamp_min(1e-30)))
        assert worst <= 1e-6, f"{what}: averaged gradient differs from the mean of the per-rank gradients: {worst:.3e}"
        run = {k: v for k, v in model.state_dict().items() if "running" in k}
        assert all(torch.equal(run[k], local_running[k]) for k in run), f"{what}: BatchNorm statistics are not per replica"
        return worst

    def fresh():
        mm = make(0)
        mm.load_state_dict(st)
        return mm

    # (a) hooks, buckets reduced while backward runs
    m1 = fresh()
    red = ddp.GradAllReducer(m1, bucket_bytes=8 << 20)
    for _ in range(2):                                   # two steps: bucket state must reset (weights unchanged)
        m1.load_state_dict(st)
        m1.zero_grad(set_to_none=True)
        fwd_loss(m1)(x, t).backward()
        red.finish()
        wa = check(m1, "hook mode")
    assert red.launched == 2 * len(red.buckets) and len(red.buckets) > 1
    red.remove()
    # (b) overlap off (all buckets from finish())
    m2 = fresh()
    red = ddp.GradAllReducer(m2, bucket_bytes=8 << 20, overlap=False)
    fwd_loss(m2)(x, t).backward()
    red.finish()
    wb = check(m2, "finish mode")
    red.remove()
    # (c) graph-replayed segmented step; the "optimizer" leaves the weights alone
    m3 = fresh()
    ticks = torch.zeros((), device=dev)
    step = ddp.SegmentedStep(m3, fwd_loss(m3), lambda: ticks.add_(1), (x, t), warmup=1)
    m3.load_state_dict(st)                               # (warm-up and capture moved the running statistics)
    step(x, t)
    torch.cuda.synchronize()
    wc = check(m3, "segmented graphs")
    step.release()
    # the ranks really had different gradients / statistics
    k0 = "inc.double_conv.0.weight"
    g0 = local_grads[k0].reshape(-1).contiguous()
    parts = [torch.empty_like(g0) for _ in range(world)]
    dist.all_gather(parts, g0)
    assert not torch.equal(parts[0], parts[1]), "ranks computed identical gradients: the test is vacuous"
    dist.barrier()
    print(f"RANK {rank}: OK hook {wa:.2e} finish {wb:.2e} segmented {wc:.2e} (max-rel vs mean of per-rank gradients)",
          flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
