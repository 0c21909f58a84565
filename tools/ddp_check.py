"""2+ GPU check of the data-parallel training step (run under torchrun): every rank runs DDPM-style
optimize_parameters pieces on its shard with the FlatGradReducer; the all-reduced flat gradient must equal the sum of
the gradients of all shards computed locally on one GPU (for 'resdiff' the independent unit is the LOCAL batch).
usage: torchrun --nproc-per-node 2 tools/ddp_check.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import wsr


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle.cases import LINEAR_1000, unet_cfg
    from oracle.weights import fill_module, seeded_randn
    cfg = unet_cfg(32, 64, attn_res=(4,))
    U = wsr.sub("models.diffusion_models.resdiff.unet").UNet
    D = wsr.sub("models.diffusion_models.resdiff.resdiff_diffusion").ResDiffDiffusion
    par = wsr.sub("parallel")
    net = U(in_channel=5, out_channel=1, norm_groups=32, inner_channel=64, channel_mults=cfg["channel_mults"], attn_res=cfg["attn_res"],
            res_blocks=2, dropout=0, image_height=32, image_width=64, image_channels=1, precision="fp32")
    net = fill_module(net, 7).to(dev).train()
    diff = D(net, image_height=32, image_width=64, channels=1, conditional=True).to(dev)
    diff.set_new_noise_schedule(LINEAR_1000, dev)
    diff.set_loss(dev)
    B = 2                                   # per rank
    sr = seeded_randn("ddp.sr", (B * world, 1, 32, 64), 1).to(dev)
    hr = sr + 0.3 * seeded_randn("ddp.hr", (B * world, 1, 32, 64), 1).to(dev)
    noise = seeded_randn("ddp.noise", (B * world, 1, 32, 64), 1).to(dev)
    numel = hr.numel()
    plan = net.train_plan(B, dev)

    def grads_of_shard(r, reducer_on):
        np.random.seed(100 + r)             # same t / levels for shard r wherever it is computed
        lo, hi = r * B, (r + 1) * B
        for p in diff.parameters():
            p.grad = None
        loss = diff.p_losses({"HR": hr[lo:hi], "SR": sr[lo:hi]}, noise=noise[lo:hi])
        (loss.sum() / numel).backward()
        return plan.gflat.clone()

    # reference: all shards locally, no reduction
    plan.on_ready = None
    local_sum = sum(grads_of_shard(r, False) for r in range(world))
    # data-parallel: own shard, bucketed all-reduce launched during backward
    red = par.FlatGradReducer(plan, bucket_mb=8.0)
    g = grads_of_shard(rank, True)
    buckets = red.finish()
    torch.cuda.synchronize(dev)
    got = plan.gflat
    err = float((got - local_sum).norm() / local_sum.norm())
    print("rank %d: %d buckets, all-reduced gradient vs sum of local shard gradients rel-L2 = %.3e" % (rank, len(buckets), err), flush=True)
    assert err < 1e-5 and len(buckets) >= 3
    del g
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
