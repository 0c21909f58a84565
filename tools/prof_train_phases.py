"""Where does a training step spend its time?  Host wall-clock and device time (CUDA events) per phase, steady state."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import wsr

LINEAR_1000 = {"schedule": "linear", "n_timestep": 1000, "linear_start": 1e-6, "linear_end": 1e-2}
CFG_A = dict(in_channel=5, out_channel=1, norm_groups=32, inner_channel=64, channel_mults=[1, 2, 4, 8, 8], attn_res=[16],
             res_blocks=2, dropout=0.2, image_height=128, image_width=256, image_channels=1)


def main():
    dev = torch.device("cuda:0")
    B = 4
    U = wsr.sub("models.diffusion_models.resdiff.unet").UNet
    D = wsr.sub("models.diffusion_models.resdiff.resdiff_diffusion").ResDiffDiffusion
    glue = wsr.sub("autograd_glue")
    networks = wsr.sub("models.diffusion_models.networks")
    torch.manual_seed(0); np.random.seed(0)
    net = U(precision="bf16", **CFG_A)
    networks.init_weights(net, "orthogonal")
    net = net.to(dev).train()
    diff = D(net, image_height=128, image_width=256, channels=1, conditional=True).to(dev)
    diff.set_new_noise_schedule(LINEAR_1000, dev); diff.set_loss(dev)
    plan = net.train_plan(B, dev)
    opt = glue.FusedAdam(list(diff.parameters()), lr=1e-4)
    opt.attach_flat(plan)
    sr = torch.nn.functional.interpolate(torch.randn(B, 1, 32, 64), scale_factor=4, mode="bicubic").to(dev)
    hr = sr + 0.3 * torch.randn_like(sr)
    numel = hr.numel()
    marks = []

    def mark(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((name, time.perf_counter(), ev))

    orig_refresh = plan.refresh_weights

    def timed_refresh():
        mark("pre_refresh")
        orig_refresh()
        mark("refresh")
    plan.refresh_weights = timed_refresh

    for it in range(7):
        marks.clear()
        mark("start")
        opt.zero_grad()
        loss = diff.p_losses({"HR": hr, "SR": sr})
        mark("forward+loss")
        (loss.sum() / numel).backward()
        mark("backward")
        opt.step()
        mark("adam")
        torch.cuda.synchronize()
        t_end = time.perf_counter()
        if it >= 5:
            print("step %d: host wall %.2f ms" % (it, 1e3 * (t_end - marks[0][1])))
            for (n0, t0, e0), (n1, t1, e1) in zip(marks[:-1], marks[1:]):
                print("   %-14s host %6.2f ms   device %6.2f ms" % (n1, 1e3 * (t1 - t0), e0.elapsed_time(e1)))


if __name__ == "__main__":
    main()
