"""Throughput of the ResDiff Cfg-A reverse step over the WeatherBench-shaped field sizes BASELINE names (32x64 ... 128x256), batch 64 per GPU,
bf16, CUDA-graph replay (the 128x256 row is bench.py's headline).  usage: python tools/size_sweep.py [batch]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wsr

LINEAR_1000 = {"schedule": "linear", "n_timestep": 1000, "linear_start": 1e-6, "linear_end": 1e-2}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda:0")
    U = wsr.sub("models.diffusion_models.resdiff.unet").UNet
    D = wsr.sub("models.diffusion_models.resdiff.resdiff_diffusion").ResDiffDiffusion
    networks = wsr.sub("models.diffusion_models.networks")
    builder = wsr.sub("data.dataset_builder")
    for H, W in ((32, 64), (64, 128), (128, 256)):
        torch.manual_seed(0)
        net = U(in_channel=5, out_channel=1, norm_groups=32, inner_channel=64, channel_mults=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2,
                dropout=0.2, image_height=H, image_width=W, image_channels=1)
        networks.init_weights(net, "orthogonal")
        net = net.to(dev).eval()
        diff = D(net, image_height=H, image_width=W, channels=1, conditional=True).to(dev)
        diff.set_new_noise_schedule(LINEAR_1000, dev)
        g = torch.Generator().manual_seed(1234)
        cond = builder.bicubic_sr(torch.randn(B, 1, H // 4, W // 4, generator=g).to(dev), 4)
        plan = diff._plan(B, dev)
        diff._set_condition(plan, cond)
        loop = diff.begin_loop(plan, (B, 1, H, W), seed=7)
        for _ in range(3):
            loop.step()
        loop.capture()
        loop.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        K = 30
        for _ in range(K):
            loop.replay()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / K
        ok = bool(torch.isfinite(loop.x).all())
        print("ResDiff Cfg-A %3dx%-3d B=%d: %.3f ms per reverse step, %.2f samples/s for 1000 steps, %.1f us per sample-step, finite=%s"
              % (H, W, B, ms, B / ms, 1e3 * ms / B, ok))
        del loop, plan, diff, net
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
