"""(1) bf16 vs fp32-check-mode drift of the FULL 1000-step reverse chain (same weights, same Philox noise), reported as
final-field RMSE in standardised units and in Kelvin (sigma_K = 21.26 K, the usual WeatherBench t2m standard deviation).
(2) smoke + timing of BASELINE configs[3] (SRDiff + RRDB encoder, batch 32) and configs[4] (3-variable stress config).
usage: python tools/drift_and_configs.py [drift] [c4] [c5]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wsr

LINEAR_1000 = {"schedule": "linear", "n_timestep": 1000, "linear_start": 1e-6, "linear_end": 1e-2}
dev = torch.device("cuda:0")
networks = wsr.sub("models.diffusion_models.networks")


def resdiff(precision, c_img=1, inner=64, in_ch=5):
    U = wsr.sub("models.diffusion_models.resdiff.unet").UNet
    D = wsr.sub("models.diffusion_models.resdiff.resdiff_diffusion").ResDiffDiffusion
    torch.manual_seed(0)
    net = U(in_channel=in_ch, out_channel=c_img, norm_groups=32, inner_channel=inner, channel_mults=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2,
            dropout=0.2, image_height=128, image_width=256, image_channels=c_img, precision=precision)
    networks.init_weights(net, "orthogonal")
    net = net.to(dev).eval()
    diff = D(net, image_height=128, image_width=256, channels=c_img, conditional=True).to(dev)
    diff.set_new_noise_schedule(LINEAR_1000, dev)
    return diff


def drift():
    g = torch.Generator().manual_seed(1234)
    B = 2
    sr = torch.nn.functional.interpolate(torch.randn(B, 1, 32, 64, generator=g), scale_factor=4, mode="bicubic").to(dev)
    outs = {}
    for prec in ("fp32", "bf16"):
        diff = resdiff(prec)
        diff.sample_seed = 77
        t0 = time.time()
        outs[prec] = diff.super_resolution({"SR": sr}).float().cpu()
        torch.cuda.synchronize()
        print("drift: %s chain of 1000 steps, B=%d: %.1f s" % (prec, B, time.time() - t0), flush=True)
    d = outs["bf16"] - outs["fp32"]
    rmse = float(d.pow(2).mean().sqrt())
    rel = float(d.norm() / outs["fp32"].norm())
    print("drift: 1000-step final field bf16 vs fp32 check mode: RMSE %.4e (std units) = %.4f K at sigma 21.26 K; rel-L2 %.3e; field rms %.3f"
          % (rmse, rmse * 21.26, rel, float(outs["fp32"].pow(2).mean().sqrt())), flush=True)


def c4():
    U = wsr.sub("models.diffusion_models.srdiff.unet").UNet
    D = wsr.sub("models.diffusion_models.srdiff.srdiff_diffusion").SRDiffDiffusion
    R = wsr.sub("models.rrdb_encoder.RRDBNet").RRDBNet
    torch.manual_seed(0)
    net = U(in_channel=1, out_channel=1, norm_groups=32, inner_channel=64, channel_mults=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2, dropout=0.2,
            image_height=128, image_width=256, image_channels=1, precision="bf16")
    networks.init_weights(net, "orthogonal")
    diff = D(net.to(dev).eval(), image_height=128, image_width=256, channels=1, conditional=True).to(dev)
    diff.init_rrdb_encoder(None, lock_weights=True)          # random-init RRDB-17 encoder (no checkpoint ships)
    diff.rrdb_encoder.to(dev)
    del R
    sched = dict(LINEAR_1000)
    sched["n_timestep"] = 100
    diff.set_new_noise_schedule(sched, dev)
    B = 32
    lr = torch.randn(B, 1, 32, 64, device=dev)
    sr = torch.nn.functional.interpolate(lr, scale_factor=4, mode="bicubic")
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        out = diff.super_resolution({"LR": lr, "SR": sr, "INTERPOLATED": sr})
        torch.cuda.synchronize(); el = time.time() - t0
    print("c4: SRDiff + RRDB-17 encoder, B=%d, 100-step loop: %.2f s -> %.2f ms/step incl. encoder, finite=%s, %.2f samples/s extrapolated to 1000 steps"
          % (B, el, 1e3 * el / 100, bool(torch.isfinite(out).all()), B / (el * 10)), flush=True)


def c5():
    diff = resdiff("bf16", c_img=3, inner=128, in_ch=15)
    sched = dict(LINEAR_1000)
    sched["n_timestep"] = 50
    diff.set_new_noise_schedule(sched, dev)
    B = 8
    lr = torch.randn(B, 3, 16, 32, device=dev)
    sr = torch.nn.functional.interpolate(lr, scale_factor=8, mode="bicubic")
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        out = diff.super_resolution({"SR": sr})
        torch.cuda.synchronize(); el = time.time() - t0
    fl = 781.32e9 * B * 50 / el / 1e12
    print("c5: 3-variable 8x stress config (inner 128, C_img 3), B=%d, 50-step loop: %.2f s -> %.1f ms/step, %.0f TFLOP/s (781.32 GF/sample-step), finite=%s"
          % (B, el, 1e3 * el / 50, fl, bool(torch.isfinite(out).all())), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["drift", "c4", "c5"]
    for w in which:
        try:
            {"drift": drift, "c4": c4, "c5": c5}[w]()
        except Exception as e:          # keep going: this is a survey of what runs
            import traceback
            traceback.print_exc()
            print("%s: FAILED %s: %s" % (w, type(e).__name__, e), flush=True)
