"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

fp32 torch-CPU restatement of the diffusion process around the denoiser: the reverse chain and the training loss.
Noise is always INJECTED (a pre-generated tensor), never drawn here, so that runs are comparable across devices.
"""
import numpy as np
import torch

from . import nets
from .schedule import ddpm_tables


def _tables(schedule_opt):
    t32, sqrt_abar_prev = ddpm_tables(schedule_opt)
    return {k: torch.from_numpy(v) for k, v in t32.items()}, sqrt_abar_prev


def p_sample_step(denoise, tab, sqrt_abar_prev, x, t, z, clip=True):
    """diffusion.py:144-192.  ``denoise(x_t, level)`` -> eps_hat.  z is the injected noise for this step (ignored
    at t == 0, where the reference uses zeros)."""
    b = x.shape[0]
    level = torch.full((b, 1), float(np.float32(sqrt_abar_prev[t + 1])), dtype=torch.float32)  # :159-160
    eps = denoise(x, level)
    x0 = tab["sqrt_recip_alphas_cumprod"][t] * x - tab["sqrt_recipm1_alphas_cumprod"][t] * eps   # :124-125
    if clip:
        x0 = x0.clamp(-1.0, 1.0)                                                                 # :168-169
    mean = tab["posterior_mean_coef1"][t] * x0 + tab["posterior_mean_coef2"][t] * x            # :139-140
    logvar = tab["posterior_log_variance_clipped"][t]
    noise = z if t > 0 else torch.zeros_like(x)                                                 # :191
    return mean + noise * (0.5 * logvar).exp(), eps


def resdiff_chain(sd, cfg, schedule_opt, cond, noise, return_eps=False):
    """resdiff/resdiff_diffusion.py:58-108 (conditional branch).  ``noise``: [T+1,B,C,H,W]; noise[0] is the initial
    image (the reference's ``torch.randn(shape)`` at :83), noise[1+k] the k-th ``randn_like`` (steps t=T-1..1)."""
    tab, sap = _tables(schedule_opt)
    T = int(schedule_opt["n_timestep"])
    img = noise[0].clone()
    eps_all = []

    def denoise(x, level):
        return nets.resdiff_unet(sd, torch.cat([cond, x], dim=1), level, cfg)

    k = 1
    for t in reversed(range(T)):
        img, eps = p_sample_step(denoise, tab, sap, img, t, noise[k] if t > 0 else None)
        k += 1
        if return_eps:
            eps_all.append(eps)
    out = img + cond                                                                             # :94
    return (out, eps_all) if return_eps else out


def srdiff_chain(unet_sd, rrdb_sd, cfg, schedule_opt, lr, sr_up, noise, return_eps=False):
    """srdiff/srdiff_diffusion.py:77-159: RRDB features once, then T steps of the SRDiff UNet."""
    tab, sap = _tables(schedule_opt)
    T = int(schedule_opt["n_timestep"])
    _, feas = nets.rrdb_net(rrdb_sd, lr)
    img = noise[0].clone()
    eps_all = []

    def denoise(x, level):
        return nets.srdiff_unet(unet_sd, feas, x, level, cfg)

    k = 1
    for t in reversed(range(T)):
        img, eps = p_sample_step(denoise, tab, sap, img, t, noise[k] if t > 0 else None)
        k += 1
        if return_eps:
            eps_all.append(eps)
    out = img + sr_up
    return (out, eps_all) if return_eps else out


def cond_chain(unet_fn, sd, cfg, schedule_opt, cond, noise, add_cond, return_eps=False):
    """Conditional reverse chain of the SR3 (sr3_diffusion.py:49-84, returns the image) and PhyDiff
    (phydiff_diffusion.py:49-82, returns image + condition) processes; noise layout as in resdiff_chain."""
    tab, sap = _tables(schedule_opt)
    T = int(schedule_opt["n_timestep"])
    img = noise[0].clone()
    eps_all = []

    def denoise(x, level):
        return unet_fn(sd, torch.cat([cond, x], dim=1), level, cfg)

    k = 1
    for t in reversed(range(T)):
        img, eps = p_sample_step(denoise, tab, sap, img, t, noise[k] if t > 0 else None)
        k += 1
        if return_eps:
            eps_all.append(eps)
    out = img + cond if add_cond else img
    return (out, eps_all) if return_eps else out


def q_sample(x0, a, noise):
    """diffusion.py:209-228: a*x0 + sqrt(1-a^2)*noise with a of shape (B,1,1,1)."""
    return a * x0 + (1 - a ** 2).sqrt() * noise


def resdiff_p_losses(sd, cfg, hr, sr, level, noise, loss_type="l1"):
    """resdiff/resdiff_diffusion.py:111-152 with the random draws (t, continuous level, noise) injected:
    ``level`` (B,) is the already-drawn continuous sqrt(alpha_bar).  Returns the SUM-reduced loss (diffusion.py:105-108)
    and eps_hat.  Dropout is not modelled (tests use dropout = 0)."""
    x0 = hr - sr
    x_noisy = q_sample(x0, level.view(-1, 1, 1, 1), noise)
    eps = nets.resdiff_unet(sd, torch.cat([sr, x_noisy], dim=1), level.view(-1, 1), cfg)
    if loss_type == "l1":
        loss = (noise - eps).abs().sum()
    else:
        loss = ((noise - eps) ** 2).sum()
    return loss, eps


def arch_p_losses(arch, sd, cfg, hr, sr, level, noise, loss_type="l1"):
    """p_losses of the 'resdiff' (:111-152), 'phydiff' (phydiff_diffusion.py:98-139; same residual target) and 'sr3'
    (sr3_diffusion.py:101-137; target = HR itself) processes with the random draws injected."""
    if arch == "resdiff":
        return resdiff_p_losses(sd, cfg, hr, sr, level, noise, loss_type)
    x0 = hr if arch == "sr3" else hr - sr
    fn = nets.sr3_unet if arch == "sr3" else nets.phydiff_unet
    x_noisy = q_sample(x0, level.view(-1, 1, 1, 1), noise)
    eps = fn(sd, torch.cat([sr, x_noisy], dim=1), level.view(-1, 1), cfg)
    loss = (noise - eps).abs().sum() if loss_type == "l1" else ((noise - eps) ** 2).sum()
    return loss, eps


def srdiff_param_grads(unet_sd, rrdb_sd, cfg, lr, hr, sr, level, noise, loss_type="l1"):
    """SRDiff training step with the frozen RRDB encoder (srdiff_diffusion.py:161-216, lock_weights=True: no extra RRDB
    loss term): gradients of the UNet parameters (incl. cond_proj) only."""
    with torch.no_grad():
        _, feas = nets.rrdb_net(rrdb_sd, lr)
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in unet_sd.items()}
    x_noisy = q_sample(hr - sr, level.view(-1, 1, 1, 1), noise)
    eps = nets.srdiff_unet(leaf, feas, x_noisy, level.view(-1, 1), cfg)
    loss = (noise - eps).abs().sum() if loss_type == "l1" else ((noise - eps) ** 2).sum()
    (loss / hr.numel()).backward()
    return loss.detach(), {k: v.grad for k, v in leaf.items() if v.grad is not None}


def srdiff_joint_param_grads(unet_sd, rrdb_sd, cfg, lr, hr, sr, level, noise, loss_type="l1"):
    """SRDiff training step with a TRAINABLE encoder (srdiff_diffusion.py:161-216, lock_weights=False): the loss gains
    ``F.l1_loss(rrdb_sr, HR)`` (:212-214) and both parameter sets receive gradients.  Returns (loss, unet grads, rrdb grads)."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in unet_sd.items()}
    rleaf = {k: v.detach().clone().requires_grad_(True) for k, v in rrdb_sd.items()}
    rrdb_sr, feas = nets.rrdb_net(rleaf, lr)
    x_noisy = q_sample(hr - sr, level.view(-1, 1, 1, 1), noise)
    eps = nets.srdiff_unet(leaf, feas, x_noisy, level.view(-1, 1), cfg)
    loss = (noise - eps).abs().sum() if loss_type == "l1" else ((noise - eps) ** 2).sum()
    loss = loss + (rrdb_sr - hr).abs().mean()
    (loss / hr.numel()).backward()
    return loss.detach(), {k: v.grad for k, v in leaf.items() if v.grad is not None}, {k: v.grad for k, v in rleaf.items() if v.grad is not None}


def arch_param_grads(arch, sd, cfg, hr, sr, level, noise, loss_type="l1"):
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    loss, _ = arch_p_losses(arch, leaf, cfg, hr, sr, level, noise, loss_type)
    (loss / hr.numel()).backward()
    return loss.detach(), {k: v.grad for k, v in leaf.items() if v.grad is not None}


def resdiff_param_grads(sd, cfg, hr, sr, level, noise, loss_type="l1"):
    """The reference's training step up to the optimizer (models/diffusion_models/model.py:61-68): sum-loss / numel ->
    backward.  Returns (sum-reduced loss, {name: gradient}) for every entry of ``sd`` that received a gradient."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    loss, _ = resdiff_p_losses(leaf, cfg, hr, sr, level, noise, loss_type)
    (loss / hr.numel()).backward()
    return loss.detach(), {k: v.grad for k, v in leaf.items() if v.grad is not None}
