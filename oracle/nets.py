"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functional fp32 torch-CPU restatement of the networks on the hot path.  Every function takes a reference-layout
``state_dict`` (OIHW fp32 conv weights, the reference's key names) plus tensors, and returns tensors; no module
classes of the reference are used.  Each function cites the reference lines it restates.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------
# small pieces
# ----------------------------------------------------------------------------------------------------------------
def swish(x):
    """nn_modules/functional_layers.py:44-47."""
    return x * torch.sigmoid(x)


def mish(x):
    """nn_modules/functional_layers.py:49-52."""
    return x * torch.tanh(F.softplus(x))


def positional_encoding(level, dim):
    """nn_modules/functional_layers.py:33-41.  level (B,1) -> (B,1,dim) (the middle dim is kept)."""
    half = dim // 2
    k = torch.arange(half, dtype=level.dtype, device=level.device) / half
    arg = level.unsqueeze(1) * torch.exp(-math.log(1e4) * k.unsqueeze(0))
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)


def linear(sd, p, x):
    return F.linear(x, sd[p + "weight"], sd.get(p + "bias"))


def conv(sd, p, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + "weight"], sd.get(p + "bias"), stride=stride, padding=padding)


def group_norm(sd, p, x, groups):
    return F.group_norm(x, groups, sd[p + "weight"], sd[p + "bias"], eps=1e-5)


def noise_level_mlp(sd, level, inner, act=swish, p="noise_level_mlp."):
    """resdiff/unet.py:46-53,135 (Swish) and srdiff/unet.py:49-54 (Mish)."""
    e = positional_encoding(level, inner)
    return linear(sd, p + "3.", act(linear(sd, p + "1.", e)))


# Training-mode dropout (nn_modules/resnet.py:23, only block2 has p != 0): an iterator of keep-scale masks (0 or 1/(1-p), shape of
# the activation), consumed in execution order by every block2; None = eval mode / p = 0.  The GPU tests export the masks the
# CUDA path used (wsr_dropout_mask) so that both sides apply the SAME mask.
DROP_MASKS = None


def block(sd, p, x, groups, drop=False):
    """nn_modules/resnet.py:19-28: GN -> Swish -> (Dropout) -> conv3x3."""
    a = swish(group_norm(sd, p + "block.0.", x, groups))
    if drop and DROP_MASKS is not None:
        a = a * next(DROP_MASKS)
    return conv(sd, p + "block.3.", a, padding=1)


def resnet_block(sd, p, x, t_emb, groups):
    """nn_modules/resnet.py:44-59 (use_affine_level False: resnet.py:155-156)."""
    b = x.shape[0]
    h = block(sd, p + "block1.", x, groups)
    h = h + linear(sd, p + "noise_func.noise_func.0.", t_emb).view(b, -1, 1, 1)
    h = block(sd, p + "block2.", h, groups, drop=True)
    skip = conv(sd, p + "res_conv.", x) if (p + "res_conv.weight") in sd else x
    return h + skip


def _attend(q, k, v, channel):
    """Single-head dense attention over flattened pixels; scale 1/sqrt(C) after QK^T
    (resnet.py:90-97, guided_cross_attention.py:34-41).  q,k,v: (B,C,H,W)."""
    b, c, h, w = k.shape
    qf = q.reshape(b, c, -1)                      # (B,C,Nq)
    kf = k.reshape(b, c, -1)
    vf = v.reshape(b, c, -1)
    s = torch.einsum("bcq,bck->bqk", qf, kf) / math.sqrt(channel)
    pr = torch.softmax(s, dim=-1)
    o = torch.einsum("bqk,bck->bcq", pr, vf)
    return o.reshape(b, c, q.shape[2], q.shape[3])


def self_attention(sd, p, x, groups):
    """nn_modules/resnet.py:81-100, n_head = 1."""
    c = x.shape[1]
    n = group_norm(sd, p + "norm.", x, groups)
    q, k, v = conv(sd, p + "qkv.", n).chunk(3, dim=1)
    return conv(sd, p + "out.", _attend(q, k, v, c)) + x


def hf_guided_ca(sd, p, feat, query_img, groups=32):
    """resdiff/guided_cross_attention.py:24-44."""
    c = feat.shape[1]
    n = group_norm(sd, p + "norm.", feat, groups)
    k, v = conv(sd, p + "kv.", n).chunk(2, dim=1)
    q = conv(sd, p + "q.", query_img)
    return conv(sd, p + "out.", _attend(q, k, v, c)) + feat


def res_block_with_attn(sd, p, x, t_emb, groups):
    """nn_modules/resnet.py:124-128."""
    x = resnet_block(sd, p + "res_block.", x, t_emb, groups)
    if (p + "attn.qkv.weight") in sd:
        x = self_attention(sd, p + "attn.", x, groups)
    return x


def res_se(sd, p, x):
    """resdiff/fd_info_spliter.py:142-148: x * sigmoid(fc(avgpool(x))) + x."""
    y = x.mean(dim=(2, 3))
    y = torch.sigmoid(F.linear(torch.relu(F.linear(y, sd[p + "fc.0.weight"])), sd[p + "fc.2.weight"]))
    return x * y[:, :, None, None] + x


def haar_detail_sums(img, levels=4):
    """resdiff/unet.py:124-132 with pytorch_wavelets.DWTForward(J, 'haar', 'symmetric'): per level the SUM of the
    three detail bands.  Convention (pytorch_wavelets 1.3.0 / pywt Haar; see oracle/ref_shims.py): on a 2x2 block
    ``a b / c d`` LL=(a+b+c+d)/2, LH=(a+b-c-d)/2, HL=(a-b+c-d)/2, HH=(a-b-c+d)/2."""
    out = []
    ll = img
    for _ in range(levels):
        a = ll[:, :, 0::2, 0::2]
        b = ll[:, :, 0::2, 1::2]
        c = ll[:, :, 1::2, 0::2]
        d = ll[:, :, 1::2, 1::2]
        out.append((3 * a - b - c - d) / 2)
        ll = (a + b + c + d) / 2
    return out


def fd_info_spliter(sd, p, x_cat, t_emb, in_ch, height, width):
    """resdiff/fd_info_spliter.py:37-117.  x_cat = cat([cnn_x, x], 1); returns cat([x, cnn_x, denoise_x, lf, hf])."""
    cnn_x, x = torch.split(x_cat, in_ch, dim=1)
    b = x.shape[0]
    # noise-image suppression (:43-47): Linear(dim -> W) tiled over channels and rows, ResSE gate
    ne = linear(sd, p + "noise_func.", t_emb.view(b, -1))
    ne = ne[:, None, None, :].repeat(1, in_ch, height, 1)
    denoise_x = x * res_se(sd, p + "noise_resSE.", ne)
    # FFT over ALL FOUR dims (no dim= argument, :61-63)
    n, m = x.shape[-2:]
    u = torch.arange(n, dtype=torch.float32, device=x.device)[:, None] - n / 2
    v = torch.arange(m, dtype=torch.float32, device=x.device)[None, :] - m / 2
    spec = torch.fft.fftn(torch.complex(cnn_x, torch.zeros_like(cnn_x)))
    x_fd = torch.cat([spec.real, spec.imag], dim=1)
    ell = min(height, width)
    sig = torch.abs(res_se(sd, p + "sigma_resSE.", x_fd).mean(dim=(2, 3), keepdim=True).mean(dim=1)) + ell / 2
    sig = torch.minimum(sig, torch.full_like(sig, float(ell - 10)))          # (B,1,1)
    dist = torch.sqrt(u ** 2 + v ** 2)
    hp = (1 - torch.exp(-dist ** 2 / (2 * sig ** 2))).unsqueeze(1)           # (B,1,H,W), unshifted spectrum
    if in_ch > 1:
        hp = torch.cat([hp] * in_ch, dim=1)
    spec_f = spec * hp
    gate = res_se(sd, p + "HF_guided_resSE.", torch.cat([spec_f.real, spec_f.imag], dim=1))
    x_lf = cnn_x * conv(sd, p + "channel_transform.", gate)
    x_hf = torch.abs(torch.fft.ifftn(spec_f))
    return torch.cat([x, cnn_x, denoise_x, x_lf, x_hf], dim=1)


# ----------------------------------------------------------------------------------------------------------------
# UNets
# ----------------------------------------------------------------------------------------------------------------
def _layer_kinds(sd, prefix):
    """Reconstruct the module list of ``downs`` / ``ups`` from the state_dict keys."""
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix)})
    kinds = []
    for i in idx:
        base = "%s%d." % (prefix, i)
        if (base + "res_block.block1.block.3.weight") in sd:
            kinds.append("res")
        elif (base + "conv.weight") in sd:
            kinds.append("resample")
        else:
            kinds.append("conv")
    return kinds


def resdiff_unet(sd, x, level, cfg):
    """resdiff/unet.py:121-177 (eval mode).  x = cat([cond, x_t], 1) (B, 2*C_img, H, W); level (B,1)."""
    g = cfg.get("norm_groups", 32)
    inner = cfg["inner_channel"]
    c_img = cfg["image_channels"]
    cond = x[:, :c_img]
    queries = haar_detail_sums(cond, 4)
    t = noise_level_mlp(sd, level, inner, swish)
    x = fd_info_spliter(sd, "fd_spliter.", x, t, c_img, cfg["image_height"], cfg["image_width"])
    feats = []
    hf_i = 0
    for i, kind in enumerate(_layer_kinds(sd, "downs.")):
        p = "downs.%d." % i
        if kind == "conv":
            x = conv(sd, p, x, padding=1)
        elif kind == "res":
            x = res_block_with_attn(sd, p, x, t, g)
        else:
            x = conv(sd, p + "conv.", x, stride=2, padding=1)               # functional_layers.py:79-82
        if feats and feats[-1].shape[2:] != x.shape[2:]:
            feats.append(hf_guided_ca(sd, "hf_ca_list.%d." % hf_i, x, queries[hf_i], 32))
            hf_i += 1
        else:
            feats.append(x)
    for i in range(2):
        x = res_block_with_attn(sd, "mid.%d." % i, x, t, g)
    for i, kind in enumerate(_layer_kinds(sd, "ups.")):
        p = "ups.%d." % i
        if kind == "res":
            x = res_block_with_attn(sd, p, torch.cat([x, feats.pop()], dim=1), t, g)
        else:
            x = conv(sd, p + "conv.", F.interpolate(x, scale_factor=2, mode="nearest"), padding=1)
    return block(sd, "final_conv.", x, g)


def srdiff_unet(sd, feas, x, level, cfg):
    """srdiff/unet.py:112-141 (eval mode).  feas: list of 18 (B,64,H/4,W/4); x = x_t (B,C,H,W)."""
    g = cfg.get("norm_groups", 32)
    t = noise_level_mlp(sd, level, cfg["inner_channel"], mish)
    cond = F.conv_transpose2d(torch.cat(feas[2::3], 1), sd["cond_proj.weight"], sd["cond_proj.bias"],
                              stride=4, padding=2)
    feats = []
    for i, kind in enumerate(_layer_kinds(sd, "downs.")):
        p = "downs.%d." % i
        if kind == "conv":
            x = conv(sd, p, x, padding=1)
        elif kind == "res":
            x = res_block_with_attn(sd, p, x, t, g)
        else:
            x = conv(sd, p + "conv.", x, stride=2, padding=1)
        if i == 2:
            x = x + cond
        feats.append(x)
    for i in range(2):
        x = res_block_with_attn(sd, "mid.%d." % i, x, t, g)
    for i, kind in enumerate(_layer_kinds(sd, "ups.")):
        p = "ups.%d." % i
        if kind == "res":
            x = res_block_with_attn(sd, p, torch.cat([x, feats.pop()], dim=1), t, g)
        else:
            x = conv(sd, p + "conv.", F.interpolate(x, scale_factor=2, mode="nearest"), padding=1)
    return block(sd, "final_conv.", x, g)


def haar_detail_bands(img, levels=4):
    """phydiff/unet.py:265-276: per level cat([LH, HL, HH], dim=1) of the Haar analysis of the condition (same band
    convention as haar_detail_sums; PhyDiff keeps the three bands apart, so their ORDER matters here)."""
    out = []
    ll = img
    for _ in range(levels):
        a = ll[:, :, 0::2, 0::2]
        b = ll[:, :, 0::2, 1::2]
        c = ll[:, :, 1::2, 0::2]
        d = ll[:, :, 1::2, 1::2]
        out.append(torch.cat([(a + b - c - d) / 2, (a - b + c - d) / 2, (a - b - c + d) / 2], dim=1))
        ll = (a + b + c + d) / 2
    return out


def phy_stencils(cond):
    """phydiff/unet.py:189-196,311-314: forward differences in x and y and the 5-point Laplacian of the REFLECT-padded
    condition; each is a conv2d with a (1, C_img, 3, 3) kernel, i.e. the stencil summed over the image channels.
    Returns (B, 3, H, W)."""
    c = cond.shape[1]
    kx = torch.tensor([[0, 0, 0], [0, -1, 1], [0, 0, 0]], dtype=torch.float32).view(1, 1, 3, 3).repeat(1, c, 1, 1)
    ky = torch.tensor([[0, 0, 0], [0, -1, 0], [0, 1, 0]], dtype=torch.float32).view(1, 1, 3, 3).repeat(1, c, 1, 1)
    kxy = torch.tensor([[0, 1, 0], [1, -4, 1], [0, 1, 0]], dtype=torch.float32).view(1, 1, 3, 3).repeat(1, c, 1, 1)
    padded = F.pad(cond, (1, 1, 1, 1), mode="reflect")
    return torch.cat([F.conv2d(padded, kx), F.conv2d(padded, ky), F.conv2d(padded, kxy)], dim=1)


def _unet_trunk(sd, x, t, g, queries=None):
    """downs / mid / ups / final_conv shared by the SR3 and PhyDiff UNets; ``queries`` (PhyDiff) = the HF-guided
    cross-attention query images, one per Downsample (the attended tensor replaces the SKIP only)."""
    feats = []
    hf_i = 0
    for i, kind in enumerate(_layer_kinds(sd, "downs.")):
        p = "downs.%d." % i
        if kind == "conv":
            x = conv(sd, p, x, padding=1)
        elif kind == "res":
            x = res_block_with_attn(sd, p, x, t, g)
        else:
            x = conv(sd, p + "conv.", x, stride=2, padding=1)
        if queries is not None and feats and feats[-1].shape[2:] != x.shape[2:]:
            feats.append(hf_guided_ca(sd, "hf_ca_list.%d." % hf_i, x, queries[hf_i], 32))
            hf_i += 1
        else:
            feats.append(x)
    for i in range(len(_layer_kinds(sd, "mid."))):            # SR3 has ONE mid block (sr3/unet.py:77-81), PhyDiff two
        x = res_block_with_attn(sd, "mid.%d." % i, x, t, g)
    for i, kind in enumerate(_layer_kinds(sd, "ups.")):
        p = "ups.%d." % i
        if kind == "res":
            x = res_block_with_attn(sd, p, torch.cat([x, feats.pop()], dim=1), t, g)
        else:
            x = conv(sd, p + "conv.", F.interpolate(x, scale_factor=2, mode="nearest"), padding=1)
    return block(sd, "final_conv.", x, g)


def sr3_unet(sd, x, level, cfg):
    """sr3/unet.py:100-124 (eval mode).  x = cat([cond, x_t], 1) (B, 2*C_img, H, W); level (B,1)."""
    t = noise_level_mlp(sd, level, cfg["inner_channel"], swish)
    return _unet_trunk(sd, x, t, cfg.get("norm_groups", 32))


def phydiff_unet(sd, x, level, cfg):
    """phydiff/unet.py:262-346 (eval mode).  x = cat([cond, x_t], 1); the stem sees cat([x, Kx, Ky, Kxy](cond))."""
    c_img = cfg["image_channels"]
    cond = x[:, :c_img]
    queries = haar_detail_bands(cond, 4)
    x = torch.cat([x, phy_stencils(cond)], dim=1)
    t = noise_level_mlp(sd, level, cfg["inner_channel"], swish)
    return _unet_trunk(sd, x, t, cfg.get("norm_groups", 32), queries)


# ----------------------------------------------------------------------------------------------------------------
# priors
# ----------------------------------------------------------------------------------------------------------------
def simple_cnn(sd, lr):
    """simple_cnn/Simple_CNN.py:24-32 (bicubic x4 hard-coded at :25)."""
    up = F.interpolate(lr, scale_factor=4, mode="bicubic", align_corners=False)
    h = torch.relu(conv(sd, "conv1.", lr, padding=1))
    h = torch.relu(conv(sd, "conv2.", h, padding=1))
    return F.pixel_shuffle(conv(sd, "conv3.", h, padding=1), 4) + up


def _rdb(sd, p, x):
    """rrdb_encoder/RRDBNet.py:105-111."""
    feats = [x]
    for i in range(1, 5):
        feats.append(F.leaky_relu(conv(sd, "%sconv%d." % (p, i), torch.cat(feats, 1), padding=1), 0.2))
    return conv(sd, p + "conv5.", torch.cat(feats, 1), padding=1) * 0.2 + x


def rrdb_net(sd, lr, n_blocks=17):
    """rrdb_encoder/RRDBNet.py:38-59.  Returns (sr_image, list of n_blocks+1 feature maps)."""
    x = (lr + 1) / 2
    first = fea = conv(sd, "conv_first.", x, padding=1)
    feas = []
    for i in range(n_blocks):
        p = "RRDB_trunk.%d." % i
        out = _rdb(sd, p + "RDB3.", _rdb(sd, p + "RDB2.", _rdb(sd, p + "RDB1.", fea)))
        fea = out * 0.2 + fea
        feas.append(fea)
    fea = first + conv(sd, "trunk_conv.", fea, padding=1)
    feas.append(fea)
    fea = F.leaky_relu(conv(sd, "upconv1.", F.interpolate(fea, scale_factor=2, mode="nearest"), padding=1), 0.2)
    fea = F.leaky_relu(conv(sd, "upconv2.", F.interpolate(fea, scale_factor=2, mode="nearest"), padding=1), 0.2)
    out = conv(sd, "conv_last.", F.leaky_relu(conv(sd, "HRconv.", fea, padding=1), 0.2), padding=1)
    return out.clamp(0, 1) * 2 - 1, feas
