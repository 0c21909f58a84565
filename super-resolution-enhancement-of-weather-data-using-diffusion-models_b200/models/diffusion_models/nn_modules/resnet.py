"""Parameter containers mirroring models/diffusion_models/nn_modules/resnet.py of the reference (Block :7-28,
ResnetBlock :31-59, SelfAttention :62-100, ResnetBlocWithAttn :103-128, FeatureWiseAffine :131-157): identical class
names, constructor arguments, sub-module names and therefore state_dict keys.  Arithmetic: ``unet_plan.py``.
"""
from torch import nn

from .functional_layers import EngineOnly, Swish


class Block(EngineOnly):
    def __init__(self, dim, dim_out, groups=32, dropout=0):
        super().__init__()
        self.dropout = dropout
        self.block = nn.Sequential(
            nn.GroupNorm(groups, dim),
            Swish(),
            nn.Dropout(dropout) if dropout != 0 else nn.Identity(),
            nn.Conv2d(dim, dim_out, 3, padding=1),
        )


class FeatureWiseAffine(EngineOnly):
    def __init__(self, in_channels, out_channels, use_affine_level=False):
        super().__init__()
        if use_affine_level:
            raise NotImplementedError("use_affine_level=True is never used by the reference (resnet.py:41-42)")
        self.use_affine_level = use_affine_level
        self.noise_func = nn.Sequential(nn.Linear(in_channels, out_channels))


class ResnetBlock(EngineOnly):
    def __init__(self, dim, dim_out, noise_level_emb_dim=None, dropout=0, use_affine_level=False, norm_groups=32):
        super().__init__()
        self.noise_func = FeatureWiseAffine(noise_level_emb_dim, dim_out, use_affine_level)
        self.block1 = Block(dim, dim_out, groups=norm_groups)
        self.block2 = Block(dim_out, dim_out, groups=norm_groups, dropout=dropout)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()


class SelfAttention(EngineOnly):
    def __init__(self, in_channel, n_head=1, norm_groups=32):
        super().__init__()
        if n_head != 1:
            raise NotImplementedError("the reference only instantiates n_head=1 (resnet.py:72,122)")
        self.n_head = n_head
        self.norm = nn.GroupNorm(norm_groups, in_channel)
        self.qkv = nn.Conv2d(in_channel, in_channel * 3, 1, bias=False)
        self.out = nn.Conv2d(in_channel, in_channel, 1)


class ResnetBlocWithAttn(EngineOnly):
    def __init__(self, dim, dim_out, *, noise_level_emb_dim=None, norm_groups=32, dropout=0, with_attn=False):
        super().__init__()
        self.with_attn = with_attn
        self.res_block = ResnetBlock(dim, dim_out, noise_level_emb_dim, norm_groups=norm_groups, dropout=dropout)
        if with_attn:
            self.attn = SelfAttention(dim_out, norm_groups=norm_groups)
