"""Architecture registry -- drop-in for the reference's models/diffusion_models/networks.py:56-169.
``define_diffusion(opt)`` maps ``opt['model']['architecture']`` to a (UNet, *Diffusion) pair with the reference's
constructor arguments.  Accelerated architectures: ``resdiff``, ``srdiff`` and ``sr3``, ``phydiff`` (8f N1); ``physrdiff`` raises
NotImplementedError, like an unknown name does in the reference (:133-134).

Extra optional config key (default reproduces the reference): ``model.precision`` = ``"bf16"`` | ``"fp32"``.
"""
import functools
import logging

from torch.nn import init

logger = logging.getLogger('base')


def weights_init_normal(m, std=0.02):
    name = type(m).__name__
    if 'Conv' in name or 'Linear' in name:
        if getattr(m, 'weight', None) is None:
            return
        init.normal_(m.weight.data, 0.0, std)
        if m.bias is not None:
            m.bias.data.zero_()
    elif 'BatchNorm2d' in name:
        init.normal_(m.weight.data, 1.0, std)
        init.constant_(m.bias.data, 0.0)


def weights_init_kaiming(m, scale=1):
    name = type(m).__name__
    if 'Conv2d' in name or 'Linear' in name:
        if getattr(m, 'weight', None) is None:
            return
        init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
        m.weight.data *= scale
        if m.bias is not None:
            m.bias.data.zero_()
    elif 'BatchNorm2d' in name:
        init.constant_(m.weight.data, 1.0)
        init.constant_(m.bias.data, 0.0)


def weights_init_orthogonal(m):
    """reference :56-77 -- orthogonal weights / zero bias for every *Conv* and *Linear* module (GroupNorm untouched)."""
    name = type(m).__name__
    if 'PhyConv' in name:
        return
    if 'Conv' in name or 'Linear' in name:
        if getattr(m, 'weight', None) is None:
            return
        init.orthogonal_(m.weight.data, gain=1)
        if m.bias is not None:
            m.bias.data.zero_()
    elif 'BatchNorm2d' in name:
        init.constant_(m.weight.data, 1.0)
        init.constant_(m.bias.data, 0.0)


def init_weights(net, init_type='kaiming', scale=1, std=0.02):
    logger.info('Initialization method [{:s}]'.format(init_type))
    if init_type == 'normal':
        net.apply(functools.partial(weights_init_normal, std=std))
    elif init_type == 'kaiming':
        net.apply(functools.partial(weights_init_kaiming, scale=scale))
    elif init_type == 'orthogonal':
        net.apply(weights_init_orthogonal)
    else:
        raise NotImplementedError('initialization method [{:s}] not implemented'.format(init_type))


def define_diffusion(opt):
    model_opt = opt['model']
    arch = model_opt['architecture']
    if arch == 'resdiff':
        from .resdiff import unet
        from .resdiff.resdiff_diffusion import ResDiffDiffusion as Diffusion
    elif arch == 'srdiff':
        from .srdiff import unet
        from .srdiff.srdiff_diffusion import SRDiffDiffusion as Diffusion
    elif arch == 'sr3':
        from .sr3 import unet
        from .sr3.sr3_diffusion import SR3Diffusion as Diffusion
    elif arch == 'phydiff':
        from .phydiff import unet
        from .phydiff.phydiff_diffusion import PhyDiffDiffusion as Diffusion
    elif arch == 'physrdiff':
        # the reference's own physrdiff UNet raises AttributeError on its first forward (physrdiff/unet.py:150)
        raise NotImplementedError('Architecture [{:s}] is broken in the reference itself and outside the accelerated hot path (SURVEY.md 8f).'.format(arch))
    else:
        raise NotImplementedError('Architecture [{:s}] is not implemented.'.format(arch))

    u = model_opt['unet']
    if u.get('norm_groups') is None:
        u['norm_groups'] = 32
    d = model_opt['diffusion']
    model = unet.UNet(in_channel=u['in_channel'], out_channel=u['out_channel'], norm_groups=u['norm_groups'],
                      inner_channel=u['inner_channel'], channel_mults=u['channel_multiplier'], attn_res=u['attn_res'],
                      res_blocks=u['res_blocks'], dropout=u['dropout'], image_height=d['image_height'],
                      image_width=d['image_width'], image_channels=d['image_channels'],
                      precision=model_opt.get('precision', 'bf16'))
    diffusion_model = Diffusion(model, image_height=d['image_height'], image_width=d['image_width'],
                                channels=d['image_channels'], loss_type='l1', conditional=d['conditional'],
                                schedule_opt=model_opt['beta_schedule']['train'],
                                pretrained_model_path=model_opt['pretrained_model']['model_path'],
                                lock_weights=model_opt['pretrained_model']['lock_weights'])
    if opt['phase'] == 'train':
        init_weights(diffusion_model, init_type='orthogonal')
    # The reference wraps the model in nn.DataParallel when several GPU ids are given (:166-168).  Here multi-GPU runs
    # are one process per GPU (torch.distributed / NCCL): the launcher shards the batch, see parallel.py.
    return diffusion_model
