import logging

logger = logging.getLogger('base')


def create_model(opt):
    """reference models/diffusion_models/__init__.py:5-18."""
    from .model import DDPM
    model = DDPM(opt)
    logger.info('Model [{:s}] is created.'.format(model.__class__.__name__))
    return model
