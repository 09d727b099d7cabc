"""Solver-side helpers mirroring utils.py of the reference (only what the hot path uses)."""
import math
import os

import torch
import torch.nn.init as init
import yaml
from torch.optim import lr_scheduler

from .flat import ema_update


def get_config(path):
    with open(path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f)


def moving_average(model, model_copy, beta=0.999):
    """p_copy = lerp(p, p_copy, beta) for every parameter (utils.py:52-54), one kernel per network."""
    ema_update(model.ensure_flat(), model_copy.ensure_flat(), beta)


def get_scheduler(optimizer, hyperparameters, iterations=-1):
    """utils.py:220-231."""
    policy = hyperparameters.get('lr_policy', 'const')
    if policy == 'const':
        return None
    if iterations != -1:
        for g in optimizer.param_groups:
            g.setdefault('initial_lr', g['lr'])
    if policy == 'step':
        return lr_scheduler.StepLR(optimizer, step_size=hyperparameters['step_size'], gamma=hyperparameters['gamma'],
                                   last_epoch=iterations)
    if policy == 'cosa':
        return lr_scheduler.CosineAnnealingLR(optimizer, T_max=hyperparameters['step_size'],
                                              eta_min=hyperparameters['eta_min'], last_epoch=iterations)
    raise NotImplementedError('learning rate policy [%s] is not implemented' % policy)


def weights_init(init_type='gaussian'):
    """utils.py:234-254: modules whose class name starts with Conv/Linear and that own a .weight."""
    def init_fun(m):
        name = m.__class__.__name__
        if (name.find('Conv') == 0 or name.find('Linear') == 0) and hasattr(m, 'weight') and \
                isinstance(getattr(m, 'weight'), torch.nn.Parameter):
            if init_type == 'gaussian':
                init.normal_(m.weight.data, 0.0, 0.02)
            elif init_type == 'xavier':
                init.xavier_normal_(m.weight.data, gain=math.sqrt(2))
            elif init_type == 'kaiming':
                init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
            elif init_type == 'orthogonal':
                init.orthogonal_(m.weight.data, gain=math.sqrt(2))
            elif init_type != 'default':
                raise AssertionError("Unsupported initialization: {}".format(init_type))
            if hasattr(m, 'bias') and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
    return init_fun


def get_model_list(dirname, key):
    """newest checkpoint of a kind (utils.py:169-178)."""
    if not os.path.exists(dirname):
        return None
    models = sorted(os.path.join(dirname, f) for f in os.listdir(dirname)
                    if os.path.isfile(os.path.join(dirname, f)) and key in f and ".pt" in f)
    return models[-1] if models else None
