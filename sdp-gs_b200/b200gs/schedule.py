"""Optimisation defaults and the position learning-rate schedule of the reference (arguments/__init__.py:75-101,
utils/general_utils.py:get_expon_lr_func).  Pure Python on purpose: bench.py's reference arm and the torch oracle
import these without mapping libb200gs.so."""
from __future__ import annotations

import math

DEFAULTS = dict(  # arguments/__init__.py:75-101
    position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=30_000,
    feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001, language_feature_lr=0.013,
    lambda_dssim=0.2, depth_weight=0.05, spatial_lr_scale=1.0, beta1=0.9, beta2=0.999, eps=1e-15)


def expon_lr(step, lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """utils/general_utils.py:get_expon_lr_func"""
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0), 1))
    else:
        delay_rate = 1.0
    t = min(max(step / max_steps, 0), 1)
    return delay_rate * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)
