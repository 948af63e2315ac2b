"""ORACLE helper: deterministic, name-independent parameter fill.

There is no checkpoint in this environment (models/pose_higher_hrnet_w48_640.pth.tar is
absent), so parity runs use synthetic weights.  Default PyTorch init depends on module
construction order; this fill depends only on the ORDER AND SHAPES of ``state_dict()``
(identical between the reference module and the drop-in), so the golden fixtures made from
the reference can be reproduced without shipping 255 MB of weights.
"""
from __future__ import annotations

import torch


@torch.no_grad()
def fill_params_deterministic(module: torch.nn.Module, seed: int = 0):
    sd = module.state_dict()
    for idx, (name, t) in enumerate(sd.items()):
        if not t.dtype.is_floating_point:
            continue
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        leaf = name.rsplit(".", 1)[-1]
        if t.dim() == 4:                                     # conv / deconv weight
            fan_in = t.shape[1] * t.shape[2] * t.shape[3]
            v = torch.randn(t.shape, generator=g) * (0.3 / fan_in) ** 0.5
        elif t.dim() == 2:                                   # Linear weight (SE layers)
            v = torch.randn(t.shape, generator=g) * (1.0 / t.shape[1]) ** 0.5
        elif leaf == "running_var":
            v = torch.rand(t.shape, generator=g) * 0.5 + 0.75
        elif leaf == "running_mean":
            v = torch.randn(t.shape, generator=g) * 0.1
        elif leaf == "weight":                               # BN gamma
            v = torch.rand(t.shape, generator=g) * 0.5 + 0.75
        else:                                                # biases
            v = torch.randn(t.shape, generator=g) * 0.1
        t.copy_(v.to(t.dtype))
    return module
