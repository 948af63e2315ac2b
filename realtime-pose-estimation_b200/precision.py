"""Mirror of the inference part of rtpe/third_party/fp16_utils/fp16util.py:40-91 and of
``get_hrnet_w48_teacher`` (rtpe/helpers.py:32-73).

``network_to_half(net)`` returns ``Sequential(tofp16(), net.half() with float32 BatchNorm,
tofp32())`` exactly like the reference, which is why teacher checkpoints carry the
``"1."`` key prefix.  With this package's ``PoseHigherResolutionNet`` inside, half
parameters select the bf16 tcgen05 path (fp32 accumulation, fp32 folded-BN epilogue).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .hhrnet import PoseHigherResolutionNet


class tofp16(nn.Module):
    def forward(self, input):
        return input.half()


class tofp32(nn.Module):
    def forward(self, input):
        if isinstance(input, list):
            return [x.float() for x in input]
        return input.float()


def BN_convert_float(module):
    if isinstance(module, torch.nn.modules.batchnorm._BatchNorm):
        module.float()
    for child in module.children():
        BN_convert_float(child)
    return module


def network_to_half(network):
    return nn.Sequential(tofp16(), BN_convert_float(network.half()), tofp32())


W48_KWARGS = dict(
    num_joints=17, tag_per_joint=True, final_conv_ksize=1, pretrained_layers=["*"], inplanes=64,
    s2_modules=1, s2_branches=2, s2_block_type="BASIC", s2_blocks=[4, 4], s2_chans=[48, 96],
    s3_modules=4, s3_branches=3, s3_block_type="BASIC", s3_blocks=[4, 4, 4],
    s3_chans=[48, 96, 192],
    s4_modules=3, s4_branches=4, s4_block_type="BASIC", s4_blocks=[4, 4, 4, 4],
    s4_chans=[48, 96, 192, 384],
    deconvs=1, deconv_chans=[48], deconv_ksize=[4], deconv_num_blocks=4, deconv_cat=[True],
    with_ae_loss=(True, False))


def get_hrnet_w48_teacher(w48_statedict_path=None, half=True):
    """rtpe/helpers.py:32-73.  ``w48_statedict_path=None`` keeps the random init (there is
    no checkpoint in this environment); otherwise the state dict is loaded strictly."""
    model = PoseHigherResolutionNet(**{k: (list(v) if isinstance(v, list) else v)
                                       for k, v in W48_KWARGS.items()})
    if half:
        model = network_to_half(model)
    if w48_statedict_path is not None:
        model.load_state_dict(torch.load(w48_statedict_path), strict=True)
    model.eval()
    return model
