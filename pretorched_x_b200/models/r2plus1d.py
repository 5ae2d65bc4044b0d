"""R(2+1)D networks (reference: pretorched/models/r2plus1d.py).

Every "convolution" of the 3-D ResNet is replaced by :class:`SpatioTemporalConv` -- a (1,k,k) conv to
``Mi`` intermediate channels, BatchNorm, ReLU, then a (k,1,1) conv (r2plus1d.py:29-88).  The containers
below reproduce the reference's parameter names (``*.spatial_conv.weight``, ``*.bn.*``,
``*.temporal_conv.weight``) and init order; the engine runs each half as one implicit-GEMM launch with
the surrounding BN / residual / ReLU folded into the epilogues (engine.conv_bn_act).

Deviation, on purpose: in the reference R2Plus1D.forward breaks as soon as any
``resnet3d*`` factory has been called in the process, because ``modify_resnets`` patches the shared base
class.  Here R2Plus1D always works; it keeps ``fc`` as the head's state_dict name and additionally
offers ``features / logits / last_linear`` (alias of ``fc``).
"""
import math

import torch.nn as nn
from torch.nn.modules.utils import _triple

from . import resnet3d
from .resnet3d import EngineModule, ResNet3D
from .. import engine

__all__ = [
    'R2Plus1D', 'SpatioTemporalConv', 'r2plus1d10', 'r2plus1d18', 'r2plus1d34', 'r2plus1d50', 'r2plus1d101',
    'r2plus1d152', 'r2plus1d200',
]


def intermediate_channels(in_channels, out_channels, kernel_size):
    """Mi = floor(t*d*d*Ni*No / (d*d*Ni + t*No))  (r2plus1d.py:68-69; paper section 3.5)."""
    kt, kh, kw = kernel_size
    return int(math.floor((kt * kh * kw * in_channels * out_channels) /
                          (kh * kw * in_channels + kt * out_channels)))


class SpatioTemporalConv(EngineModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        kt, kh, kw = _triple(kernel_size)
        st, sh, sw = _triple(stride)
        pt, ph, pw = _triple(padding)
        mid = intermediate_channels(in_channels, out_channels, (kt, kh, kw))
        self.spatial_conv = nn.Conv3d(in_channels, mid, [1, kh, kw], stride=[1, sh, sw], padding=[0, ph, pw], bias=bias)
        self.bn = nn.BatchNorm3d(mid)
        self.relu = nn.ReLU()
        self.temporal_conv = nn.Conv3d(mid, out_channels, [kt, 1, 1], stride=[st, 1, 1], padding=[pt, 0, 0], bias=bias)

    def _run(self, a):
        return engine.conv_bn_act(self, None, a, relu=False)


def conv3x3x3(in_planes, out_planes, stride=1):
    return SpatioTemporalConv(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(resnet3d.BasicBlock):
    Conv3d = staticmethod(conv3x3x3)


class Bottleneck(resnet3d.Bottleneck):
    Conv3d = SpatioTemporalConv


class R2Plus1D(ResNet3D):
    Conv3d = SpatioTemporalConv
    head_name = 'fc'

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, SpatioTemporalConv):
                nn.init.kaiming_normal_(m.spatial_conv.weight, mode='fan_out')
                nn.init.kaiming_normal_(m.temporal_conv.weight, mode='fan_out')
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    @property
    def last_linear(self):
        return self.fc

    def __setattr__(self, name, value):
        # nn.Module.__setattr__ would register a second module called `last_linear`; keep one head: `fc`
        super().__setattr__('fc' if name == 'last_linear' else name, value)


def r2plus1d10(**kwargs):
    return R2Plus1D(BasicBlock, [1, 1, 1, 1], **kwargs)


def r2plus1d18(**kwargs):
    return R2Plus1D(BasicBlock, [2, 2, 2, 2], **kwargs)


def r2plus1d34(**kwargs):
    return R2Plus1D(BasicBlock, [3, 4, 6, 3], **kwargs)


def r2plus1d50(**kwargs):
    return R2Plus1D(Bottleneck, [3, 4, 6, 3], **kwargs)


def r2plus1d101(**kwargs):
    return R2Plus1D(Bottleneck, [3, 4, 23, 3], **kwargs)


def r2plus1d152(**kwargs):
    return R2Plus1D(Bottleneck, [3, 8, 36, 3], **kwargs)


def r2plus1d200(**kwargs):
    return R2Plus1D(Bottleneck, [3, 24, 36, 3], **kwargs)
