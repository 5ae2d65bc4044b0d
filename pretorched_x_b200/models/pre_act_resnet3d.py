"""Pre-activation 3-D ResNets (reference: pretorched/models/pre_act_resnet3D.py).

``PreActivationBasicBlock`` (pre_act_resnet3D.py:27-57) and ``PreActivationBottleneck`` (:60-96) apply BN -> ReLU *before*
each convolution and add the shortcut without a final ReLU.  The containers keep the reference's registration order
(bn1, conv1, bn2, conv2, ...) so a seeded init and the ``state_dict`` are identical; on the engine the BN + ReLU that
precedes conv k+1 rides in the epilogue of conv k, the block's own bn1 + ReLU is one element-wise pass (its input is also
the raw shortcut), and the closing convolution's epilogue adds the residual (``engine._preact_body``).

Like ``r2plus1d.py`` the upstream file does ``import resnet3D`` (absolute, pre_act_resnet3D.py:8) and is not exported by
``pretorched/__init__.py``; the net is a plain ``ResNet3D`` subclass with an ``fc`` head (resnet3D.py:203-218).
"""
import torch.nn as nn

from .resnet3d import EngineModule, ResNet3D
from .. import engine

__all__ = [
    'PreActivationResNet3D', 'preact_resnet3d10', 'preact_resnet3d18', 'preact_resnet3d34',
    'preact_resnet3d50', 'preact_resnet3d101', 'preact_resnet3d152', 'preact_resnet3d200',
]


def conv3x3x3(in_planes, out_planes, stride=1):
    return nn.Conv3d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class _PreActBlock(EngineModule):
    preactivation = True

    def _run(self, a):
        return engine.run_block(self, a)


class PreActivationBasicBlock(_PreActBlock):
    expansion = 1
    Conv3d = staticmethod(conv3x3x3)

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.bn1 = nn.BatchNorm3d(inplanes)
        self.conv1 = self.Conv3d(inplanes, planes, stride)
        self.bn2 = nn.BatchNorm3d(planes)
        self.conv2 = self.Conv3d(planes, planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class PreActivationBottleneck(_PreActBlock):
    expansion = 4
    Conv3d = nn.Conv3d

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.bn1 = nn.BatchNorm3d(inplanes)
        self.conv1 = self.Conv3d(inplanes, planes, kernel_size=1, bias=False)
        self.bn2 = nn.BatchNorm3d(planes)
        self.conv2 = self.Conv3d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn3 = nn.BatchNorm3d(planes)
        self.conv3 = self.Conv3d(planes, planes * 4, kernel_size=1, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class PreActivationResNet3D(ResNet3D):
    head_name = 'fc'

    @property
    def last_linear(self):
        return self.fc

    def __setattr__(self, name, value):
        super().__setattr__('fc' if name == 'last_linear' else name, value)


def preact_resnet3d10(**kwargs):
    return PreActivationResNet3D(PreActivationBasicBlock, [1, 1, 1, 1], **kwargs)


def preact_resnet3d18(**kwargs):
    return PreActivationResNet3D(PreActivationBasicBlock, [2, 2, 2, 2], **kwargs)


def preact_resnet3d34(**kwargs):
    return PreActivationResNet3D(PreActivationBasicBlock, [3, 4, 6, 3], **kwargs)


def preact_resnet3d50(**kwargs):
    return PreActivationResNet3D(PreActivationBottleneck, [3, 4, 6, 3], **kwargs)


def preact_resnet3d101(**kwargs):
    return PreActivationResNet3D(PreActivationBottleneck, [3, 4, 23, 3], **kwargs)


def preact_resnet3d152(**kwargs):
    return PreActivationResNet3D(PreActivationBottleneck, [3, 8, 36, 3], **kwargs)


def preact_resnet3d200(**kwargs):
    return PreActivationResNet3D(PreActivationBottleneck, [3, 24, 36, 3], **kwargs)
