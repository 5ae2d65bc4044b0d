"""Non-local 3-D ResNets (reference: pretorched/models/nonlocalnet.py).

``NonLocalBlock3D`` holds the theta / phi / g / W(+BN) parameters of the embedded-gaussian block
(nonlocalnet.py:51-131); its body (nonlocalnet.py:143-166) runs in ``engine.run_nonlocal``: one GEMM for
theta|phi, one swap-AB GEMM for g^T, the fused QK^T-softmax-V kernel, and the W conv with BN and the
``+x`` residual in its epilogue.  The THW x THW matrix is never written to memory.

Reference behaviours reproduced on purpose:
  * ``nonlocalresnet3d50(num_classes=...)`` does not forward ``num_classes`` (nonlocalnet.py:553-561):
    the network always has 339 outputs unless ``num_classes`` arrives through ``**kwargs`` of the class.
  * default ``shortcut_type='A'`` even for bottlenecks -> parameter-free shortcuts, no ``downsample.*`` keys.
  * ``init_weights`` re-initialises *all* Conv3d (theta/phi/g/W too) and sets every BN weight to 1, so the
    non-local branch is not an identity at random init (nonlocalnet.py:448-454).
"""
from collections import defaultdict

import torch.nn as nn

from .. import engine
from .resnet3d import EngineModule, ShortcutA, _attach_settings, conv3x3x3

__all__ = ['NonLocalBlock3D', 'NonLocalResNet3D', 'nonlocalresnet3d', 'nonlocalresnet3d18', 'nonlocalresnet3d34',
           'nonlocalresnet3d50', 'nonlocalresnet3d101', 'nonlocalresnet3d152', 'nonlocalresnet3d200']

_URL = 'http://pretorched-x.csail.mit.edu/models/resnet3d50_kinetics-aad059c9.pth'
pretrained_settings = defaultdict(dict)
for _dataset, _n in (('kinetics-400', 400), ('moments', 339)):
    pretrained_settings['nonlocalresnet3d50'][_dataset] = {
        'input_space': 'RGB', 'input_range': [0, 1],
        'url': _URL if _dataset == 'kinetics-400' else None,
        'std': [0.229, 0.224, 0.225], 'mean': [0.485, 0.456, 0.406],
        'num_classes': _n, 'input_size': [3, 224, 224],
    }


class NonLocalBlock3D(EngineModule):
    """Embedded-gaussian non-local block over (T,H,W); parameters only."""

    def __init__(self, in_channels, inter_channels=None, mode='embedded_gaussian', sub_sample=False, bn_layer=True):
        super().__init__()
        if mode != 'embedded_gaussian' or sub_sample:
            raise NotImplementedError("engine scope: embedded_gaussian mode without sub-sampling "
                                      "(the configuration every reference 3-D net uses, nonlocalnet.py:395)")
        self.mode, self.dimension, self.sub_sample = mode, 3, sub_sample
        self.in_channels = in_channels
        self.inter_channels = inter_channels if inter_channels is not None else max(in_channels // 2, 1)
        d = self.inter_channels
        self.g = nn.Conv3d(in_channels, d, kernel_size=1, stride=1, padding=0)
        if bn_layer:
            self.W = nn.Sequential(nn.Conv3d(d, in_channels, kernel_size=1, stride=1, padding=0),
                                   nn.BatchNorm3d(in_channels))
            nn.init.constant_(self.W[1].weight, 0)
            nn.init.constant_(self.W[1].bias, 0)
        else:
            self.W = nn.Conv3d(d, in_channels, kernel_size=1, stride=1, padding=0)
            nn.init.constant_(self.W.weight, 0)
            nn.init.constant_(self.W.bias, 0)
        self.theta = nn.Conv3d(in_channels, d, kernel_size=1, stride=1, padding=0)
        self.phi = nn.Conv3d(in_channels, d, kernel_size=1, stride=1, padding=0)

    def _run(self, a):
        return engine.run_nonlocal(self, a)


class _NLBlockBase(EngineModule):
    def _run(self, a):
        return engine.run_block(self, a)


class NonLocalBasicBlock(_NLBlockBase):
    expansion = 1
    Conv3d = staticmethod(conv3x3x3)

    def __init__(self, inplanes, planes, stride=1, downsample=None, nonlocal_layer=False):
        super().__init__()
        self.conv1, self.bn1 = self.Conv3d(inplanes, planes, stride), nn.BatchNorm3d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2, self.bn2 = self.Conv3d(planes, planes), nn.BatchNorm3d(planes)
        self.stride, self.downsample, self.nonlocal_layer = stride, downsample, nonlocal_layer
        if nonlocal_layer:
            self.nonlocalblock = NonLocalBlock3D(planes)


class NonLocalBottleneck(_NLBlockBase):
    expansion = 4
    Conv3d = nn.Conv3d

    def __init__(self, inplanes, planes, stride=1, downsample=None, nonlocal_layer=False):
        super().__init__()
        mk = self.Conv3d
        self.conv1, self.bn1 = mk(inplanes, planes, kernel_size=1, bias=False), nn.BatchNorm3d(planes)
        self.conv2 = mk(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm3d(planes)
        self.conv3, self.bn3 = mk(planes, planes * 4, kernel_size=1, bias=False), nn.BatchNorm3d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.stride, self.downsample, self.nonlocal_layer = stride, downsample, nonlocal_layer
        if nonlocal_layer:
            self.nonlocalblock = NonLocalBlock3D(planes * 4)


class NonLocalResNet3D(nn.Module):
    Conv3d = nn.Conv3d
    head_name = 'last_linear'

    def __init__(self, block, layers, nonlocal_layers, shortcut_type='A', num_classes=339):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv3d(3, 64, kernel_size=7, stride=(1, 2, 2), padding=(3, 3, 3), bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool3d(kernel_size=(3, 3, 3), stride=2, padding=1)
        for i, (planes, nblocks, n_nl) in enumerate(zip((64, 128, 256, 512), layers, nonlocal_layers)):
            setattr(self, 'layer%d' % (i + 1),
                    self._make_layer(block, planes, nblocks, n_nl, shortcut_type, stride=1 if i == 0 else 2))
        self.avgpool = nn.AdaptiveAvgPool3d(1)
        self.last_linear = nn.Linear(512 * block.expansion, num_classes)
        self.init_weights()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, self.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out')
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, nonlocal_blocks, shortcut_type, stride=1):
        out_planes = planes * block.expansion
        downsample = None
        if stride != 1 or self.inplanes != out_planes:
            if shortcut_type == 'A':
                downsample = ShortcutA(out_planes, stride)
            else:
                downsample = nn.Sequential(
                    self.Conv3d(self.inplanes, out_planes, kernel_size=1, stride=stride, bias=False),
                    nn.BatchNorm3d(out_planes))
        # a non-local block follows every `freq`-th residual block (nonlocalnet.py:474-479)
        freq = blocks // nonlocal_blocks if nonlocal_blocks != 0 else -1
        seq = []
        for i in range(blocks):
            seq.append(block(self.inplanes, planes, stride=stride, downsample=downsample,
                             nonlocal_layer=(freq > 0 and i % freq == 0)))
            if i == 0:
                stride, downsample, self.inplanes = 1, None, out_planes
        return nn.Sequential(*seq)

    def features_act(self, x):
        if self.training:
            raise RuntimeError("the forward engine is inference-only: call model.eval() first")
        return engine.run_trunk(self, x)

    def features(self, x):
        from .. import ops
        return ops.to_ncdhw(self.features_act(x))

    def logits(self, features):
        from .. import ops
        a = features if isinstance(features, ops.Act) else ops.from_ncdhw(features, pitch=ops._round_up(features.shape[1], 8))
        return engine.run_head(self, a, self.last_linear)

    def forward(self, input):
        return self.logits(self.features_act(input))


def nonlocalresnet3d(**kwargs):
    return NonLocalResNet3D(NonLocalBasicBlock, [1, 1, 1, 1], **kwargs)


def nonlocalresnet3d18(**kwargs):
    return NonLocalResNet3D(NonLocalBasicBlock, [2, 2, 2, 2], **kwargs)


def nonlocalresnet3d34(**kwargs):
    return NonLocalResNet3D(NonLocalBasicBlock, [3, 4, 6, 3], **kwargs)


def nonlocalresnet3d50(num_classes=339, num_nonlocal_blocks=5, pretrained='kinetics-400', **kwargs):
    if num_nonlocal_blocks == 5:
        nonlocal_blocks = [0, 2, 3, 0]
    elif num_nonlocal_blocks == 10:
        nonlocal_blocks = [0, 4, 6, 0]
    else:
        raise ValueError("num_nonlocal_blocks must be 5 or 10 (nonlocalnet.py:556-559)")
    # num_classes is intentionally NOT forwarded: reference quirk, see module docstring
    model = NonLocalResNet3D(NonLocalBottleneck, [3, 4, 6, 3], nonlocal_blocks, **kwargs)
    if pretrained is not None:
        import torch.utils.model_zoo as model_zoo
        settings = pretrained_settings['nonlocalresnet3d50'][pretrained]
        model.load_state_dict(model_zoo.load_url(settings['url']), strict=False)
        _attach_settings(model, settings)
    return model


def nonlocalresnet3d101(**kwargs):
    return NonLocalResNet3D(NonLocalBottleneck, [3, 4, 23, 3], **kwargs)


def nonlocalresnet3d152(**kwargs):
    return NonLocalResNet3D(NonLocalBottleneck, [3, 8, 36, 3], **kwargs)


def nonlocalresnet3d200(**kwargs):
    return NonLocalResNet3D(NonLocalBottleneck, [3, 24, 36, 3], **kwargs)
