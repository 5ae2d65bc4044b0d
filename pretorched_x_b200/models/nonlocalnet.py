"""Non-local 3-D ResNets (reference: pretorched/models/nonlocalnet.py).

``NonLocalBlock{1,2,3}D`` hold the theta / phi / g / W(+BN) parameters of the non-local block
(nonlocalnet.py:51-131); the body (nonlocalnet.py:143-211) runs in ``engine.run_nonlocal``: one GEMM for
theta|phi|g, the fused QK^T-softmax-V kernel (V consumed in place as an MN-major operand), and the W projection
with BN and the ``+x`` residual in its epilogue.  The positions x positions matrix is never written to memory.

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

__all__ = ['NonLocalBlock1D', 'NonLocalBlock2D', 'NonLocalBlock3D', 'NonLocalResNet3D', 'nonlocalresnet3d', 'nonlocalresnet3d18', 'nonlocalresnet3d34',
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


class _NonLocalBlockND(EngineModule):
    """Non-local block over 1, 2 or 3 position axes (nonlocalnet.py:51-131); parameters only.

    Modes on the engine: ``embedded_gaussian`` (:143-166), ``gaussian`` (:168-190), ``dot_product`` (:192-211), each
    with or without ``sub_sample`` (max-pooled phi / g, :126-131) and with or without the output BatchNorm.
    ``concatenation`` (:213-243) runs on the same fused kernel through a rank-2 encoding of a_i + b_j (engine.run_nonlocal)."""

    def __init__(self, in_channels, inter_channels=None, dimension=3, mode='embedded_gaussian', sub_sample=False,
                 bn_layer=True):
        super().__init__()
        assert dimension in [1, 2, 3]
        assert mode in ['embedded_gaussian', 'gaussian', 'dot_product', 'concatenation']
        self.mode, self.dimension, self.sub_sample = mode, dimension, sub_sample
        self.in_channels = in_channels
        self.inter_channels = inter_channels if inter_channels is not None else max(in_channels // 2, 1)
        conv_nd = (None, nn.Conv1d, nn.Conv2d, nn.Conv3d)[dimension]
        max_pool = (None, nn.MaxPool1d, nn.MaxPool2d, nn.MaxPool3d)[dimension]
        bn = (None, nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)[dimension]
        d = self.inter_channels
        self.g = conv_nd(in_channels, d, kernel_size=1, stride=1, padding=0)
        if bn_layer:
            self.W = nn.Sequential(conv_nd(d, in_channels, kernel_size=1, stride=1, padding=0), bn(in_channels))
            nn.init.constant_(self.W[1].weight, 0)
            nn.init.constant_(self.W[1].bias, 0)
        else:
            self.W = conv_nd(d, in_channels, kernel_size=1, stride=1, padding=0)
            nn.init.constant_(self.W.weight, 0)
            nn.init.constant_(self.W.bias, 0)
        self.theta = None
        self.phi = None
        self.concat_project = None
        if mode in ('embedded_gaussian', 'dot_product', 'concatenation'):
            self.theta = conv_nd(in_channels, d, kernel_size=1, stride=1, padding=0)
            self.phi = conv_nd(in_channels, d, kernel_size=1, stride=1, padding=0)
            if mode == 'concatenation':        # nonlocalnet.py:117-121
                self.concat_project = nn.Sequential(nn.Conv2d(d * 2, 1, 1, 1, 0, bias=False), nn.ReLU())
        if sub_sample:
            self.g = nn.Sequential(self.g, max_pool(kernel_size=2))
            self.phi = max_pool(kernel_size=2) if self.phi is None else nn.Sequential(self.phi, max_pool(kernel_size=2))

    def _run(self, a):
        return engine.run_nonlocal(self, a)

    def forward(self, x):
        from .. import ops
        if isinstance(x, ops.Act):
            return self._run(x)
        shape = x.shape
        x5 = x.reshape(shape[0], shape[1], *([1] * (3 - self.dimension)), *shape[2:])      # -> [N, C, T, H, W]
        out = ops.to_ncdhw(self._run(ops.from_ncdhw(x5, pitch=ops._round_up(shape[1], 8))))
        return out.reshape(shape)


class NonLocalBlock1D(_NonLocalBlockND):
    def __init__(self, in_channels, inter_channels=None, mode='embedded_gaussian', sub_sample=False, bn_layer=True):
        super().__init__(in_channels, inter_channels=inter_channels, dimension=1, mode=mode, sub_sample=sub_sample,
                         bn_layer=bn_layer)


class NonLocalBlock2D(_NonLocalBlockND):
    def __init__(self, in_channels, inter_channels=None, mode='embedded_gaussian', sub_sample=False, bn_layer=True):
        super().__init__(in_channels, inter_channels=inter_channels, dimension=2, mode=mode, sub_sample=sub_sample,
                         bn_layer=bn_layer)


class NonLocalBlock3D(_NonLocalBlockND):
    def __init__(self, in_channels, inter_channels=None, mode='embedded_gaussian', sub_sample=False, bn_layer=True):
        super().__init__(in_channels, inter_channels=inter_channels, dimension=3, mode=mode, sub_sample=sub_sample,
                         bn_layer=bn_layer)


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


class NonLocalResNet3D(engine.CacheOwner, nn.Module):
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
