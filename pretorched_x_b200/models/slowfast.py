"""SlowFast / SlowOnly / FastOnly networks (reference: pretorched/models/slowfast.py).

Two ResNet pathways over the same clip: ``fast`` sees every ``fast_stride``-th frame with 1/8 of the channels,
``slow`` sees every ``slow_stride``-th frame; after the stem pool and after res2/res3/res4 the fast pathway is
projected by a (5,1,1) stride-(8,1,1) convolution and concatenated to the slow pathway's channels
(slowfast.py:136-153, 273-315); the head concatenates both pooled feature vectors (slowfast.py:390-396).

As elsewhere in this package the classes below are parameter containers that reproduce the reference's tensors,
names and construction order (the reference uses PyTorch's default initialisation here, so a seeded build is
bit-identical); the forward bodies run on the engine: the (1,7,7)/(5,7,7) stems on the Toeplitz stem kernel, the
(1,3,3)/(3,1,1) convolutions on the slab kernel, 1x1x1 convolutions on the persistent GEMM, the stride-8 lateral
projections on the gather kernel, channel concatenation on ``b2_concat_channels``.  ``SlowFastV0`` (an older
variant of the same network kept upstream for reference) is not provided.
"""
import torch.nn as nn

from .. import engine, ops
from ..ops import Act
from .resnet3d import EngineModule

__all__ = ['SlowFast', 'SlowOnly', 'FastOnly', 'resnet18', 'resnet50', 'resnet101', 'resnet152', 'resnet200']


class BasicBlock(EngineModule):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, head_conv=1):
        super().__init__()
        if head_conv == 1:
            self.conv1 = nn.Conv3d(inplanes, planes, kernel_size=(1, 3, 3), padding=(0, 1, 1),
                                   stride=(1, stride, stride), bias=False)
        elif head_conv == 3:
            self.conv1 = nn.Conv3d(inplanes, planes, kernel_size=(3, 1, 1), padding=(1, 0, 0), bias=False)
        else:
            raise ValueError('Unsupported head_conv')
        self.bn1 = nn.BatchNorm3d(planes)
        self.relu = nn.ReLU(inplace=True)
        # the reference leaves bias=True on this one (slowfast.py:31-34)
        self.conv2 = nn.Conv3d(planes, planes, kernel_size=(1, 3, 3), padding=(0, 1, 1), stride=(1, stride, stride))
        self.bn2 = nn.BatchNorm3d(planes)
        self.downsample = downsample
        self.stride = stride

    def _run(self, a):
        return engine.run_basic(self, a)


class Bottleneck(EngineModule):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, head_conv=1):
        super().__init__()
        if head_conv == 1:
            self.conv1 = nn.Conv3d(inplanes, planes, kernel_size=1, bias=False)
        elif head_conv == 3:
            self.conv1 = nn.Conv3d(inplanes, planes, kernel_size=(3, 1, 1), bias=False, padding=(1, 0, 0))
        else:
            raise ValueError("Unsupported head_conv!")
        self.bn1 = nn.BatchNorm3d(planes)
        self.conv2 = nn.Conv3d(planes, planes, kernel_size=(1, 3, 3), stride=(1, stride, stride), padding=(0, 1, 1),
                               bias=False)
        self.bn2 = nn.BatchNorm3d(planes)
        self.conv3 = nn.Conv3d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm3d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def _run(self, a):
        return engine.run_bottleneck(self, a)


def _res_layer(block, inplanes, planes, blocks, stride, head_conv):
    """One residual stage; the (1,s,s)-strided 1x1x1 projection is built before the blocks, as upstream."""
    downsample = None
    if stride != 1 or inplanes != planes * block.expansion:
        downsample = nn.Sequential(
            nn.Conv3d(inplanes, planes * block.expansion, kernel_size=1, stride=(1, stride, stride), bias=False),
            nn.BatchNorm3d(planes * block.expansion))
    seq = [block(inplanes, planes, stride, downsample, head_conv=head_conv)]
    seq += [block(planes * block.expansion, planes, head_conv=head_conv) for _ in range(1, blocks)]
    return nn.Sequential(*seq)


def _stem_and_pool(m, a):
    a = engine.conv_bn_act(m.conv1, m.bn1, a, relu=True)
    k, s, p = engine._pool_args(m.maxpool)
    return ops.maxpool3d(a, k, s, p)


def _subsample_frames(x, stride):
    if isinstance(x, Act):
        raise TypeError("SlowFast networks take the NCDHW clip tensor (frame sub-sampling happens on it)")
    return x[:, :, ::stride]


class Slow(engine.CacheOwner, nn.Module):
    """Slow pathway with lateral inputs (slowfast.py:104-180)."""
    lateral_factor = 2            # channels gained from the fast pathway: out // 8 * 2

    def __init__(self, block=Bottleneck, layers=(2, 2, 2, 2)):
        super().__init__()
        self.inplanes = 64 + (64 // 8 * 2 if self.lateral_factor else 0)
        self._make_layers(block, layers)

    def _widen(self, planes_out):
        return planes_out + (planes_out // 8 * 2 if self.lateral_factor else 0)

    def _make_layers(self, block, layers):
        self.conv1 = nn.Conv3d(3, 64, kernel_size=(1, 7, 7), stride=(1, 2, 2), padding=(0, 3, 3), bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        res3_stride = 2 if issubclass(block, Bottleneck) else 1
        for name, planes, n, stride, head in (('res2', 64, layers[0], 1, 1), ('res3', 128, layers[1], res3_stride, 1),
                                              ('res4', 256, layers[2], 2, 3), ('res5', 512, layers[3], 2, 3)):
            setattr(self, name, _res_layer(block, self.inplanes, planes, n, stride, head))
            self.inplanes = self._widen(planes * block.expansion)

    def run(self, a, lateral):
        """a: NDHWC4 clip Act (already frame-sub-sampled); lateral: 4 Acts from the fast pathway (or None)."""
        a = _stem_and_pool(self, a)
        for i, name in enumerate(('res2', 'res3', 'res4', 'res5')):
            if lateral is not None:
                a = ops.concat_channels(a, lateral[i])
            for blk in getattr(self, name):
                a = engine.run_block(blk, a)
        return a


class SlowOnly(Slow):
    lateral_factor = 0

    def __init__(self, block=Bottleneck, layers=(2, 2, 2, 2), num_classes=400, dropout=0.5, slow_stride=16):
        nn.Module.__init__(self)
        self.inplanes = 64
        self.slow_stride = slow_stride
        self._make_layers(block, layers)
        self.dropout = nn.Dropout(dropout)
        self.last_linear = nn.Linear(self.inplanes, num_classes)

    def input_transform(self, input):
        return _subsample_frames(input, self.slow_stride)

    def forward(self, input):
        _require_eval(self)
        a = self.run(ops.from_ncdhw(self.input_transform(input)), None)
        return engine.run_head(self, a, self.last_linear)


class Fast(engine.CacheOwner, nn.Module):
    """Fast pathway + the lateral projections (slowfast.py:244-345)."""

    def __init__(self, block=Bottleneck, layers=(2, 2, 2, 2)):
        super().__init__()
        self.inplanes = 8
        self.conv1 = nn.Conv3d(3, 8, kernel_size=(5, 7, 7), stride=(1, 2, 2), padding=(2, 3, 3), bias=False)
        self.bn1 = nn.BatchNorm3d(8)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        res3_stride = 2 if issubclass(block, Bottleneck) else 1
        for name, planes, n, stride in (('res2', 8, layers[0], 1), ('res3', 16, layers[1], res3_stride),
                                        ('res4', 32, layers[2], 2), ('res5', 64, layers[3], 2)):
            setattr(self, name, _res_layer(block, self.inplanes, planes, n, stride, 3))
            self.inplanes = planes * block.expansion
        self._make_lateral_layers(4 if issubclass(block, Bottleneck) else 1)

    def _make_lateral_layers(self, expansion):
        for name, c in (('lateral_p1', 8), ('lateral_res2', 8 * expansion), ('lateral_res3', 16 * expansion),
                        ('lateral_res4', 32 * expansion)):
            setattr(self, name, nn.Conv3d(c, c * 2, kernel_size=(5, 1, 1), stride=(8, 1, 1), bias=False,
                                          padding=(2, 0, 0)))

    def run(self, a, with_lateral=True):
        lateral = []
        a = _stem_and_pool(self, a)
        if with_lateral:
            lateral.append(engine.conv_bn_act(self.lateral_p1, None, a))
        for name in ('res2', 'res3', 'res4', 'res5'):
            for blk in getattr(self, name):
                a = engine.run_block(blk, a)
            if with_lateral and name != 'res5':
                lateral.append(engine.conv_bn_act(getattr(self, 'lateral_' + name), None, a))
        return a, lateral


class FastOnly(Fast):
    def __init__(self, block=Bottleneck, layers=(2, 2, 2, 2), num_classes=400, dropout=0.5, fast_stride=2):
        super().__init__(block=block, layers=layers)
        self.fast_stride = fast_stride
        self.dropout = nn.Dropout(dropout)
        self.last_linear = nn.Linear(self.inplanes, num_classes)

    def _make_lateral_layers(self, expansion):
        return None

    def input_transform(self, input):
        return _subsample_frames(input, self.fast_stride)

    def forward(self, input):
        _require_eval(self)
        a, _ = self.run(ops.from_ncdhw(self.input_transform(input)), with_lateral=False)
        return engine.run_head(self, a, self.last_linear)


class SlowFast(engine.CacheOwner, nn.Module):
    def __init__(self, block=Bottleneck, layers=(2, 2, 2, 2), num_classes=400, dropout=0.5, slow_stride=16,
                 fast_stride=2):
        super().__init__()
        self.slow_stride, self.fast_stride = slow_stride, fast_stride
        self.expansion = 4 if issubclass(block, Bottleneck) else 1
        self.slow = Slow(block=block, layers=layers)
        self.fast = Fast(block=block, layers=layers)
        self.dropout = nn.Dropout(dropout)
        self.last_linear = nn.Linear(self.fast.inplanes + 512 * self.expansion, num_classes, bias=False)

    def forward(self, input):
        _require_eval(self)
        fast, lateral = self.fast.run(ops.from_ncdhw(_subsample_frames(input, self.fast_stride)))
        slow = self.slow.run(ops.from_ncdhw(_subsample_frames(input, self.slow_stride)), lateral)
        ps, pf = ops.avgpool_global(slow), ops.avgpool_global(fast)            # fp16 [B][C]
        feats = ops.concat_rows(ps, slow.C, pf, fast.C)                         # torch.cat([slow, fast], dim=1)
        head = self.last_linear
        if isinstance(head, nn.Linear):
            pl = engine._cached(head, "pl", engine._sig(head.weight, head.bias),
                                lambda: ops.PackedLinear(head.weight, head.bias))
            return ops.linear(feats, pl, out_f32=True)
        return head(feats[:, :slow.C + fast.C].float())


def _require_eval(m):
    if m.training:
        raise RuntimeError("the forward engine is inference-only: call model.eval() first")


def _by_mode(mode):
    return {'sf': SlowFast, 'f': FastOnly, 's': SlowOnly}.get(mode.lower(), SlowFast)


def resnet18(mode='SF', **kwargs):
    return _by_mode(mode)(BasicBlock, [2, 2, 2, 2], **kwargs)


def resnet50(mode='SF', **kwargs):
    return _by_mode(mode)(Bottleneck, [3, 4, 6, 3], **kwargs)


def resnet101(**kwargs):
    return SlowFast(Bottleneck, [3, 4, 23, 3], **kwargs)


def resnet152(**kwargs):
    return SlowFast(Bottleneck, [3, 8, 36, 3], **kwargs)


def resnet200(**kwargs):
    return SlowFast(Bottleneck, [3, 24, 36, 3], **kwargs)
