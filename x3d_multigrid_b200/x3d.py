"""Drop-in nn.Module surface of the reference's x3d.py, executed by sm_100a kernels.

Same public names, constructor signatures, attribute names (hence ``state_dict`` keys,
SURVEY.md A4) and methods as KiyoshiKAWASAKI/X3D-Multigrid ``x3d.py``:
``generate_model`` (x3d.py:366), ``ResNet`` (:174), ``Bottleneck`` (:106),
``SubBatchNorm3d`` (:9), ``Swish``/``SwishEfficient`` (:61,:71), ``conv3x3x3`` (:87),
``conv1x1x1`` (:98), ``get_inplanes``/``get_blocks`` (:352,:359).

The modules are parameter containers: ``ResNet.forward`` and ``Bottleneck.forward`` hand the
whole computation to ``engine.Engine`` (hand-written CUDA kernels behind the C ABI of
libx3d_b200.so).  Inputs must be CUDA tensors -- there is no CPU or ATen fallback.

Extra knob (not in the reference): ``compute_dtype`` attribute / ``set_compute_dtype``:
torch.bfloat16 (default: bf16 activations, fp32 accumulation and statistics) or
torch.float32 (parity mode).
"""
from __future__ import annotations

from functools import partial
from typing import List, Optional

import torch
import torch.nn as nn

from . import engine as _engine
from . import ops as _ops


class SubBatchNorm3d(nn.Module):
    """Split batch-norm (x3d.py:9-58): ``num_splits`` independent statistics groups
    (samples b::num_splits) in training, one shared affine, aggregated stats for eval."""

    def __init__(self, num_splits, **args):
        super().__init__()
        self.num_splits = num_splits
        self.num_features = args['num_features']
        if args.get('affine', True):
            self.affine = True
            args['affine'] = False
            self.weight = nn.Parameter(torch.ones(self.num_features))
            self.bias = nn.Parameter(torch.zeros(self.num_features))
        else:
            self.affine = False
        self.bn = nn.BatchNorm3d(**args)
        args['num_features'] = self.num_features * self.num_splits
        self.split_bn = nn.BatchNorm3d(**args)

    def set_split_bn(self, num_splits):
        """Fresh ``split_bn`` for ``num_splits`` groups (x3d.py:301-302 builds a new BatchNorm3d).  One module per
        split count is kept and reset in place when it comes back, so that its buffers keep their addresses and
        CUDA graphs captured for an earlier visit of the same long cycle stay valid."""
        cache = self.__dict__.setdefault('_split_cache', {})
        cache.setdefault(self.split_bn.num_features // self.num_features, self.split_bn)
        bn = cache.get(num_splits)
        if bn is None:
            bn = nn.BatchNorm3d(num_features=self.num_features * num_splits, affine=False).to(self.weight.device)
            cache[num_splits] = bn
        else:
            bn.to(self.weight.device)
            bn.reset_running_stats()
        bn.train(self.training)
        self.split_bn = bn

    def _get_aggregated_mean_std(self, means, stds, n):
        # x3d.py:27-33 ("std" is a variance there too)
        mean = means.view(n, -1).sum(0) / n
        std = stds.view(n, -1).sum(0) / n + ((means.view(n, -1) - mean) ** 2).view(n, -1).sum(0) / n
        return mean.detach(), std.detach()

    def aggregate_stats(self):
        """x3d.py:35-45 -- fold the split running stats into ``bn`` before eval."""
        if self.split_bn.track_running_stats:
            self.bn.running_mean.data, self.bn.running_var.data = self._get_aggregated_mean_std(
                self.split_bn.running_mean, self.split_bn.running_var, self.num_splits)

    def forward(self, x):
        return _ops.sub_batch_norm(self, x)


class SwishEfficient(torch.autograd.Function):
    """x * sigmoid(x) with the hand-written backward of x3d.py:71-84 (CUDA kernels)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return _ops.swish_fwd(x)

    @staticmethod
    def backward(ctx, grad_output):
        (x,) = ctx.saved_tensors
        return _ops.swish_bwd(x, grad_output)


class Swish(nn.Module):
    def forward(self, x):
        return SwishEfficient.apply(x)


class _DepthwiseConv3d(nn.Conv3d):
    def forward(self, x):
        return _ops.depthwise_conv(self, x)


class _PointwiseConv3d(nn.Conv3d):
    def forward(self, x):
        return _ops.pointwise_conv(self, x)


class _StemSpatialConv3d(nn.Conv3d):
    def forward(self, x):
        return _ops.stem_conv(self, x)


def conv3x3x3(in_planes, out_planes, stride=1):
    """channelwise 3x3x3, stride (1,s,s), pad 1 (x3d.py:87-95)"""
    return _DepthwiseConv3d(in_planes, out_planes, kernel_size=3, stride=(1, stride, stride), padding=1,
                            bias=False, groups=in_planes)


def conv1x1x1(in_planes, out_planes, stride=1):
    """pointwise conv, stride (1,s,s) (x3d.py:98-103)"""
    return _PointwiseConv3d(in_planes, out_planes, kernel_size=1, stride=(1, stride, stride), bias=False)


class Bottleneck(nn.Module):
    """x3d.py:106-171."""

    def __init__(self, in_planes, planes, stride=1, downsample=None, index=0, base_bn_splits=8):
        super().__init__()
        self.index = index
        self.base_bn_splits = base_bn_splits
        mid, out = planes[0], planes[1]
        self.conv1 = conv1x1x1(in_planes, mid)
        self.bn1 = SubBatchNorm3d(num_splits=base_bn_splits, num_features=mid, affine=True)
        self.conv2 = conv3x3x3(mid, mid, stride)
        self.bn2 = SubBatchNorm3d(num_splits=base_bn_splits, num_features=mid, affine=True)
        self.conv3 = conv1x1x1(mid, out)
        self.bn3 = SubBatchNorm3d(num_splits=base_bn_splits, num_features=out, affine=True)
        self.swish = Swish()
        self.relu = nn.ReLU(inplace=True)
        if self.index % 2 == 0:
            width = self.round_width(mid)
            self.global_pool = nn.AdaptiveAvgPool3d((1, 1, 1))
            self.fc1 = _PointwiseConv3d(mid, width, kernel_size=1, stride=1)
            self.fc2 = _PointwiseConv3d(width, mid, kernel_size=1, stride=1)
            self.sigmoid = nn.Sigmoid()
        self.downsample = downsample
        self.stride = stride
        # engine metadata (not part of the reference surface)
        self.in_planes, self.mid_planes, self.out_planes = in_planes, mid, out
        self.prefix = 'blk'   # overwritten by ResNet ('layerL.i')
        self.stage = 0
        self.compute_dtype = torch.float32   # stand-alone use; ResNet has its own knob

    @property
    def has_se(self):
        return self.index % 2 == 0

    @property
    def se_width(self):
        return self.fc1.out_channels if self.has_se else 0

    def round_width(self, width, multiplier=0.0625, min_width=8, divisor=8):
        """x3d.py:129-140"""
        if not multiplier:
            return width
        width *= multiplier
        min_width = min_width or divisor
        width_out = max(min_width, int(width + divisor / 2) // divisor * divisor)
        if width_out < 0.9 * width:
            width_out += divisor
        return int(width_out)

    def forward(self, x):
        return _ops.bottleneck_standalone(self, x)


class ResNet(nn.Module):
    """x3d.py:174-345."""

    def __init__(self, block, layers, block_inplanes, n_input_channels=3, shortcut_type='B', widen_factor=1.0,
                 dropout=0.5, n_classes=400, base_bn_splits=8, task='class'):
        super().__init__()
        block_inplanes = [(int(x * widen_factor), int(y * widen_factor)) for x, y in block_inplanes]
        self.index = 0
        self.base_bn_splits = base_bn_splits
        self.task = task
        if shortcut_type != 'B':
            # x3d.py:252-261,266-269: the reference's type-'A' shortcut is F.avg_pool3d(x, kernel_size=1, stride=stride)
            # with an INT stride, i.e. it also halves T while the main branch keeps T (stride (1,s,s)) -- the residual add
            # raises "size of tensor a must match tensor b at dimension 2" on the first forward of any clip with T > 1
            # (checked against the reference module).  No script uses it; there is no working behaviour to reproduce.
            raise NotImplementedError("shortcut_type='A' does not run in the reference either (its shortcut strides T, "
                                      "x3d.py:253 vs :91); only 'B' (the scripts' default) is built")
        self.in_planes = block_inplanes[0][1]
        self.conv1_s = _StemSpatialConv3d(n_input_channels, self.in_planes, kernel_size=(1, 3, 3), stride=(1, 2, 2),
                                          padding=(0, 1, 1), bias=False)
        self.conv1_t = _DepthwiseConv3d(self.in_planes, self.in_planes, kernel_size=(5, 1, 1), stride=(1, 1, 1),
                                        padding=(2, 0, 0), bias=False, groups=self.in_planes)
        self.bn1 = SubBatchNorm3d(num_splits=base_bn_splits, num_features=self.in_planes, affine=True)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer(block, block_inplanes[0], layers[0], shortcut_type, stride=2)
        self.layer2 = self._make_layer(block, block_inplanes[1], layers[1], shortcut_type, stride=2)
        self.layer3 = self._make_layer(block, block_inplanes[2], layers[2], shortcut_type, stride=2)
        self.layer4 = self._make_layer(block, block_inplanes[3], layers[3], shortcut_type, stride=2)
        self.conv5 = _PointwiseConv3d(block_inplanes[3][1], block_inplanes[3][0], kernel_size=(1, 1, 1),
                                      stride=(1, 1, 1), padding=(0, 0, 0), bias=False)
        self.bn5 = SubBatchNorm3d(num_splits=base_bn_splits, num_features=block_inplanes[3][0], affine=True)
        if task == 'class':
            self.avgpool = nn.AdaptiveAvgPool3d((1, 1, 1))
        elif task == 'loc':
            self.avgpool = nn.AdaptiveAvgPool3d((None, 1, 1))
        self.fc1 = _PointwiseConv3d(block_inplanes[3][0], 2048, bias=False, kernel_size=1, stride=1)
        self.fc2 = nn.Linear(2048, n_classes)
        self.dropout = nn.Dropout(dropout)
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        # engine wiring
        self.compute_dtype = torch.bfloat16
        self._engines = {}
        for li in range(1, 5):
            for i, blk in enumerate(getattr(self, f'layer{li}')):
                blk.prefix = f'layer{li}.{i}'
                blk.stage = li

    def _make_layer(self, block, planes, blocks, shortcut_type, stride=1):
        downsample = None
        if stride != 1 or self.in_planes != planes[1]:
            downsample = nn.Sequential(
                conv1x1x1(self.in_planes, planes[1], stride),
                SubBatchNorm3d(num_splits=self.base_bn_splits, num_features=planes[1], affine=True))
        layers = [block(in_planes=self.in_planes, planes=planes, stride=stride, downsample=downsample,
                        index=self.index, base_bn_splits=self.base_bn_splits)]
        self.in_planes = planes[1]
        self.index += 1
        for _ in range(1, blocks):
            layers.append(block(self.in_planes, planes, index=self.index, base_bn_splits=self.base_bn_splits))
            self.index += 1
        self.index = 0
        return nn.Sequential(*layers)

    # ---- reference methods ------------------------------------------------------------
    def replace_logits(self, n_classes):
        """x3d.py:294-295"""
        self.fc2 = nn.Linear(2048, n_classes)

    def update_bn_splits_long_cycle(self, long_cycle_bn_scale):
        """x3d.py:298-303 -- re-creates every split_bn (its running stats restart)."""
        for m in self.modules():
            if isinstance(m, SubBatchNorm3d):
                m.num_splits = self.base_bn_splits * long_cycle_bn_scale
                m.set_split_bn(m.num_splits)
        return self.base_bn_splits * long_cycle_bn_scale

    def aggregate_sub_bn_stats(self):
        """x3d.py:306-313"""
        count = 0
        for m in self.modules():
            if isinstance(m, SubBatchNorm3d):
                m.aggregate_stats()
                count += 1
        return count

    # ---- engine -----------------------------------------------------------------------
    def blocks(self) -> List[Bottleneck]:
        return [b for li in range(1, 5) for b in getattr(self, f'layer{li}')]

    def set_compute_dtype(self, dtype):
        assert dtype in (torch.float32, torch.bfloat16)
        self.compute_dtype = dtype
        return self

    def engine(self) -> '_engine.Engine':
        e = self._engines.get(self.compute_dtype)
        if e is None:
            e = self._engines[self.compute_dtype] = _engine.Engine(self, self.compute_dtype)
        return e

    def forward(self, x):
        return _ops.resnet_forward(self, x)

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica._engines = {}
        return replica


def get_inplanes(version):
    return {'S': [(54, 24), (108, 48), (216, 96), (432, 192)],
            'M': [(54, 24), (108, 48), (216, 96), (432, 192)],
            'XL': [(72, 32), (162, 72), (306, 136), (630, 280)]}[version]


def get_blocks(version):
    return {'S': [3, 5, 11, 7], 'M': [3, 5, 11, 7], 'XL': [5, 10, 25, 15]}[version]


def generate_model(x3d_version, **kwargs):
    """x3d.py:366-368"""
    return ResNet(Bottleneck, get_blocks(x3d_version), get_inplanes(x3d_version), **kwargs)
