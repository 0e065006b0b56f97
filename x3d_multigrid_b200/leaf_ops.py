"""Stand-alone forward/backward of the leaf modules (SubBatchNorm3d, the Conv3d flavours) for
callers that use them outside Bottleneck/ResNet.  fp32 storage, same kernels as the fused path."""
from __future__ import annotations


def _todo(name):
    raise NotImplementedError(
        f'{name}: stand-alone leaf execution is not wired yet; call it through Bottleneck / ResNet. '
        f'(No ATen fallback is provided on purpose.)')


def sub_batch_norm(mod, x):
    _todo('SubBatchNorm3d.forward')


def depthwise_conv(mod, x):
    _todo('depthwise Conv3d.forward')


def pointwise_conv(mod, x):
    _todo('pointwise Conv3d.forward')


def stem_conv(mod, x):
    _todo('stem Conv3d.forward')
