"""x3d_multigrid_b200 -- B200-native (sm_100a) implementation of the X3D training hot path of
KiyoshiKAWASAKI/X3D-Multigrid, behind the reference's own nn.Module surface."""
from .x3d import (Bottleneck, ResNet, SubBatchNorm3d, Swish, SwishEfficient, conv1x1x1, conv3x3x3,
                  generate_model, get_blocks, get_inplanes)

from .input_pipeline import UInt8Clips, clip_from_uint8, crop_table

__all__ = ['UInt8Clips', 'clip_from_uint8', 'crop_table', 'Bottleneck', 'ResNet', 'SubBatchNorm3d', 'Swish', 'SwishEfficient', 'conv1x1x1', 'conv3x3x3',
           'generate_model', 'get_blocks', 'get_inplanes']
