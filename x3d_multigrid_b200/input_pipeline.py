"""GPU-side input pipeline: decoded uint8 frames -> what the reference's DataLoader delivers.

The reference prepares every clip on CPU workers (12 of them, train_x3d_kinetics_multigrid.py:99) with PIL:
``MultiScaleRandomCropMultigrid`` (square window of the frame, resized to the crop size), ``RandomHorizontalFlip``,
``ToTensor(255)`` and ``Normalize(KINETICS_MEAN, KINETICS_STD)`` (transforms/spatial_transforms.py:472-501, 331-349,
35-119), then ships fp32 [B,3,T,H,W] clips to the GPU (154 MB per 16-clip batch).  Here the frames stay uint8
([B,T,Hs,Ws,3], 4x fewer bytes over PCIe) and the integer crop window, the flip, the 1/255 scaling and the
normalisation happen on the device with the reference's exact fp32 arithmetic:

* ``clip_from_uint8`` materialises the fp32 NCDHW clip (drop-in input for ``ResNet.forward``);
* ``UInt8Clips`` handed to ``ResNet.forward`` instead of a tensor: the clip crosses PCIe as uint8 and is expanded on
  the device (one kernel, bit-identical to the reference's transforms); with ``fused=True`` the stem kernels (conv1_s
  forward and weight gradient) read the frames directly and the fp32 clip never exists (saves 154 MB of HBM per
  16-clip batch at the price of slower, byte-granular stem kernels).

The resize of the window to the crop size (PIL's antialiased bilinear filter) is NOT done here: frames are expected at
the scale the crop is taken at (decoder- or loader-side resize); windows are S x S pixels of the source frames."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib

KINETICS_MEAN = (110.63666788 / 255, 103.16065604 / 255, 96.29023126 / 255)    # train_x3d_kinetics_multigrid.py:45-46
KINETICS_STD = (38.7568578 / 255, 37.88248729 / 255, 40.02898126 / 255)


class CropDesc(ctypes.Structure):
    """x3d_crop_t"""
    _fields_ = [('x1', ctypes.c_int32), ('y1', ctypes.c_int32), ('flip', ctypes.c_int32), ('reserved', ctypes.c_int32)]


def crop_table(x1: Sequence[int], y1: Sequence[int], flip: Sequence[bool], device) -> torch.Tensor:
    """per-clip crop windows / flips as a device int32 [B,4] table (x3d_crop_t layout)"""
    t = torch.tensor([[int(a), int(b), int(bool(f)), 0] for a, b, f in zip(x1, y1, flip)], dtype=torch.int32)
    return t.to(device, non_blocking=True)


@dataclass
class UInt8Clips:
    """A batch of clips still in decoded form: ``frames`` uint8 [B,T,Hs,Ws,3] on the GPU, ``crops`` int32 [B,4]
    (x1, y1, flip, 0) on the GPU, crop size ``size`` (= H = W of the clip the network sees)."""
    frames: torch.Tensor
    crops: torch.Tensor
    size: int
    mean: Sequence[float] = KINETICS_MEAN
    std: Sequence[float] = KINETICS_STD
    norm_value: float = 255.0
    fused: bool = False       # True: conv1_s reads the frames directly (no fp32 clip in HBM at all, but byte-granular
                              # loads make the stem kernels ~1 ms slower per 16-clip step); False (default): ONE kernel
                              # materialises the fp32 clip on the device (50 us), then the regular stem kernels run

    def __post_init__(self):
        f = self.frames
        if not (isinstance(f, torch.Tensor) and f.is_cuda and f.dtype == torch.uint8 and f.dim() == 5 and f.shape[-1] == 3):
            raise RuntimeError('UInt8Clips.frames must be a CUDA uint8 tensor [B, T, Hs, Ws, 3]')
        if not (self.crops.is_cuda and self.crops.dtype == torch.int32 and tuple(self.crops.shape) == (f.shape[0], 4)):
            raise RuntimeError('UInt8Clips.crops must be a CUDA int32 tensor [B, 4] (x1, y1, flip, 0)')
        if self.size > f.shape[2] or self.size > f.shape[3]:
            raise RuntimeError('crop size larger than the source frames')
        self.frames = f.contiguous()
        self.crops = self.crops.contiguous()
        self._ms = (ctypes.c_float * 6)(*[float(v) for v in self.mean], *[float(v) for v in self.std])

    # what ResNet.forward needs to know about its "input tensor"
    @property
    def shape(self):
        B, T = self.frames.shape[:2]
        return torch.Size((B, 3, T, self.size, self.size))

    @property
    def device(self):
        return self.frames.device

    def mean_std_ptr(self):
        return ctypes.addressof(self._ms)

    def check_windows(self):
        """host-side validation of the crop windows (synchronises; debugging aid)"""
        c = self.crops.cpu()
        Hs, Ws = self.frames.shape[2:4]
        ok = bool(((c[:, 0] >= 0) & (c[:, 1] >= 0) & (c[:, 0] + self.size <= Ws) & (c[:, 1] + self.size <= Hs)).all())
        if not ok:
            raise RuntimeError('crop window outside the source frames')


def clip_from_uint8(clips: UInt8Clips) -> torch.Tensor:
    """fp32 [B,3,T,S,S] clip exactly as the reference's spatial transform chain produces it from these frames"""
    B, T, Hs, Ws, _ = clips.frames.shape
    S = clips.size
    out = torch.empty(B, 3, T, S, S, dtype=torch.float32, device=clips.device)
    with torch.cuda.device(clips.device):
        _lib.lib().call('x3d_clip_u8_to_f32', clips.frames.data_ptr(), clips.crops.data_ptr(), out.data_ptr(), B, T, Hs, Ws, S,
                        clips.mean_std_ptr(), float(clips.norm_value), torch.cuda.current_stream(clips.device).cuda_stream)
    return out
