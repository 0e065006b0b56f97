"""autograd glue between the nn.Module surface (x3d.py) and the kernel engine.

Only CUDA tensors are accepted; everything numeric happens in libx3d_b200.so.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib
from . import engine as _engine
from ._lib import F32
from .engine import pad8, _ptr


def _require_cuda(x, what):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f'{what}: x3d_multigrid_b200 needs CUDA tensors (there is no CPU fallback)')


# ---------------------------------------------------------------------------------------
# whole network
# ---------------------------------------------------------------------------------------
class _NetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, model, mask, *params):
        eng = model.engine()
        logits, save = eng.forward(x, model.training, True, mask)
        ctx.eng, ctx.save = eng, save
        ctx.n_params = len(params)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.save is None:
            raise RuntimeError('x3d backward called twice (saved activations were released)')
        grads = ctx.eng.backward(ctx.save, dlogits)
        ctx.save = None
        out = [g if need else None for g, need in zip(grads, ctx.needs_input_grad[3:])]
        return (None, None, None, *out)


def _dropout_mask(model, rows, device):
    p = float(model.dropout.p)
    if not model.training or p <= 0.0:
        return None
    if p >= 1.0:
        return torch.zeros(rows, model.fc1.out_channels, device=device)
    keep = torch.empty(rows, model.fc1.out_channels, device=device).bernoulli_(1.0 - p)
    return keep.mul_(1.0 / (1.0 - p))


def resnet_forward(model, x, dropout_mask: Optional[torch.Tensor] = None):
    """ResNet.forward (x3d.py:316-345)."""
    from .input_pipeline import UInt8Clips
    if not isinstance(x, UInt8Clips):
        _require_cuda(x, 'ResNet.forward')
        if x.dim() != 5:
            raise RuntimeError('expected a [B, C, T, H, W] clip batch')
        if x.requires_grad:
            raise NotImplementedError('gradients w.r.t. the input clip are not produced by this path')
    B, _, T = x.shape[:3]
    rows = B if model.task == 'class' else B * T
    mask = dropout_mask if dropout_mask is not None else _dropout_mask(model, rows, x.device)
    params = list(model.parameters())
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return _NetFunction.apply(x, model, mask, *params)
    logits, _ = model.engine().forward(x, model.training, False, mask)
    return logits


# ---------------------------------------------------------------------------------------
# stand-alone Bottleneck (NCDHW fp32 in / out)
# ---------------------------------------------------------------------------------------
class _BlockHost:
    """Minimal 'model' view of one Bottleneck for engine.Engine."""

    def __init__(self, blk):
        self.blk = blk

    def named_parameters(self):
        return [(f'{self.blk.prefix}.{k}', p) for k, p in self.blk.named_parameters()]

    def parameters(self):
        return list(self.blk.parameters())

    def blocks(self):
        return [self.blk]


def _block_engine(blk, dtype):
    cache = blk.__dict__.setdefault('_engines', {})
    e = cache.get(dtype)
    if e is None:
        e = cache[dtype] = _engine.Engine(_BlockHost(blk), dtype)
    return e


class _BlockFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, blk, dtype, *params):
        eng = _block_engine(blk, dtype)
        with torch.cuda.device(x.device):
            eng.prepare(x.device)
            eng.arena_f.begin()
            eng.pack_weights()
            N, C, T, H, W = x.shape
            xin = eng.to_ndhwc(x.contiguous().float())
            save: List = []
            out, (_, _, Ho, Wo) = eng.block_fwd(blk, xin, (N, T, H, W), blk.training, save)
            ctx.eng, ctx.rec = eng, save[0]
            ctx.arena = eng.arena_f.frame()               # keeps st2 (read by the backward pass) alive
            ctx.in_shape = x.shape
            return eng.to_ncdhw(out, blk.out_planes)

    @staticmethod
    def backward(ctx, dy):
        eng = ctx.eng
        with torch.cuda.device(eng.device):
            eng.arena_b.begin()
            eng.new_grad_buffer()
            d = eng.to_ndhwc(dy.contiguous().float())
            dx = eng.block_bwd(ctx.rec, d, need_dx=True)
            eng._join_side()
            grads = eng.param_grads()
            out = [g if need else None for g, need in zip(grads, ctx.needs_input_grad[3:])]
            dxo = eng.to_ncdhw(dx, ctx.in_shape[1]) if ctx.needs_input_grad[0] else None
        return (dxo, None, None, *out)


def bottleneck_standalone(blk, x, dtype=None):
    """Bottleneck.forward (x3d.py:143-171) on an NCDHW fp32 tensor."""
    _require_cuda(x, 'Bottleneck.forward')
    dtype = dtype or getattr(blk, 'compute_dtype', torch.float32)
    return _BlockFunction.apply(x, blk, dtype, *list(blk.parameters()))


# ---------------------------------------------------------------------------------------
# leaf modules used on their own (not on the fused network path)
# ---------------------------------------------------------------------------------------
def _flat8(x):
    """view any fp32 CUDA tensor as a [1][P][8] channels-last slab (zero padded)"""
    n = x.numel()
    flat = x.contiguous().view(-1)
    pad = (-n) % 8
    if pad:
        flat = torch.cat([flat, flat.new_zeros(pad)])
    return flat, n


def swish_fwd(x):
    _require_cuda(x, 'Swish')
    L = _lib.lib()
    flat, n = _flat8(x.float())
    one = torch.ones(8, device=x.device)
    zero = torch.zeros(8, device=x.device)
    out = torch.empty_like(flat)
    st = torch.cuda.current_stream(x.device).cuda_stream
    L.call('x3d_swish_gate_fwd', _ptr(flat), _ptr(one), _ptr(zero), 1, None, _ptr(out), 1, flat.numel() // 8, 8, F32, st)
    return out[:n].view(x.shape).to(x.dtype)


def swish_bwd(x, dy):
    L = _lib.lib()
    flat, n = _flat8(x.float())
    dflat, _ = _flat8(dy.float())
    one = torch.ones(8, device=x.device)
    zero = torch.zeros(8, device=x.device)
    coef = torch.tensor([1.0, 0.0, 0.0], device=x.device).repeat_interleave(8).contiguous()    # planar [3][1][8]
    out = torch.empty_like(flat)
    st = torch.cuda.current_stream(x.device).cuda_stream
    L.call('x3d_swish_gate_bwd_apply', _ptr(dflat), _ptr(flat), _ptr(one), _ptr(zero), 1, None, _ptr(coef), _ptr(out),
           1, flat.numel() // 8, 8, F32, st)
    return out[:n].view(x.shape).to(dy.dtype)


def sub_batch_norm(mod, x):
    from . import leaf_ops
    return leaf_ops.sub_batch_norm(mod, x)


def depthwise_conv(mod, x):
    from . import leaf_ops
    return leaf_ops.depthwise_conv(mod, x)


def pointwise_conv(mod, x):
    from . import leaf_ops
    return leaf_ops.pointwise_conv(mod, x)


def stem_conv(mod, x):
    from . import leaf_ops
    return leaf_ops.stem_conv(mod, x)
