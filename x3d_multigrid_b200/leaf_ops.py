"""Stand-alone forward/backward of the leaf modules (SubBatchNorm3d and the Conv3d flavours of x3d.py) for callers
that use them outside Bottleneck/ResNet, e.g. ``model.layer1[0].conv2(x)`` in a debugging session.

NCDHW fp32 in / out (the reference's user-facing layout); inside, the same C-ABI kernels as the fused network path run
on NDHWC fp32 buffers.  The fused path never goes through here.  There is no ATen fallback: CPU tensors raise."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import F32, PackDesc

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _pad8(c):
    return (c + 7) // 8 * 8


def _st(x):
    return torch.cuda.current_stream(x.device).cuda_stream


def _check(x, what):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f'{what}: x3d_multigrid_b200 needs CUDA tensors (there is no CPU fallback)')
    if x.dim() != 5:
        raise RuntimeError(f'{what}: expected a [B, C, T, H, W] tensor')
    return x.contiguous().float()


def _to_ndhwc(x):
    N, C, T, H, W = x.shape
    out = torch.empty(N, T, H, W, _pad8(C), dtype=torch.float32, device=x.device)
    _lib.lib().call('x3d_ncdhw_to_ndhwc', x.data_ptr(), out.data_ptr(), N, C, _pad8(C), T, H, W, F32, _st(x))
    return out


def _to_ncdhw(x, C):
    N, T, H, W, Cp = x.shape
    out = torch.empty(N, C, T, H, W, dtype=torch.float32, device=x.device)
    _lib.lib().call('x3d_ndhwc_to_ncdhw', x.data_ptr(), out.data_ptr(), N, C, Cp, T, H, W, F32, _st(x))
    return out


def _pack(w2d, dst_rows, dst_cols, transpose):
    """fp32 [rows][cols] -> zero-padded fp32 operand buffer (x3d_pack_params)"""
    rows, cols = w2d.shape
    dst = torch.empty(dst_rows, dst_cols, dtype=torch.float32, device=w2d.device)
    desc = (PackDesc * 1)(PackDesc(w2d.data_ptr(), dst.data_ptr(), rows, cols, dst_rows, dst_cols, int(transpose), F32))
    host = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).pin_memory()
    dev = host.to(w2d.device, non_blocking=True)
    _lib.lib().call('x3d_pack_params', dev.data_ptr(), 1, dst_rows * dst_cols, _st(w2d))
    dst._keepalive = (host, dev)
    return dst


# ---------------------------------------------------------------------------------------------------------
class _SubBN(torch.autograd.Function):
    """SubBatchNorm3d.forward (x3d.py:47-58)"""

    @staticmethod
    def forward(ctx, x, mod, gamma, beta):
        L = _lib.lib()
        N, C, T, H, W = x.shape
        Cp, P, st = _pad8(C), T * H * W, _st(x)
        xn = _to_ndhwc(x)
        dev = x.device
        train = mod.training
        splits = mod.num_splits if train else 1
        if train and N % splits:
            raise RuntimeError(f'batch size {N} is not divisible by num_splits {splits} (x3d.py:50)')
        scale, shift, mean, rstd = (torch.empty(splits, Cp, device=dev) for _ in range(4))
        if train:
            stats = torch.zeros(N, Cp, 2, dtype=torch.float64, device=dev)
            L.call('x3d_bn_bwd_reduce', xn.data_ptr(), None, xn.data_ptr(), stats.data_ptr(), N, P, Cp, F32, st)  # sum x, sum x*x
            sb = mod.split_bn
            track = sb.track_running_stats and sb.running_mean is not None
            L.call('x3d_bn_finalize', stats.data_ptr(), N, splits, P, C, Cp, gamma.data_ptr(), beta.data_ptr(),
                   sb.running_mean.data_ptr() if track else None, sb.running_var.data_ptr() if track else None,
                   sb.num_batches_tracked.data_ptr() if track else None,
                   float(sb.momentum if sb.momentum is not None else BN_MOMENTUM), float(sb.eps),
                   scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), st)
        else:
            L.call('x3d_bn_eval_params', gamma.data_ptr(), beta.data_ptr(), mod.bn.running_mean.data_ptr(),
                   mod.bn.running_var.data_ptr(), C, Cp, float(mod.bn.eps), scale.data_ptr(), shift.data_ptr(),
                   mean.data_ptr(), rstd.data_ptr(), st)
        out = torch.empty_like(xn)
        L.call('x3d_bn_act_fwd', xn.data_ptr(), scale.data_ptr(), shift.data_ptr(), splits, None, None, None, 0,
               out.data_ptr(), N, P, Cp, F32, st)
        ctx.save_for_backward(xn, gamma, mean, rstd)
        ctx.meta = (N, C, Cp, P, splits, train)
        return _to_ncdhw(out, C)

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        xn, gamma, mean, rstd = ctx.saved_tensors
        N, C, Cp, P, splits, train = ctx.meta
        dyn = _to_ndhwc(dy.contiguous().float())
        st, dev = _st(dy), dy.device
        stats = torch.zeros(N, Cp, 2, dtype=torch.float64, device=dev)
        L.call('x3d_bn_bwd_reduce', dyn.data_ptr(), None, xn.data_ptr(), stats.data_ptr(), N, P, Cp, F32, st)
        coef = torch.empty(3, splits, Cp, device=dev)
        dgamma, dbeta = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        L.call('x3d_bn_bwd_finalize', stats.data_ptr(), N, splits, P, C, Cp, gamma.data_ptr(), mean.data_ptr(),
               rstd.data_ptr(), int(train), coef.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), st)
        dx = torch.empty_like(xn)
        L.call('x3d_bn_bwd_apply', dyn.data_ptr(), None, xn.data_ptr(), coef.data_ptr(), splits, dx.data_ptr(), N, P, Cp,
               F32, st)
        return _to_ncdhw(dx, C), None, dgamma, dbeta


def sub_batch_norm(mod, x):
    x = _check(x, 'SubBatchNorm3d.forward')
    if mod.affine:
        gamma, beta = mod.weight, mod.bias
    else:
        gamma, beta = torch.ones(mod.num_features, device=x.device), torch.zeros(mod.num_features, device=x.device)
    return _SubBN.apply(x, mod, gamma, beta)


# ---------------------------------------------------------------------------------------------------------
class _Depthwise(torch.autograd.Function):
    """conv3x3x3 / conv1_t (x3d.py:87-95, 202-208): groups == channels, pad k/2, stride (1,s,s)"""

    @staticmethod
    def forward(ctx, x, w, stride):
        L = _lib.lib()
        N, C, T, H, W = x.shape
        kt, kh, kw = w.shape[2:]
        Cp, st = _pad8(C), _st(x)
        xn = _to_ndhwc(x)
        wp = _pack(w.detach().reshape(C, kt * kh * kw).contiguous(), kt * kh * kw, Cp, True)
        Ho, Wo = (H + 2 * (kh // 2) - kh) // stride + 1, (W + 2 * (kw // 2) - kw) // stride + 1
        y = torch.empty(N, T, Ho, Wo, Cp, dtype=torch.float32, device=x.device)
        L.call('x3d_dwconv_fwd', xn.data_ptr(), wp.data_ptr(), y.data_ptr(), N, T, H, W, Cp, kt, kh, kw, stride, None,
               None, 1, 0, None, F32, st)
        ctx.save_for_backward(xn, wp)
        ctx.meta = (N, C, Cp, T, H, W, kt, kh, kw, stride, tuple(w.shape))
        return _to_ncdhw(y, C)

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        xn, wp = ctx.saved_tensors
        N, C, Cp, T, H, W, kt, kh, kw, stride, wshape = ctx.meta
        dyn = _to_ndhwc(dy.contiguous().float())
        st = _st(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dxn = torch.empty_like(xn)
            L.call('x3d_dwconv_dgrad', dyn.data_ptr(), wp.data_ptr(), dxn.data_ptr(), N, T, H, W, Cp, kt, kh, kw, stride,
                   None, None, None, 1, None, F32, st)
            dx = _to_ncdhw(dxn, C)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(wshape, dtype=torch.float32, device=dy.device)
            L.call('x3d_dwconv_wgrad', xn.data_ptr(), dyn.data_ptr(), dw.data_ptr(), N, T, H, W, C, Cp, kt, kh, kw, stride,
                   None, None, 1, 0, F32, st)
        return dx, dw, None


def depthwise_conv(mod, x):
    x = _check(x, 'depthwise Conv3d.forward')
    return _Depthwise.apply(x, mod.weight, int(mod.stride[1]))


# ---------------------------------------------------------------------------------------------------------
class _Pointwise(torch.autograd.Function):
    """conv1x1x1 (x3d.py:98-103): y[m][n] = sum_k x[row(m)][k] w[n][k], rows gathered with stride (1,s,s)"""

    @staticmethod
    def forward(ctx, x, w, bias, stride):
        L = _lib.lib()
        N, K, T, H, W = x.shape
        Nn = w.shape[0]
        Kp, Np, st = _pad8(K), _pad8(Nn), _st(x)
        xn = _to_ndhwc(x)
        w2 = w.detach().reshape(Nn, K).contiguous()
        wf = _pack(w2, Np, Kp, False)
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        y = torch.empty(N, T, Ho, Wo, Np, dtype=torch.float32, device=x.device)
        L.call('x3d_pwconv_fwd', xn.data_ptr(), wf.data_ptr(), y.data_ptr(), N, T, H, W, Kp, Np, stride, None, F32, st)
        ctx.save_for_backward(xn, w2)
        ctx.meta = (N, K, Kp, Nn, Np, T, H, W, stride, tuple(w.shape))
        if bias is not None:        # SE fc1 / fc2 are the only biased 1x1x1 convs (x3d.py:123-124)
            ones = torch.ones(Np, device=x.device)
            shift = torch.zeros(Np, device=x.device)
            shift[:Nn] = bias.detach()
            yb = torch.empty_like(y)   # y*1 + bias[c] through the BN-apply kernel (scale 1, shift bias)
            L.call('x3d_bn_act_fwd', y.data_ptr(), ones.data_ptr(), shift.data_ptr(), 1, None, None, None, 0,
                   yb.data_ptr(), N, T * Ho * Wo, Np, F32, st)
            y = yb
        return _to_ncdhw(y, Nn)

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        xn, w2 = ctx.saved_tensors
        N, K, Kp, Nn, Np, T, H, W, stride, wshape = ctx.meta
        dyn = _to_ndhwc(dy.contiguous().float())
        st = _st(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wt = _pack(w2, Kp, Np, True)
            dxn = torch.zeros_like(xn) if stride > 1 else torch.empty_like(xn)
            L.call('x3d_pwconv_dgrad', dyn.data_ptr(), wt.data_ptr(), dxn.data_ptr(), N, T, H, W, Kp, Np, stride, 0, F32,
                   st)
            dx = _to_ncdhw(dxn, K)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(wshape, dtype=torch.float32, device=dy.device)
            L.call('x3d_pwconv_wgrad', xn.data_ptr(), dyn.data_ptr(), dw.data_ptr(), N, T, H, W, K, Kp, Nn, Np, stride,
                   F32, st)
        db = None
        if ctx.needs_input_grad[2]:
            Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
            acc = torch.zeros(Np, dtype=torch.float32, device=dy.device)
            L.call('x3d_colsum', dyn.data_ptr(), N * T * Ho * Wo, Np, acc.data_ptr(), st)
            db = acc[:Nn].clone()
        return dx, dw, db, None


def pointwise_conv(mod, x):
    x = _check(x, 'pointwise Conv3d.forward')
    return _Pointwise.apply(x, mod.weight, mod.bias, int(mod.stride[1]))


# ---------------------------------------------------------------------------------------------------------
class _StemSpatial(torch.autograd.Function):
    """conv1_s (x3d.py:196-201): dense 1x3x3, stride (1,2,2), pad (0,1,1), reads the NCDHW clip directly"""

    @staticmethod
    def forward(ctx, x, w):
        L = _lib.lib()
        N, Ci, T, H, W = x.shape
        Co = w.shape[0]
        Cop, st = _pad8(Co), _st(x)
        H1, W1 = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        y = torch.empty(N, T, H1, W1, Cop, dtype=torch.float32, device=x.device)
        wc = w.detach().contiguous()
        L.call('x3d_stem_conv_s_fwd', x.data_ptr(), wc.data_ptr(), y.data_ptr(), N, Ci, T, H, W, Co, Cop, F32, st)
        ctx.save_for_backward(x)
        ctx.meta = (N, Ci, T, H, W, Co, Cop, tuple(w.shape))
        return _to_ncdhw(y, Co)

    @staticmethod
    def backward(ctx, dy):
        if ctx.needs_input_grad[0]:
            raise NotImplementedError('gradients w.r.t. the input clip are not produced by the stem kernels')
        L = _lib.lib()
        (x,) = ctx.saved_tensors
        N, Ci, T, H, W, Co, Cop, wshape = ctx.meta
        dyn = _to_ndhwc(dy.contiguous().float())
        dw = torch.zeros(wshape, dtype=torch.float32, device=dy.device)
        L.call('x3d_stem_conv_s_wgrad', x.data_ptr(), dyn.data_ptr(), dw.data_ptr(), N, Ci, T, H, W, Co, Cop, F32,
               _st(dy))
        return None, dw


def stem_conv(mod, x):
    x = _check(x, 'stem Conv3d.forward')
    return _StemSpatial.apply(x, mod.weight)
