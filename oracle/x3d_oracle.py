"""CPU oracle for the X3D hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, as plain functional tensor arithmetic on the CPU, the
algorithm of the reference's ``x3d.py`` (KiyoshiKAWASAKI/X3D-Multigrid).  It is
the checker the CUDA path is compared against; it is never imported by the
product package ``x3d_multigrid_b200`` (only ``tests/``, ``__graft_entry__.smoke``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it).

Parity status: the reference has no tests, golden vectors or checkpoints of its
own -- the oracle is pinned instead against OUTPUTS OF THE
REFERENCE ITSELF: ``oracle/make_golden.py`` imports ``/root/reference/x3d.py``
in the build container, runs it in fp64 on deterministic weights / clips and
commits the results under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this restatement against those vectors (and, when /root/reference is
present, against the live reference module).

Everything takes / returns NCDHW tensors like the reference, any float dtype
(fp64 is the anchor).  ``sd`` is a flat ``{state_dict key: tensor}`` mapping with
the reference's key names.

Two convolution back-ends restate the same math:
  * ``conv_impl="explicit"`` -- shifted-slice sums / matmul, no conv library;
  * ``conv_impl="aten"``     -- ``torch.nn.functional.conv3d``, i.e. exactly the
    ATen call the reference's ``nn.Conv3d`` modules make (used for CPU timing).
``bn_impl="aten"`` additionally issues the reference's exact BatchNorm / Swish calls
(``F.batch_norm`` on the ``(n//s, c*s, ...)`` view + the two affine ops, x3d.py:50-57; the
saved-x ``SwishEfficient`` backward, :71-84) -- the op sequence the timing arms must execute.
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # nn.BatchNorm3d default used by x3d.py:23,25
BN_MOMENTUM = 0.1      # idem

# ----------------------------------------------------------------------------
# architecture tables (x3d.py:352-363)
# ----------------------------------------------------------------------------
_PLANES = {'S': [(54, 24), (108, 48), (216, 96), (432, 192)],
           'M': [(54, 24), (108, 48), (216, 96), (432, 192)],
           'XL': [(72, 32), (162, 72), (306, 136), (630, 280)]}
_BLOCKS = {'S': [3, 5, 11, 7], 'M': [3, 5, 11, 7], 'XL': [5, 10, 25, 15]}


def get_inplanes(version: str):
    """x3d.py:352-356"""
    return list(_PLANES[version])


def get_blocks(version: str):
    """x3d.py:359-363"""
    return list(_BLOCKS[version])


def round_width(width, multiplier=0.0625, min_width=8, divisor=8):
    """SE bottleneck width, x3d.py:129-140."""
    if not multiplier:
        return width
    width *= multiplier
    min_width = min_width or divisor
    width_out = max(min_width, int(width + divisor / 2) // divisor * divisor)
    if width_out < 0.9 * width:
        width_out += divisor
    return int(width_out)


def block_specs(version: str, widen_factor: float = 1.0):
    """[(prefix, in_planes, mid, out, stride, has_se, has_downsample)] in forward
    order -- restates ResNet.__init__/_make_layer, x3d.py:189-230,263-291."""
    planes = [(int(a * widen_factor), int(b * widen_factor)) for a, b in _PLANES[version]]
    blocks = _BLOCKS[version]
    specs = []
    in_planes = planes[0][1]
    for li, ((mid, out), nb) in enumerate(zip(planes, blocks)):
        for i in range(nb):
            stride = 2 if i == 0 else 1
            has_ds = (i == 0)  # stride != 1 always true for block 0 (x3d.py:265)
            specs.append((f'layer{li + 1}.{i}', in_planes, mid, out, stride, i % 2 == 0, has_ds))
            in_planes = out
    return specs


# ----------------------------------------------------------------------------
# deterministic, RNG-free parameter / clip fill (bit-reproducible everywhere)
# ----------------------------------------------------------------------------
def det_uniform(n: int, salt: int) -> np.ndarray:
    """splitmix64 hash of (index+salt) -> float64 uniform in [-1, 1)."""
    with np.errstate(over='ignore'):
        z = (np.arange(n, dtype=np.uint64) + np.uint64(salt & 0xFFFFFFFFFFFFFFFF)) \
            * np.uint64(0x9E3779B97F4A7C15)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53) * 2.0 - 1.0


def det_tensor(shape, key: str, scale: float = 1.0, dtype=torch.float64) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    u = det_uniform(n, zlib.crc32(key.encode()) * 1000003)
    return torch.from_numpy(u * scale).reshape(tuple(shape)).to(dtype)


def det_fill_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Deterministic values for every tensor of an X3D state_dict (same variance
    law as the reference's kaiming fan_out init, x3d.py:246-250).  Returns a new
    dict of fp64 tensors (int64 counters are zeroed)."""
    out = {}
    for k, v in sd.items():
        shp = tuple(v.shape)
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros(shp, dtype=torch.int64)
        elif k.endswith('running_mean'):
            out[k] = torch.zeros(shp, dtype=torch.float64)
        elif k.endswith('running_var'):
            out[k] = torch.ones(shp, dtype=torch.float64)
        elif v.dim() == 5:  # conv weight [O, I/g, kt, kh, kw]
            fan_out = shp[0] * shp[2] * shp[3] * shp[4]
            std = (2.0 / fan_out) ** 0.5
            out[k] = det_tensor(shp, k, scale=std * 3 ** 0.5)
        elif k == 'fc2.weight':
            out[k] = det_tensor(shp, k, scale=(1.0 / shp[1]) ** 0.5 * 3 ** 0.5)
        elif k.endswith('.weight') and v.dim() == 1:  # SubBN gamma
            out[k] = 1.0 + det_tensor(shp, k, scale=0.2)
        elif k.endswith('.bias'):
            out[k] = det_tensor(shp, k, scale=0.1)
        else:
            raise KeyError(f'unexpected state_dict entry {k} {shp}')
    return out


def det_clip(shape, key='clip', dtype=torch.float64) -> torch.Tensor:
    """Synthetic clip, ~unit variance (stands in for Normalize'd frames)."""
    return det_tensor(shape, key, scale=3 ** 0.5, dtype=dtype)


# ----------------------------------------------------------------------------
# primitive ops
# ----------------------------------------------------------------------------
def dwconv3d(x, w, stride: int, conv_impl='explicit'):
    """Channelwise conv, kernel (kt,kh,kw) from w[C,1,kt,kh,kw], pad k//2, stride
    (1,s,s): conv3x3x3 x3d.py:87-95 and conv1_t x3d.py:202-208."""
    C = x.shape[1]
    kt, kh, kw = w.shape[2:]
    pt, ph, pw = kt // 2, kh // 2, kw // 2
    if conv_impl == 'aten':
        return F.conv3d(x, w, None, (1, stride, stride), (pt, ph, pw), 1, C)
    N, _, T, H, W = x.shape
    Ho = (H + 2 * ph - kh) // stride + 1
    Wo = (W + 2 * pw - kw) // stride + 1
    xp = F.pad(x, (pw, pw, ph, ph, pt, pt))
    y = None
    for i in range(kt):
        for j in range(kh):
            for k in range(kw):
                sl = xp[:, :, i:i + T,
                        j:j + (Ho - 1) * stride + 1:stride,
                        k:k + (Wo - 1) * stride + 1:stride]
                term = sl * w[:, 0, i, j, k].view(1, C, 1, 1, 1)
                y = term if y is None else y + term
    return y


def pwconv(x, w, stride: int = 1, bias=None, conv_impl='explicit'):
    """1x1x1 conv, stride (1,s,s): conv1x1x1 x3d.py:98-103; SE fc x3d.py:123-124."""
    if conv_impl == 'aten':
        return F.conv3d(x, w, bias, (1, stride, stride))
    xs = x[:, :, :, ::stride, ::stride]
    y = torch.einsum('nkthw,ok->nothw', xs, w.reshape(w.shape[0], w.shape[1]))
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    return y


def stem_conv_s(x, w, conv_impl='explicit'):
    """conv1_s: kernel (1,3,3), stride (1,2,2), pad (0,1,1), x3d.py:196-201."""
    if conv_impl == 'aten':
        return F.conv3d(x, w, None, (1, 2, 2), (0, 1, 1))
    N, Ci, T, H, W = x.shape
    Ho = (H + 2 - 3) // 2 + 1
    Wo = (W + 2 - 3) // 2 + 1
    xp = F.pad(x, (1, 1, 1, 1, 0, 0))
    y = None
    for j in range(3):
        for k in range(3):
            sl = xp[:, :, :, j:j + (Ho - 1) * 2 + 1:2, k:k + (Wo - 1) * 2 + 1:2]
            term = torch.einsum('nithw,oi->nothw', sl, w[:, :, 0, j, k])
            y = term if y is None else y + term
    return y


def sub_bn_aten(x, prefix, sd, splits: int, training: bool, new_stats: Optional[dict]):
    """SubBatchNorm3d.forward with the reference's exact ATen calls (x3d.py:47-58): F.batch_norm on the
    (n//s, c*s, t, h, w) view with running buffers, then the two elementwise affine ops.  Same math as ``sub_bn``;
    used where the reference's op sequence matters (CPU / GPU timing arms of bench.py)."""
    gamma, beta = sd[prefix + '.weight'], sd[prefix + '.bias']
    N, C = x.shape[:2]
    if training:
        rm = sd.get(prefix + '.split_bn.running_mean')
        rv = sd.get(prefix + '.split_bn.running_var')
        if rm is None or rm.numel() != splits * C:
            rm = torch.zeros(splits * C, dtype=x.dtype, device=x.device)
            rv = torch.ones(splits * C, dtype=x.dtype, device=x.device)
        rm, rv = rm.detach().clone().to(torch.float32 if x.dtype != torch.float64 else x.dtype), \
            rv.detach().clone().to(torch.float32 if x.dtype != torch.float64 else x.dtype)
        y = F.batch_norm(x.view(N // splits, C * splits, *x.shape[2:]), rm, rv, None, None, True, BN_MOMENTUM, BN_EPS)
        y = y.view(x.shape)
        if new_stats is not None:
            new_stats[prefix + '.split_bn.running_mean'] = rm
            new_stats[prefix + '.split_bn.running_var'] = rv
    else:
        y = F.batch_norm(x, sd[prefix + '.bn.running_mean'], sd[prefix + '.bn.running_var'], None, None, False,
                         BN_MOMENTUM, BN_EPS)
    y = y * gamma.view(-1, 1, 1, 1)
    y = y + beta.view(-1, 1, 1, 1)
    return y


class _SwishEfficient(torch.autograd.Function):
    """x3d.py:71-84: saves x, hand-written backward"""

    @staticmethod
    def forward(ctx, x):
        result = x * torch.sigmoid(x)
        ctx.save_for_backward(x)
        return result

    @staticmethod
    def backward(ctx, grad_output):
        (x,) = ctx.saved_tensors
        sigmoid_x = torch.sigmoid(x)
        return grad_output * (sigmoid_x * (1 + x * (1 - sigmoid_x)))


def sub_bn(x, prefix, sd, splits: int, training: bool, new_stats: Optional[dict], impl: str = 'explicit'):
    """SubBatchNorm3d.forward, x3d.py:47-58.  Training: split b = samples b::s
    (the reference's NCDHW view (n//s, c*s, ...)), biased variance, then the
    shared affine.  Running statistics are returned through ``new_stats``
    (momentum 0.1, unbiased variance; index b*C+c) instead of being mutated."""
    if impl == 'aten':
        return sub_bn_aten(x, prefix, sd, splits, training, new_stats)
    gamma, beta = sd[prefix + '.weight'], sd[prefix + '.bias']
    N, C = x.shape[:2]
    if training:
        assert N % splits == 0, 'per-replica batch must be divisible by num_splits (x3d.py:50)'
        xv = x.reshape(N // splits, splits, C, *x.shape[2:])
        m = xv.numel() // (splits * C)
        mean = xv.mean(dim=(0, 3, 4, 5))                              # [s, C]
        var = ((xv - mean.view(1, splits, C, 1, 1, 1)) ** 2).mean(dim=(0, 3, 4, 5))
        xh = (xv - mean.view(1, splits, C, 1, 1, 1)) / torch.sqrt(var.view(1, splits, C, 1, 1, 1) + BN_EPS)
        y = xh.reshape(x.shape)
        if new_stats is not None:
            rm = sd.get(prefix + '.split_bn.running_mean')
            rv = sd.get(prefix + '.split_bn.running_var')
            if rm is None or rm.numel() != splits * C:      # freshly re-split (x3d.py:302)
                rm = torch.zeros(splits * C, dtype=x.dtype)
                rv = torch.ones(splits * C, dtype=x.dtype)
            unb = var.detach() * (m / max(m - 1, 1))
            new_stats[prefix + '.split_bn.running_mean'] = \
                (1 - BN_MOMENTUM) * rm.to(x.dtype) + BN_MOMENTUM * mean.detach().reshape(-1)
            new_stats[prefix + '.split_bn.running_var'] = \
                (1 - BN_MOMENTUM) * rv.to(x.dtype) + BN_MOMENTUM * unb.reshape(-1)
    else:
        rm, rv = sd[prefix + '.bn.running_mean'], sd[prefix + '.bn.running_var']
        y = (x - rm.view(1, C, 1, 1, 1)) / torch.sqrt(rv.view(1, C, 1, 1, 1) + BN_EPS)
    y = y * gamma.view(1, C, 1, 1, 1)
    y = y + beta.view(1, C, 1, 1, 1)
    return y


def aggregate_stats(split_mean, split_var, splits: int):
    """SubBatchNorm3d._get_aggregated_mean_std, x3d.py:27-33 (the 'std' there is
    a variance)."""
    means = split_mean.view(splits, -1)
    mean = means.sum(0) / splits
    var = split_var.view(splits, -1).sum(0) / splits + ((means - mean) ** 2).sum(0) / splits
    return mean, var


def swish(x):
    """x * sigmoid(x), x3d.py:71-84 (autograd reproduces the hand-written bwd)."""
    return x * torch.sigmoid(x)


def bottleneck(x, prefix, sd, stride, has_se, has_ds, splits, training, new_stats, conv_impl,
               taps: Optional[dict] = None, bn_impl: str = 'explicit'):
    """Bottleneck.forward, x3d.py:143-171."""
    out = pwconv(x, sd[prefix + '.conv1.weight'], 1, None, conv_impl)
    out = sub_bn(out, prefix + '.bn1', sd, splits, training, new_stats, bn_impl)
    out = torch.relu(out)
    out = dwconv3d(out, sd[prefix + '.conv2.weight'], stride, conv_impl)
    out = sub_bn(out, prefix + '.bn2', sd, splits, training, new_stats, bn_impl)
    if has_se:
        se = out.mean(dim=(2, 3, 4), keepdim=True)
        se = pwconv(se, sd[prefix + '.fc1.weight'], 1, sd[prefix + '.fc1.bias'], conv_impl)
        se = torch.relu(se)
        se = pwconv(se, sd[prefix + '.fc2.weight'], 1, sd[prefix + '.fc2.bias'], conv_impl)
        se = torch.sigmoid(se)
        out = out * se
    out = _SwishEfficient.apply(out) if bn_impl == 'aten' else swish(out)
    out = pwconv(out, sd[prefix + '.conv3.weight'], 1, None, conv_impl)
    out = sub_bn(out, prefix + '.bn3', sd, splits, training, new_stats, bn_impl)
    if has_ds:
        res = pwconv(x, sd[prefix + '.downsample.0.weight'], stride, None, conv_impl)
        res = sub_bn(res, prefix + '.downsample.1', sd, splits, training, new_stats, bn_impl)
    else:
        res = x
    out = torch.relu(out + res)
    if taps is not None:
        taps[prefix] = out
    return out


def forward(sd, x, version='M', splits=1, training=True, dropout_mask=None, task='class',
            conv_impl='explicit', widen_factor=1.0, new_stats: Optional[dict] = None,
            taps: Optional[dict] = None, bn_impl: str = 'explicit'):
    """ResNet.forward, x3d.py:316-345.  ``dropout_mask`` ([B,2048] or [B,T,2048],
    already scaled by 1/(1-p)) replaces nn.Dropout's RNG; None = no dropout."""
    out = stem_conv_s(x, sd['conv1_s.weight'], conv_impl)
    out = dwconv3d(out, sd['conv1_t.weight'], 1, conv_impl)
    out = sub_bn(out, 'bn1', sd, splits, training, new_stats, bn_impl)
    out = torch.relu(out)
    if taps is not None:
        taps['stem'] = out
    for (prefix, _cin, _mid, _cout, stride, has_se, has_ds) in block_specs(version, widen_factor):
        out = bottleneck(out, prefix, sd, stride, has_se, has_ds, splits, training, new_stats,
                         conv_impl, taps, bn_impl)
    out = pwconv(out, sd['conv5.weight'], 1, None, conv_impl)
    out = sub_bn(out, 'bn5', sd, splits, training, new_stats, bn_impl)
    out = torch.relu(out)
    if task == 'class':
        out = out.mean(dim=(2, 3, 4), keepdim=True)
    else:
        out = out.mean(dim=(3, 4), keepdim=True)
    out = pwconv(out, sd['fc1.weight'], 1, None, conv_impl)
    out = torch.relu(out)
    if task == 'class':
        out = out.flatten(1)                                    # B C
        if dropout_mask is not None:
            out = out * dropout_mask
        out = (out @ sd['fc2.weight'].t() + sd['fc2.bias']).unsqueeze(2)   # B C 1
    else:
        out = out.squeeze(4).squeeze(3).permute(0, 2, 1)        # B T C
        if dropout_mask is not None:
            out = out * dropout_mask
        out = (out @ sd['fc2.weight'].t() + sd['fc2.bias']).permute(0, 2, 1)  # B C T
    return out


def ce_loss(logits, labels):
    """nn.CrossEntropyLoss on [B,C,1] vs [B,1], train_x3d_kinetics_multigrid.py:189,245,259."""
    return F.cross_entropy(logits, labels)


def bce_cls_loss(logits, labels):
    """Charades classification: nn.BCEWithLogitsLoss on logits.squeeze(2) [B,C] vs multi-hot [B,C],
    train_x3d_charades.py:122,162,176."""
    return F.binary_cross_entropy_with_logits(logits.squeeze(2), labels.to(logits.dtype))


def bce_loc_loss(logits, labels):
    """Charades localisation, train_x3d_charades_loc.py:168-189: per-frame logits [B,C,T] are linearly
    interpolated to the label length TL; loss = (BCE(max_t logits, max_t labels) + BCE(logits, labels)) / 2."""
    per_frame = F.interpolate(logits, labels.shape[2], mode='linear')
    lab = labels.to(logits.dtype)
    cls_loss = F.binary_cross_entropy_with_logits(torch.max(per_frame, dim=2)[0], torch.max(lab, dim=2)[0])
    loc_loss = F.binary_cross_entropy_with_logits(per_frame, lab)
    return (cls_loss + loc_loss) / 2


LOSSES = {'ce': ce_loss, 'bce': bce_cls_loss, 'bce_loc': bce_loc_loss}


def loss_and_grads(sd, x, labels, loss='ce', **kw):
    """One fwd+bwd: returns (logits, loss, {param key: grad}, new running stats)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and 'running_' not in k}
    full = dict(sd)
    full.update(params)
    new_stats: dict = {}
    logits = forward(full, x, new_stats=new_stats, **kw)
    loss = LOSSES[loss](logits, labels)
    loss.backward()
    grads = {k: p.grad for k, p in params.items() if p.grad is not None}
    return logits.detach(), loss.detach(), grads, new_stats


def param_keys(sd) -> List[str]:
    return [k for k, v in sd.items() if v.is_floating_point() and 'running_' not in k]


def state_dict_manifest(version='M', n_classes=400, base_bn_splits=1, widen_factor=1.0,
                        n_input_channels=3) -> List[Tuple[str, Tuple[int, ...], str]]:
    """Key/shape/dtype list of the reference's state_dict, derived
    from the constructor logic x3d.py:106-126,176-244 -- used to pin the
    drop-in layout without needing the reference at run time."""
    ent: List[Tuple[str, Tuple[int, ...], str]] = []
    s = base_bn_splits

    def conv(name, o, i, k, bias=False):
        ent.append((name + '.weight', (o, i) + tuple(k), 'float32'))
        if bias:
            ent.append((name + '.bias', (o,), 'float32'))

    def bn(name, c):
        ent.append((name + '.weight', (c,), 'float32'))
        ent.append((name + '.bias', (c,), 'float32'))
        ent.append((name + '.bn.running_mean', (c,), 'float32'))
        ent.append((name + '.bn.running_var', (c,), 'float32'))
        ent.append((name + '.bn.num_batches_tracked', (), 'int64'))
        ent.append((name + '.split_bn.running_mean', (c * s,), 'float32'))
        ent.append((name + '.split_bn.running_var', (c * s,), 'float32'))
        ent.append((name + '.split_bn.num_batches_tracked', (), 'int64'))

    specs = block_specs(version, widen_factor)
    c0 = specs[0][1]
    conv('conv1_s', c0, n_input_channels, (1, 3, 3))
    conv('conv1_t', c0, 1, (5, 1, 1))
    bn('bn1', c0)
    for (p, cin, mid, cout, stride, has_se, has_ds) in specs:
        conv(p + '.conv1', mid, cin, (1, 1, 1))
        bn(p + '.bn1', mid)
        conv(p + '.conv2', mid, 1, (3, 3, 3))
        bn(p + '.bn2', mid)
        conv(p + '.conv3', cout, mid, (1, 1, 1))
        bn(p + '.bn3', cout)
        if has_se:
            w = round_width(mid)
            conv(p + '.fc1', w, mid, (1, 1, 1), bias=True)
            conv(p + '.fc2', mid, w, (1, 1, 1), bias=True)
        if has_ds:
            conv(p + '.downsample.0', cout, cin, (1, 1, 1))
            bn(p + '.downsample.1', cout)
    last_mid, last_out = specs[-1][2], specs[-1][3]
    conv('conv5', last_mid, last_out, (1, 1, 1))
    bn('bn5', last_mid)
    conv('fc1', 2048, last_mid, (1, 1, 1))
    ent.append(('fc2.weight', (n_classes, 2048), 'float32'))
    ent.append(('fc2.bias', (n_classes,), 'float32'))
    return ent


def make_state_dict(version='M', n_classes=400, base_bn_splits=1, widen_factor=1.0,
                    dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """Deterministically filled state_dict with the reference's layout."""
    shapes = {k: torch.empty(shp, dtype=torch.int64 if dt == 'int64' else torch.float64)
              for k, shp, dt in state_dict_manifest(version, n_classes, base_bn_splits, widen_factor)}
    sd = det_fill_state_dict(shapes)
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# ----------------------------------------------------------------------------
# multigrid shape law (kinetics_multigrid.py:205-237, cycle_batch_sampler.py:98-113)
# ----------------------------------------------------------------------------
def multigrid_shapes(base_batch: int, frames: int, crop: int, long_cycle=(8, 4, 2, 1)):
    """{long_ind: [(batch, T, H), ...] per short-cycle state} -- SURVEY.md A3."""
    root2 = int(np.floor(crop / np.sqrt(2)))
    long_shapes = [(frames // 4, root2), (frames // 2, root2), (frames // 2, crop), (frames, crop)]
    table = {}
    for li, (t, c) in enumerate(long_shapes):
        bs = base_batch * long_cycle[li]
        if li in (0, 1):
            table[li] = [(bs * 2, t, int(np.floor(c / np.sqrt(2)))), (bs, t, c)]
        else:
            table[li] = [(bs * 4, t, c // 2), (bs * 2, t, int(np.floor(c / np.sqrt(2)))), (bs, t, c)]
    return table
