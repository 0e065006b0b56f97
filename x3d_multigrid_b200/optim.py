"""Fused SGD (momentum + weight decay) over all parameters in ONE kernel launch.

Same update rule as ``torch.optim.SGD(params, lr, momentum, weight_decay)`` as used by the
reference loop (train_x3d_kinetics_multigrid.py:183,277): d = g*grad_scale + wd*p;
buf = d (first step) | momentum*buf + d;  p -= lr*buf.  One param group semantics are kept
(``param_groups[0]['lr']`` can be edited by LR schedulers / the long-cycle LR law)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import SgdDesc


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self._tables = {}
        self.grad_scale = 1.0       # e.g. 1/world_size when gradients were summed, not averaged

    def _table(self, gi, group):
        ps = [p for p in group['params'] if p.grad is not None]
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        tab = self._tables.get(gi)
        if tab is not None and tab[0] == key:
            return tab
        first = False
        arr = (SgdDesc * len(ps))()
        mx = 1
        for i, p in enumerate(ps):
            if p.dtype != torch.float32 or not p.is_cuda:
                raise RuntimeError('FusedSGD needs fp32 CUDA parameters')
            st = self.state[p]
            if 'momentum_buffer' not in st:
                st['momentum_buffer'] = torch.zeros_like(p)
                st['fresh'] = True
                first = True
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            arr[i] = SgdDesc(p.data_ptr(), g.data_ptr(), st['momentum_buffer'].data_ptr(), p.numel())
            mx = max(mx, p.numel())
        dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(ps[0].device) if ps else None
        tab = (key, dev, len(ps), mx, ps)
        self._tables[gi] = tab
        return tab

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            key, dev, n, mx, ps = self._table(gi, group)
            if n == 0:
                continue
            fresh = [self.state[p].get('fresh', False) for p in ps]
            if any(fresh) and not all(fresh):
                raise RuntimeError('FusedSGD: parameters joined the group after the first step')
            first = all(fresh)
            st = torch.cuda.current_stream(ps[0].device).cuda_stream
            L.call('x3d_sgd_step', dev.data_ptr(), n, mx, float(group['lr']), float(group['momentum']),
                   float(group['weight_decay']), float(self.grad_scale), 1 if first else 0, st)
            if first:
                for p in ps:
                    self.state[p]['fresh'] = False
        return loss
