"""Fused SGD (momentum + weight decay) over all parameters in ONE kernel launch.

Same update rule as ``torch.optim.SGD(params, lr, momentum, weight_decay)`` as used by the
reference loop (train_x3d_kinetics_multigrid.py:183,277): d = g*grad_scale + wd*p;
buf = d (first step) | momentum*buf + d;  p -= lr*buf.  One param group semantics are kept
(``param_groups[0]['lr']`` can be edited by LR schedulers / the long-cycle LR law).

``capturable=True`` reads the hyper-parameters from a small device tensor, so that a captured
CUDA graph (graphs.GraphedTrainStep) follows later LR changes: ``sync_hyper()`` (a 16-byte H2D copy,
skipped when nothing changed) is called by GraphedTrainStep before every replay, so LR schedulers that
edit ``param_groups`` directly (MultiStepLR, lr_warmup, the long-cycle law) reach the captured step."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import SgdDesc


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, capturable=False):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self._tables = {}
        self.grad_scale = 1.0       # e.g. 1/world_size when gradients were summed, not averaged
        self.capturable = capturable
        self._hyper = {}            # group index -> device tensor + ring of pinned staging buffers

    def _table(self, gi, group):
        ps = [p for p in group['params'] if p.grad is not None]
        key = tuple(p.grad.data_ptr() for p in ps)
        tab = self._tables.get(gi)
        if tab is not None and tab[0] == key:
            return tab
        arr = (SgdDesc * max(len(ps), 1))()
        mx = 1
        for i, p in enumerate(ps):
            if p.dtype != torch.float32 or not p.is_cuda:
                raise RuntimeError('FusedSGD needs fp32 CUDA parameters')
            st = self.state[p]
            if 'momentum_buffer' not in st:
                st['momentum_buffer'] = torch.zeros_like(p)
                st['fresh'] = True
            if not p.grad.is_contiguous():
                raise RuntimeError('FusedSGD needs contiguous gradients')
            arr[i] = SgdDesc(p.data_ptr(), p.grad.data_ptr(), st['momentum_buffer'].data_ptr(), p.numel())
            mx = max(mx, p.numel())
        dev = host = None
        if ps:
            # pinned staging + async copy: legal inside CUDA-graph capture (the host buffer stays alive in the table)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).pin_memory()
            dev = torch.empty(host.numel(), dtype=torch.uint8, device=ps[0].device)
            dev.copy_(host, non_blocking=True)
        tab = (key, dev, len(ps), mx, ps, host)
        self._tables[gi] = tab
        return tab

    _RING = 4

    def sync_hyper(self, force=False):
        """push lr / momentum / weight_decay / grad_scale of every group to the device (capturable mode).

        Cheap to call before every replay: nothing is copied when the values are unchanged.  The copy is
        asynchronous from a small RING of pinned staging buffers, each guarded by an event, so a step that is
        still queued on the stream never sees the hyper-parameters of a later step."""
        for gi, group in enumerate(self.param_groups):
            if gi not in self._hyper:
                continue
            st = self._hyper[gi]
            vals = (float(group['lr']), float(group['momentum']), float(group['weight_decay']), float(self.grad_scale))
            if vals == st['last'] and not force:
                continue
            slot = st['next']
            st['next'] = (slot + 1) % self._RING
            host, ev = st['hosts'][slot], st['events'][slot]
            if ev is not None:
                ev.synchronize()                      # the copy that last used this slot has executed
            host[0], host[1], host[2], host[3] = vals
            st['dev'].copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(st['dev'].device))
            st['events'][slot] = ev
            st['last'] = vals

    def _hyper_dev(self, gi, group, device):
        if gi not in self._hyper:
            self._hyper[gi] = {'dev': torch.zeros(4, dtype=torch.float32, device=device),
                               'hosts': [torch.zeros(4, dtype=torch.float32).pin_memory() for _ in range(self._RING)],
                               'events': [None] * self._RING, 'next': 0, 'last': None}
            self.sync_hyper()
        return self._hyper[gi]['dev']

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            key, dev, n, mx, ps, _host = self._table(gi, group)
            if n == 0:
                continue
            fresh = [self.state[p].get('fresh', False) for p in ps]
            if any(fresh) and not all(fresh):
                raise RuntimeError('FusedSGD: parameters joined the group after the first step')
            first = all(fresh)
            st = torch.cuda.current_stream(ps[0].device).cuda_stream
            if self.capturable:
                hy = self._hyper_dev(gi, group, ps[0].device)
                L.call('x3d_sgd_step_dev', dev.data_ptr(), n, mx, hy.data_ptr(), 1 if first else 0, st)
            else:
                L.call('x3d_sgd_step', dev.data_ptr(), n, mx, float(group['lr']), float(group['momentum']),
                       float(group['weight_decay']), float(self.grad_scale), 1 if first else 0, st)
            if first:
                for p in ps:
                    self.state[p]['fresh'] = False
        return loss
