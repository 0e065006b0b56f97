#!/usr/bin/env python
"""Micro-benchmark of the element-wise / reduction kernels (bn apply, swish + SE gate, residual) through the C ABI at
the four stage shapes of X3D-M B=16 16x224x224.  GB/s = tensors read + written / CUDA-event time; buffers rotate so
that successive launches do not hit in the 126 MB L2."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from x3d_multigrid_b200 import _lib  # noqa: E402

STAGES = [('s1', 54, 24, 56), ('s2', 108, 48, 28), ('s3', 216, 96, 14), ('s4', 432, 192, 7)]   # name, Cmid, Cout, H


def pad8(c):
    return (c + 7) // 8 * 8


def main():
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    N, T, splits = 16, 16, 2
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
    iters = 10
    for name, Cm, Co, H in STAGES:
        P = T * H * H
        for what, C in (('mid', Cm), ('out', Co)):
            Cp = pad8(C)
            nb = N * P * Cp * 2
            nbuf = max(2, int(400e6 // (3 * nb)) + 1)
            mk = lambda: [torch.randn(N, P, Cp, device='cuda').bfloat16() for _ in range(nbuf)]
            a, b, c = mk(), mk(), mk()
            sc = torch.rand(splits, Cp, device='cuda') + 0.5
            sh = torch.randn(splits, Cp, device='cuda') * 0.1
            coef = torch.randn(3, splits, Cp, device='cuda') * 0.1
            coefn = torch.randn(3, N, Cp, device='cuda') * 0.1
            gate = torch.rand(N, Cp, device='cuda')
            stats = torch.zeros(N, Cp, 2, dtype=torch.float64, device='cuda')
            calls = {
                'bn_bwd_apply(2r1w)': (3, lambda j: L.call('x3d_bn_bwd_apply', a[j].data_ptr(), None, b[j].data_ptr(), coef.data_ptr(), splits, c[j].data_ptr(), N, P, Cp, 1, st)),
                'bn_bwd_reduce_store(3r1w)': (4, lambda j: L.call('x3d_bn_bwd_reduce_store', a[j].data_ptr(), b[j].data_ptr(), c[j].data_ptr(), stats.data_ptr(), c[(j + 1) % nbuf].data_ptr(), N, P, Cp, 1, st)),
                'bn_act_fwd(2r1w)': (3, lambda j: L.call('x3d_bn_act_fwd', a[j].data_ptr(), sc.data_ptr(), sh.data_ptr(), splits, b[j].data_ptr(), None, None, 1, c[j].data_ptr(), N, P, Cp, 1, st)),
                'swish_gate_fwd(1r1w)': (2, lambda j: L.call('x3d_swish_gate_fwd', a[j].data_ptr(), sc.data_ptr(), sh.data_ptr(), splits, gate.data_ptr(), c[j].data_ptr(), N, P, Cp, 1, st)),
                'swish_gate_bwd_reduce(2r)': (2, lambda j: L.call('x3d_swish_gate_bwd_reduce', a[j].data_ptr(), b[j].data_ptr(), sc.data_ptr(), sh.data_ptr(), splits, gate.data_ptr(), stats.data_ptr(), N, P, Cp, 1, st)),
                'swish_gate_bwd_apply(2r1w)': (3, lambda j: L.call('x3d_swish_gate_bwd_apply', a[j].data_ptr(), b[j].data_ptr(), sc.data_ptr(), sh.data_ptr(), splits, gate.data_ptr(), coefn.data_ptr(), c[j].data_ptr(), N, P, Cp, 1, st)),
            }
            for k, (nt, fn) in calls.items():
                if what == 'out' and k.startswith('swish'):
                    continue
                if what == 'mid' and k in ('bn_bwd_reduce_store(3r1w)', 'bn_act_fwd(2r1w)'):
                    continue
                for i in range(3):
                    fn(i % nbuf)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(iters):
                    fn(i % nbuf)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / iters
                gbs = nt * nb / us / 1e3
                print(json.dumps({'stage': name, 'tensor': what, 'Cp': Cp, 'kernel': k, 'us': round(us, 1), 'mb': round(nt * nb / 1e6, 1),
                                  'gbs': round(gbs), 'frac_hbm': round(gbs / peak, 3)}), flush=True)
            del a, b, c
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
