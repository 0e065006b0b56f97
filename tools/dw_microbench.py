#!/usr/bin/env python
"""Micro-benchmark of the depthwise-conv kernels through the C ABI (one layer shape per line).

  python tools/dw_microbench.py [--layers m_b16] [--iters 20] [--only fwd] [--json out.json]

Each kernel is timed with CUDA events on the launching stream; inputs rotate through enough buffers
to exceed the 126 MB L2 between timed launches.  GB/s = algorithmic bytes / time.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from x3d_multigrid_b200 import _lib  # noqa: E402

# X3D-M, batch 16, 16x224x224: (name, C, H_in, stride, count) of the conv2 layers
M_B16 = [('l1.0', 54, 112, 2, 1), ('l1.x', 54, 56, 1, 2), ('l2.0', 108, 56, 2, 1), ('l2.x', 108, 28, 1, 4),
         ('l3.0', 216, 28, 2, 1), ('l3.x', 216, 14, 1, 10), ('l4.0', 432, 14, 2, 1), ('l4.x', 432, 7, 1, 6)]


def pad8(c):
    return (c + 7) // 8 * 8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--frames', type=int, default=16)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--only', default='fwd,dgrad,wgrad')
    ap.add_argument('--layer', default='')
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--json', default='')
    a = ap.parse_args()
    L = _lib.lib()
    dt = torch.bfloat16 if a.dtype == 'bf16' else torch.float32
    dti = 1 if a.dtype == 'bf16' else 0
    eb = 2 if a.dtype == 'bf16' else 4
    st = torch.cuda.current_stream().cuda_stream
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
    N, T = a.batch, a.frames
    out = []
    tot = {'fwd': [0.0, 0.0], 'dgrad': [0.0, 0.0], 'wgrad': [0.0, 0.0]}
    for name, C, H, s, cnt in M_B16:
        if a.layer and a.layer != name:
            continue
        Cp = pad8(C)
        Ho = (H + 2 - 3) // s + 1
        in_bytes = N * T * H * H * Cp * eb
        nbuf = max(2, int(300e6 // in_bytes) + 1)
        xs = [torch.randn(N, T, H, H, Cp, device='cuda').to(dt) for _ in range(nbuf)]
        ys = [torch.empty(N, T, Ho, Ho, Cp, device='cuda', dtype=dt) for _ in range(nbuf)]
        dys = [torch.randn(N, T, Ho, Ho, Cp, device='cuda').to(dt) for _ in range(nbuf)]
        dxs = [torch.empty(N, T, H, H, Cp, device='cuda', dtype=dt) for _ in range(nbuf)]
        w = torch.randn(27, Cp, device='cuda') * 0.2
        sc = torch.rand(2, Cp, device='cuda') + 0.5
        sh = torch.randn(2, Cp, device='cuda') * 0.3
        stats = torch.zeros(N, Cp, 2, dtype=torch.float64, device='cuda')
        dw = torch.zeros(C, 27, device='cuda')
        alg = {'fwd': N * T * C * (H * H + Ho * Ho) * eb + 27 * C * eb,
               'dgrad': N * T * C * (Ho * Ho + 2 * H * H) * eb,       # + the saved conv1 output for the mask/BN epilogue
               'wgrad': N * T * C * (H * H + Ho * Ho) * eb + 27 * C * 4}

        def run(kind, i):
            j = i % nbuf
            if kind == 'fwd':
                L.call('x3d_dwconv_fwd', xs[j].data_ptr(), w.data_ptr(), ys[j].data_ptr(), N, T, H, H, Cp, 3, 3, 3, s,
                       sc.data_ptr(), sh.data_ptr(), 2, 1, stats.data_ptr(), dti, st)
            elif kind == 'dgrad':
                L.call('x3d_dwconv_dgrad', dys[j].data_ptr(), w.data_ptr(), dxs[j].data_ptr(), N, T, H, H, Cp, 3, 3, 3, s,
                       xs[j].data_ptr(), sc.data_ptr(), sh.data_ptr(), 2, stats.data_ptr(), dti, st)
            else:
                L.call('x3d_dwconv_wgrad', xs[j].data_ptr(), dys[j].data_ptr(), dw.data_ptr(), N, T, H, H, C, Cp, 3, 3, 3,
                       s, sc.data_ptr(), sh.data_ptr(), 2, 1, dti, st)

        row = {'layer': name, 'C': C, 'H': H, 'stride': s, 'count': cnt}
        for kind in a.only.split(','):
            for i in range(3):
                run(kind, i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.iters):
                run(kind, i)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / a.iters
            gbs = alg[kind] / us / 1e3
            row[kind] = {'us': us, 'gbs': gbs, 'frac': gbs / peaks['hbm_gbs'], 'alg_mb': alg[kind] / 1e6}
            tot[kind][0] += us * cnt
            tot[kind][1] += alg[kind] * cnt
        out.append(row)
        print(json.dumps(row), flush=True)
        del xs, ys, dys, dxs
        torch.cuda.empty_cache()
    summary = {k: {'ms_per_step': v[0] / 1e3, 'gbs': v[1] / v[0] / 1e3 if v[0] else None,
                   'frac': v[1] / v[0] / 1e3 / peaks['hbm_gbs'] if v[0] else None} for k, v in tot.items()}
    print(json.dumps({'summary_x3d_m_b16': summary, 'hbm_peak_gbs': peaks['hbm_gbs']}), flush=True)
    if a.json:
        with open(a.json, 'w') as f:
            json.dump({'layers': out, 'summary': summary, 'hbm_peak_gbs': peaks['hbm_gbs']}, f, indent=1)


if __name__ == '__main__':
    main()
