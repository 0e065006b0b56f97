#!/usr/bin/env python
"""Micro-benchmark of the pointwise-conv (1x1x1) GEMM kernels through the C ABI, per X3D-M layer shape.

  python tools/pw_microbench.py [--only fwd,dgrad,wgrad] [--layer l1.c1] [--json out.json]

GB/s = algorithmic bytes (M*(K+N)*2 + K*N*2 for fwd) / CUDA-event time; buffers rotate so that
successive launches do not hit in the 126 MB L2.  TFLOP/s is reported next to it.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from x3d_multigrid_b200 import _lib  # noqa: E402

# X3D-M, batch 16, 16x224x224: (name, K, N, H(=W), count); M = 16*16*H*H  
M_B16 = [('l1.0.c1', 24, 54, 112, 1), ('l1.c3', 54, 24, 56, 3), ('l1.c1', 24, 54, 56, 2), ('l2.0.c1', 24, 108, 56, 1),
         ('l2.c3', 108, 48, 28, 5), ('l2.c1', 48, 108, 28, 4), ('l3.0.c1', 48, 216, 28, 1), ('l3.c3', 216, 96, 14, 11),
         ('l3.c1', 96, 216, 14, 10), ('l4.0.c1', 96, 432, 14, 1), ('l4.c3', 432, 192, 7, 7), ('l4.c1', 192, 432, 7, 7)]


def pad8(c):
    return (c + 7) // 8 * 8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--frames', type=int, default=16)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--only', default='fwd,dgrad,wgrad')
    ap.add_argument('--layer', default='')
    ap.add_argument('--json', default='')
    a = ap.parse_args()
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    WS = torch.empty(32 << 20, dtype=torch.uint8, device='cuda')      # scratch of the two-stage wgrad reduction
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
    N, T = a.batch, a.frames
    tot = {k: [0.0, 0.0] for k in ('fwd', 'dgrad', 'wgrad')}
    rows = []
    for name, K, Nn, H, cnt in M_B16:
        if a.layer and a.layer != name:
            continue
        Kp, Np = pad8(K), pad8(Nn)
        M = N * T * H * H
        nbuf = max(2, int(300e6 // (M * (Kp + Np) * 2)) + 1)
        xs = [torch.randn(N, T, H, H, Kp, device='cuda').bfloat16() for _ in range(nbuf)]
        ys = [torch.empty(N, T, H, H, Np, device='cuda', dtype=torch.bfloat16) for _ in range(nbuf)]
        dys = [torch.randn(N, T, H, H, Np, device='cuda').bfloat16() for _ in range(nbuf)]
        dxs = [torch.empty(N, T, H, H, Kp, device='cuda', dtype=torch.bfloat16) for _ in range(nbuf)]
        w = (torch.randn(Np, Kp, device='cuda') * 0.1).bfloat16()
        wt = w.t().contiguous()
        stats = torch.zeros(N, Np, 2, dtype=torch.float64, device='cuda')
        dw = torch.zeros(Nn, K, device='cuda')
        alg = {'fwd': M * (K + Nn) * 2 + K * Nn * 2, 'dgrad': M * (K + Nn) * 2 + K * Nn * 2,
               'wgrad': M * (K + Nn) * 2 + K * Nn * 4}

        def run(kind, i):
            j = i % nbuf
            if kind == 'fwd':
                L.call('x3d_pwconv_fwd', xs[j].data_ptr(), w.data_ptr(), ys[j].data_ptr(), N, T, H, H, Kp, Np, 1,
                       stats.data_ptr(), 1, st)
            elif kind == 'dgrad':
                L.call('x3d_pwconv_dgrad', dys[j].data_ptr(), wt.data_ptr(), dxs[j].data_ptr(), N, T, H, H, Kp, Np, 1, 0,
                       1, st)
            else:
                if os.environ.get('X3D_WG_ATOMIC'):
                    L.call('x3d_pwconv_wgrad', xs[j].data_ptr(), dys[j].data_ptr(), dw.data_ptr(), N, T, H, H, K, Kp, Nn, Np,
                           1, 1, st)
                else:
                    L.call('x3d_pwconv_wgrad_ws', xs[j].data_ptr(), dys[j].data_ptr(), dw.data_ptr(), N, T, H, H, K, Kp, Nn,
                           Np, 1, WS.data_ptr(), WS.numel(), 1, st)

        row = {'layer': name, 'K': K, 'N': Nn, 'M': M, 'count': cnt}
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
            row[kind] = {'us': round(us, 2), 'gbs': round(gbs, 1), 'frac_hbm': round(gbs / peaks['hbm_gbs'], 3),
                         'tflops': round(2.0 * M * K * Nn / us / 1e6, 1)}
            tot[kind][0] += us * cnt
            tot[kind][1] += alg[kind] * cnt
        rows.append(row)
        print(json.dumps(row), flush=True)
        del xs, ys, dys, dxs
        torch.cuda.empty_cache()
    summary = {k: {'ms_per_step': round(v[0] / 1e3, 3), 'gbs': round(v[1] / v[0] / 1e3, 1) if v[0] else None,
                   'frac_hbm': round(v[1] / v[0] / 1e3 / peaks['hbm_gbs'], 3) if v[0] else None} for k, v in tot.items()}
    print(json.dumps({'summary_x3d_m_b16': summary, 'hbm_peak_gbs': peaks['hbm_gbs']}), flush=True)
    if a.json:
        with open(a.json, 'w') as f:
            json.dump({'layers': rows, 'summary': summary, 'hbm_peak_gbs': peaks['hbm_gbs']}, f, indent=1)


if __name__ == '__main__':
    main()
