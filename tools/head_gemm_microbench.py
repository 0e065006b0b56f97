#!/usr/bin/env python
"""Times the six fp32 head GEMMs (fc1 / fc2 forward, dgrad, wgrad of X3D-M, x3d.py:338-345) through x3d_small_gemm.

  python tools/head_gemm_microbench.py [--rows 16]

Each call is timed alone (CUDA events around 20 launches, a 256 MB buffer written between launches to evict L2)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from x3d_multigrid_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=16)
    a = ap.parse_args()
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    R, C5, F1, NC = a.rows, 432, 2048, 400
    pooled = torch.randn(R, C5, device='cuda')
    w1 = torch.randn(F1, C5, device='cuda')
    h1 = torch.randn(R, F1, device='cuda')
    w2 = torch.randn(NC, F1, device='cuda')
    b2 = torch.randn(NC, device='cuda')
    logits = torch.empty(R, NC, device='cuda')
    dl = torch.randn(R, NC, device='cuda')
    dh = torch.empty(R, F1, device='cuda')
    dp = torch.empty(R, C5, device='cuda')
    g1 = torch.zeros(F1, C5, device='cuda')
    g2 = torch.zeros(NC, F1, device='cuda')
    p = lambda t: t.data_ptr()
    calls = {
        'fc1 fwd   [R,432]x[432,2048]': (p(pooled), C5, 1, p(w1), 1, C5, p(h1), F1, R, F1, C5, None, 1, None, 0, st),
        'fc2 fwd   [R,2048]x[2048,400]': (p(h1), F1, 1, p(w2), 1, F1, p(logits), NC, R, NC, F1, p(b2), 0, None, 0, st),
        'fc2 wgrad [400,R]x[R,2048]': (p(dl), 1, NC, p(h1), F1, 1, p(g2), F1, NC, F1, R, None, 0, None, 1, st),
        'fc2 dgrad [R,400]x[400,2048]': (p(dl), NC, 1, p(w2), F1, 1, p(dh), F1, R, F1, NC, None, 0, None, 0, st),
        'fc1 wgrad [2048,R]x[R,432]': (p(dh), 1, F1, p(pooled), C5, 1, p(g1), C5, F1, C5, R, None, 0, None, 1, st),
        'fc1 dgrad [R,2048]x[2048,432]': (p(dh), F1, 1, p(w1), C5, 1, p(dp), C5, R, C5, F1, None, 0, None, 0, st),
    }
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ws = torch.zeros(16 << 20, dtype=torch.uint8, device='cuda')

    def timed(fn_name, args):
        for _ in range(3):
            L.call(fn_name, *args)
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.call(fn_name, *args)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    tot = [0.0, 0.0]
    for name, args in calls.items():
        t0 = timed('x3d_small_gemm', args)
        t1 = timed('x3d_small_gemm_ws', args[:-1] + (ws.data_ptr(), ws.numel(), st))
        tot[0] += t0
        tot[1] += t1
        print(f'{name:34s} x3d_small_gemm {t0:7.1f} us   x3d_small_gemm_ws {t1:7.1f} us')
    print(f'total without workspace {tot[0]:.1f} us, with workspace (split-K where the dispatch picks it) {tot[1]:.1f} us   '
          f'(an empty launch + two event records cost ~6 us here)')


if __name__ == '__main__':
    main()
