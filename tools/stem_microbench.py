#!/usr/bin/env python
"""Times the stem kernels of X3D-M at B=16, 16x224x224 (x3d.py:196-208,317-320) through the C ABI: conv1_s forward /
weight gradient, conv1_t (5x1x1 depthwise) forward / dgrad / wgrad.  GB/s = algorithmic bytes / CUDA-event time, buffers
rotate so that successive launches miss the 126 MB L2.

  python tools/stem_microbench.py [--batch 16] [--frames 16] [--crop 224]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from x3d_multigrid_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--frames', type=int, default=16)
    ap.add_argument('--crop', type=int, default=224)
    ap.add_argument('--iters', type=int, default=10)
    a = ap.parse_args()
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    N, T, S = a.batch, a.frames, a.crop
    H1 = (S + 2 - 3) // 2 + 1
    C0, C0p = 24, 24
    nb = 3
    xs = [torch.randn(N, 3, T, S, S, device='cuda') for _ in range(nb)]
    acts = [torch.randn(N, T, H1, H1, C0p, device='cuda').bfloat16() for _ in range(nb)]
    outs = [torch.empty(N, T, H1, H1, C0p, device='cuda', dtype=torch.bfloat16) for _ in range(nb)]
    ws = torch.randn(C0, 3, 1, 3, 3, device='cuda') * 0.2
    wt = torch.randn(5, C0p, device='cuda') * 0.3          # packed [kt][Cp]
    gws = torch.zeros(C0, 3, 1, 3, 3, device='cuda')
    gwt = torch.zeros(C0, 1, 5, 1, 1, device='cuda')
    stats = torch.zeros(N, C0p, 2, dtype=torch.float64, device='cuda')
    clip_b, act_b = xs[0].numel() * 4, acts[0].numel() * 2
    p = lambda t: t.data_ptr()
    cases = {
        'conv1_s fwd': (clip_b + act_b, lambda i: L.call('x3d_stem_conv_s_fwd', p(xs[i]), p(ws), p(outs[i]), N, 3, T, S, S, C0,
                                                           C0p, 1, st)),
        'conv1_s wgrad': (clip_b + act_b, lambda i: L.call('x3d_stem_conv_s_wgrad', p(xs[i]), p(acts[i]), p(gws), N, 3, T, S, S,
                                                             C0, C0p, 1, st)),
        'conv1_t fwd (+bn1 statistics)': (2 * act_b, lambda i: L.call('x3d_dwconv_fwd', p(acts[i]), p(wt), p(outs[i]), N, T, H1,
                                                                       H1, C0p, 5, 1, 1, 1, None, None, 1, 0, p(stats), 1, st)),
        'conv1_t dgrad': (2 * act_b, lambda i: L.call('x3d_dwconv_dgrad', p(acts[i]), p(wt), p(outs[i]), N, T, H1, H1, C0p, 5, 1,
                                                       1, 1, None, None, None, 1, None, 1, st)),
        'conv1_t wgrad': (2 * act_b, lambda i: L.call('x3d_dwconv_wgrad', p(acts[i]), p(outs[(i + 1) % nb]), p(gwt), N, T, H1,
                                                       H1, C0, C0p, 5, 1, 1, 1, None, None, 1, 0, 1, st)),
    }
    peak = 6547.2
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peak = json.load(open(pk)).get('hbm_gbs', peak)
    tot = 0.0
    for name, (nbytes, fn) in cases.items():
        for i in range(3):
            fn(i % nb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.iters):
            fn(i % nb)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.iters
        tot += us
        print(json.dumps({'kernel': name, 'us': round(us, 1), 'mb': round(nbytes / 1e6, 1), 'gbs': round(nbytes / us / 1e3),
                          'frac_hbm': round(nbytes / us / 1e3 / peak, 3)}), flush=True)
    print(json.dumps({'total_us': round(tot, 1)}))


if __name__ == '__main__':
    main()
