#!/usr/bin/env python
"""Feasibility probe: the BN splits of a batch are independent sub-batches (x3d.py:50, split = n % s).  Would running them
as concurrent half-size passes on two streams hide the launch / dependency latency of the ~700-kernel chain?

Times (a) one graph of the full step at B=16, 2 splits, (b) one graph at B=8, 1 split, (c) two independent B=8 graphs
replayed concurrently on two streams.  (c) < (a) would make the restructuring worth building."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import x3d_multigrid_b200 as X  # noqa: E402
from x3d_multigrid_b200.graphs import GraphedTrainStep  # noqa: E402
from x3d_multigrid_b200.optim import FusedSGD  # noqa: E402


def make(B, splits, seed):
    torch.manual_seed(seed)
    m = X.generate_model('M', n_classes=400, base_bn_splits=splits, dropout=0.5).cuda()
    m = m.set_compute_dtype(torch.bfloat16).train()
    opt = FusedSGD(m.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-5, capturable=True)
    x = torch.randn(B, 3, 16, 224, 224, device='cuda')
    y = torch.randint(0, 400, (B, 1), device='cuda')
    return GraphedTrainStep(m, opt, torch.nn.CrossEntropyLoss(), x, y)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    full = make(16, 2, 0)
    print('B=16, 2 splits, one graph      : %.3f ms' % timed(lambda: full.replay(clone_loss=False)))
    a, b = make(8, 1, 1), make(8, 1, 2)
    print('B=8, 1 split, one graph        : %.3f ms' % timed(lambda: a.replay(clone_loss=False)))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    cur = torch.cuda.current_stream()

    def both():
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            a.replay(clone_loss=False)
        with torch.cuda.stream(s2):
            b.replay(clone_loss=False)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    print('2 x (B=8, 1 split) concurrently: %.3f ms' % timed(both))

    def serial():
        a.replay(clone_loss=False)
        b.replay(clone_loss=False)

    print('2 x (B=8, 1 split) back to back: %.3f ms' % timed(serial))


if __name__ == '__main__':
    main()
