#!/usr/bin/env python
"""In-graph marginal cost of each C-ABI entry point: capture the X3D-M training step (16 x 3x16x224x224, bf16) as a CUDA
graph with one entry point turned into a no-op, replay it, and report  full step - step without it.

  python tools/marginal_cost.py [--names x3d_bn_finalize,x3d_bn_bwd_finalize] [--json out.json]

A name prefixed with "2x:" is issued twice instead (skipping a kernel can zero the data flow behind it, which makes
later kernels skip their zero-valued statistics atomics; doubling keeps the data sane and prices one more launch in the
dependency chain).  This is the number that says what removing / fusing a kernel class could buy: per-call CUDA events of an eager pass
(bench.py --kernel-table) include launch gaps and ignore the overlap with the weight-gradient stream.  The numerical
results of the crippled graphs are garbage by construction -- nothing here is a bench value.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import x3d_multigrid_b200 as X  # noqa: E402
from x3d_multigrid_b200 import _lib  # noqa: E402
from x3d_multigrid_b200.graphs import GraphedTrainStep  # noqa: E402
from x3d_multigrid_b200.optim import FusedSGD  # noqa: E402

DEFAULT = ['x3d_pwconv_fwd', 'x3d_pwconv_dgrad', 'x3d_pwconv_wgrad_ws', 'x3d_dwconv_fwd', 'x3d_dwconv_dgrad',
           'x3d_dwconv_wgrad', 'x3d_bn_finalize', 'x3d_bn_bwd_finalize', 'x3d_bn_bwd_apply', 'x3d_bn_bwd_reduce',
           'x3d_bn_bwd_reduce_store', 'x3d_bn_act_fwd', 'x3d_swish_gate_fwd', 'x3d_swish_gate_bwd_reduce',
           'x3d_swish_gate_bwd_apply', 'x3d_se_fwd', 'x3d_se_bn_bwd', 'x3d_stem_conv_s_fwd', 'x3d_stem_conv_s_wgrad',
           'x3d_small_gemm', 'x3d_sgd_step_dev']


def timed_step(skip, steps, B=16, T=16, S=224, zero_input=False):
    L = _lib.lib()
    orig = L.call
    counts = {}

    def call(name, *args):
        if name in skip:
            counts[name] = counts.get(name, 0) + 1
            return
        if '2x:' + name in skip:             # issue the call twice: what one more launch in the chain costs, data kept sane
            counts[name] = counts.get(name, 0) + 1
            orig(name, *args)
        return orig(name, *args)

    L.call = call
    try:
        torch.manual_seed(0)
        model = X.generate_model('M', n_classes=400, base_bn_splits=2, dropout=0.5).cuda()
        model = model.set_compute_dtype(torch.bfloat16).train()
        opt = FusedSGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-5, capturable=True)
        crit = torch.nn.CrossEntropyLoss()
        g = torch.Generator().manual_seed(1)
        xs = [torch.randn(B, 3, T, S, S, generator=g).cuda() for _ in range(2)]
        ys = [torch.randint(0, 400, (B, 1), generator=g).cuda() for _ in range(2)]
        if zero_input:
            xs = [torch.zeros_like(x) for x in xs]
        step = GraphedTrainStep(model, opt, crit, xs[0], ys[0])
        for i in range(3):
            step(xs[i % 2], ys[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(xs[i % 2], ys[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    finally:
        L.call = orig
    del step, model, opt
    torch.cuda.empty_cache()
    return ms, sum(counts.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--names', default=','.join(DEFAULT))
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--json', default='')
    ap.add_argument('--zero-input', action='store_true', help='also time the full step on an all-zero clip (data-dependent cost)')
    a = ap.parse_args()
    full, _ = timed_step(set(), a.steps)
    full2, _ = timed_step(set(), a.steps)
    out = {'full_ms': round(full, 3), 'full_ms_repeat': round(full2, 3), 'marginal_ms': {}}
    print(json.dumps({'full_ms': full, 'repeat': full2}), flush=True)
    if a.zero_input:
        z, _ = timed_step(set(), a.steps, zero_input=True)
        out['zero_input_ms'] = round(z, 3)
        print(json.dumps({'zero_input_ms': z}), flush=True)
    for n in a.names.split(','):
        grp = set(n.split('+'))
        ms, k = timed_step(grp, a.steps)
        out['marginal_ms'][n] = {'step_without_ms': round(ms, 3), 'marginal_ms': round((ms - full) if n.startswith('2x:') else (full - ms), 3),
                                 'calls_per_capture': k}
        print(json.dumps({n: out['marginal_ms'][n]}), flush=True)
    if a.json:
        with open(a.json, 'w') as f:
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
