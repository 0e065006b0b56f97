#!/usr/bin/env python
"""Whole-network bf16 noise floor of the REFERENCE itself: the unmodified /root/reference/x3d.py under
torch.autocast(cpu, bfloat16) against its own fp64 golden (tests/golden/*.npz) on the golden cases.

Build-container tool (needs /root/reference).  Output: tests/golden/bf16_network_noise.json -- the GPU tests grade the
bf16 logits of the CUDA path against max(2e-2, this number) and print both."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get('X3D_REFERENCE', '/root/reference'))
from oracle import x3d_oracle as O  # noqa: E402
from oracle.make_golden import CASES, case_clip, case_labels, case_loss  # noqa: E402
import x3d as ref  # noqa: E402


def run(c, x, labels, splits, rows=None):
    torch.manual_seed(0)
    m = ref.generate_model(c['version'], n_classes=c['n_classes'], dropout=0.0, base_bn_splits=splits, task=c['task'])
    sd = O.det_fill_state_dict(m.state_dict())
    m.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd.items()})
    m.train()
    xs = (x if rows is None else x[rows]).float()
    ls = labels if rows is None else labels[rows]
    with torch.no_grad(), torch.autocast('cpu', dtype=torch.bfloat16):
        logits = m(xs)
    return logits.double(), case_loss(c, logits.double(), ls)


def main():
    out = {}
    for name in sys.argv[1:] or ['s_config1', 'm_mg_111', 'xl_small', 'm_charades_cls', 'm_config2']:
        c = CASES[name]
        gold = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
        x, labels = case_clip(c['shape']), case_labels(c)
        if c.get('by_split'):
            B, s = c['shape'][0], c['splits']
            logits = torch.zeros(tuple(gold['logits'].shape), dtype=torch.float64)
            loss = 0.0
            for b in range(s):
                rows = torch.arange(b, B, s)
                lg, ls = run(c, x, labels, 1, rows)
                logits[rows] = lg
                loss += float(ls) / s
        else:
            logits, loss = run(c, x, labels, c['splits'])
            loss = float(loss)
        refl = torch.from_numpy(gold['logits']).double()
        e = float((logits - refl).norm() / refl.norm())
        out[name] = {'autocast_bf16_logits_rel_l2': e, 'autocast_bf16_loss': loss, 'golden_loss': float(gold['loss'])}
        print(name, out[name], flush=True)
        path = os.path.join(ROOT, 'tests', 'golden', 'bf16_network_noise.json')
        with open(path, 'w') as f:
            json.dump({'what': 'reference x3d.py under torch.autocast(cpu, bf16), train-mode forward, vs its fp64 golden',
                       'cases': out}, f, indent=1)


if __name__ == '__main__':
    main()
