#!/usr/bin/env python
"""How well can ANY bf16 execution of one reference Bottleneck reproduce its fp64 gradients?

Runs the UNMODIFIED reference Bottleneck (/root/reference/x3d.py) on the three block configurations of
tests/test_network_gpu.py::test_bottleneck_standalone, once in fp64 (anchor) and once under
torch.autocast(bfloat16) on the CPU, and prints relative-L2 errors of the output, the input gradient and
every parameter gradient.  Build-container tool (needs /root/reference); its output is committed as
profiles/r02_bf16_block_noise.json and is what the bf16 block tolerances of the GPU tests are set against."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get('X3D_REFERENCE', '/root/reference'))
from oracle import x3d_oracle as O  # noqa: E402
import x3d as ref  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def run(cfg, shape, autocast):
    cin, planes, stride, index, ds = cfg
    splits = 2
    down = None
    if ds:
        down = torch.nn.Sequential(ref.conv1x1x1(cin, planes[1], stride),
                                   ref.SubBatchNorm3d(num_splits=splits, num_features=planes[1], affine=True))
    blk = ref.Bottleneck(cin, planes, stride=stride, downsample=down, index=index, base_bn_splits=splits)
    sd = O.det_fill_state_dict(blk.state_dict())
    x0 = torch.relu(O.det_clip(shape, 'blkx', torch.float32)).to(torch.bfloat16)
    if autocast:
        blk = blk.float()
        blk.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd.items()})
        x = x0.float().requires_grad_(True)
    else:
        blk = blk.double()
        blk.load_state_dict(sd)
        x = x0.double().requires_grad_(True)
    blk.train()
    with torch.autocast('cpu', dtype=torch.bfloat16, enabled=autocast):
        y = blk(x)
    dy = O.det_clip(tuple(y.shape), 'blkdy', torch.float32).to(torch.bfloat16).to(y.dtype)
    y.backward(dy)
    return y.detach(), x.grad, {k: p.grad for k, p in blk.named_parameters()}


def main():
    cfgs = [(24, (54, 24), 1, 1, False), (24, (54, 24), 1, 2, False), (24, (108, 48), 2, 0, True)]
    out = {}
    for shape in ((4, 0, 3, 9, 7), (8, 0, 8, 28, 28)):
        for cfg in cfgs:
            shp = (shape[0], cfg[0]) + shape[2:]
            y64, dx64, g64 = run(cfg, shp, False)
            ya, dxa, ga = run(cfg, shp, True)
            e = {'y': rel(ya, y64), 'dx': rel(dxa, dx64)}
            e.update({k: rel(ga[k], g64[k]) for k in g64})
            key = f'cfg{cfg} x{shp}'
            out[key] = e
            worst = max(v for k, v in e.items() if k not in ('y', 'dx'))
            print(f'{key}: y {e["y"]:.2e} dx {e["dx"]:.2e} worst param grad {worst:.2e}', flush=True)
    path = os.path.join(ROOT, 'profiles', 'r02_bf16_block_noise.json')
    with open(path, 'w') as f:
        json.dump({'what': 'reference Bottleneck under torch.autocast(cpu, bf16) vs its own fp64 run; relative L2',
                   'cases': out}, f, indent=1)


if __name__ == '__main__':
    main()
