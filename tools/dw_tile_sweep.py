#!/usr/bin/env python
"""Sweep the (TH, TW, CC) tile of the tiled depthwise kernels per X3D-M layer shape (X3D_DW_FORCE knob) and print the
best tile per (layer, kernel) next to the planner's own choice.  One subprocess per tile (the knob is read per call but
kernel attributes are cached per instantiation, so a fresh process keeps runs independent).

  python tools/dw_tile_sweep.py [--layers l1.x,l2.x,l3.x,l4.x] [--json out.json]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(layer, force):
    env = dict(os.environ)
    if force:
        env['X3D_DW_FORCE'] = force
    else:
        env.pop('X3D_DW_FORCE', None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'dw_microbench.py'), '--layer', layer, '--iters', '10'],
                         env=env, capture_output=True, text=True, timeout=120)
    for line in out.stdout.splitlines():
        try:
            d = json.loads(line)
        except Exception:
            continue
        if d.get('layer') == layer:
            return {k: d[k]['us'] for k in ('fwd', 'dgrad', 'wgrad')}
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--layers', default='l1.x,l2.x,l3.x,l4.x,l1.0,l2.0,l3.0,l4.0')
    ap.add_argument('--json', default='')
    a = ap.parse_args()
    res = {}
    for layer in a.layers.split(','):
        base = run(layer, None)
        rows = {'planner': base}
        for CC in (48, 56, 72):
            for TW in (4, 8, 16):
                for TH in (2, 4, 6, 8):
                    threads = (TH // 2) * (TW // 2) * (CC // 2)
                    if threads < 64 or threads > 256:
                        continue
                    r = run(layer, f'{TH},{TW},{CC}')
                    if r:
                        rows[f'{TH},{TW},{CC}'] = r
        res[layer] = rows
        best = {k: min(((v[k], t) for t, v in rows.items() if v), key=lambda z: z[0]) for k in ('fwd', 'dgrad', 'wgrad')}
        print(layer, 'planner', {k: round(v, 1) for k, v in base.items()}, 'best',
              {k: (round(v[0], 1), v[1]) for k, v in best.items()}, flush=True)
    if a.json:
        with open(a.json, 'w') as f:
            json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()
