#!/usr/bin/env python
"""Turn gpurun_out/<round>_* (written by tools/collect_profiles.sh) into the tracked files under profiles/."""
import collections
import csv
import gzip
import json
import os
import shutil
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else 'r02'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
os.makedirs(P, exist_ok=True)
for name in ('bench.json', 'bench_reference.json', 'kernel_table.json', 'dw_microbench.json', 'pw_microbench.json', 'smi.csv',
             'multigrid_shapes.jsonl', 'bench_n2.json', 'bench_n4.json', 'bench_n8.json', 'fma_rate.txt', 'bench_xl.json', 'bench_s.json',
             'bench_multigrid.json', 'bench_charades.json', 'sass_mnemonics.txt', 'bench_xl_n8.json', 'bench_multigrid_n8.json',
             'marginal_cost.json', 'stem_microbench.jsonl', 'pytest_gpu.log', 'head_gemm_microbench.txt', 'split_concurrency_probe.txt', 'ew_microbench.jsonl'):
    src = os.path.join(G, f'{R}_{name}')
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f'{R}_{name}'))

# ---- ncu launch list -> per-kernel shares (cold-cache, serialised: compare SHARES, not absolutes)
lc = os.path.join(G, f'{R}_launches.csv')
if os.path.exists(lc):
    lines = [l for l in open(lc) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = row['Kernel Name'].split('(')[0][:90]
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        v = float(row['Metric Value'].replace(',', ''))
        d[1] += v / 1e3 if row['Metric Unit'] == 'ns' else v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f'{R}_launches_summary.csv'), 'w') as f:
        f.write('kernel,launches,total_us,share\n')
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{n},{us:.1f},{us / tot:.4f}\n')
    with open(lc, 'rb') as fi, gzip.open(os.path.join(P, f'{R}_launches.csv.gz'), 'wb') as fo:
        shutil.copyfileobj(fi, fo)

# ---- DRAM traffic of the conv C-ABI calls: ncu dram counters per launch over one eager step
tc = os.path.join(G, f'{R}_conv_traffic.csv')
if os.path.exists(tc):
    import re
    lines = [l for l in open(tc) if not l.startswith('==')]
    per = collections.defaultdict(dict)
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(row['Metric Unit'], 1.0)
        per[row['ID']][row['Metric Name']] = v * scale
        per[row['ID']]['kernel'] = row['Kernel Name']
    calls = collections.defaultdict(list)
    extra = collections.defaultdict(list)          # second-stage kernels that belong to a call (counted into its bytes)
    for d in per.values():
        k = d['kernel']
        if 'dw3_wgrad' in k:
            calls['x3d_dwconv_wgrad'].append(d)
        elif 'pw_wgrad_reduce' in k:
            extra['x3d_pwconv_wgrad'].append(d)
        elif 'pw_wgrad' in k:
            calls['x3d_pwconv_wgrad'].append(d)
        elif 'pw_tc_kernel' in k:
            calls['x3d_pwconv_fwd' if '<1>' in k or 'true' in k else 'x3d_pwconv_dgrad'].append(d)
        else:
            m = re.search(r'dw3_tiled_kernel<[^,]+, (\d)', k)
            if m:
                calls['x3d_dwconv_fwd' if int(m.group(1)) < 2 else 'x3d_dwconv_dgrad'].append(d)
    out = {'how': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:dw3_|pw_ on bench.py --no-graph, one eager '
                  'step (tiled depthwise kernels and tcgen05 GEMMs; the second-stage reduce kernel of the pointwise wgrad is '
                  'counted into its call)', 'calls': {}}
    for call, ds in calls.items():
        n = len(ds)
        allk = ds + extra.get(call, [])
        rd = sum(d.get('dram__bytes_read.sum', 0.0) for d in allk)
        wr = sum(d.get('dram__bytes_write.sum', 0.0) for d in allk)
        us = sum(d.get('gpu__time_duration.sum', 0.0) for d in allk)
        out['calls'][call] = {'launches': n, 'dram_read_bytes_per_launch': rd / n, 'dram_write_bytes_per_launch': wr / n,
                              'dram_bytes_per_launch': (rd + wr) / n, 'avg_launch_us_under_ncu': us / n}
    with open(os.path.join(P, f'{R}_traffic.json'), 'w') as f:
        json.dump(out, f, indent=1)

# ---- ncu --set full captures -> small json of the metrics DESIGN.md / bench.py cite
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'smsp__inst_executed.sum',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']
for fn in sorted(os.listdir(G)):
    if not (fn.startswith(R) and fn.endswith('.ncu-rep')):
        continue
    raw = subprocess.run(['ncu', '-i', os.path.join(G, fn), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for d in data:
        rec = {'kernel': d[hdr.index('Kernel Name')][:120]}
        for w in want:
            if w in hdr:
                rec[w] = f'{d[hdr.index(w)]} {units[hdr.index(w)]}'.strip()
        stalls = {}
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
                try:
                    stalls[h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')] = round(float(d[i]), 3)
                except ValueError:
                    pass
        rec['top_stalls'] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        out.append(rec)
    with open(os.path.join(P, fn.replace('.ncu-rep', '_ncu.json')), 'w') as f:
        json.dump(out, f, indent=1)
print(sorted(os.listdir(P)))
