#!/usr/bin/env python
"""Benchmark of the X3D training hot path (BASELINE.json metric: X3D-M train clips/sec; dwconv HBM GB/s).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

A "step" is one full training iteration on one batch of synthetic clips: forward, cross-entropy,
backward (+ bucketed gradient allreduce when N > 1) and the SGD(momentum, weight-decay) update --
the body of the reference loop, train_x3d_kinetics_multigrid.py:244-279.  Workload at N=1 is
BASELINE.json configs[1]: X3D-M, batch 16/GPU, 16x224x224 clips, bf16 storage, base_bn_splits=2.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'x3d_m_train_clips_per_sec'
UNIT = 'clips/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--version', default='M')
    ap.add_argument('--batch', type=int, default=16, help='clips per GPU')
    ap.add_argument('--frames', type=int, default=16)
    ap.add_argument('--crop', type=int, default=224)
    ap.add_argument('--classes', type=int, default=400)
    ap.add_argument('--bn-splits', type=int, default=2)
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-eager', action='store_true', help='reference arm: skip the informational ATen-on-GPU timing')
    ap.add_argument('--no-graph', action='store_true', help='enqueue every kernel from Python instead of replaying a CUDA graph')
    ap.add_argument('--kernel-table', default='', help='write a per-kernel timing table (JSON) here')
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------
def cpu_reference_rate(args, steps, warmup, batch):
    """clips/s of the reference's CPU implementation (oracle port, ATen/oneDNN convs, fp32, all host
    threads) on a bounded sample: `batch` clips of the benchmark's shape per step."""
    import torch
    from oracle import x3d_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {k: (v.float() if v.is_floating_point() else v)
          for k, v in O.make_state_dict(args.version, args.classes, 1).items()}
    x = O.det_clip((batch, 3, args.frames, args.crop, args.crop), dtype=torch.float32)
    labels = torch.arange(batch).unsqueeze(1) % args.classes
    mom = {}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, loss, grads, stats = O.loss_and_grads(sd, x, labels, version=args.version, splits=1, training=True,
                                                 conv_impl='aten')
        with torch.no_grad():       # SGD(momentum .9, wd 5e-5), as the reference loop does
            for k, g in grads.items():
                d = g + 5e-5 * sd[k]
                mom[k] = d if k not in mom else 0.9 * mom[k] + d
                sd[k] = sd[k] - 0.01 * mom[k]
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return batch * len(times) / total, total / len(times), cores


def gpu_eager_rate(args, steps=5, warmup=2):
    """Informational: the same oracle port (plain ATen / cuDNN calls, i.e. what stock PyTorch eager launches for the
    reference's x3d.py) on cuda:0 at the benchmark's full batch, fp32 and bf16 autocast.  Never the product path."""
    import torch
    from oracle import x3d_oracle as O
    if not torch.cuda.is_available():
        return None
    out = {}
    B = args.batch
    dev = torch.device('cuda', 0)
    sd = {k: (v.float() if v.is_floating_point() else v).to(dev)
          for k, v in O.make_state_dict(args.version, args.classes, args.bn_splits).items()}
    x = torch.randn(B, 3, args.frames, args.crop, args.crop, device=dev)
    labels = (torch.arange(B, device=dev).unsqueeze(1) % args.classes)
    for name, ctx in (('fp32', None), ('bf16_autocast', torch.bfloat16)):
        try:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for it in range(warmup + steps):
                if it == warmup:
                    torch.cuda.synchronize()
                    e0.record()
                with torch.autocast('cuda', dtype=ctx, enabled=ctx is not None):
                    O.loss_and_grads(sd, x, labels, version=args.version, splits=args.bn_splits, training=True,
                                     conv_impl='aten')
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {'clips_per_s': B * 1e3 / ms, 'ms_per_step': ms}
        except Exception as ex:       # the oracle is CPU test infrastructure; a device/dtype hiccup is not fatal here
            out[name] = {'error': f'{type(ex).__name__}: {ex}'[:200]}
    out['what'] = (f'oracle port of x3d.py as stock ATen/cuDNN eager calls on 1 GPU, batch {B}, fwd+bwd (no optimizer), '
                   f'CUDA-event timed')
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = 2 if args.steps + args.warmup <= 8 else 1
    rate, sec, cores = cpu_reference_rate(args, args.steps, args.warmup, batch)
    sample = (f'{batch} clip(s) of 3x{args.frames}x{args.crop}x{args.crop} per step, fwd+bwd+SGD, fp32, '
              f'oracle port of x3d.py (ATen/oneDNN), {cores} threads')
    line = {'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'X3D-{args.version} train step, {args.frames}x{args.crop}x{args.crop}, '
                                   f'{args.classes} classes (CPU sample: batch {batch})'},
            'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    if not args.no_gpu_eager:
        try:
            line['gpu_eager_torch'] = gpu_eager_rate(args)
        except Exception as ex:
            line['gpu_eager_torch'] = {'error': f'{type(ex).__name__}: {ex}'[:200]}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# algorithmic bytes of the depthwise kernels (SURVEY.md 8d / BASELINE.md section 3)
# --------------------------------------------------------------------------------------------
def dw_bytes(name, a, eb):
    """a = positional args of the C-ABI call.  Unpadded channel counts are not visible at this level;
    padded Cp is used for dgrad/fwd (<= 3.7% more than algorithmic) -- see DESIGN.md."""
    if name == 'x3d_dwconv_fwd':
        N, T, H, W, Cp, kt, kh, kw, s = a[3:12]
    elif name == 'x3d_dwconv_dgrad':
        N, T, H, W, Cp, kt, kh, kw, s = a[3:12]
    else:  # wgrad: x, dy, dw, N, T, H, W, C, Cp, kt, kh, kw, stride
        N, T, H, W, _, Cp, kt, kh, kw, s = a[3:13]
    Ho, Wo = (H + 2 * (kh // 2) - kh) // s + 1, (W + 2 * (kw // 2) - kw) // s + 1
    taps = kt * kh * kw
    io = N * T * Cp * (H * W + Ho * Wo) * eb
    if name == 'x3d_dwconv_dgrad' and a[12] is not None:
        io += N * T * Cp * H * W * eb          # fused relu-mask/BN epilogue also reads the saved conv1 output
    return io + taps * Cp * 4


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import x3d_multigrid_b200 as X
    from x3d_multigrid_b200 import _lib
    from x3d_multigrid_b200.optim import FusedSGD
    from x3d_multigrid_b200.parallel import DistributedX3D

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback for the product path)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus or world == 1, f'--gpus {args.gpus} but WORLD_SIZE={world}'

    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    torch.manual_seed(0)
    model = X.generate_model(args.version, n_classes=args.classes, base_bn_splits=args.bn_splits, dropout=0.5)
    model = model.to(dev).set_compute_dtype(dtype).train()
    use_graph = not args.no_graph
    # eager: bucketed allreduce launched from inside backward (parallel.DistributedX3D); graph: forward+backward are
    # replayed as one CUDA graph, then ONE allreduce of the flat gradient buffer + the one-kernel SGD run eagerly
    net = DistributedX3D(model) if (world > 1 and not use_graph) else model
    if world > 1 and use_graph:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)
    opt = FusedSGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-5, capturable=use_graph and world == 1)
    if world > 1 and use_graph:
        opt.grad_scale = 1.0 / world
    crit = torch.nn.CrossEntropyLoss()

    B, T, S = args.batch, args.frames, args.crop
    gen = torch.Generator().manual_seed(1234 + rank)
    n_host = 2
    host_x = [torch.randn(B, 3, T, S, S, generator=gen).pin_memory() for _ in range(n_host)]
    host_y = [torch.randint(0, args.classes, (B, 1), generator=gen).pin_memory() for _ in range(n_host)]
    dev_x = [h.to(dev) for h in host_x]
    dev_y = [h.to(dev) for h in host_y]
    L = _lib.lib()

    def eager_step(x, y):
        logits = net(x)
        loss = crit(logits, y)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    if use_graph:
        # one CUDA graph for the whole step (forward, CE, backward, allreduce, SGD) -- x3d_multigrid_b200.graphs
        from x3d_multigrid_b200.graphs import GraphedTrainStep
        reduce_fn = (lambda: dist.all_reduce(model.engine().gflat)) if world > 1 else None
        graphed = GraphedTrainStep(net, opt, crit, dev_x[0], dev_y[0], reduce_fn=reduce_fn)
        train_step = graphed
    else:
        train_step = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm -------------------------------------------------
    for i in range(max(args.warmup, 3)):
        train_step(dev_x[i % n_host], dev_y[i % n_host])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    prof_names = ('x3d_dwconv_fwd', 'x3d_dwconv_dgrad', 'x3d_dwconv_wgrad')
    if not use_graph:
        L.prof_names, L.prof_records = set(prof_names), []
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss = train_step(dev_x[i % n_host], dev_y[i % n_host])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    final_loss = float(loss.item())
    roofline_note = 'per-call CUDA events inside the timed region'
    if use_graph:
        # the timed region replays a CUDA graph (no per-kernel host calls): the SAME kernels on the same shapes are
        # timed here, in this process, with CUDA events around each depthwise call of 3 eager steps
        # (weight-gradient kernels back on the main stream, so that every kernel is timed running alone)
        L.prof_names, L.prof_records = set(prof_names), []
        l0 = L.launch_count()
        side, model.engine().side = model.engine().side, None
        for i in range(3):
            eager_step(dev_x[i % n_host], dev_y[i % n_host])
        torch.cuda.synchronize()
        model.engine().side = side
        launches = (L.launch_count() - l0) // 3 * args.steps       # kernels per step x replayed steps
        roofline_note = ('timed region = CUDA-graph replays; kernel times from per-call CUDA events over 3 eager steps '
                         'run right after it in the same process')
    records, L.prof_names = L.prof_records, set()

    # ---------------- end-to-end arm: pinned host clips -> H2D each step, loss -> D2H ---------
    copy_stream = torch.cuda.Stream(dev)
    stage_x = [torch.empty_like(dev_x[0]) for _ in range(2)]
    stage_y = [torch.empty_like(dev_y[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            stage_x[s].copy_(host_x[i % n_host], non_blocking=True)
            stage_y[s].copy_(host_y[i % n_host], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n):
        for s in range(2):
            freed[s].record()
        prefetch(0)
        tot = 0.0
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            s = i % 2
            torch.cuda.current_stream().wait_event(ready[s])
            ls = train_step(stage_x[s], stage_y[s])
            freed[s].record()
            tot += float(ls.item())          # D2H read of the step's loss
        return tot

    e2e_loop(2)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(args.steps)
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)

    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant depthwise kernel ------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
    eb = 2 if dtype == torch.bfloat16 else 4
    agg = {}
    for name, a, ev0, ev1 in records:
        d = agg.setdefault(name, {'ms': 0.0, 'bytes': 0.0, 'n': 0})
        d['ms'] += ev0.elapsed_time(ev1)
        d['bytes'] += dw_bytes(name, a, eb)
        d['n'] += 1
    table = {k: {'launches': v['n'], 'ms_total': v['ms'], 'gbytes': v['bytes'] / 1e9,
                 'gbs': v['bytes'] / 1e6 / v['ms'] if v['ms'] > 0 else None} for k, v in agg.items()}
    dom = max(agg, key=lambda k: agg[k]['ms']) if agg else None
    roofline = None
    # measured DRAM traffic per launch of the dominant call: committed ncu summary (profiles/rNN_traffic.json, written by
    # tools/summarize_profiles.py from an ncu dram-counter pass of this same command); null when it is for another kernel
    traffic, traffic_src = None, None
    try:
        cand = sorted(f for f in os.listdir(os.path.join(ROOT, 'profiles')) if f.endswith('_traffic.json'))
        if cand:
            tj = json.load(open(os.path.join(ROOT, 'profiles', cand[-1])))
            tc = tj.get('calls', {}).get(dom)
            if tc and args.version == 'M' and args.batch == 16 and args.frames == 16 and args.crop == 224 and dtype == torch.bfloat16:
                traffic = tc['dram_bytes_per_launch']
                traffic_src = f'profiles/{cand[-1]} (ncu dram__bytes_read+write, mean over the tiled launches of {dom})'
    except Exception:
        pass
    if dom:
        d = agg[dom]
        ach = d['bytes'] / 1e6 / d['ms']
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': ach / hbm_peak, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                    'launches': d['n'], 'avg_launch_us': 1e3 * d['ms'] / d['n'],
                    'algorithmic_bytes_per_launch': d['bytes'] / d['n'],
                    'share_of_step': (d['ms'] / (3 if use_graph else args.steps)) / (ms / args.steps),
                    'how': roofline_note, 'dw_kernels': table}
    if args.kernel_table:
        # two extra (untimed) EAGER steps with EVERY C-ABI call bracketed by events: where the step goes
        L.prof_names, L.prof_records = set(L.fn), []
        side, model.engine().side = model.engine().side, None
        for i in range(2):
            eager_step(dev_x[i % n_host], dev_y[i % n_host])
        torch.cuda.synchronize()
        model.engine().side = side
        full = {}
        for name, a, ev0, ev1 in L.prof_records:
            d = full.setdefault(name, {'launches': 0, 'ms_total': 0.0})
            d['launches'] += 1
            d['ms_total'] += ev0.elapsed_time(ev1)
        L.prof_names = set()
        tot = sum(v['ms_total'] for v in full.values())
        for v in full.values():
            v['share'] = v['ms_total'] / tot
            v['ms_per_step'] = v['ms_total'] / 2
        with open(args.kernel_table, 'w') as f:
            json.dump({'per_step_ms_sum': tot / 2, 'dw': table,
                       'all': dict(sorted(full.items(), key=lambda kv: -kv[1]['ms_total']))}, f, indent=1)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, cores = cpu_reference_rate(args, steps=2, warmup=1, batch=2)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': f'2 clips of 3x{T}x{S}x{S} per step, 1 warm-up + 2 timed fwd+bwd+SGD steps, fp32, oracle '
                         f'port of x3d.py (ATen/oneDNN convs), {cores} threads'}

    clips = B * world * args.steps
    h2d = host_x[0].numel() * 4 + host_y[0].numel() * 8
    line = {
        'metric': METRIC, 'value': clips / (ms / 1e3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': f'X3D-{args.version} training step (fwd+CE+bwd+SGD), batch {B}/GPU, '
                               f'{T}x{S}x{S} clips, {args.classes} classes, base_bn_splits={args.bn_splits}',
                   'global_batch': B * world, 'parallelism': f'dp{world}',
                   'l2_policy': 'inputs_exceed_l2 (clip batch 154 MB, activations > 1 GB vs 126 MB L2)'},
        'e2e': {'value': clips / (ms_e2e / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches), 'cuda_graph': bool(use_graph), 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
        'final_loss': final_loss,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
