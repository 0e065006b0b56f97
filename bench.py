#!/usr/bin/env python
"""Benchmark of the X3D training hot path (BASELINE.json metric: X3D-M train clips/sec; dwconv HBM GB/s).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, BASELINE config 2
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port, exact ATen calls)
  python bench.py --config multigrid                       # BASELINE config 3: every clip shape of the multigrid schedule
  python bench.py --config xl | charades                   # configs 4 (X3D-XL 16x312^2, batch 8) / 5 (157-way BCE, 1 split)

A "step" is one full training iteration on one batch of synthetic clips: forward, loss, backward (+ bucketed gradient
allreduce when N > 1) and the SGD(momentum, weight-decay) update -- the body of the reference loop,
train_x3d_kinetics_multigrid.py:244-279.  Workload at N=1 is BASELINE.json configs[1]: X3D-M, batch 16/GPU, 16x224x224
clips, bf16 storage, base_bn_splits=2.  Prints ONE JSON line (rank 0).
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
    ap.add_argument('--config', default='train', choices=['train', 'multigrid', 'xl', 'charades'],
                    help='train = BASELINE config 2 (default); multigrid = config 3; xl = config 4; charades = config 5')
    ap.add_argument('--version', default=None)
    ap.add_argument('--batch', type=int, default=None, help='clips per GPU')
    ap.add_argument('--frames', type=int, default=16)
    ap.add_argument('--crop', type=int, default=None)
    ap.add_argument('--classes', type=int, default=None)
    ap.add_argument('--bn-splits', type=int, default=None)
    ap.add_argument('--loss', default=None, choices=['ce', 'bce'])
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--ddp', default='graph_tail', choices=['graph_tail', 'graph_nccl', 'eager'],
                    help='N > 1: captured fwd+bwd followed by one NCCL allreduce of the flat gradient buffer + the one-kernel SGD '
                         '(default); whole step incl. the bucketed NCCL allreduces captured in the CUDA graph (measured equal at '
                         '2 GPUs: 12.88 vs 12.86 ms); or the bucketed allreduce issued eagerly from inside backward')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity', action='store_true', help='skip the golden-logits check of the benchmarked configuration')
    ap.add_argument('--no-gpu-eager', action='store_true', help='reference arm: skip the informational ATen-on-GPU timing')
    ap.add_argument('--no-graph', action='store_true', help='enqueue every kernel from Python instead of replaying a CUDA graph')
    ap.add_argument('--kernel-table', default='', help='write a per-kernel timing table (JSON) here')
    a = ap.parse_args()
    d = {'train': dict(version='M', batch=16, crop=224, classes=400, bn_splits=2, loss='ce'),
         'multigrid': dict(version='M', batch=16, crop=224, classes=400, bn_splits=2, loss='ce'),
         'xl': dict(version='XL', batch=8, crop=312, classes=400, bn_splits=2, loss='ce'),
         'charades': dict(version='M', batch=16, crop=224, classes=157, bn_splits=1, loss='bce')}[a.config]
    for k, v in d.items():
        if getattr(a, k) is None:
            setattr(a, k, v)
    return a


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
# reference arm / cpu baseline: the oracle port (exact ATen calls of x3d.py) on the host cores
# --------------------------------------------------------------------------------------------
def _labels(args, batch, torch):
    if args.loss == 'bce':
        g = torch.Generator().manual_seed(7)
        return (torch.rand(batch, args.classes, generator=g) < 0.05).double()       # Bernoulli(0.05) multi-hot
    return torch.arange(batch).unsqueeze(1) % args.classes


def cpu_reference_rate(args, steps, warmup, batch):
    """clips/s of the reference's CPU implementation on a bounded sample: `batch` clips of the benchmark's shape per
    step, SAME BN split count as the GPU arm, fp32, all host threads.  The oracle port issues the reference's exact
    ATen calls (conv3d, F.batch_norm on the (n/s, c*s) view + the two affine ops, the saved-x Swish backward)."""
    import torch
    from oracle import x3d_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    splits = args.bn_splits if batch % args.bn_splits == 0 else 1
    sd = {k: (v.float() if v.is_floating_point() else v)
          for k, v in O.make_state_dict(args.version, args.classes, splits).items()}
    x = O.det_clip((batch, 3, args.frames, args.crop, args.crop), dtype=torch.float32)
    labels = _labels(args, batch, torch)
    wd = 1e-5 if args.loss == 'bce' else 5e-5
    mom = {}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, loss, grads, stats = O.loss_and_grads(sd, x, labels, loss=args.loss, version=args.version, splits=splits,
                                                 training=True, conv_impl='aten', bn_impl='aten')
        with torch.no_grad():       # SGD(momentum .9, wd), as the reference loop does
            for k, g in grads.items():
                d = g + wd * sd[k]
                mom[k] = d if k not in mom else 0.9 * mom[k] + d
                sd[k] = sd[k] - 0.01 * mom[k]
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return batch * len(times) / total, total / len(times), cores, splits


def gpu_eager_rate(args, steps=5, warmup=2):
    """Informational: the incumbent -- what stock PyTorch eager launches for the reference's x3d.py on the same B200
    (the oracle port issuing the reference's exact ATen / cuDNN calls), full batch, fwd+bwd+SGD: fp32, bf16 autocast
    and channels_last_3d (base_bn_splits=1 only: the reference's (n/s, c*s) view does not exist in that layout).  Also
    the torch.profiler device time of ATen's conv_depthwise3d kernels with their algorithmic GB/s.  Never the product."""
    import torch
    from oracle import x3d_oracle as O
    if not torch.cuda.is_available():
        return None
    out = {}
    B = args.batch
    dev = torch.device('cuda', 0)
    x = torch.randn(B, 3, args.frames, args.crop, args.crop, device=dev)
    labels = _labels(args, B, torch).to(dev)
    wd = 1e-5 if args.loss == 'bce' else 5e-5

    def make_sd(splits):
        return {k: (v.float() if v.is_floating_point() else v).to(dev)
                for k, v in O.make_state_dict(args.version, args.classes, splits).items()}

    def one_step(sd, mom, xin, splits, ctx):
        with torch.autocast('cuda', dtype=ctx, enabled=ctx is not None):
            _, loss, grads, _ = O.loss_and_grads(sd, xin, labels, loss=args.loss, version=args.version, splits=splits,
                                                 training=True, conv_impl='aten', bn_impl='aten')
        with torch.no_grad():
            keys = list(grads)
            gs = [grads[k] for k in keys]
            ps = [sd[k] for k in keys]
            torch._foreach_add_(gs, ps, alpha=wd)
            if not mom:
                mom.update({k: g.clone() for k, g in zip(keys, gs)})
            else:
                ms = [mom[k] for k in keys]
                torch._foreach_mul_(ms, 0.9)
                torch._foreach_add_(ms, gs)
            torch._foreach_add_(ps, [mom[k] for k in keys], alpha=-0.01)

    arms = [('fp32', None, args.bn_splits, False), ('bf16_autocast', torch.bfloat16, args.bn_splits, False),
            ('fp32_channels_last_3d_splits1', None, 1, True), ('bf16_autocast_channels_last_3d_splits1', torch.bfloat16, 1, True)]
    for name, ctx, splits, cl in arms:
        try:
            sd, mom = make_sd(splits), {}
            if cl:
                sd = {k: (v.contiguous(memory_format=torch.channels_last_3d) if v.dim() == 5 else v) for k, v in sd.items()}
            xin = x.contiguous(memory_format=torch.channels_last_3d) if cl else x
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for it in range(warmup + steps):
                if it == warmup:
                    torch.cuda.synchronize()
                    e0.record()
                one_step(sd, mom, xin, splits, ctx)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {'clips_per_s': B * 1e3 / ms, 'ms_per_step': ms, 'bn_splits': splits}
            if name == 'fp32':
                # profiler table of ATen's depthwise kernels
                from torch.profiler import ProfilerActivity, profile
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    one_step(sd, mom, xin, splits, ctx)
                    torch.cuda.synchronize()
                dw = {}
                for ev in prof.key_averages():
                    if 'conv_depthwise3d' in ev.key:
                        kind = ('dgrad' if 'backward_input' in ev.key else 'wgrad' if 'backward_weight' in ev.key else 'fwd')
                        d = dw.setdefault(kind, {'calls': 0, 'ms': 0.0})
                        d['calls'] += ev.count
                        d['ms'] += getattr(ev, 'device_time_total', getattr(ev, 'cuda_time_total', 0.0)) / 1e3
                by = dw_algorithmic_bytes(args.version, B, args.frames, args.crop, 4)
                for kind, d in dw.items():
                    d['algorithmic_gbytes'] = by[kind] / 1e9
                    d['gbs'] = by[kind] / 1e6 / d['ms'] if d['ms'] > 0 else None
                out['aten_conv_depthwise3d_fp32'] = dw
            del sd, mom
            torch.cuda.empty_cache()
        except Exception as ex:       # the oracle is CPU test infrastructure; a device/dtype hiccup is not fatal here
            out[name] = {'error': f'{type(ex).__name__}: {ex}'[:200]}
    out['what'] = (f'oracle port of x3d.py issuing the reference\'s exact ATen/cuDNN calls (conv3d, F.batch_norm on the '
                   f'(n/s,c*s) view, saved-x Swish) eagerly on 1 GPU, batch {B}, fwd+bwd+SGD, CUDA-event timed')
    return out


def dw_algorithmic_bytes(version, B, T, crop, eb):
    """sum over the 3x3x3 depthwise layers + the 5x1x1 stem conv of B*T*C*(H*W + Ho*Wo)*eb (unpadded C)"""
    from x3d_multigrid_b200.x3d import get_blocks, get_inplanes
    planes, blocks = get_inplanes(version), get_blocks(version)
    H = (crop + 2 - 3) // 2 + 1
    fwd = B * T * planes[0][1] * (H * H * 2) * eb                  # conv1_t 5x1x1
    for (mid, _), nb in zip(planes, blocks):
        for i in range(nb):
            s = 2 if i == 0 else 1
            Ho = (H + 2 - 3) // s + 1
            fwd += B * T * mid * (H * H + Ho * Ho) * eb
            H = Ho
    return {'fwd': fwd, 'dgrad': fwd, 'wgrad': fwd}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = 2
    rate, sec, cores, splits = cpu_reference_rate(args, args.steps, args.warmup, batch)
    sample = (f'{batch} clips of 3x{args.frames}x{args.crop}x{args.crop} per step, base_bn_splits={splits}, fwd+bwd+SGD, fp32, '
              f'oracle port of x3d.py issuing its exact ATen/oneDNN calls, {cores} threads')
    line = {'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'X3D-{args.version} train step, {args.frames}x{args.crop}x{args.crop}, '
                                   f'{args.classes} classes, base_bn_splits={splits} (CPU sample: batch {batch})'},
            'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    if not args.no_gpu_eager:
        try:
            line['gpu_eager_torch'] = gpu_eager_rate(args)
        except Exception as ex:
            line['gpu_eager_torch'] = {'error': f'{type(ex).__name__}: {ex}'[:200]}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# algorithmic bytes / flops of the conv calls, from the positional args of the C-ABI call
# --------------------------------------------------------------------------------------------
def conv_cost(name, a, eb, unpad):
    """-> (class, algorithmic bytes with UNPADDED channels, bytes with padded channels, flops)"""
    if name.startswith('x3d_dwconv'):
        if name == 'x3d_dwconv_wgrad':          # x, dy, dw, N, T, H, W, C, Cp, kt, kh, kw, stride
            N, T, H, W, C, Cp, kt, kh, kw, s = a[3:13]
        else:                                   # x|dy, w, y|dx, N, T, H, W, Cp, kt, kh, kw, stride
            N, T, H, W, Cp, kt, kh, kw, s = a[3:12]
            C = unpad.get(Cp, Cp)
        Ho, Wo = (H + 2 * (kh // 2) - kh) // s + 1, (W + 2 * (kw // 2) - kw) // s + 1
        taps = kt * kh * kw
        io = N * T * (H * W + Ho * Wo) * eb
        cls = {'x3d_dwconv_fwd': 'dw_fwd', 'x3d_dwconv_dgrad': 'dw_dgrad', 'x3d_dwconv_wgrad': 'dw_wgrad'}[name]
        return cls, io * C + taps * C * 4, io * Cp + taps * Cp * 4, 2 * N * T * Ho * Wo * C * taps
    # pointwise: x, w, y, N, T, H, W, Kp, Np, stride  |  dy, wT, dx, N, T, H, W, Kp, Np, stride, acc  |  wgrad: ..., K, Kp, Nn, Np, stride
    if name in ('x3d_pwconv_wgrad', 'x3d_pwconv_wgrad_ws'):
        N, T, H, W, K, Kp, Nn, Np, s = a[3:12]
    else:
        N, T, H, W, Kp, Np, s = a[3:10]
        K, Nn = unpad.get(Kp, Kp), unpad.get(Np, Np)
    M = N * T * ((H - 1) // s + 1) * ((W - 1) // s + 1)
    cls = {'x3d_pwconv_fwd': 'pw_fwd', 'x3d_pwconv_dgrad': 'pw_dgrad', 'x3d_pwconv_wgrad': 'pw_wgrad',
           'x3d_pwconv_wgrad_ws': 'pw_wgrad'}[name]
    wbytes = 4 if cls == 'pw_wgrad' else eb
    return cls, M * (K + Nn) * eb + K * Nn * wbytes, M * (Kp + Np) * eb + Kp * Np * wbytes, 2 * M * K * Nn


CONV_CALLS = ('x3d_dwconv_fwd', 'x3d_dwconv_dgrad', 'x3d_dwconv_wgrad', 'x3d_pwconv_fwd', 'x3d_pwconv_dgrad',
              'x3d_pwconv_wgrad', 'x3d_pwconv_wgrad_ws')


def parity_check(args, torch, X, dev):
    """BASELINE config 2 only: the model under test on the golden clip vs the logits the UNMODIFIED reference produced in
    fp64 (tests/golden/m_config2.npz, oracle/make_golden.py).  The oracle package is used here as the checker only
    (deterministic weight / clip generators); nothing of it runs inside a timed region."""
    import numpy as np
    from oracle import x3d_oracle as O
    path = os.path.join(ROOT, 'tests', 'golden', 'm_config2.npz')
    if not os.path.exists(path):
        return None
    gold = np.load(path)
    m = X.generate_model('M', n_classes=400, base_bn_splits=2, dropout=0.0)
    sd = O.make_state_dict('M', 400, 2)
    m.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd.items()})
    m = m.to(dev).set_compute_dtype(torch.bfloat16 if args.dtype == 'bf16' else torch.float32).train()
    x = O.det_clip((16, 3, 16, 224, 224), dtype=torch.float32).to(dev)
    labels = torch.from_numpy(gold['labels']).to(dev)
    logits = m(x)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    ref = torch.from_numpy(gold['logits']).double()
    err = float((logits.detach().double().cpu() - ref).norm() / ref.norm())
    del m, x
    torch.cuda.empty_cache()
    return {'case': 'tests/golden/m_config2.npz (reference x3d.py, fp64, X3D-M B=16 16x224x224, 2 BN splits)',
            'logits_rel_l2_err': err, 'loss': float(loss), 'golden_loss': float(gold['loss']),
            'tolerance': 2e-2 if args.dtype == 'bf16' else 1e-4}


def teardown(dist, world, *graph_holders):
    """NCCL kernels captured in a CUDA graph pin the communicator: drop the graphs first, then leave the group.  The
    process exits right after; a rank that would block in communicator teardown must not keep the launcher alive."""
    import gc
    import torch
    for h in graph_holders:
        if isinstance(h, dict):
            h.clear()
    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)        # no collective teardown: the other ranks may already be gone (rank 0 prints last)


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import x3d_multigrid_b200 as X
    from x3d_multigrid_b200 import _lib
    from x3d_multigrid_b200 import multigrid as MG
    from x3d_multigrid_b200.graphs import GraphedTrainStep
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
    if world > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)
    wd = 1e-5 if args.loss == 'bce' else 5e-5
    opt = FusedSGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=wd, capturable=use_graph)
    crit = torch.nn.CrossEntropyLoss() if args.loss == 'ce' else torch.nn.BCEWithLogitsLoss()
    if args.loss == 'bce':
        crit_fn = lambda lg, y: crit(lg.squeeze(2), y)            # train_x3d_charades.py:162,176
    else:
        crit_fn = crit
    L = _lib.lib()

    # ---- data-parallel wiring ------------------------------------------------------------
    ddp_mode = 'single' if world == 1 else (args.ddp if use_graph else 'eager')
    if ddp_mode == 'eager':
        use_graph = False
    net, reduce_fn = model, None
    if ddp_mode == 'eager':
        net = DistributedX3D(model, broadcast_from=None)
    elif ddp_mode == 'graph_nccl':
        net = DistributedX3D(model, broadcast_from=None, side_stream=True, defer_scale=True)
        opt.grad_scale = 1.0 / world
    elif ddp_mode == 'graph_tail':
        opt.grad_scale = 1.0 / world
        reduce_fn = lambda flat: dist.all_reduce(flat)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == 'multigrid':
        return run_multigrid(args, torch, dist, model, net, opt, crit_fn, reduce_fn, MG, dev, world, rank, local, L, barrier)

    B, T, S = args.batch, args.frames, args.crop
    gen = torch.Generator().manual_seed(1234 + rank)
    n_host = 2
    host_x = [torch.randn(B, 3, T, S, S, generator=gen).pin_memory() for _ in range(n_host)]
    if args.loss == 'bce':
        host_y = [(torch.rand(B, args.classes, generator=gen) < 0.05).float().pin_memory() for _ in range(n_host)]
    else:
        host_y = [torch.randint(0, args.classes, (B, 1), generator=gen).pin_memory() for _ in range(n_host)]
    dev_x = [h.to(dev) for h in host_x]
    dev_y = [h.to(dev) for h in host_y]

    def eager_step(x, y):
        opt.zero_grad(set_to_none=True)       # a graph replay leaves .grad set: do not accumulate on top of it
        logits = net(x)
        loss = crit_fn(logits, y)
        loss.backward()
        if reduce_fn is not None:
            reduce_fn(model.engine().gflat)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    ddp_note = ddp_mode
    reduce_fn_keep = reduce_fn
    if use_graph:
        try:
            graphed = GraphedTrainStep(net, opt, crit_fn, dev_x[0], dev_y[0], reduce_fn=reduce_fn)
        except Exception as ex:
            if ddp_mode != 'graph_nccl':
                raise
            # NCCL work could not be captured on this stack: captured fwd+bwd, eager allreduce + SGD after the replay
            ddp_note = f'graph_tail (capture with NCCL failed: {type(ex).__name__}: {str(ex)[:120]})'
            torch.cuda.synchronize()
            model.engine().grad_hook = None
            model.engine().use_side = True
            net = model
            reduce_fn = lambda flat: dist.all_reduce(flat)
            reduce_fn_keep = reduce_fn
            graphed = GraphedTrainStep(net, opt, crit_fn, dev_x[0], dev_y[0], reduce_fn=reduce_fn)
        train_step = lambda x, y: graphed(x, y)
    else:
        train_step = eager_step

    # ---------------- device-resident arm -------------------------------------------------
    for i in range(max(args.warmup, 3)):
        train_step(dev_x[i % n_host], dev_y[i % n_host])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if not use_graph:
        L.prof_names, L.prof_records = set(CONV_CALLS), []
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
    prof_steps = args.steps
    if use_graph:
        # the timed region replays a CUDA graph (no per-kernel host calls): the SAME kernels on the same shapes are
        # timed here, in this process, with CUDA events around each conv call of 3 eager steps
        # (weight-gradient kernels back on the main stream, so that every kernel is timed running alone)
        L.prof_names, L.prof_records = set(CONV_CALLS), []
        l0 = L.launch_count()
        eng = model.engine()
        was_side, eng.use_side = eng.use_side, False
        prof_steps = 3
        for i in range(prof_steps):
            eager_step(dev_x[i % n_host], dev_y[i % n_host])
        torch.cuda.synchronize()
        eng.use_side = was_side
        launches = (L.launch_count() - l0) // prof_steps * args.steps       # kernels per step x replayed steps
        roofline_note = ('timed region = CUDA-graph replays; kernel times from per-call CUDA events over 3 eager steps '
                         'run right after it in the same process (all kernels on one stream)')
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

    # ---------------- end-to-end arm with the GPU-side input pipeline: pinned uint8 frames -> H2D -> fused stem ------
    # (decoded 256 x 256 frames, per-clip crop window + flip; crop / flip / ToTensor / Normalize run on the device)
    ms_u8, h2d_u8 = None, None
    if use_graph and args.config in ('train', 'charades') and S <= 256:
        Hs = 256
        hu = [torch.randint(0, 256, (B, T, Hs, Hs, 3), dtype=torch.uint8, generator=gen).pin_memory() for _ in range(n_host)]
        hc = [torch.stack([torch.randint(0, Hs - S + 1, (B,), generator=gen), torch.randint(0, Hs - S + 1, (B,), generator=gen),
                           torch.randint(0, 2, (B,), generator=gen), torch.zeros(B, dtype=torch.int64)], 1).int().pin_memory()
              for _ in range(n_host)]
        ex = X.UInt8Clips(hu[0].to(dev), hc[0].to(dev), S)
        g8 = GraphedTrainStep(net, opt, crit_fn, ex, dev_y[0], reduce_fn=reduce_fn_keep)
        st8 = [X.UInt8Clips(torch.empty_like(ex.frames), torch.empty_like(ex.crops), S) for _ in range(2)]

        def prefetch8(i):
            sl = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[sl])
                st8[sl].frames.copy_(hu[i % n_host], non_blocking=True)
                st8[sl].crops.copy_(hc[i % n_host], non_blocking=True)
                stage_y[sl].copy_(host_y[i % n_host], non_blocking=True)
                ready[sl].record(copy_stream)

        def loop8(n):
            for sl in range(2):
                freed[sl].record()
            prefetch8(0)
            tot = 0.0
            for i in range(n):
                if i + 1 < n:
                    prefetch8(i + 1)
                sl = i % 2
                torch.cuda.current_stream().wait_event(ready[sl])
                ls = g8(st8[sl], stage_y[sl])
                freed[sl].record()
                tot += float(ls.item())
            return tot

        loop8(2)
        barrier()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        loop8(args.steps)
        u1.record()
        barrier()
        ms_u8 = u0.elapsed_time(u1)
        h2d_u8 = hu[0].numel() + hc[0].numel() * 4 + host_y[0].numel() * host_y[0].element_size()

    if world > 1:
        tt = torch.tensor([ms, ms_e2e, ms_u8 if ms_u8 is not None else 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])
        if ms_u8 is not None:
            ms_u8 = float(tt[2])

    if rank != 0:
        return teardown(dist, world)
    reduce_fn = None                      # rank 0 is on its own from here on: no collectives below

    # ---------------- roofline: every conv class, dominant one reported -------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    tc_peak = float(peaks.get('bf16_tflops_sustained', 1350.0))
    peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
    eb = 2 if dtype == torch.bfloat16 else 4
    unpad = {}
    for mid, out_ in X.get_inplanes(args.version):
        for c in (mid, out_):
            unpad[(c + 7) // 8 * 8] = c
    agg = {}
    for name, a, ev0, ev1 in records:
        cls, by, by_pad, fl = conv_cost(name, a, eb, unpad)
        d = agg.setdefault(cls, {'ms': 0.0, 'bytes': 0.0, 'bytes_padded': 0.0, 'flops': 0.0, 'n': 0})
        d['ms'] += ev0.elapsed_time(ev1)
        d['bytes'] += by
        d['bytes_padded'] += by_pad
        d['flops'] += fl
        d['n'] += 1
    step_ms = ms / args.steps
    classes = {}
    for k, v in agg.items():
        gbs = v['bytes'] / 1e6 / v['ms'] if v['ms'] > 0 else None
        tfl = v['flops'] / 1e9 / v['ms'] if v['ms'] > 0 else None
        t_hbm, t_tc = v['bytes'] / (hbm_peak * 1e6), v['flops'] / (tc_peak * 1e9)        # ms at the two peaks
        classes[k] = {'launches_per_step': v['n'] / prof_steps, 'ms_per_step': v['ms'] / prof_steps,
                      'algorithmic_gbytes_per_step': v['bytes'] / 1e9 / prof_steps,
                      'padded_channel_gbytes_per_step': v['bytes_padded'] / 1e9 / prof_steps, 'gbs': gbs,
                      'hbm_frac': gbs / hbm_peak if gbs else None,
                      'tflops': tfl if k.startswith('pw') else None,
                      'tensor_frac': (tfl / tc_peak if tfl else None) if k.startswith('pw') else None,
                      'roofline_frac': (max(t_hbm, t_tc) / v['ms']) if v['ms'] > 0 else None,
                      'share_of_step': (v['ms'] / prof_steps) / step_ms}
    dom = max(agg, key=lambda k: agg[k]['ms']) if agg else None
    roofline = None
    traffic, traffic_src = None, None
    try:
        cand = sorted(f for f in os.listdir(os.path.join(ROOT, 'profiles')) if f.endswith('_traffic.json'))
        if cand and dom:
            tj = json.load(open(os.path.join(ROOT, 'profiles', cand[-1])))
            key = {'dw_fwd': 'x3d_dwconv_fwd', 'dw_dgrad': 'x3d_dwconv_dgrad', 'dw_wgrad': 'x3d_dwconv_wgrad',
                   'pw_fwd': 'x3d_pwconv_fwd', 'pw_dgrad': 'x3d_pwconv_dgrad', 'pw_wgrad': 'x3d_pwconv_wgrad'}[dom]
            tc = tj.get('calls', {}).get(key)
            if tc and args.config == 'train' and args.version == 'M' and B == 16 and T == 16 and S == 224 and dtype == torch.bfloat16:
                traffic = tc['dram_bytes_per_launch']
                traffic_src = f'profiles/{cand[-1]} (ncu dram__bytes_read+write, mean over the launches of {key})'
    except Exception:
        pass
    if dom:
        d = agg[dom]
        ach = d['bytes'] / 1e6 / d['ms']
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': ach / hbm_peak, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                    'launches': d['n'], 'avg_launch_us': 1e3 * d['ms'] / d['n'],
                    'algorithmic_bytes_per_launch': d['bytes'] / d['n'],
                    'algorithmic_bytes_formula': 'SURVEY.md 8d with UNPADDED channel counts (dw: B*T*C*(H*W+Ho*Wo)*eb + taps*C*4; '
                                                 'pw: M*(K+N)*eb + K*N*eb)',
                    'share_of_step': (d['ms'] / prof_steps) / step_ms,
                    'how': roofline_note, 'tensor_peak_tflops': tc_peak, 'classes': classes}
    if args.kernel_table and world == 1:
        # two extra (untimed) EAGER steps with EVERY C-ABI call bracketed by events: where the step goes
        L.prof_names, L.prof_records = set(L.fn), []
        eng = model.engine()
        was_side, eng.use_side = eng.use_side, False
        for i in range(2):
            eager_step(dev_x[i % n_host], dev_y[i % n_host])
        torch.cuda.synchronize()
        eng.use_side = was_side
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
            json.dump({'per_step_ms_sum': tot / 2, 'classes': classes, 'paths': L.path_counts(),
                       'all': dict(sorted(full.items(), key=lambda kv: -kv[1]['ms_total']))}, f, indent=1)

    parity = None
    if world == 1 and args.config == 'train' and not args.no_parity and args.version == 'M':
        try:
            parity = parity_check(args, torch, X, dev)
        except Exception as ex:
            parity = {'error': f'{type(ex).__name__}: {ex}'[:200]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, cores, splits = cpu_reference_rate(args, steps=2, warmup=1, batch=2)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': f'2 clips of 3x{T}x{S}x{S} per step, base_bn_splits={splits}, 1 warm-up + 2 timed fwd+bwd+SGD steps, '
                         f'fp32, oracle port of x3d.py issuing its exact ATen/oneDNN calls, {cores} threads'}

    clips = B * world * args.steps
    h2d = host_x[0].numel() * 4 + host_y[0].numel() * host_y[0].element_size()
    line = {
        'metric': METRIC, 'value': clips / (ms / 1e3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': f'X3D-{args.version} training step (fwd+{args.loss.upper()}+bwd+SGD), batch {B}/GPU, '
                               f'{T}x{S}x{S} clips, {args.classes} classes, base_bn_splits={args.bn_splits}',
                   'global_batch': B * world, 'parallelism': f'dp{world}', 'ddp_mode': ddp_note,
                   'l2_policy': 'inputs_exceed_l2 (clip batch 154 MB, activations > 1 GB vs 126 MB L2)'},
        'e2e': {'value': clips / (ms_e2e / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                'ms_per_step': ms_e2e / args.steps},
        'e2e_uint8_frames': None if ms_u8 is None else {
            'value': clips / (ms_u8 / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': h2d_u8, 'd2h_bytes_per_step': 4,
            'ms_per_step': ms_u8 / args.steps,
            'what': 'same step fed with pinned uint8 256x256 frames + per-clip crop window / flip: crop, flip, ToTensor(255) and '
                    'Normalize run on the device (input_pipeline.UInt8Clips), 3x fewer bytes over PCIe'},
        'gpu_launches': int(launches), 'cuda_graph': bool(use_graph), 'clocks': clocks, 'roofline': roofline,
        'cpu_baseline': cpu, 'final_loss': final_loss, 'parity': parity,
    }
    print(json.dumps(line), flush=True)
    teardown(dist, world)


def run_multigrid(args, torch, dist, model, net, opt, crit_fn, reduce_fn, MG, dev, world, rank, local, L, barrier):
    """BASELINE config 3: the clip shapes of the multigrid long/short-cycle schedule (kinetics_multigrid.py:205-237,
    cycle_batch_sampler.py:76-113; B x T x H x W held constant), driven through multigrid.iteration_plan and
    MultigridTrainer: one captured CUDA graph per shape, BN re-split and the LR law on every long-cycle change."""
    base_b, frames, crop = args.batch, args.frames, args.crop
    per_shape = args.steps
    # a schedule with one LR phase whose four quarters are the four long cycles, `k` iterations each
    k = 3 * (max(args.warmup, 3) + per_shape)
    plan = list(MG.iteration_plan(base_b, [0, 4 * k, 4 * k + 4], frames, crop, 4 * k))
    trainer = MG.MultigridTrainer(net, opt, crit_fn, use_graphs=not args.no_graph, reduce_fn=reduce_fn)
    gen = torch.Generator().manual_seed(1234 + rank)
    clips_cache = {}
    results = {}
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.launch_count()
    tot_ms, tot_clips = 0.0, 0
    seen = {}
    for it in plan:
        key = (it['long_index'], it['batch'], it['frames'], it['crop'])
        if key not in clips_cache:
            B, T, S = it['batch'], it['frames'], it['crop']
            x = torch.randn(B, 3, T, S, S, generator=gen).to(dev)
            y = torch.randint(0, args.classes, (B, 1), generator=gen).to(dev)
            clips_cache = {key: (x, y)} if len(clips_cache) > 3 else {**clips_cache, key: (x, y)}
        x, y = clips_cache[key]
        n = seen.get(key, 0)
        seen[key] = n + 1
        timed = n >= max(args.warmup, 3) and n < max(args.warmup, 3) + per_shape
        if timed:
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        trainer.step(x, y, it['long_index'])
        if timed:
            e1.record()
            barrier()
            msv = e0.elapsed_time(e1)
            if world > 1:
                tt = torch.tensor([msv], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                msv = float(tt[0])
            r = results.setdefault(key, {'ms': 0.0, 'steps': 0})
            r['ms'] += msv
            r['steps'] += 1
            tot_ms += msv
            tot_clips += it['batch'] * world
    launches = L.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return teardown(dist, world, trainer.graphs)
    shapes = []
    for (li, B, T, S), r in results.items():
        shapes.append({'long_index': li, 'batch_per_gpu': B, 'frames': T, 'crop': S, 'bn_splits': args.bn_splits * MG.LONG_CYCLE[li],
                       'ms_per_step': r['ms'] / r['steps'], 'clips_per_s': B * world * r['steps'] / (r['ms'] / 1e3)})
    steps = sum(r['steps'] for r in results.values())
    line = {'metric': METRIC, 'value': tot_clips / (tot_ms / 1e3), 'unit': UNIT, 'n_gpus': world, 'steps': steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': tot_ms / steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': f'X3D-{args.version} multigrid training (BASELINE config 3): {len(shapes)} clip shapes of the '
                                   f'long/short-cycle schedule, base batch {base_b}/GPU, T0={frames}, crop {crop}, base_bn_splits='
                                   f'{args.bn_splits}, one CUDA graph per shape, BN re-split + LR law per long cycle',
                       'parallelism': f'dp{world}', 'graphs_captured': len(trainer.graphs),
                       'value_is': 'clips of all timed steps / their summed device time (equal step count per shape)'},
            'shapes': shapes, 'gpu_launches': int(launches), 'cuda_graph': not args.no_graph, 'clocks': clocks,
            'e2e': None, 'roofline': None, 'cpu_baseline': None}
    print(json.dumps(line), flush=True)
    teardown(dist, world, trainer.graphs)


if __name__ == '__main__':
    main()
