"""World-size-2 gloo test of the data-parallel host logic (bucketed allreduce-mean)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from x3d_multigrid_b200.parallel import BucketReducer
    ranges = [(0, 40), (40, 40), (40, 100), (100, 128)]     # includes an empty bucket
    flat = torch.arange(128, dtype=torch.float32) * (rank + 1)
    red = BucketReducer(ranges)
    for b in range(len(ranges)):
        red.launch(flat, b)
    red.finish(flat)
    want = torch.arange(128, dtype=torch.float32) * (1 + 2) / 2
    ok = torch.allclose(flat, want)
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_bucket_reducer_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_engine_bucket_layout_cpu():
    """bucket ranges tile the flat gradient buffer in reverse-stage order (no CUDA needed)"""
    import x3d_multigrid_b200 as X
    from x3d_multigrid_b200.engine import Engine
    m = X.generate_model('S', n_classes=10, base_bn_splits=1)
    e = Engine(m, torch.float32)
    # prepare() needs device memory only for the buffers; exercise the layout logic on CPU tensors
    e.lib = None
    try:
        e.prepare(torch.device('cpu'))
    except Exception as ex:  # pack table upload may fail on CPU-only builds of torch; layout is done before
        pytest.skip(f'prepare needs CUDA here: {ex}')
    lo = 0
    for a, b in e.bucket_ranges:
        assert a == lo and b >= a
        lo = b
    assert lo == e.gflat_numel
    first = e.refs['fc2.weight'].goff
    assert e.bucket_ranges[0][0] <= first < e.bucket_ranges[0][1]
    assert e.bucket_ranges[3][0] <= e.refs['conv1_s.weight'].goff < e.bucket_ranges[3][1]
    assert sum(r.numel for r in e.param_order) == sum(p.numel() for p in m.parameters())
