"""Worker of tests/test_parallel_gpu.py: one NCCL rank of a 2-GPU data-parallel training step.

Every rank builds the same X3D-S (deterministic weights), takes its shard of the global batch and runs ONE
training step (forward, CE, backward with the bucketed gradient allreduce, SGD) in the requested mode:

  eager        parallel.DistributedX3D: bucket allreduces launched from inside the backward pass
  graph        the whole step -- NCCL kernels and the SGD kernel included -- captured as one CUDA graph
  graph_tail   forward+backward captured, one allreduce of the flat buffer + SGD run eagerly after the replay

and saves the allreduced gradients, the updated parameters and its BN running statistics for the parent."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CROP = 96      # stage-4 BN groups then hold 2 x 4 x 3 x 3 = 72 values (tiny groups make the gradients pure noise)


def shard_rows(rank, world, per_rank, local_splits):
    """global positions of rank r's samples in the single-process batch with S = s*G splits:
    rank-r sample b + k*s sits at global position (r*s + b) + k*S"""
    s, S = local_splits, local_splits * world
    return [(rank * s + (i % s)) + (i // s) * S for i in range(per_rank)]


def main(rank, world, port, mode, outdir, per_rank=4, local_splits=2, dtype='fp32'):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    import x3d_multigrid_b200 as X
    from oracle import x3d_oracle as O
    from x3d_multigrid_b200.graphs import GraphedTrainStep
    from x3d_multigrid_b200.optim import FusedSGD
    from x3d_multigrid_b200.parallel import DistributedX3D

    ncls = 7
    m = X.generate_model('S', n_classes=ncls, base_bn_splits=local_splits, dropout=0.0)
    sd = O.make_state_dict('S', ncls, local_splits)
    m.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd.items()})
    m = m.to(dev).set_compute_dtype(torch.float32 if dtype == 'fp32' else torch.bfloat16).train()
    B = per_rank * world
    xg = O.det_clip((B, 3, 4, CROP, CROP), 'dpclip', torch.float32)
    yg = (torch.arange(B) * 3 % ncls).view(B, 1)
    rows = shard_rows(rank, world, per_rank, local_splits)
    x, y = xg[rows].to(dev), yg[rows].to(dev)
    crit = torch.nn.CrossEntropyLoss()
    graph = mode != 'eager'
    opt = FusedSGD(m.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-5, capturable=graph)
    if mode == 'eager':
        net = DistributedX3D(m)
        loss = crit(net(x), y)
        loss.backward()
        grads = m.engine().gflat.clone()
        opt.step()
    elif mode == 'graph':
        net = DistributedX3D(m, side_stream=True, defer_scale=True)
        opt.grad_scale = 1.0 / world
        step = GraphedTrainStep(net, opt, crit, x, y)          # preserve_state: warm-up steps are undone
        loss = step(x, y)
        grads = step.flat.clone() / world
    else:
        opt.grad_scale = 1.0 / world
        step = GraphedTrainStep(m, opt, crit, x, y, reduce_fn=lambda flat: dist.all_reduce(flat))
        loss = step(x, y)
        grads = step.flat.clone() / world
    torch.cuda.synchronize()
    eng = m.engine()
    out = {'loss': float(loss), 'gflat': grads.cpu(),
           'goff': {k: (r.goff, r.numel) for k, r in eng.refs.items()},
           'params': {k: p.detach().cpu() for k, p in m.named_parameters()},
           'stats': {k: v.detach().cpu() for k, v in m.named_buffers() if 'split_bn.running' in k}}
    torch.save(out, os.path.join(outdir, f'{mode}_{rank}.pt'))
    # CUDA graphs that captured NCCL kernels pin the communicator: leave without a collective teardown
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == '__main__':
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5])
