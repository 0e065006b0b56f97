"""Data-parallel training of X3D: one process per GPU, NCCL allreduce of the flat gradient
buffer in four reverse-stage buckets, each launched as soon as its last wgrad kernel has been
enqueued so that it overlaps with the backward pass of the earlier stages.

Replaces the reference's single-process ``nn.DataParallel`` (train_x3d_kinetics_multigrid.py:
175-177): same semantics -- the loss is a mean over the global batch (gradients are AVERAGED
over ranks), BatchNorm statistics stay per replica (no SyncBN), buffers are not reduced.
The path shards by batch; the only exchange step is this allreduce.

Works with any torch.distributed backend: ``nccl`` on GPUs (NVLink/NVSwitch), ``gloo`` for the
CPU tests of the host logic (tests/test_parallel_cpu.py drive ``BucketReducer`` directly)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


class BucketReducer:
    """Asynchronous bucketed allreduce-mean over slices of one flat buffer."""

    def __init__(self, bucket_ranges: List[Tuple[int, int]], process_group=None):
        self.ranges = list(bucket_ranges)
        self.group = process_group
        self.works: List = []
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1

    def launch(self, flat: torch.Tensor, bucket: int):
        if self.world == 1:
            return
        lo, hi = self.ranges[bucket]
        if hi > lo:
            self.works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self, flat: torch.Tensor, scale: bool = True):
        """wait for every bucket, then scale to the mean (``scale=False``: the optimizer applies 1/world)"""
        if self.world == 1:
            return
        for w in self.works:
            w.wait()
        self.works = []
        if scale:
            flat.mul_(1.0 / self.world)


class DistributedX3D(torch.nn.Module):
    """``DistributedX3D(model)`` -- drop-in for ``nn.DataParallel(model)`` in the reference loop
    (keeps the ``.module`` attribute the scripts use, train_x3d_kinetics_multigrid.py:205,228,287)."""

    def __init__(self, module, process_group=None, broadcast_from: Optional[int] = 0, side_stream: bool = False,
                 defer_scale: bool = False):
        """``side_stream``: keep the weight-gradient kernels on the engine's second stream while the bucket
        allreduces are in flight (used by the CUDA-graph step, where the whole step -- NCCL kernels included -- is
        one captured graph); ``defer_scale``: leave the 1/world factor to the optimizer (FusedSGD.grad_scale)
        instead of a separate pass over the flat gradient buffer."""
        super().__init__()
        self.module = module
        self.group = process_group
        self.side_stream = side_stream
        self.defer_scale = defer_scale
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        if self.world > 1 and broadcast_from is not None:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=broadcast_from, group=process_group)
        self._reducer: Optional[BucketReducer] = None

    def _attach(self, eng):
        if self.world == 1 or getattr(eng, '_ddp_owner', None) is self:
            return
        eng._ddp_owner = self
        # The bucket allreduces overlap the backward pass here.  With NCCL kernels resident, weight-gradient kernels on
        # the side stream AND programmatic dependent launch together were measured 2.5x slower (48 vs 18-20 ms/step at
        # 2 GPUs: early-launched dependents and NCCL compete for the same SM slots), each alone is fine -- this mode
        # keeps PDL and runs the weight gradients on the main stream.  (Graph mode, bench.py's default, replays
        # forward+backward without any NCCL kernel in flight and keeps both.)
        eng.use_side = bool(self.side_stream)

        def hook(bucket: int, eng=eng):
            if self._reducer is None or self._reducer.ranges != eng.bucket_ranges:
                self._reducer = BucketReducer(eng.bucket_ranges, self.group)
            self._reducer.launch(eng.gflat, bucket)
            if bucket == len(eng.bucket_ranges) - 1:
                self._reducer.finish(eng.gflat, scale=not self.defer_scale)

        eng.grad_hook = hook

    def forward(self, x):
        self._attach(self.module.engine())
        return self.module(x)
