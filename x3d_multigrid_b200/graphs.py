"""CUDA-graph capture of one whole training step (forward, loss, backward, optimizer) per clip shape.

A training step of X3D-M is ~700 kernel launches of 2-400 microseconds; replaying them as one CUDA graph
removes the host from the critical path.  The multigrid schedule (kinetics_multigrid.py:205-237) alternates
between 2-3 shapes inside a long cycle: keep one ``GraphedTrainStep`` per (B, T, H, W).

    step = GraphedTrainStep(model, optimizer, criterion, example_clip, example_labels)
    loss = step(clip, labels)          # copies into the static buffers (async), replays, returns the loss tensor

``optimizer`` should be ``optim.FusedSGD(..., capturable=True)`` so that LR changes reach the graph: the
hyper-parameters live in a small device tensor that is refreshed (only when ``param_groups`` changed) before
every replay, so ``lr_sched.step()`` of the reference loop (train_x3d_kinetics_multigrid.py:279) just works.

Building a step runs ``warmup`` real training steps on the example batch; by default the parameters, BN
buffers and momentum buffers are restored afterwards (``preserve_state=True``).

Data parallel, two ways:
* wrap the model in ``parallel.DistributedX3D(model, side_stream=True, defer_scale=True)`` and set
  ``optimizer.grad_scale = 1/world``: the bucket allreduces are issued from inside the backward pass and the WHOLE
  step -- NCCL kernels and the SGD kernel included -- is captured; a replay overlaps every bucket's allreduce with
  the backward pass of the earlier stages, nothing runs from the host between the first and the last kernel;
* or pass ``reduce_fn(flat)`` (e.g. ``lambda flat: dist.all_reduce(flat)``): the graph then holds forward + backward
  only; one NCCL call on the flat gradient buffer and the one-kernel optimizer step run eagerly after the replay.
  ``flat`` is the gradient buffer THIS graph was captured with (the engine may move on to another buffer when eager
  steps run in between)."""
from __future__ import annotations

import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, criterion, example_x, example_y, warmup=3, reduce_fn=None,
                 preserve_state=True):
        self.model, self.opt, self.crit = model, optimizer, criterion
        self.reduce_fn = reduce_fn
        if warmup < 1:
            # the first optimizer step initialises the momentum buffers (first_step is baked into a capture)
            raise ValueError('GraphedTrainStep needs warmup >= 1')
        dev = example_x.device if isinstance(example_x.device, torch.device) else torch.device(example_x.device)
        # the warm-up steps below are real training steps on the example batch; with preserve_state the parameters,
        # BN buffers and momentum buffers are put back afterwards (in place: the graph keeps their addresses)
        snap = self._snapshot() if preserve_state else None
        from .input_pipeline import UInt8Clips
        self.static_y = torch.empty_like(example_y, device=dev)
        self.static_y.copy_(example_y)
        if isinstance(example_x, UInt8Clips):
            # decoded uint8 frames: the graph reads static frame / crop-table buffers through the fused stem kernels
            self.static_x = UInt8Clips(example_x.frames.clone(), example_x.crops.clone(), example_x.size, example_x.mean,
                                       example_x.std, example_x.norm_value)
        else:
            self.static_x = torch.empty_like(example_x, device=dev)
            self.static_x.copy_(example_x)
        # warm-up on a side stream (allocator pools, parameter tables, kernel attributes, arena sizes)
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._one_step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.static_loss = self._one_step(captured=True)
        torch.cuda.synchronize(dev)
        # everything the captured kernels address must outlive the graph: the flat gradient buffer belongs to the
        # engine, which replaces it when an eager backward finds views of it still alive (Engine.new_grad_buffer)
        net = getattr(self.model, 'module', self.model)
        self.flat = net.engine().gflat if hasattr(net, 'engine') else None
        if snap is not None:
            self._restore(snap)

    def _snapshot(self):
        net = getattr(self.model, 'module', self.model)
        tensors = list(net.parameters()) + list(net.buffers())
        mom = {id(p): self.opt.state[p]['momentum_buffer'].clone() for p in net.parameters()
               if 'momentum_buffer' in self.opt.state.get(p, {})}
        return [t.detach().clone() for t in tensors], mom

    @torch.no_grad()
    def _restore(self, snap):
        net = getattr(self.model, 'module', self.model)
        saved, mom = snap
        for t, s0 in zip(list(net.parameters()) + list(net.buffers()), saved):
            t.copy_(s0)
        for p in net.parameters():
            st = self.opt.state.get(p, {})
            if 'momentum_buffer' in st:
                if id(p) in mom:
                    st['momentum_buffer'].copy_(mom[id(p)])
                else:
                    st['momentum_buffer'].zero_()     # momentum*0 + d == the first-step rule buf = d

    def _one_step(self, captured=False):
        self.opt.zero_grad(set_to_none=True)
        logits = self.model(self.static_x)
        loss = self.crit(logits, self.static_y)
        loss.backward()
        if self.reduce_fn is None:
            self.opt.step()
        elif not captured:
            self._finish()
        return loss.detach()

    def _finish(self):
        """eager tail of a data-parallel step: gradient allreduce + optimizer"""
        net = getattr(self.model, 'module', self.model)
        flat = getattr(self, 'flat', None)
        if flat is None and hasattr(net, 'engine'):
            flat = net.engine().gflat              # warm-up steps before the capture
        self.reduce_fn(flat)
        self.opt.step()

    def __call__(self, x, y):
        if x.shape != self.static_x.shape:
            raise RuntimeError('GraphedTrainStep is bound to one clip shape; keep one instance per multigrid shape')
        from .input_pipeline import UInt8Clips
        if isinstance(self.static_x, UInt8Clips):
            if not isinstance(x, UInt8Clips) or x.frames.shape != self.static_x.frames.shape:
                raise RuntimeError('this step was captured for uint8 frames of another geometry')
            self.static_x.frames.copy_(x.frames, non_blocking=True)
            self.static_x.crops.copy_(x.crops, non_blocking=True)
        else:
            self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        return self.replay()

    def replay(self, clone_loss=True):
        """replay on whatever currently sits in static_x / static_y (inputs staged by the caller).  Returns the loss
        of THIS step (a copy: the graph overwrites its static loss tensor at the next replay)."""
        if hasattr(self.opt, 'sync_hyper'):
            self.opt.sync_hyper()                  # LR / momentum / weight-decay edits since the last step
        self.graph.replay()
        if self.reduce_fn is not None:
            self._finish()
        return self.static_loss.clone() if clone_loss else self.static_loss
