"""Multigrid training schedule: which clip shape and batch size every iteration uses, what happens
on a long-cycle change, and a trainer that keeps one captured CUDA graph per shape.

Data-free restatement of the reference's scheduling logic:

* ``CycleSchedule`` / ``CycleBatchSampler`` / ``RandomEpochSampler`` -- cycle_batch_sampler.py:4-113.  The long cycle
  index walks 0,1,2,3 inside every LR phase (each phase split in four equal chunks); the last phase runs with index -1
  (no long-cycle changes).  The per-iteration batch is ``batch_size * long_cycle[long] * short multiplier`` with a
  2-step short cycle (x2, x1) for long indices 0/1 and a 3-step one (x4, x2, x1) otherwise.
* ``clip_shape`` -- kinetics_multigrid.py:205-237: frames and crop of the clip for (long index, iteration in epoch).
* ``LongCycleController`` -- train_x3d_kinetics_multigrid.py:226-234: BN split count and LR law on a long-cycle change;
  ``lr_warmup`` -- :300-305.
* ``MultigridTrainer`` -- one ``graphs.GraphedTrainStep`` per (clip shape, BN splits); the multigrid schedule alternates
  between 2-3 shapes inside a long cycle, so every shape is captured once and replayed afterwards.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Iterator, List, Sequence, Tuple

import torch
from torch.utils.data import sampler as _sampler

LONG_CYCLE = (8, 4, 2, 1)                  # batch / BN-split scale per long cycle (train_...multigrid.py:55)
LONG_CYCLE_LR_SCALE = (8, 0.5, 0.5, 0.5)   # LR factor on entering a long cycle (:56)


# ---------------------------------------------------------------------------------------------------------
# schedule state machine
# ---------------------------------------------------------------------------------------------------------
class CycleSchedule:
    """Long-cycle index as a function of the global iteration counter (cycle_batch_sampler.py:72-95).

    ``schedule`` = [0, end of LR phase 1, ..., last iteration].  Stateful like the reference: ``advance(it)`` applies at
    most one transition per call (a phase change when ``it`` has passed the phase end, else the next quarter of the
    phase), so callers query it once per iteration (and five times at start-up to catch up after a resume)."""

    def __init__(self, schedule: Sequence[int], n_long: int = 4):
        self.schedule = list(schedule)
        self.n_long = n_long
        self.phase = 1
        self.offset = 0.0
        self.long_index = 0
        self._quarter = (self.schedule[1] - self.schedule[0]) / n_long

    @property
    def last_phase(self) -> bool:
        return self.phase == len(self.schedule) - 1

    def advance(self, iteration: int) -> bool:
        """returns True when the long-cycle index may have changed"""
        if not self.last_phase and iteration > self.schedule[self.phase]:
            self.offset = self.schedule[self.phase]
            self.phase += 1
            self._quarter = (self.schedule[self.phase] - self.schedule[self.phase - 1]) / self.n_long
            self.long_index = -1 if self.last_phase else 0
            return True
        if iteration >= self._quarter + self.offset:
            self.offset += self._quarter
            self.long_index = -1 if self.last_phase else min(self.long_index + 1, self.n_long - 1)
            return True
        return False


def short_cycle_len(long_index: int) -> int:
    return 2 if long_index in (0, 1) else 3


def short_cycle_batch_scale(long_index: int, it_in_epoch: int) -> int:
    """cycle_batch_sampler.py:98-113"""
    if long_index in (0, 1):
        return 2 if it_in_epoch % 2 == 0 else 1
    return (4, 2, 1)[it_in_epoch % 3]


def clip_shape(long_index: int, it_in_epoch: int, frames: int, crop: int) -> Tuple[int, int]:
    """(T, H=W) of the clips of this iteration (kinetics_multigrid.py:205-237)."""
    small = int(math.floor(crop / math.sqrt(2)))
    t, c = ((frames // 4, small), (frames // 2, small), (frames // 2, crop), (frames, crop))[long_index]
    if long_index in (0, 1):
        if it_in_epoch % 2 == 0:
            c = int(math.floor(c / math.sqrt(2)))
    else:
        k = it_in_epoch % 3
        if k == 0:
            c = c // 2
        elif k == 1:
            c = int(math.floor(c / math.sqrt(2)))
    return t, c


class RandomEpochSampler(_sampler.RandomSampler):
    """Endless stream of random permutations of the data set (cycle_batch_sampler.py:4-25); ``len`` = epochs x size."""

    def __init__(self, data_source, replacement=False, num_samples=None, epochs=1):
        self.epochs = epochs
        super().__init__(data_source, replacement, num_samples)

    @property
    def num_samples(self):
        n = len(self.data_source) if self._num_samples is None else self._num_samples
        return n * self.epochs

    def __len__(self):
        return self.num_samples

    def __iter__(self):
        n = len(self.data_source)
        while True:
            yield from torch.randperm(n).tolist()


class CycleBatchSampler(_sampler.BatchSampler):
    """Batches of ``(sample index, long-cycle index)`` whose size follows the multigrid schedule
    (same constructor as the reference, cycle_batch_sampler.py:28-70)."""

    def __init__(self, sampler, batch_size, drop_last, schedule, cur_iterations, long_cycle_bs_scale):
        super().__init__(sampler, batch_size, drop_last)
        self.long_cycle_bs_scale = list(long_cycle_bs_scale)
        self.state = CycleSchedule(schedule, len(self.long_cycle_bs_scale))
        self.iteration_counter = cur_iterations
        self.short_iteration_counter = 0

    @property
    def long_cycle_index(self):
        return self.state.long_index

    def _batch_now(self) -> int:
        li = self.state.long_index
        return self.batch_size * self.long_cycle_bs_scale[li] * short_cycle_batch_scale(li, self.short_iteration_counter)

    def __iter__(self):
        self.short_iteration_counter = 0
        for _ in range(5):                      # catch up with a resumed iteration counter
            self.state.advance(self.iteration_counter)
        want = self._batch_now()
        batch: List[Tuple[int, int]] = []
        for idx in self.sampler:
            batch.append((idx, self.state.long_index))
            if len(batch) == want:
                yield batch
                batch = []
                self.iteration_counter += 1
                self.short_iteration_counter += 1
                self.state.advance(self.iteration_counter)
                want = self._batch_now()
        if batch and not self.drop_last:
            yield batch


def iteration_plan(batch_size: int, schedule: Sequence[int], frames: int, crop: int, n_iterations: int,
                   long_cycle: Sequence[int] = LONG_CYCLE, start_iteration: int = 0) -> Iterator[Dict]:
    """Data-free view of one epoch of the schedule: per iteration the long index, batch size and clip shape."""
    st = CycleSchedule(schedule, len(long_cycle))
    for _ in range(5):
        st.advance(start_iteration)
    for k in range(n_iterations):
        li = st.long_index
        t, h = clip_shape(li, k, frames, crop)
        yield {'iteration': start_iteration + k, 'long_index': li,
               'batch': batch_size * long_cycle[li] * short_cycle_batch_scale(li, k), 'frames': t, 'crop': h}
        st.advance(start_iteration + k + 1)


# ---------------------------------------------------------------------------------------------------------
# long-cycle side effects
# ---------------------------------------------------------------------------------------------------------
class LongCycleController:
    """BN split count and learning-rate law on a long-cycle change (train_x3d_kinetics_multigrid.py:226-234)."""

    def __init__(self, model, optimizer, long_cycle: Sequence[int] = LONG_CYCLE,
                 lr_scale: Sequence[float] = LONG_CYCLE_LR_SCALE):
        self.model, self.opt = model, optimizer
        self.long_cycle, self.lr_scale = tuple(long_cycle), tuple(lr_scale)
        self.last_long = -2                     # "not started" (also the state after a restart)
        self.bn_splits = None

    def on_batch(self, long_index: int) -> bool:
        """call with the long index of the incoming batch; returns True when it changed"""
        if long_index == self.last_long:
            return False
        net = getattr(self.model, 'module', self.model)
        self.bn_splits = net.update_bn_splits_long_cycle(self.long_cycle[long_index])
        first_or_last = self.last_long == -2 or long_index == -1
        factor = self.long_cycle[long_index] if first_or_last else self.lr_scale[long_index]
        self.last_long = long_index
        for g in self.opt.param_groups:
            g['lr'] *= factor
        if hasattr(self.opt, 'sync_hyper'):
            self.opt.sync_hyper()
        return True


def lr_warmup(init_lr: float, cur_steps: int, warmup_steps: int, opt) -> None:
    """linear warm-up of the first ``warmup_steps`` updates (train_x3d_kinetics_multigrid.py:300-305)"""
    if 1 < cur_steps < warmup_steps:
        scale = min(1.0, float(cur_steps + 1) / warmup_steps)
        for g in opt.param_groups:
            g['lr'] = scale * init_lr
        if hasattr(opt, 'sync_hyper'):
            opt.sync_hyper()


# ---------------------------------------------------------------------------------------------------------
# trainer: one captured graph per clip shape
# ---------------------------------------------------------------------------------------------------------
class MultigridTrainer:
    """Runs training steps whose clip shape follows the multigrid schedule.

        trainer = MultigridTrainer(model, FusedSGD(..., capturable=True), criterion)
        for clips, labels, long_index in loader:
            loss = trainer.step(clips, labels, long_index)

    Per (clip shape, BN split count) one ``GraphedTrainStep`` is captured at first use and replayed afterwards;
    ``use_graphs=False`` runs every step eagerly (same arithmetic)."""

    def __init__(self, model, optimizer, criterion, long_cycle: Sequence[int] = LONG_CYCLE,
                 lr_scale: Sequence[float] = LONG_CYCLE_LR_SCALE, use_graphs: bool = True, reduce_fn=None):
        self.model, self.opt, self.crit = model, optimizer, criterion
        self.ctrl = LongCycleController(model, optimizer, long_cycle, lr_scale)
        self.use_graphs = use_graphs
        self.reduce_fn = reduce_fn
        self.graphs: Dict[Tuple, object] = {}
        self.steps = 0

    def _eager(self, x, y):
        self.opt.zero_grad(set_to_none=True)
        loss = self.crit(self.model(x), y)
        loss.backward()
        if self.reduce_fn is not None:
            net = getattr(self.model, 'module', self.model)
            self.reduce_fn(net.engine().gflat)
        self.opt.step()
        return loss.detach()

    def step(self, x: torch.Tensor, y: torch.Tensor, long_index: int):
        self.ctrl.on_batch(long_index)
        self.steps += 1
        if not self.use_graphs:
            return self._eager(x, y)
        from .graphs import GraphedTrainStep
        key = (tuple(x.shape), tuple(y.shape), self.ctrl.bn_splits)
        g = self.graphs.get(key)
        if g is None:
            g = self.graphs[key] = GraphedTrainStep(self.model, self.opt, self.crit, x, y, reduce_fn=self.reduce_fn,
                                                    preserve_state=True)
        return g(x, y)
