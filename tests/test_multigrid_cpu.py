"""Multigrid scheduler against the reference's golden schedule -- CPU only."""
import json
import os

import pytest
import torch

from x3d_multigrid_b200 import multigrid as MG

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_cycle_batch_sampler_matches_reference_schedule():
    """tests/golden/sampler_schedule.json was recorded from the reference CycleBatchSampler
    (cycle_batch_sampler.py:28-113) by oracle/make_golden.py: rows = [iteration, long index, batch length]."""
    g = json.load(open(os.path.join(GOLD, 'sampler_schedule.json')))

    class DS:
        def __len__(self):
            return 100000
    smp = MG.RandomEpochSampler(DS(), epochs=1)
    bs = MG.CycleBatchSampler(smp, g['batch_size'], False, schedule=g['schedule'], cur_iterations=0,
                              long_cycle_bs_scale=g['long_cycle'])
    rows = []
    for it, batch in enumerate(bs):
        assert all(li == batch[0][1] for _, li in batch)
        rows.append([it, batch[0][1], len(batch)])
        if it >= len(g['rows']) - 1:
            break
    assert rows == g['rows']
    # SURVEY.md A3 spot checks
    assert rows[0][1:] == [0, 64] and rows[40][1] == 1 and rows[80][1] == 2 and rows[120][1] == 3
    assert rows[161][1] == 0 and rows[341][1] == -1


def test_sampler_resume_catches_up():
    """cur_iterations > 0 (restart from a checkpoint): the first batches already carry the right long index"""
    g = json.load(open(os.path.join(GOLD, 'sampler_schedule.json')))
    bs = MG.CycleBatchSampler(range(10 ** 6), 4, False, schedule=g['schedule'], cur_iterations=125,
                              long_cycle_bs_scale=g['long_cycle'])
    first = next(iter(bs))
    assert first[0][1] == 3 and len(first) == 4 * 1 * 4


def test_clip_shape_law():
    """kinetics_multigrid.py:205-237 with frames=16, crop=224 (BASELINE config 3) and the log's T0=8"""
    assert [MG.clip_shape(0, k, 16, 224) for k in range(2)] == [(4, 111), (4, 158)]
    assert [MG.clip_shape(1, k, 16, 224) for k in range(2)] == [(8, 111), (8, 158)]
    assert [MG.clip_shape(2, k, 16, 224) for k in range(3)] == [(8, 112), (8, 158), (8, 224)]
    assert [MG.clip_shape(3, k, 16, 224) for k in range(3)] == [(16, 112), (16, 158), (16, 224)]
    assert [MG.clip_shape(-1, k, 16, 224) for k in range(3)] == [(16, 112), (16, 158), (16, 224)]
    # batch x shape keeps B*T*H*W roughly constant
    plan = list(MG.iteration_plan(16, [0, 160, 260, 340, 400], 16, 224, 6))
    assert [(p['batch'], p['frames'], p['crop']) for p in plan[:2]] == [(256, 4, 111), (128, 4, 158)]


class _FakeNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.calls = []

    def update_bn_splits_long_cycle(self, scale):
        self.calls.append(scale)
        return 2 * scale


def test_long_cycle_lr_and_bn_law():
    """train_x3d_kinetics_multigrid.py:226-234 / SURVEY.md A3: x8 on the first batch, x0.5 on 0->1->2->3, x8 on the
    3->0 wrap, x1 (LONG_CYCLE[-1]) on entering the last phase"""
    net = _FakeNet()
    opt = torch.optim.SGD(net.parameters(), lr=0.2)
    ctl = MG.LongCycleController(net, opt)
    lrs = []
    for li in (0, 0, 1, 2, 3, 3, 0, 1, -1, -1):
        ctl.on_batch(li)
        lrs.append(opt.param_groups[0]['lr'])
    assert lrs == pytest.approx([1.6, 1.6, 0.8, 0.4, 0.2, 0.2, 1.6, 0.8, 0.8, 0.8])
    assert net.calls == [8, 4, 2, 1, 8, 4, 1]
    assert ctl.bn_splits == 2


def test_lr_warmup():
    net = _FakeNet()
    opt = torch.optim.SGD(net.parameters(), lr=1.0)
    MG.lr_warmup(1.0, 1, 100, opt)
    assert opt.param_groups[0]['lr'] == 1.0                       # starts after step 1
    MG.lr_warmup(1.0, 49, 100, opt)
    assert opt.param_groups[0]['lr'] == pytest.approx(0.5)
    MG.lr_warmup(1.0, 100, 100, opt)
    assert opt.param_groups[0]['lr'] == pytest.approx(0.5)        # untouched once past the warm-up


def test_iteration_plan_equals_sampler_schedule():
    """the data-free plan (what MultigridTrainer pre-captures graphs for) reproduces the recorded sampler schedule"""
    g = json.load(open(os.path.join(GOLD, 'sampler_schedule.json')))
    plan = list(MG.iteration_plan(g['batch_size'], g['schedule'], 16, 224, len(g['rows']), g['long_cycle']))
    assert [[p['iteration'], p['long_index'], p['batch']] for p in plan] == g['rows']
    shapes = {(p['long_index'], p['batch'], p['frames'], p['crop']) for p in plan}
    # 2 + 2 + 3 + 3 shapes of the four long cycles, plus the three of the final phase (long index -1)
    assert len(shapes) == 13
    assert (3, 4, 16, 224) in shapes and (0, 64, 4, 111) in shapes and (-1, 16, 16, 112) in shapes
