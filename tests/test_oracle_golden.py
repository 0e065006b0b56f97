"""Pin the CPU oracle (oracle/x3d_oracle.py) against outputs of the reference itself.

tests/golden/*.npz were produced by oracle/make_golden.py, which imports the
unmodified /root/reference/x3d.py and runs it in fp64.  No GPU needed.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import x3d_oracle as O
from oracle.make_golden import CASES, case_clip

REF = '/root/reference'


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def _run_oracle(case, conv_impl):
    c = CASES[case]
    sd = O.make_state_dict(c['version'], c['n_classes'], c['splits'])
    x = case_clip(c['shape'])
    gold = np.load(os.path.join(GOLDEN, case + '.npz'))
    labels = torch.from_numpy(gold['labels'])
    logits, loss, grads, stats = O.loss_and_grads(sd, x, labels, loss=c.get('loss', 'ce'), version=c['version'],
                                                  splits=c['splits'], training=True, task=c['task'],
                                                  conv_impl=conv_impl)
    return gold, sd, x, logits, loss, grads, stats


@pytest.mark.parametrize('case,conv_impl', [('s_small_split2', 'explicit'), ('s_small_split2', 'aten'),
                                            ('m_odd_loc', 'explicit'), ('m_odd_loc', 'aten'),
                                            ('m_charades_cls', 'aten'), ('m_charades_loc', 'aten'),
                                            ('xl_small', 'aten'), ('m_mg_111', 'aten')])
def test_oracle_matches_reference_golden(case, conv_impl):
    gold, sd, x, logits, loss, grads, stats = _run_oracle(case, conv_impl)
    assert _rel(logits.numpy(), gold['logits']) < 1e-10
    assert abs(loss.item() - float(gold['loss'])) < 1e-10
    for k, g in grads.items():
        # fp64 noise floor of whole-network grads through train-mode BN is ~1e-9
        assert _rel(g.reshape(-1)[:16].numpy(), gold['ghead/' + k]) < 1e-6, k
        assert abs(g.norm().item() - float(gold['gnorm/' + k])) <= 1e-6 * float(gold['gnorm/' + k]) + 1e-12, k
    for k in gold.files:
        if k.startswith('gfull/'):
            assert _rel(grads[k[6:]].numpy(), gold[k]) < 1e-6, k
        if k.startswith('stat/'):
            assert _rel(stats[k[5:]].numpy(), gold[k]) < 1e-10, k


def test_oracle_eval_and_aggregate():
    case = 's_small_split2'
    c = CASES[case]
    gold, sd, x, logits, loss, grads, stats = _run_oracle(case, 'explicit')
    sd2 = dict(sd)
    sd2.update(stats)
    for k in list(sd2):
        if k.endswith('split_bn.running_mean'):
            p = k[:-len('.split_bn.running_mean')]
            m, v = O.aggregate_stats(sd2[k], sd2[p + '.split_bn.running_var'], c['splits'])
            sd2[p + '.bn.running_mean'], sd2[p + '.bn.running_var'] = m, v
    for k in gold.files:
        if k.startswith('agg/'):
            assert _rel(sd2[k[4:]].numpy(), gold[k]) < 1e-12, k
    with torch.no_grad():
        ev = O.forward(sd2, x, version=c['version'], splits=c['splits'], training=False, task=c['task'])
    assert _rel(ev.numpy(), gold['eval_logits']) < 1e-10


@pytest.mark.slow
def test_oracle_config1_logits():
    """BASELINE config 1 (X3D-S, B=2, 13x160x160): forward only, aten convs (fast)."""
    c = CASES['s_config1']
    gold = np.load(os.path.join(GOLDEN, 's_config1.npz'))
    sd = O.make_state_dict(c['version'], c['n_classes'], c['splits'])
    x = case_clip(c['shape'])
    with torch.no_grad():
        logits = O.forward(sd, x, version='S', splits=1, training=True, conv_impl='aten')
    assert _rel(logits.numpy(), gold['logits']) < 1e-10


def test_by_split_equivalence():
    """oracle/make_golden.py produces the BASELINE config-2 golden (batch 16, 2 BN splits) split by split: the
    network on x with s splits equals s independent runs on x[b::s] with one split (x3d.py:47-52).  Checked
    here on a small case: logits interleave, loss / gradients average, running statistics concatenate."""
    sd2 = O.make_state_dict('S', 13, 2)
    sd1 = O.make_state_dict('S', 13, 1)
    x = case_clip((4, 3, 4, 24, 28))
    labels = torch.tensor([[1], [5], [9], [12]])
    lg, ls, g, st = O.loss_and_grads(sd2, x, labels, version='S', splits=2, training=True, conv_impl='aten')
    acc, loss = {}, 0.0
    for b in range(2):
        rows = torch.arange(b, 4, 2)
        lgb, lsb, gb, stb = O.loss_and_grads(sd1, x[rows], labels[rows], version='S', splits=1, training=True,
                                             conv_impl='aten')
        assert _rel(lgb.numpy(), lg[rows].numpy()) < 1e-12
        loss += lsb.item() / 2
        for k, v in gb.items():
            acc[k] = v / 2 if k not in acc else acc[k] + v / 2
        for k, v in stb.items():
            C = v.numel()
            assert _rel(v.numpy(), st[k][b * C:(b + 1) * C].numpy()) < 1e-12, k
    assert abs(loss - ls.item()) < 1e-12
    for k, v in g.items():
        assert _rel(acc[k].numpy(), v.numpy()) < 1e-9, k


@pytest.mark.slow
def test_golden_config2_summary_is_consistent():
    """the committed BASELINE config-2 fixtures exist and carry what the GPU tests read"""
    for case in ('m_config2', 'm_config2_b4'):
        gold = np.load(os.path.join(GOLDEN, case + '.npz'))
        B = CASES[case]['shape'][0]
        assert gold['logits'].shape == (B, 400, 1) and gold['eval_logits'].shape == (B, 400, 1)
        assert abs(float(gold['loss']) - np.log(400)) < 0.2          # near-uniform logits at init
        assert gold['stat/bn1.split_bn.running_mean'].shape == (2 * 24,)


def test_manifest_matches_reference():
    with open(os.path.join(GOLDEN, 'state_dict_manifest.json')) as f:
        man = json.load(f)
    for v in ('S', 'M', 'XL'):
        for s in (1, 2, 4, 8):
            mine = [[k, list(shp), dt] for k, shp, dt in O.state_dict_manifest(v, 400, s)]
            assert mine == man[f'{v}_s{s}'], (v, s)
    mine = [[k, list(shp), dt] for k, shp, dt in O.state_dict_manifest('M', 157, 1)]
    assert mine == man['M_loc157']
    assert len(man['M_s4']) == 820 and len(man['XL_s1']) == 1659


def test_multigrid_shape_law():
    """SURVEY.md A3 / log lines 15,82,158,234 (global batch 128, T0=8, crop 224)."""
    t = O.multigrid_shapes(128, 8, 224)
    assert t[0] == [(2048, 2, 111), (1024, 2, 158)]
    assert t[1] == [(1024, 4, 111), (512, 4, 158)]
    assert t[2] == [(1024, 4, 112), (512, 4, 158), (256, 4, 224)]
    assert t[3] == [(512, 8, 112), (256, 8, 158), (128, 8, 224)]


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present on this machine')
def test_oracle_matches_live_reference_block():
    """Live check against the reference module itself (build container only)."""
    sys.path.insert(0, REF)
    import x3d as ref
    torch.manual_seed(0)
    m = ref.generate_model('S', n_classes=11, dropout=0.0, base_bn_splits=2).double()
    sd = O.det_fill_state_dict(m.state_dict())
    m.load_state_dict(sd)
    m.train()
    x = case_clip((4, 3, 3, 20, 24))
    got = O.forward(sd, x, version='S', splits=2, training=True)
    assert _rel(got.detach().numpy(), m(x).detach().numpy()) < 1e-10
