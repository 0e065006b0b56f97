"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/x3d_b200.h
declares; the nn.Module surface has the reference's state_dict layout; host logic."""
import ctypes
import json
import os

import pytest
import torch

from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
    from x3d_multigrid_b200 import _lib
    protos = _lib.parse_header()
    names = [p[0] for p in protos]
    assert len(names) >= 30 and len(set(names)) == len(names)
    assert os.path.exists(_lib.LIB_PATH), 'run __graft_entry__.build() first'
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(cdll, n), f'{n} declared in include/x3d_b200.h but not exported'
    L = _lib.lib()
    assert L.fn['x3d_abi_version']() == 2
    assert L.launch_count() == 0            # nothing may launch without a GPU
    paths = L.path_counts()                 # kernel-family counters (x3d_path_t)
    assert len(paths) == 15 and 'pw_fwd_tc' in paths and 'dw_wgrad_tiled' in paths
    assert all(v == 0 for v in paths.values())


def test_workspace_size_queries_are_host_only():
    """the workspace-size entry points are plain host arithmetic (usable before any GPU work): head split-K scratch =
    4 KB of tile tickets + slices x tiles x 16 KB, never more than the engine's fixed 16 MB for any multigrid batch"""
    from x3d_multigrid_b200 import _lib
    L = _lib.lib()
    q = L.fn['x3d_small_gemm_workspace_bytes']
    assert q(16, 400, 2048) == 4096 + 16 * 7 * 64 * 64 * 4          # fc2 forward: 7 column tiles, 16 slices of K
    assert q(16, 2048, 432) == 4096 + 6 * 32 * 64 * 64 * 4          # fc1 forward: K / 64 = 6 slices
    assert q(64, 2048, 16) == 4096 + 1 * 32 * 64 * 64 * 4           # K of one k-tile: no split
    for rows in (1, 16, 64, 128, 256):
        for n, k in ((2048, 432), (400, 2048), (432, 2048), (2048, 400)):
            assert 4096 < q(rows, n, k) <= (16 << 20)
    assert int(L.fn['x3d_pwconv_wgrad_workspace_bytes']()) >= (1 << 20)
    assert L.launch_count() == 0


def test_struct_layouts_match_header():
    from x3d_multigrid_b200 import _lib
    assert ctypes.sizeof(_lib.PackDesc) == 40
    assert ctypes.sizeof(_lib.SgdDesc) == 32


def test_state_dict_layout_matches_reference_manifest():
    import x3d_multigrid_b200 as X
    with open(os.path.join(GOLDEN, 'state_dict_manifest.json')) as f:
        man = json.load(f)
    for v, s in (('S', 1), ('M', 4), ('XL', 2)):
        m = X.generate_model(v, n_classes=400, base_bn_splits=s)
        mine = [[k, list(t.shape), str(t.dtype).replace('torch.', '')] for k, t in m.state_dict().items()]
        assert mine == man[f'{v}_s{s}'], (v, s)
    m = X.generate_model('M', n_classes=157, base_bn_splits=1, task='loc')
    mine = [[k, list(t.shape), str(t.dtype).replace('torch.', '')] for k, t in m.state_dict().items()]
    assert mine == man['M_loc157']
    # update_bn_splits_long_cycle (x3d.py:298-303)
    m = X.generate_model('M', n_classes=400, base_bn_splits=2)
    ret = m.update_bn_splits_long_cycle(4)
    mine = [[k, list(t.shape), str(t.dtype).replace('torch.', '')] for k, t in m.state_dict().items()]
    assert ret == man['M_s2_resplit4']['ret'] and mine == man['M_s2_resplit4']['entries']
    m.replace_logits(157)
    assert m.fc2.weight.shape == (157, 2048)


def test_aggregate_stats_matches_oracle():
    import x3d_multigrid_b200 as X
    from oracle import x3d_oracle as O
    bn = X.SubBatchNorm3d(num_splits=4, num_features=6)
    bn.split_bn.running_mean.copy_(O.det_tensor((24,), 'rm', dtype=torch.float32))
    bn.split_bn.running_var.copy_(O.det_tensor((24,), 'rv', dtype=torch.float32).abs() + 0.5)
    bn.aggregate_stats()
    m, v = O.aggregate_stats(bn.split_bn.running_mean.double(), bn.split_bn.running_var.double(), 4)
    assert torch.allclose(bn.bn.running_mean.double(), m, atol=1e-6)
    assert torch.allclose(bn.bn.running_var.double(), v, atol=1e-6)


def test_cpu_input_fails_loudly():
    import x3d_multigrid_b200 as X
    m = X.generate_model('S', n_classes=5, base_bn_splits=1)
    with pytest.raises(RuntimeError, match='CUDA'):
        m(torch.zeros(1, 3, 4, 32, 32))


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference tree not present')
def test_same_seed_gives_reference_init():
    import sys
    sys.path.insert(0, '/root/reference')
    import x3d as R
    import x3d_multigrid_b200 as X
    torch.manual_seed(3)
    a = R.generate_model('S', n_classes=11, base_bn_splits=2).state_dict()
    torch.manual_seed(3)
    b = X.generate_model('S', n_classes=11, base_bn_splits=2).state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
