"""GPU-side input pipeline against the reference's own spatial transform chain.

tests/golden/input_pipeline.npz holds uint8 frames and the fp32 clips that MultiScaleRandomCropMultigrid ->
RandomHorizontalFlip -> ToTensor(255) -> Normalize (transforms/spatial_transforms.py, run by oracle/make_golden.py in
the build container) produce from them.  The CUDA path must reproduce those clips BIT FOR BIT, and the stem kernels
that read the frames directly must give exactly the network output of the materialised clip."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import x3d_oracle as O

pytestmark = pytest.mark.gpu


def _load():
    import x3d_multigrid_b200 as X
    g = np.load(os.path.join(GOLDEN, 'input_pipeline.npz'))
    frames = torch.from_numpy(g['frames']).cuda()
    crops = torch.from_numpy(g['crops']).cuda()
    clips = X.UInt8Clips(frames, crops, int(g['clip'].shape[-1]), tuple(g['mean']), tuple(g['std']))
    return X, g, clips


def test_clip_from_uint8_is_bit_exact():
    X, g, clips = _load()
    clips.check_windows()
    got = X.clip_from_uint8(clips)
    want = torch.from_numpy(g['clip'])
    assert got.shape == want.shape and got.dtype == torch.float32
    assert torch.equal(got.cpu(), want)                      # crop window, flip, /255, (x - mean) / std: same roundings
    # crop_table helper builds the same table
    c = g['crops']
    assert torch.equal(X.crop_table(c[:, 0], c[:, 1], c[:, 2], 'cuda'), clips.crops)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_network_on_uint8_frames_equals_network_on_clip(dtype):
    X, g, clips = _load()
    sd = O.make_state_dict('S', 9, 1)
    y = torch.tensor([[1], [4], [7]]).cuda()
    outs = []
    fused = X.UInt8Clips(clips.frames, clips.crops, clips.size, clips.mean, clips.std, fused=True)
    for source in ('clip', 'frames', 'frames_fused'):
        m = X.generate_model('S', n_classes=9, base_bn_splits=1, dropout=0.0)
        m.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd.items()})
        m = m.cuda().set_compute_dtype(dtype).train()
        x = X.clip_from_uint8(clips) if source == 'clip' else (clips if source == 'frames' else fused)
        logits = m(x)
        torch.nn.functional.cross_entropy(logits, y).backward()
        outs.append((logits.detach().clone(), m.conv1_s.weight.grad.clone(), m.fc2.weight.grad.clone()))
    for o in outs[1:]:
        assert torch.equal(outs[0][0], o[0])                 # the stem saw bit-identical input values
        for a, b in zip(outs[0][1:], o[1:]):
            assert float((a - b).norm() / b.norm()) < 1e-5   # weight-gradient sums: fp32 reds of leaf outputs


def test_uint8_clips_validation():
    X, g, clips = _load()
    with pytest.raises(RuntimeError):
        X.UInt8Clips(clips.frames.cpu(), clips.crops, clips.size)
    with pytest.raises(RuntimeError):
        X.UInt8Clips(clips.frames, clips.crops, 64)          # larger than the 40 x 48 source frames
    bad = clips.crops.clone()
    bad[0, 0] = 40
    with pytest.raises(RuntimeError):
        X.UInt8Clips(clips.frames, bad, clips.size).check_windows()
