"""Validation pass of the reference loops on the CUDA path.

* ``prepare_eval`` -- what the scripts do before the 'val' phase: ``aggregate_sub_bn_stats()`` folds the split running
  statistics into ``bn`` and the model goes to eval mode (train_x3d_kinetics_multigrid.py:203-206).  In eval mode the
  engine derives scale/shift from ``bn.running_*`` once per forward and applies them where the training path applies
  the batch statistics: on load in the depthwise conv, in the Swish/SE pass and in the residual pass -- i.e. BatchNorm
  is already folded into its consumers, no statistics kernels run.
* ``predict_multicrop`` -- the multi-crop reduction: Kinetics validation feeds ``[b, n, c, t, h, w]`` (n temporal crops
  per video, kinetics.py:218-233), views it as ``[b*n, ...]``, and averages the per-crop softmax
  (train_x3d_kinetics_multigrid.py:240-257); Charades takes the maximum over crops of sigmoid scores and logits
  (train_x3d_charades.py:156-174)."""
from __future__ import annotations

from typing import Tuple

import torch


def prepare_eval(model) -> int:
    net = getattr(model, 'module', model)
    count = net.aggregate_sub_bn_stats()
    model.eval()
    return count


@torch.no_grad()
def predict_multicrop(model, inputs: torch.Tensor, reduce: str = 'softmax_mean') -> Tuple[torch.Tensor, torch.Tensor]:
    """inputs [b, n, c, t, h, w] -> (scores [b, C, 1], logits [b, C, 1]) reduced over the n crops.

    ``softmax_mean`` (Kinetics): scores = mean_n softmax_C(logits), logits = mean_n logits; predictions are
    ``scores.argmax(1)``.  ``max`` (Charades): scores = max_n sigmoid(logits), logits = max_n logits."""
    if inputs.dim() != 6:
        raise RuntimeError('expected [b, n_crops, c, t, h, w]')
    b, n, c, t, h, w = inputs.shape
    logits = model(inputs.reshape(b * n, c, t, h, w).contiguous())          # [b*n, C, 1]
    logits = logits.view(b, n, logits.shape[1], logits.shape[2])
    if reduce == 'softmax_mean':
        return torch.softmax(logits, dim=2).mean(1), logits.mean(1)
    if reduce == 'max':
        return torch.sigmoid(logits).max(dim=1)[0], logits.max(dim=1)[0]
    raise ValueError(f'unknown reduction {reduce!r}')
