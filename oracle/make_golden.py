"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/x3d.py).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

The reference module is imported from where it lies (never copied); it is fed
RNG-free deterministic weights and clips (oracle.x3d_oracle.det_*) and executed in
fp64 on the CPU.  Outputs are small summaries so the fixtures stay a few hundred KB.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import x3d_oracle as O  # noqa: E402

REF = os.environ.get('X3D_REFERENCE', '/root/reference')
GOLD = os.path.join(ROOT, 'tests', 'golden')

# name -> (version, n_classes, base_bn_splits, long_cycle_scale, clip shape, task, dropout_mask?)
CASES = {
    # small: every code path (SE / no-SE / downsample / odd sizes / 2 splits)
    's_small_split2': dict(version='S', n_classes=37, splits=2, shape=(4, 3, 4, 36, 44), task='class'),
    # odd spatial sizes as in the multigrid schedule (H=158-like -> 79 -> 40 ...), loc head
    'm_odd_loc': dict(version='M', n_classes=19, splits=1, shape=(2, 3, 5, 30, 38), task='loc'),
    # BASELINE config 1: X3D-S fwd+bwd, batch 2, 13x160x160, 400 classes
    's_config1': dict(version='S', n_classes=400, splits=1, shape=(2, 3, 13, 160, 160), task='class'),
}


def load_reference():
    sys.path.insert(0, REF)
    import x3d as ref_x3d  # the reference's own file
    return ref_x3d


def run_case(ref, name, version, n_classes, splits, shape, task):
    torch.manual_seed(0)
    model = ref.generate_model(version, n_classes=n_classes, dropout=0.0, base_bn_splits=splits, task=task)
    sd = O.det_fill_state_dict(model.state_dict())
    model = model.double()
    model.load_state_dict(sd)
    model.train()
    x = O.det_clip(shape)
    B = shape[0]
    if task == 'class':
        labels = torch.tensor([(7 * i + 3) % n_classes for i in range(B)]).unsqueeze(1)
    else:
        T = shape[2]
        labels = torch.tensor([[(7 * i + 3 * t + 1) % n_classes for t in range(T)] for i in range(B)])
    logits = model(x)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()

    out = {'logits': logits.detach().numpy(), 'loss': np.array(loss.item()), 'labels': labels.numpy()}
    for k, p in model.named_parameters():
        g = p.grad
        assert g is not None, k
        out['gnorm/' + k] = np.array(g.norm().item())
        out['ghead/' + k] = g.reshape(-1)[:16].numpy().copy()
        out['gsum/' + k] = np.array(g.sum().item())
    after = model.state_dict()
    for k, v in after.items():
        if 'split_bn.running_' in k:
            out['stat/' + k] = v.numpy().copy()
    # a few full gradients that are cheap to keep and discriminate well
    for k in ('conv1_s.weight', 'conv1_t.weight', 'layer1.0.conv2.weight', 'layer2.1.conv2.weight',
              'layer1.0.fc1.weight', 'layer1.0.fc2.bias', 'layer4.6.conv2.weight', 'bn1.weight',
              'layer3.0.downsample.1.bias', 'fc2.bias'):
        out['gfull/' + k] = dict(model.named_parameters())[k].grad.numpy().copy()

    # eval-mode logits after aggregate_sub_bn_stats (x3d.py:306-313)
    model.aggregate_sub_bn_stats()
    model.eval()
    with torch.no_grad():
        out['eval_logits'] = model(x).numpy()
    for k in ('bn1.bn.running_mean', 'bn1.bn.running_var', 'layer2.0.bn2.bn.running_var'):
        out['agg/' + k] = model.state_dict()[k].numpy().copy()
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)
    print(name, 'loss', loss.item(), 'logit absmax', logits.abs().max().item())


def manifests(ref):
    man = {}
    for v in ('S', 'M', 'XL'):
        for s in (1, 2, 4, 8):
            m = ref.generate_model(v, n_classes=400, base_bn_splits=s)
            man[f'{v}_s{s}'] = [[k, list(t.shape), str(t.dtype).replace('torch.', '')]
                                for k, t in m.state_dict().items()]
    m = ref.generate_model('M', n_classes=157, base_bn_splits=1, task='loc')
    man['M_loc157'] = [[k, list(t.shape), str(t.dtype).replace('torch.', '')] for k, t in m.state_dict().items()]
    # long-cycle re-split (x3d.py:298-303)
    m = ref.generate_model('M', n_classes=400, base_bn_splits=2)
    ret = m.update_bn_splits_long_cycle(4)
    man['M_s2_resplit4'] = {'ret': ret, 'entries': [[k, list(t.shape), str(t.dtype).replace('torch.', '')]
                                                   for k, t in m.state_dict().items()]}
    with open(os.path.join(GOLD, 'state_dict_manifest.json'), 'w') as f:
        json.dump(man, f)


def sampler_golden():
    """Reference CycleBatchSampler schedule (cycle_batch_sampler.py:28-113), SURVEY.md A3."""
    sys.path.insert(0, REF)
    import cycle_batch_sampler as cbs

    class DS:
        def __len__(self):
            return 100000
    torch.manual_seed(0)
    smp = cbs.RandomEpochSampler(DS(), epochs=1)
    bs = cbs.CycleBatchSampler(smp, 4, False, schedule=[0, 160, 260, 340, 400], cur_iterations=0,
                               long_cycle_bs_scale=[8, 4, 2, 1])
    rows = []
    for it, batch in enumerate(bs):
        rows.append([it, batch[0][1], len(batch)])
        if it >= 399:
            break
    with open(os.path.join(GOLD, 'sampler_schedule.json'), 'w') as f:
        json.dump({'batch_size': 4, 'schedule': [0, 160, 260, 340, 400], 'long_cycle': [8, 4, 2, 1],
                   'rows': rows}, f)


if __name__ == '__main__':
    ref = load_reference()
    os.makedirs(GOLD, exist_ok=True)
    manifests(ref)
    sampler_golden()
    only = sys.argv[1:] or list(CASES)
    for name in only:
        run_case(ref, name, **CASES[name])
