"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/x3d.py).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

The reference module is imported from where it lies (never copied); it is fed
RNG-free deterministic weights and clips (oracle.x3d_oracle.det_*; clip values are
fp32-representable because the CUDA path takes fp32 clips) and executed in fp64 on the
CPU.  Outputs are small summaries so the fixtures stay a few hundred KB each.
`python oracle/make_golden.py m_config2` needs ~30 GB of RAM and a few minutes.
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

# name -> version, n_classes, base_bn_splits, clip shape, task [, loss, by_split]
CASES = {
    # small: every code path (SE / no-SE / downsample / odd sizes / 2 splits)
    's_small_split2': dict(version='S', n_classes=37, splits=2, shape=(4, 3, 4, 36, 44), task='class'),
    # odd spatial sizes as in the multigrid schedule (H=158-like -> 79 -> 40 ...), loc head
    'm_odd_loc': dict(version='M', n_classes=19, splits=1, shape=(2, 3, 5, 30, 38), task='loc'),
    # BASELINE config 1: X3D-S fwd+bwd, batch 2, 13x160x160, 400 classes
    's_config1': dict(version='S', n_classes=400, splits=1, shape=(2, 3, 13, 160, 160), task='class'),
    # BASELINE config 2 (the benchmarked one): X3D-M, batch 16, 16x224x224, base_bn_splits=2.  Run split by split
    # (see run_case: the two BN groups are independent sub-batches), 25 GB of fp64 activations each.
    'm_config2': dict(version='M', n_classes=400, splits=2, shape=(16, 3, 16, 224, 224), task='class', by_split=True),
    # same clip size, batch 4 (fp32 parity run of the production-size kernels)
    'm_config2_b4': dict(version='M', n_classes=400, splits=2, shape=(4, 3, 16, 224, 224), task='class'),
    # BASELINE config 3: one multigrid shape (long cycle 1 / short cycle 0: B x8, T/4, H=111), BN splits x4
    'm_mg_111': dict(version='M', n_classes=400, splits=4, shape=(8, 3, 4, 111, 111), task='class'),
    # BASELINE config 4: X3D-XL widths (72/162/306/630), reduced clip
    'xl_small': dict(version='XL', n_classes=400, splits=1, shape=(2, 3, 4, 156, 156), task='class'),
    # BASELINE config 5: Charades heads -- 157-way multi-label BCE (train_x3d_charades.py:97-122) and the
    # localisation variant (train_x3d_charades_loc.py:168-189: interpolate to the label length, cls + loc BCE)
    'm_charades_cls': dict(version='M', n_classes=157, splits=1, shape=(2, 3, 8, 112, 112), task='class', loss='bce'),
    'm_charades_loc': dict(version='M', n_classes=157, splits=1, shape=(2, 3, 8, 64, 64), task='loc', loss='bce_loc',
                           tl=12),
}


def load_reference():
    sys.path.insert(0, REF)
    import x3d as ref_x3d  # the reference's own file
    return ref_x3d


def case_clip(shape):
    """The clip of a case: fp32-representable values (the CUDA path takes fp32 clips, x3d.py:316), as fp64."""
    return O.det_clip(shape, dtype=torch.float32).double()


def case_labels(c):
    """Deterministic targets of a case (also used by the tests through the stored 'labels' array)."""
    B, n_classes = c['shape'][0], c['n_classes']
    kind = c.get('loss', 'ce')
    if kind == 'bce':          # multi-hot [B, C] (charades.py:91-98 style per-clip label vector)
        return (O.det_tensor((B, n_classes), 'multihot') > 0.8).double()
    if kind == 'bce_loc':      # per-frame multi-hot [B, C, TL]
        return (O.det_tensor((B, n_classes, c['tl']), 'multihot_loc') > 0.8).double()
    if c['task'] == 'class':
        return torch.tensor([(7 * i + 3) % n_classes for i in range(B)]).unsqueeze(1)
    T = c['shape'][2]
    return torch.tensor([[(7 * i + 3 * t + 1) % n_classes for t in range(T)] for i in range(B)])


def case_loss(c, logits, labels):
    """The loss the reference's training loops apply to the network output."""
    kind = c.get('loss', 'ce')
    if kind == 'ce':           # train_x3d_kinetics_multigrid.py:189,259
        return torch.nn.functional.cross_entropy(logits, labels)
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    if kind == 'bce':          # train_x3d_charades.py:162,176: logits.squeeze(2) vs [B,C]
        return bce(logits.squeeze(2), labels.to(logits.dtype))
    # train_x3d_charades_loc.py:168-189
    per_frame = torch.nn.functional.interpolate(logits, labels.shape[2], mode='linear')
    lab = labels.to(logits.dtype)
    cls_loss = bce(torch.max(per_frame, dim=2)[0], torch.max(lab, dim=2)[0])
    loc_loss = bce(per_frame, lab)
    return (cls_loss + loc_loss) / 2


def _fwd_bwd(ref, c, x, labels, splits, rows=None):
    """one reference forward+backward; ``rows`` selects a sub-batch (by_split mode)"""
    torch.manual_seed(0)
    model = ref.generate_model(c['version'], n_classes=c['n_classes'], dropout=0.0, base_bn_splits=splits,
                               task=c['task'])
    sd = O.det_fill_state_dict(model.state_dict())
    model = model.double()
    model.load_state_dict(sd)
    model.train()
    xs = x if rows is None else x[rows]
    ls = labels if rows is None else labels[rows]
    logits = model(xs)
    loss = case_loss(c, logits, ls)
    loss.backward()
    return model, logits.detach(), loss.detach()


def run_case(ref, name, c):
    shape, splits = c['shape'], c['splits']
    x = case_clip(shape)
    labels = case_labels(c)
    if not c.get('by_split'):
        model, logits, loss = _fwd_bwd(ref, c, x, labels, splits)
        grads = {k: p.grad for k, p in model.named_parameters()}
        stats = {k: v for k, v in model.state_dict().items() if 'split_bn.running_' in k}
    else:
        # SubBatchNorm3d (x3d.py:47-52) views the batch as (n//s, c*s, ...): BN group b is exactly the sub-batch
        # x[b::s] and nothing else couples samples, so the network on the full batch with s splits IS s
        # independent runs of the same reference module on x[b::s] with one split each: logits interleave, the
        # mean loss is the mean of the group losses, gradients average, split_bn.running_*[b*C+c] come from run b.
        # (tests/test_oracle_golden.py::test_by_split_equivalence checks this identity on a small case.)
        B = shape[0]
        logits = torch.zeros(B, c['n_classes'], 1 if c['task'] == 'class' else shape[2], dtype=torch.float64)
        loss = torch.zeros((), dtype=torch.float64)
        grads, per_split = {}, []
        for b in range(splits):
            rows = torch.arange(b, B, splits)
            model, lg, ls = _fwd_bwd(ref, c, x, labels, 1, rows)
            logits[rows] = lg
            loss += ls / splits
            for k, p in model.named_parameters():
                grads[k] = p.grad / splits if k not in grads else grads[k] + p.grad / splits
            per_split.append({k: v.clone() for k, v in model.state_dict().items() if 'split_bn.running_' in k})
            del model
        stats = {k: torch.cat([ps[k] for ps in per_split]) for k in per_split[0]}
        # eval pass needs a module with the s-split buffers: build it, fill the split statistics, aggregate
        torch.manual_seed(0)
        model = ref.generate_model(c['version'], n_classes=c['n_classes'], dropout=0.0, base_bn_splits=splits,
                                   task=c['task'])
        sd = O.det_fill_state_dict(model.state_dict())
        sd.update(stats)
        model = model.double()
        model.load_state_dict(sd)

    out = {'logits': logits.numpy(), 'loss': np.array(loss.item()), 'labels': labels.numpy()}
    for k, g in grads.items():
        assert g is not None, k
        out['gnorm/' + k] = np.array(g.norm().item())
        out['ghead/' + k] = g.reshape(-1)[:16].numpy().copy()
        out['gsum/' + k] = np.array(g.sum().item())
    for k, v in stats.items():
        out['stat/' + k] = v.numpy().copy()
    # a few full gradients that are cheap to keep and discriminate well
    pn = dict(model.named_parameters())
    last4 = max(int(k.split('.')[1]) for k in pn if k.startswith('layer4.'))
    for k in ('conv1_s.weight', 'conv1_t.weight', 'layer1.0.conv2.weight', 'layer2.1.conv2.weight',
              'layer1.0.fc1.weight', 'layer1.0.fc2.bias', f'layer4.{last4}.conv2.weight', 'bn1.weight',
              'layer3.0.downsample.1.bias', 'fc2.bias'):
        out['gfull/' + k] = grads[k].numpy().copy()

    # eval-mode logits after aggregate_sub_bn_stats (x3d.py:306-313)
    model.aggregate_sub_bn_stats()
    model.eval()
    with torch.no_grad():
        out['eval_logits'] = model(x).numpy()
    for k in ('bn1.bn.running_mean', 'bn1.bn.running_var', 'layer2.0.bn2.bn.running_var'):
        out['agg/' + k] = model.state_dict()[k].numpy().copy()
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)
    print(name, 'loss', loss.item(), 'logit absmax', logits.abs().max().item(), flush=True)


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


def input_pipeline_golden():
    """tests/golden/input_pipeline.npz: uint8 frames pushed through the REFERENCE's own spatial transform chain
    (train_x3d_kinetics_multigrid.py:70-73: MultiScaleRandomCropMultigrid -> RandomHorizontalFlip -> ToTensor(255)
    -> Normalize(KINETICS_MEAN, KINETICS_STD), transforms/spatial_transforms.py) with the crop scale chosen so that
    the window already has the target size (PIL's resize is then the identity): the fp32 clips the GPU-side input
    pipeline must reproduce bit for bit."""
    sys.path.insert(0, REF)
    from PIL import Image
    from transforms import spatial_transforms as ST
    mean = [110.63666788 / 255, 103.16065604 / 255, 96.29023126 / 255]
    std = [38.7568578 / 255, 37.88248729 / 255, 40.02898126 / 255]
    B, T, Hs, Ws, S = 3, 3, 40, 48, 32
    u = O.det_uniform(B * T * Hs * Ws * 3, 4242)
    frames = np.clip(np.floor((u * 0.5 + 0.5) * 256), 0, 255).astype(np.uint8).reshape(B, T, Hs, Ws, 3)
    crop = ST.MultiScaleRandomCropMultigrid([S / min(Hs, Ws)], S)
    flip = ST.RandomHorizontalFlip()
    chain = ST.Compose([crop, flip, ST.ToTensor(255), ST.Normalize(mean, std)])
    params = [(0.31, 0.77, 0.2), (0.0, 1.0, 0.9), (0.999, 0.5, 0.49)]        # (tl_x, tl_y, flip p) per clip
    clips, table = [], []
    for b in range(B):
        crop.size, crop.scale = S, S / min(Hs, Ws)
        crop.tl_x, crop.tl_y = params[b][0], params[b][1]
        flip.p = params[b][2]
        cs = int(min(Hs, Ws) * crop.scale)
        assert cs == S
        x1, y1 = int(crop.tl_x * (Ws - cs)), int(crop.tl_y * (Hs - cs))
        table.append([x1, y1, int(flip.p < 0.5), 0])
        per_t = [chain(Image.fromarray(frames[b, t])) for t in range(T)]     # [3,S,S] each
        clips.append(torch.stack(per_t, 1))                                  # [3,T,S,S]
    clip = torch.stack(clips, 0).numpy()
    np.savez_compressed(os.path.join(GOLD, 'input_pipeline.npz'), frames=frames, crops=np.array(table, dtype=np.int32),
                        clip=clip, mean=np.array(mean), std=np.array(std))
    print('input_pipeline', clip.shape, float(clip.mean()), table)


if __name__ == '__main__':
    ref = load_reference()
    os.makedirs(GOLD, exist_ok=True)
    manifests(ref)
    sampler_golden()
    input_pipeline_golden()
    only = sys.argv[1:] or list(CASES)
    for name in only:
        run_case(ref, name, CASES[name])
