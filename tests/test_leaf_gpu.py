"""Stand-alone leaf modules (SubBatchNorm3d and the Conv3d flavours of x3d.py) against plain fp64 math on the CPU:
the modules a user reaches as ``model.layer1[0].conv2`` etc. run the same C-ABI kernels outside the fused network."""
import pytest
import torch
import torch.nn.functional as F

from oracle import x3d_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


@pytest.mark.parametrize('stride,kernel', [(1, (3, 3, 3)), (2, (3, 3, 3)), (1, (5, 1, 1))])
def test_depthwise_module(stride, kernel):
    import x3d_multigrid_b200 as X
    C = 54
    if kernel == (3, 3, 3):
        conv = X.conv3x3x3(C, C, stride).cuda()
    else:
        m = X.generate_model('S', n_classes=5)
        conv, C = m.conv1_t.cuda(), m.conv1_t.in_channels
    x = O.det_clip((2, C, 4, 9, 11), 'leafdw', torch.float32).cuda().requires_grad_(True)
    y = conv(x)
    dy = O.det_clip(tuple(y.shape), 'leafdwdy', torch.float32).cuda()
    y.backward(dy)
    xr = x.detach().double().cpu().requires_grad_(True)
    wr = conv.weight.detach().double().cpu().requires_grad_(True)
    yr = F.conv3d(xr, wr, stride=(1, stride, stride), padding=tuple(k // 2 for k in kernel), groups=C)
    yr.backward(dy.double().cpu())
    assert rel(y, yr) < 1e-5 and rel(x.grad, xr.grad) < 1e-5 and rel(conv.weight.grad, wr.grad) < 1e-5


@pytest.mark.parametrize('stride', [1, 2])
def test_pointwise_module(stride):
    import x3d_multigrid_b200 as X
    conv = X.conv1x1x1(24, 54, stride).cuda()
    x = O.det_clip((2, 24, 3, 8, 7), 'leafpw', torch.float32).cuda().requires_grad_(True)
    y = conv(x)
    dy = O.det_clip(tuple(y.shape), 'leafpwdy', torch.float32).cuda()
    y.backward(dy)
    xr = x.detach().double().cpu().requires_grad_(True)
    wr = conv.weight.detach().double().cpu().requires_grad_(True)
    yr = F.conv3d(xr, wr, stride=(1, stride, stride))
    yr.backward(dy.double().cpu())
    assert rel(y, yr) < 1e-5 and rel(x.grad, xr.grad) < 1e-5 and rel(conv.weight.grad, wr.grad) < 1e-5


def test_stem_spatial_module():
    import x3d_multigrid_b200 as X
    m = X.generate_model('S', n_classes=5)
    conv = m.conv1_s.cuda()
    x = O.det_clip((2, 3, 3, 12, 10), 'leafstem', torch.float32).cuda()
    y = conv(x)
    dy = O.det_clip(tuple(y.shape), 'leafstemdy', torch.float32).cuda()
    y.backward(dy)
    wr = conv.weight.detach().double().cpu().requires_grad_(True)
    yr = F.conv3d(x.double().cpu(), wr, stride=(1, 2, 2), padding=(0, 1, 1))
    yr.backward(dy.double().cpu())
    assert rel(y, yr) < 1e-5 and rel(conv.weight.grad, wr.grad) < 1e-5


@pytest.mark.parametrize('training', [True, False])
def test_sub_batch_norm_module(training):
    import x3d_multigrid_b200 as X
    C, splits = 24, 2
    bn = X.SubBatchNorm3d(num_splits=splits, num_features=C, affine=True).cuda()
    with torch.no_grad():
        bn.weight.copy_(1.0 + O.det_tensor((C,), 'leafg', scale=0.3, dtype=torch.float32))
        bn.bias.copy_(O.det_tensor((C,), 'leafb', scale=0.5, dtype=torch.float32))
        bn.bn.running_mean.copy_(O.det_tensor((C,), 'leafrm', scale=0.2, dtype=torch.float32))
        bn.bn.running_var.copy_(1.0 + 0.5 * O.det_tensor((C,), 'leafrv', scale=1.0, dtype=torch.float32).abs())
    bn.train(training)
    x = O.det_clip((4, C, 3, 6, 5), 'leafbn', torch.float32).cuda().requires_grad_(True)
    y = bn(x)
    dy = O.det_clip(tuple(y.shape), 'leafbndy', torch.float32).cuda()
    y.backward(dy)
    sd = {'p.' + k: v.detach().double().cpu() for k, v in bn.state_dict().items()}
    sd['p.split_bn.running_mean'] = torch.zeros(splits * C, dtype=torch.float64)     # state BEFORE the step
    sd['p.split_bn.running_var'] = torch.ones(splits * C, dtype=torch.float64)
    g = sd['p.weight'].clone().requires_grad_(True)
    b = sd['p.bias'].clone().requires_grad_(True)
    sd['p.weight'], sd['p.bias'] = g, b
    xr = x.detach().double().cpu().requires_grad_(True)
    new = {}
    yr = O.sub_bn(xr, 'p', sd, splits, training, new)
    yr.backward(dy.double().cpu())
    assert rel(y, yr) < 1e-5 and rel(x.grad, xr.grad) < 1e-4
    assert rel(bn.weight.grad, g.grad) < 1e-5 and rel(bn.bias.grad, b.grad) < 1e-5
    if training:
        assert rel(bn.split_bn.running_mean, new['p.split_bn.running_mean']) < 1e-5
        assert rel(bn.split_bn.running_var, new['p.split_bn.running_var']) < 1e-5
        assert int(bn.split_bn.num_batches_tracked) == 1
