"""Per-kernel parity of the CUDA path (through the C ABI) against the CPU oracle (fp64).

Tolerances (relative L2): fp32 storage 1e-5 per kernel (north_star budget 1e-4 end to end);
bf16 storage 1e-2 per kernel against the oracle evaluated on the bf16-rounded inputs
(north_star budget 2e-2)."""
import numpy as np
import pytest
import torch

from oracle import x3d_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 1e-2}
DT = {torch.float32: 0, torch.bfloat16: 1}


def L():
    from x3d_multigrid_b200 import _lib
    return _lib.lib()


def stream():
    return torch.cuda.current_stream().cuda_stream


def pad8(c):
    return (c + 7) // 8 * 8


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def to_ndhwc(x, dtype):
    """x: NCDHW fp32 cuda -> NDHWC padded, through the library's converter"""
    N, C, T, H, W = x.shape
    out = torch.empty(N, T, H, W, pad8(C), dtype=dtype, device='cuda')
    L().call('x3d_ncdhw_to_ndhwc', x.data_ptr(), out.data_ptr(), N, C, pad8(C), T, H, W, DT[dtype], stream())
    return out


def to_ncdhw(y, C):
    N, T, H, W, Cp = y.shape
    out = torch.empty(N, C, T, H, W, dtype=torch.float32, device='cuda')
    L().call('x3d_ndhwc_to_ncdhw', y.data_ptr(), out.data_ptr(), N, C, Cp, T, H, W, DT[y.dtype], stream())
    return out


def q(x, dtype):
    """value as stored in the activation dtype, back in fp64 on the CPU"""
    return x.to(dtype).double().cpu()


def pack_dw(w, Cp):
    """[C,1,kt,kh,kw] -> fp32 [taps][Cp] (host-side stand-in for x3d_pack_params in kernel tests)"""
    C = w.shape[0]
    out = torch.zeros(w[0].numel(), Cp, dtype=torch.float32, device='cuda')
    out[:, :C] = w.reshape(C, -1).t().float()
    return out.contiguous()


def pack_pw(w, dtype):
    n, k = w.shape[:2]
    f = torch.zeros(pad8(n), pad8(k), dtype=dtype, device='cuda')
    f[:n, :k] = w.reshape(n, k).to(dtype)
    t = torch.zeros(pad8(k), pad8(n), dtype=dtype, device='cuda')
    t[:k, :n] = w.reshape(n, k).t().to(dtype)
    return f.contiguous(), t.contiguous()


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_layout_roundtrip(dtype):
    x = O.det_clip((2, 54, 3, 5, 7), 'lay', torch.float32).cuda()
    y = to_ndhwc(x, dtype)
    assert y.shape == (2, 3, 5, 7, 56)
    assert torch.all(y[..., 54:] == 0)
    assert torch.equal(y[..., :54].permute(0, 4, 1, 2, 3).float(), x.to(dtype).float())
    back = to_ncdhw(y, 54)
    assert torch.equal(back, x.to(dtype).float())


def test_pack_params_matches_host_pack():
    import ctypes
    from x3d_multigrid_b200._lib import PackDesc
    w = O.det_tensor((54, 24, 1, 1, 1), 'pw', dtype=torch.float32).cuda()
    d = O.det_tensor((54, 1, 3, 3, 3), 'dw', dtype=torch.float32).cuda()
    f = torch.empty(56, 24, dtype=torch.bfloat16, device='cuda')
    t = torch.empty(24, 56, dtype=torch.bfloat16, device='cuda')
    dd = torch.empty(27, 56, dtype=torch.float32, device='cuda')
    arr = (PackDesc * 3)(PackDesc(w.data_ptr(), f.data_ptr(), 54, 24, 56, 24, 0, 1),
                         PackDesc(w.data_ptr(), t.data_ptr(), 54, 24, 24, 56, 1, 1),
                         PackDesc(d.data_ptr(), dd.data_ptr(), 54, 27, 27, 56, 1, 0))
    dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    L().call('x3d_pack_params', dev.data_ptr(), 3, 56 * 27, stream())
    ef, et = pack_pw(w, torch.bfloat16)
    assert torch.equal(f, ef) and torch.equal(t, et) and torch.equal(dd, pack_dw(d, 56))


DW_CASES = [  # N, C, T, H, W, stride, kernel
    (2, 54, 3, 7, 9, 1, (3, 3, 3)),
    (2, 54, 4, 8, 10, 2, (3, 3, 3)),
    (3, 108, 2, 7, 7, 2, (3, 3, 3)),      # odd -> ceil (7 -> 4), SURVEY A2
    (2, 24, 5, 6, 5, 1, (5, 1, 1)),
    (1, 432, 4, 4, 4, 1, (3, 3, 3)),
    (2, 16, 13, 20, 12, 1, (3, 3, 3)),
    (2, 56, 3, 20, 37, 1, (3, 3, 3)),     # several ragged tiles of the tiled kernel
    (2, 56, 2, 21, 35, 2, (3, 3, 3)),
    (1, 216, 2, 14, 14, 1, (3, 3, 3)),    # channel chunks that do not divide Cp evenly
    (2, 630, 2, 10, 10, 1, (3, 3, 3)),    # X3D-XL stage 4 (Cp = 632)
    (2, 216, 3, 10, 10, 1, (3, 3, 3)),    # 158-pixel multigrid shape, stage 3 (narrow image: 4-wide tiles)
    (2, 432, 2, 5, 5, 1, (3, 3, 3)),      # 158-pixel multigrid shape, stage 4
    (2, 216, 2, 20, 20, 2, (3, 3, 3)),    # 158-pixel multigrid shape, stage 3 entry (20 -> 10)
    # production sizes (BASELINE config 2 tilings: 112^2 -> 56^2 stride-2 entry, 28^2 stride-1) and X3D-XL widths
    (2, 54, 16, 112, 112, 2, (3, 3, 3)),
    (2, 108, 16, 28, 28, 1, (3, 3, 3)),
    (2, 72, 3, 156, 156, 2, (3, 3, 3)),   # XL stage 1 entry (Cp = 72)
    (2, 162, 4, 39, 39, 1, (3, 3, 3)),    # XL stage 2 (Cp = 168, odd image)
    (2, 306, 3, 20, 20, 1, (3, 3, 3)),    # XL stage 3 (Cp = 312)
]
# shapes that must be served by the TMA-tiled kernels (every 3x3x3 case above is a hot shape of some config)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('case', DW_CASES)
@pytest.mark.parametrize('fused', [False, True])
def test_dwconv_fwd_dgrad_wgrad(case, dtype, fused):
    N, C, T, H, W, s, k = case
    Cp = pad8(C)
    splits = 2 if (fused and N % 2 == 0) else 1
    paths0 = L().path_counts()
    x = O.det_clip((N, C, T, H, W), f'dwx{case}', torch.float32).cuda()
    w = O.det_tensor((C, 1) + k, f'dww{case}', scale=0.4, dtype=torch.float32).cuda()
    xn = to_ndhwc(x, dtype)
    wp = pack_dw(w, Cp)
    xq = q(x, dtype)
    sc = sh = None
    if fused:
        sc = (1.0 + O.det_tensor((splits, Cp), 'sc', scale=0.3, dtype=torch.float32)).cuda()
        sh = O.det_tensor((splits, Cp), 'sh', scale=0.5, dtype=torch.float32).cuda()
        sc[:, C:] = 0
        sh[:, C:] = 0
        bidx = torch.arange(N) % splits
        xt = torch.relu(xq * sc.double().cpu()[bidx][:, :C].view(N, C, 1, 1, 1)
                        + sh.double().cpu()[bidx][:, :C].view(N, C, 1, 1, 1))
    else:
        xt = xq
    xt.requires_grad_(True)
    w64 = w.double().cpu().requires_grad_(True)
    ref = O.dwconv3d(xt, w64, s)
    Ho, Wo = ref.shape[3], ref.shape[4]
    y = torch.empty(N, T, Ho, Wo, Cp, dtype=dtype, device='cuda')
    stats = torch.zeros(N, Cp, 2, dtype=torch.float64, device='cuda')
    L().call('x3d_dwconv_fwd', xn.data_ptr(), wp.data_ptr(), y.data_ptr(), N, T, H, W, Cp, k[0], k[1], k[2], s,
             sc.data_ptr() if fused else None, sh.data_ptr() if fused else None, splits, 1 if fused else 0,
             stats.data_ptr(), DT[dtype], stream())
    got = to_ncdhw(y, C)
    assert rel(got, ref.detach()) < TOL[dtype]
    assert torch.all(y[..., C:] == 0)
    # epilogue statistics are those of the stored tensor
    ys = got.double().cpu()
    assert rel(stats[:, :C, 0], ys.sum(dim=(2, 3, 4))) < 1e-5
    assert rel(stats[:, :C, 1], (ys * ys).sum(dim=(2, 3, 4))) < 1e-5

    # backward
    dy = O.det_clip(tuple(ref.shape), f'dwdy{case}', torch.float32).cuda()
    dyn = to_ndhwc(dy, dtype)
    ref.backward(q(dy, dtype))
    dx = torch.empty_like(xn)
    if fused:
        bst = torch.zeros(N, Cp, 2, dtype=torch.float64, device='cuda')
        L().call('x3d_dwconv_dgrad', dyn.data_ptr(), wp.data_ptr(), dx.data_ptr(), N, T, H, W, Cp, k[0], k[1], k[2], s,
                 xn.data_ptr(), sc.data_ptr(), sh.data_ptr(), splits, bst.data_ptr(), DT[dtype], stream())
        # xt.grad is d/d(relu output); the kernel additionally applies the relu mask
        mask = (xt.detach() > 0).double()
        want = xt.grad * mask
        gdx = to_ncdhw(dx, C)
        assert rel(gdx, want) < TOL[dtype]
        gs = gdx.double().cpu()
        assert rel(bst[:, :C, 0], gs.sum(dim=(2, 3, 4))) < 1e-5
        assert rel(bst[:, :C, 1], (gs * xq).sum(dim=(2, 3, 4))) < 1e-5
    else:
        L().call('x3d_dwconv_dgrad', dyn.data_ptr(), wp.data_ptr(), dx.data_ptr(), N, T, H, W, Cp, k[0], k[1], k[2], s,
                 None, None, None, 1, None, DT[dtype], stream())
        assert rel(to_ncdhw(dx, C), xt.grad) < TOL[dtype]
    dw = torch.zeros_like(w)
    L().call('x3d_dwconv_wgrad', xn.data_ptr(), dyn.data_ptr(), dw.data_ptr(), N, T, H, W, C, Cp, k[0], k[1], k[2], s,
             sc.data_ptr() if fused else None, sh.data_ptr() if fused else None, splits, 1 if fused else 0,
             DT[dtype], stream())
    assert rel(dw, w64.grad) < TOL[dtype]
    # which kernel family served the three calls: 3x3x3 -> TMA-tiled, 5x1x1 -> streaming temporal (never the
    # shape-generic direct kernels for these hot shapes)
    d = {kk: v - paths0[kk] for kk, v in L().path_counts().items()}
    fam = 'tiled' if k == (3, 3, 3) else 'temporal'
    if k == (3, 3, 3) or not fused:
        assert d[f'dw_fwd_{fam}'] == 1 and d[f'dw_dgrad_{fam}'] == 1 and d[f'dw_wgrad_{fam}'] == 1, d


PW_CASES = [  # N, K, Nout, T, H, W, stride
    (2, 24, 54, 3, 6, 7, 1),
    (2, 54, 24, 2, 5, 5, 1),
    (2, 24, 48, 3, 7, 9, 2),       # downsample, odd sizes
    (2, 24, 24, 4, 20, 18, 2),     # downsample: several 128-row tiles of gathered rows (tensor-core path, K < 64)
    (3, 48, 96, 2, 14, 15, 2),
    (2, 96, 192, 2, 6, 6, 2),      # K >= 64: strided rows stay on the SIMT kernel
    (3, 192, 432, 2, 4, 4, 1),
    (2, 432, 192, 2, 3, 3, 1),
    (4, 96, 216, 1, 2, 2, 1),      # tiny P: tiles span several samples
    (70, 8, 16, 1, 1, 1, 1),       # P = 1: more than MAXS samples per tile
    # production size: M = 401 k rows -> 3136 tiles on <= 444 persistent CTAs, statistics flushed per sample
    (2, 24, 54, 16, 112, 112, 1),
    (2, 54, 24, 16, 56, 56, 1),
    # X3D-XL widths (K, N in {72, 136, 162, 280, 306, 630})
    (2, 72, 162, 4, 39, 39, 1),
    (2, 306, 136, 2, 20, 20, 1),
    (2, 280, 630, 2, 10, 10, 1),
    (2, 630, 280, 2, 10, 10, 1),
    (2, 32, 72, 3, 40, 40, 2),     # XL downsample branch (K < 64: gathered rows on the tensor-core path)
]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('case', PW_CASES)
def test_pwconv_fwd_dgrad_wgrad(case, dtype):
    N, K, Nn, T, H, W, s = case
    Kp, Np = pad8(K), pad8(Nn)
    paths0 = L().path_counts()
    x = O.det_clip((N, K, T, H, W), f'pwx{case}', torch.float32).cuda()
    w = O.det_tensor((Nn, K, 1, 1, 1), f'pww{case}', scale=(3.0 / K) ** 0.5, dtype=torch.float32).cuda()
    xn = to_ndhwc(x, dtype)
    wf, wt = pack_pw(w, dtype)
    x64 = q(x, dtype).requires_grad_(True)
    w64 = q(w, dtype).requires_grad_(True)
    ref = O.pwconv(x64, w64, s)
    Ho, Wo = ref.shape[3], ref.shape[4]
    y = torch.empty(N, T, Ho, Wo, Np, dtype=dtype, device='cuda')
    stats = torch.zeros(N, Np, 2, dtype=torch.float64, device='cuda')
    L().call('x3d_pwconv_fwd', xn.data_ptr(), wf.data_ptr(), y.data_ptr(), N, T, H, W, Kp, Np, s, stats.data_ptr(),
             DT[dtype], stream())
    got = to_ncdhw(y, Nn)
    assert rel(got, ref.detach()) < TOL[dtype]
    assert torch.all(y[..., Nn:] == 0)
    ys = got.double().cpu()
    assert rel(stats[:, :Nn, 0], ys.sum(dim=(2, 3, 4))) < 1e-5
    assert rel(stats[:, :Nn, 1], (ys * ys).sum(dim=(2, 3, 4))) < 1e-5

    dy = O.det_clip(tuple(ref.shape), f'pwdy{case}', torch.float32).cuda()
    dyn = to_ndhwc(dy, dtype)
    ref.backward(q(dy, dtype))
    dx = torch.zeros_like(xn)
    L().call('x3d_pwconv_dgrad', dyn.data_ptr(), wt.data_ptr(), dx.data_ptr(), N, T, H, W, Kp, Np, s, 0, DT[dtype],
             stream())
    assert rel(to_ncdhw(dx, K), x64.grad) < TOL[dtype]
    # accumulate flavour: dx += ...
    base = O.det_clip((N, K, T, H, W), 'base', torch.float32).cuda()
    dx2 = to_ndhwc(base, dtype)
    L().call('x3d_pwconv_dgrad', dyn.data_ptr(), wt.data_ptr(), dx2.data_ptr(), N, T, H, W, Kp, Np, s, 1, DT[dtype],
             stream())
    assert rel(to_ncdhw(dx2, K), x64.grad + q(base, dtype)) < 2 * TOL[dtype]
    dw = torch.zeros(Nn, K, dtype=torch.float32, device='cuda')
    L().call('x3d_pwconv_wgrad', xn.data_ptr(), dyn.data_ptr(), dw.data_ptr(), N, T, H, W, K, Kp, Nn, Np, s, DT[dtype],
             stream())
    assert rel(dw, w64.grad.reshape(Nn, K)) < TOL[dtype]
    # two-stage variant (scratch buffer, ordered second stage): same result, bit-identical from run to run, and it
    # ACCUMULATES into dw like the plain entry point
    ws = torch.empty(8 << 20, dtype=torch.uint8, device='cuda')
    dws = []
    for _ in range(2):
        d2 = torch.ones(Nn, K, dtype=torch.float32, device='cuda')
        L().call('x3d_pwconv_wgrad_ws', xn.data_ptr(), dyn.data_ptr(), d2.data_ptr(), N, T, H, W, K, Kp, Nn, Np, s,
                 ws.data_ptr(), ws.numel(), DT[dtype], stream())
        dws.append(d2)
    assert rel(dws[0] - 1.0, w64.grad.reshape(Nn, K)) < 2 * TOL[dtype]
    if dtype == torch.bfloat16 and s == 1:
        assert torch.equal(dws[0], dws[1])
    d = {kk: v - paths0[kk] for kk, v in L().path_counts().items()}
    if dtype == torch.bfloat16:      # bf16 storage runs the tcgen05 kernels (fp32 storage is the SIMT parity path)
        assert d['pw_dgrad_tc'] == 2 and d['pw_dgrad_simt'] == 0, d
        assert d['pw_fwd_tc'] == 1 or (s == 2 and K >= 64), d
        assert d['pw_wgrad_tc'] == 3 or (s == 2 and K >= 64), d
    else:
        assert d['pw_fwd_simt'] == 1 and d['pw_wgrad_simt'] == 3, d


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('hw', [(11, 14), (12, 16), (21, 72), (40, 140)])   # W % 4 == 0: cp.async-pipelined wgrad kernel
def test_stem_conv_s(dtype, hw):
    N, T, Co = 2, 3, 24
    H, W = hw
    x = O.det_clip((N, 3, T, H, W), 'stemx', torch.float32).cuda()
    w = O.det_tensor((Co, 3, 1, 3, 3), 'stemw', scale=0.3, dtype=torch.float32).cuda()
    w64 = w.double().cpu().requires_grad_(True)
    ref = O.stem_conv_s(x.double().cpu(), w64)
    Ho, Wo = ref.shape[3], ref.shape[4]
    y = torch.empty(N, T, Ho, Wo, Co, dtype=dtype, device='cuda')
    L().call('x3d_stem_conv_s_fwd', x.data_ptr(), w.data_ptr(), y.data_ptr(), N, 3, T, H, W, Co, Co, DT[dtype], stream())
    assert rel(to_ncdhw(y, Co), ref.detach()) < TOL[dtype]
    dy = O.det_clip(tuple(ref.shape), 'stemdy', torch.float32).cuda()
    ref.backward(q(dy, dtype))
    dw = torch.zeros_like(w)
    L().call('x3d_stem_conv_s_wgrad', x.data_ptr(), to_ndhwc(dy, dtype).data_ptr(), dw.data_ptr(), N, 3, T, H, W, Co, Co,
             DT[dtype], stream())
    assert rel(dw, w64.grad) < TOL[dtype]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('splits', [1, 2])
def test_split_bn_forward_backward(dtype, splits):
    """SubBatchNorm3d train-mode forward (+running stats) and two-pass backward vs the oracle."""
    N, C, T, H, W = 4, 54, 2, 5, 6
    Cp, P = pad8(C), T * H * W
    x = (O.det_clip((N, C, T, H, W), 'bnx', torch.float32) * 1.7 + 0.3).cuda()
    gamma = (1 + O.det_tensor((C,), 'bng', scale=0.2, dtype=torch.float32)).cuda()
    beta = O.det_tensor((C,), 'bnb', scale=0.2, dtype=torch.float32).cuda()
    res = O.det_clip((N, C, T, H, W), 'bnres', torch.float32).cuda()
    xn, rn = to_ndhwc(x, dtype), to_ndhwc(res, dtype)
    x64 = q(x, dtype).requires_grad_(True)
    g64, b64 = gamma.double().cpu().requires_grad_(True), beta.double().cpu().requires_grad_(True)
    sd = {'bn.weight': g64, 'bn.bias': b64}
    new_stats = {}
    yref = torch.relu(O.sub_bn(x64, 'bn', sd, splits, True, new_stats) + q(res, dtype))
    # statistics through the reduce kernel (dout = a = x gives sum x, sum x^2)
    stats = torch.zeros(N, Cp, 2, dtype=torch.float64, device='cuda')
    L().call('x3d_bn_bwd_reduce', xn.data_ptr(), None, xn.data_ptr(), stats.data_ptr(), N, P, Cp, DT[dtype], stream())
    rm = torch.zeros(splits * C, device='cuda')
    rv = torch.ones(splits * C, device='cuda')
    nbt = torch.zeros((), dtype=torch.int64, device='cuda')
    buf = torch.empty(4, splits, Cp, device='cuda')
    L().call('x3d_bn_finalize', stats.data_ptr(), N, splits, P, C, Cp, gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(),
             rv.data_ptr(), nbt.data_ptr(), 0.1, 1e-5, buf[0].data_ptr(), buf[1].data_ptr(), buf[2].data_ptr(),
             buf[3].data_ptr(), stream())
    assert int(nbt.item()) == 1
    assert rel(rm, new_stats['bn.split_bn.running_mean']) < 1e-6
    assert rel(rv, new_stats['bn.split_bn.running_var']) < 1e-6
    out = torch.empty_like(xn)
    L().call('x3d_bn_act_fwd', xn.data_ptr(), buf[0].data_ptr(), buf[1].data_ptr(), splits, rn.data_ptr(), None, None, 1,
             out.data_ptr(), N, P, Cp, DT[dtype], stream())
    assert rel(to_ncdhw(out, C), yref.detach()) < TOL[dtype]
    # backward
    dy = O.det_clip((N, C, T, H, W), 'bndy', torch.float32).cuda()
    dyn = to_ndhwc(dy, dtype)
    yref.backward(q(dy, dtype))
    bst = torch.zeros(N, Cp, 2, dtype=torch.float64, device='cuda')
    L().call('x3d_bn_bwd_reduce', dyn.data_ptr(), out.data_ptr(), xn.data_ptr(), bst.data_ptr(), N, P, Cp, DT[dtype],
             stream())
    coef = torch.empty(3, splits, Cp, device='cuda')
    dg, db = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    L().call('x3d_bn_bwd_finalize', bst.data_ptr(), N, splits, P, C, Cp, gamma.data_ptr(), buf[2].data_ptr(),
             buf[3].data_ptr(), 1, coef.data_ptr(), dg.data_ptr(), db.data_ptr(), stream())
    da = torch.empty_like(xn)
    L().call('x3d_bn_bwd_apply', dyn.data_ptr(), out.data_ptr(), xn.data_ptr(), coef.data_ptr(), splits, da.data_ptr(), N,
             P, Cp, DT[dtype], stream())
    tol = TOL[dtype] * (3 if dtype == torch.bfloat16 else 10)   # relu-mask flips on rounded outputs
    assert rel(to_ncdhw(da, C), x64.grad) < tol
    assert rel(dg, g64.grad) < tol and rel(db, b64.grad) < tol


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('with_se', [True, False])
@pytest.mark.parametrize('N', [4, 44])       # 44 samples: SE parameter gradients summed by lane groups (multigrid batches)
def test_se_swish_forward_backward(dtype, with_se, N):
    """bn2 -> SE gate -> swish (x3d.py:151-160) forward and backward incl. SE/BN parameter grads."""
    C, T, H, W, sw, splits = 54, 2, 4, 5, 8, 2
    Cp, P = pad8(C), T * H * W
    a2 = (O.det_clip((N, C, T, H, W), 'sea', torch.float32) * 1.3 - 0.2).cuda()
    gamma = (1 + O.det_tensor((C,), 'seg', scale=0.2, dtype=torch.float32)).cuda()
    beta = O.det_tensor((C,), 'seb', scale=0.2, dtype=torch.float32).cuda()
    W1 = O.det_tensor((sw, C, 1, 1, 1), 'sew1', scale=0.3, dtype=torch.float32).cuda()
    b1 = O.det_tensor((sw,), 'seb1', scale=0.1, dtype=torch.float32).cuda()
    W2 = O.det_tensor((C, sw, 1, 1, 1), 'sew2', scale=0.5, dtype=torch.float32).cuda()
    b2 = O.det_tensor((C,), 'seb2', scale=0.1, dtype=torch.float32).cuda()
    an = to_ndhwc(a2, dtype)
    leaves = {k: v.double().cpu().requires_grad_(True) for k, v in
              dict(g=gamma, b=beta, W1=W1, b1=b1, W2=W2, b2=b2).items()}
    a64 = q(a2, dtype).requires_grad_(True)
    u = O.sub_bn(a64, 'bn', {'bn.weight': leaves['g'], 'bn.bias': leaves['b']}, splits, True, None)
    if with_se:
        se = u.mean(dim=(2, 3, 4), keepdim=True)
        se = torch.relu(O.pwconv(se, leaves['W1'], 1, leaves['b1']))
        se = torch.sigmoid(O.pwconv(se, leaves['W2'], 1, leaves['b2']))
        u = u * se
    vref = O.swish(u)
    # forward through the kernels
    st2 = torch.zeros(N, Cp, 2, dtype=torch.float64, device='cuda')
    L().call('x3d_bn_bwd_reduce', an.data_ptr(), None, an.data_ptr(), st2.data_ptr(), N, P, Cp, DT[dtype], stream())
    buf = torch.empty(4, splits, Cp, device='cuda')
    L().call('x3d_bn_finalize', st2.data_ptr(), N, splits, P, C, Cp, gamma.data_ptr(), beta.data_ptr(), None, None, None,
             0.1, 1e-5, buf[0].data_ptr(), buf[1].data_ptr(), buf[2].data_ptr(), buf[3].data_ptr(), stream())
    pooled = hidden = gate = None
    if with_se:
        pooled = torch.empty(N, C, device='cuda')
        hidden = torch.empty(N, sw, device='cuda')
        gate = torch.empty(N, Cp, device='cuda')
        L().call('x3d_se_fwd', st2.data_ptr(), buf[0].data_ptr(), buf[1].data_ptr(), splits, N, P, C, Cp, sw,
                 W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), pooled.data_ptr(), hidden.data_ptr(),
                 gate.data_ptr(), stream())
    v = torch.empty_like(an)
    gp = gate.data_ptr() if with_se else None
    L().call('x3d_swish_gate_fwd', an.data_ptr(), buf[0].data_ptr(), buf[1].data_ptr(), splits, gp, v.data_ptr(), N, P, Cp,
             DT[dtype], stream())
    assert rel(to_ncdhw(v, C), vref.detach()) < TOL[dtype]
    # backward
    dv = O.det_clip((N, C, T, H, W), 'sedv', torch.float32).cuda()
    dvn = to_ndhwc(dv, dtype)
    vref.backward(q(dv, dtype))
    bst = torch.zeros(N, Cp, 2, dtype=torch.float64, device='cuda')
    L().call('x3d_swish_gate_bwd_reduce', dvn.data_ptr(), an.data_ptr(), buf[0].data_ptr(), buf[1].data_ptr(), splits, gp,
             bst.data_ptr(), N, P, Cp, DT[dtype], stream())
    gW1, gb1 = torch.zeros(sw, C, device='cuda'), torch.zeros(sw, device='cuda')
    gW2, gb2 = torch.zeros(C, sw, device='cuda'), torch.zeros(C, device='cuda')
    dg, db = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    work = torch.empty(N, 2 * Cp + max(sw, 1), device='cuda')
    coef = torch.empty(N, Cp, 3, device='cuda')
    if with_se:
        se_args = (W1.data_ptr(), W2.data_ptr(), pooled.data_ptr(), hidden.data_ptr(), gate.data_ptr(), gW1.data_ptr(),
                   gb1.data_ptr(), gW2.data_ptr(), gb2.data_ptr())
    else:
        se_args = (None,) * 9
    L().call('x3d_se_bn_bwd', st2.data_ptr(), bst.data_ptr(), N, splits, P, C, Cp, sw if with_se else 0, gamma.data_ptr(),
             buf[2].data_ptr(), buf[3].data_ptr(), buf[0].data_ptr(), buf[1].data_ptr(), 1, *se_args, dg.data_ptr(),
             db.data_ptr(), work.data_ptr(), coef.data_ptr(), stream())
    da = torch.empty_like(an)
    L().call('x3d_swish_gate_bwd_apply', dvn.data_ptr(), an.data_ptr(), buf[0].data_ptr(), buf[1].data_ptr(), splits, gp,
             coef.data_ptr(), da.data_ptr(), N, P, Cp, DT[dtype], stream())
    tol = TOL[dtype] * 3
    assert rel(to_ncdhw(da, C), a64.grad) < tol
    assert rel(dg, leaves['g'].grad) < tol and rel(db, leaves['b'].grad) < tol
    if with_se:
        assert rel(gW1, leaves['W1'].grad.reshape(sw, C)) < tol and rel(gb1, leaves['b1'].grad) < tol
        assert rel(gW2, leaves['W2'].grad.reshape(C, sw)) < tol and rel(gb2, leaves['b2'].grad) < tol


def test_small_gemm_and_sgd():
    A = O.det_tensor((5, 70), 'ga', dtype=torch.float32).cuda()
    B = O.det_tensor((33, 70), 'gb', dtype=torch.float32).cuda()
    bias = O.det_tensor((33,), 'gbias', dtype=torch.float32).cuda()
    C = torch.empty(5, 33, device='cuda')
    L().call('x3d_small_gemm', A.data_ptr(), 70, 1, B.data_ptr(), 1, 70, C.data_ptr(), 33, 5, 33, 70, bias.data_ptr(), 1,
             None, 0, stream())
    want = torch.relu(A.double() @ B.double().t() + bias.double())
    assert rel(C, want) < 1e-6
    # transposed-A accumulate: C2 += A^T @ A
    C2 = torch.ones(70, 70, device='cuda')
    L().call('x3d_small_gemm', A.data_ptr(), 1, 70, A.data_ptr(), 70, 1, C2.data_ptr(), 70, 70, 70, 5, None, 0, None, 1,
             stream())
    assert rel(C2, 1 + A.double().t() @ A.double()) < 1e-6
    # the four head-GEMM flavours at a large-batch (multigrid) size: M = 200 rows is not "skinny" -> tiled kernel
    R, F1, C5, ncls = 200, 150, 70, 37
    X = O.det_tensor((R, C5), 'hx', dtype=torch.float32).cuda()
    W1 = O.det_tensor((F1, C5), 'hw1', dtype=torch.float32).cuda()
    Dl = O.det_tensor((R, ncls), 'hdl', dtype=torch.float32).cuda()
    W2 = O.det_tensor((ncls, F1), 'hw2', dtype=torch.float32).cuda()
    msk = (O.det_tensor((R, F1), 'hm', dtype=torch.float32) > 0).float().cuda() * 2
    H1 = torch.empty(R, F1, device='cuda')          # NT + relu + dropout mask
    L().call('x3d_small_gemm', X.data_ptr(), C5, 1, W1.data_ptr(), 1, C5, H1.data_ptr(), F1, R, F1, C5, None, 1,
             msk.data_ptr(), 0, stream())
    assert rel(H1, torch.relu(X.double() @ W1.double().t()) * msk.double()) < 1e-6
    G2 = torch.zeros(ncls, F1, device='cuda')       # TN, accumulate: dW2 = Dl^T H1
    L().call('x3d_small_gemm', Dl.data_ptr(), 1, ncls, H1.data_ptr(), F1, 1, G2.data_ptr(), F1, ncls, F1, R, None, 0, None, 1,
             stream())
    assert rel(G2, Dl.double().t() @ H1.double()) < 1e-6
    Dh = torch.empty(R, F1, device='cuda')          # NN: dh = Dl W2
    L().call('x3d_small_gemm', Dl.data_ptr(), ncls, 1, W2.data_ptr(), F1, 1, Dh.data_ptr(), F1, R, F1, ncls, None, 0, None, 0,
             stream())
    assert rel(Dh, Dl.double() @ W2.double()) < 1e-6
    # fused SGD == torch.optim.SGD
    from x3d_multigrid_b200._lib import SgdDesc
    p = O.det_tensor((1000,), 'p', dtype=torch.float32).cuda()
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.SGD([ref_p], lr=0.1, momentum=0.9, weight_decay=5e-5)
    mom = torch.zeros_like(p)
    for step in range(3):
        g = O.det_tensor((1000,), f'g{step}', dtype=torch.float32).cuda()
        ref_p.grad = g.clone()
        opt.step()
        arr = (SgdDesc * 1)(SgdDesc(p.data_ptr(), g.data_ptr(), mom.data_ptr(), 1000))
        dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
        L().call('x3d_sgd_step', dev.data_ptr(), 1, 1000, 0.1, 0.9, 5e-5, 1.0, 1 if step == 0 else 0, stream())
        torch.cuda.synchronize()
    assert rel(p, ref_p.detach()) < 1e-6


@pytest.mark.parametrize('R', [16, 24, 64, 200])
def test_head_gemms_split_k(R):
    """The six fc1 / fc2 GEMMs of the X3D-M head (x3d.py:338-345) through x3d_small_gemm_ws at the production widths
    (432 -> 2048 -> 400): values against fp64, bit-identical from call to call (ordered split-K reduction), the
    workspace tickets back at zero, and the same answer as the workspace-free entry point."""
    C5, F1, ncls = 432, 2048, 400
    g = torch.Generator().manual_seed(R)
    pooled = torch.randn(R, C5, generator=g).cuda()
    w1 = (torch.randn(F1, C5, generator=g) * 0.05).cuda()
    w2 = (torch.randn(ncls, F1, generator=g) * 0.05).cuda()
    b2 = torch.randn(ncls, generator=g).cuda()
    dl = torch.randn(R, ncls, generator=g).cuda()
    msk = (torch.rand(R, F1, generator=g) > 0.5).float().cuda() * 2
    ws = torch.zeros(16 << 20, dtype=torch.uint8, device='cuda')
    need = L().fn['x3d_small_gemm_workspace_bytes'](R, ncls, F1)
    assert 4096 < need <= ws.numel()

    def gemm(args, out_shape, accumulate=False):
        outs = []
        for use_ws in (True, True, False):
            out = torch.full(out_shape, 0.5 if accumulate else float('nan'), device='cuda')
            if use_ws:
                L().call('x3d_small_gemm_ws', *args(out), ws.data_ptr(), ws.numel(), stream())
            else:
                L().call('x3d_small_gemm', *args(out), stream())
            outs.append(out)
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]), 'split-K result differs between two identical calls'
        assert int(ws[:4096].view(torch.int32).abs().sum()) == 0, 'tickets not reset'
        assert rel(outs[0], outs[2]) < 1e-6
        return outs[0]

    p = lambda t: t.data_ptr()
    h1 = gemm(lambda o: (p(pooled), C5, 1, p(w1), 1, C5, p(o), F1, R, F1, C5, None, 1, p(msk), 0), (R, F1))
    assert rel(h1, torch.relu(pooled.double() @ w1.double().t()) * msk.double()) < 1e-6
    lg = gemm(lambda o: (p(h1), F1, 1, p(w2), 1, F1, p(o), ncls, R, ncls, F1, p(b2), 0, None, 0), (R, ncls))
    assert rel(lg, h1.double() @ w2.double().t() + b2.double()) < 1e-6
    g2 = gemm(lambda o: (p(dl), 1, ncls, p(h1), F1, 1, p(o), F1, ncls, F1, R, None, 0, None, 1), (ncls, F1), True)
    assert rel(g2, 0.5 + dl.double().t() @ h1.double()) < 1e-6
    dh = gemm(lambda o: (p(dl), ncls, 1, p(w2), F1, 1, p(o), F1, R, F1, ncls, None, 0, None, 0), (R, F1))
    assert rel(dh, dl.double() @ w2.double()) < 1e-6
    g1 = gemm(lambda o: (p(dh), 1, F1, p(pooled), C5, 1, p(o), C5, F1, C5, R, None, 0, None, 1), (F1, C5), True)
    assert rel(g1, 0.5 + dh.double().t() @ pooled.double()) < 1e-6
    dp = gemm(lambda o: (p(dh), F1, 1, p(w1), C5, 1, p(o), C5, R, C5, F1, None, 0, None, 0), (R, C5))
    assert rel(dp, dh.double() @ w1.double()) < 1e-6
    # a workspace that only holds the tickets degrades to no split, not to an error
    tiny = torch.zeros(4096, dtype=torch.uint8, device='cuda')
    out = torch.empty(R, C5, device='cuda')
    L().call('x3d_small_gemm_ws', p(dh), F1, 1, p(w1), C5, 1, p(out), C5, R, C5, F1, None, 0, None, 0, tiny.data_ptr(), 4096,
             stream())
    assert rel(out, dp) < 1e-6
