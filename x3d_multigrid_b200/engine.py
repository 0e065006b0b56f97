"""Host-side orchestration of the X3D forward / backward over the C ABI (libx3d_b200.so).

PyTorch is used here for device memory (torch.empty), streams and nothing else: every
arithmetic step of the network is a kernel of the shared library.  Dataflow per Bottleneck
(x3d.py:143-171), NDHWC, "raw" = conv output before its SubBatchNorm3d:

  x --pw GEMM(+stats)--> a1 --[bn1+relu fused on load] dw 3x3x3 (+stats,+SE sums)--> a2
    --bn2 finalize, SE fc--> swish(gate*bn2(a2)) = v --pw GEMM(+stats)--> a3
    --relu(bn3(a3) + residual)--> out

Backward mirrors it with two-pass BN backward (reduce -> tiny finalize -> apply); all
parameter gradients are accumulated by the kernels into one flat fp32 buffer laid out in
reverse-stage bucket order (so a data-parallel allreduce can start while earlier stages
are still running backward).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import BF16, F32, PackDesc
from .input_pipeline import UInt8Clips

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def pad8(c: int) -> int:
    return (c + 7) // 8 * 8


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


@dataclass
class BNState:
    scale: torch.Tensor    # [s][Cp] fp32
    shift: torch.Tensor
    mean: torch.Tensor
    rstd: torch.Tensor
    splits: int
    train: bool


class Arena:
    """Zero-initialised bump allocator for the small per-step statistic buffers.

    The first pass of a given step signature measures the need with individual torch.zeros; later passes carve
    ONE buffer that is cleared with a single memset.  Every ``begin()`` takes a FRESH buffer from the caching
    allocator instead of recycling the previous one: statistics saved for the backward pass (``st2`` of the SE
    blocks) stay intact when another forward runs before that backward, and a captured CUDA graph keeps the
    buffer it was captured with (allocations made during capture live in the graph's private pool), whatever
    batch sizes are visited afterwards.  ``frame()`` returns the tensors of the current pass; whoever needs them
    later (the autograd save record) holds on to that."""

    def __init__(self, device):
        self.device = device
        self.buf: Optional[torch.Tensor] = None
        self.off = 0
        self.need = 0
        self.keep: List[torch.Tensor] = []

    def begin(self):
        self.buf = torch.zeros(self.need, dtype=torch.uint8, device=self.device) if self.need > 0 else None
        self.off = 0
        self.keep = []

    def zeros(self, nbytes: int) -> torch.Tensor:
        nbytes = (nbytes + 255) // 256 * 256
        if self.buf is not None and self.off + nbytes <= self.buf.numel():
            t = self.buf[self.off:self.off + nbytes]
            self.off += nbytes
        else:
            t = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            self.keep.append(t)
            self.off += nbytes
        self.need = max(self.need, self.off)
        return t

    def frame(self):
        """the tensors backing everything handed out since ``begin()`` (keep a reference to keep them alive)"""
        return (self.buf, self.keep)


class _ParamRef:
    """One nn.Parameter of the model plus its slot in the flat gradient buffer."""
    __slots__ = ('name', 'param', 'goff', 'numel')

    def __init__(self, name, param, goff):
        self.name, self.param, self.goff, self.numel = name, param, goff, param.numel()


class Engine:
    """Runs ResNet.forward / backward of ``model`` (x3d_multigrid_b200.x3d.ResNet) on CUDA."""

    def __init__(self, model, dtype: torch.dtype):
        assert dtype in (torch.float32, torch.bfloat16)
        self.model = model
        self.dtype = dtype
        self.dt = BF16 if dtype == torch.bfloat16 else F32
        self.lib = _lib.lib()
        self.device = None
        self._sig = None
        self.arena_f: Optional[Arena] = None
        self.arena_b: Optional[Arena] = None
        self._side = None
        self.wg_ws = None           # scratch buffer of the two-stage pointwise weight-gradient reduction
        self.use_side = True        # parallel.DistributedX3D switches it off in eager mode (see there)
        self.grad_hook: Optional[Callable[[int], None]] = None   # called after each bucket's grads are final

    @property
    def side(self):
        """second stream for work that feeds nothing downstream (weight gradients, the downsample branch)"""
        return self._side if self.use_side else None

    @side.setter
    def side(self, value):        # bench.py's per-kernel timing parks the stream (None) and puts it back
        if value is None:
            self.use_side = False
        else:
            self._side, self.use_side = value, True

    # ------------------------------------------------------------------ parameter tables
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _signature(self):
        return tuple((p.data_ptr(), tuple(p.shape)) for p in self.model.parameters())

    def prepare(self, device):
        """(Re)build the parameter tables when parameters moved / were replaced."""
        sig = self._signature()
        if self._sig == sig and self.device == device:
            return
        self._sig = sig
        self.device = device
        m = self.model
        for p in m.parameters():
            if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('all parameters must be contiguous fp32 tensors on the input device')
        self.arena_f, self.arena_b = Arena(device), Arena(device)
        self.wg_ws = None
        if device.type == 'cuda' and self.dt == BF16 and not os.environ.get('X3D_WG_ATOMIC'):
            self.wg_ws = torch.empty(int(self.lib.fn['x3d_pwconv_wgrad_workspace_bytes']()), dtype=torch.uint8, device=device)
        # split-K scratch of the fp32 head GEMMs (main stream only); its ticket area must start out zero
        self.head_ws = torch.zeros(16 << 20, dtype=torch.uint8, device=device)
        # weight-gradient kernels feed nothing downstream: they run on a side stream, concurrently with the
        # dgrad / BN chain of the main stream (fills the SMs that the small late-stage kernels leave idle)
        self._side = torch.cuda.Stream(device) if (device.type == 'cuda' and not os.environ.get('X3D_NO_SIDE')) else None

        # ---- flat gradient buffer, bucket order: head+stage4 first ... stem last
        named = dict(m.named_parameters())
        self.is_net = hasattr(m, 'conv1_s')
        if self.is_net:
            head = [k for k in named if k.split('.')[0] in ('conv5', 'bn5', 'fc1', 'fc2')]
            buckets = [head + [k for k in named if k.startswith('layer4.')],
                       [k for k in named if k.startswith('layer3.')],
                       [k for k in named if k.startswith('layer2.')],
                       [k for k in named if k.startswith('layer1.')] +
                       [k for k in named if k.split('.')[0] in ('conv1_s', 'conv1_t', 'bn1')]]
        else:
            buckets = [list(named)]
        self.refs: Dict[str, _ParamRef] = {}
        off = 0
        self.bucket_ranges: List[Tuple[int, int]] = []
        for b in buckets:
            start = off
            for k in b:
                self.refs[k] = _ParamRef(k, named[k], off)
                off += (named[k].numel() + 3) // 4 * 4
            self.bucket_ranges.append((start, off))
        assert len(self.refs) == len(named), 'unassigned parameters'
        self.gflat_numel = off
        self.gflat = torch.zeros(off, dtype=torch.float32, device=device)
        self.param_order = [self.refs[k] for k in named]          # named_parameters() order

        # ---- packed operand buffers
        esz = 2 if self.dt == BF16 else 4
        descs: List[PackDesc] = []
        plan: List[Tuple[str, str, int, int, int, int, int, int, int]] = []
        cur = 0

        def reserve(nbytes):
            nonlocal cur
            o = cur
            cur += (nbytes + 255) // 256 * 256
            return o

        self.packed: Dict[str, Tuple[int, int, int]] = {}   # key -> (byte offset, rows, cols)

        def add_pw(name):     # [N][K][1,1,1] -> Wf [Np][Kp] and Wt [Kp][Np] in activation dtype
            w = named[name + '.weight']
            n, k = w.shape[0], w.shape[1]
            np_, kp = pad8(n), pad8(k)
            of = reserve(np_ * kp * esz)
            ot = reserve(np_ * kp * esz)
            plan.append((name + '.weight', 'f', of, n, k, np_, kp, 0, self.dt))
            plan.append((name + '.weight', 't', ot, n, k, kp, np_, 1, self.dt))
            self.packed[name + '.f'] = (of, np_, kp)
            self.packed[name + '.t'] = (ot, kp, np_)

        def add_dw(name):     # [C][1][kt][kh][kw] -> [taps][Cp] fp32
            w = named[name + '.weight']
            c, taps = w.shape[0], w.shape[2] * w.shape[3] * w.shape[4]
            cp = pad8(c)
            o = reserve(taps * cp * 4)
            plan.append((name + '.weight', 'd', o, c, taps, taps, cp, 1, F32))
            self.packed[name + '.d'] = (o, taps, cp)

        if self.is_net:
            add_dw('conv1_t')
        for blk in m.blocks():
            add_pw(blk.prefix + '.conv1')
            add_dw(blk.prefix + '.conv2')
            add_pw(blk.prefix + '.conv3')
            if blk.downsample is not None:
                add_pw(blk.prefix + '.downsample.0')
        if self.is_net:
            add_pw('conv5')
        self.pack_buf = torch.zeros(max(cur, 256), dtype=torch.uint8, device=device)
        base = self.pack_buf.data_ptr()
        arr = (PackDesc * len(plan))()
        max_elems = 1
        for i, (pname, _kind, o, rows, cols, drows, dcols, tr, dtp) in enumerate(plan):
            arr[i] = PackDesc(named[pname].data_ptr(), base + o, rows, cols, drows, dcols, tr, dtp)
            max_elems = max(max_elems, drows * dcols)
        raw = bytes(arr)
        self.pack_descs = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        self.n_pack = len(plan)
        self.pack_max = max_elems

    def pk(self, key):
        return self.pack_buf.data_ptr() + self.packed[key][0]

    def g(self, name):
        """device pointer of the gradient slot of parameter ``name``"""
        r = self.refs[name]
        return self.gflat.data_ptr() + 4 * r.goff

    def p(self, name):
        return self.refs[name].param.data_ptr()

    def _gflat_has_outside_views(self) -> bool:
        """does anybody else hold a view of the flat gradient buffer?  (.grad of the parameters after a backward pass,
        or gradients an earlier backward of the SAME autograd graph has returned and autograd has not yet accumulated,
        e.g. (loss1 + loss2).backward() over two forward passes)"""
        try:
            # the tensor itself + the temporary storage handle made here = 2
            return torch._C._storage_Use_Count(self.gflat.untyped_storage()._cdata) > 2
        except Exception:       # private API missing: be conservative
            return True

    def new_grad_buffer(self):
        """Zeroed flat gradient buffer for this backward pass.  The persistent buffer is reused (stable addresses:
        fused optimizer tables and CUDA graphs stay valid) unless views of it are still alive outside the engine --
        then a caller is accumulating, or autograd still owes the previous result to the parameters, and a fresh
        buffer is taken (whoever holds the old views keeps the old buffer alive)."""
        if self._gflat_has_outside_views():
            self.gflat = torch.zeros(self.gflat_numel, dtype=torch.float32, device=self.device)
        else:
            self.gflat.zero_()

    def param_grads(self):
        return [self.gflat[r.goff:r.goff + r.numel].view(r.param.shape) for r in self.param_order]

    def to_ndhwc(self, x: torch.Tensor) -> torch.Tensor:
        """NCDHW fp32 -> internal NDHWC (activation dtype, channels padded to 8)"""
        N, C, T, H, W = x.shape
        out = self._act(N, T, H, W, pad8(C))
        self.lib.call('x3d_ncdhw_to_ndhwc', _ptr(x), _ptr(out), N, C, pad8(C), T, H, W, self.dt, self._stream())
        return out

    def to_ncdhw(self, x: torch.Tensor, C: int) -> torch.Tensor:
        N, T, H, W, Cp = x.shape
        out = self._f32(N, C, T, H, W)
        self.lib.call('x3d_ndhwc_to_ncdhw', _ptr(x), _ptr(out), N, C, Cp, T, H, W, self.dt, self._stream())
        return out

    def _wgrad(self, name, tensors, *args):
        """launch a weight-gradient kernel on the side stream once everything enqueued so far has run"""
        if os.environ.get('X3D_DBG_SKIP_WGRAD'):          # profiling experiment: main chain alone
            return
        if name == 'x3d_pwconv_wgrad' and self.wg_ws is not None:
            # two-stage deterministic reduction through the engine's scratch buffer (all weight-gradient kernels of a
            # pass are serialised on one stream, so one buffer serves them all)
            name = 'x3d_pwconv_wgrad_ws'
            args = args[:-1] + (self.wg_ws.data_ptr(), self.wg_ws.numel()) + args[-1:]
        main = torch.cuda.current_stream(self.device)
        if self.side is None:
            self.lib.call(name, *args, main.cuda_stream)
            return
        ev = torch.cuda.Event()
        ev.record(main)
        self.side.wait_event(ev)
        for t in tensors:
            if t is not None:
                t.record_stream(self.side)
        with torch.cuda.stream(self.side):
            self.lib.call(name, *args, self.side.cuda_stream)

    def _join_side(self):
        if self.side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.side)

    def pack_weights(self):
        self.lib.call('x3d_pack_params', self.pack_descs.data_ptr(), self.n_pack, self.pack_max, self._stream())

    # ------------------------------------------------------------------ small helpers
    def _act(self, *shape):
        return torch.empty(shape, dtype=self.dtype, device=self.device)

    def _f32(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def _bn_forward_state(self, bn_mod, prefix, stats, N, P, C, Cp, training) -> BNState:
        """SubBatchNorm3d statistics -> scale/shift (x3d.py:47-58)."""
        lib, st = self.lib, self._stream()
        if training:
            s = int(bn_mod.num_splits)
            if N % s != 0:
                raise RuntimeError(f'batch {N} not divisible by num_splits {s} (x3d.py:50)')
            buf = self._f32(4, s, Cp)
            sb = bn_mod.split_bn
            track = sb.track_running_stats and sb.running_mean is not None
            lib.call('x3d_bn_finalize', _ptr(stats), N, s, P, C, Cp, self.p(prefix + '.weight'),
                     self.p(prefix + '.bias'), _ptr(sb.running_mean) if track else None,
                     _ptr(sb.running_var) if track else None,
                     _ptr(sb.num_batches_tracked) if track else None,
                     float(sb.momentum if sb.momentum is not None else BN_MOMENTUM), float(sb.eps),
                     _ptr(buf[0]), _ptr(buf[1]), _ptr(buf[2]), _ptr(buf[3]), st)
            return BNState(buf[0], buf[1], buf[2], buf[3], s, True)
        buf = self._f32(4, 1, Cp)
        bn = bn_mod.bn
        lib.call('x3d_bn_eval_params', self.p(prefix + '.weight'), self.p(prefix + '.bias'),
                 _ptr(bn.running_mean), _ptr(bn.running_var), C, Cp, float(bn.eps),
                 _ptr(buf[0]), _ptr(buf[1]), _ptr(buf[2]), _ptr(buf[3]), st)
        return BNState(buf[0], buf[1], buf[2], buf[3], 1, False)

    def _stats(self, arena: Arena, N, Cp):
        return arena.zeros(N * Cp * 2 * 8)

    def _bn_backward(self, arena, prefix, bn: BNState, dout, mask_out, a, N, P, C, Cp, da):
        """Two-pass BN backward: returns da (may alias dout)."""
        lib, st = self.lib, self._stream()
        bst = self._stats(arena, N, Cp)
        lib.call('x3d_bn_bwd_reduce', _ptr(dout), _ptr(mask_out), _ptr(a), _ptr(bst), N, P, Cp, self.dt, st)
        coef = self._f32(3, bn.splits, Cp)
        lib.call('x3d_bn_bwd_finalize', _ptr(bst), N, bn.splits, P, C, Cp, self.p(prefix + '.weight'),
                 _ptr(bn.mean), _ptr(bn.rstd), int(bn.train), _ptr(coef), self.g(prefix + '.weight'),
                 self.g(prefix + '.bias'), st)
        lib.call('x3d_bn_bwd_apply', _ptr(dout), _ptr(mask_out), _ptr(a), _ptr(coef), bn.splits, _ptr(da),
                 N, P, Cp, self.dt, st)
        return da

    # ------------------------------------------------------------------ stem
    def _stem_fwd(self, x, training, save):
        m, lib, st = self.model, self.lib, self._stream()
        N, Ci, T, H, W = x.shape
        C0 = m.conv1_s.out_channels
        C0p = pad8(C0)
        H1, W1 = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        a_s = self._act(N, T, H1, W1, C0p)
        if isinstance(x, UInt8Clips):       # decoded frames: crop / flip / ToTensor / Normalize fused into the conv
            if m.conv1_s.in_channels != 3:
                raise RuntimeError('uint8 frame input needs a 3-channel stem')
            lib.call('x3d_stem_conv_s_fwd_u8', _ptr(x.frames), _ptr(x.crops), self.p('conv1_s.weight'), _ptr(a_s), N, T,
                     x.frames.shape[2], x.frames.shape[3], x.size, x.mean_std_ptr(), float(x.norm_value), C0, C0p,
                     self.dt, st)
        else:
            lib.call('x3d_stem_conv_s_fwd', _ptr(x), self.p('conv1_s.weight'), _ptr(a_s), N, Ci, T, H, W, C0, C0p,
                     self.dt, st)
        a_t = self._act(N, T, H1, W1, C0p)
        P = T * H1 * W1
        st0 = self._stats(self.arena_f, N, C0p) if training else None
        lib.call('x3d_dwconv_fwd', _ptr(a_s), self.pk('conv1_t.d'), _ptr(a_t), N, T, H1, W1, C0p, 5, 1, 1, 1,
                 None, None, 1, 0, _ptr(st0), self.dt, st)
        bn0 = self._bn_forward_state(m.bn1, 'bn1', st0, N, P, C0, C0p, training)
        x0 = self._act(N, T, H1, W1, C0p)
        lib.call('x3d_bn_act_fwd', _ptr(a_t), _ptr(bn0.scale), _ptr(bn0.shift), bn0.splits, None, None, None, 1,
                 _ptr(x0), N, P, C0p, self.dt, st)
        if save is not None:
            save['stem'] = (x, a_s, a_t, bn0, x0)
        return x0, (N, T, H1, W1)

    def _stem_bwd(self, save, dx0):
        m, lib, st = self.model, self.lib, self._stream()
        x, a_s, a_t, bn0, x0 = save['stem']
        N, Ci, T, H, W = x.shape
        _, _, H1, W1, C0p = a_s.shape
        C0 = m.conv1_s.out_channels
        P = T * H1 * W1
        da_t = self._bn_backward(self.arena_b, 'bn1', bn0, dx0, x0, a_t, N, P, C0, C0p, dx0)
        self._wgrad('x3d_dwconv_wgrad', (a_s, da_t), _ptr(a_s), _ptr(da_t), self.g('conv1_t.weight'), N, T, H1, W1, C0,
                    C0p, 5, 1, 1, 1, None, None, 1, 0, self.dt)
        da_s = self._act(N, T, H1, W1, C0p)
        lib.call('x3d_dwconv_dgrad', _ptr(da_t), self.pk('conv1_t.d'), _ptr(da_s), N, T, H1, W1, C0p, 5, 1, 1, 1,
                 None, None, None, 1, None, self.dt, st)
        if isinstance(x, UInt8Clips):
            self._wgrad('x3d_stem_conv_s_wgrad_u8', (x.frames, x.crops, da_s), _ptr(x.frames), _ptr(x.crops), _ptr(da_s),
                        self.g('conv1_s.weight'), N, T, x.frames.shape[2], x.frames.shape[3], x.size, x.mean_std_ptr(),
                        float(x.norm_value), C0, C0p, self.dt)
        else:
            self._wgrad('x3d_stem_conv_s_wgrad', (x, da_s), _ptr(x), _ptr(da_s), self.g('conv1_s.weight'), N, Ci, T, H, W,
                        C0, C0p, self.dt)

    # ------------------------------------------------------------------ bottleneck
    def block_fwd(self, blk, x, geom, training, save_list):
        lib, st, dt = self.lib, self._stream(), self.dt
        N, T, H, W = geom
        pre = blk.prefix
        s = blk.stride
        Cin, Cm, Co = blk.in_planes, blk.mid_planes, blk.out_planes
        Cinp, Cmp, Cop = pad8(Cin), pad8(Cm), pad8(Co)
        Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
        P_in, P_out = T * H * W, T * Ho * Wo
        af = self.arena_f
        # downsample branch (x3d.py:165-166): independent of the main chain until the residual add -> side stream
        ad = bnd = None
        if blk.downsample is not None:
            std = self._stats(af, N, Cop) if training else None
            main = torch.cuda.current_stream(self.device)
            if self.side is not None:
                ev = torch.cuda.Event()
                ev.record(main)
                self.side.wait_event(ev)
                x.record_stream(self.side)
            with torch.cuda.stream(self.side if self.side is not None else main):
                ad = self._act(N, T, Ho, Wo, Cop)
                lib.call('x3d_pwconv_fwd', _ptr(x), self.pk(pre + '.downsample.0.f'), _ptr(ad), N, T, H, W, Cinp, Cop, s,
                         _ptr(std), dt, self._stream())
                bnd = self._bn_forward_state(blk.downsample[1], pre + '.downsample.1', std, N, P_out, Co, Cop, training)
            if self.side is not None:
                ad.record_stream(main)
                bnd.scale.record_stream(main)      # scale/shift/mean/rstd share one allocation
        # conv1 + bn1 statistics
        a1 = self._act(N, T, H, W, Cmp)
        st1 = self._stats(af, N, Cmp) if training else None
        lib.call('x3d_pwconv_fwd', _ptr(x), self.pk(pre + '.conv1.f'), _ptr(a1), N, T, H, W, Cinp, Cmp, 1,
                 _ptr(st1), dt, st)
        bn1 = self._bn_forward_state(blk.bn1, pre + '.bn1', st1, N, P_in, Cm, Cmp, training)
        # depthwise conv2 on relu(bn1(a1)); statistics for bn2 and the SE pool
        a2 = self._act(N, T, Ho, Wo, Cmp)
        need2 = training or blk.has_se
        st2 = self._stats(af, N, Cmp) if need2 else None
        lib.call('x3d_dwconv_fwd', _ptr(a1), self.pk(pre + '.conv2.d'), _ptr(a2), N, T, H, W, Cmp, 3, 3, 3, s,
                 _ptr(bn1.scale), _ptr(bn1.shift), bn1.splits, 1, _ptr(st2), dt, st)
        bn2 = self._bn_forward_state(blk.bn2, pre + '.bn2', st2, N, P_out, Cm, Cmp, training)
        pooled = hidden = gate = None
        if blk.has_se:
            sw = blk.se_width
            pooled, hidden, gate = self._f32(N, Cm), self._f32(N, sw), self._f32(N, Cmp)
            lib.call('x3d_se_fwd', _ptr(st2), _ptr(bn2.scale), _ptr(bn2.shift), bn2.splits, N, P_out, Cm, Cmp, sw,
                     self.p(pre + '.fc1.weight'), self.p(pre + '.fc1.bias'), self.p(pre + '.fc2.weight'),
                     self.p(pre + '.fc2.bias'), _ptr(pooled), _ptr(hidden), _ptr(gate), st)
        v = self._act(N, T, Ho, Wo, Cmp)
        lib.call('x3d_swish_gate_fwd', _ptr(a2), _ptr(bn2.scale), _ptr(bn2.shift), bn2.splits, _ptr(gate), _ptr(v),
                 N, P_out, Cmp, dt, st)
        # conv3 + bn3
        a3 = self._act(N, T, Ho, Wo, Cop)
        st3 = self._stats(af, N, Cop) if training else None
        lib.call('x3d_pwconv_fwd', _ptr(v), self.pk(pre + '.conv3.f'), _ptr(a3), N, T, Ho, Wo, Cmp, Cop, 1,
                 _ptr(st3), dt, st)
        bn3 = self._bn_forward_state(blk.bn3, pre + '.bn3', st3, N, P_out, Co, Cop, training)
        out = self._act(N, T, Ho, Wo, Cop)
        if blk.downsample is not None:
            self._join_side()                 # the downsample branch (side stream) meets the main chain here
            lib.call('x3d_bn_act_fwd', _ptr(a3), _ptr(bn3.scale), _ptr(bn3.shift), bn3.splits, _ptr(ad),
                     _ptr(bnd.scale), _ptr(bnd.shift), 1, _ptr(out), N, P_out, Cop, dt, st)
        else:
            lib.call('x3d_bn_act_fwd', _ptr(a3), _ptr(bn3.scale), _ptr(bn3.shift), bn3.splits, _ptr(x), None, None,
                     1, _ptr(out), N, P_out, Cop, dt, st)
        if save_list is not None:
            save_list.append((blk, geom, x, a1, bn1, a2, st2, bn2, pooled, hidden, gate, v, a3, bn3, ad, bnd, out))
        return out, (N, T, Ho, Wo)

    def block_bwd(self, rec, dout, need_dx=True):
        lib, st, dt = self.lib, self._stream(), self.dt
        (blk, geom, x, a1, bn1, a2, st2, bn2, pooled, hidden, gate, v, a3, bn3, ad, bnd, out) = rec
        N, T, H, W = geom
        pre = blk.prefix
        s = blk.stride
        Cin, Cm, Co = blk.in_planes, blk.mid_planes, blk.out_planes
        Cinp, Cmp, Cop = pad8(Cin), pad8(Cm), pad8(Co)
        Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
        P_in, P_out = T * H * W, T * Ho * Wo
        ab = self.arena_b
        # ---- downsample branch: its BN backward only needs dout / out / ad -> side stream, met again at the dgrad
        dad = ev_dad = None
        if blk.downsample is not None:
            main = torch.cuda.current_stream(self.device)
            if self.side is not None:
                ev = torch.cuda.Event()
                ev.record(main)
                self.side.wait_event(ev)
                for t in (dout, out, ad):
                    t.record_stream(self.side)
            with torch.cuda.stream(self.side if self.side is not None else main):
                dad = self._act(N, T, Ho, Wo, Cop)
                self._bn_backward(ab, pre + '.downsample.1', bnd, dout, out, ad, N, P_out, Co, Cop, dad)
            if self.side is not None:
                dad.record_stream(main)
                ev_dad = torch.cuda.Event()
                ev_dad.record(self.side)
        # ---- bn3 backward (dpre = dout * [out > 0])
        da3 = self._act(N, T, Ho, Wo, Cop)
        dx = None
        if blk.downsample is None and need_dx:
            # identity residual (x3d.py:168-169): dx = dpre + conv1-dgrad.  The reduce pass writes dpre straight into
            # the dx buffer (same shape as dout), the apply pass reads it unmasked, and the conv1 dgrad at the end of
            # this function accumulates on top -- no separate "dx += dout*(out>0)" pass, one tensor read less per pass
            dx = self._act(N, T, H, W, Cinp)
            bst3 = self._stats(ab, N, Cop)
            lib.call('x3d_bn_bwd_reduce_store', _ptr(dout), _ptr(out), _ptr(a3), _ptr(bst3), _ptr(dx), N, P_out, Cop,
                     dt, st)
            coef3 = self._f32(3, bn3.splits, Cop)
            lib.call('x3d_bn_bwd_finalize', _ptr(bst3), N, bn3.splits, P_out, Co, Cop, self.p(pre + '.bn3.weight'),
                     _ptr(bn3.mean), _ptr(bn3.rstd), int(bn3.train), _ptr(coef3), self.g(pre + '.bn3.weight'),
                     self.g(pre + '.bn3.bias'), st)
            lib.call('x3d_bn_bwd_apply', _ptr(dx), None, _ptr(a3), _ptr(coef3), bn3.splits, _ptr(da3), N, P_out, Cop,
                     dt, st)
        else:
            self._bn_backward(ab, pre + '.bn3', bn3, dout, out, a3, N, P_out, Co, Cop, da3)
        # ---- conv3
        self._wgrad('x3d_pwconv_wgrad', (v, da3), _ptr(v), _ptr(da3), self.g(pre + '.conv3.weight'), N, T, Ho, Wo, Cm,
                    Cmp, Co, Cop, 1, dt)
        dv = self._act(N, T, Ho, Wo, Cmp)
        lib.call('x3d_pwconv_dgrad', _ptr(da3), self.pk(pre + '.conv3.t'), _ptr(dv), N, T, Ho, Wo, Cmp, Cop, 1, 0,
                 dt, st)
        del da3
        # ---- swish, SE gate, bn2
        bst2 = self._stats(ab, N, Cmp)
        lib.call('x3d_swish_gate_bwd_reduce', _ptr(dv), _ptr(a2), _ptr(bn2.scale), _ptr(bn2.shift), bn2.splits,
                 _ptr(gate), _ptr(bst2), N, P_out, Cmp, dt, st)
        coef2, work = self._f32(3, N, Cmp), self._f32(N, 2 * Cmp + max(blk.se_width, 1))
        if blk.has_se:
            se_args = (self.p(pre + '.fc1.weight'), self.p(pre + '.fc2.weight'), _ptr(pooled), _ptr(hidden), _ptr(gate),
                       self.g(pre + '.fc1.weight'), self.g(pre + '.fc1.bias'), self.g(pre + '.fc2.weight'),
                       self.g(pre + '.fc2.bias'))
            sw = blk.se_width
        else:
            se_args = (None,) * 9
            sw = 0
        fwd_stats = st2 if st2 is not None else bst2   # only dereferenced when the block has SE
        lib.call('x3d_se_bn_bwd', _ptr(fwd_stats), _ptr(bst2), N, bn2.splits, P_out, Cm, Cmp, sw,
                 self.p(pre + '.bn2.weight'), _ptr(bn2.mean), _ptr(bn2.rstd), _ptr(bn2.scale), _ptr(bn2.shift),
                 int(bn2.train), *se_args, self.g(pre + '.bn2.weight'), self.g(pre + '.bn2.bias'), _ptr(work),
                 _ptr(coef2), st)
        da2 = dv   # in place
        lib.call('x3d_swish_gate_bwd_apply', _ptr(dv), _ptr(a2), _ptr(bn2.scale), _ptr(bn2.shift), bn2.splits,
                 _ptr(gate), _ptr(coef2), _ptr(da2), N, P_out, Cmp, dt, st)
        # ---- depthwise conv2
        self._wgrad('x3d_dwconv_wgrad', (a1, da2, bn1.scale, bn1.shift), _ptr(a1), _ptr(da2), self.g(pre + '.conv2.weight'), N, T,
                    H, W, Cm, Cmp, 3, 3, 3, s, _ptr(bn1.scale), _ptr(bn1.shift), bn1.splits, 1, dt)
        d1 = self._act(N, T, H, W, Cmp)
        bst1 = self._stats(ab, N, Cmp)
        lib.call('x3d_dwconv_dgrad', _ptr(da2), self.pk(pre + '.conv2.d'), _ptr(d1), N, T, H, W, Cmp, 3, 3, 3, s,
                 _ptr(a1), _ptr(bn1.scale), _ptr(bn1.shift), bn1.splits, _ptr(bst1), dt, st)
        del dv, da2
        # ---- bn1 (relu mask already applied by the dgrad epilogue)
        coef1 = self._f32(3, bn1.splits, Cmp)
        lib.call('x3d_bn_bwd_finalize', _ptr(bst1), N, bn1.splits, P_in, Cm, Cmp, self.p(pre + '.bn1.weight'),
                 _ptr(bn1.mean), _ptr(bn1.rstd), int(bn1.train), _ptr(coef1), self.g(pre + '.bn1.weight'),
                 self.g(pre + '.bn1.bias'), st)
        da1 = d1
        lib.call('x3d_bn_bwd_apply', _ptr(d1), None, _ptr(a1), _ptr(coef1), bn1.splits, _ptr(da1), N, P_in, Cmp, dt, st)
        # ---- conv1
        self._wgrad('x3d_pwconv_wgrad', (x, da1), _ptr(x), _ptr(da1), self.g(pre + '.conv1.weight'), N, T, H, W, Cin,
                    Cinp, Cm, Cmp, 1, dt)
        if need_dx:
            acc1 = 1 if dx is not None else 0        # identity block: dx already holds dpre
            if dx is None:
                dx = self._act(N, T, H, W, Cinp)
            lib.call('x3d_pwconv_dgrad', _ptr(da1), self.pk(pre + '.conv1.t'), _ptr(dx), N, T, H, W, Cinp, Cmp, 1, acc1,
                     dt, st)
        # ---- residual branch
        if blk.downsample is not None:
            if ev_dad is not None:
                torch.cuda.current_stream(self.device).wait_event(ev_dad)
            self._wgrad('x3d_pwconv_wgrad', (x, dad), _ptr(x), _ptr(dad), self.g(pre + '.downsample.0.weight'), N, T, H,
                        W, Cin, Cinp, Co, Cop, s, dt)
            if need_dx:
                lib.call('x3d_pwconv_dgrad', _ptr(dad), self.pk(pre + '.downsample.0.t'), _ptr(dx), N, T, H, W, Cinp,
                         Cop, s, 1, dt, st)
        return dx

    # ------------------------------------------------------------------ head
    def _head_fwd(self, xL, geom, training, dropout_mask, save):
        m, lib, st, dt = self.model, self.lib, self._stream(), self.dt
        N, T, H, W = geom
        C5in, C5 = m.conv5.in_channels, m.conv5.out_channels
        C5inp, C5p = pad8(C5in), pad8(C5)
        P = T * H * W
        a5 = self._act(N, T, H, W, C5p)
        st5 = self._stats(self.arena_f, N, C5p) if training else None
        lib.call('x3d_pwconv_fwd', _ptr(xL), self.pk('conv5.f'), _ptr(a5), N, T, H, W, C5inp, C5p, 1, _ptr(st5), dt, st)
        bn5 = self._bn_forward_state(m.bn5, 'bn5', st5, N, P, C5, C5p, training)
        pool_t = 1 if m.task == 'class' else 0
        R = N if pool_t else N * T
        pooled = self._f32(R, C5)
        lib.call('x3d_bn_relu_pool_fwd', _ptr(a5), _ptr(bn5.scale), _ptr(bn5.shift), bn5.splits, _ptr(pooled), N, T,
                 H * W, pool_t, C5, C5p, dt, st)
        F1 = m.fc1.out_channels
        h1 = self._f32(R, F1)
        lib.call('x3d_small_gemm_ws', _ptr(pooled), C5, 1, self.p('fc1.weight'), 1, C5, _ptr(h1), F1, R, F1, C5, None, 1,
                 _ptr(dropout_mask), 0, self.head_ws.data_ptr(), self.head_ws.numel(), st)
        ncls = m.fc2.out_features
        logits = self._f32(R, ncls)
        lib.call('x3d_small_gemm_ws', _ptr(h1), F1, 1, self.p('fc2.weight'), 1, F1, _ptr(logits), ncls, R, ncls, F1,
                 self.p('fc2.bias'), 0, None, 0, self.head_ws.data_ptr(), self.head_ws.numel(), st)
        if save is not None:
            save['head'] = (xL, geom, a5, bn5, pooled, h1, dropout_mask)
        if pool_t:
            return logits.view(N, ncls, 1)
        return logits.view(N, T, ncls).permute(0, 2, 1)

    def _head_bwd(self, save, dlogits):
        m, lib, st, dt = self.model, self.lib, self._stream(), self.dt
        xL, geom, a5, bn5, pooled, h1, dropout_mask = save['head']
        N, T, H, W = geom
        C5in, C5 = m.conv5.in_channels, m.conv5.out_channels
        C5inp, C5p = pad8(C5in), pad8(C5)
        P = T * H * W
        pool_t = 1 if m.task == 'class' else 0
        R = N if pool_t else N * T
        F1 = m.fc1.out_channels
        ncls = m.fc2.out_features
        if pool_t:
            dl = dlogits.reshape(N, ncls)
        else:
            dl = dlogits.permute(0, 2, 1).reshape(R, ncls)
        dl = dl.to(torch.float32).contiguous()
        # fc2
        # parameter gradients feed nothing downstream: side stream (K = batch rows: the plain tile kernel, no workspace)
        self._wgrad('x3d_small_gemm', (dl, h1), _ptr(dl), 1, ncls, _ptr(h1), F1, 1, self.g('fc2.weight'), F1, ncls, F1, R,
                    None, 0, None, 1)
        self._wgrad('x3d_colsum', (dl,), _ptr(dl), R, ncls, self.g('fc2.bias'))
        dh = self._f32(R, F1)
        lib.call('x3d_small_gemm_ws', _ptr(dl), ncls, 1, self.p('fc2.weight'), F1, 1, _ptr(dh), F1, R, F1, ncls, None, 0,
                 None, 0, self.head_ws.data_ptr(), self.head_ws.numel(), st)
        dz = dh
        lib.call('x3d_relu_mask_mul', _ptr(dh), _ptr(h1), _ptr(dropout_mask), _ptr(dz), R * F1, st)
        # fc1
        self._wgrad('x3d_small_gemm', (dz, pooled), _ptr(dz), 1, F1, _ptr(pooled), C5, 1, self.g('fc1.weight'), C5, F1, C5,
                    R, None, 0, None, 1)
        dpooled = self._f32(R, C5)
        lib.call('x3d_small_gemm_ws', _ptr(dz), F1, 1, self.p('fc1.weight'), C5, 1, _ptr(dpooled), C5, R, C5, F1, None, 0,
                 None, 0, self.head_ws.data_ptr(), self.head_ws.numel(), st)
        # relu + pool + bn5
        bst5 = self._stats(self.arena_b, N, C5p)
        lib.call('x3d_bn_relu_pool_bwd_reduce', _ptr(a5), _ptr(bn5.scale), _ptr(bn5.shift), bn5.splits, _ptr(dpooled),
                 _ptr(bst5), N, T, H * W, pool_t, C5, C5p, dt, st)
        coef5 = self._f32(3, bn5.splits, C5p)
        lib.call('x3d_bn_bwd_finalize', _ptr(bst5), N, bn5.splits, P, C5, C5p, self.p('bn5.weight'), _ptr(bn5.mean),
                 _ptr(bn5.rstd), int(bn5.train), _ptr(coef5), self.g('bn5.weight'), self.g('bn5.bias'), st)
        da5 = self._act(N, T, H, W, C5p)
        lib.call('x3d_bn_relu_pool_bwd_apply', _ptr(a5), _ptr(bn5.scale), _ptr(bn5.shift), bn5.splits, _ptr(dpooled),
                 _ptr(coef5), _ptr(da5), N, T, H * W, pool_t, C5, C5p, dt, st)
        # conv5
        self._wgrad('x3d_pwconv_wgrad', (xL, da5), _ptr(xL), _ptr(da5), self.g('conv5.weight'), N, T, H, W, C5in, C5inp,
                    C5, C5p, 1, dt)
        dxL = self._act(N, T, H, W, C5inp)
        lib.call('x3d_pwconv_dgrad', _ptr(da5), self.pk('conv5.t'), _ptr(dxL), N, T, H, W, C5inp, C5p, 1, 0, dt, st)
        return dxL

    # ------------------------------------------------------------------ whole network
    def forward(self, x: torch.Tensor, training: bool, need_grad: bool,
                dropout_mask: Optional[torch.Tensor] = None):
        """x: [B,3,T,H,W] fp32 NCDHW on CUDA -> logits ([B,C,1] or [B,C,T], fp32), saved state."""
        if isinstance(x, UInt8Clips):
            if not x.fused:                   # expand on the device (bit-identical to the reference's transform chain)
                from .input_pipeline import clip_from_uint8
                x = clip_from_uint8(x)
        else:
            if not x.is_cuda:
                raise RuntimeError('x3d_multigrid_b200 runs on CUDA only (no CPU fallback)')
            if x.dtype != torch.float32:
                raise RuntimeError('input clips must be fp32 NCDHW (x3d.py:316) or input_pipeline.UInt8Clips')
            x = x.contiguous()
        # kernels launch on the CURRENT device: make it the clip's device (a process may drive several GPUs, e.g.
        # the nn.DataParallel replicas the reference supports, x3d.py:278-281 / train_..._multigrid.py:175-177)
        with torch.cuda.device(x.device):
            self.prepare(x.device)
            self.arena_f.begin()
            self.pack_weights()
            save = {'blocks': []} if need_grad else None
            h, geom = self._stem_fwd(x, training, save)
            for blk in self.model.blocks():
                h, geom = self.block_fwd(blk, h, geom, training, save['blocks'] if save is not None else None)
            logits = self._head_fwd(h, geom, training, dropout_mask, save)
            if save is not None:
                save['arena'] = self.arena_f.frame()      # st2 of the SE blocks is read again by the backward pass
        return logits, save

    def backward(self, save, dlogits: torch.Tensor):
        """Accumulates parameter gradients into ``self.gflat`` (zeroed here first)."""
        with torch.cuda.device(self.device):
            return self._backward(save, dlogits)

    def _backward(self, save, dlogits: torch.Tensor):
        self.arena_b.begin()
        self.new_grad_buffer()
        d = self._head_bwd(save, dlogits)
        blocks = save['blocks']
        stage_of = [rec[0].stage for rec in blocks]
        for i in range(len(blocks) - 1, -1, -1):
            d = self.block_bwd(blocks[i], d)
            blocks[i] = None
            # bucket b (0 = head+stage4, 1 = stage3, 2 = stage2) is final once its first block is done
            if self.grad_hook and (i == 0 or stage_of[i - 1] != stage_of[i]) and stage_of[i] > 1:
                self._join_side()
                self.grad_hook(4 - stage_of[i])
        self._stem_bwd(save, d)
        self._join_side()
        if self.grad_hook:
            self.grad_hook(3)
        return self.param_grads()
