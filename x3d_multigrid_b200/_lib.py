"""ctypes binding of libx3d_b200.so (the C ABI declared in include/x3d_b200.h).

The prototypes are parsed from the header itself, so the binding cannot drift from the
ABI.  There is no fallback: if the shared library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(ROOT, 'include', 'x3d_b200.h')
LIB_PATH = os.path.join(_HERE, 'libx3d_b200.so')

F32, BF16 = 0, 1


class PackDesc(ctypes.Structure):
    """x3d_pack_desc_t"""
    _fields_ = [('src', ctypes.c_void_p), ('dst', ctypes.c_void_p), ('rows', ctypes.c_int32),
                ('cols', ctypes.c_int32), ('dst_rows', ctypes.c_int32), ('dst_cols', ctypes.c_int32),
                ('transpose', ctypes.c_int32), ('dtype', ctypes.c_int32)]


class SgdDesc(ctypes.Structure):
    """x3d_sgd_desc_t"""
    _fields_ = [('param', ctypes.c_void_p), ('grad', ctypes.c_void_p), ('momentum_buf', ctypes.c_void_p),
                ('numel', ctypes.c_int64)]


def _ctype(decl: str):
    d = decl.strip()
    if '*' in d:
        return ctypes.c_char_p if d.replace(' ', '') == 'constchar*' else ctypes.c_void_p
    base = d.replace('const', '').split()[0]
    return {'int64_t': ctypes.c_int64, 'int': ctypes.c_int, 'int32_t': ctypes.c_int32,
            'float': ctypes.c_float, 'double': ctypes.c_double, 'size_t': ctypes.c_size_t,
            'x3d_dtype_t': ctypes.c_int, 'x3d_stream_t': ctypes.c_void_p, 'void': None}[base]


def parse_header(path: str = HEADER) -> List[Tuple[str, object, List[object]]]:
    """[(symbol, restype, [argtypes])] for every function the header declares."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    src = re.sub(r'//[^\n]*', ' ', src)
    src = re.sub(r'typedef\s+struct\s*\{.*?\}\s*\w+\s*;', ' ', src, flags=re.S)
    protos = []
    for m in re.finditer(r'([A-Za-z_][\w\s\*]*?)\b(x3d_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith('typedef'):
            continue
        argtypes = []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                # drop the parameter name (last identifier) unless the decl ends with '*'
                decl = a if a.endswith('*') else re.sub(r'\b\w+$', '', a).strip()
                argtypes.append(_ctype(decl))
        protos.append((name, _ctype(ret), argtypes))
    return protos


class Lib:
    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise RuntimeError(
                f'{path} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                f'(or make -C x3d_multigrid_b200/csrc).  There is no CPU / PyTorch fallback.')
        self.cdll = ctypes.CDLL(path)
        self.fn: Dict[str, object] = {}
        for name, restype, argtypes in parse_header():
            f = getattr(self.cdll, name)       # AttributeError if the .so lacks a declared symbol
            f.restype = restype
            f.argtypes = argtypes
            self.fn[name] = f
        self._err = self.fn['x3d_last_error']
        # optional per-kernel timing (bench.py): CUDA events around the calls named in prof_names,
        # recorded on torch's current stream, which is the stream every call is launched on
        self.prof_names = set()
        self.prof_records = []

    def call(self, name: str, *args):
        if name in self.prof_names:
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = self.fn[name](*args)
            e1.record()
            self.prof_records.append((name, args, e0, e1))
        else:
            rc = self.fn[name](*args)
        if rc != 0:
            raise RuntimeError(f'{name} failed (rc={rc}): {self._err().decode()}')

    def launch_count(self) -> int:
        return int(self.fn['x3d_launch_count']())

    def path_counts(self) -> Dict[str, int]:
        """{kernel family: calls served} -- which implementation each conv call took (x3d_path_t); tests assert
        with this that the hot shapes run the TMA-tiled / tcgen05 kernels, not the shape-generic fallbacks"""
        n = 0
        out = {}
        while True:
            c = int(self.fn['x3d_path_count'](n))
            if c < 0:
                return out
            out[self.fn['x3d_path_name'](n).decode()] = c
            n += 1


_LIB = None


def lib() -> Lib:
    global _LIB
    if _LIB is None:
        _LIB = Lib()
    return _LIB
