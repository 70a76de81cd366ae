"""TEST INFRASTRUCTURE: g++ build of the CUDA library's device-function headers (see harness.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'libglamr_hostmath.so')
SRC = [os.path.join(HERE, 'harness.cpp'), os.path.join(HERE, 'emu.cpp')]
DEPS = SRC + [os.path.join(HERE, '..', '..', 'glamr_b200', 'csrc', f) for f in ['glamr_math.cuh', 'rowops.cuh', 'globalopt_frames.cuh', 'eval_math.cuh']] + \
    [os.path.join(HERE, '..', '..', 'include', 'glamr_b200.h')]


def build(force=False):
    newest = max(os.path.getmtime(p) for p in DEPS)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-shared', '-fPIC', '-x', 'c++'] + SRC + ['-o', SO])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _fp(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def rowop_dims(op):
    d0, d1, do = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib().glamr_host_rowop_dims(op, ctypes.byref(d0), ctypes.byref(d1), ctypes.byref(do))
    return d0.value, d1.value, do.value


def rowop_fwd(op, a, b=None):
    d0, d1, do = rowop_dims(op)
    a = np.ascontiguousarray(a, np.float32).reshape(-1, d0)
    b = None if b is None else np.ascontiguousarray(b, np.float32).reshape(-1, d1)
    out = np.zeros((a.shape[0], do), np.float32)
    assert lib().glamr_host_rowop_fwd(op, a.shape[0], _fp(a), _fp(b), _fp(out)) == 0
    return out


def rowop_vjp(op, a, b, g, want_b=False):
    d0, d1, do = rowop_dims(op)
    a = np.ascontiguousarray(a, np.float32).reshape(-1, d0)
    b = None if b is None else np.ascontiguousarray(b, np.float32).reshape(-1, d1)
    g = np.ascontiguousarray(g, np.float32).reshape(-1, do)
    ga = np.zeros_like(a)
    gb = np.zeros_like(b) if (b is not None and want_b) else None
    assert lib().glamr_host_rowop_vjp(op, a.shape[0], _fp(a), _fp(b), _fp(g), _fp(ga), _fp(gb)) == 0
    return ga, gb
