"""Test-only host build + ctypes driver of the kernel emulation (tests/emu/emu.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from sod100k_b200 import ir

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcsnet_emu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "emu.cpp")
        deps = [src, os.path.join(HERE, "..", "..", "sod100k_b200", "csrc", "generic_ops.cuh"),
                os.path.join(HERE, "..", "..", "include", "csnet_b200.h")]
        if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
            subprocess.run(["g++", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c++17", "-o", LIB, src], check=True)
        _lib = C.CDLL(LIB)
        _lib.csnet_emu_run.restype = C.c_int
    return _lib


def run_ext(prog: ir.Program, ext_arrays, N: int, taps=()):
    """Generic variant: `ext_arrays[i]` is bound to external i (float32 numpy arrays, outputs are written in place)."""
    arena = np.zeros(prog.arena_bytes_per_image * N + 256, np.uint8)
    ext = (C.c_void_p * len(ext_arrays))(*[a.ctypes.data for a in ext_arrays])
    blob = np.ascontiguousarray(prog.blob, np.float32)
    rc = lib().csnet_emu_run(prog.tensor_array(), len(prog.tensors), prog.op_array(), len(prog.ops),
                             blob.ctypes.data_as(C.POINTER(C.c_float)), N, ext, arena.ctypes.data_as(C.c_char_p))
    if rc != 0:
        raise RuntimeError(f"emu failed rc={rc}")
    got = {}
    for name in taps:
        t = prog.tensors[prog.taps[name]]
        off = N * t.arena_offset
        got[name] = arena[off:off + N * t.bytes_per_image].view(np.float32).reshape(N, t.C, t.H, t.W).copy()
    return got


def run(prog: ir.Program, x: np.ndarray, taps=()):
    """Execute an fp32 program on the host.  Returns (logits, {tap name: array}); compile the program with
    reuse_arena=False when taps are wanted."""
    N = x.shape[0]
    x = np.ascontiguousarray(x, np.float32)
    out_t = prog.tensors[prog.output]
    y = np.zeros((N, out_t.C, out_t.H, out_t.W), np.float32)
    arena = np.zeros(prog.arena_bytes_per_image * N + 256, np.uint8)
    ext = (C.c_void_p * 2)(x.ctypes.data, y.ctypes.data)
    blob = np.ascontiguousarray(prog.blob, np.float32)
    rc = lib().csnet_emu_run(prog.tensor_array(), len(prog.tensors), prog.op_array(), len(prog.ops),
                             blob.ctypes.data_as(C.POINTER(C.c_float)), N, ext, arena.ctypes.data_as(C.c_char_p))
    if rc != 0:
        raise RuntimeError(f"emu failed rc={rc}")
    got = {}
    for name in taps:
        t = prog.tensors[prog.taps[name]]
        off = N * t.arena_offset
        got[name] = arena[off:off + N * t.bytes_per_image].view(np.float32).reshape(N, t.C, t.H, t.W).copy()
    return y, got
