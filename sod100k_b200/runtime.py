"""ctypes binding of libcsnet_b200.so (include/csnet_b200.h).  PyTorch is only plumbing here: it owns the
device buffers and the stream; every kernel that runs is ours.  There is no CPU or library fallback — a
missing library or a missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import ir

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcsnet_b200.so")
ABI_VERSION = 5
PARAM_EPOCH = 0      # bumped by in-place parameter updates that bypass torch's version counters (FusedAdam)
_lib = None

# every symbol include/csnet_b200.h declares (tests check the library exports exactly these)
SYMBOLS = ("csnet_abi_version", "csnet_last_error", "csnet_device_count", "csnet_plan_create",
           "csnet_plan_set_blob", "csnet_plan_run", "csnet_plan_profile", "csnet_plan_tensor_ptr", "csnet_plan_read_tensor", "csnet_plan_op_kernel", "csnet_plan_launches",
           "csnet_plan_arena_bytes", "csnet_plan_destroy", "csnet_plan_run_host", "csnet_plan_run_host_u8",
           "csnet_train_last_error", "csnet_train_bn_stats", "csnet_train_bn_prelu_fwd", "csnet_train_bn_prelu_bwd",
           "csnet_train_dw_conv", "csnet_train_dw_wgrad", "csnet_train_dw_bwd", "csnet_train_mix_fwd", "csnet_train_mix_dgrad",
           "csnet_train_mix_wgrad", "csnet_train_pool_fwd", "csnet_train_pool_bwd", "csnet_slim_gather", "csnet_train_bce", "csnet_train_adam", "csnet_salmetric_hist")


class EngineError(RuntimeError):
    pass


def load_library(path: Optional[str] = None):
    """dlopen the engine.  Raises EngineError (never falls back) if it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise EngineError(f"{path} not found — build it with `python -m sod100k_b200.build` "
                          "(__graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    lib.csnet_abi_version.restype = C.c_int
    lib.csnet_last_error.restype = C.c_char_p
    lib.csnet_device_count.restype = C.c_int
    lib.csnet_plan_create.restype = C.c_int
    lib.csnet_plan_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(ir.TensorDesc), C.c_int32,
                                      C.POINTER(ir.OpDesc), C.c_int32, C.c_int64, C.c_int32, C.c_int32]
    lib.csnet_plan_set_blob.restype = C.c_int
    lib.csnet_plan_set_blob.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.csnet_plan_run.restype = C.c_int
    lib.csnet_plan_run.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]
    lib.csnet_plan_profile.restype = C.c_int
    lib.csnet_plan_profile.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                                       C.POINTER(C.c_float), C.c_int32]
    lib.csnet_plan_tensor_ptr.restype = C.c_void_p
    lib.csnet_plan_tensor_ptr.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.csnet_plan_read_tensor.restype = C.c_int
    lib.csnet_plan_read_tensor.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.csnet_plan_op_kernel.restype = C.c_char_p
    lib.csnet_plan_op_kernel.argtypes = [C.c_void_p, C.c_int32]
    lib.csnet_plan_launches.restype = C.c_int32
    lib.csnet_plan_launches.argtypes = [C.c_void_p]
    lib.csnet_plan_arena_bytes.restype = C.c_int64
    lib.csnet_plan_arena_bytes.argtypes = [C.c_void_p]
    lib.csnet_plan_destroy.restype = None
    lib.csnet_plan_destroy.argtypes = [C.c_void_p]
    lib.csnet_plan_run_host.restype = C.c_int
    lib.csnet_plan_run_host.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.csnet_plan_run_host_u8.restype = C.c_int
    lib.csnet_plan_run_host_u8.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
    lib.csnet_slim_gather.restype = C.c_int
    lib.csnet_slim_gather.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.csnet_train_last_error.restype = C.c_char_p
    if lib.csnet_abi_version() != ABI_VERSION:
        raise EngineError(f"{path}: ABI {lib.csnet_abi_version()} != expected {ABI_VERSION}; rebuild")
    if path == LIB_PATH:
        _lib = lib
    return lib


def _check(lib, rc: int, what: str):
    if rc != 0:
        raise EngineError(f"{what} failed ({rc}): {lib.csnet_last_error().decode()}")


class Plan:
    """One compiled program resident on one GPU."""

    def __init__(self, prog: ir.Program, max_batch: int, device: int = 0):
        self.lib = load_library()
        n = self.lib.csnet_device_count()
        if n <= 0:
            raise EngineError("no CUDA device visible: the CSNet engine runs on the GPU only "
                              f"({self.lib.csnet_last_error().decode()})")
        self.prog = prog
        self.max_batch = int(max_batch)
        self.device = int(device)
        self._h = C.c_void_p()
        self._tensors, self._ops = prog.tensor_array(), prog.op_array()
        _check(self.lib, self.lib.csnet_plan_create(C.byref(self._h), self._tensors, len(prog.tensors), self._ops,
                                                    len(prog.ops), int(prog.blob.size), self.max_batch, self.device),
               "csnet_plan_create")
        self.set_blob(prog.blob)

    def set_blob(self, blob: np.ndarray, stream: int = 0):
        blob = np.ascontiguousarray(blob, np.float32)
        _check(self.lib, self.lib.csnet_plan_set_blob(self._h, blob.ctypes.data, blob.size, stream), "csnet_plan_set_blob")

    def run(self, N: int, ext_ptrs, stream: int = 0):
        arr = (C.c_void_p * len(ext_ptrs))(*[int(p) for p in ext_ptrs])
        _check(self.lib, self.lib.csnet_plan_run(self._h, int(N), arr, len(ext_ptrs), stream), "csnet_plan_run")

    def profile(self, N: int, ext_ptrs, stream: int = 0):
        """Per-op device milliseconds of one run (CUDA events around every launch)."""
        arr = (C.c_void_p * len(ext_ptrs))(*[int(p) for p in ext_ptrs])
        ms = (C.c_float * len(self.prog.ops))()
        _check(self.lib, self.lib.csnet_plan_profile(self._h, int(N), arr, len(ext_ptrs), stream, ms, len(self.prog.ops)),
               "csnet_plan_profile")
        return list(ms)

    def run_host(self, N: int, x_host_ptr: int, y_host_ptr: int, stream: int = 0):
        _check(self.lib, self.lib.csnet_plan_run_host(self._h, int(N), x_host_ptr, y_host_ptr, stream), "csnet_plan_run_host")

    def run_host_u8(self, N: int, x_host_ptr: int, y_host_ptr: int, mean, std, stream: int = 0):
        m, s_ = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
        _check(self.lib, self.lib.csnet_plan_run_host_u8(self._h, int(N), x_host_ptr, y_host_ptr, m, s_, stream), "csnet_plan_run_host_u8")

    def tensor_ptr(self, tensor: int, N: int) -> int:
        return int(self.lib.csnet_plan_tensor_ptr(self._h, tensor, N) or 0)

    def op_kernel(self, i: int) -> str:
        """Kernel (family) that runs op i of this plan."""
        return (self.lib.csnet_plan_op_kernel(self._h, int(i)) or b"").decode()

    @property
    def launches(self) -> int:
        return int(self.lib.csnet_plan_launches(self._h))

    @property
    def arena_bytes(self) -> int:
        return int(self.lib.csnet_plan_arena_bytes(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.csnet_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- torch plumbing ---------------------------------------------------------------------------
    def forward(self, x):
        """x: CUDA float32 [N,3,H,W] contiguous -> float32 logits [N,1,H,W] on the current stream."""
        import torch

        if not x.is_cuda:
            raise EngineError("input must be a CUDA tensor: the engine has no CPU path")
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        t_in, t_out = self.prog.tensors[self.prog.input], self.prog.tensors[self.prog.output]
        if tuple(x.shape[1:]) != (t_in.C, t_in.H, t_in.W):
            raise EngineError(f"plan compiled for {(t_in.C, t_in.H, t_in.W)}, got {tuple(x.shape[1:])}")
        N = x.shape[0]
        y = torch.empty((N, t_out.C, t_out.H, t_out.W), dtype=torch.float32, device=x.device)
        self.run(N, [x.data_ptr(), y.data_ptr()], torch.cuda.current_stream(x.device).cuda_stream)
        return y

    def read_tensor(self, tensor: int, N: int):
        """Copy an arena tensor of the last run (batch N) into a float32 torch tensor (tests / taps)."""
        import torch

        t = self.prog.tensors[tensor]
        tdt = {ir.F32: torch.float32, ir.F16: torch.float16, ir.BF16: torch.bfloat16}[t.dtype]
        out = torch.empty((N, t.C, t.H, t.W), dtype=tdt, device=f"cuda:{self.device}")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _check(self.lib, self.lib.csnet_plan_read_tensor(self._h, tensor, N, out.data_ptr(), stream), "csnet_plan_read_tensor")
        return out.float()
