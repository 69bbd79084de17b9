"""Python face of the NVLink transport's device executor (csrc/cuda/nvl_exec.cu): the
semi-persistent cluster kernels that serve isend for registered CUDA buffers, plus the
fused move+accumulate / move+cast variants (K1, K4, K5, K7, K8)."""
from __future__ import annotations

import ctypes as C
import time

import torch

from ..utils.native import load

OPS = {"copy": 0, "red_add_f32": 1, "red_add_bf16": 2, "cast_bf16_to_f32": 3, "cast_f32_to_bf16": 4, "flush": 5,
       "acc_bf16_to_f32": 6, "cast_bf16_to_e4m3": 7, "acc_e4m3_to_f32": 8, "cast_f32_to_e4m3": 9,
       "cast_bf16_to_e5m2": 10, "acc_e5m2_to_f32": 11, "cast_f32_to_e5m2": 12, "cast_e4m3_to_f32": 13, "cast_e5m2_to_f32": 14}


class P2PExecutor:
    """Submit copy/reduce/cast jobs between device buffers (local or peer-mapped) and wait on
    the per-chunk completion words, exactly as the plugin's isend/test do."""

    MAX_CHUNKS = 16

    def __init__(self, device: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("P2PExecutor needs a CUDA device")
        self.lib = load()
        self.lib.bnet_exec_op.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                          C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
        self.lib.bnet_exec_op_scaled.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                                 C.c_uint64, C.c_float, C.POINTER(C.c_int)]
        self.device = torch.cuda.current_device() if device is None else device
        # pinned + UVA: the same address is valid on host and device
        self.flags = torch.zeros(self.MAX_CHUNKS * 64, dtype=torch.int64).pin_memory()
        self.seq = 0
        self.slot = 0

    def submit(self, op: str, src: torch.Tensor, dst: torch.Tensor, src_bytes: int | None = None, sync: bool = True,
               scale: float | None = None):
        # The executor's kernels run on their own streams: whatever produced `src` / initialised
        # `dst` on torch's stream must be finished first (NCCL gives the plugin the same guarantee).
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        nbytes = src.numel() * src.element_size() if src_bytes is None else src_bytes
        self.seq += 1
        slot = self.slot
        self.slot = (self.slot + 1) % 64
        base = self.flags.data_ptr() + slot * self.MAX_CHUNKS * 8
        n = C.c_int(0)
        if scale is not None:
            rc = self.lib.bnet_exec_op_scaled(self.device, OPS[op], C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                              nbytes, C.c_void_p(base), self.seq, float(scale), C.byref(n))
        else:
            rc = self.lib.bnet_exec_op(self.device, OPS[op], C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), nbytes,
                                       C.c_void_p(base), C.c_void_p(base), self.seq, C.byref(n))
        if rc != 0:
            raise RuntimeError(f"bnet_exec_op({op}) failed")
        return slot, n.value, self.seq

    def done(self, ticket) -> bool:
        slot, n, seq = ticket
        f = self.flags[slot * self.MAX_CHUNKS: slot * self.MAX_CHUNKS + n]
        return bool((f == seq).all())

    def wait(self, ticket, timeout: float = 20.0):
        t0 = time.time()
        while not self.done(ticket):
            if time.time() - t0 > timeout:
                raise TimeoutError("device executor did not complete the job")

    def run(self, op: str, src: torch.Tensor, dst: torch.Tensor, scale: float | None = None):
        self.wait(self.submit(op, src, dst, scale=scale))
        return dst

    MODES = {"default": -1, "msg": 0, "persistent": 1, "oneshot": 2, "ce": 3}

    def run_mode(self, mode: str, op: str, src: torch.Tensor, dst: torch.Tensor, scale: float = 1.0):
        """The same job through an explicit executor mode: "msg" = one launch per message (what NCCL's isend uses),
        "persistent" = resident cluster queues, "oneshot" = one launch per chunk, "ce" = DMA copy engines."""
        if not hasattr(self.lib, "_bnet_mode_decl"):
            self.lib.bnet_exec_op_mode.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                                   C.c_uint64, C.c_float, C.POINTER(C.c_int)]
            self.lib._bnet_mode_decl = True
        torch.cuda.current_stream(self.device).synchronize()
        self.seq += 1
        slot = self.slot
        self.slot = (self.slot + 1) % 64
        base = self.flags.data_ptr() + slot * self.MAX_CHUNKS * 8
        n = C.c_int(0)
        rc = self.lib.bnet_exec_op_mode(self.device, self.MODES[mode], OPS[op], C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                        src.numel() * src.element_size(), C.c_void_p(base), self.seq, float(scale), C.byref(n))
        if rc != 0:
            raise RuntimeError(f"bnet_exec_op_mode({mode}, {op}) failed")
        self.wait((slot, n.value, self.seq))
        return dst
