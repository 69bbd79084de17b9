"""Functional wrappers over the symmetric-memory kernels (csrc/cuda/coll.cu)."""
from __future__ import annotations

import ctypes as C

import torch

from ..parallel.comm import DT, SymmComm


def all_reduce(comm: SymmComm, t: torch.Tensor, op: str = "sum", algo: str = "auto", **kw) -> torch.Tensor:
    """In-place all-reduce of a tensor allocated with ``comm.alloc``.
    algo: "nvls" (in-switch multimem reduction), "p2p" (peer loads + peer stores) or "auto"."""
    return comm.all_reduce(t, op=op, algo=algo, **kw)


def all_reduce_oneshot(comm: SymmComm, t: torch.Tensor, out: torch.Tensor | None = None, op: str = "sum", **kw):
    """Latency-optimal out-of-place all-reduce: every rank reads every peer once."""
    if out is None:
        out = torch.empty_like(t)
    return comm.all_reduce_oneshot(t, out, op=op, **kw)


def fused_allreduce_sgd(comm: SymmComm, grad, param, master, mom, lr, momentum=0.0, weight_decay=0.0, **kw):
    """grad mean-reduce + SGD(momentum, wd) on the fp32 shard + parameter broadcast, one kernel."""
    return comm.fused_allreduce_sgd(grad, param, master, mom, lr, momentum, weight_decay, **kw)


class _Item(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst_elem_off", C.c_uint64), ("numel", C.c_uint64)]


def pack_cast(comm: SymmComm, tensors: list[torch.Tensor], dst: torch.Tensor, offsets: list[int] | None = None,
              scale: float = 1.0, stream=None) -> torch.Tensor:
    """Pack ``tensors`` (one dtype) into the flat ``dst`` at element ``offsets`` with a dtype cast and
    scale, in ONE launch (K5: bf16<->fp32 cast fused with the bucket fill)."""
    if not tensors:
        return dst
    src_dt = tensors[0].dtype
    assert all(t.dtype == src_dt and t.is_contiguous() and t.is_cuda for t in tensors)
    if offsets is None:
        offsets, o = [], 0
        for t in tensors:
            offsets.append(o)
            o += t.numel()
    assert offsets[-1] + tensors[-1].numel() <= dst.numel()
    n = len(tensors)
    host = torch.empty(n * 3, dtype=torch.int64).pin_memory()
    for i, (t, o) in enumerate(zip(tensors, offsets)):
        host[3 * i], host[3 * i + 1], host[3 * i + 2] = t.data_ptr(), o, t.numel()
    dev_items = host.to(dst.device, non_blocking=True)
    s = stream if stream is not None else torch.cuda.current_stream(dst.device)
    rc = comm.lib.bnet_pack_cast(C.c_void_p(dev_items.data_ptr()), n, C.c_void_p(dst.data_ptr()), DT[src_dt],
                                 DT[dst.dtype], scale, max(t.numel() for t in tensors), C.c_void_p(s.cuda_stream))
    if rc < 0:
        raise RuntimeError(comm.lib.bnet_coll_last_error().decode())
    comm.launches += rc
    dev_items.record_stream(s)
    return dst
