"""Functional wrappers over the symmetric-memory kernels (csrc/cuda/coll.cu)."""
from __future__ import annotations

import ctypes as C

import torch

from ..parallel.comm import DT, SymmComm


def all_reduce(comm: SymmComm, t: torch.Tensor, op: str = "sum", algo: str = "auto", stream=None) -> torch.Tensor:
    """In-place all-reduce of ANY CUDA tensor over the symmetric-memory kernels, by the cheapest route that applies:

    * small messages (<= 4 * comm.ll_words bytes, any memory): the barrier-free LL kernel — one NVLink store latency;
    * tensors that live in the symmetric heap (``comm.alloc``) with a size the kernels can shard (a multiple of
      16 bytes * world): in place, NVLS in-switch reduction or peer loads/stores (``algo``);
    * everything else: staged through the heap in chunks (``comm.all_reduce_tensor``).

    ``algo`` only steers the heap kernels ("nvls", "p2p", "auto")."""
    if comm.world == 1:
        return t
    nbytes = t.numel() * t.element_size()
    if algo == "auto" and getattr(comm, "_ll", None) is not None and t.is_contiguous() and nbytes <= 4 * comm.ll_words \
            and t.dtype in DT and op in ("sum", "avg", "max", "min"):
        return comm.all_reduce_ll(t, op=op, stream=stream)
    in_heap = t.is_contiguous() and comm.heap_ptr <= t.data_ptr() and t.data_ptr() + nbytes <= comm.heap_ptr + comm.heap_bytes
    if in_heap and nbytes % (16 * comm.world) == 0 and (t.data_ptr() - comm.heap_ptr) % 16 == 0:
        return comm.all_reduce(t, op=op, algo=algo if algo != "auto" else "auto", stream=stream)
    return comm.all_reduce_tensor(t, op=op, algo=algo, stream=stream)


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
