"""All-reduce that rides the transport (csrc/coll/transport_ring.cc): a ring over the plugin's own connections whose
reduce-scatter hops are *fused isends* — the sending GPU's kernel accumulates into the next rank's buffer while it moves
the data over NVLink (``red.global.add``), so there is no reduce kernel and no staging buffer anywhere.

    ring = TransportRing()                      # every rank; uses torch.distributed to pass the connection handles
    buf = ring.buffer(numel, torch.bfloat16)    # device memory registered with both connections
    ring.all_reduce(buf)                        # in-place sum

The reference only moves bytes for NCCL; this is the collective BASELINE.json's north star asks
to ride the new transport, next to the symmetric-heap collectives of ``SymmComm``."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ..utils.native import load

HANDLE_BYTES = 128
_DT = {torch.float32: 0, torch.bfloat16: 1}
WIRE_FORMATS = {"bf16": 1, "e4m3": 2, "e5m2": 3}       # what travels between the ranks of a compressed all-reduce
_WIRE_BYTES = {"bf16": 2, "e4m3": 1, "e5m2": 1}


def _lib():
    lib = load()
    if not hasattr(lib, "_tring_decl"):
        vp, i = C.c_void_p, C.c_int
        lib.bnet_tring_create.argtypes = [i, i, i, vp, C.POINTER(vp)]
        lib.bnet_tring_connect.argtypes = [vp, vp, i]
        lib.bnet_tring_register.argtypes = [vp, vp, C.c_size_t]
        lib.bnet_tring_allreduce.argtypes = [vp, vp, C.c_size_t, i, C.c_size_t, i, i]
        lib.bnet_tring_register_wire.argtypes = [vp, vp, C.c_size_t, i]
        lib.bnet_tring_allreduce_compressed.argtypes = [vp, vp, C.c_size_t, i, C.c_float, i, C.c_size_t, i, i]
        lib.bnet_tring_stats.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
        lib.bnet_tring_destroy.argtypes = [vp]
        lib.bnet_tring_last_error.argtypes = [vp]
        lib.bnet_tring_last_error.restype = C.c_char_p
        lib.bnet_tring_transport.argtypes = [vp]
        lib.bnet_tring_transport.restype = C.c_char_p
        lib.bnet_tmesh_create.argtypes = [i, i, i, vp, C.POINTER(vp)]
        lib.bnet_tmesh_connect.argtypes = [vp, vp, i]
        lib.bnet_tmesh_register.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t]
        lib.bnet_tmesh_register2.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, i]
        lib.bnet_tmesh_allreduce.argtypes = [vp, vp, vp, C.c_size_t, i, i, C.c_size_t, i, i]
        lib.bnet_tmesh_allreduce2.argtypes = [vp, vp, vp, C.c_size_t, i, i, i, C.c_size_t, i, i]
        lib.bnet_tmesh_stats.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
        lib.bnet_tmesh_destroy.argtypes = [vp]
        lib.bnet_tmesh_last_error.argtypes = [vp]
        lib.bnet_tmesh_last_error.restype = C.c_char_p
        lib.bnet_tmesh_transport.argtypes = [vp]
        lib.bnet_tmesh_transport.restype = C.c_char_p
        lib._tring_decl = True
    return lib


class RingCore:
    """The native ring without torch: callers exchange the 128-byte handles themselves (used by the CPU tests)."""

    def __init__(self, rank: int, world: int, net_dev: int = 0):
        self.lib = _lib()
        self.rank, self.world = rank, world
        self._h = C.c_void_p()
        self._handle = C.create_string_buffer(HANDLE_BYTES)
        if self.lib.bnet_tring_create(rank, world, net_dev, self._handle, C.byref(self._h)) != 0:
            raise RuntimeError("bnet_tring_create failed")

    @property
    def handle(self) -> bytes:
        return bytes(self._handle.raw)

    def _err(self) -> str:
        return self.lib.bnet_tring_last_error(self._h).decode()

    def connect(self, next_handle: bytes, timeout_ms: int = 30000):
        buf = C.create_string_buffer(next_handle, HANDLE_BYTES)
        if self.lib.bnet_tring_connect(self._h, buf, timeout_ms) != 0:
            raise RuntimeError(f"ring connect: {self._err()}")

    @property
    def transport(self) -> str:
        return self.lib.bnet_tring_transport(self._h).decode()

    def register(self, ptr: int, nbytes: int):
        if self.lib.bnet_tring_register(self._h, C.c_void_p(ptr), nbytes) != 0:
            raise RuntimeError(f"ring register: {self._err()}")

    def all_reduce(self, ptr: int, count: int, dtype_code: int, piece_bytes: int = 2 << 20, inflight: int = 16,
                   timeout_ms: int = 60000):
        if self.lib.bnet_tring_allreduce(self._h, C.c_void_p(ptr), count, dtype_code, piece_bytes, inflight, timeout_ms) != 0:
            raise RuntimeError(f"ring all-reduce: {self._err()}")

    def register_wire(self, ptr: int, nbytes: int, device_memory: bool = True):
        """The wire-format mirror of the data buffer (count * 2 bytes for bf16, count bytes for fp8)."""
        if self.lib.bnet_tring_register_wire(self._h, C.c_void_p(ptr), nbytes, 2 if device_memory else 1) != 0:
            raise RuntimeError(f"ring register_wire: {self._err()}")

    def all_reduce_compressed(self, ptr: int, count: int, wire: str = "bf16", scale: float = 1.0, fused: bool = True,
                              piece_elems: int = 1 << 19, inflight: int = 16, timeout_ms: int = 60000):
        if self.lib.bnet_tring_allreduce_compressed(self._h, C.c_void_p(ptr), count, WIRE_FORMATS[wire], float(scale),
                                                    1 if fused else 0, piece_elems, inflight, timeout_ms) != 0:
            raise RuntimeError(f"compressed ring all-reduce: {self._err()}")

    def stats(self) -> dict:
        m, b = C.c_ulonglong(0), C.c_ulonglong(0)
        self.lib.bnet_tring_stats(self._h, C.byref(m), C.byref(b))
        return {"messages": int(m.value), "bytes_sent": int(b.value)}

    def close(self):
        if self._h:
            self.lib.bnet_tring_destroy(self._h)
            self._h = C.c_void_p()


class TransportRing:
    """Ring all-reduce over the bnet transport between the ranks of a torch.distributed group (one GPU per rank)."""

    def __init__(self, group=None, net_dev: int = 0):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world < 2:
            raise ValueError("a ring needs at least two ranks")
        self.core = RingCore(self.rank, self.world, net_dev)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.core.handle, group=group)
        self.core.connect(handles[(self.rank + 1) % self.world])
        dist.barrier(group=group)
        self._buf = None

    @property
    def transport(self) -> str:
        return self.core.transport

    def buffer(self, numel: int, dtype=torch.bfloat16, device=None) -> torch.Tensor:
        """Device memory the ring can reduce in place (registered: the previous rank's kernels write it over NVLink)."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        t = torch.zeros(numel, dtype=dtype, device=device)
        self.core.register(t.data_ptr(), t.numel() * t.element_size())
        self._buf = t
        return t

    def all_reduce(self, t: torch.Tensor, piece_bytes: int = 2 << 20, inflight: int = 16) -> torch.Tensor:
        """In-place sum over the ranks.  `t` must lie inside the registered buffer.  The caller's stream is synchronised
        first (the transport's kernels run on their own streams, like under NCCL's proxy)."""
        if t.dtype not in _DT:
            raise TypeError("transport ring all-reduce supports fp32 and bf16")
        torch.cuda.current_stream(t.device).synchronize()
        self.core.all_reduce(t.data_ptr(), t.numel(), _DT[t.dtype], piece_bytes, inflight)
        return t

    def all_reduce_compressed(self, t: torch.Tensor, wire: str = "bf16", scale: float = 1.0, fused: bool | None = None,
                              piece_elems: int = 1 << 19, inflight: int = 16) -> torch.Tensor:
        """In-place sum of an fp32 tensor with a narrower format on the wire: "bf16" (half the bytes), "e4m3" / "e5m2"
        (a quarter; values are multiplied by `scale` before quantisation and by 1/scale on the way back — choose it so
        that the partial sums stay inside the format: |x| * scale <= 448 for e4m3, 57344 for e5m2).  Every rank ends with
        the same bits.  On the NVLink transport the quantisation is fused into the send (`fused`, default there): the
        sending GPU's kernel converts while it moves, so NVLink carries the narrow format and nothing is staged."""
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError("compressed all-reduce works on contiguous fp32 tensors")
        if wire not in WIRE_FORMATS:
            raise ValueError(f"wire format {wire!r} (one of {sorted(WIRE_FORMATS)})")
        need = t.numel() * _WIRE_BYTES[wire]
        if getattr(self, "_wire", None) is None or self._wire.numel() < need:
            self._wire = torch.zeros(max(need, 64), dtype=torch.uint8, device=t.device)
            self.core.register_wire(self._wire.data_ptr(), self._wire.numel(), device_memory=True)
        if fused is None:
            fused = self.transport == "nvl"
        torch.cuda.current_stream(t.device).synchronize()
        self.core.all_reduce_compressed(t.data_ptr(), t.numel(), wire, scale, fused, piece_elems, inflight)
        return t

    def close(self):
        self.core.close()


MESH_ALGOS = {"one-shot": 0, "two-shot": 1}


class MeshCore:
    """The native all-reduces over a full mesh of plugin connections (csrc/coll/transport_mesh.cc: one-shot and two-shot)
    without torch: callers exchange the 128-byte handles themselves (used by the CPU tests)."""

    def __init__(self, rank: int, world: int, net_dev: int = 0):
        self.lib = _lib()
        self.rank, self.world = rank, world
        self._h = C.c_void_p()
        self._handle = C.create_string_buffer(HANDLE_BYTES)
        if self.lib.bnet_tmesh_create(rank, world, net_dev, self._handle, C.byref(self._h)) != 0:
            raise RuntimeError("bnet_tmesh_create failed")

    @property
    def handle(self) -> bytes:
        return bytes(self._handle.raw)

    def _err(self) -> str:
        return self.lib.bnet_tmesh_last_error(self._h).decode()

    def connect(self, handles: list, timeout_ms: int = 30000):
        assert len(handles) == self.world
        blob = C.create_string_buffer(b"".join(bytes(h).ljust(HANDLE_BYTES, b"\0") for h in handles), HANDLE_BYTES * self.world)
        if self.lib.bnet_tmesh_connect(self._h, blob, timeout_ms) != 0:
            raise RuntimeError(f"mesh connect: {self._err()}")

    @property
    def transport(self) -> str:
        return self.lib.bnet_tmesh_transport(self._h).decode()

    def register(self, in_ptr: int, in_bytes: int, out_ptr: int, out_bytes: int, host_memory: bool = False):
        """host_memory: the buffers are ordinary host memory — reduced with the two-shot schedule over any transport (TCP
        between hosts included), the reduce-scatter pieces added on the host."""
        if self.lib.bnet_tmesh_register2(self._h, C.c_void_p(in_ptr), in_bytes, C.c_void_p(out_ptr), out_bytes,
                                         1 if host_memory else 2) != 0:
            raise RuntimeError(f"mesh register: {self._err()}")

    def all_reduce(self, in_ptr: int, out_ptr: int, count: int, in_dtype: int, out_dtype: int, piece_bytes: int = 1 << 20,
                   inflight: int = 8, timeout_ms: int = 60000, algo: str = "one-shot"):
        if self.lib.bnet_tmesh_allreduce2(self._h, C.c_void_p(in_ptr), C.c_void_p(out_ptr), count, in_dtype, out_dtype,
                                          MESH_ALGOS[algo], piece_bytes, inflight, timeout_ms) != 0:
            raise RuntimeError(f"mesh all-reduce ({algo}): {self._err()}")

    def stats(self) -> dict:
        m, b = C.c_ulonglong(0), C.c_ulonglong(0)
        self.lib.bnet_tmesh_stats(self._h, C.byref(m), C.byref(b))
        return {"messages": int(m.value), "bytes_sent": int(b.value)}

    def close(self):
        if self._h:
            self.lib.bnet_tmesh_destroy(self._h)
            self._h = C.c_void_p()


class TransportMesh:
    """All-reduce over a full mesh of bnet connections; the sending GPU's kernel accumulates into the peer's buffer while
    it moves the data (``red.global.add`` over NVLink).

    ``algo="one-shot"``: every rank sends its input to every peer once — one network step, the latency-optimal companion
    of :class:`TransportRing` for messages up to a few MiB.  ``algo="two-shot"``: reduce-scatter into the slice owners, then
    all-gather by copy — two steps whatever the world size, ``2 (n-1)/n`` x size on the wire (what an NVSwitch wants), every
    rank ends with the same bits, and ``x is out`` (in place) is allowed.  The CollNet table of the plugin runs this one.

        mesh = TransportMesh()
        x, y = mesh.buffers(numel, torch.bfloat16, torch.float32)     # registered input / output
        x.copy_(grad); mesh.all_reduce(x, y)                           # y = sum over ranks of x, accumulated in fp32
        g = mesh.buffer(numel, torch.float32)                          # one registered buffer
        mesh.all_reduce(g, g, algo="two-shot")                         # in place
    """

    def __init__(self, group=None, net_dev: int = 0):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world < 2:
            raise ValueError("a mesh needs at least two ranks")
        self.core = MeshCore(self.rank, self.world, net_dev)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.core.handle, group=group)
        self.core.connect(handles)
        dist.barrier(group=group)
        self._in = self._out = None

    @property
    def transport(self) -> str:
        return self.core.transport

    def buffers(self, numel: int, in_dtype=torch.bfloat16, out_dtype=None, device=None):
        out_dtype = in_dtype if out_dtype is None else out_dtype
        if (in_dtype, out_dtype) not in ((torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32)):
            raise TypeError("supported: fp32 -> fp32, bf16 -> bf16, bf16 -> fp32")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._in = torch.zeros(numel, dtype=in_dtype, device=device)
        self._out = torch.zeros(numel, dtype=out_dtype, device=device)
        self.core.register(self._in.data_ptr(), self._in.numel() * self._in.element_size(), self._out.data_ptr(),
                           self._out.numel() * self._out.element_size(), host_memory=device.type == "cpu")
        return self._in, self._out

    def buffer(self, numel: int, dtype=torch.float32, device=None) -> torch.Tensor:
        """One registered buffer, for in-place two-shot all-reduces.  ``device="cpu"``: host memory — reduced over any
        transport (TCP between hosts, shared memory), the reduce-scatter pieces added on the host."""
        if dtype not in _DT:
            raise TypeError("supported: fp32, bf16")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._in = self._out = torch.zeros(numel, dtype=dtype, device=device)
        nbytes = self._in.numel() * self._in.element_size()
        self.core.register(self._in.data_ptr(), nbytes, self._in.data_ptr(), nbytes, host_memory=device.type == "cpu")
        return self._in

    def all_reduce(self, x: torch.Tensor, out: torch.Tensor, piece_bytes: int = 1 << 20, inflight: int = 8,
                   algo: str = "one-shot") -> torch.Tensor:
        """out = sum over ranks of x (both inside the registered buffers, same number of elements)."""
        if x.numel() != out.numel() or x.dtype not in _DT or out.dtype not in _DT:
            raise TypeError("mesh all-reduce: fp32 / bf16 tensors of equal length")
        if x.is_cuda:
            torch.cuda.current_stream(x.device).synchronize()
        elif algo != "two-shot":
            raise ValueError("host memory is reduced with algo='two-shot'")
        self.core.all_reduce(x.data_ptr(), out.data_ptr(), x.numel(), _DT[x.dtype], _DT[out.dtype], piece_bytes, inflight,
                             algo=algo)
        return out

    def close(self):
        self.core.close()
