"""An independent peer that speaks the REFERENCE's TCP wire format, written from its specification only (SURVEY.md
section 2.2 / 2.3 and Appendix B; reference src/implement/nthread_per_socket_backend.rs:305-522 for BASIC,
tokio_backend.rs:336-592 for TOKIO) with plain Python sockets — none of this repository's native code.

    handle   the listener's `struct sockaddr` in the first 16 bytes of NCCL's handle buffer (lib.rs:157-158)
    connect  nstreams data connections, each announcing itself with its 8-byte big-endian stream id, then the control
             connection announcing id == nstreams
    BASIC    per message: u64 big-endian length on the control stream; payload cut into max(ceil(len / nstreams),
             min_chunksize)-byte chunks dealt round-robin over the data streams, the cursor persisting across messages
    TOKIO    per message: u32 big-endian length; at most nstreams chunks, chunk j on stream j, every message from stream 0

The loopback tests put it on one end of a connection and the plugin (BNET_WIRE_COMPAT=1) on the other: a host that still
runs bagua-net and a host that runs this library can be in one job (docs/MIGRATION.md)."""
from __future__ import annotations

import socket
import struct


def chunk_size(total: int, min_chunk: int, nstreams: int) -> int:
    return max(-(-total // nstreams), min_chunk)


def _read_exact(s: socket.socket, n: int) -> bytes:
    out = bytearray()
    while len(out) < n:
        b = s.recv(min(n - len(out), 1 << 20))
        if not b:
            raise EOFError(f"peer closed after {len(out)} of {n} bytes")
        out += b
    return bytes(out)


def parse_handle(handle: bytes):
    """(ip, port) from a `struct sockaddr_in` at the start of the handle."""
    family = struct.unpack_from("=H", handle, 0)[0]
    assert family == socket.AF_INET, f"address family {family}"
    port = struct.unpack_from("!H", handle, 2)[0]
    return socket.inet_ntoa(handle[4:8]), port


def make_handle(ip: str, port: int, size: int = 64) -> bytes:
    sa = struct.pack("=H", socket.AF_INET) + struct.pack("!H", port) + socket.inet_aton(ip) + b"\0" * 8
    return sa + b"\0" * (size - len(sa))


class RefSender:
    def __init__(self, handle: bytes, impl: str = "BASIC", nstreams: int = 2, min_chunk: int | None = None):
        self.impl, self.n = impl, nstreams
        self.min_chunk = min_chunk if min_chunk is not None else (1048576 if impl == "BASIC" else 65535)
        addr = parse_handle(handle)
        self.streams = []
        for i in range(nstreams):
            s = socket.create_connection(addr, timeout=30)
            s.sendall(struct.pack("!Q", i))
            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            self.streams.append(s)
        self.ctrl = socket.create_connection(addr, timeout=30)
        self.ctrl.sendall(struct.pack("!Q", nstreams))
        self.ctrl.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        self.cursor = 0

    def send(self, data: bytes):
        if self.impl == "BASIC":
            self.ctrl.sendall(struct.pack("!Q", len(data)))
            if not data:
                return
            c = chunk_size(len(data), self.min_chunk, self.n)
            for off in range(0, len(data), c):
                self.streams[self.cursor].sendall(data[off:off + c])
                self.cursor = (self.cursor + 1) % self.n
        else:
            self.ctrl.sendall(struct.pack("!I", len(data)))
            if not data:
                return
            c = chunk_size(len(data), self.min_chunk, self.n)
            for j, off in enumerate(range(0, len(data), c)):
                self.streams[j].sendall(data[off:off + c])

    def close(self):
        for s in self.streams + [self.ctrl]:
            s.close()


class RefReceiver:
    def __init__(self, impl: str = "BASIC", nstreams: int = 2, min_chunk: int | None = None):
        self.impl, self.n = impl, nstreams
        self.min_chunk = min_chunk if min_chunk is not None else (1048576 if impl == "BASIC" else 65535)
        self.lsock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.lsock.bind(("127.0.0.1", 0))
        self.lsock.listen(64)
        self.handle = make_handle(*self.lsock.getsockname())
        self.streams, self.ctrl, self.cursor = {}, None, 0

    def accept(self):
        self.lsock.settimeout(30)
        for _ in range(self.n + 1):
            s, _ = self.lsock.accept()
            s.settimeout(60)
            sid = struct.unpack("!Q", _read_exact(s, 8))[0]
            if sid == self.n:
                self.ctrl = s
            else:
                self.streams[sid] = s          # keyed by the announced id: arrival order does not matter
        assert self.ctrl is not None and sorted(self.streams) == list(range(self.n))

    def recv(self) -> bytes:
        if self.impl == "BASIC":
            n = struct.unpack("!Q", _read_exact(self.ctrl, 8))[0]
        else:
            n = struct.unpack("!I", _read_exact(self.ctrl, 4))[0]
        if n == 0:
            return b""
        c = chunk_size(n, self.min_chunk, self.n)
        out = bytearray()
        for j, off in enumerate(range(0, n, c)):
            if self.impl == "BASIC":
                sid, self.cursor = self.cursor, (self.cursor + 1) % self.n
            else:
                sid = j
            out += _read_exact(self.streams[sid], min(c, n - off))
        return bytes(out)

    def close(self):
        for s in list(self.streams.values()) + [self.ctrl, self.lsock]:
            if s is not None:
                s.close()
