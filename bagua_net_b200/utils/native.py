"""Loading the native library and thin ctypes wrappers over its C API (csrc/capi.cc)."""
from __future__ import annotations

import ctypes
import json
import os
import threading

from .. import LIB_DIR, LIB_NAME

_lock = threading.Lock()
_libs: dict[str, ctypes.CDLL] = {}


def load(name: str = LIB_NAME) -> ctypes.CDLL:
    """dlopen the in-tree library.  Fails loudly when it cannot be built or loaded —
    there is no pure-Python fallback for the data path."""
    with _lock:
        if name in _libs:
            return _libs[name]
        path = os.path.join(LIB_DIR, name)
        if not os.path.exists(path):
            from .._build import build_native

            build_native()      # takes an inter-process file lock: ranks of one job do not build concurrently
        lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
        _declare(lib)
        _libs[name] = lib
        return lib


def _declare(lib: ctypes.CDLL) -> None:
    c = ctypes
    lib.bnet_version.restype = c.c_char_p
    lib.bnet_chunk_size.restype = c.c_ulonglong
    lib.bnet_chunk_size.argtypes = [c.c_ulonglong] * 3
    lib.bnet_chunk_count.restype = c.c_ulonglong
    lib.bnet_chunk_count.argtypes = [c.c_ulonglong] * 3
    lib.bnet_parse_user_pass_addr.argtypes = [c.c_char_p, c.c_char_p, c.c_char_p, c.c_char_p, c.c_int]
    lib.bnet_sockaddr_roundtrip.argtypes = [c.c_char_p, c.c_char_p, c.c_int]
    lib.bnet_base64.argtypes = [c.c_char_p, c.c_char_p, c.c_int]
    lib.bnet_if_filter_accepts.argtypes = [c.c_char_p, c.c_char_p]
    lib.bnet_find_interfaces.argtypes = [c.c_char_p, c.c_int, c.c_char_p, c.c_int]
    lib.bnet_config_json.argtypes = [c.c_char_p, c.c_int]
    lib.bnet_metrics_text.argtypes = [c.c_char_p, c.c_int]
    lib.bnet_trace_json.argtypes = [c.c_char_p, c.c_int]
    lib.bnet_http_send.argtypes = [c.c_char_p] * 6
    lib.bnet_comm_transport.restype = c.c_char_p
    lib.bnet_comm_transport.argtypes = [c.c_void_p]
    lib.bnet_exec_stats.argtypes = [c.POINTER(c.c_ulonglong)]


def _text(fn, *args, cap: int = 1 << 20) -> str:
    buf = ctypes.create_string_buffer(cap)
    fn(*args, buf, cap)
    return buf.value.decode()


def version() -> str:
    return load().bnet_version().decode()


def chunk_size(total: int, min_chunksize: int, expected_nchunks: int) -> int:
    """max(ceil(total/n), min) — reference: src/utils.rs:200-205."""
    return int(load().bnet_chunk_size(total, min_chunksize, expected_nchunks))


def chunk_count(total: int, min_chunksize: int, expected_nchunks: int) -> int:
    return int(load().bnet_chunk_count(total, min_chunksize, expected_nchunks))


def parse_user_pass_and_addr(raw: str):
    """'[user:pass@]host:port' -> (user, pass, addr) — reference: src/utils.rs:180-198."""
    lib = load()
    u, p, a = (ctypes.create_string_buffer(256) for _ in range(3))
    if lib.bnet_parse_user_pass_addr(raw.encode(), u, p, a, 256) != 0:
        return None
    return u.value.decode(), p.value.decode(), a.value.decode()


def sockaddr_roundtrip(addr: str) -> str | None:
    out = ctypes.create_string_buffer(128)
    if load().bnet_sockaddr_roundtrip(addr.encode(), out, 128) != 0:
        return None
    return out.value.decode()


def base64(s: str) -> str:
    return _text(load().bnet_base64, s.encode(), cap=4096)


def if_filter_accepts(spec: str, ifname: str) -> bool:
    return bool(load().bnet_if_filter_accepts(spec.encode(), ifname.encode()))


def find_interfaces(spec: str | None = None, family: int = -2) -> list[dict]:
    """NIC discovery with NCCL_SOCKET_IFNAME syntax — reference: src/utils.rs:32-130."""
    return json.loads(_text(load().bnet_find_interfaces, (spec or "").encode(), family))


def config() -> dict:
    return json.loads(_text(load().bnet_config_json))


def reload_config() -> None:
    load().bnet_config_reload()


def metrics_text() -> str:
    return _text(load().bnet_metrics_text)


def trace_json() -> dict:
    return json.loads(_text(load().bnet_trace_json, cap=8 << 20))


def telemetry_flush() -> int:
    return int(load().bnet_telemetry_flush())


def exec_stats() -> dict:
    arr = (ctypes.c_ulonglong * 5)()
    load().bnet_exec_stats(arr)
    return dict(zip(["jobs", "chunks", "bytes", "launches", "persistent"], [int(x) for x in arr]))
