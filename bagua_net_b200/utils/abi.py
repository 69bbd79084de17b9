"""ctypes view of the exported ncclNet tables + a small driver class.

This lets Python call the plugin exactly the way NCCL does — through the
``ncclNetPlugin_vN`` data symbols — which is how the loopback tests exercise the
ABI (SURVEY.md §4 "ABI test", BASELINE.json config #1).  Layouts mirror
include/bnet/nccl_net_abi.h.
"""
from __future__ import annotations

import ctypes as C
import time

from .native import load

ncclSuccess = 0
NCCL_PTR_HOST, NCCL_PTR_CUDA = 1, 2
HANDLE_BYTES = {3: 64, 4: 64, 5: 128, 6: 128, 7: 128, 8: 128, 9: 128, 10: 128}

_vp, _vpp, _i, _ip = C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int)
_sz, _szp = C.c_size_t, C.POINTER(C.c_size_t)
_res = C.c_int
_logger_t = C.c_void_p


def _fn(*args):
    return C.CFUNCTYPE(_res, *args)


class PropsV4(C.Structure):
    _fields_ = [("name", C.c_char_p), ("pciPath", C.c_char_p), ("guid", C.c_uint64), ("ptrSupport", _i),
                ("speed", _i), ("port", _i), ("maxComms", _i)]


class PropsV6(C.Structure):
    _fields_ = [("name", C.c_char_p), ("pciPath", C.c_char_p), ("guid", C.c_uint64), ("ptrSupport", _i),
                ("speed", _i), ("port", _i), ("latency", C.c_float), ("maxComms", _i), ("maxRecvs", _i)]


class PropsV8(C.Structure):
    _fields_ = [("name", C.c_char_p), ("pciPath", C.c_char_p), ("guid", C.c_uint64), ("ptrSupport", _i),
                ("regIsGlobal", _i), ("speed", _i), ("port", _i), ("latency", C.c_float), ("maxComms", _i),
                ("maxRecvs", _i), ("netDeviceType", _i), ("netDeviceVersion", _i)]


class VProps(C.Structure):
    _fields_ = [("ndevs", _i), ("devs", _i * 4)]


class PropsV9(C.Structure):
    _fields_ = [("name", C.c_char_p), ("pciPath", C.c_char_p), ("guid", C.c_uint64), ("ptrSupport", _i),
                ("regIsGlobal", _i), ("forceFlush", _i), ("speed", _i), ("port", _i), ("latency", C.c_float),
                ("maxComms", _i), ("maxRecvs", _i), ("netDeviceType", _i), ("netDeviceVersion", _i),
                ("vProps", VProps), ("maxP2pBytes", _sz), ("maxCollBytes", _sz)]


class NetV4(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("init", _fn(_logger_t)),
        ("devices", _fn(_ip)),
        ("getProperties", _fn(_i, C.POINTER(PropsV4))),
        ("listen", _fn(_i, _vp, _vpp)),
        ("connect", _fn(_i, _vp, _vpp)),
        ("accept", _fn(_vp, _vpp)),
        ("regMr", _fn(_vp, _vp, _i, _i, _vpp)),
        ("deregMr", _fn(_vp, _vp)),
        ("isend", _fn(_vp, _vp, _i, _vp, _vpp)),
        ("irecv", _fn(_vp, _vp, _i, _vp, _vpp)),
        ("iflush", _fn(_vp, _vp, _i, _vp, _vpp)),   # v3: flush(recvComm, data, size, mhandle)
        ("test", _fn(_vp, _ip, _ip)),
        ("closeSend", _fn(_vp)),
        ("closeRecv", _fn(_vp)),
        ("closeListen", _fn(_vp)),
    ]


class NetV6(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("init", _fn(_logger_t)),
        ("devices", _fn(_ip)),
        ("getProperties", _fn(_i, C.POINTER(PropsV6))),
        ("listen", _fn(_i, _vp, _vpp)),
        ("connect", _fn(_i, _vp, _vpp)),
        ("accept", _fn(_vp, _vpp)),
        ("regMr", _fn(_vp, _vp, _i, _i, _vpp)),
        ("regMrDmaBuf", _fn(_vp, _vp, _sz, _i, C.c_uint64, _i, _vpp)),
        ("deregMr", _fn(_vp, _vp)),
        ("isend", _fn(_vp, _vp, _i, _i, _vp, _vpp)),
        ("irecv", _fn(_vp, _i, _vpp, _ip, _ip, _vpp, _vpp)),
        ("iflush", _fn(_vp, _i, _vpp, _ip, _vpp, _vpp)),
        ("test", _fn(_vp, _ip, _ip)),
        ("closeSend", _fn(_vp)),
        ("closeRecv", _fn(_vp)),
        ("closeListen", _fn(_vp)),
    ]


class NetV8(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("init", _fn(_logger_t)),
        ("devices", _fn(_ip)),
        ("getProperties", _fn(_i, C.POINTER(PropsV8))),
        ("listen", _fn(_i, _vp, _vpp)),
        ("connect", _fn(_i, _vp, _vpp, _vpp)),
        ("accept", _fn(_vp, _vpp, _vpp)),
        ("regMr", _fn(_vp, _vp, _sz, _i, _vpp)),
        ("regMrDmaBuf", _fn(_vp, _vp, _sz, _i, C.c_uint64, _i, _vpp)),
        ("deregMr", _fn(_vp, _vp)),
        ("isend", _fn(_vp, _vp, _i, _i, _vp, _vpp)),
        ("irecv", _fn(_vp, _i, _vpp, _ip, _ip, _vpp, _vpp)),
        ("iflush", _fn(_vp, _i, _vpp, _ip, _vpp, _vpp)),
        ("test", _fn(_vp, _ip, _ip)),
        ("closeSend", _fn(_vp)),
        ("closeRecv", _fn(_vp)),
        ("closeListen", _fn(_vp)),
        ("getDeviceMr", _fn(_vp, _vp, _vpp)),
        ("irecvConsumed", _fn(_vp, _i, _vp)),
    ]


class NetV10(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("init", _fn(_logger_t, C.c_void_p)),
        ("devices", _fn(_ip)),
        ("getProperties", _fn(_i, C.POINTER(PropsV9))),
        ("listen", _fn(_i, _vp, _vpp)),
        ("connect", _fn(_i, _vp, _vp, _vpp, _vpp)),
        ("accept", _fn(_vp, _vpp, _vpp)),
        ("regMr", _fn(_vp, _vp, _sz, _i, _vpp)),
        ("regMrDmaBuf", _fn(_vp, _vp, _sz, _i, C.c_uint64, _i, _vpp)),
        ("deregMr", _fn(_vp, _vp)),
        ("isend", _fn(_vp, _vp, _sz, _i, _vp, _vp, _vpp)),
        ("irecv", _fn(_vp, _i, _vpp, _szp, _ip, _vpp, _vpp, _vpp)),
        ("iflush", _fn(_vp, _i, _vpp, _ip, _vpp, _vpp)),
        ("test", _fn(_vp, _ip, _ip)),
        ("closeSend", _fn(_vp)),
        ("closeRecv", _fn(_vp)),
        ("closeListen", _fn(_vp)),
        ("getDeviceMr", _fn(_vp, _vp, _vpp)),
        ("irecvConsumed", _fn(_vp, _i, _vp)),
        ("makeVDevice", C.c_void_p),
    ]


_TABLES = {3: NetV4, 4: NetV4, 5: None, 6: NetV6, 8: NetV8, 10: NetV10}
_PROPS = {3: PropsV4, 4: PropsV4, 6: PropsV6, 8: PropsV8, 10: PropsV9}


class CollNetV6(C.Structure):       # include/bnet/nccl_net_abi.h: ncclCollNet_v6_t (v7 differs in the properties struct only)
    _fields_ = [
        ("name", C.c_char_p),
        ("init", _fn(_logger_t)),
        ("devices", _fn(_ip)),
        ("getProperties", _fn(_i, C.POINTER(PropsV6))),
        ("listen", _fn(_i, _vp, _vpp)),
        ("connect", _fn(_vpp, _i, _i, _vp, _vpp)),
        ("reduceSupport", _fn(_i, _i, _ip)),
        ("regMr", _fn(_vp, _vp, _i, _i, _vpp)),
        ("regMrDmaBuf", _fn(_vp, _vp, _sz, _i, C.c_uint64, _i, _vpp)),
        ("deregMr", _fn(_vp, _vp)),
        ("iallreduce", _fn(_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vpp)),
        ("iflush", _fn(_vp, _vp, _i, _vp, _vpp)),
        ("test", _fn(_vp, _ip, _ip)),
        ("closeColl", _fn(_vp)),
        ("closeListen", _fn(_vp)),
    ]


class CollNetV8(C.Structure):       # ncclCollNet_v8_t
    _fields_ = [
        ("name", C.c_char_p),
        ("init", _fn(_logger_t)),
        ("devices", _fn(_ip)),
        ("getProperties", _fn(_i, C.POINTER(PropsV8))),
        ("listen", _fn(_i, _vp, _vpp)),
        ("connect", _fn(_vpp, _i, _i, _vp, _vpp)),
        ("reduceSupport", _fn(_i, _i, _ip)),
        ("regMr", _fn(_vp, _vp, _sz, _i, _vpp)),
        ("regMrDmaBuf", _fn(_vp, _vp, _sz, _i, C.c_uint64, _i, _vpp)),
        ("deregMr", _fn(_vp, _vp)),
        ("iallreduce", _fn(_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vpp)),
        ("iallgather", _fn(_vp, _vp, _i, _vp, _sz, _sz, _sz, _vp, _vpp)),
        ("ireducescatter", _fn(_vp, _i, _vp, _vp, _sz, _sz, _sz, _i, _i, _vp, _vpp)),
        ("iflush", _fn(_vp, _vp, _i, _vp, _vpp)),
        ("test", _fn(_vp, _ip, _ip)),
        ("closeColl", _fn(_vp)),
        ("closeListen", _fn(_vp)),
    ]


class CollNetV10(C.Structure):      # ncclCollNet_v9_t / _v10_t (exported by libnccl-net-bnetx.so)
    _fields_ = [
        ("name", C.c_char_p),
        ("init", _fn(_logger_t)),
        ("devices", _fn(_ip)),
        ("getProperties", _fn(_i, C.POINTER(PropsV9))),
        ("listen", _fn(_i, _vp, _vpp)),
        ("connect", _fn(_vpp, _i, _i, _vp, _vpp)),
        ("reduceSupport", _fn(_i, _i, _ip)),
        ("regMr", _fn(_vp, _vp, _sz, _i, _vpp)),
        ("regMrDmaBuf", _fn(_vp, _vp, _sz, _i, C.c_uint64, _i, _vpp)),
        ("deregMr", _fn(_vp, _vp)),
        ("iallreduce", _fn(_vp, _vp, _vp, _sz, _i, _i, _vp, _vp, _vpp)),
        ("iallgather", _fn(_vp, _vp, _i, _vp, _sz, _sz, _sz, _vp, _vpp)),
        ("ireducescatter", _fn(_vp, _i, _vp, _vp, _sz, _sz, _sz, _i, _i, _vp, _vpp)),
        ("iflush", _fn(_vp, _vp, _i, _vp, _vpp)),
        ("test", _fn(_vp, _ip, _ip)),
        ("closeColl", _fn(_vp)),
        ("closeListen", _fn(_vp)),
        ("makeVDevice", _fn(_ip, _vp)),
    ]


_COLL_TABLES = {6: CollNetV6, 8: CollNetV8, 9: CollNetV10, 10: CollNetV10}
_COLL_PROPS = {6: PropsV6, 8: PropsV8, 9: PropsV9, 10: PropsV9}
ncclSum, ncclFloat32, ncclBfloat16 = 0, 7, 9


# ncclProfilerCallback_t(void** eHandle, int type, void* pHandle, int64_t pluginId, void* extData)
PROFILER_CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64, C.c_void_p)


class ProfilerEventDescr(C.Structure):   # include/bnet/bnet_profiler.h: bnetProfilerEventDescr_v1_t
    _fields_ = [("type", C.c_uint8), ("path", C.c_uint8), ("reserved", C.c_uint16), ("tag", C.c_int32),
                ("comm_id", C.c_uint64), ("request_id", C.c_uint64), ("length", C.c_size_t)]


BNET_PROFILER_PLUGIN_ID = (0x42 << 16) | 1


class PluginError(RuntimeError):
    def __init__(self, what: str, code: int):
        super().__init__(f"{what} -> ncclResult {code}")
        self.code = code


class NetPlugin:
    """Drives one exported table.  Sizes/handles follow the ABI version chosen."""

    def __init__(self, version: int = 8, lib_name: str | None = None):
        if _TABLES.get(version) is None:
            raise ValueError(f"ABI v{version} has no ctypes table here")
        self.version = version
        self.lib = load(lib_name) if lib_name else load()
        self.tab = _TABLES[version].in_dll(self.lib, f"ncclNetPlugin_v{version}")
        self.name = self.tab.name.decode()

    def _chk(self, what, rc):
        if rc != ncclSuccess:
            raise PluginError(what, rc)

    def init(self, profiler=None):
        """profiler (v10 only): a PROFILER_CB(...) object — NCCL's ncclProfilerCallback_t — kept alive by the caller."""
        if self.version >= 10:
            self._prof = profiler
            self._chk("init", self.tab.init(None, C.cast(profiler, C.c_void_p) if profiler is not None else None))
        else:
            self._chk("init", self.tab.init(None))

    def devices(self) -> int:
        n = C.c_int(0)
        self._chk("devices", self.tab.devices(C.byref(n)))
        return n.value

    def get_properties(self, dev: int) -> dict:
        p = _PROPS[self.version]()
        self._chk("getProperties", self.tab.getProperties(dev, C.byref(p)))
        out = {}
        for f, _t in p._fields_:
            v = getattr(p, f)
            if isinstance(v, bytes):
                v = v.decode()
            if isinstance(v, VProps):
                v = {"ndevs": v.ndevs, "devs": list(v.devs)}
            out[f] = v
        return out

    def listen(self, dev: int = 0):
        handle = C.create_string_buffer(HANDLE_BYTES[self.version])
        comm = C.c_void_p()
        self._chk("listen", self.tab.listen(dev, handle, C.byref(comm)))
        return bytes(handle.raw), comm

    def connect(self, handle: bytes, dev: int = 0, timeout: float = 30.0):
        hbuf = C.create_string_buffer(handle, HANDLE_BYTES[self.version])
        comm = C.c_void_p()
        t0 = time.time()
        while True:
            if self.version >= 10:
                rc = self.tab.connect(dev, None, hbuf, C.byref(comm), None)
            elif self.version >= 7:
                rc = self.tab.connect(dev, hbuf, C.byref(comm), None)
            else:
                rc = self.tab.connect(dev, hbuf, C.byref(comm))
            self._chk("connect", rc)
            if comm.value:
                return comm
            if time.time() - t0 > timeout:
                raise TimeoutError("connect")

    def accept(self, lcomm, timeout: float = 30.0, poll: bool = True):
        comm = C.c_void_p()
        t0 = time.time()
        while True:
            if self.version >= 7:
                rc = self.tab.accept(lcomm, C.byref(comm), None)
            else:
                rc = self.tab.accept(lcomm, C.byref(comm))
            self._chk("accept", rc)
            if comm.value or not poll:
                return comm if comm.value else None
            if time.time() - t0 > timeout:
                raise TimeoutError("accept")
            time.sleep(0.0005)

    def reg_mr(self, comm, addr: int, size: int, ptr_type: int = NCCL_PTR_HOST):
        mh = C.c_void_p()
        self._chk("regMr", self.tab.regMr(comm, C.c_void_p(addr), size, ptr_type, C.byref(mh)))
        return mh

    def dereg_mr(self, comm, mh):
        self._chk("deregMr", self.tab.deregMr(comm, mh))

    def isend(self, comm, addr: int, size: int, mh=None, tag: int = 0, phandle: int | None = None):
        req = C.c_void_p()
        if self.version >= 10:
            rc = self.tab.isend(comm, C.c_void_p(addr), size, tag, mh, C.c_void_p(phandle) if phandle else None, C.byref(req))
        elif self.version >= 5:
            rc = self.tab.isend(comm, C.c_void_p(addr), size, tag, mh, C.byref(req))
        else:
            rc = self.tab.isend(comm, C.c_void_p(addr), size, mh, C.byref(req))
        self._chk("isend", rc)
        return req if req.value else None

    def irecv(self, comm, addr: int, size: int, mh=None, tag: int = 0, phandle: int | None = None):
        req = C.c_void_p()
        if self.version >= 5:
            data = (C.c_void_p * 1)(addr)
            tags = (C.c_int * 1)(tag)
            mhs = (C.c_void_p * 1)(mh.value if mh is not None and mh.value else None)
            if self.version >= 10:
                sizes = (C.c_size_t * 1)(size)
                phs = (C.c_void_p * 1)(phandle) if phandle else None
                rc = self.tab.irecv(comm, 1, data, sizes, tags, mhs, phs, C.byref(req))
            else:
                sizes = (C.c_int * 1)(size)
                rc = self.tab.irecv(comm, 1, data, sizes, tags, mhs, C.byref(req))
        else:
            rc = self.tab.irecv(comm, C.c_void_p(addr), size, mh, C.byref(req))
        self._chk("irecv", rc)
        return req if req.value else None

    def irecv_group(self, comm, addrs: list, sizes: list, mhs: list, tags: list | None = None):
        """One grouped receive (ncclNet v5+, ``maxRecvs`` > 1): ``len(addrs)`` buffers under ONE request, matched in order to
        the sender's next isends.  Returns the request, or None when the plugin asks to try again."""
        if self.version < 5:
            raise ValueError("grouped receives exist from ncclNet v5 on")
        n = len(addrs)
        req = C.c_void_p()
        data = (C.c_void_p * n)(*addrs)
        tg = (C.c_int * n)(*(tags or [0] * n))
        mh = (C.c_void_p * n)(*[(m.value if m is not None and m.value else None) for m in mhs])
        if self.version >= 10:
            rc = self.tab.irecv(comm, n, data, (C.c_size_t * n)(*sizes), tg, mh, None, C.byref(req))
        else:
            rc = self.tab.irecv(comm, n, data, (C.c_int * n)(*sizes), tg, mh, C.byref(req))
        self._chk("irecv(group)", rc)
        return req if req.value else None

    def test_group(self, req, n: int):
        """(done, [size of every entry]) of a grouped receive."""
        done, sizes = C.c_int(0), (C.c_int * n)()
        self._chk("test", self.tab.test(req, C.byref(done), sizes))
        return bool(done.value), list(sizes)

    def iflush(self, comm, addr: int, size: int, mh=None):
        req = C.c_void_p()
        if self.version == 3:
            self._chk("flush", self.tab.iflush(comm, C.c_void_p(addr), size, mh, None))
            return None
        if self.version >= 5:
            data = (C.c_void_p * 1)(addr)
            sizes = (C.c_int * 1)(size)
            mhs = (C.c_void_p * 1)(mh.value if mh is not None and mh.value else None)
            rc = self.tab.iflush(comm, 1, data, sizes, mhs, C.byref(req))
        else:
            rc = self.tab.iflush(comm, C.c_void_p(addr), size, mh, C.byref(req))
        self._chk("iflush", rc)
        return req if req.value else None

    def test(self, req):
        done, size = C.c_int(0), C.c_int(0)
        self._chk("test", self.tab.test(req, C.byref(done), C.byref(size)))
        return bool(done.value), size.value

    def wait(self, req, timeout: float = 60.0) -> int:
        t0 = time.time()
        while True:
            done, size = self.test(req)
            if done:
                return size
            if time.time() - t0 > timeout:
                raise TimeoutError("request did not complete")

    def close_send(self, comm):
        self._chk("closeSend", self.tab.closeSend(comm))

    def close_recv(self, comm):
        self._chk("closeRecv", self.tab.closeRecv(comm))

    def close_listen(self, comm):
        self._chk("closeListen", self.tab.closeListen(comm))

    def transport_of(self, comm) -> str:
        return self.lib.bnet_comm_transport(comm).decode()


class CollNetPlugin:
    """Drives one exported ``ncclCollNetPlugin_vN`` table the way NCCL does (csrc/plugin/collnet.cc): listen on every rank,
    connect with everybody's handles, register the send and receive buffers, ``iallreduce`` + ``test``."""

    def __init__(self, version: int = 8, lib_name: str | None = None):
        if version not in _COLL_TABLES:
            raise ValueError(f"CollNet ABI v{version} has no ctypes table here")
        self.version = version
        self.lib = load(lib_name) if lib_name else load()
        self.tab = _COLL_TABLES[version].in_dll(self.lib, f"ncclCollNetPlugin_v{version}")
        self.name = self.tab.name.decode()

    def _chk(self, what, rc):
        if rc != ncclSuccess:
            raise PluginError(what, rc)

    def init(self):
        self._chk("init", self.tab.init(None))

    def devices(self) -> int:
        n = C.c_int(0)
        self._chk("devices", self.tab.devices(C.byref(n)))
        return n.value

    def get_properties(self, dev: int) -> dict:
        p = _COLL_PROPS[self.version]()
        self._chk("getProperties", self.tab.getProperties(dev, C.byref(p)))
        return {f: (getattr(p, f).decode() if isinstance(getattr(p, f), bytes) else getattr(p, f)) for f, _t in p._fields_
                if not isinstance(getattr(p, f), VProps)}

    def listen(self, dev: int = 0):
        handle = C.create_string_buffer(HANDLE_BYTES[self.version])
        comm = C.c_void_p()
        self._chk("listen", self.tab.listen(dev, handle, C.byref(comm)))
        return bytes(handle.raw), comm

    def connect(self, handles: list, rank: int, listen_comm):
        bufs = [C.create_string_buffer(bytes(h), HANDLE_BYTES[self.version]) for h in handles]
        arr = (C.c_void_p * len(bufs))(*[C.cast(b, C.c_void_p) for b in bufs])
        comm = C.c_void_p()
        self._chk("connect", self.tab.connect(arr, len(bufs), rank, listen_comm, C.byref(comm)))
        return comm

    def reduce_support(self, dtype: int, op: int = ncclSum) -> bool:
        ok = C.c_int(0)
        self._chk("reduceSupport", self.tab.reduceSupport(dtype, op, C.byref(ok)))
        return bool(ok.value)

    def reg_mr(self, comm, ptr: int, nbytes: int, ptr_type: int = NCCL_PTR_CUDA):
        mh = C.c_void_p()
        self._chk("regMr", self.tab.regMr(comm, C.c_void_p(ptr), nbytes, ptr_type, C.byref(mh)))
        return mh

    def dereg_mr(self, comm, mh):
        self._chk("deregMr", self.tab.deregMr(comm, mh))

    def iallreduce(self, comm, send_ptr: int, recv_ptr: int, count: int, dtype: int, send_mh, recv_mh, op: int = ncclSum):
        """Returns the request, or None when the plugin asks to try again later."""
        req = C.c_void_p()
        self._chk("iallreduce", self.tab.iallreduce(comm, C.c_void_p(send_ptr), C.c_void_p(recv_ptr), count, dtype, op, send_mh,
                                                   recv_mh, C.byref(req)))
        return req if req.value else None

    def iflush(self, comm, ptr: int, nbytes: int, mh):
        req = C.c_void_p()
        self._chk("iflush", self.tab.iflush(comm, C.c_void_p(ptr), nbytes, mh, C.byref(req)))
        return req

    def test(self, req):
        done, size = C.c_int(0), C.c_int(0)
        self._chk("test", self.tab.test(req, C.byref(done), C.byref(size)))
        return bool(done.value), size.value

    def close_coll(self, comm):
        self._chk("closeColl", self.tab.closeColl(comm))

    def close_listen(self, comm):
        self._chk("closeListen", self.tab.closeListen(comm))
