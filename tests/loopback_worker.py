"""Two-process loopback driver for the ncclNet tables (used by test_loopback.py and
usable by hand):  python tests/loopback_worker.py <role> <rendezvous dir> [options]

role 0 = receiver (listen/accept/irecv), role 1 = sender (connect/isend).
The handle is exchanged through a file like NCCL's bootstrap would do.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bagua_net_b200.utils.abi import NCCL_PTR_CUDA, NCCL_PTR_HOST, NetPlugin, PluginError  # noqa: E402


def pattern(size: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=max(size, 1), dtype=np.uint8)


class HostBuf:
    def __init__(self, n):
        self.arr = np.zeros(max(n, 1), dtype=np.uint8)
        self.addr = self.arr.ctypes.data
        self.type = NCCL_PTR_HOST

    def view(self):
        return self.arr


class FakeCudaBuf:
    """'Device' memory under BNET_FAKE_CUDA=1: a shm segment the library can export."""

    def __init__(self, lib, n):
        lib.bnet_fake_cuda_alloc.restype = ctypes.c_void_p
        lib.bnet_fake_cuda_alloc.argtypes = [ctypes.c_size_t]
        lib.bnet_fake_cuda_free.argtypes = [ctypes.c_void_p]
        self._lib = lib
        self.n = max(n, 16)
        self.addr = lib.bnet_fake_cuda_alloc(self.n)
        assert self.addr
        self.arr = np.ctypeslib.as_array((ctypes.c_uint8 * self.n).from_address(self.addr))
        self.type = NCCL_PTR_CUDA

    def view(self):
        return self.arr

    def close(self):
        if self.addr:
            self.arr = None
            self._lib.bnet_fake_cuda_free(ctypes.c_void_p(self.addr))
            self.addr = 0


class CudaBuf:
    """Real device memory (torch caching allocator -> cudaMalloc -> CUDA IPC export)."""

    def __init__(self, n):
        import torch

        self.t = torch.zeros(max(n, 16), dtype=torch.uint8, device="cuda")
        self.addr = self.t.data_ptr()
        self.type = NCCL_PTR_CUDA

    def view(self):
        return _CudaProxy(self.t)


class _CudaProxy:
    """numpy-ish slice assignment/reads on a CUDA tensor, enough for this driver."""

    def __init__(self, t):
        self.t = t

    def __setitem__(self, sl, val):
        import torch

        self.t[sl] = torch.from_numpy(np.ascontiguousarray(val)).to(self.t.device)
        torch.cuda.synchronize()

    def __getitem__(self, sl):
        import torch

        torch.cuda.synchronize()
        return self.t[sl].cpu().numpy()


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("role", type=int)
    ap.add_argument("dir")
    ap.add_argument("--abi", type=int, default=8)
    ap.add_argument("--sizes", default="0,1,8,4096,524288,1048575,1048577,4194304")
    ap.add_argument("--inflight", type=int, default=8)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--mem", default="host", choices=["host", "fakecuda", "cuda"])
    ap.add_argument("--send-mem", default=None, choices=["host", "fakecuda", "cuda"], help="sender's memory kind (default: --mem)")
    ap.add_argument("--mix", action="store_true", help="sender alternates pinned-host and device buffers message by message")
    ap.add_argument("--die-after", type=int, default=-1, help="sender exits abruptly after N messages")
    ap.add_argument("--expect-error", action="store_true")
    ap.add_argument("--bw", action="store_true", help="print a bandwidth line for the largest size")
    ap.add_argument("--group", type=int, default=1, help="receiver posts its receives in groups of this many (one request each)")
    a = ap.parse_args()

    sizes = [int(s) for s in a.sizes.split(",") if s != ""]
    p = NetPlugin(a.abi)
    p.init()
    assert p.devices() >= 1, "no device"
    hfile = os.path.join(a.dir, "handle.bin")
    result = {"role": a.role, "abi": a.abi, "name": p.name}
    if a.mem == "cuda":
        import torch

        torch.cuda.set_device(int(os.environ.get("BNET_TEST_CUDA_DEV", a.role)) % torch.cuda.device_count())
        torch.zeros(1, device="cuda")
    mem = a.send_mem if (a.role == 1 and a.send_mem) else a.mem
    alloc = (lambda n: FakeCudaBuf(p.lib, n)) if mem == "fakecuda" else (CudaBuf if mem == "cuda" else HostBuf)

    if a.role == 0:
        handle, lcomm = p.listen(0)
        with open(hfile + ".tmp", "wb") as f:
            f.write(handle)
        os.rename(hfile + ".tmp", hfile)
        comm = p.accept(lcomm)
        p.close_listen(lcomm)
    else:
        t0 = time.time()
        while not os.path.exists(hfile):
            assert time.time() - t0 < 30, "no handle"
            time.sleep(0.005)
        with open(hfile, "rb") as f:
            handle = f.read()
        comm = p.connect(handle)
    result["transport"] = p.transport_of(comm)

    nmsg = 0
    try:
        for rnd in range(a.rounds):
            for size in sizes:
                # `inflight` messages of this size posted back to back, then drained
                bufs, reqs, mhs = [], [], []
                extra = 64 if a.role == 0 else 0       # recv buffer larger than the message
                for j in range(a.inflight):
                    if a.mix and a.role == 1:        # LL-style (host) and Simple-style (device) messages on ONE connection
                        b = HostBuf(size + extra) if (j + rnd) % 2 else FakeCudaBuf(p.lib, size + extra)
                    else:
                        b = alloc(size + extra)
                    mh = p.reg_mr(comm, b.addr, size + extra, b.type)
                    if a.role == 1:
                        b.view()[: max(size, 1)] = pattern(size, 1000 * rnd + 7 * j + size % 97)
                    bufs.append(b)
                    mhs.append(mh)
                t0 = time.perf_counter()
                if a.role == 0 and a.group > 1:
                    # grouped receives: `group` buffers under one request, the sender's isends fill them in order
                    groups = []
                    for j0 in range(0, a.inflight, a.group):
                        idx = list(range(j0, min(j0 + a.group, a.inflight)))
                        while True:
                            r = p.irecv_group(comm, [bufs[j].addr for j in idx], [size + extra] * len(idx), [mhs[j] for j in idx],
                                              tags=[5] * len(idx))
                            if r is not None:
                                break
                        groups.append((r, len(idx)))
                        nmsg += len(idx)
                    for r, n in groups:
                        t1 = time.time()
                        while True:
                            done, got = p.test_group(r, n)
                            if done:
                                break
                            if time.time() - t1 > 60:
                                raise TimeoutError("grouped receive did not complete")
                        assert got == [size] * n, f"sizes of a grouped receive: {got} != {size} x {n}"
                    result["grouped"] = result.get("grouped", 0) + len(groups)
                else:
                    for j in range(a.inflight):
                        while True:
                            r = (p.irecv(comm, bufs[j].addr, size + extra, mhs[j]) if a.role == 0
                                 else p.isend(comm, bufs[j].addr, size, mhs[j], tag=5 if a.group > 1 else 0))
                            if r is not None:
                                break
                        reqs.append(r)
                        nmsg += 1
                        if a.role == 1 and a.die_after >= 0 and nmsg >= a.die_after:
                            os._exit(17)        # simulate a crashed peer mid-transfer
                    for j in range(a.inflight):
                        got = p.wait(reqs[j], timeout=60)
                        assert got == size, f"size mismatch: {got} != {size}"
                dt = time.perf_counter() - t0
                if a.role == 0:
                    for j in range(a.inflight):
                        exp = pattern(size, 1000 * rnd + 7 * j + size % 97)
                        if size and not np.array_equal(bufs[j].view()[:size], exp[:size]):
                            bad = int(np.argmax(bufs[j].view()[:size] != exp[:size]))
                            raise AssertionError(f"payload mismatch size={size} msg={j} first bad byte {bad}")
                        if a.mem == "host":
                            assert not bufs[j].view()[size:size + extra].any(), "wrote past the message"
                for j in range(a.inflight):
                    p.dereg_mr(comm, mhs[j])
                for b in bufs:                      # emulated device segments are shm files: give them back
                    if hasattr(b, "close"):
                        b.close()
                if a.bw and size == max(sizes):
                    result["gbps"] = a.inflight * size / dt / 1e9
        result["ok"] = True
    except PluginError as e:
        result["ok"] = False
        result["error"] = str(e)
        result["code"] = e.code
    except TimeoutError as e:
        result["ok"] = False
        result["error"] = f"timeout: {e}"
        result["code"] = -1
    result["messages"] = nmsg
    try:
        from bagua_net_b200.utils import native

        result["exec"] = native.exec_stats()
        result["cma_messages"] = sum(int(float(ln.split()[-1])) for ln in native.metrics_text().splitlines()
                                     if ln.startswith("bnet_cma_messages_total"))
        result["kernel_chunks"] = sum(int(float(ln.split()[-1])) for ln in native.metrics_text().splitlines()
                                      if ln.startswith("bnet_nvl_kernel_chunks_total"))
    except Exception:
        pass
    (p.close_recv if a.role == 0 else p.close_send)(comm)
    print(json.dumps(result), flush=True)
    if a.expect_error:
        return 0 if not result["ok"] and result.get("code", -1) > 0 else 1
    return 0 if result["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
