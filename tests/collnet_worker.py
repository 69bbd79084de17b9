"""One rank of the CollNet test: drives ncclCollNetPlugin_vN the way NCCL does — listen, connect with everybody's handles,
regMr of a send and a receive buffer, several iallreduce calls in flight (they queue), test until done.  Device memory is
the CPU emulation (BNET_FAKE_CUDA=1); the all-reduce underneath is the two-shot mesh with fused isends.
usage: collnet_worker.py <rank> <world> <dir> <abi version> <count> <f32|bf16> <lib name or -> [host]
host: ordinary host memory (NCCL_PTR_HOST) instead of emulated device memory — works over TCP as well"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

rank, world, d, ver, count, dt = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
lib_name = None if sys.argv[7] == "-" else sys.argv[7]
host_mem = len(sys.argv) > 8 and sys.argv[8] == "host"

from bagua_net_b200.utils.abi import NCCL_PTR_HOST, CollNetPlugin, PluginError, ncclBfloat16, ncclFloat32  # noqa: E402

p = CollNetPlugin(ver, lib_name)
p.init()
ndev = p.devices()
assert ndev >= 1, ndev
props = p.get_properties(0)
lib = p.lib
lib.bnet_fake_cuda_alloc.restype = C.c_void_p
lib.bnet_fake_cuda_alloc.argtypes = [C.c_size_t]
es = 4 if dt == "f32" else 2
nb = max(count * es, 64)
NCALLS = 3
if host_mem:
    _keep = [np.zeros(nb + 64, dtype=np.uint8) for _ in range(NCALLS)] + [np.zeros(NCALLS * nb + 64, dtype=np.uint8)]
    sptr, rptr = [a.ctypes.data for a in _keep[:NCALLS]], _keep[NCALLS].ctypes.data
else:
    sptr = [lib.bnet_fake_cuda_alloc(nb + 64) for _ in range(NCALLS)]
    rptr = lib.bnet_fake_cuda_alloc(NCALLS * nb + 64)
PTR = NCCL_PTR_HOST if host_mem else 2
sraw = [(C.c_char * nb).from_address(x) for x in sptr]
rraw = (C.c_char * (NCALLS * nb)).from_address(rptr)

handle, lcomm = p.listen(0)
with open(os.path.join(d, f"h{rank}.tmp"), "wb") as f:
    f.write(handle)
os.replace(os.path.join(d, f"h{rank}.tmp"), os.path.join(d, f"h{rank}"))
handles, t0 = [], time.time()
for r in range(world):
    path = os.path.join(d, f"h{r}")
    while not os.path.exists(path):
        assert time.time() - t0 < 60
        time.sleep(0.01)
    handles.append(open(path, "rb").read())
comm = p.connect(handles, rank, lcomm)
nccl_dt = ncclFloat32 if dt == "f32" else ncclBfloat16
support = {"sum_this_type": p.reduce_support(nccl_dt), "max": p.reduce_support(nccl_dt, 2), "int32": p.reduce_support(2)}
bad_type_refused = False
try:
    p.reg_mr(comm, sptr[0], nb, 4)                 # NCCL_PTR_DMABUF is not offered
except PluginError:
    bad_type_refused = True
smh = [p.reg_mr(comm, x, nb, PTR) for x in sptr]
rmh = p.reg_mr(comm, rptr, NCALLS * nb, PTR)


def put(raw, vals):
    if dt == "f32":
        np.frombuffer(raw, dtype=np.float32, count=count)[:] = vals
    else:
        np.frombuffer(raw, dtype=np.uint16, count=count)[:] = (vals.astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)


def get(raw, k):
    if dt == "f32":
        return np.frombuffer(raw, dtype=np.float32, count=count, offset=k * nb).copy()
    return (np.frombuffer(raw, dtype=np.uint16, count=count, offset=k * nb).astype(np.uint32) << 16).view(np.float32)


gen = lambda r, k: ((np.arange(count) * 3 + r + 2 * k) % 7 - 3).astype(np.float32)   # noqa: E731  (exact in bf16, any order)
ok = True
for rnd in range(2):
    for k in range(NCALLS):
        put(sraw[k], gen(rank, k + rnd))
    np.frombuffer(rraw, dtype=np.uint8)[:] = 0xEE
    reqs = []
    for k in range(NCALLS):           # all posted before any is tested: the later ones wait in the plugin's queue
        while True:
            q = p.iallreduce(comm, sptr[k], rptr + k * nb, count, nccl_dt, smh[k], rmh)
            if q is not None:
                break
        reqs.append(q)
    sizes = []
    for q in reversed(reqs):          # tested out of order on purpose
        t1 = time.time()
        while True:
            done, size = p.test(q)
            if done:
                sizes.append(size)
                break
            assert time.time() - t1 < 120
    fl = p.iflush(comm, rptr, NCALLS * nb, rmh)
    fdone, _ = p.test(fl)
    for k in range(NCALLS):
        want = sum(gen(r, k + rnd) for r in range(world))
        ok = ok and bool(np.array_equal(get(rraw, k), want))
    ok = ok and fdone and sizes == [count * es] * NCALLS
    open(os.path.join(d, f"done{rnd}_{rank}"), "w").close()     # nobody rewrites its send buffers before all have checked
    for r in range(world):
        while not os.path.exists(os.path.join(d, f"done{rnd}_{r}")):
            time.sleep(0.002)
for h in smh + [rmh]:
    p.dereg_mr(comm, h)
p.close_coll(comm)
p.close_listen(lcomm)
print(json.dumps({"ok": ok, "name": p.name, "ndev": ndev, "ptr_support": props["ptrSupport"], "support": support,
                  "bad_type_refused": bad_type_refused}))
