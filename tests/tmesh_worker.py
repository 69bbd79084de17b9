"""One rank of the one-shot (full-mesh) all-reduce test: every rank sends its input to every peer once, the sender's kernel —
here its CPU emulation over the NVL transport with emulated device memory — accumulates into the peer's output.
usage: tmesh_worker.py <rank> <world> <dir> <count> <in f32|bf16> <out f32|bf16> <piece_bytes> <inflight> [rounds] [algo] [inplace]
algo: one-shot | two-shot (reduce-scatter into the slice owners, all-gather by copy); inplace = 1: input and output are ONE buffer"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

rank, world, d = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
count, idt, odt = int(sys.argv[4]), sys.argv[5], sys.argv[6]
piece, inflight = int(sys.argv[7]), int(sys.argv[8])
rounds = int(sys.argv[9]) if len(sys.argv) > 9 else 2
algo = sys.argv[10] if len(sys.argv) > 10 else "one-shot"
inplace = len(sys.argv) > 11 and sys.argv[11] == "1"
host_mem = len(sys.argv) > 12 and sys.argv[12] == "host"      # ordinary host memory (any transport) instead of emulated device memory

from bagua_net_b200.parallel.transport_ring import MeshCore  # noqa: E402
from bagua_net_b200.utils.native import load  # noqa: E402

lib = load()
lib.bnet_fake_cuda_alloc.restype = C.c_void_p
lib.bnet_fake_cuda_alloc.argtypes = [C.c_size_t]
ies, oes = (4 if idt == "f32" else 2), (4 if odt == "f32" else 2)
ib, ob = max(count * ies, 64), max(count * oes, 64)
if host_mem:
    _keep = [np.zeros(ib + 64, dtype=np.uint8), None]
    _keep[1] = _keep[0] if inplace else np.zeros(ob + 64, dtype=np.uint8)
    iptr, optr = _keep[0].ctypes.data, _keep[1].ctypes.data
else:
    iptr = lib.bnet_fake_cuda_alloc(ib + 64)
    optr = iptr if inplace else lib.bnet_fake_cuda_alloc(ob + 64)
assert iptr and optr and (not inplace or idt == odt)
iraw, oraw = (C.c_char * ib).from_address(iptr), (C.c_char * ob).from_address(optr)

core = MeshCore(rank, world)
with open(os.path.join(d, f"h{rank}.tmp"), "wb") as f:
    f.write(core.handle)
os.replace(os.path.join(d, f"h{rank}.tmp"), os.path.join(d, f"h{rank}"))
handles = []
t0 = time.time()
for r in range(world):
    p = os.path.join(d, f"h{r}")
    while not os.path.exists(p):
        assert time.time() - t0 < 60
        time.sleep(0.01)
    handles.append(open(p, "rb").read())
core.connect(handles)
core.register(iptr, ib, optr, ob, host_memory=host_mem)


def put(raw, dt, vals):
    if dt == "f32":
        np.frombuffer(raw, dtype=np.float32, count=count)[:] = vals
    else:
        np.frombuffer(raw, dtype=np.uint16, count=count)[:] = (vals.astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)


def get(raw, dt):
    if dt == "f32":
        return np.frombuffer(raw, dtype=np.float32, count=count).copy()
    return (np.frombuffer(raw, dtype=np.uint16, count=count).astype(np.uint32) << 16).view(np.float32)


ok = True
for rnd in range(rounds):
    gen = lambda r: ((np.arange(count) * 5 + r + rnd) % 9 - 4).astype(np.float32)   # noqa: E731  (small integers: exact in bf16, any order)
    if not inplace:
        put(oraw, odt, np.full(count, 99.0, dtype=np.float32))        # stale output: every element must be overwritten
    put(iraw, idt, gen(rank))
    want = sum(gen(r) for r in range(world))
    core.all_reduce(iptr, optr, count, 0 if idt == "f32" else 1, 0 if odt == "f32" else 1, piece, inflight, algo=algo)
    ok = ok and bool(np.array_equal(get(oraw, odt), want)) and (inplace or bool(np.array_equal(get(iraw, idt), gen(rank))))
    # nobody may start the next round (and overwrite its input) before every rank has checked this one
    open(os.path.join(d, f"done{rnd}_{rank}"), "w").close()
    for r in range(world):
        while not os.path.exists(os.path.join(d, f"done{rnd}_{r}")):
            time.sleep(0.002)
print(json.dumps({"ok": ok, "transport": core.transport, "stats": core.stats()}))
core.close()
