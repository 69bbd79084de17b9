"""One rank of the transport-ring all-reduce test (tests/test_loopback.py::test_transport_ring_*): the ring rides the
plugin's NVL transport with emulated device memory (BNET_FAKE_CUDA=1), so the fused isend-reduce protocol — op codes
in the request, accumulate into the receiver's registered buffer, FIFO matching, piece pipelining — runs without a GPU.
usage: tring_worker.py <rank> <world> <dir> <count> <dtype f32|bf16> <piece_bytes> <inflight> [rounds]
       dtype "c:<wire>:<fused 0|1>" runs the compressed all-reduce (fp32 data, bf16 / e4m3 / e5m2 on the wire)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

rank, world, d = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
count, dtype = int(sys.argv[4]), sys.argv[5]
piece, inflight = int(sys.argv[6]), int(sys.argv[7])
rounds = int(sys.argv[8]) if len(sys.argv) > 8 else 2

from bagua_net_b200.parallel.transport_ring import RingCore  # noqa: E402
from bagua_net_b200.utils.native import load  # noqa: E402

lib = load()
lib.bnet_fake_cuda_alloc.restype = C.c_void_p
lib.bnet_fake_cuda_alloc.argtypes = [C.c_size_t]
compressed = dtype.startswith("c:")
wire, fused = (dtype.split(":")[1], dtype.split(":")[2] == "1") if compressed else (None, False)
es = 2 if dtype == "bf16" else 4
nbytes = max(count * es, 64)
ptr = lib.bnet_fake_cuda_alloc(nbytes + 64)
assert ptr
raw = (C.c_char * (nbytes + 64)).from_address(ptr)

core = RingCore(rank, world)
with open(os.path.join(d, f"h{rank}.tmp"), "wb") as f:
    f.write(core.handle)
os.replace(os.path.join(d, f"h{rank}.tmp"), os.path.join(d, f"h{rank}"))
nxt = os.path.join(d, f"h{(rank + 1) % world}")
t0 = time.time()
while not os.path.exists(nxt):
    assert time.time() - t0 < 60
    time.sleep(0.01)
core.connect(open(nxt, "rb").read())
if not compressed or fused:
    core.register(ptr, nbytes)
if compressed:
    wbytes = max(count * (2 if wire == "bf16" else 1), 64)
    wptr = lib.bnet_fake_cuda_alloc(wbytes + 64)
    assert wptr
    core.register_wire(wptr, wbytes, device_memory=os.environ.get("TRING_WIRE_HOST") != "1")

ok = True
for rnd in range(rounds):
    # small integers: exact in bf16 and in any summation order
    if compressed:
        # values {-1, 0, 1, 2} x 0.25: every partial sum of up to 8 ranks is k/4 with |k| <= 16 — exact in bf16 and, scaled by
        # 4, an integer <= 16 that e4m3 (3 mantissa bits) and e5m2 (2 bits: {0..4, 6, 8, 12, 16} only) may have to round
        a = np.frombuffer(raw, dtype=np.float32, count=count)
        gen = lambda r: (((np.arange(count) * 7 + r * 3 + rnd) % 4) - 1).astype(np.float32) * 0.25   # noqa: E731
        a[:] = gen(rank)
        want = sum(gen(r) for r in range(world)).astype(np.float32)
    elif dtype == "f32":
        a = np.frombuffer(raw, dtype=np.float32, count=count)
        a[:] = ((np.arange(count) * 7 + rank * 3 + rnd) % 17 - 8).astype(np.float32)
        want = sum(((np.arange(count) * 7 + r * 3 + rnd) % 17 - 8) for r in range(world)).astype(np.float32)
    else:
        a16 = np.frombuffer(raw, dtype=np.uint16, count=count)
        vals = ((np.arange(count) * 5 + rank + rnd) % 9 - 4).astype(np.float32)
        a16[:] = (vals.view(np.uint32) >> 16).astype(np.uint16)
        want = sum(((np.arange(count) * 5 + r + rnd) % 9 - 4) for r in range(world)).astype(np.float32)
    # crude barrier through the file system: nobody starts reducing into a buffer that is still being filled
    open(os.path.join(d, f"ready{rnd}_{rank}"), "w").close()
    for r in range(world):
        while not os.path.exists(os.path.join(d, f"ready{rnd}_{r}")):
            time.sleep(0.002)
    if compressed:
        core.all_reduce_compressed(ptr, count, wire, 4.0, fused, piece, inflight)
    else:
        core.all_reduce(ptr, count, 0 if dtype == "f32" else 1, piece, inflight)
    if dtype == "f32" or compressed:
        got = np.frombuffer(raw, dtype=np.float32, count=count).copy()
    else:
        got = (np.frombuffer(raw, dtype=np.uint16, count=count).astype(np.uint32) << 16).view(np.float32)
    if compressed and wire == "e5m2":
        # 2 mantissa bits: a partial sum may be rounded at every hop; the result stays within the format's spacing of the
        # true sum, and — what the protocol promises — it is the SAME on every rank (checked by the parent through `digest`)
        ok = ok and bool(np.all(np.abs(got - want) <= 1.5))
    else:
        ok = ok and bool(np.array_equal(got, want))
    digest = int(np.frombuffer(got.tobytes(), dtype=np.uint32).astype(np.uint64).sum() % (1 << 61))
    open(os.path.join(d, f"done{rnd}_{rank}"), "w").close()
    for r in range(world):
        while not os.path.exists(os.path.join(d, f"done{rnd}_{r}")):
            time.sleep(0.002)
st = core.stats()
print(json.dumps({"ok": ok, "transport": core.transport, "stats": st, "digest": digest}))
core.close()
