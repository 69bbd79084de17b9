"""N-rank ring all-reduce over the ncclNet table, driven the way NCCL's proxy drives a net plugin:
one listen/connect/accept per (channel, neighbour), registered buffers, non-blocking isend/irecv with
several requests in flight per connection, one thread polling `test` over every connection.

    python tests/ring_worker.py <rank> <world> <rendezvous dir> [--channels C] [--mem host|fakecuda] ...

Every rank talks to two different peers (next and previous), which the two-process loopback driver
(tests/loopback_worker.py) cannot exercise; this is the CPU stand-in for the reference's manual multi-host
`all_reduce_perf` procedure (reference README.md:14-46).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bagua_net_b200.utils.abi import NetPlugin  # noqa: E402
from loopback_worker import FakeCudaBuf, HostBuf  # noqa: E402


def wait_file(path: str, timeout: float = 60.0) -> bytes:
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise TimeoutError(path)
        time.sleep(0.002)
    with open(path, "rb") as f:
        return f.read()


def publish(path: str, data: bytes) -> None:
    with open(path + ".tmp", "wb") as f:
        f.write(data)
    os.rename(path + ".tmp", path)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("rank", type=int)
    ap.add_argument("world", type=int)
    ap.add_argument("dir")
    ap.add_argument("--abi", type=int, default=8)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--mem", default="host", choices=["host", "fakecuda"])
    ap.add_argument("--counts", default="1,64,1000,65536,300001", help="int32 elements per all-reduce")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--slices", type=int, default=4, help="messages in flight per connection and step")
    ap.add_argument("--gap-ms", type=float, default=3.0, help="idle time between all-reduces")
    a = ap.parse_args()
    rank, n, C = a.rank, a.world, a.channels
    nxt, prv = (rank + 1) % n, (rank - 1) % n

    p = NetPlugin(a.abi)
    p.init()
    alloc = (lambda k: FakeCudaBuf(p.lib, k)) if a.mem == "fakecuda" else HostBuf

    # ---- connection setup: channel c of rank r listens for rank r-1, connects to rank r+1
    listens = []
    for c in range(C):
        handle, lcomm = p.listen(0)
        publish(os.path.join(a.dir, f"h-{rank}-{c}"), handle)
        listens.append(lcomm)
    sends = [p.connect(wait_file(os.path.join(a.dir, f"h-{nxt}-{c}"))) for c in range(C)]
    recvs = [p.accept(listens[c]) for c in range(C)]
    for lc in listens:
        p.close_listen(lc)
    transports = sorted({p.transport_of(x) for x in sends + recvs})

    counts = [int(x) for x in a.counts.split(",") if x]
    maxc = max(counts)
    # per-rank chunking pads to a multiple of world * channels * slices elements
    unit = n * C * a.slices
    cap = (maxc + unit - 1) // unit * unit
    work = alloc(cap * 4)                                   # the all-reduce buffer (registered on every comm)
    stage = [alloc(cap // n // C * 4 + 64) for _ in range(C)]   # landing zone per channel
    wv = work.view()[: cap * 4].view(np.int32)
    sv = [s.view()[: cap // n // C * 4 + 64].view(np.int32) for s in stage]
    smh = [p.reg_mr(sends[c], work.addr, cap * 4, work.type) for c in range(C)]
    rmh_stage = [p.reg_mr(recvs[c], stage[c].addr, cap // n // C * 4 + 64, stage[c].type) for c in range(C)]
    rmh_work = [p.reg_mr(recvs[c], work.addr, cap * 4, work.type) for c in range(C)]

    def exchange(send_chunk: int, recv_chunk: int, per: int, into_stage: bool):
        """One ring step: chunk `send_chunk` goes to next, chunk `recv_chunk` arrives from prev; each chunk is cut
        into C channel pieces of `a.slices` messages.  Everything is posted non-blocking and polled round-robin."""
        piece = per // C
        sl = piece // a.slices
        pend = []
        for c in range(C):
            for s in range(a.slices):
                off_r = (recv_chunk * per + c * piece + s * sl) * 4
                if into_stage:
                    pend.append(["r", c, stage[c].addr + s * sl * 4, sl * 4 + 16, rmh_stage[c], None, sl * 4])
                else:
                    pend.append(["r", c, work.addr + off_r, sl * 4, rmh_work[c], None, sl * 4])
                off_s = (send_chunk * per + c * piece + s * sl) * 4
                pend.append(["s", c, work.addr + off_s, sl * 4, smh[c], None, sl * 4])
        t0 = time.time()
        done = 0
        while done < len(pend):
            blocked = set()          # FIFO matching: never post message s+1 of a connection before message s
            for e in pend:
                if e[5] is None:
                    if (e[0], e[1]) in blocked:
                        continue
                    e[5] = (p.irecv(recvs[e[1]], e[2], e[3], e[4]) if e[0] == "r" else p.isend(sends[e[1]], e[2], e[3], e[4]))
                    if e[5] is None:
                        blocked.add((e[0], e[1]))
                        continue
                if e[5] == "done":
                    continue
                fin, got = p.test(e[5])
                if fin:
                    assert got == e[6], f"rank {rank}: {e[0]} got {got} bytes, expected {e[6]}"
                    e[5] = "done"
                    done += 1
            if time.time() - t0 > 60:
                raise TimeoutError(f"rank {rank}: ring step stuck ({done}/{len(pend)} requests done)")

    nred = 0
    for it in range(a.iters):
        for cnt in counts:
            padded = (cnt + unit - 1) // unit * unit
            per = padded // n
            wv[:padded] = 0
            wv[:cnt] = (np.arange(cnt, dtype=np.int64) * (rank + 1) + it).astype(np.int32)
            # reduce-scatter
            for step in range(n - 1):
                sc, rc = (rank - step) % n, (rank - step - 1) % n
                exchange(sc, rc, per, into_stage=True)
                piece = per // C
                for c in range(C):
                    lo = rc * per + c * piece
                    wv[lo:lo + piece] += sv[c][:piece]
            # all-gather
            for step in range(n - 1):
                sc, rc = (rank + 1 - step) % n, (rank - step) % n
                exchange(sc, rc, per, into_stage=False)
            exp = (np.arange(cnt, dtype=np.int64) * (n * (n + 1) // 2) + it * n).astype(np.int32)
            if not np.array_equal(wv[:cnt], exp):
                bad = int(np.argmax(wv[:cnt] != exp))
                raise AssertionError(f"rank {rank}: all-reduce mismatch count={cnt} iter={it} first bad element {bad}")
            nred += 1
            time.sleep(a.gap_ms / 1e3)

    for c in range(C):
        p.dereg_mr(sends[c], smh[c])
        p.dereg_mr(recvs[c], rmh_stage[c])
        p.dereg_mr(recvs[c], rmh_work[c])
    for c in range(C):
        p.close_send(sends[c])
        p.close_recv(recvs[c])
    print(json.dumps({"rank": rank, "ok": True, "allreduces": nred, "transports": transports}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
