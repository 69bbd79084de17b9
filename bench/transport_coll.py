#!/usr/bin/env python
"""The all-reduces that ride the TRANSPORT (the plugin's own connections, reduction fused into the isends; no NCCL on the
data path), verified and timed: ring, two-shot mesh (in place), one-shot mesh (bf16 -> fp32).

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 4 bench/transport_coll.py [--json out.json]

Per size the all-reduce runs on integer-valued data (the sum must be exact on every rank), then `iters` timed calls; a call
is host-driven like NCCL's proxy (it returns when every request of this rank has completed), so the time is taken on the
host around the calls, between barriers, and the MAX over ranks is reported.  busbw = size x 2(n-1)/n / time.
`bench.py` runs this as a bounded child job at N > 1 (`extra.transport_allreduce`)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--sizes-mib", default="1,16,64")
    ap.add_argument("--iters", type=int, default=8)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from bagua_net_b200.parallel import init_process_group_from_env
    from bagua_net_b200.parallel.transport_ring import TransportMesh, TransportRing

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world < 2:
        print(json.dumps({"status": "needs at least two ranks"}))
        return 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    init_process_group_from_env("nccl")
    sizes = [int(float(s) * (1 << 20)) for s in args.sizes_mib.split(",") if s]
    nmax = max(sizes) // 4

    def timed(fn, iters):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        t = torch.tensor([(time.perf_counter() - t0) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def pattern(n, r, salt):
        return ((torch.arange(n, device=dev) * 5 + r + salt) % 9 - 4).float()

    def checkpoint(stage):
        # what has been measured so far survives a later stage that hangs or faults (the parent reads the file whatever the
        # child's end was); `exact` is rank 0's own view until the final all-reduce over the ranks
        if rank == 0 and args.json:
            with open(args.json + ".tmp", "w") as f:
                f.write(json.dumps(dict(out, partial=f"stopped after: {stage}")))
            os.replace(args.json + ".tmp", args.json)

    bus = 2 * (world - 1) / world
    out = {"world": world, "ring_fp32": {}, "two_shot_fp32_in_place": {}, "exact": True}
    t_start = time.time()
    ring = TransportRing()
    out["transport"] = ring.transport
    g = ring.buffer(nmax, torch.float32)
    for nbytes in sizes:
        n = nbytes // 4
        g[:n].copy_(pattern(n, rank, 1))
        torch.cuda.synchronize()
        dist.barrier()
        ring.all_reduce(g[:n])
        want = sum(pattern(n, r, 1) for r in range(world))
        out["exact"] = out["exact"] and bool(torch.equal(g[:n], want))
        dist.barrier()
        t = timed(lambda: ring.all_reduce(g[:n]), args.iters)
        out["ring_fp32"][str(nbytes)] = {"us": round(t * 1e6, 1), "busbw_gbs": round(nbytes * bus / t / 1e9, 1)}
    ring.close()
    del g
    checkpoint("ring")
    mesh = TransportMesh()
    h = mesh.buffer(nmax, torch.float32)
    for nbytes in sizes:
        n = nbytes // 4
        h[:n].copy_(pattern(n, rank, 2))
        torch.cuda.synchronize()
        dist.barrier()
        mesh.all_reduce(h[:n], h[:n], algo="two-shot")
        want = sum(pattern(n, r, 2) for r in range(world))
        out["exact"] = out["exact"] and bool(torch.equal(h[:n], want))
        dist.barrier()
        t = timed(lambda: mesh.all_reduce(h[:n], h[:n], algo="two-shot"), args.iters)
        out["two_shot_fp32_in_place"][str(nbytes)] = {"us": round(t * 1e6, 1), "busbw_gbs": round(nbytes * bus / t / 1e9, 1)}
    checkpoint("ring, two-shot mesh")
    # latency shape: one network step, 256 Ki bf16 gradients accumulated in fp32
    x, y = mesh.buffers(1 << 18, torch.bfloat16, torch.float32)
    x.copy_(pattern(1 << 18, rank, 3).to(torch.bfloat16))
    torch.cuda.synchronize()
    dist.barrier()
    mesh.all_reduce(x, y)
    out["exact"] = out["exact"] and bool(torch.equal(y, sum(pattern(1 << 18, r, 3) for r in range(world))))
    dist.barrier()
    t = timed(lambda: mesh.all_reduce(x, y), 20)
    out["one_shot_bf16_to_fp32_512KiB_us"] = round(t * 1e6, 1)
    ok = torch.tensor([1 if out["exact"] else 0], device=dev, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["exact"] = bool(int(ok.item()))
    out["wall_s"] = round(time.time() - t_start, 1)
    out["note"] = ("host-timed around blocking calls (one polling thread per rank drives isend_op / irecv / test), max over ranks; "
                   "busbw = bytes x 2(n-1)/n / time")
    mesh.close()
    if rank == 0:
        line = json.dumps(out)
        if args.json:
            with open(args.json + ".tmp", "w") as f:
                f.write(line)
            os.replace(args.json + ".tmp", args.json)
        print(line, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
