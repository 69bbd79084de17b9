#!/usr/bin/env python
"""Does NCCL accept the plugin's CollNet table, and what does an all-reduce cost through it?

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 4 bench/nccl_collnet_probe.py [--json out.json]

Every rank re-executes itself with the plugin's environment (the loader path must be set before the process starts),
BNET_COLLNET=1 (the CollNet table reports its devices), NCCL_COLLNET_ENABLE=1 and its own NCCL_HOSTID — one rank per
"node", which is the topology NCCL's CollNet algorithms are built for (and sends every byte through the plugin).  Then:
torch.distributed all-reduces of fp32 sums (the table offers sum of fp32 / bf16), checked against the closed form and timed
on the device (max over ranks).  NCCL_DEBUG=INFO goes to the log; the lines that mention CollNet are returned with the
numbers, together with the plugin's own count of all-reduces it executed (`collnet_allreduces`: 0 = NCCL loaded the plugin
but kept its ring / tree over the net path).  `bench.py` runs this as a bounded child job at N > 1 (`extra.nccl_collnet`)."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--sizes-kib", default="64,1024,16384,65536")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("BNET_COLLNET_PROBE_REEXEC") != "1":
        from bagua_net_b200.utils.env import nccl_plugin_env

        env = dict(os.environ)
        env.update(nccl_plugin_env(force_net=True, eager_modules=False))
        env.update(BNET_COLLNET="1", NCCL_COLLNET_ENABLE="1", NCCL_HOSTID=f"bnet-probe-{rank}", NCCL_DEBUG="INFO",
                   NCCL_DEBUG_SUBSYS="INIT,NET,COLL,TUNING", BNET_COLLNET_PROBE_REEXEC="1", BNET_WATCHDOG_MS="8000")
        env.pop("NCCL_ALGO", None)
        os.execve(sys.executable, [sys.executable] + sys.argv, env)

    import torch
    import torch.distributed as dist

    from bagua_net_b200.parallel import init_process_group_from_env
    from bagua_net_b200.utils.native import load

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # every kernel used below is launched once before a collective can be in flight (CUDA's lazy module loading and a
    # transport that launches from NCCL's proxy thread do not mix: bagua_net_b200/utils/env.py)
    w = torch.ones(1 << 16, device=dev)
    (w * 2 + 1).sum().item()
    torch.full((4,), 3.0, device=dev).eq(3.0).all().item()
    torch.cuda.synchronize()
    init_process_group_from_env("nccl")
    out = {"world": world, "busbw_gbs_fp32": {}, "time_us": {}, "exact": True}
    bus = 2 * (world - 1) / world
    for kib in [int(s) for s in args.sizes_kib.split(",") if s]:
        nbytes = kib << 10
        t = torch.full((nbytes // 4,), float(rank + 1), device=dev)
        torch.cuda.synchronize()
        dist.all_reduce(t)
        torch.cuda.synchronize()
        out["exact"] = out["exact"] and bool(t.eq(float(world * (world + 1) // 2)).all().item())
        t.fill_(1.0)
        iters = 10 if nbytes <= (16 << 20) else 5
        for _ in range(2):
            dist.all_reduce(t)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            dist.all_reduce(t)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize()
        us = float(ms.item()) * 1e3
        out["time_us"][str(nbytes)] = round(us, 1)
        out["busbw_gbs_fp32"][str(nbytes)] = round(nbytes * bus / (us * 1e-6) / 1e9, 2)
        del t
    # the plugin's own evidence: how many all-reduces did the CollNet table execute in this process?
    try:
        lib = load()
        lib.bnet_collnet_allreduces.restype = __import__("ctypes").c_ulonglong
        mine = int(lib.bnet_collnet_allreduces())
    except Exception:   # noqa: BLE001
        mine = -1
    cnt = torch.tensor([mine], device=dev, dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    out["collnet_allreduces_min_over_ranks"] = int(cnt.item())
    if rank == 0:
        line = json.dumps(out)
        if args.json:
            with open(args.json + ".tmp", "w") as f:
                f.write(line)
            os.replace(args.json + ".tmp", args.json)
        print(line, flush=True)
    dist.barrier()
    import threading

    threading.Timer(8.0, lambda: os._exit(0)).start()     # never let teardown outlive the measurement
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    sys.exit(main())
