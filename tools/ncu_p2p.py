"""One process, two GPUs: the transport's executor kernels moving a buffer of GPU 0 into GPU 1 over NVLink — the form ncu
can profile (one-sided: no peer process has to take part in a kernel replay).

    ncu --set full --clock-control none --import-source on -k regex:bnet_nvl -c 6 -o gpurun_out/x/nvl python tools/ncu_p2p.py
    python tools/ncu_p2p.py --time          # CUDA-event bandwidth table instead (no profiler)
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bagua_net_b200.ops import P2PExecutor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbytes", type=int, default=256)
    ap.add_argument("--time", action="store_true")
    args = ap.parse_args()
    assert torch.cuda.device_count() >= 2, "needs two GPUs in one process"
    torch.cuda.set_device(0)
    n = args.mbytes << 20
    src = torch.randint(0, 255, (n,), device="cuda:0", dtype=torch.uint8)
    dst = torch.zeros(n, device="cuda:1", dtype=torch.uint8)
    dst[:16].copy_(src[:16])                      # (torch enables peer access between the two devices here)
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    ex = P2PExecutor(0)
    f32 = torch.randn(n // 8, device="cuda:0")
    acc = torch.zeros(n // 8, device="cuda:1")
    bf = torch.randn(n // 8, device="cuda:0").bfloat16()
    acc2 = torch.zeros(n // 8, device="cuda:1")
    jobs = [("msg", "copy", src, dst), ("persistent", "copy", src, dst), ("msg", "red_add_f32", f32, acc),
            ("msg", "acc_bf16_to_f32", bf, acc2)]
    for mode, op, s, d in jobs:
        ex.run_mode(mode, op, s, d)
        if not args.time:
            continue
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            ex.run_mode(mode, op, s, d)
        dt = (time.perf_counter() - t0) / reps
        nb = s.numel() * s.element_size()
        print(f"{mode:10s} {op:16s} {nb >> 20:5d} MiB  {dt * 1e6:9.1f} us  {nb / dt / 1e9:7.1f} GB/s (source bytes, host-timed incl. launch)")
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    ok = torch.equal(dst.cpu(), src.cpu())
    print("copy correct:", ok)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
