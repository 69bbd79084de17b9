#!/usr/bin/env python
"""Where does a training step go?  torch.profiler (CUPTI) over a few steady-state steps of the
flagship engine: per-kernel device time, share of the step, and GPU-busy fraction (launch gaps).
Single GPU:  python tools/step_profile.py [--comm bnet|torch] [--batch 32] [--out profiles/step_profile.txt]"""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bagua_net_b200.models import build_model  # noqa: E402
from bagua_net_b200.parallel import BnetDDP  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--comm", default="bnet")
    ap.add_argument("--model", default="vgg16")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--fused", action="store_true", help="fused conv blocks (bnet arm default in bench.py)")
    ap.add_argument("--json", default="", help="also write the top kernels as one JSON object (bench.py: extra.step_profile)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    torch.backends.cudnn.benchmark = True
    model = build_model(a.model, **({"fused": True} if a.fused else {})).cuda().to(torch.bfloat16).to(memory_format=torch.channels_last)
    x = torch.randn(a.batch, 3, 224, 224, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (a.batch,), device="cuda")
    if a.comm == "bnet":
        eng = BnetDDP(model, lr=0.01, momentum=0.9, weight_decay=1e-4)
        step = lambda: eng.train_step(x, y)  # noqa: E731
    else:
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(model(x).float(), y)
            loss.backward()
            opt.step()
            return loss
    for _ in range(8):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    wall_ms = e0.elapsed_time(e1) / a.steps
    agg = collections.defaultdict(lambda: [0, 0.0])
    busy = 0.0
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time_total > 0:
            name = ev.name[:90]
            agg[name][0] += 1
            agg[name][1] += ev.device_time_total
            busy += ev.device_time_total
    busy_ms = busy / 1e3 / a.steps
    lines = [f"# {a.model} batch {a.batch} comm={a.comm} fused={a.fused}: step {wall_ms:.3f} ms (CUDA events), sum of kernel time {busy_ms:.3f} ms "
             f"per step ({100 * busy_ms / wall_ms:.1f}% of the step; >100% means kernels overlap on several streams)"]
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        lines.append(f"{100 * t / busy:5.1f}%  {t / 1e3 / a.steps:8.3f} ms/step  n/step={n / a.steps:6.1f}  {name}")
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        open(a.out, "w").write(txt + "\n")
    if a.json:
        import json

        top = [{"kernel": name[:64], "ms_per_step": round(t / 1e3 / a.steps, 3), "launches_per_step": round(n / a.steps, 1),
                "share": round(t / busy, 3)} for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]]
        ours = sum(t for name, (n, t) in agg.items() if "bnet" in name or "tc_linear_kernel" in name) / busy if busy else 0.0
        obj = {"note": "torch.profiler (CUPTI) over eager steps, no CUDA graph: explains the step, is not a bench value",
               "model": a.model, "batch": a.batch, "fused": bool(a.fused), "step_ms_under_profiler": round(wall_ms, 3),
               "kernel_ms_per_step": round(busy_ms, 3), "share_of_kernel_time_in_our_kernels": round(ours, 3), "top_kernels": top}
        with open(a.json + ".tmp", "w") as f:
            f.write(json.dumps(obj))
        os.replace(a.json + ".tmp", a.json)


if __name__ == "__main__":
    main()
