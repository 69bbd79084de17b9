#!/bin/bash
# Round-2 session 1 (1 GPU): tcgen05 first contact, what changed since the last hardware run, bench, ncu.
TAG=${1:-r2s1}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD BNET_TEST_TC=1
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG $(date -u)"; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv | head -3
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|OMP_NUM" $OUT/$name.log | tail -${TAILN:-12} | cut -c1-${CUT:-260}; return $rc; }
step build 300 make -j16
step tc_first 120 python - <<'PY'
import torch
from bagua_net_b200.ops import tc_linear as t
print("supported", t.supported())
x = torch.randn(32, 64, device="cuda").bfloat16(); w = torch.randn(128, 64, device="cuda").bfloat16()
y = t.linear(x, w); torch.cuda.synchronize()
print("err flag", t.last_error(), "max abs err", (y.float() - x.float() @ w.float().t()).abs().max().item())
x = torch.randn(128, 64, device="cuda").bfloat16()
y = t.linear(x, w); torch.cuda.synchronize()
print("no-swap: err flag", t.last_error(), "max abs err", (y.float() - x.float() @ w.float().t()).abs().max().item())
print("self_check", t.self_check(verbose=True))
PY
TAILN=80 step tc_probe 200 python tools/tc_probe.py
TAILN=8 step tc_gpu_test 300 python -m pytest tests/test_gpu.py -q -x -k "tcgen05_linear"
TAILN=30 step tc_timing 200 python tools/tc_linear_bench.py
CUT=2000 TAILN=2 step bench1 400 python bench.py --gpus 1 --steps 30 --warmup 5
TAILN=40 step nn_kernels 120 python tools/nn_kernel_bench.py
TAILN=8 step gpu_tests 700 python -m pytest tests -q -m gpu -x
step ncu_tc 300 ncu --set full --clock-control none --import-source on -k "regex:tc_linear_kernel" -c 4 -o $OUT/tc_linear python tools/tc_linear_bench.py --iters 1 --warmup 0 --shapes 4096x4096x4096,32x4096x25088
step ncu_nn 300 ncu --set full --clock-control none --import-source on -k "regex:relu_bwd_bias_grad|pool_relu_bwd" -c 6 -o $OUT/nn_bwd python tools/nn_kernel_bench.py --iters 1 --warmup 0
echo "== done $(date -u)"
