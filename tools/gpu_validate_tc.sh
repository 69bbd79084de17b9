#!/bin/bash
# First run of the tcgen05 linear / fused GEMM + all-reduce kernel on hardware (csrc/cuda/tc_gemm.cu).
#   gpurun --timeout 600 -- 'bash tools/gpu_validate_tc.sh'            # 1 GPU: numerics + timing + ncu
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_validate_tc.sh 2' # + the row-parallel (cross-rank adds) test
# Everything runs under `timeout`; the kernel's own watchdog (2 s per wait) reports a stuck pipeline as an error code.
N=${1:-1}
OUT=gpurun_out/tc; mkdir -p $OUT
export PYTHONPATH=$PWD BNET_TEST_TC=1
exec > >(tee $OUT/session.log) 2>&1
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name]"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$" $OUT/$name.log | tail -${TAILN:-15} | cut -c1-240; return $rc; }
step build 300 make -j16
step desc 120 python -m pytest tests/test_utils.py -q -k tc
# smallest possible first contact: one tile, one K block
step first 120 python - <<'PY'
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
TAILN=60 step probe 200 python tools/tc_probe.py
step sanitizer 300 compute-sanitizer --tool memcheck python -c "
import torch
from bagua_net_b200.ops import tc_linear as t
print(t.self_check(verbose=True))"
TAILN=8 step gpu_test 300 python -m pytest tests/test_gpu.py -q -x -k "tcgen05_linear"
[ "$N" -ge 2 ] && TAILN=8 step gpu_test_2 300 python -m pytest tests/test_gpu.py -q -x -k "tcgen05_row_parallel"
TAILN=30 step timing 200 python tools/tc_linear_bench.py
step ncu 400 ncu --set full --clock-control none --import-source on -k "regex:tc_linear_kernel" -c 4 -o $OUT/tc_linear python tools/tc_linear_bench.py --iters 1 --warmup 0 --shapes 4096x4096x4096,32x4096x25088
[ -f tools/summarize_ncu.sh ] && step ncu_summary 120 bash tools/summarize_ncu.sh $OUT/tc_linear.ncu-rep $OUT/tc_linear
echo "== done"
