#!/bin/bash
# First GPU session after a stretch of CPU-only work (1 GPU is enough): validate, then measure, what changed
# since the last run on hardware — the re-batched backward layer kernels, the executor's batched reduce/cast
# loops, the copy-engine mode, the prefetching end-to-end loop — and bring back ncu evidence for the new kernels.
#   gpurun --timeout 900 -- 'bash tools/gpu_session_next.sh s10'
TAG=${1:-next}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG $(date -u)"; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv | head -3
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|OMP_NUM" $OUT/$name.log | tail -${TAILN:-12} | cut -c1-${CUT:-260}; return $rc; }
step build 300 make -j16
TAILN=6 step gpu_tests 600 python -m pytest tests -q -m gpu -x
TAILN=40 step nn_kernels 120 python tools/nn_kernel_bench.py
CUT=2000 TAILN=2 step bench1 300 python bench.py --gpus 1 --steps 30 --warmup 5
CUT=2000 TAILN=2 step bench1_noprefetch 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-prefetch --no-extra
CUT=2000 TAILN=2 step bench1_resnet50 300 python bench.py --gpus 1 --steps 30 --warmup 5 --model resnet50 --no-extra
CUT=2000 TAILN=2 step bench1_resnet50_eager 300 python bench.py --gpus 1 --steps 30 --warmup 5 --model resnet50 --no-extra --no-fused
TAILN=30 step step_profile 200 python tools/step_profile.py --fused
step ncu_nn 400 ncu --set full --clock-control none --import-source on -k "regex:relu_bwd_bias_grad|pool_relu_bwd" -c 6 -o $OUT/nn_bwd python tools/nn_kernel_bench.py --iters 1 --warmup 0
step ncu_summary 120 bash tools/summarize_ncu.sh $OUT/nn_bwd.ncu-rep $OUT/nn_bwd
echo "== done $(date -u)"
