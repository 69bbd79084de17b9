#!/bin/bash
# One GPU-box session: run everything worth measuring, keep every log under gpurun_out/<tag>/.
# usage: tools/gpu_session.sh <tag> [ngpus]
TAG=${1:-s1}; NG=${2:-1}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
nvidia-smi -L; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
nvidia-smi topo -m 2>&1 | head -20
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; tail -${TAILN:-15} $OUT/$name.log; return $rc; }
step build 600 make -j16
step gpu_tests_1 900 python -m pytest tests/test_gpu.py -x -q -m "gpu and not multigpu" -p no:cacheprovider
step bench1 600 python bench.py --gpus 1 --steps 10 --warmup 3
step bench1_nccl 600 python bench.py --gpus 1 --steps 10 --warmup 3 --comm nccl --no-e2e
if [ "$NG" -ge 2 ]; then
  TAILN=40 step gpu_tests_multi 1200 python -m pytest tests/test_gpu.py -x -q -m "multigpu" -p no:cacheprovider -s
  step bench2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps 10 --warmup 3
  step bench2_nccl 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $NG --steps 10 --warmup 3 --comm nccl --no-e2e
  step bench2_plugin 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $NG --steps 6 --warmup 3 --comm nccl-plugin --no-e2e
fi
# profiles: launch list + one full capture of the fused kernel (single GPU only)
step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file $OUT/launches.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-e2e --batch 32
step ncu_fused 600 ncu --set full --clock-control none --import-source on -k regex:bnet_fused -s 6 -c 2 -o $OUT/fused_sgd python bench.py --gpus 1 --steps 2 --warmup 3 --no-e2e
step ncu_exec 600 ncu --set full --clock-control none --import-source on -k regex:bnet_nvl -c 2 -o $OUT/nvl_exec env BNET_PERSISTENT=0 python tests/gpu_worker.py executor
echo "== done $(date -u)"
