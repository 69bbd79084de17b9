#!/bin/bash
# Scaling session on an N-GPU box (gpurun --gpus N): flagship bench at every power of two up to N (ours and torch DDP
# over stock NCCL), ResNet-50, and the all-reduce sweep at N.   usage: tools/gpu_session_scale.sh [tag] [ngpus]
TAG=${1:-scale}; NG=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|OMP_NUM" $OUT/$name.log | tail -${TAILN:-2} | cut -c1-${CUT:-1800}; return $rc; }
step build 300 make -j16
port=29700
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
  port=$((port+1)); step bench_${n} 300 $TR --master-port $port bench.py --gpus $n --steps 30 --warmup 5
  port=$((port+1)); step bench_${n}_nccl 300 $TR --master-port $port bench.py --gpus $n --steps 20 --warmup 5 --comm nccl --no-e2e
  port=$((port+1)); step bench_${n}_resnet50 300 $TR --master-port $port bench.py --gpus $n --steps 20 --warmup 5 --model resnet50 --no-extra
  port=$((port+1)); step bench_${n}_resnet50_nccl 300 $TR --master-port $port bench.py --gpus $n --steps 20 --warmup 5 --model resnet50 --comm nccl --no-e2e
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
TAILN=90 CUT=200 step sweep_all 400 $TR --master-port 29790 bench/allreduce_sweep.py --min-bytes 1K --max-bytes 1G --json $OUT/sweep_all.json
echo "== done $(date -u)"
