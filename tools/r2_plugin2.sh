#!/bin/bash
# 2-GPU reproduction of the 4-GPU crash: which of the round-2 changes kills the ranks (backtraces on)
TAG=${1:-r2p2}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
make -j16 >/dev/null 2>&1; make bench >/dev/null 2>&1
BASE="$(python -m bagua_net_b200.utils.env) BNET_WATCHDOG_MS=5000 NCCL_DEBUG=WARN"
ARP="build/bench/all_reduce_perf -N $NG -d bfloat16"
run() { local name=$1 tmo=$2 pargs=$3; shift 3; echo "---- [$name] $(date -u +%T) $*"; timeout -k 5 $tmo env $BASE "$@" $ARP $pargs > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "^$" $OUT/$name.log | tail -${TAILN:-18} | cut -c1-330; return $rc; }
SWEEP="-b 8 -e 128M -f 4 -n 10 -w 3"
BIG="-b 1M -e 128M -f 8 -n 10 -w 3"
run default 60 "$SWEEP"
TAILN=40 run pci_gpu 40 "-b 8 -e 1M -f 8 -n 5 -w 2" BNET_GPU_DEVICE_PCI=gpu
run no_gpu_devs 60 "$SWEEP" BNET_GPU_DEVICES=0
TAILN=60 run info 40 "-b 8 -e 1M -f 8 -n 5 -w 2" NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,NET,GRAPH BNET_LOG_LEVEL=2
run persistent 60 "$BIG" BNET_EXEC_MODE=persistent
run ce 60 "$BIG" BNET_EXEC_MODE=ce
run vhost 60 "$SWEEP -H"
for bs in 8388608 33554432; do for ch in 8 16 32; do
  TAILN=4 run tune_b${bs}_c${ch} 45 "$BIG" NCCL_BUFFSIZE=$bs NCCL_MIN_NCHANNELS=$ch
done; done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" $OUT/$name.log | tail -${TAILN:-6} | cut -c1-${CUT:-2500}; return $rc; }
TAILN=3 step bench2 400 $TR --master-port 29621 bench.py --gpus $NG --steps 20 --warmup 5
TAILN=12 step tc_conv_check 200 python -c "
import torch
from bagua_net_b200.ops import tc_conv, tc_linear
print('tc_linear self_check', tc_linear.self_check(verbose=True))
print('tc_conv self_check', tc_conv.self_check(verbose=True))"
TAILN=30 step tc_timing 200 python tools/tc_linear_bench.py
TAILN=8 step gpu_tests 600 python -m pytest tests -q -m gpu -x
echo "== done $(date -u)"
