#!/bin/bash
# Round-2 session 9 (8 GPUs, one minute): where do the 8-rank ring's ~90 us per step go?  Launch statistics of the device
# executor behind NCCL, and three knobs (batching window, channel count) at 32 MiB / 128 MiB.
TAG=${1:-r2s9}; NG=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
BASE="$(python -m bagua_net_b200.utils.env) BNET_WATCHDOG_MS=5000 NCCL_DEBUG=WARN BNET_EXEC_STATS=1"
ARP="build/bench/all_reduce_perf -N $NG -d bfloat16 -b 32M -e 128M -f 4 -n 8 -w 2"
run() { local name=$1; shift; echo "---- [$name] $(date -u +%T) $*"; timeout -k 3 19 env $BASE "$@" $ARP > $OUT/$name.log 2>&1; echo "rc=$?"; grep -E "^ +[0-9]+ +[0-9]+ +bf16|bnet stats" $OUT/$name.log | head -4 | cut -c1-200; }
run base
run window20 BNET_MSG_BATCH_US=20
run ch8 NCCL_MIN_NCHANNELS=8 NCCL_MAX_NCHANNELS=8
[ -n "$MORE" ] && run buf8 NCCL_BUFFSIZE=8388608
echo "== done $(date -u)"
