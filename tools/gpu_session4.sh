#!/bin/bash
# Session 4: DDP over the plugin after arming kernels at setup, NCCL tuning knobs, sweeps, profiles.
TAG=${1:-s6}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" $OUT/$name.log | tail -${TAILN:-12} | cut -c1-${CUT:-300}; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
step build 600 make -j16
TAILN=20 step gpu_test_fused 300 python -m pytest tests/test_gpu.py -q -m "gpu and not multigpu" -p no:cacheprovider -k "fused_conv or executor_copy_reduce_cast"
PENV=$(python -m bagua_net_b200.utils.env --debug)
nccl_case() { local name=$1; shift; TAILN=4 step $name 75 env $PENV BNET_LOG_LEVEL=INFO BNET_WATCHDOG_MS=8000 "$@" $TR --master-port 29551 tests/gpu_worker.py nccl_allreduce; grep -E "nccl-over-plugin|watchdog|WARN" $OUT/$name.log | head -6 | cut -c1-400; }
nccl_case nccl_tcp_gdr_staged BNET_NVL=0
CUT=1500 TAILN=3 step bench2_plugin 200 env BNET_WATCHDOG_MS=8000 $TR --master-port 29543 bench.py --gpus $NG --steps 10 --warmup 3 --comm nccl-plugin --no-e2e
grep -E "watchdog" $OUT/bench2_plugin.log | head -4 | cut -c1-300
PQ=$(python -m bagua_net_b200.utils.env)
TAILN=16 step nccl_perf_plugin 200 env $PQ build/bench/all_reduce_perf -b 8 -e 128M -f 4 -N $NG -d bfloat16
TAILN=16 step nccl_perf_plugin_simple 200 env $PQ NCCL_PROTO=Simple build/bench/all_reduce_perf -b 8 -e 128M -f 4 -N $NG -d bfloat16
TAILN=8 step nccl_perf_plugin_bigbuf 200 env $PQ NCCL_PROTO=Simple NCCL_BUFFSIZE=33554432 build/bench/all_reduce_perf -b 1M -e 512M -f 4 -N $NG -d bfloat16
TAILN=8 step nccl_perf_plugin_chan 200 env $PQ NCCL_PROTO=Simple NCCL_BUFFSIZE=33554432 NCCL_MIN_NCHANNELS=16 build/bench/all_reduce_perf -b 1M -e 512M -f 4 -N $NG -d bfloat16
TAILN=60 step sweep_blocks 300 $TR --master-port 29561 bench/allreduce_sweep.py --min-bytes 16M --max-bytes 1G --algos nvls,p2p --blocks 32,64,96,148,200,296 --json $OUT/sweep_blocks.json
TAILN=80 step sweep_all 300 $TR --master-port 29562 bench/allreduce_sweep.py --min-bytes 1K --max-bytes 1G --json $OUT/sweep_all.json
TAILN=32 step step_profile_fused 200 python tools/step_profile.py --fused --out $OUT/step_profile_fused.txt
step ncu_nn 400 ncu --set full --clock-control none --import-source on -k regex:bnet::nn -s 40 -c 8 -o $OUT/nn_kernels python bench.py --gpus 1 --steps 2 --warmup 3 --no-e2e
step ncu_fused 300 ncu --set full --clock-control none --import-source on -k regex:bnet_fused -s 6 -c 3 -o $OUT/fused_sgd python bench.py --gpus 1 --steps 2 --warmup 3 --no-e2e
echo "== done $(date -u)"
