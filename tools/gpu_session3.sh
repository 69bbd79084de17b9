#!/bin/bash
# Session 3: NCCL-over-plugin after the lazy-loading fix, fused conv blocks, tuning sweeps.
TAG=${1:-s4}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" $OUT/$name.log | tail -${TAILN:-12} | cut -c1-${CUT:-300}; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
step build 600 make -j16
TAILN=30 step gpu_tests_1 700 python -m pytest tests/test_gpu.py -q -m "gpu and not multigpu" -p no:cacheprovider
PENV=$(python -m bagua_net_b200.utils.env --debug)
nccl_case() { local name=$1; shift; TAILN=6 step $name 75 env $PENV BNET_LOG_LEVEL=INFO BNET_WATCHDOG_MS=8000 "$@" $TR --master-port 29551 tests/gpu_worker.py nccl_allreduce; grep -E "nccl-over-plugin|watchdog|WARN" $OUT/$name.log | head -8 | cut -c1-400; }
nccl_case nccl_nvl_gdr_persistent BNET_NVL=1
nccl_case nccl_nvl_gdr_oneshot BNET_NVL=1 BNET_PERSISTENT=0
nccl_case nccl_nvl_gdr_tma BNET_NVL=1 BNET_COPY_ENGINE=tma
nccl_case nccl_tcp_gdr_staged BNET_NVL=0
TAILN=40 step gpu_tests_multi 700 python -m pytest tests/test_gpu.py -q -m "multigpu" -p no:cacheprovider
CUT=1500 step bench1 300 python bench.py --gpus 1 --steps 20 --warmup 5
CUT=1500 step bench1_nofused 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-fused --no-e2e
CUT=1500 step bench2 400 $TR --master-port 29541 bench.py --gpus $NG --steps 20 --warmup 5
TAILN=32 step step_profile_fused 200 python tools/step_profile.py --out $OUT/step_profile_fused.txt
CUT=1500 TAILN=3 step bench2_plugin 240 env BNET_WATCHDOG_MS=8000 $TR --master-port 29543 bench.py --gpus $NG --steps 10 --warmup 3 --comm nccl-plugin --no-e2e
TAILN=60 step sweep_blocks 400 $TR --master-port 29561 bench/allreduce_sweep.py --min-bytes 16M --max-bytes 1G --algos nvls,p2p --blocks 32,64,96,148,200,296 --json $OUT/sweep_blocks.json
TAILN=80 step sweep_all 400 $TR --master-port 29562 bench/allreduce_sweep.py --min-bytes 1K --max-bytes 1G --json $OUT/sweep_all.json
TAILN=20 step p2p_default 200 $TR --master-port 29563 bench/p2p_bw.py
TAILN=20 step p2p_tma 200 env BNET_COPY_ENGINE=tma $TR --master-port 29564 bench/p2p_bw.py
TAILN=40 step nccl_perf_plugin 300 env $(python -m bagua_net_b200.utils.env) build/bench/all_reduce_perf -b 8 -e 128M -f 4 -N $NG -d bfloat16
echo "== done $(date -u)"
