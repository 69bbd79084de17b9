#!/bin/bash
# Round-2 session 6 (2 GPUs, short): NVLink-side ncu of the transport's executor kernels (one process, two GPUs), and the
# kernel-side sanitizers (memcheck / racecheck / synccheck) over the executor, the fused all-reduce + SGD kernel, the
# fused layer kernels and the tcgen05 kernels.  Everything bounded; logs -> gpurun_out/<tag>/.
TAG=${1:-r2s6}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD BNET_TEST_QUICK=1
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG $(date -u)"
make -j16 >/dev/null 2>&1
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|\*\*\*" $OUT/$name.log | tail -${TAILN:-6} | cut -c1-300; return $rc; }
# 0) the bench with both DDP arms (the plugin arm with lazy module loading + the single-rank DDP dry run)
export BNET_BENCH_STACKS=100 BNET_BENCH_LOG_DIR=$PWD/$OUT/arms
TAILN=12 step bench_full 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29634 bench.py --gpus 2 --steps 20 --warmup 5 --arm-timeout 100
unset BNET_BENCH_STACKS
# 1) bandwidth table without a profiler, then the ncu capture (6 launches of the executor kernels, GPU0 -> GPU1)
TAILN=8 step p2p_time 120 python tools/ncu_p2p.py --time
step ncu_p2p 240 ncu --set full --clock-control none --import-source on -k regex:bnet_nvl -c 6 -f -o $OUT/nvl_p2p python tools/ncu_p2p.py --mbytes 128
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,nvltx__bytes.sum,nvlrx__bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__cluster_x,launch__registers_per_thread"
ncu -i $OUT/nvl_p2p.ncu-rep --page raw --csv --metrics $M > $OUT/nvl_p2p.raw.csv 2>$OUT/nvl_p2p.raw.err; head -c 3000 $OUT/nvl_p2p.raw.csv
ncu -i $OUT/nvl_p2p.ncu-rep --page raw --csv 2>/dev/null | head -1 | tr ',' '\n' | grep -i "nvl\|pcie" | head -20
# 2) sanitizers
CS="compute-sanitizer --error-exitcode 66 --launch-timeout 0"
san() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?;
        echo "---- [$name] rc=$rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$name.log | tail -1)"; grep -E "Invalid|Race reported|hazard|Barrier error|misaligned" $OUT/$name.log | head -4 | cut -c1-250; }
san memcheck_executor_msg 120 env BNET_EXEC_MODE=msg $CS --tool memcheck python tests/gpu_worker.py executor
san racecheck_executor_msg 120 env BNET_EXEC_MODE=msg $CS --tool racecheck python tests/gpu_worker.py executor
san memcheck_fused_sgd 120 $CS --tool memcheck python tests/gpu_worker.py fused_sgd
san racecheck_fused_sgd 120 $CS --tool racecheck python tests/gpu_worker.py fused_sgd
if [ "$2" = "2gpu" ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  san memcheck_allreduce_2gpu 150 $CS --tool memcheck --target-processes all $TR --master-port 29655 tests/gpu_worker.py allreduce
fi
echo "== summary"; grep -H -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/*.log | sed "s#$OUT/##"
echo "== done $(date -u)"
