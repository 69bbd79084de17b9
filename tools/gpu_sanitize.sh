#!/bin/bash
# Kernel-side sanitizer runs: compute-sanitizer memcheck, racecheck, synccheck
# over the transport executor, the fused layer kernels, the fused all-reduce + SGD kernel and the tcgen05 kernels
# (1 GPU), and memcheck over the cross-GPU collectives when two GPUs are there.  Logs -> gpurun_out/<tag>/, the
# summaries are committed under profiles/.
#   gpurun --timeout 1200 -- 'bash tools/gpu_sanitize.sh san1'
TAG=${1:-san}; NG=${2:-1}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD BNET_TEST_QUICK=1
exec > >(tee $OUT/session.log) 2>&1
echo "== sanitizer session $TAG ngpus=$NG $(date -u)"
make -j16 >/dev/null 2>&1
CS="compute-sanitizer --error-exitcode 66 --launch-timeout 0"
run() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?;
        echo "---- [$name] rc=$rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$name.log | tail -1)"; grep -E "Invalid|Race reported|hazard|Barrier error|misaligned" $OUT/$name.log | head -5; }
for tool in memcheck synccheck racecheck; do
  run ${tool}_executor 300 $CS --tool $tool python tests/gpu_worker.py executor
  run ${tool}_executor_msg 300 env BNET_EXEC_MODE=msg $CS --tool $tool python tests/gpu_worker.py executor
  run ${tool}_fused_sgd 300 $CS --tool $tool python tests/gpu_worker.py fused_sgd
  run ${tool}_fused_nn 400 $CS --tool $tool python tests/gpu_worker.py fused_nn
  run ${tool}_pack_cast 200 $CS --tool $tool python tests/gpu_worker.py pack_cast
done
run memcheck_tc_linear 400 $CS --tool memcheck python tests/gpu_worker.py tc_linear
run memcheck_tc_conv 400 $CS --tool memcheck python tests/gpu_worker.py tc_conv
run synccheck_tc_conv 400 $CS --tool synccheck python tests/gpu_worker.py tc_conv
if [ "$NG" -ge 2 ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  run memcheck_allreduce_2gpu 600 $CS --tool memcheck --target-processes all $TR --master-port 29655 tests/gpu_worker.py allreduce
  run memcheck_fused_sgd_2gpu 400 $CS --tool memcheck --target-processes all $TR --master-port 29656 tests/gpu_worker.py fused_sgd
  run racecheck_fused_sgd_2gpu 400 $CS --tool racecheck --target-processes all $TR --master-port 29657 tests/gpu_worker.py fused_sgd
fi
echo "== summary"; grep -H -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/*.log | sed "s#$OUT/##"
echo "== done $(date -u)"
