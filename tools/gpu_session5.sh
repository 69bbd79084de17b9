#!/bin/bash
# Session 5: residency-aware stream kernels under DDP, optimised backward nn kernels (ncu), p2p unrolled.
TAG=${1:-s7}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" $OUT/$name.log | tail -${TAILN:-12} | cut -c1-${CUT:-300}; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
step build 600 make -j16
TAILN=20 step gpu_tests_1 600 python -m pytest tests/test_gpu.py -q -m "gpu and not multigpu" -p no:cacheprovider
TAILN=20 step gpu_tests_multi 600 python -m pytest tests/test_gpu.py -q -m "multigpu" -p no:cacheprovider
CUT=1500 TAILN=3 step bench2_plugin 200 env BNET_WATCHDOG_MS=8000 $TR --master-port 29543 bench.py --gpus $NG --steps 10 --warmup 3 --comm nccl-plugin --no-e2e
grep -E "watchdog" $OUT/bench2_plugin.log | sed 's/\[bnet watchdog\]/\n[bnet watchdog]/g' | grep watchdog | cut -c1-260 | sort | uniq | head -12
CUT=1500 step bench1 300 python bench.py --gpus 1 --steps 20 --warmup 5
CUT=1500 step bench2 400 $TR --master-port 29541 bench.py --gpus $NG --steps 20 --warmup 5
CUT=1500 step bench2_nccl 400 $TR --master-port 29542 bench.py --gpus $NG --steps 20 --warmup 5 --comm nccl --no-e2e
CUT=1500 step bench2_resnet50 400 $TR --master-port 29544 bench.py --gpus $NG --steps 20 --warmup 5 --model resnet50 --no-extra
CUT=1500 step bench2_resnet50_nccl 400 $TR --master-port 29545 bench.py --gpus $NG --steps 20 --warmup 5 --model resnet50 --comm nccl --no-e2e
TAILN=40 step sweep_p2p 300 $TR --master-port 29561 bench/allreduce_sweep.py --min-bytes 4M --max-bytes 1G --algos nvls,p2p,nccl --blocks 0 --json $OUT/sweep_p2p.json
TAILN=30 step sweep_nvls_small 200 $TR --master-port 29562 bench/allreduce_sweep.py --min-bytes 64M --max-bytes 256M --algos nvls --blocks 8,16,24,32,48 --json $OUT/sweep_nvls.json
TAILN=32 step step_profile_fused 200 python tools/step_profile.py --fused --out $OUT/step_profile_fused.txt
step ncu_nn 400 ncu --set full --clock-control none --import-source on -k "regex:relu_bwd_bias_grad|pool_relu_bwd|bias_relu" -s 26 -c 10 -o $OUT/nn_kernels python bench.py --gpus 1 --steps 2 --warmup 3 --no-e2e
echo "== done $(date -u)"
