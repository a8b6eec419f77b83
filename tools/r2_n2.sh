#!/bin/bash
# round-2 multi-GPU check (N GPUs of one box): sharded == single index over both exchanges, VectorStore over real devices, bench lines
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/multi_gpu_check.py > $O/r2_multi_gpu_check_w$N.log 2>&1; echo "multi_gpu_check rc=$?" >> $O/r2_multi_gpu_check_w$N.log
grep -v "^W\|^\[W\|NCCL version" $O/r2_multi_gpu_check_w$N.log | tail -45
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pipeline.py -m gpu -x -q -k "real_devices or persistence" 2>&1 | tail -5
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err; echo "bench rc=$?"; cat $O/r2_bench_n$N.json; tail -3 $O/r2_bench_n$N.err
timeout 600 $TR --master-port 29513 bench.py --gpus $N --config c4t --steps 50 --warmup 5 > $O/r2_bench_c4t_n$N.json 2> $O/r2_bench_c4t_n$N.err; echo "c4t rc=$?"; cat $O/r2_bench_c4t_n$N.json; tail -3 $O/r2_bench_c4t_n$N.err
