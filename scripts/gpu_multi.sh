#!/bin/bash
N=${N:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/sharded_check.py > gpurun_out/multi_check_$N.log 2>&1
echo "sharded_check exit $?"; tail -4 gpurun_out/multi_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/multi_bench_$N.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/multi_bench_$N.log | cut -c1-1500
