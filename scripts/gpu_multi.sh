#!/bin/bash
# N-GPU evidence: sharded search == unsharded, the default bench, and (BIG=1) BASELINE configs[3] at true scale:
# 500 000 pages x 1030 tokens per GPU (131.8 GB each; 4 M pages on 8 GPUs), single-query top-10 p50 incl. all-gather + merge.
N=${N:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/sharded_check.py > gpurun_out/multi_check_$N.log 2>&1
echo "sharded_check exit $?"; tail -4 gpurun_out/multi_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/multi_bench_$N.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/multi_bench_$N.log | cut -c1-1500
if [ "${BIG:-0}" = "1" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $N --steps 5 --warmup 3 --pages 500000 --search-iters 200 --no-cpu > gpurun_out/multi_bench_big_$N.log 2>&1
  echo "bench_big exit $?"; tail -1 gpurun_out/multi_bench_big_$N.log | cut -c1-300; tail -1 gpurun_out/multi_bench_big_$N.log | grep -o '"search": {[^}]*}'
fi
if [ "${C3:-0}" = "1" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
      scripts/gpu_c3_sharded.py > gpurun_out/multi_c3_$N.log 2>&1
  echo "c3_sharded exit $?"; grep '^{' gpurun_out/multi_c3_$N.log | cut -c1-420
fi
