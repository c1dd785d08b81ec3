#!/bin/bash
# N-GPU evidence (round 2): step trace of the communicator path, the world-invariance check through the C-ABI sharded
# search, then the default bench at N GPUs (configs[1] weak scaling + configs[3] search at 500 000 pages per GPU +
# configs[2] ragged top-100) exactly as the driver launches it, and the reference arm.  Tight timeouts everywhere.
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/multi_gpus_$N.txt 2>&1
NCCL_DEBUG=WARN timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    scripts/gpu_comm_diag.py > gpurun_out/comm_diag_$N.log 2>&1
echo "comm diag exit $?"; grep "rank 0" gpurun_out/comm_diag_$N.log | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/sharded_check.py > gpurun_out/multi_check_$N.log 2>&1
echo "sharded_check exit $?"; tail -3 gpurun_out/multi_check_$N.log | cut -c1-600
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/multi_bench_$N.log 2> gpurun_out/multi_bench_$N.err
echo "bench exit $?"; tail -1 gpurun_out/multi_bench_$N.log | cut -c1-7000; tail -3 gpurun_out/multi_bench_$N.err | cut -c1-300
if [ -n "$REF" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/multi_ref_$N.log 2>&1
echo "reference arm exit $?"; tail -1 gpurun_out/multi_ref_$N.log | cut -c1-600
fi
