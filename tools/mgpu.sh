#!/bin/bash
# on a multi-GPU box (gpurun --gpus N): sharded parity check + the bench line at N GPUs
set -u
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
EGS_MGC_FAST=${EGS_MGC_FAST:-} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 \
    tools/multi_gpu_check.py > gpurun_out/multi_gpu_check_${TAG}_n${N}.log 2>&1
echo "check rc=$?"; grep -E "MULTI_GPU_OK|mismatch=\[.+\]|rank 0" gpurun_out/multi_gpu_check_${TAG}_n${N}.log | cut -c1-220 | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29742 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-roofline > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n${N}.err
echo "bench rc=$?"; tail -c 3500 gpurun_out/bench_${TAG}_n${N}.json; tail -3 gpurun_out/bench_${TAG}_n${N}.err
