#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list, one full ncu capture of the
# evaluate kernel.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 4000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
tail -c 1500 gpurun_out/bench_ref_${TAG}.json
# launch list (per-launch device time; cold-cache, serialised: compare shares)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --pods 30000 --no-cpu > gpurun_out/ncu_bench_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_bench_${TAG}.log | cut -c1-300
# full capture of the evaluate kernel on the HBM-resident 4M-node cluster
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_evaluate -s 3 -c 2 -o gpurun_out/prof_evaluate_${TAG} -f \
    python bench.py --roofline-only > gpurun_out/ncu_eval_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_eval_${TAG}.log | cut -c1-300
ls -la gpurun_out | tail -12
