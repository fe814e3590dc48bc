#!/bin/bash
# GPU box: tests + a short bench (all under timeouts)
set -u
mkdir -p gpurun_out
TAG=${1:-r02b}
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
