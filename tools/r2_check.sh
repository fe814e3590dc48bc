#!/bin/bash
# GPU box: differential check of the engines + tests + a short bench (all under timeouts)
set -u
mkdir -p gpurun_out
TAG=${1:-r02a}
timeout 900 python tools/quick_rounds.py > gpurun_out/quick_${TAG}.log 2>&1
echo "quick rc=$?"; tail -40 gpurun_out/quick_${TAG}.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-roofline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
