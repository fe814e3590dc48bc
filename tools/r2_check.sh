#!/bin/bash
# GPU box: tests + a short bench (all under timeouts; per-test timeout so that a hang costs minutes, not the budget)
set -u
mkdir -p gpurun_out
TAG=${1:-r02c}
timeout 1500 python -m pytest tests -m gpu -q --timeout=240 --durations=25 > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -c 7000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
