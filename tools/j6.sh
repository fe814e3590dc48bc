#!/bin/bash
# targeted re-validation after a resolver edit: sharded lock step, late-window oracle parity, mutations, full-size properties, bench
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_sharding.py tests/test_gpu_mutations.py tests/test_gpu_parity.py -q -m gpu --timeout=200 \
  -k "in_process or mutations or late_windows and (250000 or 990000 or 90000 or 400000) or mixed_shapes or full_size or fit_and_score" \
  > gpurun_out/pytest_r02e.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r02e.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r02e.json
