#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list of a short bench, one full ncu capture of the resolver, the
# cycle counters of the -DEGS_RESOLVE_PROF build, and the evaluate kernel capture.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --pods 30000 --no-cpu --no-roofline --no-configs > gpurun_out/ncu_bench_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_bench_${TAG}.log | cut -c1-300
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_resolve_mw -s 3 -c 1 -o gpurun_out/prof_resolve_${TAG} -f \
    python tools/prof_sections.py 60000 4 > gpurun_out/ncu_resolve_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_resolve_${TAG}.log | cut -c1-300
EGS_LIB=libegs_prof.so timeout 200 python tools/prof_sections.py 300000 4 > gpurun_out/resolve_sections_${TAG}.txt 2>&1
cat gpurun_out/resolve_sections_${TAG}.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_evaluate -s 3 -c 2 -o gpurun_out/prof_evaluate_${TAG} -f \
    python bench.py --roofline-only > gpurun_out/ncu_eval_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_eval_${TAG}.log | cut -c1-300
ls -la gpurun_out | tail -8
