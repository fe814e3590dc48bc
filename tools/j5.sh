#!/bin/bash
# smoke() + compute-sanitizer (memcheck, then racecheck) over tools/sanitize_run.py
set -u
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke_r02.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r02.log
timeout 280 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck_r02.txt python tools/sanitize_run.py > gpurun_out/memcheck_r02.out 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_r02.out; tail -3 gpurun_out/memcheck_r02.txt
timeout 170 compute-sanitizer --tool racecheck --racecheck-report analysis --log-file gpurun_out/racecheck_r02.txt python tools/sanitize_run.py > gpurun_out/racecheck_r02.out 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck_r02.out; tail -5 gpurun_out/racecheck_r02.txt
