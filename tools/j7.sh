#!/bin/bash
# does the nvidia-smi sampler perturb the timed steps?  same box, three sampler settings, twice
mkdir -p gpurun_out; : > gpurun_out/clock_sampler_ab.txt
for rep in 1 2; do for lms in 200 1000 0; do
  EGS_CLOCKS_LMS=$lms timeout 60 python bench.py --steps 3 --warmup 2 --no-cpu --no-configs --no-roofline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('lms=$lms', 'value_ms', round(d['ms_per_step'],1), 'e2e_ms', round(d['e2e']['ms_per_step'],1), 'samples', d['clocks']['samples'] if d.get('clocks') else None)" >> gpurun_out/clock_sampler_ab.txt
done; done
cat gpurun_out/clock_sampler_ab.txt
