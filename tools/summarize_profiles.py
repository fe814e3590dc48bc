"""Turns gpurun_out/ ncu artefacts into small tracked summaries under profiles/."""
import collections, csv, json, subprocess, sys, os
tag = sys.argv[1]
out = {}
# launch list
rows = list(csv.reader(open(f'gpurun_out/launches_{tag}.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    n = r[ki].split('(')[0]
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
lines = [f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 1 --warmup 1 --pods 30000 --no-cpu`",
         "# per-launch times are cold-cache and serialised: compare SHARES", f"total_us={tot:.1f}", "kernel,launches,sum_us,share_pct,avg_us"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k},{v[0]},{v[1]:.1f},{100*v[1]/tot:.2f},{v[1]/v[0]:.2f}")
open(f'profiles/launches_{tag}_summary.csv', 'w').write("\n".join(lines) + "\n")
# evaluate kernel full capture
raw = subprocess.run(['ncu', '-i', f'gpurun_out/prof_evaluate_{tag}.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]; units = rr[1]
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
ev = []
for r in rr[2:]:
    ev.append({k: (r[h.index(k)] + " " + units[h.index(k)]).strip() for k in keys if k in h})
json.dump({"source": f"ncu --set full --clock-control none -k regex:k_evaluate -s 3 -c 2 over bench.py roofline leg (4M nodes)", "launches": ev},
          open(f'profiles/evaluate_{tag}_ncu.json', 'w'), indent=1)
print(open(f'profiles/launches_{tag}_summary.csv').read())
print(json.dumps(ev[0], indent=1))
