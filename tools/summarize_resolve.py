"""profiles/resolve_<tag>_ncu.json from gpurun_out/prof_resolve_<tag>.ncu-rep (one k_resolve_mw launch, ncu --set full)."""
import csv, json, subprocess, sys
tag = sys.argv[1]
pods = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(['ncu', '-i', f'gpurun_out/prof_resolve_{tag}.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units, v = rr[0], rr[1], rr[2]
d = {k: (x, u) for k, u, x in zip(h, units, v)}
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "sm__cycles_active.max", "smsp__inst_executed.sum", "smsp__inst_issued.sum",
        "sm__inst_executed.sum.per_cycle_active", "smsp__issue_active.avg.per_cycle_active", "sm__icc_request_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum"]
stalls = sorted(((k.split("issue_stalled_")[1].split("_per_issue")[0], float(x)) for k, (x, u) in d.items()
                 if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")), key=lambda t: -t[1])
out = {"source": "ncu --set full --import-source on --clock-control none -k regex:k_resolve_mw -s 3 -c 1 over tools/prof_sections.py "
                 "(config 4, 100 000 nodes; one round of the batch)",
       "launch": {k: " ".join(d[k]).strip() for k in keys if k in d},
       "warp_stalls_per_issued_instruction": {k: round(x, 3) for k, x in stalls},
       "reading": "ONE CTA on one SM: 16 owner warps, of which one holds the ticket at any time; the others sleep on their mbarrier "
                  "(try_wait) -- ncu books that wait under long_scoreboard -- or prepare / post-process their shape. The kernel is bound by "
                  "the dependent-instruction latency of the warp that holds the ticket (wait = fixed-latency dependency, "
                  "branch_resolving, short_scoreboard = shared memory), not by memory: DRAM/L2 traffic is negligible."}
if pods:
    out["pods_in_this_launch"] = pods
json.dump(out, open(f'profiles/resolve_{tag}_ncu.json', 'w'), indent=1)
print(json.dumps(out, indent=1)[:3000])
