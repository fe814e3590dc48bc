import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np, egs_b200
w = egs_b200.workloads.config(4)
e = egs_b200.Egs(w.policy, w.n_nodes)
e.state_load_bulk(0, w.gpus, w.mem_total, w.core, w.mem)
e.profile_reset(True)
prev = (0, 0.0, 0.0, 0.0, {})
for chunk in range(10):
    sub_off = w.c_off[chunk*100000:(chunk+1)*100000+1] - w.c_off[chunk*100000]
    sub_units = w.units[w.c_off[chunk*100000]:w.c_off[(chunk+1)*100000]]
    t = time.perf_counter(); e.schedule_batch(sub_off, sub_units, mode=2); dt = time.perf_counter() - t
    st = e.rounds_stats(); sel, mer, res = e.profile_get(2)[1], e.profile_get(4)[1], e.profile_get(3)[1]
    print(f"pods {chunk*100}k-{(chunk+1)*100}k: wall {dt*1e3:6.0f} ms  select {sel-prev[1]:6.1f} merge {mer-prev[2]:5.1f} resolve {res-prev[3]:6.1f} ms  rounds {st['rounds']-prev[0]:4d} tracked {st['tracked']-prev[4].get('tracked',0):6d} dry {st['stop_list_dry']-prev[4].get('stop_list_dry',0)} full {st['stop_tracked_full']-prev[4].get('stop_tracked_full',0)}", flush=True)
    prev = (st['rounds'], sel, mer, res, st)
