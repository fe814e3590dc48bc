"""Short rounds-engine run for ncu captures (cfg4 cluster, pod prefix)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import egs_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
w = egs_b200.workloads.config(4, n_pods=n)
e = egs_b200.Egs(w.policy, w.n_nodes)
e.state_load_bulk(0, w.gpus, w.mem_total, w.core, w.mem)
out = e.schedule_batch(w.c_off, w.units, mode=2)
print(e.rounds_stats())
