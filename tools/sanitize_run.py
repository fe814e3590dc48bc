"""Small end-to-end run for compute-sanitizer: verbs + both batch engines incl. mixed shapes."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import numpy as np
import egs_b200
w = egs_b200.workloads.config(4, n_nodes=3000, n_pods=2500)
for mode in (1, 2):
    e = egs_b200.Egs(w.policy, w.n_nodes)
    e.state_load_bulk(0, w.gpus, w.mem_total, w.core, w.mem)
    e.schedule_batch(w.c_off, w.units, mode=mode)
    e.filter(None, [(10, 4096, 0)]); e.score(None, [(10, 4096, 0)]); e.bind(3, [(10, 4096, 0)], 9)
    e.pod_apply(5, [(0, 0, 1), (20, 100, 0)], [[1], [2]], 11); e.pod_cancel(5, [(0, 0, 1), (20, 100, 0)], [[1], [2]], 11)
    e.profile_evaluate([(25, 8192, 0)], iters=2)
    e.close()
w3 = egs_b200.workloads.config(3, n_nodes=500, n_pods=600)
e = egs_b200.Egs(0, w3.n_nodes); e.state_load_bulk(0, w3.gpus, w3.mem_total, w3.core, w3.mem)
e.schedule_batch(w3.c_off, w3.units, mode=2); e.close()
print("sanitize run done")
