"""Evaluate-kernel roofline sweep on the GPU box (4M nodes, > L2)."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np
import egs_b200
w = egs_b200.workloads.config(4, n_pods=64)
N = 4_000_000
reps = (N + w.n_nodes - 1) // w.n_nodes
for pol in (0, 1):
    e = egs_b200.Egs(pol, N)
    e.state_load_bulk(0, w.gpus, w.mem_total, np.tile(w.core, (reps, 1))[:N], np.tile(w.mem, (reps, 1))[:N])
    for items in (1, 2, 4):
        os.environ["EGS_EVAL_ITEMS"] = str(items)
        e.profile_evaluate([(25, 8192, 0)], iters=3)
        ms = e.profile_evaluate([(25, 8192, 0)], iters=30)
        print(f"policy {pol} items {items}: {ms*1e3:.1f} us/launch  {N*70/ms/1e6:.0f} GB/s  frac {N*70/ms/1e6/6566.7:.3f}", flush=True)
    e.close()
