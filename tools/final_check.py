"""Seconds-long GPU check without torch: smoke(), then batch / restore / uid bookkeeping cycles against the oracle
(exercises the pinned result-buffer pool), then three full-size host-buffer steps timed on the host clock."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np
import __graft_entry__ as g
t0 = time.time()
g.smoke()
import egs_b200, oracle_c
cap = egs_b200.capi
F = ["node", "status", "alloc_mask", "fit_count", "fit_digest", "score_digest"]
w = egs_b200.workloads.config(4, n_nodes=2000, n_pods=20000)
e = egs_b200.Egs(w.policy, w.n_nodes); e.state_load_bulk(0, w.gpus, w.mem_total, w.core, w.mem); e.snapshot()
o = oracle_c.OracleC(w.policy)
for n in range(w.n_nodes):
    o.add_node(800, 8 * w.mem_total); o.set_rows(n, w.core[n], w.mem[n])
uids = np.arange(1, w.n_pods + 1, dtype=np.uint64)
ref = o.schedule_batch(w.c_off, w.units64(), uids=uids)
for cyc in range(4):
    e.restore()
    got = e.schedule_batch(w.c_off, w.units, uids=uids if cyc % 2 == 0 else None)
    assert all(np.array_equal(ref[f], got[f]) for f in F), f"cycle {cyc}"
    if cyc % 2 == 0:
        assert e.pod_known(5) == o.known_pod(5)
        half = egs_b200.workloads.window(w, 0, 500)       # a second, small batch in the same state (new uids)
        e.schedule_batch(half.c_off, half.units, uids=np.arange(10**6, 10**6 + 500, dtype=np.uint64))
        assert e.pod_known(10**6)
    else:
        e.schedule_batch(w.c_off[:101], w.units[: int(w.c_off[100])])      # library uids twice: two auto batches
        e.state_load_bulk(0, w.gpus, w.mem_total, w.core, w.mem)             # node reload drops them
e.close()
print("cycles ok", round(time.time() - t0, 1), "s")
wf = egs_b200.workloads.config(4)
e = egs_b200.Egs(wf.policy, wf.n_nodes)
ms = []
for i in range(4):
    t = time.perf_counter()
    e.state_load_bulk(0, wf.gpus, wf.mem_total, wf.core, wf.mem)
    r = e.schedule_batch(wf.c_off, wf.units)
    ms.append((time.perf_counter() - t) * 1e3)
print("full-size host-buffer steps ms:", [round(x, 1) for x in ms], "FINAL_CHECK_OK", round(time.time() - t0, 1), "s")
