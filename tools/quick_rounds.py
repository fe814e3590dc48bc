"""GPU quick check: rounds engine vs rescan vs oracle on a few workloads + timing."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np
import egs_b200, oracle_c
F = ["node", "status", "alloc_mask", "fit_count", "fit_digest", "score_digest"]
def gpu(w, mode):
    e = egs_b200.Egs(w.policy, w.n_nodes)
    e.state_load_bulk(0, w.gpus, w.mem_total, w.core, w.mem)
    e.profile_reset(True)
    t = time.perf_counter(); out = e.schedule_batch(w.c_off, w.units, mode=mode); dt = time.perf_counter() - t
    return e, out, dt
for cfg, nn, npods, pol in [(0, None, None, None), (1, 40, 4000, 0), (2, 40, 4000, 1), (4, 40, 4000, 0), (3, 40, 3000, 1), (3, 40, 3000, 0),
                            (1, None, None, None), (2, None, 20000, None), (3, None, 3000, None), (4, None, 50000, None), (4, None, 200000, None)]:
    w = egs_b200.workloads.config(cfg, n_nodes=nn, n_pods=npods, policy=pol)
    e1, a, t1 = gpu(w, 1)
    e2, b, t2 = gpu(w, 2)
    bad = [f for f in F if not np.array_equal(a[f], b[f])]
    rows_ok = all(np.array_equal(x, y) for x, y in zip(e1.state_dump()[:2], e2.state_dump()[:2]))
    print(f"cfg{cfg} N={w.n_nodes} P={w.n_pods} pol={w.policy}: rescan {w.n_pods/t1:,.0f}/s rounds {w.n_pods/t2:,.0f}/s  mismatch={bad} rows_ok={rows_ok} {e2.rounds_stats()}", flush=True)
    print("   ms: select %.2f merge %.2f resolve %.2f  total %.2f" % (e2.profile_get(2)[1], e2.profile_get(4)[1], e2.profile_get(3)[1], t2*1e3))
    if bad:
        f = bad[0]; i = int(np.argwhere(a[f] != b[f])[0][0])
        print("  first diff pod", i, {k: (a[k][i], b[k][i]) for k in F})
