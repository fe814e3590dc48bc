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
print("sanitize run (round 1 paths) done")

# ---- round 2: multi-warp resolver with many shapes, vectors, option dump, mutations, in-process shards ----
import threading
cap = egs_b200.capi
w2 = egs_b200.workloads.config(2, n_nodes=1500, n_pods=4000)            # spread, core+memory shapes
e = egs_b200.Egs(w2.policy, w2.n_nodes); e.state_load_bulk(0, w2.gpus, w2.mem_total, w2.core, w2.mem)
e.schedule_batch_vec(w2.c_off, w2.units, 64)
for req in egs_b200.workloads.shapes_of(w2)[:3]:
    e.option_dump(req)
e.close()
w4 = egs_b200.workloads.config(4, n_nodes=1500, n_pods=3000)
e = egs_b200.Egs(w4.policy, w4.n_nodes); e.state_load_bulk(0, w4.gpus, w4.mem_total, w4.core, w4.mem)
recs = [(cap.EGS_MUT_ADD, n, [(10, 4096, 0)] * 6, [[n % 8]] * 6, 900000 + n) for n in range(0, 300, 7)]
recs += [(cap.EGS_MUT_FORGET, n, [(10, 4096, 0)] * 6, [[n % 8]] * 6, 900000 + n) for n in range(0, 300, 14)]
e.mutations_apply(recs[:20])
mut_at = sorted(int(x) for x in np.random.default_rng(3).integers(1, w4.n_pods, len(recs) - 20))
e.schedule_batch_mut(w4.c_off, w4.units, mut_at, recs[20:], uids=np.arange(1, w4.n_pods + 1, dtype=np.uint64))
e.close()
for world in (2, 4):
    hs = []
    for r in range(world):
        h = egs_b200.Egs(w4.policy, w4.n_nodes); h.shard_set(r, world); hs.append(h)
    cap.comm_init_local(hs)
    for h in hs:
        h.state_load_bulk(0, w4.gpus, w4.mem_total, w4.core, w4.mem)
    th = [threading.Thread(target=lambda h=h: h.schedule_batch(w4.c_off, w4.units, mode=cap.EGS_MODE_ROUNDS)) for h in hs]
    [t.start() for t in th]; [t.join() for t in th]
    for h in hs:
        h.close()
print("sanitize run (round 2 paths) done")
