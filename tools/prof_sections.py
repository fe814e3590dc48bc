import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np, egs_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = egs_b200.workloads.config(cfg, n_pods=n)
e = egs_b200.Egs(w.policy, w.n_nodes)
e.state_load_bulk(0, w.gpus, w.mem_total, w.core, w.mem)
e.profile_reset(True)
e.schedule_batch(w.c_off, w.units, mode=2)
out = (C.c_longlong * 16)()
e.L.egs_debug_resolve_prof.argtypes = [C.c_void_p, C.c_void_p]
e.L.egs_debug_resolve_prof(e.h, out)
v = [int(x) for x in out]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
names = ["fast:precond", "fast:decide", "fast:hazard", "fast:commit", "gen:decide", "gen:commit"]
tot = sum(v[:6])
print(e.rounds_stats(), "resolve ms", e.profile_get(3)[1])
for nme, x in zip(names, v[:6]):
    print(f"  {nme:14s} {x/1e6:9.2f} Mcyc  {100*x/max(tot,1):5.1f}%")
print(f"  fast iterations {v[6]}, pods committed by fast path {v[7]} (avg {v[7]/max(v[6],1):.2f} of W avg {v[8]/max(v[6],1):.2f}), general-path pods {v[9]}")
n=max(v[11],1)
print(f"  head-win events {v[11]}: wait {v[12]/n:.0f}  install {v[13]/n:.0f}  heads {v[14]/n:.0f}  prefetch {v[15]/n:.0f} cycles each")
print(f"  cycles/pod fast {sum(v[:4])/max(v[7],1):.0f}   general {sum(v[4:6])/max(v[9],1):.0f}")
