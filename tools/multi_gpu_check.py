"""torchrun entry: node-sharded rounds engine (NCCL allgather of candidate buffers) must equal the
unsharded result bit for bit.  Usage: python -m torch.distributed.run --nproc-per-node N tools/multi_gpu_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.distributed as dist

import egs_b200

F = ["node", "status", "alloc_mask", "fit_count", "fit_digest", "score_digest"]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cap = egs_b200.capi
    fails = 0
    cases = [(1, None, None, None), (2, None, 20000, None), (4, None, 60000, None), (3, None, 1500, 1), (4, 300, 6000, 0), (3, 300, 3000, 0)]
    if os.environ.get("EGS_MGC_FAST"):
        cases = [(4, None, 40000, None), (1, None, None, None), (2, 3000, 6000, None)]
    for cfg, nn, npods, pol in cases:
        w = egs_b200.workloads.config(cfg, n_nodes=nn, n_pods=npods, policy=pol)
        ref = None
        if rank == 0:
            e0 = egs_b200.Egs(w.policy, w.n_nodes, 8, local)
            e0.state_load_bulk(0, w.gpus, w.mem_total, w.core, w.mem)
            ref = e0.schedule_batch(w.c_off, w.units, mode=cap.EGS_MODE_ROUNDS)
            ref_rows = e0.state_dump()[:2]
            e0.close()
        e = egs_b200.Egs(w.policy, w.n_nodes, 8, local)
        e.shard_set(rank, world)
        box = [cap.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        e.comm_init(box[0])
        e.state_load_bulk(0, w.gpus, w.mem_total, w.core, w.mem)
        got = e.schedule_batch(w.c_off, w.units, mode=cap.EGS_MODE_ROUNDS)
        lo, hi = cap.shard_range(w.n_nodes, rank, world)
        core, mem, _, _ = e.state_dump(lo, hi - lo)
        # every rank's replicated outputs must equal rank 0's unsharded run; rows of the own shard too
        obj = [ref, ref_rows if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        ref, ref_rows = obj
        bad = [f for f in F if not np.array_equal(ref[f], got[f])]
        rows_ok = np.array_equal(core, ref_rows[0][lo:hi]) and np.array_equal(mem, ref_rows[1][lo:hi])
        print(f"[rank {rank}/{world}] cfg{cfg} N={w.n_nodes} P={w.n_pods} shard=[{lo},{hi}) mismatch={bad} rows_ok={rows_ok} {e.rounds_stats()}", flush=True)
        fails += bool(bad) + (not rows_ok)
        e.close()
    t = torch.tensor([fails], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    if t.item():
        sys.exit(1)
    if rank == 0:
        print("MULTI_GPU_OK")


if __name__ == "__main__":
    main()
