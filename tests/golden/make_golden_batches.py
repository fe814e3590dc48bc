"""Generates tests/golden/batches.json: small clusters + pod batches with the per-pod outputs of the driver
rule computed by the python mirror of the reference (oracle/egs_oracle.py).

    python tests/golden/make_golden_batches.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import egs_b200  # noqa: E402
import egs_oracle as po  # noqa: E402


def run(policy, nodes, pods):
    s = po.Scheduler(policy)
    for g, m, rows in nodes:
        n = s.add_node(100 * g, m * g)
        if rows:
            s.set_rows(n, rows[0], rows[1])
    out = []
    for uid, req in enumerate(pods):
        r = s.schedule_one([tuple(u) for u in req], uid)
        masks = [0, 0, 0, 0]
        if r["status"] == 0:
            for c, a in enumerate(r["alloc"]):
                for gi in a:
                    masks[c] |= 1 << gi
        out.append([r["node"], r["status"], masks, r["fit_count"], str(r["fit_digest"]), str(r["score_digest"])])
    final = [s.rows(n) for n in range(len(nodes))]
    return out, final


def main():
    cases = []
    # BASELINE config 0 (KA-0) and small prefixes of configs 1, 2, 4 on the first nodes of their clusters
    for cfg, nn, npods in [(0, None, None), (1, 24, 400), (2, 24, 400), (4, 32, 500), (3, 16, 200)]:
        w = egs_b200.workloads.config(cfg, n_nodes=nn, n_pods=npods)
        nodes = [(w.gpus, w.mem_total, ([int(x) for x in w.core[n]], [int(x) for x in w.mem[n]])) for n in range(w.n_nodes)]
        pods = [[[int(x) for x in w.units[k]] for k in range(w.c_off[p], w.c_off[p + 1])] for p in range(w.n_pods)]
        out, final = run(w.policy, nodes, pods)
        cases.append(dict(name=f"cfg{cfg}", policy=w.policy, nodes=nodes, pods=pods, out=out, final=final))
    # mixed shapes: sentinel, whole-GPU, 1..4 containers, heterogeneous nodes
    rng = np.random.default_rng(20260921)
    for policy in (0, 1):
        nodes = []
        for _ in range(20):
            g = int(rng.choice([1, 2, 4, 8])); m = int(rng.choice([16, 40, 80]))
            rows = ([int(rng.choice([100, 100, 60, 30, 0])) for _ in range(g)], [int(rng.integers(0, m + 1)) for _ in range(g)]) if rng.integers(0, 2) else None
            nodes.append((g, m, rows))
        shapes = []
        for _ in range(7):
            units = []
            for _ in range(int(rng.integers(1, 5))):
                k = rng.integers(0, 10)
                units.append([-1, -1, 0] if k == 0 else [0, 0, int(rng.integers(1, 3))] if k == 1 else
                             [int(rng.choice([0, 5, 10, 25, 50])), int(rng.integers(1, 12)), 0])
            shapes.append(units)
        pods = [shapes[int(i)] for i in rng.integers(0, 7, 300)]
        out, final = run(policy, nodes, pods)
        cases.append(dict(name=f"mixed-policy{policy}", policy=policy, nodes=nodes, pods=pods, out=out, final=final))
    with open(os.path.join(HERE, "batches.json"), "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    print("wrote", len(cases), "cases,", os.path.getsize(os.path.join(HERE, "batches.json")), "bytes")


if __name__ == "__main__":
    main()
