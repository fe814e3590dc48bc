"""Seeded random scenarios that drive any backend (python mirror, C oracle, libegs on a GPU)
through the same verb sequence and record a comparable trace."""
from __future__ import annotations

import numpy as np

import egs_oracle as po
import oracle_c as oc


# ---------------------------------------------------------------- backends
class PyBackend:
    def __init__(self, policy):
        self.s = po.Scheduler(policy)

    def add_node(self, core_alloc, mem_alloc):
        return self.s.add_node(core_alloc, mem_alloc)

    def set_rows(self, node, core, mem):
        self.s.set_rows(node, core, mem)

    def rows(self, node):
        return self.s.rows(node)

    def filter(self, ids, req):
        return [int(x) for x in self.s.assume(ids, list(req))]

    def score(self, ids, req):
        out, st = [], 0
        for n in ids:
            try:
                out.append(self.s.score([n], list(req))[0])
            except RuntimeError:
                st = 9
                out.append(0)
        return st, out

    def bind(self, node, req, uid):
        return self.s.bind(node, list(req), uid)

    def peek(self, node, req):
        o = self.s.nodes[node].allocated.get(tuple(req))
        return None if o is None else (o.allocated, o.score)

    def add_pod(self, node, req, alloc, uid):
        self.s.add_pod(node, list(req), alloc, uid)

    def forget_pod(self, node, req, alloc, uid):
        self.s.forget_pod(node, list(req), alloc, uid)

    def known(self, uid):
        return uid in self.s.pod_maps

    def released(self, uid):
        return uid in self.s.released


class CBackend:
    def __init__(self, policy, faithful=False, threads=1):
        self.o = oc.OracleC(policy, faithful)
        self.threads = threads

    def add_node(self, core_alloc, mem_alloc):
        return self.o.add_node(core_alloc, mem_alloc)

    def set_rows(self, node, core, mem):
        self.o.set_rows(node, core, mem)

    def rows(self, node):
        return self.o.rows(node)

    def filter(self, ids, req):
        return [int(x) for x in self.o.filter(ids, req, self.threads)]

    def score(self, ids, req):
        st, sc = self.o.score(ids, req)
        return st, [int(x) for x in sc]

    def bind(self, node, req, uid):
        return self.o.bind(node, req, uid)

    def peek(self, node, req):
        return self.o.peek(node, req)

    def add_pod(self, node, req, alloc, uid):
        self.o.add_pod(node, req, alloc, uid)

    def forget_pod(self, node, req, alloc, uid):
        self.o.forget_pod(node, req, alloc, uid)

    def known(self, uid):
        return self.o.known_pod(uid)

    def released(self, uid):
        return self.o.released_pod(uid)


class GpuBackend:
    """libegs through the C ABI (needs a GPU)."""

    def __init__(self, policy, max_nodes=64):
        import egs_b200
        self.e = egs_b200.Egs(policy, max_nodes, 8, 0)
        self.n = 0

    def add_node(self, core_alloc, mem_alloc):
        nid = self.n
        self.n += 1
        st = self.e.node_set_allocatable(nid, core_alloc, mem_alloc)
        return nid if st == 0 else -1

    def set_rows(self, node, core, mem):
        assert self.e.state_load(node, core, mem) == 0

    def rows(self, node):
        return self.e.rows(node)

    def filter(self, ids, req):
        return [int(x) for x in self.e.filter(ids, req)]

    def score(self, ids, req):
        st, sc = self.e.score(ids, req)
        return st, [int(x) for x in sc]

    def bind(self, node, req, uid):
        return self.e.bind(node, req, uid)

    def peek(self, node, req):
        return self.e.peek(node, req)

    def add_pod(self, node, req, alloc, uid):
        self.e.pod_apply(node, req, alloc, uid)

    def forget_pod(self, node, req, alloc, uid):
        self.e.pod_cancel(node, req, alloc, uid)

    def known(self, uid):
        return self.e.pod_known(uid)

    def released(self, uid):
        return self.e.pod_released(uid)


# ---------------------------------------------------------------- scenario
def random_unit(rng, mem_hi):
    k = rng.integers(0, 10)
    if k == 0:
        return (-1, -1, 0)                                   # container without GPU request
    if k == 1:
        return (0, 0, int(rng.integers(1, 4)))               # whole GPUs
    core = int(rng.choice([0, 5, 10, 20, 25, 30, 50, 75, 99]))
    mem = int(rng.integers(0, mem_hi + 1)) if rng.integers(0, 4) else 0
    if core == 0 and mem == 0:
        mem = 1
    return (core, mem, 0)


def make_scenario(seed, max_nodes=6, n_ops=40, max_c=3):
    rng = np.random.default_rng(seed)
    n_nodes = int(rng.integers(1, max_nodes + 1))
    nodes = []
    for _ in range(n_nodes):
        G = int(rng.choice([1, 2, 4, 8]))
        M = int(rng.choice([12, 16, 80, 1000]))
        rows = None
        if rng.integers(0, 2):
            rows = ([int(rng.choice([100, 100, 75, 50, 20, 0])) for _ in range(G)],
                    [int(rng.integers(0, M + 1)) if rng.integers(0, 2) else M for _ in range(G)])
        nodes.append((G * 100 + int(rng.integers(0, 100)), G * M + int(rng.integers(0, G)), rows))
    if rng.integers(0, 8) == 0:
        nodes.append((50, 10, None))                         # G == 0: "no gpu available on node"
    mem_hi = 20
    shapes = [tuple(random_unit(rng, mem_hi) for _ in range(int(rng.integers(1, max_c + 1))))
              for _ in range(int(rng.integers(1, 5)))]
    ops = []
    for i in range(n_ops):
        k = rng.integers(0, 12)
        shape = shapes[int(rng.integers(0, len(shapes)))]
        if k < 7:
            ops.append(("sched", shape, 1000 + i))
        elif k == 7:
            node = int(rng.integers(0, n_nodes))
            ops.append(("add_pod", shape, 2000 + i, node, int(rng.integers(0, 1 << 30))))
        elif k == 8:
            ops.append(("forget", int(rng.integers(0, 1 << 30))))
        elif k == 9:
            ops.append(("score_cold", shape))
        elif k == 10:
            ops.append(("filter_subset", shape, int(rng.integers(0, 1 << 30))))
        else:
            ops.append(("rebind", shape, int(rng.integers(0, 1 << 30))))
    return nodes, ops


def run_scenario(backend, nodes, ops, policy):
    """Returns the trace (list of tuples) of everything observable."""
    trace = []
    ids = []
    gcount = {}
    for core_alloc, mem_alloc, rows in nodes:
        nid = backend.add_node(core_alloc, mem_alloc)
        trace.append(("node", nid))
        if nid >= 0:
            ids.append(nid)
            gcount[nid] = core_alloc // 100
            if rows is not None:
                backend.set_rows(nid, rows[0], rows[1])
    placed = []       # (uid, node, shape, alloc)
    for op in ops:
        kind = op[0]
        if kind == "sched":
            _, shape, uid = op
            fit = backend.filter(ids, shape)
            fit_ids = [n for n, f in zip(ids, fit) if f]
            st, sc = backend.score(fit_ids, shape)
            rec = ["sched", tuple(fit), st, tuple(sc)]
            if fit_ids:
                w = fit_ids[sc.index(max(sc))]
                rec.append(backend.peek(w, shape))
                bst, alloc = backend.bind(w, shape, uid)
                rec += [w, bst, alloc]
                if bst == 0:
                    placed.append((uid, w, shape, alloc))
            trace.append(tuple(map(_freeze, rec)))
        elif kind == "add_pod":
            _, shape, uid, node, r = op
            node = ids[node % len(ids)]
            rng = np.random.default_rng(r)
            alloc = []
            for u in shape:
                k = u[2] if u[2] > 0 else 1
                k = min(k, gcount[node])
                alloc.append([int(x) for x in rng.choice(gcount[node], size=k, replace=False)]
                             if rng.integers(0, 6) else [])
            backend.add_pod(node, shape, alloc, uid)
            placed.append((uid, node, shape, alloc))
            trace.append(("add_pod", node, uid, backend.known(uid)))
        elif kind == "forget":
            if placed:
                uid, node, shape, alloc = placed.pop(op[1] % len(placed))
                backend.forget_pod(node, shape, alloc, uid)
                trace.append(("forget", uid, backend.known(uid), backend.released(uid)))
        elif kind == "score_cold":
            st, sc = backend.score(ids, op[1])
            trace.append(("score_cold", st, tuple(sc)))
        elif kind == "filter_subset":
            rng = np.random.default_rng(op[2])
            sub = [int(x) for x in rng.permutation(ids)[: max(1, len(ids) // 2)]]
            trace.append(("filter_subset", tuple(sub), tuple(backend.filter(sub, op[1]))))
        elif kind == "rebind":
            # bind on an arbitrary node: entry may be missing (node.go:93-96) or the uid already known
            node = ids[op[2] % len(ids)]
            uid = placed[op[2] % len(placed)][0] if placed and (op[2] >> 8) % 2 else 5000 + (op[2] % 1000)
            st, alloc = backend.bind(node, op[1], uid)
            trace.append(("rebind", node, uid, st, _freeze(alloc)))
        trace.append(("rows", tuple(tuple(backend.rows(n)) for n in ids)))
    return trace


def _freeze(x):
    if isinstance(x, (list, tuple)):
        return tuple(_freeze(y) for y in x)
    if isinstance(x, np.generic):
        return x.item()
    return x
