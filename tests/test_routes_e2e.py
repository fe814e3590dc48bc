"""Extender routes end to end: an ExtenderArgs / ExtenderBindingArgs HTTP body goes through the JSON
decoder, the plugin verbs (libegs on the GPU) and the Go-byte-exact encoders; the response bodies must equal what
encoding/json would emit for the results the ORACLE computes on the same cluster (pkg/routes/routes.go:39-163,
pkg/server/{predicate,priority,bind}.go)."""
import json

import pytest

import oracle_c as oc

pytestmark = pytest.mark.gpu

CORE, MEM = "elasticgpu.io/gpu-core", "elasticgpu.io/gpu-memory"
NOFIT = "no enough resource to allocate"


def _go(obj) -> str:
    """encoding/json of plain ASCII data: no spaces; map keys sorted (Go sorts map keys; struct order is given)."""
    return json.dumps(obj, separators=(",", ":"))


def _pod_json(name, containers, uid):
    return {"metadata": {"name": name, "namespace": "default", "uid": uid},
            "spec": {"containers": [{"name": n, "resources": {"requests": r, "limits": r}} for n, r in containers]},
            "status": {"phase": "Pending"}}


def test_filter_priorities_bind_bodies_match_the_oracle():
    import egs_b200.host as H
    N = 300
    names = [f"node-{i:06d}" for i in range(N)]
    s = H.CudaUnitScheduler(0, max_nodes=1024)
    o = oc.OracleC(0)
    for i, n in enumerate(names):
        core_alloc, mem_alloc = 100 * (1 + i % 8), 16 * (1 + i % 8) + (i % 3)
        s.register_node(n, core_alloc, mem_alloc)
        o.add_node(core_alloc, mem_alloc)
    R = H.ExtenderRoutes(s)
    shapes = [[("main", {CORE: "30", MEM: "8"})],
              [("main", {CORE: "50", MEM: "12"}), ("side", {})],                      # sidecar without GPU request
              [("a", {CORE: "200"})],                                                 # two whole GPUs
              [("a", {MEM: "16"}), ("b", {CORE: "20", MEM: "1"})]]
    for k in range(24):
        cont = shapes[k % len(shapes)]
        req = []
        for _, r in cont:
            c, m = int(r.get(CORE, 0)), int(r.get(MEM, 0))
            req.append(oc.unit_from_requests(c, m))
        pj = _pod_json(f"pod-{k}", cont, f"uid-{k}")
        body = _go({"pod": pj, "nodes": None, "nodenames": names}).encode()
        # ---- /scheduler/filter
        fit = o.filter(None, req)
        filtered = [n for n, f in zip(names, fit) if f]
        want = {"nodenames": filtered}
        failed = {n: NOFIT for n, f in zip(names, fit) if not f}
        if failed:
            want["failedNodes"] = dict(sorted(failed.items()))
        st, got = R.filter(body)
        assert st == 200 and got == _go(want), f"filter body differs at pod {k}"
        if not filtered:
            continue
        # ---- /scheduler/priorities (kube-scheduler sends the filtered names)
        body2 = _go({"pod": pj, "nodenames": filtered}).encode()
        ost, sc = o.score([names.index(n) for n in filtered], req)
        assert ost == 0
        st, got = R.priorities(body2)
        assert st == 200 and got == _go([{"host": n, "score": int(x)} for n, x in zip(filtered, sc)])
        # ---- /scheduler/bind on the first node with the maximum score
        w = filtered[list(sc).index(max(sc))]
        pod = H.Pod(f"pod-{k}", [(n, {kk: int(v) for kk, v in (("core", r.get(CORE)), ("memory", r.get(MEM))) if v is not None})
                                 for n, r in cont], uid=f"uid-{k}")
        R.register_pod(pod)
        bst, _ = o.bind(names.index(w), req, k)
        st, got = R.bind(_go({"podName": f"pod-{k}", "podNamespace": "default", "podUID": f"uid-{k}", "node": w}).encode())
        if bst == 0:
            assert (st, got) == (200, "{}")
        else:
            assert st == 500 and json.loads(got)["error"]
    # final rows equal
    status = json.loads(s.Status())
    for i, n in enumerate(names):
        if n in status:
            assert [(g["CoreAvailable"], g["MemoryAvailable"]) for g in status[n]] == o.rows(i)


def test_error_bodies():
    import egs_b200.host as H
    s = H.CudaUnitScheduler(0, max_nodes=16)
    s.register_node("n0", 200, 32)
    R = H.ExtenderRoutes(s)
    pj = _pod_json("p", [("main", {CORE: "10"})], "u")
    # nodeCacheCapable=false: kube-scheduler sends "nodes", no "nodenames" (routes.go:59-64)
    st, got = R.filter(_go({"pod": pj, "nodes": {"items": []}}).encode())
    assert (st, got) == (200, '{"error":"elastic-gpu-scheduler extender must be configured with nodeCacheCapable=true"}')
    # undecodable body: error member only (routes.go:51-58)
    st, got = R.filter(b'{"pod": ')
    assert st == 200 and set(json.loads(got)) == {"error"}
    # no container asks for a managed resource (predicate.go:19-24)
    st, got = R.filter(_go({"pod": _pod_json("q", [("main", {"cpu": "1"})], "u2"), "nodenames": ["n0"]}).encode())
    assert (st, got) == (200, '{"error":"cannot find scheduler for pod default/q"}')
    # unknown node: per-node message, nodenames present and empty (scheduler.go:124,158-167)
    st, got = R.filter(_go({"pod": pj, "nodenames": ["ghost"]}).encode())
    assert st == 200 and json.loads(got) == {"nodenames": [], "failedNodes": {"ghost": 'elastic gpu scheduler get node failed: nodes "ghost" not found'}}
    assert got.startswith('{"nodenames":[],"failedNodes":{')
    # bind of a pod the apiserver does not know / without a cached option: HTTP 500 with the error (routes.go:146-158)
    st, got = R.bind(_go({"podName": "nope", "podNamespace": "default", "podUID": "x", "node": "n0"}).encode())
    assert st == 500 and json.loads(got) == {"error": 'pods "nope" not found'}
    pod = H.Pod("p", [("main", {"core": 10})], uid="u")
    R.register_pod(pod)
    st, got = R.bind(_go({"podName": "p", "podNamespace": "default", "podUID": "u", "node": "n0"}).encode())
    assert st == 500 and json.loads(got)["error"].startswith("cannot find option of GPU request")
    # priorities without nodenames: the reference dereferences nil and panics (priority.go:19)
    st, got = R.priorities(_go({"pod": pj}).encode())
    assert st == -1 and got.startswith("panic:")
