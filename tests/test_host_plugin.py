"""The C++ mirror of the reference's ResourceScheduler plugin (csrc/host) driven like the reference's
own test (pkg/scheduler/scheduler_test.go) -- nodes with allocatable gpu-core/gpu-memory, pods with
container requests -- and checked against the oracle's pod-level semantics."""
import json

import pytest

import egs_oracle as po
from ka_vectors import KA0_FINAL, KA0_MEM, KA0_TRACE

pytestmark = pytest.mark.gpu

NOFIT = "no enough resource to allocate"


def _host():
    import egs_b200.host as host
    return host


def test_reference_test_assume_shape():
    """scheduler_test.go:11-24: node 400/48, Spread rater, pod {gpu-core 0, gpu-memory 4};
    Allocate (Bind) without a prior Assume errors and leaves the 4 x (100,12) rows untouched."""
    H = _host()
    s = H.CudaUnitScheduler(po.POLICY_SPREAD)
    s.register_node("node-0", 400, 48)
    pod = H.Pod("test-pod--0", [("c0", {"core": 0, "memory": 4})])
    err = s.Bind("node-0", pod)
    rows = '[' + ','.join(['{"CoreAvailable":100,"MemoryAvailable":12,"CoreTotal":100,"MemoryTotal":12}'] * 4) + ']'
    assert err == "cannot find option of GPU request (core: 0, memory: 4, gpu count: 0) on " + rows
    assert json.loads(s.Status()) == {"node-0": json.loads(rows)}
    # with the Assume first it binds to the last GPU (spread == last feasible option)
    assert s.Assume(["node-0"], pod) == (["node-0"], {}, None)
    assert s.Bind("node-0", pod) is None
    ann, lab = s.pod_meta(pod)
    assert ann == {"elasticgpu.io/container-c0": "3", "elasticgpu.io/assumed": "true"}
    assert lab == {"elasticgpu.io/assumed": "true"}
    assert s.KnownPod(pod)


def test_ka0_through_the_plugin_interface():
    H = _host()
    s = H.CudaUnitScheduler(po.POLICY_BINPACK)
    names = [f"node-{i:06d}" for i in range(4)]
    for n in names:
        s.register_node(n, 200, 32)
    for i, (m, (fit, scores, node, status, alloc)) in enumerate(zip(KA0_MEM, KA0_TRACE)):
        pod = H.Pod(f"pod-{i}", [("main", {"memory": m})])
        assert H.CudaUnitScheduler.handles(pod)
        filtered, failed, err = s.Assume(names, pod)
        assert err is None
        assert filtered == [n for n, f in zip(names, fit) if f]
        assert failed == {n: NOFIT for n, f in zip(names, fit) if not f}
        sc = s.Score(filtered, pod)
        assert sc == scores
        w = filtered[sc.index(max(sc))]
        assert w == names[node]
        e = s.Bind(w, pod)
        if status == 0:
            assert e is None
            assert s.pod_meta(pod)[0][f"elasticgpu.io/container-main"] == ",".join(map(str, alloc[0]))
        else:
            assert e.startswith("can't trade option &{Request:(core: 0, memory: 8, gpu count: 0) Allocated:[[1]] Score:200} on [")
            assert e.endswith("because the GPU's residual memory or core can't satisfy the container")
            assert not s.KnownPod(pod)
    st = json.loads(s.Status())
    assert [[(g["CoreAvailable"], g["MemoryAvailable"]) for g in st[n]] for n in names] == KA0_FINAL


def test_unknown_node_and_no_gpu_node_messages():
    H = _host()
    s = H.CudaUnitScheduler(po.POLICY_BINPACK)
    s.register_node("good", 200, 32)
    s.register_node("cpu-only", 50, 0)
    pod = H.Pod("p", [("c", {"core": 20, "memory": 4})])
    filtered, failed, err = s.Assume(["ghost", "good", "cpu-only"], pod)
    assert err is None and filtered == ["good"]
    assert failed == {"ghost": 'elastic gpu scheduler get node failed: nodes "ghost" not found',
                      "cpu-only": "elastic gpu scheduler get node failed: no gpu available on node cpu-only"}
    assert s.Score(["ghost", "good"], pod)[0] == 0          # ScoreMin for a node that cannot be loaded


def test_get_resource_scheduler_rule():
    H = _host()
    assert not H.CudaUnitScheduler.handles(H.Pod("plain", [("c", {})]))
    assert H.CudaUnitScheduler.handles(H.Pod("sidecar", [("side", {}), ("gpu", {"core": 10})]))


def test_add_forget_and_restart_replay():
    """AddPod / ForgetPod with the option rebuilt from annotations (allocate.go:75-93), and the replay of
    assumed pods when a node is first loaded (node.go:52-54)."""
    H = _host()
    s = H.CudaUnitScheduler(po.POLICY_BINPACK)
    s.register_node("n0", 400, 64)
    running = H.Pod("running", [("a", {"core": 30, "memory": 4}), ("b", {"core": 100})], node_name="n0",
                    annotations={"elasticgpu.io/container-a": "2", "elasticgpu.io/container-b": "0"})
    s.register_assumed_pod("n0", running)                     # the apiserver lists it for n0
    probe = H.Pod("probe", [("c", {"core": 10, "memory": 1})])
    s.Assume(["n0"], probe)                                   # loads n0 -> replays `running`
    rows = [(g["CoreAvailable"], g["MemoryAvailable"]) for g in json.loads(s.Status())["n0"]]
    assert rows == [(0, 0), (100, 16), (70, 12), (100, 16)]
    assert s.AddPod(running) is None                          # informer delivers it too: node podsMap already has it
    rows2 = [(g["CoreAvailable"], g["MemoryAvailable"]) for g in json.loads(s.Status())["n0"]]
    assert rows2 == rows and s.KnownPod(running)
    assert s.ForgetPod(running) is None
    rows3 = [(g["CoreAvailable"], g["MemoryAvailable"]) for g in json.loads(s.Status())["n0"]]
    assert rows3 == [(100, 16)] * 4 and s.ReleasedPod(running) and not s.KnownPod(running)
    nameless = H.Pod("nameless", [("c", {"core": 10})])
    assert s.AddPod(nameless) == "pod default/nameless nodename is empty"


def test_plugin_matches_oracle_on_random_pods():
    import numpy as np
    H = _host()
    for policy in (0, 1):
        rng = np.random.default_rng(policy)
        s = H.CudaUnitScheduler(policy)
        o = po.Scheduler(policy)
        names = []
        for i in range(12):
            g = int(rng.choice([1, 2, 4, 8]))
            core, mem = 100 * g + int(rng.integers(0, 99)), g * int(rng.choice([16, 24, 80]))
            names.append(f"n{i}")
            s.register_node(names[-1], core, mem)
            assert o.add_node(core, mem) == i
        for k in range(300):
            nc = int(rng.integers(1, 4))
            reqs = []
            for _ in range(nc):
                t = rng.integers(0, 8)
                reqs.append({} if t == 0 else {"core": int(rng.choice([100, 200]))} if t == 1 else
                            {"core": int(rng.choice([0, 10, 25, 50])), "memory": int(rng.integers(0, 12))})
            if not any(reqs):
                continue
            pod = H.Pod(f"p{k}", [(f"c{j}", r) for j, r in enumerate(reqs)])
            req = po.new_gpu_request([(r.get("core", 0), r.get("memory", 0)) for r in reqs])
            filtered, failed, err = s.Assume(names, pod)
            fit = o.assume(range(12), req)
            assert filtered == [n for n, f in zip(names, fit) if f] and set(failed) == {n for n, f in zip(names, fit) if not f}
            if not filtered:
                continue
            ids = [i for i in range(12) if fit[i]]
            assert s.Score(filtered, pod) == o.score(ids, req)
            sc = o.score(ids, req)
            w = ids[sc.index(max(sc))]
            st, alloc = o.bind(w, req, k)
            e = s.Bind(names[w], pod)
            assert (e is None) == (st == 0)
            if st == 0:
                ann = s.pod_meta(pod)[0]
                assert [ann[f"elasticgpu.io/container-c{j}"] for j in range(nc)] == [",".join(map(str, a)) for a in alloc]
        status = json.loads(s.Status())
        for i, n in enumerate(names):
            assert [(g["CoreAvailable"], g["MemoryAvailable"]) for g in status[n]] == o.rows(i)
