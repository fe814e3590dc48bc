"""The extender JSON codec (csrc/host/extender_json.cc): wire format of /scheduler/filter, /priorities, /bind
(pkg/routes/routes.go:39-163).  CPU only."""
import ctypes as C
import json

import numpy as np
import pytest


@pytest.fixture(scope="module")
def J():
    import egs_b200
    L = C.CDLL(egs_b200._build.build_host())
    L.egsj_new.restype = C.c_void_p
    for f in ("egsj_parse_args", "egsj_pod_dump", "egsj_parse_binding", "egsj_encode_filter", "egsj_encode_priorities",
              "egsj_encode_binding", "egsj_node_name", "egsj_encode_filter_error"):
        getattr(L, f).restype = C.c_char_p
    L.egsj_parse_args.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.egsj_parse_binding.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.egsj_pod_dump.argtypes = [C.c_void_p]
    L.egsj_node_ids.restype = C.POINTER(C.c_int32); L.egsj_node_ids.argtypes = [C.c_void_p]
    for f in ("egsj_has_nodenames", "egsj_n_nodes", "egsj_interned"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.egsj_node_name.argtypes = [C.c_void_p, C.c_int]
    L.egsj_quantity.argtypes = [C.c_char_p, C.POINTER(C.c_int64)]
    L.egsj_encode_filter.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
    L.egsj_encode_filter_error.argtypes = [C.c_void_p, C.c_char_p]
    L.egsj_encode_priorities.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
    L.egsj_encode_binding.argtypes = [C.c_void_p, C.c_char_p]
    L.egsj_free.argtypes = [C.c_void_p]
    return L


def _pod(containers, **meta):
    return {"metadata": dict({"name": "p", "namespace": "default", "uid": "u-1"}, **meta),
            "spec": {"containers": containers, "schedulerName": "default-scheduler"}, "status": {"phase": "Pending"}}


def test_quantity_value_rounds_up(J):
    cases = {"100": 100, "4": 4, "0": 0, "1Gi": 1 << 30, "1500m": 2, "500m": 1, "1.5": 2, "2e3": 2000, "1k": 1000,
             "12Mi": 12 << 20, "0.1": 1, "100m": 1, "1000m": 1, "1001m": 2, "3.0": 3}
    for q, want in cases.items():
        v = C.c_int64()
        assert J.egsj_quantity(q.encode(), C.byref(v)) == 1 and v.value == want, q
    for bad in ["", "abc", "1Qi", "1..2"]:
        assert J.egsj_quantity(bad.encode(), C.byref(C.c_int64())) == 0


def test_parse_filter_request_and_intern_nodes(J):
    ctx = J.egsj_new()
    names = [f"node-{i:06d}" for i in range(2000)]
    pod = _pod([{"name": "main", "resources": {"requests": {"elasticgpu.io/gpu-core": "50", "elasticgpu.io/gpu-memory": "4",
                                                                "cpu": "500m"}, "limits": {"elasticgpu.io/gpu-core": "50"}}},
                {"name": "side", "resources": {}},
                {"name": "big", "resources": {"requests": {"elasticgpu.io/gpu-core": 200}}}],
               annotations={"elasticgpu.io/container-main": "1", "note": "a \"quoted\" \\ value é"})
    body = json.dumps({"pod": pod, "nodenames": names, "nodes": None}).encode()
    assert J.egsj_parse_args(ctx, body, len(body)) == b""
    assert J.egsj_has_nodenames(ctx) == 1 and J.egsj_n_nodes(ctx) == 2000 and J.egsj_interned(ctx) == 2000
    ids = np.ctypeslib.as_array(J.egsj_node_ids(ctx), shape=(2000,)).copy()
    assert list(ids) == list(range(2000)) and J.egsj_node_name(ctx, 1234) == b"node-001234"
    lines = J.egsj_pod_dump(ctx).decode().split("\n")
    assert lines[0] == "default\tp\tu-1\t"
    assert lines[1:4] == ["C\tmain\t1\t50\t1\t4", "C\tside\t0\t0\t0\t0", "C\tbig\t1\t200\t0\t0"]
    assert "A\telasticgpu.io/container-main\t1" in lines
    assert 'A\tnote\ta "quoted" \\ value é' in lines
    # a second request re-uses the ids (any order, partial list)
    body2 = json.dumps({"pod": pod, "nodenames": ["node-000007", "new-node", "node-000003"]}).encode()
    assert J.egsj_parse_args(ctx, body2, len(body2)) == b""
    ids2 = np.ctypeslib.as_array(J.egsj_node_ids(ctx), shape=(3,)).copy()
    assert list(ids2) == [7, 2000, 3] and J.egsj_interned(ctx) == 2001
    # nodeCacheCapable=false: no nodenames -> the route answers with an error (routes.go:59-64)
    body3 = json.dumps({"pod": pod, "nodes": {"items": []}}).encode()
    assert J.egsj_parse_args(ctx, body3, len(body3)) == b"" and J.egsj_has_nodenames(ctx) == 0
    assert J.egsj_parse_args(ctx, b'{"pod": {"metadata": ', 21) != b""
    J.egsj_free(ctx)


def test_encoders_match_go_encoding_json(J):
    ctx = J.egsj_new()
    out = J.egsj_encode_filter(ctx, b"n1\nn3", b"n2\tno enough resource to allocate\nghost\telastic gpu scheduler get node failed: nodes \"ghost\" not found", b"")
    assert out == (b'{"nodenames":["n1","n3"],"failedNodes":{"ghost":"elastic gpu scheduler get node failed: nodes \\"ghost\\" not found",'
                   b'"n2":"no enough resource to allocate"}}')
    assert json.loads(out)["failedNodes"]["n2"] == "no enough resource to allocate"
    assert J.egsj_encode_filter(ctx, b"", b"", b"") == b'{"nodenames":[]}'                      # omitempty map + error
    assert J.egsj_encode_filter(ctx, b"", b"", b"a<b & c") == b'{"nodenames":[],"error":"a\\u003cb \\u0026 c"}'  # HTML-safe
    sc = np.array([600, 0, 2050500], np.int64)
    assert J.egsj_encode_priorities(ctx, b"n1\nn2\nn3", sc.ctypes.data, 3) == \
        b'[{"host":"n1","score":600},{"host":"n2","score":0},{"host":"n3","score":2050500}]'
    assert J.egsj_encode_binding(ctx, b"") == b"{}"
    assert J.egsj_encode_binding(ctx, b"cannot find option") == b'{"error":"cannot find option"}'
    b = json.dumps({"podName": "p", "podNamespace": "ns", "podUID": "u", "node": "n9"}).encode()
    assert J.egsj_parse_binding(ctx, b, len(b)) == b"\tp\tns\tu\tn9"
    J.egsj_free(ctx)


def test_large_filter_request_throughput(J):
    """10^5 candidate names (~1.5 MB, what config 4 would put in every verb): parse + intern in well under a second."""
    import time
    ctx = J.egsj_new()
    names = [f"node-{i:06d}" for i in range(100000)]
    body = json.dumps({"pod": _pod([{"name": "c", "resources": {"requests": {"elasticgpu.io/gpu-core": "25"}}}]),
                       "nodenames": names}).encode()
    J.egsj_parse_args(ctx, body, len(body))                       # first: interning
    t = time.perf_counter()
    for _ in range(5):
        assert J.egsj_parse_args(ctx, body, len(body)) == b""
    dt = (time.perf_counter() - t) / 5
    assert J.egsj_n_nodes(ctx) == 100000 and dt < 0.5, dt
    J.egsj_free(ctx)


def test_filter_error_paths_omit_nodenames(J):
    """predicate.go:21-31 / routes.go:51-64 return ExtenderFilterResult{Error: ...} with a nil NodeNames pointer:
    encoding/json omits the member; the normal path keeps "nodenames":[] even when nothing fits."""
    ctx = J.egsj_new()
    msg = "elastic-gpu-scheduler extender must be configured with nodeCacheCapable=true"
    assert J.egsj_encode_filter_error(ctx, msg.encode()) == json.dumps({"error": msg}, separators=(",", ":")).encode()
    assert J.egsj_encode_filter(ctx, b"", b"n1\tno enough resource to allocate", b"") == \
        b'{"nodenames":[],"failedNodes":{"n1":"no enough resource to allocate"}}'
    J.egsj_free(ctx)


def test_parser_hardening(J):
    ctx = J.egsj_new()
    deep = b'{"pod":{"metadata":{"x":' + b"[" * 20000 + b"]" * 20000 + b'}},"nodenames":[]}'
    assert b"max depth" in J.egsj_parse_args(ctx, deep, len(deep))                  # encoding/json: "exceeded max depth"
    ok = b'{"pod":{"metadata":{"x":' + b"[" * 500 + b"]" * 500 + b',"name":"p"}},"nodenames":["a"]}'
    assert J.egsj_parse_args(ctx, ok, len(ok)) == b""
    bad = b'{"pod":{"metadata":{"name":"a\\u12G4"}},"nodenames":[]}'
    assert J.egsj_parse_args(ctx, bad, len(bad)) != b""                             # non-hex digit in \\u escape
    pair = json.dumps({"pod": {"metadata": {"name": "p", "annotations": {"k": "\U0001F600 ok"}}}, "nodenames": []}).encode()
    assert b"\\ud83d\\ude00" in pair                                              # json.dumps emits a surrogate pair
    assert J.egsj_parse_args(ctx, pair, len(pair)) == b""
    assert "\U0001F600 ok".encode() in J.egsj_pod_dump(ctx)
    lone = b'{"pod":{"metadata":{"name":"p","annotations":{"k":"\\ud83d!"}}},"nodenames":[]}'
    assert J.egsj_parse_args(ctx, lone, len(lone)) == b"" and "\ufffd!".encode() in J.egsj_pod_dump(ctx)
    J.egsj_free(ctx)
