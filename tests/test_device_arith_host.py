"""The kernels' integer arithmetic -- csrc/egs_device.cuh: the Trade fast path (prefix/suffix maxima,
unsigned-min PAD trick, packed q*8+g key), the general DFS Trade and Transact -- compiled for the host
(csrc/host_test/device_on_host.cu) and checked against the oracle WITHOUT a GPU.  Same source the kernels
inline; the device SASS is unaffected by the host build."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import egs_oracle as po

PAD = -(1 << 31)


@pytest.fixture(scope="module")
def DH():
    import egs_b200
    L = C.CDLL(egs_b200._build.build_devhost())
    L.egsdh_trade.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.egsdh_trade_leaves.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.egsdh_transact.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint32]
    L.egsdh_is_single.argtypes = [C.c_int, C.c_void_p]
    for f in ("egsdh_cand_key", "egsdh_fit_term", "egsdh_score_term"):
        getattr(L, f).restype = C.c_uint64
    L.egsdh_cand_key.argtypes = [C.c_int32, C.c_uint32]
    L.egsdh_fit_term.argtypes = [C.c_uint32]
    L.egsdh_score_term.argtypes = [C.c_uint32, C.c_int32]
    return L


def _pad(rows):
    core = np.full(8, PAD, np.int32); mem = np.full(8, PAD, np.int32)
    for g, (c, m) in enumerate(rows):
        core[g], mem[g] = c, m
    return core, mem


def _units(req):
    a = np.zeros((max(1, len(req)), 3), np.int32)
    for i, u in enumerate(req):
        a[i] = u
    return a


def _trade(DH, rows, mt, req, policy, path):
    core, mem = _pad(rows)
    u = _units(req)
    sc, mk = C.c_int32(0), C.c_uint32(0)
    ok = DH.egsdh_trade(core.ctypes.data, mem.ctypes.data, mt, len(req), u.ctypes.data, policy, path, C.byref(sc), C.byref(mk))
    if not ok:
        return None
    alloc = [[g for g in range(8) if (mk.value >> (8 * c + g)) & 1] for c in range(len(req))]
    return alloc, sc.value


unit = st.one_of(
    st.tuples(st.integers(0, 100), st.integers(0, 40), st.just(0)).filter(lambda u: u[0] or u[1]),
    st.tuples(st.just(0), st.just(0), st.integers(1, 3)),
    st.just((-1, -1, 0)),
)
rows_s = st.lists(st.tuples(st.integers(0, 101), st.integers(0, 41)), min_size=1, max_size=8)


@settings(max_examples=600, deadline=None)
@given(rows=rows_s, req=st.lists(unit, min_size=1, max_size=4), policy=st.integers(0, 1), mt=st.integers(1, 40))
def test_kernel_trade_equals_oracle(DH, rows, req, policy, mt):
    g = [po.GPU(c, m, 100, mt) for c, m in rows]
    opt = po.trade(g, po.RATERS[policy], list(req))
    want = None if opt is None else (opt.allocated, opt.score)
    assert _trade(DH, rows, mt, req, policy, 0) == want          # the dispatch the kernels use
    assert _trade(DH, rows, mt, req, policy, 1) == want          # general DFS on every request


@settings(max_examples=800, deadline=None)
@given(rows=rows_s, req=st.lists(unit, min_size=1, max_size=4), policy=st.integers(0, 1), mt=st.integers(1, 40))
def test_leaf_parallel_trade_equals_oracle(DH, rows, req, policy, mt):
    """trade_leaf_eval / trade_leaf_space (the Trade the resolver spreads over the lanes of a warp, one DFS leaf per
    lane): the maximal (score, leaf index) over all leaves is the option gpu.go:65-129 returns -- same Allocated,
    same Score, same 'last maximal leaf wins', including whole-GPU and sentinel containers and 3/5/6/7-GPU nodes."""
    g = [po.GPU(c, m, 100, mt) for c, m in rows]
    opt = po.trade(g, po.RATERS[policy], list(req))
    want = None if opt is None else (opt.allocated, opt.score)
    core, mem = _pad(rows)
    u = _units(req)
    sc, mk = C.c_int32(0), C.c_uint32(0)
    ok = DH.egsdh_trade_leaves(core.ctypes.data, mem.ctypes.data, mt, len(req), u.ctypes.data, policy, C.byref(sc), C.byref(mk))
    got = None if not ok else ([[gg for gg in range(8) if (mk.value >> (8 * c + gg)) & 1] for c in range(len(req))], sc.value)
    assert got == want


@settings(max_examples=600, deadline=None)
@given(rows=st.lists(st.tuples(st.integers(0, 100), st.sampled_from([0, 1, 1024, 40960, 81920, (1 << 25)])), min_size=1, max_size=8),
       core=st.integers(0, 99), mem=st.sampled_from([0, 1, 1024, 4096, 40960, 81920, (1 << 25) - 7]), policy=st.integers(0, 1))
def test_fast_path_single_container_large_values(DH, rows, core, mem, policy):
    """The fast path with realistic magnitudes (MB-scale memory up to the 2^25 guard): no int32 overflow,
    score == Range/2*100, last maximal GPU wins."""
    if core == 0 and mem == 0:
        mem = 1
    mt = 1 << 25
    req = [(core, mem, 0)]
    g = [po.GPU(c, m, 100, mt) for c, m in rows]
    opt = po.trade(g, po.RATERS[policy], req)
    want = None if opt is None else (opt.allocated, opt.score)
    u = _units(req)
    assert DH.egsdh_is_single(1, u.ctypes.data) == 1
    assert _trade(DH, rows, mt, req, policy, 0) == want


@settings(max_examples=400, deadline=None)
@given(rows=rows_s, req=st.lists(unit, min_size=1, max_size=4), mt=st.integers(1, 40), stale=st.lists(st.tuples(st.integers(0, 7), st.integers(0, 60), st.integers(0, 30)), max_size=3))
def test_kernel_transact_equals_oracle(DH, rows, req, mt, stale):
    """Transact with a possibly STALE option (rows changed after the option was made): same success/failure,
    same partial application without rollback (gpu.go:153-175)."""
    g = [po.GPU(c, m, 100, mt) for c, m in rows]
    opt = po.trade(g, po.rate_binpack, list(req))
    if opt is None:
        return
    for gi, dc, dm in stale:                                      # somebody else consumed resources meanwhile
        if gi < len(g):
            g[gi].core_avail = max(0, g[gi].core_avail - dc); g[gi].mem_avail = max(0, g[gi].mem_avail - dm)
    core, mem = _pad([(x.core_avail, x.mem_avail) for x in g])
    masks = 0
    for c, a in enumerate(opt.allocated):
        for gi in a:
            masks |= 1 << (8 * c + gi)
    ok = DH.egsdh_transact(core.ctypes.data, mem.ctypes.data, mt, len(req), _units(req).ctypes.data, masks)
    assert bool(ok) == po.transact(g, opt)
    assert [(int(core[i]), int(mem[i])) for i in range(len(g))] == [(x.core_avail, x.mem_avail) for x in g]


def test_keys_and_digest_terms(DH):
    for node, score in [(0, 0), (5, 600), (99999, 2050500), (2**31 - 1, 0)]:
        assert DH.egsdh_fit_term(node) == po.fit_digest_term(node)
        assert DH.egsdh_score_term(node, score) == po.score_digest_term(node, score)
    # ordering: higher score first, then LOWER node id
    k = DH.egsdh_cand_key
    assert k(600, 7) > k(500, 0) and k(600, 3) > k(600, 4) and k(0, 0) > 0
