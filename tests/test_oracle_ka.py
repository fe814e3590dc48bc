"""Pins both oracles (python mirror, C restatement) to the SURVEY 8c known-answer vectors."""
import hashlib

import numpy as np
import pytest

import egs_oracle as po
import oracle_c as oc
from ka_vectors import HASH_KA, KA0_FINAL, KA0_MEM, KA0_TRACE, TRADE_KA, UNIT_KA


@pytest.mark.parametrize("name,policy,mt,rows,req,exp", TRADE_KA, ids=[k[0] for k in TRADE_KA])
def test_trade_ka_python(name, policy, mt, rows, req, exp):
    g = [po.GPU(c, m, 100, mt) for c, m in rows]
    opt = po.trade(g, po.RATERS[policy], list(req))
    assert (None if opt is None else (opt.allocated, opt.score)) == exp
    assert [(x.core_avail, x.mem_avail) for x in g] == list(rows), "Trade must restore the rows"


@pytest.mark.parametrize("name,policy,mt,rows,req,exp", TRADE_KA, ids=[k[0] for k in TRADE_KA])
def test_trade_ka_c(name, policy, mt, rows, req, exp):
    o = oc.OracleC(policy)
    n = o.add_node(100 * len(rows), mt * len(rows))
    o.set_rows(n, [r[0] for r in rows], [r[1] for r in rows])
    assert o.trade(n, req) == exp
    assert o.rows(n) == list(rows)


@pytest.mark.parametrize("req,hexd", HASH_KA)
def test_request_hash(req, hexd):
    assert po.request_hash(req) == hexd
    assert oc.request_hash(req) == hexd


def test_sha256_matches_hashlib():
    rng = np.random.default_rng(7)
    for n in [0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 200, 1000]:
        msg = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        assert oc.sha256(msg) == hashlib.sha256(msg).digest()


@pytest.mark.parametrize("inp,unit", UNIT_KA)
def test_new_gpu_request(inp, unit):
    assert po.new_gpu_request([inp]) == [unit]
    assert oc.unit_from_requests(*inp) == unit


def test_ka_t_allocate_without_assume():
    """scheduler_test.go:11-24: node 400/48 -> 4 x (100,12); Allocate without Assume errors, rows untouched."""
    s = po.Scheduler(po.POLICY_SPREAD)
    n = s.add_node(400, 48)
    st, alloc = s.bind(n, [(0, 4, 0)], 1)
    assert st == po.EGS_ERR_NO_OPTION and alloc is None
    assert s.rows(n) == [(100, 12)] * 4
    o = oc.OracleC(1)
    n = o.add_node(400, 48)
    st, alloc = o.bind(n, [(0, 4, 0)], 1)
    assert st == 2 and alloc is None and o.rows(n) == [(100, 12)] * 4


def test_no_gpu_node():
    assert po.Scheduler(0).add_node(99, 48) == -1
    assert oc.OracleC(0).add_node(99, 48) == -1


def _ka0(make):
    s = make()
    for _ in range(4):
        s.add_node(200, 32)
    return s


def test_ka0_python():
    s = _ka0(lambda: po.Scheduler(po.POLICY_BINPACK))
    for uid, (m, exp) in enumerate(zip(KA0_MEM, KA0_TRACE)):
        r = s.schedule_one(po.new_gpu_request([(0, m)]), uid)
        assert (r["fit"], r["scores"], r["node"], r["status"], r["alloc"]) == exp
    assert [s.rows(i) for i in range(4)] == KA0_FINAL


@pytest.mark.parametrize("faithful", [False, True])
@pytest.mark.parametrize("threads", [1, 4])
def test_ka0_c(faithful, threads):
    o = _ka0(lambda: oc.OracleC(0, faithful))
    units = np.array([[0, m, 0] for m in KA0_MEM], np.int64)
    c_off = np.arange(9, dtype=np.int32)
    out = o.schedule_batch(c_off, units, threads=threads, vec_pods=8)
    for p, (fit, scores, node, status, alloc) in enumerate(KA0_TRACE):
        assert list(out["vec_fit"][p]) == fit
        assert [int(out["vec_score"][p][i]) for i in range(4) if fit[i]] == scores
        assert out["node"][p] == node and out["status"][p] == status
        assert out["fit_count"][p] == sum(fit)
        if alloc is not None:
            assert [g for g in range(8) if out["alloc_mask"][p][0] >> g & 1] == alloc[0]
        else:
            assert not out["alloc_mask"][p].any()
    assert [o.rows(i) for i in range(4)] == KA0_FINAL
    # uid of the failed bind stays in n1's podsMap (node.go:150) but not in podMaps
    assert not o.known_pod(7) and o.known_pod(6)
