"""Property tests (hypothesis) of the oracle pair: python mirror == C restatement on arbitrary rows and
requests, plus invariants that follow from the reference's code."""
from hypothesis import given, settings, strategies as st

import egs_oracle as po
import oracle_c as oc

unit = st.one_of(
    st.tuples(st.integers(0, 100), st.integers(0, 40), st.just(0)).filter(lambda u: u[0] or u[1]),
    st.tuples(st.just(0), st.just(0), st.integers(1, 3)),
    st.just((-1, -1, 0)),
)
rows = st.lists(st.tuples(st.integers(0, 101), st.integers(0, 41)), min_size=1, max_size=8)


@settings(max_examples=300, deadline=None)
@given(rows=rows, req=st.lists(unit, min_size=1, max_size=4), policy=st.integers(0, 1), mt=st.integers(1, 40))
def test_trade_python_equals_c(rows, req, policy, mt):
    g = [po.GPU(c, m, 100, mt) for c, m in rows]
    opt = po.trade(g, po.RATERS[policy], list(req))
    o = oc.OracleC(policy)
    n = o.add_node(100 * len(rows), mt * len(rows))
    o.set_rows(n, [r[0] for r in rows], [r[1] for r in rows])
    got = o.trade(n, req)
    assert (None if opt is None else (opt.allocated, opt.score)) == got
    # Trade restores the rows it mutates while searching (gpu.go:116-120)
    assert [(x.core_avail, x.mem_avail) for x in g] == list(rows) and o.rows(n) == list(rows)
    if opt is not None:
        assert opt.score >= 0 and (policy == 0 or opt.score == 0)        # rater.go:49-50 / :56-59
        assert opt.score % 100 == 0
        for u, a in zip(req, opt.allocated):
            assert len(a) == (u[2] if u[2] > 0 else 1) and a == sorted(a)   # ascending lists -> a GPU mask is lossless


@settings(max_examples=200, deadline=None)
@given(rows=rows, core=st.integers(0, 99), mem=st.integers(0, 40), mt=st.integers(1, 40))
def test_single_container_last_max_wins(rows, core, mem, mt):
    """One fractional container: the chosen GPU is the LAST one among those with the maximal Rate (gpu.go:85)."""
    if core == 0 and mem == 0:
        mem = 1
    g = [po.GPU(c, m, 100, mt) for c, m in rows]
    opt = po.trade(g, po.rate_binpack, [(core, mem, 0)])
    scores = []
    for i, x in enumerate(g):
        if x.can_allocate((core, mem, 0)):
            x.add((core, mem, 0)); scores.append((po.rate_binpack(g, [i]), i)); x.sub((core, mem, 0))
    if not scores:
        assert opt is None
    else:
        best = max(s for s, _ in scores)
        assert opt.score == best and opt.allocated == [[max(i for s, i in scores if s == best)]]


@settings(max_examples=100, deadline=None)
@given(rows=rows, req=st.lists(unit, min_size=1, max_size=3), mt=st.integers(1, 40))
def test_transact_then_cancel_roundtrip(rows, req, mt):
    """Transact of a fresh option succeeds and Cancel undoes it -- except for whole-GPU units, whose Sub
    resets to the totals (gpu.go:42-44), and sentinel units, which ADD on Add (gpu.go:36-37)."""
    g = [po.GPU(c, m, 100, mt) for c, m in rows]
    opt = po.trade(g, po.rate_binpack, list(req))
    if opt is None:
        return
    before = [(x.core_avail, x.mem_avail) for x in g]
    assert po.transact(g, opt)
    po.cancel(g, opt)
    after = [(x.core_avail, x.mem_avail) for x in g]
    whole = {i for u, a in zip(req, opt.allocated) if u[2] > 0 for i in a}
    for i, (b, a2) in enumerate(zip(before, after)):
        if i in whole:
            assert a2 == (100, mt)
        else:
            assert a2 == b
