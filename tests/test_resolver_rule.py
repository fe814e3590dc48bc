"""Where a round stops, and whether a list head or a tracked option wins, must be a function of the SERIAL state alone.

k_resolve_mw's owner warps maintain the heads of their untracked candidate lists lazily, outside the ticket
(csrc/egs_rounds.cuh: maintain_heads), so inside the ticket the head `bh` and the dry-list bound `dbound` may be STALE:
computed against an older, smaller tracked set.  The ticket of a fast pod decides from those stale values whenever they
suffice and re-validates otherwise (egs_rounds.cuh, "Exact rule" in the fast pod).  This file restates that decision
procedure and the exact rule in a few lines of Python and lets hypothesis search for a state -- lists, truncation flags,
tracked set then and now, best tracked key -- in which they differ.  There is none; replicated resolvers of a sharded run
therefore stay in lock step whatever the timing of their warps (the 8-rank parity failure of an
earlier version was exactly a violation of this property)."""
from hypothesis import given, settings, strategies as st


def maintain(lists, more, cur, tracked):
    """maintain_heads: advance every list's cursor past tracked nodes.  Returns (bh, bh_d, dbound); cur is updated."""
    bh, bh_d, dbound = 0, 0, 0
    for d, l in enumerate(lists):
        c = cur[d]
        while c < len(l) and l[c] in tracked:
            c += 1
        cur[d] = c
        if c < len(l):
            if l[c] > bh:
                bh, bh_d = l[c], d
        elif more[d] and l:
            dbound = max(dbound, l[-1])
    return bh, bh_d, dbound


def exact_rule(lists, more, tracked_now, best):
    """('stop' | 'head' | 'tracked' | 'nofit', winning key) from exact heads."""
    head, _, dbound = maintain(lists, more, [0] * len(lists), tracked_now)
    win = max(head, best)
    if dbound > win:
        return "stop", 0
    if head > best:
        return "head", head
    return ("tracked", best) if best else ("nofit", 0)


def device_fast_pod(lists, more, tracked_then, tracked_now, best):
    """The ticket of a fast pod: stale (hd, dbp) from the preparation, exact `best`, exact tracked set."""
    cur = [0] * len(lists)
    hd, _, dbp = maintain(lists, more, cur, tracked_then)          # preparation, outside the ticket
    reason = 0
    if max(hd, dbp) > best:
        if hd <= best:
            reason = 3
        elif hd not in tracked_now:                                  # hset_has(key_node(hd)): one probe
            if dbp > hd:
                reason = 3
        else:                                                        # the head went stale: re-validate inside the ticket
            hd, _, db = maintain(lists, more, cur, tracked_now)
            if db > max(hd, best):
                reason = 3
    if reason:
        return "stop", 0
    if hd > best:
        return "head", hd
    return ("tracked", best) if best else ("nofit", 0)


@st.composite
def states(draw):
    D = draw(st.integers(1, 4))
    keys = draw(st.lists(st.integers(1, 40), unique=True, max_size=14))
    lists = [[] for _ in range(D)]
    for k in keys:
        lists[draw(st.integers(0, D - 1))].append(2 * k)             # list keys even ...
    for l in lists:
        l.sort(reverse=True)
    more = [draw(st.booleans()) for _ in range(D)]
    universe = [2 * k for k in range(1, 41)]
    tracked_now = set(draw(st.lists(st.sampled_from(universe), unique=True, max_size=20)))
    tracked_then = set(x for x in tracked_now if draw(st.booleans()))
    best = draw(st.one_of(st.just(0), st.integers(0, 41).map(lambda x: 2 * x + 1)))   # ... tracked keys odd: never equal
    return lists, more, tracked_then, tracked_now, best


@settings(max_examples=3000, deadline=None)
@given(states())
def test_fast_pod_decision_is_timing_independent(sx):
    lists, more, tracked_then, tracked_now, best = sx
    assert device_fast_pod(lists, more, tracked_then, tracked_now, best) == exact_rule(lists, more, tracked_now, best)


@settings(max_examples=1500, deadline=None)
@given(states())
def test_stale_values_bound_every_untracked_key(sx):
    """U = max(bh, dbound) taken at ANY earlier time bounds every key an untracked candidate can have now -- the shown
    ones, and (for truncated lists) the hidden ones, which are below the list's last key."""
    lists, more, tracked_then, tracked_now, _ = sx
    bh, _, dbound = maintain(lists, more, [0] * len(lists), tracked_then)
    U = max(bh, dbound)
    for d, l in enumerate(lists):
        for k in l:
            if k not in tracked_now:
                assert k <= U
        if more[d] and l:                                            # hidden keys of list d are < l[-1]
            assert l[-1] <= U


@settings(max_examples=1500, deadline=None)
@given(states(), st.integers(0, 3))
def test_repeated_lazy_maintenance_equals_one_look_at_the_latest_time(sx, n_looks):
    lists, more, tracked_then, tracked_now, _ = sx
    cur = [0] * len(lists)
    grow = sorted(tracked_then)
    for i in range(n_looks):                                         # earlier looks at growing subsets
        maintain(lists, more, cur, set(grow[: (i + 1) * len(grow) // (n_looks + 1)]))
    a = maintain(lists, more, cur, tracked_then)
    assert a == maintain(lists, more, [0] * len(lists), tracked_then)
