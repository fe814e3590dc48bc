"""The rounds ALGORITHM (tests/rounds_model.py, a Python model of csrc/egs_rounds.cuh) is exact: on random
clusters and pod streams -- mixed shapes, sentinel and whole-GPU units, several batches on one state, tiny
list depth / tracked table / shape set so that every early-termination path fires, 1..3 shards -- it
reproduces the reference driver rule (oracle) output for output, digests included."""
import numpy as np
import pytest

import egs_oracle as po
from rounds_model import RoundsModel


def _cluster(rng, n_nodes):
    nodes = []
    for _ in range(n_nodes):
        g = int(rng.choice([1, 2, 4, 8]))
        m = int(rng.choice([16, 40, 80]))
        rows = None
        if rng.integers(0, 2):
            rows = ([int(rng.choice([100, 100, 60, 30, 0])) for _ in range(g)], [int(rng.integers(0, m + 1)) for _ in range(g)])
        nodes.append((100 * g, m * g, rows))
    return nodes


def _shapes(rng, n, mono):
    out = []
    for _ in range(n):
        units = []
        for _ in range(int(rng.integers(1, 4))):
            k = rng.integers(0, 10)
            if k == 0 and not mono:
                units.append((-1, -1, 0))
            elif k == 1:
                units.append((0, 0, int(rng.integers(1, 3))))
            else:
                units.append((int(rng.choice([0, 5, 10, 25, 50])), int(rng.integers(1, 12)), 0))
        out.append(tuple(units))
    return out


@pytest.mark.parametrize("seed", range(24))
def test_model_equals_oracle(seed):
    rng = np.random.default_rng(seed)
    policy = seed % 2
    mono = seed % 3 == 0
    K, T, RS, D = int(rng.choice([1, 2, 3, 8])), int(rng.choice([2, 3, 5, 64])), int(rng.choice([1, 2, 4, 32])), int(rng.choice([1, 2, 3]))
    nodes = _cluster(rng, int(rng.integers(3, 40)))
    shapes = _shapes(rng, int(rng.integers(1, 7)), mono)
    o = po.Scheduler(policy)
    m = RoundsModel(policy, K=K, T=T, RS=RS, shards=D)
    for core, mem, rows in nodes:
        a, b = o.add_node(core, mem), m.add_node(core, mem)
        assert a == b
        if rows:
            o.set_rows(a, *rows); m.set_rows(a, *rows)
    uid = 0
    for batch in range(3):                                        # state (rows, option caches) carries over
        pods = [shapes[int(i)] for i in rng.integers(0, len(shapes), int(rng.integers(20, 160)))]
        got = m.schedule_batch(pods)
        for p, (s, g) in enumerate(zip(pods, got)):
            r = o.schedule_one(list(s), uid)
            uid += 1
            want = dict(node=r["node"], status=r["status"], alloc=r["alloc"], fit_count=r["fit_count"],
                        fit_digest=r["fit_digest"], score_digest=r["score_digest"])
            assert g == want, (seed, batch, p, K, T, RS, D)
        for n in range(len(nodes)):
            assert m.rows(n) == o.rows(n)
        # option caches agree too: cached (CACHED) <=> present in the oracle's map, with the same option
        for s in set(pods):
            for n in range(len(nodes)):
                e = m.tables[s][n]
                opt = o.nodes[n].allocated.get(tuple(s)) if o.nodes[n] is not None else None
                assert (e.st == 1) == (opt is not None), (seed, batch, s, n, e.st)
                if opt is not None:
                    assert (e.score, e.alloc) == (opt.score, opt.allocated)
    assert m.stats["rounds"] >= 3


def test_early_termination_paths_are_exercised():
    tot = dict(rounds=0, dry=0, dry_harmless=0, full=0, shape=0, fast=0)
    for seed in range(24):
        rng = np.random.default_rng(seed)
        K, T, RS, D = int(rng.choice([1, 2, 3, 8])), int(rng.choice([2, 3, 5, 64])), int(rng.choice([1, 2, 4, 32])), int(rng.choice([1, 2, 3]))
        nodes = _cluster(rng, int(rng.integers(3, 40)))
        shapes = _shapes(rng, int(rng.integers(1, 7)), seed % 3 == 0)
        m = RoundsModel(seed % 2, K=K, T=T, RS=RS, shards=D)
        for core, mem, rows in nodes:
            a = m.add_node(core, mem)
            if rows:
                m.set_rows(a, *rows)
        m.schedule_batch([shapes[int(i)] for i in rng.integers(0, len(shapes), 150)])
        for k in tot:
            tot[k] += m.stats[k]
    assert tot["dry"] > 0 and tot["dry_harmless"] > 0 and tot["full"] > 0 and tot["shape"] > 0 and tot["fast"] > 0, tot


def _fast_regime(seed):
    rng = np.random.default_rng(5000 + seed)
    K, T, D = int(rng.choice([1, 2, 4])), int(rng.choice([2, 3, 4, 6])), int(rng.choice([1, 2, 3]))
    nodes = _cluster(rng, int(rng.integers(6, 30)))
    shapes = [((int(rng.choice([5, 10, 25, 50])), int(rng.integers(1, 12)), 0),) for _ in range(int(rng.integers(1, 5)))]
    pods = [shapes[int(i)] for i in rng.integers(0, len(shapes), 200)]
    return K, T, D, nodes, pods


@pytest.mark.parametrize("seed", range(16))
def test_fast_pods_full_table_and_dry_lists_stay_exact(seed):
    """Monotone rounds of single-container shapes (the benchmark regime): fast pods go on resolving on a FULL tracked
    table as long as tracked options win, and an exhausted truncated list ends the round only when what it hides could
    beat the winner -- output for output the oracle's."""
    policy = seed % 2
    K, T, D, nodes, pods = _fast_regime(seed)
    o = po.Scheduler(policy)
    m = RoundsModel(policy, K=K, T=T, RS=8, shards=D)
    for core, mem, rows in nodes:
        a = o.add_node(core, mem); m.add_node(core, mem)
        if rows:
            o.set_rows(a, *rows); m.set_rows(a, *rows)
    got = m.schedule_batch(pods)
    for p, (s, g) in enumerate(zip(pods, got)):
        r = o.schedule_one(list(s), p)
        want = dict(node=r["node"], status=r["status"], alloc=r["alloc"], fit_count=r["fit_count"],
                    fit_digest=r["fit_digest"], score_digest=r["score_digest"])
        assert g == want, (seed, p, K, T, D)
    for n in range(len(nodes)):
        assert m.rows(n) == o.rows(n)
    assert m.stats["fast"] > 0


def test_fast_regime_paths_are_exercised():
    tot = dict(dry=0, dry_harmless=0, full=0, fast=0)
    for seed in range(16):
        K, T, D, nodes, pods = _fast_regime(seed)
        m = RoundsModel(seed % 2, K=K, T=T, RS=8, shards=D)
        for core, mem, rows in nodes:
            a = m.add_node(core, mem)
            if rows:
                m.set_rows(a, *rows)
        m.schedule_batch(pods)
        for k in tot:
            tot[k] += m.stats[k]
    assert all(v > 0 for v in tot.values()), tot
