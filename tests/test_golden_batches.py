"""Committed golden batch vectors (tests/golden/batches.json, produced by tests/golden/make_golden_batches.py from
the python mirror): replayed through the C oracle on the CPU and through both device engines on the GPU."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "batches.json")) as f:
    CASES = json.load(f)


def _arrays(case):
    c_off = [0]
    units = []
    for req in case["pods"]:
        units.extend(req)
        c_off.append(len(units))
    return np.array(c_off, np.int32), np.array(units, np.int64)


def _check(case, out, rows_of):
    for p, (node, status, masks, fitc, fd, sd) in enumerate(case["out"]):
        assert int(out["node"][p]) == node and int(out["status"][p]) == status, (case["name"], p)
        assert [int(x) for x in out["alloc_mask"][p]] == masks, (case["name"], p)
        assert int(out["fit_count"][p]) == fitc and int(out["fit_digest"][p]) == int(fd) and int(out["score_digest"][p]) == int(sd), (case["name"], p)
    for n, want in enumerate(case["final"]):
        assert [list(r) for r in rows_of(n)] == want, (case["name"], n)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_c_oracle_reproduces_golden_batches(case):
    import oracle_c as oc
    o = oc.OracleC(case["policy"])
    for g, m, rows in case["nodes"]:
        n = o.add_node(100 * g, m * g)
        if rows:
            o.set_rows(n, rows[0], rows[1])
    c_off, units = _arrays(case)
    out = o.schedule_batch(c_off, units)
    _check(case, out, o.rows)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_gpu_reproduces_golden_batches(case, mode):
    import egs_b200
    e = egs_b200.Egs(case["policy"], len(case["nodes"]))
    for n, (g, m, rows) in enumerate(case["nodes"]):
        assert e.node_set(n, g, m) == 0
        if rows:
            assert e.state_load(n, rows[0], rows[1]) == 0
    c_off, units = _arrays(case)
    out = e.schedule_batch(c_off, units.astype(np.int32), mode=mode)
    _check(case, out, e.rows)
