"""egso_cache_load (test support of the oracle): a fresh oracle loaded with the rows and the option caches of a
scheduler that already ran must continue exactly like that scheduler.  The late-window GPU parity tests
(tests/test_gpu_parity.py::test_cfg4_late_windows ...) rely on it."""
import numpy as np
import pytest

import oracle_c as oc


def _mk(w):
    o = oc.OracleC(w.policy)
    for n in range(w.n_nodes):
        o.add_node(100 * w.gpus, w.mem_total * w.gpus)
        o.set_rows(n, w.core[n], w.mem[n])
    return o


@pytest.mark.parametrize("cfg,policy", [(4, 0), (2, 1), (3, 1), (3, 0)])
def test_cache_load_continues_like_the_original(cfg, policy):
    import egs_b200
    w = egs_b200.workloads.config(cfg, n_nodes=50, n_pods=600, policy=policy)
    a = _mk(w)
    pre = w.prefix(300)
    a.schedule_batch(pre.c_off, pre.units64())
    b = oc.OracleC(w.policy)
    for n in range(w.n_nodes):
        b.add_node(100 * w.gpus, w.mem_total * w.gpus)
        r = a.rows(n)
        b.set_rows(n, [x[0] for x in r], [x[1] for x in r])
    for sh in egs_b200.workloads.shapes_of(w):
        valid = np.zeros(w.n_nodes, np.uint8); sc = np.zeros(w.n_nodes, np.int64); am = np.zeros((w.n_nodes, 4), np.uint8)
        for n in range(w.n_nodes):
            pk = a.peek(n, list(sh))
            if pk:
                valid[n] = 1; sc[n] = pk[1]
                for c, lst in enumerate(pk[0]):
                    for g in lst:
                        am[n, c] |= 1 << g
        assert b.cache_load(list(sh), valid, sc, am) == 0
    win = egs_b200.workloads.window(w, 300, 300)
    u = np.arange(10_000, 10_300, dtype=np.uint64)
    ra = a.schedule_batch(win.c_off, win.units64(), uids=u)
    rb = b.schedule_batch(win.c_off, win.units64(), uids=u)
    for k in ra:
        assert np.array_equal(ra[k], rb[k]), k
    assert (ra["status"] != 0).sum() > 0          # the window is in the failure regime
    for n in range(w.n_nodes):
        assert a.rows(n) == b.rows(n)
