"""Differential test: python mirror vs C restatement on seeded random scenarios
(whole-GPU units, -1 sentinel units, stale-cache binds, AddPod / ForgetPod, cold Score)."""
import pytest

from scenario import CBackend, PyBackend, make_scenario, run_scenario


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("seed", range(40))
def test_python_vs_c(seed, policy):
    nodes, ops = make_scenario(seed * 2 + policy)
    a = run_scenario(PyBackend(policy), nodes, ops, policy)
    b = run_scenario(CBackend(policy), nodes, ops, policy)
    assert a == b


@pytest.mark.parametrize("seed", range(10))
def test_c_threads_and_faithful_hash(seed):
    nodes, ops = make_scenario(1000 + seed)
    base = run_scenario(CBackend(0), nodes, ops, 0)
    assert run_scenario(CBackend(0, threads=4), nodes, ops, 0) == base
    assert run_scenario(CBackend(0, faithful=True), nodes, ops, 0) == base
