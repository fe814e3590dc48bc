"""Committed golden traces (tests/golden/scenarios.json, made by tests/golden/make_golden.py from the
python mirror) replayed through the C oracle (CPU) and through libegs on the GPU."""
import json
import os

import pytest

from scenario import CBackend, GpuBackend, make_scenario, run_scenario

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "scenarios.json")) as f:
    GOLDEN = json.load(f)


def _norm(x):
    return json.loads(json.dumps(x))


@pytest.mark.parametrize("g", GOLDEN, ids=[str(g["seed"]) for g in GOLDEN])
def test_c_oracle_reproduces_golden(g):
    nodes, ops = make_scenario(g["seed"], n_ops=g["n_ops"], max_c=g["max_c"])
    assert _norm(run_scenario(CBackend(g["policy"]), nodes, ops, g["policy"])) == g["trace"]


@pytest.mark.gpu
@pytest.mark.parametrize("g", GOLDEN, ids=[str(g["seed"]) for g in GOLDEN])
def test_gpu_reproduces_golden(g):
    nodes, ops = make_scenario(g["seed"], n_ops=g["n_ops"], max_c=g["max_c"])
    assert _norm(run_scenario(GpuBackend(g["policy"]), nodes, ops, g["policy"])) == g["trace"]
