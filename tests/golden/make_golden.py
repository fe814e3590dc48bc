"""Generates tests/golden/scenarios.json from the python mirror of the reference
(oracle/egs_oracle.py).  The reference itself is Go and cannot run in this image, so the mirror
-- pinned to SURVEY 8c's known-answer vectors -- is what produces the committed fixtures.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from scenario import PyBackend, make_scenario, run_scenario  # noqa: E402


def main():
    out = []
    for seed in range(5000, 5016):
        policy = seed % 2
        nodes, ops = make_scenario(seed, n_ops=24, max_c=4 if seed % 4 == 0 else 3)
        trace = run_scenario(PyBackend(policy), nodes, ops, policy)
        out.append({"seed": seed, "policy": policy, "n_ops": 24, "max_c": 4 if seed % 4 == 0 else 3, "trace": trace})
    with open(os.path.join(HERE, "scenarios.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", len(out), "scenarios,", os.path.getsize(os.path.join(HERE, "scenarios.json")), "bytes")


if __name__ == "__main__":
    main()
