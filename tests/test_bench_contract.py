"""bench.py contract on CPU: the reference arm prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cfg", "1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "decisions/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["vs_baseline"] is None
    assert "workload" in line["config"]


def test_ncu_traffic_summary_is_readable():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    t, src = bench.ncu_traffic()
    assert t is None or (t > 2.0e8 and t < 3.2e8), (t, src)   # ~276 MB per 4M-node launch vs 280 MB algorithmic


def test_reference_arm_under_torchrun_prints_once():
    """The driver launches the reference arm like the GPU arm (torchrun, N ranks): rank 0 alone measures and prints,
    the other ranks exit 0 without work."""
    port = 29900 + os.getpid() % 500
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--cfg", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(x) for x in r.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    assert lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2 and lines[0]["value"] > 0
