timeout 300 python -m pytest tests/test_shim_double.py -q --timeout=60 2>&1 | tail -5
bash tools/gpu_round.sh r02
