timeout 200 python -m pytest tests/test_shim_double.py -q --timeout=60 2>&1 | tail -4
bash tools/mgpu.sh 2 r02
