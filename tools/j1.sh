timeout 400 python tools/ab_variants.py 300000
EGS_RESOLVER_TW=1 timeout 300 python tools/quick_rounds.py 2>&1 | grep -v "^   ms" | cut -c1-140
timeout 200 python -m pytest "tests/test_sharding.py::test_sharded_engine_in_process_equals_unsharded" -x -q --timeout=120 2>&1 | tail -15
