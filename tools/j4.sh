timeout 300 python tools/quick_rounds.py 2>&1 | grep -v "^   ms" | cut -c1-150
timeout 300 python -m pytest tests/test_sharding.py tests/test_gpu_abi_edges.py -q -m gpu --timeout=120 2>&1 | tail -4
timeout 100 python tools/prof_sections.py 300000 4 | head -1
EGS_LIB=libegs_prof.so timeout 100 python tools/prof_sections.py 300000 4 | tail -7
