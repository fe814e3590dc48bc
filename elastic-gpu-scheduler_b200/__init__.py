"""elastic-gpu-scheduler_b200: B200-native GPU bin-packing scheduler core.

Drop-in for the Filter / Score / Allocate hot path of elastic-ai/elastic-gpu-scheduler
(pkg/scheduler).  The product is csrc/ -> lib/libegs.so (C ABI in include/egs.h);
this package is the thin Python host used by tests, bench.py and __graft_entry__.py.
"""
from . import _build, capi, workloads  # noqa: F401
from .capi import Egs, EgsError  # noqa: F401
