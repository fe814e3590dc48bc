"""A/B of resolver variants on the GPU: env switches + alternative library builds (EGS_LIB)."""
import os, subprocess, sys
N = sys.argv[1] if len(sys.argv) > 1 else "300000"
variants = [
    ("owner warps (mw)", {"EGS_RESOLVER_TW": "0"}),
    ("ticket warp (tw)", {"EGS_RESOLVER_TW": "1"}),
    ("ticket warp (tw), no hpay", {"EGS_RESOLVER_TW": "1", "EGS_MW_HPAY": "0"}),
    ("ticket warp (tw), 8 helpers", {"EGS_RESOLVER_TW": "1", "EGS_MW_WARPS": "8"}),
]
for name, env in variants:
    e = dict(os.environ); e.update(env)
    try:
        out = subprocess.run([sys.executable, "tools/prof_sections.py", N, "4"], env=e, capture_output=True, text=True, timeout=90)
        line = out.stdout.splitlines()[0] if out.stdout else out.stderr[-300:]
    except subprocess.TimeoutExpired:
        line = "TIMEOUT"
    ms = line.split("resolve ms")[1].split()[0] if "resolve ms" in line else line[-200:]
    print(f"{name:32s} resolve ms {ms}", flush=True)
