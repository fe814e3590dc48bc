"""A/B of resolver variants on the GPU: env switches + alternative library builds (EGS_LIB)."""
import os, subprocess, sys
N = sys.argv[1] if len(sys.argv) > 1 else "300000"
variants = [
    ("direct lmax async", {}),
    ("direct lmax sync", {"EGS_MW_HPAY": "1"}),
    ("direct nolmax async", {"EGS_MW_LMAX": "0"}),
    ("direct nolmax sync", {"EGS_MW_LMAX": "0", "EGS_MW_HPAY": "1"}),
    ("direct nolmax nohpay", {"EGS_MW_LMAX": "0", "EGS_MW_HPAY": "0"}),
    ("direct lmax async 8 warps", {"EGS_MW_WARPS": "8"}),
    ("direct lmax async 4 warps", {"EGS_MW_WARPS": "4"}),
    ("poll lmax async", {"EGS_LIB": "libegs_poll.so"}),
    ("poll nolmax sync", {"EGS_LIB": "libegs_poll.so", "EGS_MW_LMAX": "0", "EGS_MW_HPAY": "1"}),
]
for name, env in variants:
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "tools/prof_sections.py", N, "4"], env=e, capture_output=True, text=True, timeout=60)
    line = out.stdout.splitlines()[0] if out.stdout else out.stderr[-300:]
    ms = line.split("resolve ms")[1].split()[0] if "resolve ms" in line else "?"
    print(f"{name:32s} resolve ms {ms}", flush=True)
