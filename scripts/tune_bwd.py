"""Times forward / dQ / dK/dV at (4096, 128) x 64 heads and (2048, 64) x 128 heads with the library named by
MFA_B200_LIBRARY (tuning builds: `make VARIANT=... EXTRA=-D...` in metal-flash-attention_b200/csrc)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.bench_configs import run  # noqa: E402
import mfa_b200 as mfa  # noqa: E402

P = mfa.GEMMOperandPrecision
res = {"lib": os.path.basename(mfa.library_path())}
for N, D, H in ((4096, 128, 64), (2048, 64, 128)):
    r = run(N, D, P.BF16, H, steps=30)
    res[f"N{N}D{D}"] = {k: v["tflops"] for k, v in r.items() if isinstance(v, dict)}
print(json.dumps(res), flush=True)
