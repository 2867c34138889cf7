mkdir -p gpurun_out
python - <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd())
from scripts.bench_configs import run
import mfa_b200 as mfa
P = mfa.GEMMOperandPrecision
for rep in range(3):
    for N, H in ((2048, 128), (4096, 64)):
        r = run(N, 64, P.BF16, H, steps=50)
        print(N, {k: v["tflops"] for k, v in r.items() if isinstance(v, dict)})
PY
