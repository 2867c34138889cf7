mkdir -p gpurun_out
python - <<'PY' > gpurun_out/heads_sweep.log 2>&1
import sys, os, json
sys.path.insert(0, os.getcwd())
from scripts.bench_configs import run
import mfa_b200 as mfa
P = mfa.GEMMOperandPrecision
for H in (1, 2, 4, 8, 16, 32, 64, 128):
    r = run(4096, 128, P.BF16, H, steps=40)
    print(json.dumps({"heads": H, **{k: {"ms": v["ms"], "tflops": v["tflops"]} for k, v in r.items() if isinstance(v, dict)}}), flush=True)
PY
cat gpurun_out/heads_sweep.log | cut -c1-300
