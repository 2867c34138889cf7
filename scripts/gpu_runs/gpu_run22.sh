mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tcgen05_forward.py -m gpu -q -x -k "256 or 192 or 136 or large_head" > gpurun_out/pytest_fwd.log 2>&1; echo "exit $?" >> gpurun_out/pytest_fwd.log
tail -n 5 gpurun_out/pytest_fwd.log
timeout 120 python scripts/trace_forward_d256.py 2048 1 > gpurun_out/trace_d256.log 2>&1
tail -n 6 gpurun_out/trace_d256.log | cut -c1-330
python - <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd())
from scripts.bench_configs import run
import mfa_b200 as mfa
for H in (16, 1, 16):
    print(json.dumps(run(8192, 256, mfa.GEMMOperandPrecision.BF16, H, steps=30)))
PY
