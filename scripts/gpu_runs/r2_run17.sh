# r2 call 17: D-term staging through bulk copies in the persistent dQ kernel: tests + config-3 timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_backward.py tests/test_golden_gpu.py -q -m gpu --timeout 300 -q -x 2>&1 | tail -6 > gpurun_out/bwd_tests.txt
cat gpurun_out/bwd_tests.txt
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/dq_bench.txt
import sys
sys.path.insert(0, '.')
import mfa_b200 as mfa
from scripts.bench_configs import run
P = mfa.GEMMOperandPrecision
for rep in range(2):
    for (N, D, prec, H) in ((2048, 64, None, 128), (2048, 64, P.FP16, 128), (2048, 64, P.BF16, 128), (4096, 64, P.BF16, 64), (1024, 64, P.BF16, 256), (2048, 32, P.BF16, 128)):
        r = run(N, D, prec, H, steps=20)
        print(N, D, prec, H, {k: v["tflops"] for k, v in r.items() if isinstance(v, dict)}, flush=True)
PY
