# r2 call 13: first run of the layout-generic / wide-head backward kernels (tests + a first timing)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_tcgen05_backward.py -q -m gpu -k "wide or transposed or generic" --timeout 300 -x -q 2>&1 | tail -25 > gpurun_out/generic_tests.txt
cat gpurun_out/generic_tests.txt
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/generic_bench.txt
import sys, json
sys.path.insert(0, '.')
import torch
import mfa_b200 as mfa
from scripts.bench_configs import run
P = mfa.GEMMOperandPrecision
for (N, D, H, tr) in ((4096, 256, 16, (False,)*4), (4096, 128, 32, (True, True, True, True)), (4096, 128, 32, (False, True, False, False)),
                      (4096, 192, 16, (False,)*4), (2048, 64, 64, (True,)*4)):
    try:
        r = run(N, D, P.BF16, H, steps=10, transpose=tr)
        print(N, D, H, tr, {k: (v["tflops"], v["ms"], v["kernel"]) for k, v in r.items() if isinstance(v, dict)}, flush=True)
    except Exception as e:
        print("FAILED", N, D, H, tr, e, flush=True)
PY
