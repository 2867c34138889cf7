# r2 call 35: transposed D-term loop of the generic dQ kernel with batched loads: tests + timings of the transposed configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_backward.py -q -m gpu -k "transposed or generic or wide" --timeout 300 -q 2>&1 | tail -4
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r35_bench.txt
import sys
sys.path.insert(0, '.')
import mfa_b200 as mfa
from scripts.bench_configs import run
P = mfa.GEMMOperandPrecision
for (N, D, H, tr) in ((4096, 128, 32, (True,) * 4), (4096, 128, 32, (False, False, False, True)), (4096, 128, 32, (False, True, False, False)),
                      (2048, 64, 64, (True,) * 4), (4096, 256, 16, (True,) * 4)):
    r = run(N, D, P.BF16, H, steps=20, transpose=tr)
    print(N, D, H, tr, {k: v["tflops"] for k, v in r.items() if isinstance(v, dict)}, flush=True)
PY
