# r2 call 20: fused dK/dV pass of the layout-generic kernels (D <= 128), runtime D-term switch + L prefetch in the
# persistent dQ kernel: backward tests, then timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_backward.py tests/test_golden_gpu.py tests/test_rectangular_attention.py -q -m gpu --timeout 300 -q -x 2>&1 | tail -6 > gpurun_out/bwd_tests.txt
cat gpurun_out/bwd_tests.txt
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/r20_bench.txt
import sys
sys.path.insert(0, '.')
import mfa_b200 as mfa
from scripts.bench_configs import run
P = mfa.GEMMOperandPrecision
for (N, D, prec, H, tr) in ((512, 64, P.BF16, 512, None), (1024, 64, P.BF16, 256, None), (2048, 64, P.BF16, 128, None), (2048, 64, None, 128, None),
                            (4096, 64, P.BF16, 64, None), (4096, 128, P.BF16, 32, (True,) * 4), (4096, 128, P.BF16, 32, (False, True, False, False)),
                            (2048, 64, P.BF16, 64, (True,) * 4), (4096, 256, None, 16, None), (4096, 256, P.BF16, 16, (True,) * 4)):
    r = run(N, D, prec, H, steps=20, transpose=tr or (False,) * 4)
    print(N, D, prec, H, tr, {k: v["tflops"] for k, v in r.items() if isinstance(v, dict)}, flush=True)
PY
