# r2 call 37: forward epilogue, TMA stores for short items only (run-time choice) vs the register-store build
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tcgen05_forward.py tests/test_golden_gpu.py tests/test_tcgen05_stress.py tests/test_run_host.py -q -m gpu --timeout 300 -q -x 2>&1 | tail -4
timeout 900 python scripts/variant_sweep.py --variants default,stg --rounds 2 --kernels forward \
  --configs 4096x128xBF16x64,2048x64xREFx128,2048x64xBF16x128,1024x128xBF16x128,4096x64xBF16x64,512x64xBF16x512,4096x128xBF16x1 > gpurun_out/sweep_fwd_tma_store2.jsonl 2> gpurun_out/sweep_fwd_tma_store2.err
python - <<'PY'
import json
for line in open('gpurun_out/sweep_fwd_tma_store2.jsonl'):
    d = json.loads(line)
    print(d.get('variant'), d.get('round'), {k: v.get('forward') for k, v in d.items() if isinstance(v, dict) and 'forward' in v}, d.get('error', ''))
PY
