# r2 call 23: dQ D <= 64 with the D-term offload as its own instantiation (long items take the inline form): A/B against the
# inline-only build, then the backward tests
mkdir -p gpurun_out
timeout 900 python scripts/variant_sweep.py --variants default,inlineD --rounds 2 --kernels backwardQuery \
  --configs 1024x64xBF16x256,2048x64xBF16x128,2048x64xREFx128,4096x64xBF16x64,8192x64xBF16x32,4096x128xBF16x64 > gpurun_out/sweep_dq_dterm5.jsonl 2> gpurun_out/sweep_dq_dterm5.err
python - <<'PY'
import json
for line in open('gpurun_out/sweep_dq_dterm5.jsonl'):
    d = json.loads(line)
    print(d.get('variant'), d.get('round'), {k: v.get('backwardQuery') for k, v in d.items() if isinstance(v, dict) and 'backwardQuery' in v}, d.get('error', ''))
PY
timeout 900 python -m pytest tests/test_tcgen05_backward.py tests/test_golden_gpu.py -q -m gpu --timeout 300 -q -x 2>&1 | tail -4 > gpurun_out/bwd_tests.txt
cat gpurun_out/bwd_tests.txt
