# r2 call 19: second A/B of the persistent dQ kernel's D-term (warp 11: bulk-staged, four rows per lane, no shuffles) vs inline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tcgen05_backward.py tests/test_golden_gpu.py -q -m gpu --timeout 300 -q -x 2>&1 | tail -4 > gpurun_out/bwd_tests.txt
cat gpurun_out/bwd_tests.txt
timeout 900 python scripts/variant_sweep.py --variants default,inlineD --rounds 2 --kernels backwardQuery \
  --configs 1024x64xBF16x256,2048x64xBF16x128,2048x64xREFx128,4096x64xBF16x64,2048x32xBF16x128,512x64xBF16x512 > gpurun_out/sweep_dq_dterm2.jsonl 2> gpurun_out/sweep_dq_dterm2.err
python - <<'PY'
import json
for line in open('gpurun_out/sweep_dq_dterm2.jsonl'):
    d = json.loads(line)
    print(d.get('variant'), d.get('round'), {k: v.get('backwardQuery') for k, v in d.items() if isinstance(v, dict) and 'backwardQuery' in v}, d.get('error', ''))
PY
