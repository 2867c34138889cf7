# r2 call 21: dQ D <= 64: current build (staging area allocated, runtime choice) vs inline-only build (no staging area)
mkdir -p gpurun_out
timeout 900 python scripts/variant_sweep.py --variants default,inlineD --rounds 2 --kernels backwardQuery \
  --configs 1024x64xBF16x256,2048x64xBF16x128,2048x64xREFx128,4096x64xBF16x64,8192x64xBF16x32 > gpurun_out/sweep_dq_dterm3.jsonl 2> gpurun_out/sweep_dq_dterm3.err
python - <<'PY'
import json
for line in open('gpurun_out/sweep_dq_dterm3.jsonl'):
    d = json.loads(line)
    print(d.get('variant'), d.get('round'), {k: v.get('backwardQuery') for k, v in d.items() if isinstance(v, dict) and 'backwardQuery' in v}, d.get('error', ''))
PY
