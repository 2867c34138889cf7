# r2 call 42: K / V ring depth of the D <= 64 forward (4 stages = default, 3, 2): does the smaller shared-memory footprint pay?
mkdir -p gpurun_out
timeout 900 python scripts/variant_sweep.py --variants default,stages3,stages2 --rounds 2 --kernels forward \
  --configs 2048x64xREFx128,2048x64xBF16x128,4096x64xBF16x64,512x64xBF16x512 > gpurun_out/sweep_fwd_stages_d64.jsonl 2> gpurun_out/sweep_fwd_stages_d64.err
python - <<'PY'
import json
for line in open('gpurun_out/sweep_fwd_stages_d64.jsonl'):
    d = json.loads(line)
    print(d.get('variant'), d.get('round'), {k: v.get('forward') for k, v in d.items() if isinstance(v, dict) and 'forward' in v}, d.get('error', ''))
PY
