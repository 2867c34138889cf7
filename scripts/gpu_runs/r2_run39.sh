# r2 call 39: TMA-store epilogue of the D <= 64 forward with two scratch tiles per warp vs one (variant onetile)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_forward.py tests/test_golden_gpu.py -q -m gpu --timeout 300 -q -x 2>&1 | tail -3
timeout 900 python scripts/variant_sweep.py --variants default,onetile --rounds 3 --kernels forward \
  --configs 2048x64xREFx128,2048x64xBF16x128,1024x64xBF16x256,512x64xBF16x512,2048x32xBF16x128 > gpurun_out/sweep_fwd_two_tiles.jsonl 2> gpurun_out/sweep_fwd_two_tiles.err
python - <<'PY'
import json
for line in open('gpurun_out/sweep_fwd_two_tiles.jsonl'):
    d = json.loads(line)
    print(d.get('variant'), d.get('round'), {k: v.get('forward') for k, v in d.items() if isinstance(v, dict) and 'forward' in v}, d.get('error', ''))
PY
