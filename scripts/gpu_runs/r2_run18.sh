# r2 call 18: A/B of the persistent dQ kernel's D-term (warp-11 offload vs inline), same box, two rounds; generic kernels
# re-tested after the host-side dO conversion and the DPAD = 64 instantiation
mkdir -p gpurun_out
timeout 900 python scripts/variant_sweep.py --variants default,inlineD --rounds 2 --kernels backwardQuery \
  --configs 1024x64xBF16x256,2048x64xBF16x128,2048x64xREFx128,4096x64xBF16x64,2048x32xBF16x128 > gpurun_out/sweep_dq_dterm.jsonl 2> gpurun_out/sweep_dq_dterm.err
python - <<'PY'
import json
for line in open('gpurun_out/sweep_dq_dterm.jsonl'):
    d = json.loads(line)
    print(d.get('variant'), d.get('round'), {k: v.get('backwardQuery') for k, v in d.items() if isinstance(v, dict) and 'backwardQuery' in v}, d.get('error', ''))
PY
timeout 900 python -m pytest tests/test_tcgen05_backward.py -q -m gpu -k "wide or transposed or generic" --timeout 300 -q 2>&1 | tail -6 > gpurun_out/generic_tests.txt
cat gpurun_out/generic_tests.txt
