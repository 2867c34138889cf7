# r2 call 7: persistent backwardQuery (D <= 64): correctness and A/B against the r1 build.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_backward.py tests/test_golden_gpu.py tests/test_tcgen05_stress.py tests/test_run_host.py -m gpu -q -x > gpurun_out/pytest_bwd.log 2>&1; echo "exit $?" >> gpurun_out/pytest_bwd.log
tail -n 12 gpurun_out/pytest_bwd.log | cut -c1-300
timeout 900 python scripts/variant_sweep.py --variants r1,default --configs 2048x64xREFx128,4096x64xBF16x64,2048x64xFP16x16,1024x64xBF16x256 --kernels backwardQuery,backwardKeyValue --rounds 2 > gpurun_out/sweep_bwd2.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_bwd2.jsonl'):
    d=json.loads(l)
    print(d.get('round'), d.get('variant'), {k:v for k,v in d.items() if isinstance(v,dict) and k!='clocks'}, d.get('error','')[:300])
PY
