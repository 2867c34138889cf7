# r2 call 9: persistent dK/dV, table-selected kernel instantiations: every GPU test, A/B against r1, the parameter sweep.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python scripts/variant_sweep.py --variants r1,default --configs 2048x64xREFx128,4096x64xBF16x64,4096x128xBF16x64,1024x64xBF16x256 --rounds 2 > gpurun_out/sweep_bwd3.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_bwd3.jsonl'):
    d=json.loads(l)
    print(d.get('round'), d.get('variant'), {k:v for k,v in d.items() if isinstance(v,dict) and k!='clocks'}, d.get('error','')[:300])
PY
timeout 1200 python scripts/sweep.py --quick > gpurun_out/sweep_tables.log 2>&1; echo "exit $?" >> gpurun_out/sweep_tables.log
cat gpurun_out/sweep_tables.log | cut -c1-400
cp metal-flash-attention_b200/parameters/b200.txt gpurun_out/b200_parameters.txt 2>/dev/null
