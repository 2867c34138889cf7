# r2 call 24: whole GPU suite on the current build, then the per-config timings
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 600 -q 2>&1 | tail -12 > gpurun_out/gpu_tests.txt
cat gpurun_out/gpu_tests.txt
timeout 900 python scripts/bench_configs.py 128 > gpurun_out/bench_configs.jsonl 2> gpurun_out/bench_configs.err
python - <<'PY'
import json
for line in open('gpurun_out/bench_configs.jsonl'):
    if not line.startswith('{'): continue
    d = json.loads(line)
    print(d.get('N'), d.get('D'), d.get('dtype'), d.get('heads'), d.get('transposeState(Q,K,V,O)', ''),
          {k: v['tflops'] for k, v in d.items() if isinstance(v, dict) and 'tflops' in v})
PY
