# r2 call 40: per-config timings of the final build
mkdir -p gpurun_out
timeout 900 python scripts/bench_configs.py 128 > gpurun_out/bench_configs.jsonl 2> gpurun_out/bench_configs.err
python - <<'PY'
import json
for line in open('gpurun_out/bench_configs.jsonl'):
    if not line.startswith('{'): continue
    d = json.loads(line)
    print(d.get('N'), d.get('D'), d.get('dtype'), d.get('heads'), d.get('transposeState(Q,K,V,O)', ''),
          {k: v['tflops'] for k, v in d.items() if isinstance(v, dict) and 'tflops' in v})
PY
