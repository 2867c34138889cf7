# r2 call 16: D-term offload in the persistent dQ kernel (tests + config-3 timing), f16x2 ex2 probe, generic-kernel timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_backward.py -q -m gpu --timeout 300 -q -x 2>&1 | tail -6 > gpurun_out/bwd_tests.txt
cat gpurun_out/bwd_tests.txt
timeout 300 tests/gpu_probe/_build/exp_probe > gpurun_out/exp_probe.txt 2>&1; tail -4 gpurun_out/exp_probe.txt
timeout 900 python scripts/bench_configs.py 128 > gpurun_out/bench_configs.jsonl 2> gpurun_out/bench_configs.err
python - <<'PY'
import json
for line in open('gpurun_out/bench_configs.jsonl'):
    if not line.startswith('{'): continue
    d = json.loads(line)
    print(d.get('N'), d.get('D'), d.get('dtype'), d.get('heads'), d.get('transposeState(Q,K,V,O)', ''),
          {k: v['tflops'] for k, v in d.items() if isinstance(v, dict) and 'tflops' in v})
PY
