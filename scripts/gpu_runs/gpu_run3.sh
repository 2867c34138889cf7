mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
timeout 200 python scripts/trace_forward.py 4096 64 > gpurun_out/trace_full.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log; python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','tflops','ms_per_step','clocks')}, d['roofline']['frac'], d['e2e']['value'] if d['e2e'] else None)
PY
tail -n 18 gpurun_out/trace_full.log
