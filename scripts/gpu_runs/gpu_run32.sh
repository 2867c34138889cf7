mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_backward.py tests/test_tcgen05_stress.py tests/test_golden_gpu.py tests/test_run_host.py -m gpu -q > gpurun_out/pytest_new.log 2>&1; echo "exit $?" >> gpurun_out/pytest_new.log
tail -n 4 gpurun_out/pytest_new.log
timeout 300 python scripts/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "exit $?" >> gpurun_out/bench_configs.log
python - <<'PY'
import json
for l in open("gpurun_out/bench_configs.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["N"], d["D"], d["dtype"][:12], d["heads"], {k: (v["ms"], v["tflops"]) for k, v in d.items() if isinstance(v, dict)})
    else:
        print(l.strip()[:300])
PY
