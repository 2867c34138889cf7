mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_backward.py tests/test_run_host.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "exit $?" >> gpurun_out/pytest_new.log
tail -n 25 gpurun_out/pytest_new.log
timeout 300 python scripts/bench_single.py > gpurun_out/bench_single.log 2>&1; echo "exit $?" >> gpurun_out/bench_single.log
python - <<'PY'
import json
for l in open("gpurun_out/bench_single.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["N"], d["D"], d["dtype"], {k: (v["launches"], v["eager_us"], v["graph_us"], v["tflops_graph"]) for k, v in d.items() if isinstance(v, dict)})
    else:
        print(l.strip()[:300])
PY
