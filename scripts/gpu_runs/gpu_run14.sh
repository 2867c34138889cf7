mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_backward.py tests/test_run_host.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "exit $?" >> gpurun_out/pytest_new.log
tail -n 25 gpurun_out/pytest_new.log
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
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
grep "^{" gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['tflops'], d['roofline']['frac'], d['e2e'])"
tail -n 2 gpurun_out/bench.log | cut -c1-300
