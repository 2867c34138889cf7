# r2 call 10: final-candidate build: every GPU test, the measured parity table, configs with clock records, bench line,
# single-head latency, smoke.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python scripts/parity_table.py --out gpurun_out/r2_parity.jsonl > gpurun_out/parity_table.log 2>&1; echo "exit $?" >> gpurun_out/parity_table.log
grep -c max_abs gpurun_out/parity_table.log; tail -n 3 gpurun_out/parity_table.log
timeout 400 python scripts/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "exit $?" >> gpurun_out/bench_configs.log
python - <<'PY'
import json
for l in open("gpurun_out/bench_configs.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["N"], d["D"], d["dtype"][:12], d["heads"], {k: (v["ms"], v["tflops"], v["clocks"]["sm_mhz"]) for k, v in d.items() if isinstance(v, dict)})
    else:
        print(l.strip()[:300])
PY
timeout 200 python scripts/bench_single.py > gpurun_out/bench_single.log 2>&1; cut -c1-700 gpurun_out/bench_single.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
tail -n 2 gpurun_out/bench.log | cut -c1-1500
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
tail -n 2 gpurun_out/smoke.log
