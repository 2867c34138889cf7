mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
timeout 300 python scripts/bench_configs.py > gpurun_out/bench_configs.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/smoke.log; grep "^{" gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['tflops'], d['roofline']['frac'], d['e2e']['value'], d['single_head'], d['gpu_launches'])"
tail -n 3 gpurun_out/bench_configs.log | cut -c1-330
