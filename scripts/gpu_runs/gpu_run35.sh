mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
grep "^{" gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['tflops'], d['roofline']['frac'], d['e2e']['value'], d['single_head']['ms_per_launch'], d['gpu_launches'], d['clocks'])"
tail -n 1 gpurun_out/bench.log
