mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tcgen05_forward.py -m gpu -q -x > gpurun_out/pytest_fwd.log 2>&1; echo "exit $?" >> gpurun_out/pytest_fwd.log
tail -n 15 gpurun_out/pytest_fwd.log
timeout 300 python scripts/bench_single.py > gpurun_out/bench_single.log 2>&1; echo "exit $?" >> gpurun_out/bench_single.log
cut -c1-700 gpurun_out/bench_single.log
timeout 120 python scripts/trace_forward.py 4096 1 > gpurun_out/trace_single.log 2>&1
tail -n 8 gpurun_out/trace_single.log | cut -c1-400
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
grep "^{" gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['tflops'], d['roofline']['frac'], d['e2e']['value'], d['single_head'], d['gpu_launches'])"
