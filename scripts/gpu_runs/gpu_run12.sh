mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_tcgen05_backward.py tests/test_run_host.py -m gpu -q > gpurun_out/pytest_new.log 2>&1; echo "exit $?" >> gpurun_out/pytest_new.log
tail -n 25 gpurun_out/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_tcgen05_backward.py --deselect tests/test_run_host.py > gpurun_out/pytest_rest.log 2>&1; echo "exit $?" >> gpurun_out/pytest_rest.log
tail -n 8 gpurun_out/pytest_rest.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
grep "^{" gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['tflops'], d['roofline']['frac'], d['e2e'], d['single_head'], d['gpu_launches'])"
tail -n 3 gpurun_out/bench.log | cut -c1-300
timeout 300 python scripts/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "exit $?" >> gpurun_out/bench_configs.log
cut -c1-600 gpurun_out/bench_configs.log
