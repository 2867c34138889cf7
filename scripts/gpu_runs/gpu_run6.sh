mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_configs.py > gpurun_out/bench_configs.log 2>&1
tail -n 5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_configs.log | cut -c1-700
