mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_configs.py > gpurun_out/bench_configs.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log; head -2 gpurun_out/bench_configs.log | cut -c1-400
