# r2 call 15: whole GPU suite (not stopping at the first failure)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 600 -q 2>&1 | tail -30 > gpurun_out/gpu_tests.txt
cat gpurun_out/gpu_tests.txt
