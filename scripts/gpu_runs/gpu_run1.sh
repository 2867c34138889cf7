mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
(cd tests/gpu_probe && timeout 120 ./_build/umma_probe) > gpurun_out/probe.log 2>&1; echo "probe exit $?" >> gpurun_out/probe.log
timeout 900 python -m pytest tests/test_square_attention.py -m gpu -x -q > gpurun_out/square.log 2>&1; echo "exit $?" >> gpurun_out/square.log
timeout 900 python -m pytest tests/test_tcgen05_forward.py -m gpu -q -x > gpurun_out/tc.log 2>&1; echo "exit $?" >> gpurun_out/tc.log
tail -n 12 gpurun_out/probe.log; tail -n 15 gpurun_out/square.log; tail -n 30 gpurun_out/tc.log
