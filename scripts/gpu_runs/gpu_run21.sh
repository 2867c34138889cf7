mkdir -p gpurun_out
timeout 120 python scripts/trace_forward_d256.py 2048 1 > gpurun_out/trace_d256.log 2>&1
cat gpurun_out/trace_d256.log | cut -c1-330
