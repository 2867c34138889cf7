mkdir -p gpurun_out
MFA_NO_CLUSTER=1 timeout 120 python scripts/trace_forward.py 4096 1 > gpurun_out/trace_single.log 2>&1
cut -c1-420 gpurun_out/trace_single.log | head -30
