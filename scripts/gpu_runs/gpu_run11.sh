mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tcgen05_forward.py -m gpu -q -x > gpurun_out/pytest_fwd.log 2>&1; echo "exit $?" >> gpurun_out/pytest_fwd.log
tail -n 5 gpurun_out/pytest_fwd.log
timeout 300 python scripts/bench_single.py > gpurun_out/bench_single.log 2>&1; echo "exit $?" >> gpurun_out/bench_single.log
cut -c1-420 gpurun_out/bench_single.log
MFA_NO_CLUSTER=1 timeout 120 python scripts/trace_forward.py 4096 1 > gpurun_out/trace_single_scratch.log 2>&1
timeout 120 python scripts/trace_forward.py 4096 1 > gpurun_out/trace_single_cluster.log 2>&1
timeout 120 python scripts/trace_forward.py 4096 64 > gpurun_out/trace_h64.log 2>&1
head -7 gpurun_out/trace_single_scratch.log | cut -c1-420
grep "it=0" gpurun_out/trace_single_scratch.log | cut -c1-400
