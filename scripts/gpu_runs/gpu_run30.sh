mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_backward.py tests/test_tcgen05_stress.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "exit $?" >> gpurun_out/pytest_new.log
tail -n 8 gpurun_out/pytest_new.log
for rep in 1 2; do timeout 200 python scripts/tune_bwd.py 2>&1 | cut -c1-400; done
