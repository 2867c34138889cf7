mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_stress.py -m gpu -q > gpurun_out/pytest_stress.log 2>&1; echo "exit $?" >> gpurun_out/pytest_stress.log
tail -n 40 gpurun_out/pytest_stress.log | cut -c1-300
