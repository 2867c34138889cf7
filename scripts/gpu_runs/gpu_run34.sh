mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_forward.py tests/test_rectangular_attention.py -m gpu -q > gpurun_out/pytest_fwd.log 2>&1; echo "exit $?" >> gpurun_out/pytest_fwd.log
tail -n 30 gpurun_out/pytest_fwd.log | cut -c1-250
