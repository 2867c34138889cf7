# r2 call 43: three-stage K / V ring in the D <= 64 forward: the tests that reach it, then the whole suite's forward part
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tcgen05_forward.py tests/test_golden_gpu.py tests/test_tcgen05_stress.py tests/test_square_attention.py tests/test_rectangular_attention.py tests/test_run_host.py -q -m gpu --timeout 300 -q 2>&1 | tail -5
