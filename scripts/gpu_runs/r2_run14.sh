# r2 call 14: layout-generic / wide-head backward kernels: full test selection, then the whole GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_tcgen05_backward.py -q -m gpu -k "wide or transposed or generic" --timeout 300 -q 2>&1 | tail -25 > gpurun_out/generic_tests.txt
cat gpurun_out/generic_tests.txt
timeout 2400 python -m pytest tests -q -m gpu --timeout 600 -q -x 2>&1 | tail -15 > gpurun_out/gpu_tests.txt
cat gpurun_out/gpu_tests.txt
