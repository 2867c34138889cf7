# r2 call 44: last validation of the round: whole GPU suite and smoke() on the final library
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -q 2>&1 | tail -8 > gpurun_out/gpu_tests.txt
cat gpurun_out/gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
