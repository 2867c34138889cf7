mkdir -p gpurun_out
(cd tests/gpu_probe && timeout 120 ./_build/umma_probe) > gpurun_out/probe.log 2>&1; echo "probe exit $?" >> gpurun_out/probe.log
timeout 900 python -m pytest tests/test_tcgen05_backward.py -m gpu -q -x > gpurun_out/pytest_bwd.log 2>&1; echo "exit $?" >> gpurun_out/pytest_bwd.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_tcgen05_backward.py > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 9 gpurun_out/probe.log; tail -n 25 gpurun_out/pytest_bwd.log; tail -n 4 gpurun_out/pytest_gpu.log
