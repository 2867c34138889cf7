mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_forward -s 3 -c 1 -o gpurun_out/prof_fwd_r1b python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; grep "^{" gpurun_out/bench.log | cut -c1-200
