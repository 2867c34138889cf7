mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
timeout 300 python bench.py --steps 50 --warmup 5 --heads 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_h1.log 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 --heads 16 --no-e2e --no-cpu-baseline > gpurun_out/bench_h16.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_forward -s 3 -c 1 -o gpurun_out/prof_fwd_r1a python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -n 8 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/bench.log; tail -n 2 gpurun_out/bench_h1.log; tail -n 2 gpurun_out/bench_h16.log; tail -n 5 gpurun_out/launches.csv; tail -n 3 gpurun_out/ncu_full.log
