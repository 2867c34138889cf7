mkdir -p gpurun_out
nproc > gpurun_out/nproc.log; lscpu | head -20 >> gpurun_out/nproc.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_forward -s 3 -c 1 -o gpurun_out/prof_fwd_r1 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 5 gpurun_out/smoke.log; tail -n 2 gpurun_out/bench.log | cut -c1-600; tail -n 1 gpurun_out/bench_ref.log | cut -c1-400; tail -n 3 gpurun_out/launches_r1.csv | cut -c1-300
