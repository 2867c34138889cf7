mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_golden_gpu.py -m gpu -q > gpurun_out/pytest_golden.log 2>&1; echo "exit $?" >> gpurun_out/pytest_golden.log
tail -n 6 gpurun_out/pytest_golden.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
tail -n 12 gpurun_out/smoke.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_n2.log
grep "^{" gpurun_out/bench_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['tflops'], d['ms_per_step'], d['e2e'], d['clocks'])"
tail -n 3 gpurun_out/bench_n2.log | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_ref_n2.log
tail -n 2 gpurun_out/bench_ref_n2.log | cut -c1-200
