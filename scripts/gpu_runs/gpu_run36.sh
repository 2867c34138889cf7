mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
grep "^{" gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['tflops'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
tail -n 1 gpurun_out/bench.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['tflops'], d['clocks'])"
