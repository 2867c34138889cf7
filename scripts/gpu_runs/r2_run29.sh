# r2 call 29: whole GPU suite (extended stress list), smoke(), bench line with the statistics copied back once per call
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 600 -q 2>&1 | tail -12 > gpurun_out/gpu_tests.txt
cat gpurun_out/gpu_tests.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -25 > gpurun_out/smoke.txt
cat gpurun_out/smoke.txt
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
grep "^{" gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sustained',d['sustained']['tflops_per_gpu'])"
