mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_backward.py tests/test_run_host.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "exit $?" >> gpurun_out/pytest_new.log
tail -n 25 gpurun_out/pytest_new.log
: > gpurun_out/tune_bwd.log
for lib in "" metal-flash-attention_b200/lib/variants/libmfa_b200_bwdpoly0.so metal-flash-attention_b200/lib/variants/libmfa_b200_bwdpoly1.so metal-flash-attention_b200/lib/variants/libmfa_b200_bwdpoly3.so; do
  if [ -n "$lib" ]; then export MFA_B200_LIBRARY=$PWD/$lib; else unset MFA_B200_LIBRARY; fi
  timeout 200 python scripts/tune_bwd.py >> gpurun_out/tune_bwd.log 2>&1
done
unset MFA_B200_LIBRARY
cat gpurun_out/tune_bwd.log | cut -c1-400
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
grep "^{" gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['tflops'], d['roofline']['frac'], d['e2e'])"
tail -n 2 gpurun_out/bench.log | cut -c1-300
