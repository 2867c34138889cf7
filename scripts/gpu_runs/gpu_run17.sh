mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_backward.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "exit $?" >> gpurun_out/pytest_new.log
tail -n 25 gpurun_out/pytest_new.log
: > gpurun_out/tune_bwd.log
for rep in 1 2; do
for lib in "" $(ls metal-flash-attention_b200/lib/variants/*.so 2>/dev/null); do
  if [ -n "$lib" ]; then export MFA_B200_LIBRARY=$PWD/$lib; else unset MFA_B200_LIBRARY; fi
  timeout 200 python scripts/tune_bwd.py >> gpurun_out/tune_bwd.log 2>&1
done
done
unset MFA_B200_LIBRARY
cat gpurun_out/tune_bwd.log | cut -c1-400
