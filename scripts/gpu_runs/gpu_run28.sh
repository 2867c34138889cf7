mkdir -p gpurun_out
: > gpurun_out/tune_fwd.log
for v in 0 1 0 1; do
  if [ $v = 1 ]; then export MFA_FWD_ONE_TILE=1; else unset MFA_FWD_ONE_TILE; fi
  timeout 200 python scripts/tune_fwd.py >> gpurun_out/tune_fwd.log 2>&1
done
cat gpurun_out/tune_fwd.log | cut -c1-300
MFA_FWD_ONE_TILE=1 timeout 300 python -m pytest tests/test_tcgen05_forward.py -m gpu -q -x -k "128 or 96 or 72 or 80" 2>&1 | tail -3
