mkdir -p gpurun_out
: > gpurun_out/tune_fwd.log
for rep in 1 2; do
for lib in "" $(ls metal-flash-attention_b200/lib/variants/*.so 2>/dev/null); do
  if [ -n "$lib" ]; then export MFA_B200_LIBRARY=$PWD/$lib; else unset MFA_B200_LIBRARY; fi
  timeout 200 python scripts/tune_fwd.py >> gpurun_out/tune_fwd.log 2>&1
done
done
unset MFA_B200_LIBRARY
cat gpurun_out/tune_fwd.log | cut -c1-300
