# r2 call 26: chunk-count sweep of the host-buffer path (e2e leg of bench.py), two rounds
mkdir -p gpurun_out
: > gpurun_out/e2e_chunks.txt
for round in 0 1; do
  for n in default 8 12 16 24 32; do
    if [ "$n" = default ]; then unset MFA_B200_HOST_CHUNKS; else export MFA_B200_HOST_CHUNKS=$n; fi
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sustained --no-config5 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('round $round chunks $n e2e ms/step', round(d['e2e']['ms_per_step'],3), 'value', round(d['e2e']['value']), 'kernel ms', round(d['ms_per_step'],4))" | tee -a gpurun_out/e2e_chunks.txt
  done
done
