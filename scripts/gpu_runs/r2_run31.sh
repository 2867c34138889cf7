# r2 call 31: host-buffer path with write-combined pinned buffers (experiment knob) vs default, three rounds
mkdir -p gpurun_out
: > gpurun_out/e2e_wc.txt
for round in 0 1 2; do
  for mode in 0 1; do
    export MFA_B200_HOST_WC=$mode
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sustained --no-config5 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('round $round write_combined $mode e2e ms/step', round(d['e2e']['ms_per_step'],3), 'value', round(d['e2e']['value']))" | tee -a gpurun_out/e2e_wc.txt
  done
done
