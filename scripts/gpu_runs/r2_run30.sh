# r2 call 30: host-buffer path, per-row statistics copied back once per call (default) vs once per chunk, three rounds
mkdir -p gpurun_out
: > gpurun_out/e2e_stats.txt
for round in 0 1 2; do
  for mode in 1 0; do
    export MFA_B200_HOST_BATCH_STATS=$mode
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sustained --no-config5 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('round $round batch_stats $mode e2e ms/step', round(d['e2e']['ms_per_step'],3), 'value', round(d['e2e']['value']), 'numa', d['e2e'].get('numa_node'))" | tee -a gpurun_out/e2e_stats.txt
  done
done
