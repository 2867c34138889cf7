# r2 call 32: host-buffer path with write-combined UPLOAD buffers only (mfa_host_alloc_upload for Q, K, V; O and L stay
# cacheable) against the previous allocation of everything cacheable (MFA_B200_BENCH_NO_WC=1), three rounds; then the
# run_host tests
mkdir -p gpurun_out
: > gpurun_out/e2e_wc_upload.txt
for round in 0 1 2; do
  for mode in 0 1; do
    export MFA_B200_BENCH_NO_WC=$mode
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sustained --no-config5 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('round $round cacheable_uploads $mode e2e ms/step', round(d['e2e']['ms_per_step'],3), 'value', round(d['e2e']['value']))" | tee -a gpurun_out/e2e_wc_upload.txt
  done
done
timeout 600 python -m pytest tests/test_run_host.py -q -m gpu --timeout 300 -q 2>&1 | tail -4
