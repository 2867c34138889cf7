# r2 call 4: forward v2 (one 128-row tile per item, both softmax warpgroups on it, S double/triple-buffered in TMEM,
# persistent, fused split-KV): correctness, A/B against the r1 build, S-buffer / poly variants, pipeline trace,
# single-head latency.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tcgen05_forward.py tests/test_tcgen05_stress.py tests/test_golden_gpu.py tests/test_run_host.py -m gpu -q -x > gpurun_out/pytest_fwd.log 2>&1; echo "exit $?" >> gpurun_out/pytest_fwd.log
tail -n 15 gpurun_out/pytest_fwd.log | cut -c1-300
timeout 900 python scripts/variant_sweep.py --variants r1,default,sb3,p1,sb3p1,sb3p64_0 --configs 4096x128xBF16x64,2048x64xFP16x128,4096x64xBF16x64,4096x128xBF16x1 --kernels forward --rounds 2 > gpurun_out/sweep_fwd3.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_fwd3.jsonl'):
    d=json.loads(l)
    print(d.get('round'), d.get('variant'), {k:v.get('forward') for k,v in d.items() if isinstance(v,dict) and k!='clocks'}, d.get('error','')[:300])
PY
timeout 100 python scripts/trace_forward.py 4096 64 > gpurun_out/trace_fwd_v2.txt 2>&1
cat gpurun_out/trace_fwd_v2.txt | cut -c1-250
MFA_B200_LIBRARY=$PWD/metal-flash-attention_b200/lib/variants/libmfa_b200_sb3.so timeout 100 python scripts/trace_forward.py 4096 64 > gpurun_out/trace_fwd_v2_sb3.txt 2>&1
tail -n 14 gpurun_out/trace_fwd_v2_sb3.txt | cut -c1-250
timeout 200 python scripts/bench_single.py > gpurun_out/bench_single_v2.log 2>&1; cat gpurun_out/bench_single_v2.log | cut -c1-600
