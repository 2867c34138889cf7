# r2 call 5: production build (two-tile forward + fused split-KV, packed dS arithmetic in the backward kernels, per-device
# host state): every GPU test, backward exp2-poly sweep, single-head latency, the full bench line, smoke.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python scripts/variant_sweep.py --variants r1,default,bwdpoly0,bwdpoly1,bwdpoly3 --configs 2048x64xREFx128,4096x64xBF16x64,4096x128xBF16x64 --rounds 2 > gpurun_out/sweep_bwd.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_bwd.jsonl'):
    d=json.loads(l)
    print(d.get('round'), d.get('variant'), {k:v for k,v in d.items() if isinstance(v,dict) and k!='clocks'}, d.get('error','')[:300])
PY
timeout 200 python scripts/bench_single.py > gpurun_out/bench_single.log 2>&1; cat gpurun_out/bench_single.log | cut -c1-900
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
tail -n 3 gpurun_out/bench.log | cut -c1-3000
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
tail -n 3 gpurun_out/smoke.log
