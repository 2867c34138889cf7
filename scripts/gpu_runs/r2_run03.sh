# r2 call 3: exp-stream micro-benchmark (MUFU vs FMA-pipe exp2, 1 and 2 warps per sub-partition), r1's softmax probe,
# forward A/B: r1 build, r1 loop + per-warp arrival + LDS/STS epilogue (default), same without per-warp arrival (arr0),
# the in-place restructured loop (inplace).
mkdir -p gpurun_out
timeout 300 tests/gpu_probe/_build/exp_probe > gpurun_out/exp_probe.txt 2>&1; cat gpurun_out/exp_probe.txt
make -s -C tests/gpu_probe _build/softmax_probe > /dev/null 2>&1; timeout 300 tests/gpu_probe/_build/softmax_probe > gpurun_out/softmax_probe.txt 2>&1; cat gpurun_out/softmax_probe.txt
timeout 600 python -m pytest tests/test_tcgen05_forward.py -m gpu -q -x > gpurun_out/pytest_fwd.log 2>&1; echo "exit $?" >> gpurun_out/pytest_fwd.log
tail -n 3 gpurun_out/pytest_fwd.log
timeout 900 python scripts/variant_sweep.py --variants r1,default,arr0,inplace --configs 4096x128xBF16x64,2048x64xFP16x128,4096x64xBF16x64 --kernels forward --rounds 2 > gpurun_out/sweep_fwd2.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_fwd2.jsonl'):
    d=json.loads(l)
    print(d.get('round'), d.get('variant'), {k:v.get('forward') for k,v in d.items() if isinstance(v,dict) and k!='clocks'}, d.get('error','')[:300])
PY
timeout 100 python scripts/trace_forward.py 4096 64 > gpurun_out/trace_fwd_arr.txt 2>&1
tail -n 12 gpurun_out/trace_fwd_arr.txt | cut -c1-250
