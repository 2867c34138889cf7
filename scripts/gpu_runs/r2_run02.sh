# r2 call 2: forward softmax restructure (in-place pipelined exp, per-warp arrivals) -- correctness, A/B against the r1
# build and the poly / sum variants, pipeline trace.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tcgen05_forward.py tests/test_tcgen05_stress.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/pytest_fwd.log 2>&1; echo "exit $?" >> gpurun_out/pytest_fwd.log
tail -n 5 gpurun_out/pytest_fwd.log
timeout 900 python scripts/variant_sweep.py --variants r1,default,sumraw,poly0,poly1,poly2 --configs 4096x128xBF16x64,2048x64xFP16x128,4096x64xBF16x64,8192x256xBF16x16 --kernels forward --rounds 2 > gpurun_out/sweep_fwd.jsonl 2>&1
cat gpurun_out/sweep_fwd.jsonl | cut -c1-400
timeout 100 python scripts/trace_forward.py 4096 64 > gpurun_out/trace_fwd_new.txt 2>&1
tail -n 4 gpurun_out/trace_fwd_new.txt
