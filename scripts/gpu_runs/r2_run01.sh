# r2 call 1: baseline of the round-1 kernels -- measured parity table, configs with clock records, all GPU tests
# (their measured errors land in gpurun_out/parity_tests.jsonl), forward pipeline trace, ncu of the D=64 forward.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 900 python scripts/parity_table.py --out gpurun_out/r2_parity_baseline.jsonl > gpurun_out/parity_table.log 2>&1; echo "exit $?" >> gpurun_out/parity_table.log
tail -n 60 gpurun_out/parity_table.log
timeout 400 python scripts/bench_configs.py > gpurun_out/bench_configs_baseline.log 2>&1; echo "exit $?" >> gpurun_out/bench_configs_baseline.log
timeout 100 python scripts/trace_forward.py 4096 64 > gpurun_out/trace_fwd_baseline.txt 2>&1
tail -n 8 gpurun_out/trace_fwd_baseline.txt
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
cat > /tmp/prof_fwd64.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from scripts.bench_configs import run
import mfa_b200 as mfa
print(run(2048, 64, mfa.GEMMOperandPrecision.FP16, 128, steps=1))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_forward_tcgen05 -s 3 -c 1 -f -o gpurun_out/r2_fwd_d64_baseline python /tmp/prof_fwd64.py > gpurun_out/ncu_fwd_d64.log 2>&1
ls -la gpurun_out/*.ncu-rep | cut -c30-
