# r2 call 25 (final single-GPU evidence run): bench line, reference arm, single-head latencies, parity table, ncu launch
# list of the bench command, full ncu captures of the kernels that changed since call 6.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
tail -n 2 gpurun_out/bench.log | cut -c1-3000
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
tail -n 1 gpurun_out/bench_reference.log | cut -c1-300
timeout 600 python scripts/bench_single.py > gpurun_out/single_head_latency.jsonl 2> gpurun_out/single_head_latency.err
cut -c1-300 gpurun_out/single_head_latency.jsonl
timeout 1500 python scripts/parity_table.py --out gpurun_out/parity.jsonl > gpurun_out/parity.txt 2>&1
cat gpurun_out/parity.txt | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sustained --config5-heads 128 > gpurun_out/ncu_launch.log 2>&1
cat > /tmp/prof_cfg.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from scripts.bench_configs import run
import mfa_b200 as mfa
N, D, prec, H = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
p = {"bf16": mfa.GEMMOperandPrecision.BF16, "fp16": mfa.GEMMOperandPrecision.FP16, "ref": None}[prec]
print(run(N, D, p, H, steps=1))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_backward_generic -s 3 -c 3 -f -o gpurun_out/r2_bwd_generic_d256 python /tmp/prof_cfg.py 4096 256 bf16 16 > gpurun_out/ncu_bwd_generic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_backward -s 3 -c 2 -f -o gpurun_out/r2_bwd_d64_final python /tmp/prof_cfg.py 2048 64 ref 128 > gpurun_out/ncu_bwd_d64.log 2>&1
ls -la gpurun_out/*.ncu-rep | cut -c30-
