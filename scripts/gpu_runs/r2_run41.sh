# r2 call 41: ncu evidence of the final build: launch list of the bench command, full captures of the forward at D=128
# (register-store instantiation, N=4096) and D=64 (TMA-store instantiation, N=2048)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sustained --config5-heads 128 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_forward_tcgen05 -s 4 -c 1 -f -o gpurun_out/r2_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-sustained --no-config5 > gpurun_out/ncu_fwd.log 2>&1
cat > /tmp/prof_cfg.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from scripts.bench_configs import run
import mfa_b200 as mfa
N, D, prec, H = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
p = {"bf16": mfa.GEMMOperandPrecision.BF16, "fp16": mfa.GEMMOperandPrecision.FP16, "ref": None}[prec]
print(run(N, D, p, H, steps=1))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_forward_tcgen05 -s 3 -c 1 -f -o gpurun_out/r2_fwd_d64 python /tmp/prof_cfg.py 2048 64 ref 128 > gpurun_out/ncu_fwd_d64.log 2>&1
ls -la gpurun_out/*.ncu-rep | cut -c30-
