# r2 call 6: the bench line with the new legs (sustained, config5, NUMA-local e2e), the reference arm, the ncu launch list of
# the same command and full captures of the forward at D=128 / D=64 and of the backward pair at D=64.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
tail -n 2 gpurun_out/bench.log | cut -c1-4000
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
tail -n 1 gpurun_out/bench_reference.log | cut -c1-300
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_backward -s 3 -c 2 -f -o gpurun_out/r2_bwd_d64 python /tmp/prof_cfg.py 2048 64 ref 128 > gpurun_out/ncu_bwd_d64.log 2>&1
ls -la gpurun_out/*.ncu-rep | cut -c30-
