# ncu captures of the backward kernels (N=4096 D=128 bf16, 32 heads; and N=2048 D=64 reference policy)
mkdir -p gpurun_out
cat > /tmp/prof_bwd.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from scripts.bench_configs import run
import mfa_b200 as mfa
N, D, prec, H = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
p = {"bf16": mfa.GEMMOperandPrecision.BF16, "fp16": mfa.GEMMOperandPrecision.FP16, "ref": None}[prec]
print(run(N, D, p, H, steps=1))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_backward -s 3 -c 2 -f -o gpurun_out/r1_bwd2_d128 python /tmp/prof_bwd.py 4096 128 bf16 32 > gpurun_out/ncu_bwd_d128.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_backward -s 3 -c 2 -f -o gpurun_out/r1_bwd2_d64 python /tmp/prof_bwd.py 2048 64 ref 128 > gpurun_out/ncu_bwd_d64.log 2>&1
tail -n 3 gpurun_out/ncu_bwd_d128.log gpurun_out/ncu_bwd_d64.log
ls -la gpurun_out/*.ncu-rep
