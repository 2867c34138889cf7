mkdir -p gpurun_out
cat > /tmp/prof_fwd.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from scripts.bench_configs import run
import mfa_b200 as mfa
print(run(8192, 256, mfa.GEMMOperandPrecision.BF16, 16, steps=1))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_forward_d256 -s 3 -c 1 -f -o gpurun_out/r1_fwd_d256 python /tmp/prof_fwd.py > gpurun_out/ncu_fwd_d256.log 2>&1
tail -n 3 gpurun_out/ncu_fwd_d256.log
ls -la gpurun_out/*.ncu-rep
